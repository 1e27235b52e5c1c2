O=gpurun_out/r2p; mkdir -p $O
for W in 128 256 400 512 740 1024 1480; do python tools/run_case.py $W 64 1000 0 0 0 fp32 3 2>&1 | tail -1 >> $O/auto.txt; done
for W in 1024 1332 2048 2200 2664; do python tools/run_case.py $W 32 1000 0 0 0 fp32 3 2>&1 | tail -1 >> $O/auto.txt; done
python -m pytest tests -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
cat $O/auto.txt; tail -5 $O/t.log
