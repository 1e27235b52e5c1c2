O=gpurun_out/r2l; mkdir -p $O
for W in 2048 1536 1024 3072; do for wpc in 4 3 2 1; do
  python tools/run_case.py $W 64 1000 5 $wpc 0 fp32 3 2>&1 | tail -1 | sed "s/^/wpc $wpc: /" >> $O/wpc.txt
done; done
for W in 4096 8192; do for wpc in 4 2; do
  python tools/run_case.py $W 32 1000 5 $wpc 0 fp32 3 2>&1 | tail -1 | sed "s/^/wpc $wpc: /" >> $O/wpc.txt
done; done
cat $O/wpc.txt
python -m pytest tests/test_sensor_schedule.py tests/test_gpu_replay.py -m gpu -x -q 2>&1 | tail -2
