O=gpurun_out/r2h; mkdir -p $O
python -m pytest tests/test_gpu_replay.py tests/test_gpu_dropin.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
tail -30 $O/t.log
