O=gpurun_out/r2j; mkdir -p $O
for lib in libpms_scalar libpms_b200; do
  echo "== $lib" >> $O/ab.txt
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 64 1000 0 0 0 fp32 4 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 128 300 0 0 0 fp32 3 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 96 300 0 0 0 fp32 3 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 64 1000 0 0 0 fp32_plain 3 >> $O/ab.txt 2>&1
done
python tools/bench_configs.py c5 > $O/c5.jsonl 2> $O/c5.err
python -m pytest tests/test_gpu_orca.py tests/test_gpu_warp.py tests/test_gpu_trajectory.py tests/test_gpu_hsfm.py tests/test_gpu_noise.py tests/test_sensor_schedule.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
cat $O/ab.txt; tail -6 $O/t.log; cut -c1-250 $O/c5.jsonl
