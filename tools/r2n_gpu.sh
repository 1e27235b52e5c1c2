O=gpurun_out/r2n; mkdir -p $O
for lib in libpms_t1two libpms_b200; do
  echo "== $lib" >> $O/ab.txt
  for W in 512 1024 1332 1536 2048 2664 4096; do
    PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py $W 32 1000 4 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt
  done
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 1024 20 1000 4 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt
done
echo "== warp kernel" >> $O/ab.txt
for W in 1332 1536 2048 2664 4096; do python tools/run_case.py $W 32 1000 5 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt; done
python -m pytest tests/test_gpu_warp.py tests/test_gpu_trajectory.py tests/test_sensor_schedule.py tests/test_gpu_noise.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
cat $O/ab.txt; tail -5 $O/t.log
