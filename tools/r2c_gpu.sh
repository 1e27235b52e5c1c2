O=gpurun_out/r2c; mkdir -p $O
for lib in libpms_b200 libpms_noped libpms_norobot libpms_none; do
  echo "== $lib" >> $O/ab.txt
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 64 1000 0 0 0 fp32 4 >> $O/ab.txt 2>&1
done
python -m pytest tests/test_sensor_schedule.py -m gpu -q > $O/t_sched.log 2>&1; echo "rc=$?" >> $O/t_sched.log
cat $O/ab.txt; tail -40 $O/t_sched.log
