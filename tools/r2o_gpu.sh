O=gpurun_out/r2o; mkdir -p $O
for lib in libpms_oneunit libpms_b200; do
  echo "== $lib" >> $O/ab.txt
  for W in 128 256 512 740 1024 1480 2048; do
    PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py $W 64 1000 4 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt
  done
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 512 48 1000 4 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 1024 32 1000 4 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt
  PMS_LIB=pyminisim_b200/$lib.so python tools/dump_state.py $O/$lib.npz 4
done
echo "== warp kernel" >> $O/ab.txt
for W in 1024 1480 2048; do python tools/run_case.py $W 64 1000 5 0 0 fp32 3 2>&1 | tail -1 >> $O/ab.txt; done
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/r2o/libpms_oneunit.npz'); b=np.load('gpurun_out/r2o/libpms_b200.npz')
for k in a.files:
    if not np.array_equal(a[k], b[k], equal_nan=True):
        d=np.abs(a[k].astype(float)-b[k].astype(float)); print(k, 'DIFF max', np.nanmax(d), 'n', int((d>0).sum()))
print('compared', len(a.files))
PY
python -m pytest tests/test_gpu_warp.py tests/test_gpu_trajectory.py tests/test_sensor_schedule.py tests/test_gpu_noise.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
cat $O/ab.txt; tail -5 $O/t.log
