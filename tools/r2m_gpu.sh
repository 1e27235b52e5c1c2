O=gpurun_out/r2m; mkdir -p $O
for lib in libpms_onephase libpms_blocked libpms_b200; do
  echo "== $lib" >> $O/ab.txt
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 64 1000 0 0 0 fp32 4 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 64 1000 0 0 0 fp32_plain 3 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 2048 64 1000 5 0 0 fp32 3 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/run_case.py 4096 48 1000 0 0 0 fp32 3 >> $O/ab.txt 2>&1
  PMS_LIB=pyminisim_b200/$lib.so python tools/dump_state.py $O/$lib.npz
done
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/r2m/libpms_onephase.npz')
for other in ('libpms_blocked', 'libpms_b200'):
    b=np.load(f'gpurun_out/r2m/{other}.npz')
    for k in a.files:
        eq=np.array_equal(a[k], b[k], equal_nan=True)
        if not eq:
            d=np.abs(a[k].astype(float)-b[k].astype(float)); print(other, k, 'DIFF max', np.nanmax(d), 'n', int((d>0).sum()))
print('compared', len(a.files), 'arrays; collisions', int(a['mirror64_col'].sum()), int(a['bench64_col'].sum()), 'nonfinite worlds', int((a['mirror64_nf']>0).sum()))
PY
python -m pytest tests/test_gpu_warp.py tests/test_gpu_trajectory.py tests/test_gpu_hsfm.py tests/test_gpu_noise.py tests/test_sensor_schedule.py tests/test_gpu_dropin.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
cat $O/ab.txt; tail -6 $O/t.log
