O=gpurun_out/r2f; mkdir -p $O
python -m pytest tests/test_gpu_device_io.py -m gpu -x -q > $O/t_io.log 2>&1; echo "rc=$?" >> $O/t_io.log
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -15 $O/t_io.log; tail -5 $O/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f/bench_n1.json'))
print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'], indent=1)); print(d['roofline']['traffic'], d['roofline']['fma_pipe_pct'])
PY
