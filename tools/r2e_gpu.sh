O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/t_gpu.log 2>&1; echo "rc=$?" >> $O/t_gpu.log
python tools/bench_configs.py c2 c4 > $O/configs.jsonl 2> $O/configs.err
tail -5 $O/t_gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/r2e/configs.jsonl'):
    try: d=json.loads(l)
    except: continue
    print(d.get('case'), '|', d.get('precision'), '|', round(d.get('ms',0),3), 'ms | kernel', d.get('kernel'))
PY
