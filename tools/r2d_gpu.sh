O=gpurun_out/r2d; mkdir -p $O
python -m pytest tests/test_sensor_schedule.py -m gpu -q > $O/t_sched.log 2>&1; echo "rc=$?" >> $O/t_sched.log
python tools/bench_configs.py c2 c3 c4 > $O/configs.jsonl 2> $O/configs.err
python tools/run_case.py 512 64 1000 0 0 0 fp32 3 >> $O/cases.txt 2>&1
tail -5 $O/t_sched.log; cat $O/cases.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2d/configs.jsonl'):
    try: d=json.loads(l)
    except: continue
    print(d.get('case'), '|', d.get('precision'), '|', round(d.get('ms',0),3), 'ms | kernel', d.get('kernel'))
PY
