# bench.py back to back on the GPUs of one box:  NS="2 4 8" WEAK=8 bash tools/r2_scale_gpu.sh   (outputs: gpurun_out/r2s2/)
O=gpurun_out/r2s2; mkdir -p $O
NS=${NS:-"2 4 8"}
nvidia-smi topo -m > $O/topo.txt 2>&1
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
for n in $NS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err
done
if [ -n "$WEAK" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $WEAK --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $WEAK --steps 10 --warmup 3 --no-cpu-baseline --scaling weak > $O/bench_n${WEAK}_weak.json 2> $O/bench_n${WEAK}_weak.err
fi
python - <<'PY'
import json, os
for n in (1, 2, 4, 8, '8_weak'):
    p = f'gpurun_out/r2s2/bench_n{n}.json'
    if not os.path.exists(p):
        continue
    try:
        d = json.load(open(p))
        print(n, d['scaling'], 'value %.3e' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'serial %.3e' % d['e2e']['value_serial_one_handle'],
              'kernel', d['roofline']['kernel'], 'issue', d['roofline'].get('issue_slots_pct'), [(k, '%.3e' % v['value']) for k, v in d.items() if k.endswith('_scaling')])
    except Exception as e:
        print(n, 'failed', e)
PY
