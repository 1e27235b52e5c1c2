O=gpurun_out/r2s2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --scaling weak > $O/bench_n8_weak.json 2> $O/bench_n8_weak.err
python - <<'PY'
import json
for n in (1,2,4,8,'8_weak'):
    try:
        d=json.load(open(f'gpurun_out/r2s2/bench_n{n}.json'))
        print(n, d['scaling'], 'value %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'serial %.3e'%d['e2e']['value_serial_one_handle'], 'f64 e2e %.3e'%d['e2e']['float64_io']['value'], 'kernel', d['roofline']['kernel'], [ (k, '%.3e'%v['value']) for k,v in d.items() if k.endswith('_scaling')])
    except Exception as e:
        print(n, 'failed', e)
PY
