for cfg in 0 1 2; do for wpc in 4; do
  echo "N=32 CFG=$cfg wpc=$wpc"; PMS_WARP_CFG=$cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --wpc $wpc --peds 32 --worlds 8192 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launch']['blocks_per_sm'], d['launch']['n_chunks'], d['launch']['worlds_per_block'])"
done; done
