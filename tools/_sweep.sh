for sk in 0 1 2; do for cfg in 0 1 2 3; do
  echo "C4 SKIP=$sk CFG=$cfg"; PMS_DEBUG_SKIP=$sk PMS_WARP_CFG=$cfg python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['launch']['blocks_per_sm'])"
done; done
