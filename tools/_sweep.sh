python -m pytest tests/test_gpu_warp.py -q 2>&1 | tail -3
for pr in 0 1; do for cfg in 0 1; do
  echo "C4 PRUNE=$pr CFG=$cfg"; PMS_WARP_PRUNE=$pr PMS_WARP_CFG=$cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launch']['blocks_per_sm'], d['episode_stats']['detections'], d['episode_stats']['robot_ped_collisions'], d['episode_stats']['mean_speed'])"
  echo "C2x8 PRUNE=$pr CFG=$cfg"; PMS_WARP_PRUNE=$pr PMS_WARP_CFG=$cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --peds 32 --worlds 8192 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launch']['blocks_per_sm'], d['episode_stats']['detections'], d['episode_stats']['robot_ped_collisions'], d['episode_stats']['mean_speed'])"
done; done
