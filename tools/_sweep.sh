python -m pytest tests/test_gpu_warp.py tests/test_gpu_hsfm.py tests/test_gpu_dropin.py -q 2>&1 | tail -5
for mb in 2 3 4; do for wpc in 4 8; do
  echo "MINB=$mb wpc=$wpc"; PMS_WARP_MINB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline --wpc $wpc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launch']['blocks_per_sm'], d['launch']['n_chunks'], d['episode_stats']['detections'], d['episode_stats']['robot_ped_collisions'])"
done; done
