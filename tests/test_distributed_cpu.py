"""World sharding + episode-statistics gather with world_size 2 on the gloo backend (CPU)."""
import os
import subprocess
import sys

import numpy as np

from pyminisim_b200.distributed import shard_worlds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["PMS_ROOT"])
import numpy as np
import torch.distributed as dist
from pyminisim_b200.distributed import shard_worlds, gather_episode_stats, max_over_ranks, local_episode_stats
from pyminisim_b200.scenarios import make_bench_crowd
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
first, count = shard_worlds(10, rank, ws)
sc = make_bench_crowd(count, 8, first_world=first)            # each rank builds ONLY its shard
det = np.arange(first, first + count)
stats = gather_episode_stats(local_episode_stats(det, det * 2, np.zeros(count), sc.poses))
tmax = max_over_ranks(1.0 + rank)
if rank == 0:
    print(json.dumps({"stats": stats, "tmax": tmax, "first_pose": sc.poses[0, 0].tolist()}))
dist.destroy_process_group()
'''


def test_shard_worlds_partition():
    for total in (1, 7, 4096, 4097):
        for ws in (1, 2, 3, 8):
            blocks = [shard_worlds(total, r, ws) for r in range(ws)]
            assert sum(c for _, c in blocks) == total
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_rank_gloo_stats_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PMS_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["stats"]["detections"] == sum(range(10))
    assert res["stats"]["robot_ped_collisions"] == 2 * sum(range(10))
    assert res["stats"]["worlds"] == 10 and res["stats"]["pedestrians"] == 80
    assert res["tmax"] == 2.0
    from pyminisim_b200.scenarios import make_bench_crowd
    assert np.allclose(res["first_pose"], make_bench_crowd(1, 8).poses[0, 0])   # rank 0 owns world 0
