"""Device-resident throughput of every BASELINE.json configuration (run on the GPU box):
  C2  1024 worlds x 32 pedestrians HSFM, 1000 steps
  C3  1 world x 65536 pedestrians HSFM (large-crowd kernel), 20 steps
  C4  4096 worlds x 64 pedestrians HSFM + detector, 1000 steps (bench.py's workload)
  C5  2048 worlds x 64 agents ORCA, 200 steps
  sensors: stand-alone collision/detector, lidar and omni passes over 4096 worlds (wall clock, results on the host)
CUDA events on the handle's stream, best of 3 after one warm-up.  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedOrca, BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd  # noqa: E402


def time_hsfm(name, W, N, steps, precision="fp32", tuning=None, large=False, reps=3, shuffle=False):
    if large:
        sc = make_crowd(W, N, side=2.0 * np.sqrt(N), local_goals=5.0)
    else:
        sc = make_bench_crowd(W, N)
    if shuffle:
        perm = np.random.default_rng(0).permutation(N)
        sc.poses, sc.table, sc.vmag, sc.vels = sc.poses[:, perm], sc.table[:, perm], sc.vmag[:, perm], sc.vels[:, perm]
    sim = BatchedSimulation(W, N, precision=precision)
    sim.set_state(sc.poses, sc.vels)
    sim.set_speeds(sc.vmag)
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    if tuning:
        sim.set_tuning(*tuning)
    sim.snapshot_save()
    stream = torch.cuda.ExternalStream(sim.stream)
    best = float("inf")
    for rep in range(reps + 1):
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); sim.step_async(0.01, steps); e1.record(stream); e1.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    info = sim.launch_info()
    _, _, nf = sim.get_stats()
    out = dict(case=name, worlds=W, peds=N, steps=steps, precision=precision, ms=best,
               ped_updates_per_s=W * N * steps / (best * 1e-3), us_per_step=1e3 * best / steps,
               algorithmic_tflops=W * N * steps * (30.0 * N + 120.0) / (best * 1e-3) / 1e12,
               kernel=info["kernel"], blocks=info["blocks"], threads=info["threads_per_block"],
               worlds_per_block=info["worlds_per_block"], n_chunks=info["n_chunks"], nonfinite_worlds=int((nf != 0).sum()))
    sim.close()
    print(json.dumps(out), flush=True)
    return out


def time_contact_heavy(W=4096, N=64, steps=1000, side=40.0, reps=3):
    """A crowd that really touches: MIRRORED goals (everybody crosses the centre of a 40 m square), the scenario
    the benchmark avoids because the reference's explicit Euler blows up in 17 % of such worlds.  Candidates are simulated once,
    the first W worlds that stay finite for `steps` steps make the batch; the body-contact pass (k_1, k_2 terms) then runs in
    a large share of the steps.  Reports the time next to the local-goal benchmark crowd and how often bodies touched."""
    cand = make_crowd(2 * W, N, side=side)

    def build(sc, tuning=None):
        sim = BatchedSimulation(sc.poses.shape[0], N, precision="fp32")
        sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
        sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
        sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
        if tuning:
            sim.set_tuning(*tuning)
        return sim
    probe = build(cand)
    probe.step(0.01, steps)
    nf = probe.get_stats()[2]
    probe.close()
    finite_share = float((nf == 0).mean())
    keep = np.flatnonzero(nf == 0)[:W]
    for f in ("poses", "vels", "vmag", "table", "robot_pose", "robot_control"):
        setattr(cand, f, np.ascontiguousarray(getattr(cand, f)[keep]))
    # how often do bodies touch?  closest approach per world over the episode, sampled every 10 steps on the CPU
    sim = build(cand)
    touching_steps, samples = 0, 0
    for _ in range(steps // 10):
        sim.step(0.01, 10)
        p = sim.get_state()[0][:64, :, :2]
        d = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1) + 10.0 * np.eye(N)
        touching_steps += int((d.min(axis=(1, 2)) < 0.6).sum()); samples += p.shape[0]
    sim.close()
    sim = build(cand)
    sim.snapshot_save()
    stream = torch.cuda.ExternalStream(sim.stream)
    best = float("inf")
    for rep in range(reps + 1):
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); sim.step_async(0.01, steps); e1.record(stream); e1.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    info = sim.launch_info()
    nf = sim.get_stats()[2]
    out = dict(case=f"C4 contact-heavy: mirrored goals in a {side:g} m square, worlds that stay finite", worlds=len(keep), peds=N, steps=steps,
               ms=best, ped_updates_per_s=len(keep) * N * steps / (best * 1e-3), kernel=info["kernel"],
               finite_share_of_candidate_worlds=finite_share, nonfinite_worlds_in_the_timed_batch=int((nf != 0).sum()),
               worlds_with_body_contact_per_sampled_step=touching_steps / max(samples, 1))
    sim.close()
    print(json.dumps(out), flush=True)
    return out


def time_orca(name, W, N, steps, reps=3):
    sc = make_crowd(W, N, side=16.0)
    sim = BatchedOrca(0.01, W, N)
    best = float("inf")
    for rep in range(reps + 1):
        sim.set_agents(sc.poses[..., :2], max_speeds=np.full((W, N), 1.7), pref_speeds=np.full((W, N), 1.5))
        sim.set_waypoints_fixed(sc.table, loop=True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(steps); e1.record(); e1.synchronize()   # pms_orca_step synchronises its own stream
        if rep:
            best = min(best, e0.elapsed_time(e1))
    out = dict(case=name, worlds=W, agents=N, steps=steps, ms=best, agent_updates_per_s=W * N * steps / (best * 1e-3))
    sim.close()
    print(json.dumps(out), flush=True)
    return out


def time_sensors(W=4096, N=64, n_circles=8, n_beams=100, reps=3):
    """Stand-alone sensor passes over W worlds, through the C ABI with host result buffers (their calls synchronise):
    collision list + pedestrian detector (`pms_sense`), lidar against a per-world CirclesWorld (`pms_lidar_scan`: 100 beams,
    8 m, 1 cm samples like sensors/_lidar.py), omni obstacle detector (`pms_omni_scan`).  Wall clock around the call."""
    import time
    sc = make_bench_crowd(W, N)
    sim = BatchedSimulation(W, N, precision="fp32")
    sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    rng = np.random.default_rng(3)
    circles = np.concatenate([sc.robot_pose[:, None, :2] + rng.uniform(-6.0, 6.0, (W, n_circles, 2)),
                              rng.uniform(0.3, 1.0, (W, n_circles, 1))], axis=2)
    circles[:, :, 0] += 2.0 * np.sign(circles[:, :, 0] - sc.robot_pose[:, None, 0] + 1e-9)   # keep the robot outside
    sim.set_map(capi.MAP_CIRCLES, circles)
    mask = np.empty((W, sim.words), dtype=np.uint32); vals = np.empty((W, N, 2))

    def best_of(f):
        f(); torch.cuda.synchronize()
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
        return best

    t = best_of(lambda: (sim.sense(), sim.get_detector(mask, vals), sim.get_collisions()))
    print(json.dumps(dict(case="sense: collision list + detector, results on the host", worlds=W, peds=N, ms=1e3 * t,
                          ped_checks_per_s=W * N / t)), flush=True)
    hits = sim.lidar_scan(n_beams, 8.0, 0.01)[0]
    t = best_of(lambda: sim.lidar_scan(n_beams, 8.0, 0.01))
    print(json.dumps(dict(case="lidar: 100 beams x 800 samples, CirclesWorld of 8 per world, results on the host", worlds=W,
                          beams=n_beams, circles=n_circles, ms=1e3 * t, beams_per_s=W * n_beams / t,
                          reference_point_tests_per_s=W * n_beams * 800 * n_circles / t,
                          beams_with_hit=float((hits >= 0).mean()))), flush=True)
    t = best_of(lambda: sim.omni_scan())
    print(json.dumps(dict(case="omni obstacle detector", worlds=W, circles=n_circles, ms=1e3 * t, worlds_per_s=W / t)), flush=True)
    sim.close()


def main():
    which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5", "sensors"}
    if "c2" in which:
        time_hsfm("C2 warp kernel", 1024, 32, 1000)
        time_hsfm("C2 warp kernel, plain fp32 state", 1024, 32, 1000, precision="fp32_plain")
        time_hsfm("C2 directed-pair kernel", 1024, 32, 1000, tuning=(-1, 0, 0, 0))
        time_hsfm("C2 mixed", 1024, 32, 1000, precision="mixed")
    if "c4" in which:
        time_hsfm("C4 warp kernel", 4096, 64, 1000)
        time_hsfm("C4 warp kernel, plain fp32 state", 4096, 64, 1000, precision="fp32_plain")
        time_hsfm("C4 directed-pair kernel", 4096, 64, 1000, tuning=(-1, 0, 0, 0))
        time_hsfm("C4 mixed", 4096, 64, 1000, precision="mixed")
        time_hsfm("C4 fp64", 4096, 64, 200, precision="fp64")
        time_hsfm("4096 x 128 warp kernel", 4096, 128, 300)
        time_hsfm("4096 x 128 directed-pair kernel", 4096, 128, 300, tuning=(-1, 0, 0, 0))
        time_hsfm("16384 x 32 warp kernel", 16384, 32, 1000)
    if "duo" in which:      # warp-per-world kernel (j_split 5) vs pair-warp / owner-warp kernel (j_split 4)
        for W, N in ((1024, 32), (512, 64), (256, 64), (592, 64), (1024, 64), (512, 32), (2048, 32), (128, 64), (512, 48), (1024, 20)):
            time_hsfm(f"{W} x {N} warp-per-world kernel", W, N, 1000, tuning=(5, 0, 0, 0))
            time_hsfm(f"{W} x {N} duo kernel", W, N, 1000, tuning=(4, 0, 0, 0))
    if "contact" in which:
        time_hsfm("C4 warp kernel (local looped goals: the benchmark crowd)", 4096, 64, 1000)
        time_contact_heavy()
        time_contact_heavy(side=24.0)
    if "c3" in which:
        time_hsfm("C3 large crowd", 1, 65536, 20, large=True)
        time_hsfm("C3 large crowd, 200 steps per call", 1, 65536, 200, large=True)
        time_hsfm("C3 large crowd, caller's order kept (no sorted working copy)", 1, 65536, 20, large=True, tuning=(0, 0, -2, 0))
        time_hsfm("C3 large crowd, every j-tile visited (no exact pruning)", 1, 65536, 20, large=True, tuning=(0, 0, -1, 0))
        time_hsfm("C3 large crowd, shuffled pedestrian order", 1, 65536, 20, large=True, shuffle=True)
        time_hsfm("C3 large crowd, shuffled pedestrian order, caller's order kept", 1, 65536, 20, large=True, shuffle=True, tuning=(0, 0, -2, 0))
        time_hsfm("C3 large crowd, MIXED", 1, 65536, 20, large=True, precision="mixed")
    if "c5" in which:
        time_orca("C5 ORCA", 2048, 64, 200)
    if "sensors" in which:
        time_sensors()


if __name__ == "__main__":
    main()
