#!/usr/bin/env python
"""bench.py -- HSFM pedestrian-updates/s of the fused B200 path (BASELINE.json metric).

A bench "step" = one pass of the hot path over one batch of synthetic worlds: `--sim-steps` fused
simulation steps (robot -> HSFM forces + Euler -> waypoints -> collision list + detector) of
`--worlds` worlds x `--peds` pedestrians per GPU.  Default workload = BASELINE config C4
(4096 worlds x 64 pedestrians + detector, dt = 0.01, 1000 simulation steps) IN TOTAL, sharded over the GPUs of the run
(`--scaling strong`, BASELINE.json: "4096 worlds x 64 pedestrians ... sharded across 1/2/4/8 B200"); with more than one GPU
the line also carries `weak_scaling` = the same metric with 4096 worlds PER GPU (`--scaling weak` makes that the headline).

  value : whole-job pedestrian-updates/s with state resident in HBM (CUDA events around the kernel,
          L2 flushed between timed steps, max over ranks).
  e2e   : the same metric through the public API with HOST buffers: pinned H2D upload of the initial
          state, the fused launch, D2H of final state + detector reading + statistics, every step.
  roofline    : FP32 pipe, algorithmic FLOPs (30 per directed pair, 30*N + 120 per pedestrian-update,
                SURVEY.md 8d) / kernel time vs the FP32 peak measured live by an FMA-chain kernel.
  cpu_baseline: the CPU oracle (C port of the reference's algorithm) on all host cores, bounded sample.

`--impl reference` times only that CPU port (the reference is Python/numba and cannot travel to the
GPU box) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hsfm_pedestrian_updates_per_sec"
UNIT = "pedestrian-updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds in total (strong scaling) / per GPU (weak scaling)")
    ap.add_argument("--peds", type=int, default=64)
    ap.add_argument("--sim-steps", type=int, default=1000, help="fused simulation steps per bench step")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_plain", "mixed", "fp64"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--j-split", type=int, default=0)
    ap.add_argument("--wpc", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time mixed / fp64 / single-step launches")
    return ap.parse_args()


KERNEL_NAMES = {0: "hsfm_small_kernel", 1: "hsfm_large_kernel", 3: "hsfm_warp_kernel", 4: "hsfm_duo_kernel"}
# Figures that only a profiler can measure -- DRAM bytes of the launch, executed FMA-pipe / XU-pipe / issue-slot utilisation --
# are read from profiles/ncu_captures.json: records written by `tools/ncu_summary.py --json` from an `ncu --set full` capture of
# exactly this command (kernel, worlds per GPU, pedestrians, fused steps, precision).  No matching record -> null.
def ncu_capture(kernel_name, worlds, peds, sim_steps, precision):
    path = os.path.join(ROOT, "profiles", "ncu_captures.json")
    try:
        with open(path) as f:
            recs = json.load(f)
    except Exception:
        return None
    for r in reversed(recs):            # the latest capture of this launch wins
        if (r.get("kernel") == kernel_name and r.get("worlds") == worlds and r.get("peds") == peds
                and r.get("sim_steps") == sim_steps and r.get("precision") == precision):
            return r
    return None


def flops_per_update(n_peds: int) -> float:
    return 30.0 * n_peds + 120.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(args, n_steps_timed: int, warmup: int):
    """Bounded sample of the workload on the CPU oracle, all host cores."""
    from oracle import oracle as orc
    from pyminisim_b200.scenarios import make_bench_crowd
    cores = os.cpu_count() or 1
    sample_worlds = min(args.worlds, max(16 * cores, 64))       # ~1 s of CPU work per timed step on all cores
    sample_steps = min(args.sim_steps, 1000)
    sc = make_bench_crowd(sample_worlds, args.peds)

    def fresh():
        b = orc.OracleBatch(sc.poses, vels=sc.vels, vmag=sc.vmag)
        b.set_waypoints_fixed(sc.table, loop=True)
        b.set_robot(orc.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
        b.set_detector(5.0, np.deg2rad(120.0), orc.DET_POLAR)
        return b

    for _ in range(warmup):
        fresh().rollout(0.01, sample_steps, n_threads=cores)
    t_total = 0.0
    for _ in range(n_steps_timed):
        b = fresh()
        t0 = time.perf_counter()
        b.rollout(0.01, sample_steps, n_threads=cores)
        t_total += time.perf_counter() - t0
    updates = sample_worlds * args.peds * sample_steps * n_steps_timed
    return {"value": updates / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_worlds} worlds x {args.peds} peds x {sample_steps} sim-steps per step, "
                      f"{n_steps_timed} steps, C oracle (fp64), pthreads over worlds",
            "ms_per_step": 1e3 * t_total / n_steps_timed}


def config_dict(args, world_size, precision=None):
    per = "per GPU" if args.scaling == "weak" else f"in total over {world_size} GPU(s)"
    return {"workload": f"C4: {args.worlds} worlds x {args.peds} pedestrians {per}, HSFM + unicycle robot + "
                        f"pedestrian detector (5 m, 120 deg) + collision list, jittered grid with local looped goals "
                        f"(scenarios.make_bench_crowd), {args.sim_steps} fused sim-steps "
                        f"of dt=0.01 per bench step",
            "worlds": args.worlds, "worlds_are": "per GPU" if args.scaling == "weak" else "total", "peds_per_world": args.peds, "sim_steps_per_step": args.sim_steps,
            "dt": 0.01, "precision": precision or args.precision, "parallelism": f"worlds sharded over {world_size} GPU(s), "
            "no collective on the step path", "l2": "256 MiB L2 flush between timed steps (state is 16 MB and "
            "lives on-chip during a step)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args, max(1, args.steps), max(0, min(args.warmup, 2)))
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.gpus, precision="fp64 (C port of the reference's algorithm)"),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is Python+numba and cannot travel to the GPU box; this arm is the C port of its "
                    "algorithm (oracle/), pinned to the live reference by tests/golden, on all host cores"}
    emit(line)


def emit(line: dict):
    """The ONE JSON line goes to the real stdout (saved before fd 1 was pointed at stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version banner on stdout) must not pollute the JSON line: everything written to
    fd 1 from here on lands on stderr; emit() writes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from pyminisim_b200 import _capi as capi
    from pyminisim_b200.batched import BatchedSimulation
    from pyminisim_b200.scenarios import make_bench_crowd

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pyminisim_b200.distributed import gather_episode_stats, local_episode_stats, max_over_ranks, shard_worlds
    if args.scaling == "weak":          # fixed work per GPU: rank r owns worlds [r*W, (r+1)*W)
        first_world, worlds = rank * args.worlds, args.worlds
    else:                               # fixed total work: the W worlds are split over the ranks
        first_world, worlds = shard_worlds(args.worlds, rank, world_size)
    sc = make_bench_crowd(worlds, args.peds, first_world=first_world)
    W, N, K = worlds, args.peds, args.sim_steps

    sim = BatchedSimulation(W, N, precision=args.precision, device=local_rank)
    if args.j_split or args.wpc or args.chunks or args.cluster:
        sim.set_tuning(args.j_split, args.wpc, args.chunks, args.cluster)
    # pinned host buffers for the end-to-end leg
    h_poses = capi.pinned_array((W, N, 3)); h_poses[:] = sc.poses
    h_vels = capi.pinned_array((W, N, 3)); h_vels[:] = sc.vels
    o_poses = capi.pinned_array((W, N, 3)); o_vels = capi.pinned_array((W, N, 3))
    sim.set_state(h_poses, h_vels)
    sim.set_speeds(sc.vmag)
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    sim.snapshot_save()

    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device_steps(n, n_sim, collect=True):
        total_ms = 0.0
        for _ in range(n):
            sim.snapshot_restore()
            with torch.cuda.stream(stream):
                flush.zero_()                                   # L2 flush, outside the timed events
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sim.step_async(0.01, n_sim)
                e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms

    # ---- warm-up ----
    timed_device_steps(max(args.warmup, 3), K)
    fp32_peak = capi.load().pms_measure_fp32_tflops(local_rank)

    # ---- device-resident throughput (value) ----
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches_before = sim.launch_info()["kernel_launches"]
    ms = timed_device_steps(args.steps, K)
    launches = sim.launch_info()["kernel_launches"] - launches_before
    barrier()
    ms_max = max_over_ranks(ms, device="cuda")
    total_worlds = args.worlds * world_size if args.scaling == "weak" else args.worlds
    updates_total = float(total_worlds) * N * K * args.steps
    value = updates_total / (ms_max * 1e-3)

    # ---- the other scaling mode as an extra key (N > 1): strong = BASELINE's split of 4096 worlds, weak = 4096 worlds per GPU ----
    other = None
    if world_size > 1:
        o_first, o_worlds = (rank * args.worlds, args.worlds) if args.scaling == "strong" else shard_worlds(args.worlds, rank, world_size)
        sc_o = make_bench_crowd(o_worlds, args.peds, first_world=o_first)
        sim_o = BatchedSimulation(o_worlds, N, precision=args.precision, device=local_rank)
        sim_o.set_state(sc_o.poses, sc_o.vels); sim_o.set_speeds(sc_o.vmag); sim_o.set_waypoints_fixed(sc_o.table, loop=True)
        sim_o.set_robot(capi.ROBOT_UNICYCLE, sc_o.robot_pose, control=sc_o.robot_control)
        sim_o.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
        sim_o.snapshot_save()
        st_o = torch.cuda.ExternalStream(sim_o.stream, device=torch.device("cuda", local_rank))

        def timed_other(n):
            total = 0.0
            for _ in range(n):
                sim_o.snapshot_restore()
                with torch.cuda.stream(st_o):
                    flush.zero_()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(st_o); sim_o.step_async(0.01, K); e1.record(st_o)
                e1.synchronize()
                total += e0.elapsed_time(e1)
            return total
        timed_other(3)
        barrier()
        ms_o = max_over_ranks(timed_other(args.steps), device="cuda")
        barrier()
        tot_o = args.worlds * world_size if args.scaling == "strong" else args.worlds
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "value": float(tot_o) * N * K * args.steps / (ms_o * 1e-3),
                 "unit": UNIT, "worlds_total": tot_o, "worlds_per_gpu": o_worlds, "ms_per_step": ms_o / args.steps,
                 "kernel": KERNEL_NAMES.get(sim_o.launch_info()["kernel"])}
        sim_o.close()

    # ---- end-to-end through the public API with host buffers ----
    # Every step: H2D of the step's initial state from pinned memory, the fused launch, D2H of the final state, the
    # detector reading, the collision list and the episode statistics.  Two handles alternate in asynchronous-I/O mode
    # (BatchedSimulation.set_async_io): the copies of one step overlap the kernel of the other, each step still pays
    # for its own transfers.  `value_serial_one_handle` is the same loop with one handle and blocking calls.
    # io dtype: the float32 host interface (pms_set/get_ped_state_f32, pms_get_detector_f32: the bytes the FP32 kernels
    # consume) is the headline; the reference's float64 arrays are timed next to it.
    sim2 = BatchedSimulation(W, N, precision=args.precision, device=local_rank)
    if args.j_split or args.wpc or args.chunks or args.cluster:
        sim2.set_tuning(args.j_split, args.wpc, args.chunks, args.cluster)
    sim2.set_state(h_poses, h_vels); sim2.set_speeds(sc.vmag); sim2.set_waypoints_fixed(sc.table, loop=True)
    sim2.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim2.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    sim2.snapshot_save()

    def run_e2e(io_dtype):
        f32 = io_dtype == np.float32
        in_p = capi.pinned_array((W, N, 3), io_dtype); in_p[:] = sc.poses
        in_v = capi.pinned_array((W, N, 3), io_dtype); in_v[:] = sc.vels
        bufs = [dict(poses=capi.pinned_array((W, N, 3), io_dtype), vels=capi.pinned_array((W, N, 3), io_dtype),
                     mask=capi.pinned_array((W, sim.words), np.uint32), vals=capi.pinned_array((W, N, 2), io_dtype),
                     coll=capi.pinned_array((W, sim.words), np.uint32), det=capi.pinned_array((W,), np.int64),
                     col=capi.pinned_array((W,), np.int64), nf=capi.pinned_array((W,), np.int32)) for _ in range(3)]

        def submit(s_, b_):
            s_.snapshot_restore()                  # robot / tracker / statistics back to t = 0 (device-side)
            s_.set_state(in_p, in_v)                # H2D from pinned memory
            s_.step_async(0.01, K)
            s_.get_state(b_["poses"], b_["vels"])   # D2H
            (s_.get_detector_f32 if f32 else s_.get_detector)(b_["mask"], b_["vals"])
            s_.get_collisions(b_["coll"])
            s_.get_stats((b_["det"], b_["col"], b_["nf"]))

        # one handle, blocking calls
        ser = bufs[2]
        for _ in range(2):
            submit(sim, ser)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            submit(sim, ser)
        torch.cuda.synchronize()
        serial = updates_total / max_over_ranks(time.perf_counter() - t0, device="cuda")
        # two handles, asynchronous I/O
        pipe = [(sim, bufs[0]), (sim2, bufs[1])]
        for s_, _ in pipe:
            s_.set_async_io(True)

        def pipelined(n):
            submit(*pipe[0])
            for k in range(1, n):
                submit(*pipe[k & 1])
                pipe[(k - 1) & 1][0].sync()        # the step's results are in its pinned buffers now
            pipe[(n - 1) & 1][0].sync()
        pipelined(2)
        barrier()
        t0 = time.perf_counter()
        pipelined(args.steps)
        torch.cuda.synchronize()
        piped = updates_total / max_over_ranks(time.perf_counter() - t0, device="cuda")
        for s_, _ in pipe:
            s_.set_async_io(False)
        last = pipe[(args.steps - 1) & 1][1]
        same = bool(np.array_equal(last["poses"], ser["poses"]) and np.array_equal(last["mask"], ser["mask"])
                    and np.array_equal(last["det"], ser["det"]))
        h2d_ = in_p.nbytes + in_v.nbytes
        d2h_ = sum(ser[k].nbytes for k in ("poses", "vels", "mask", "vals", "coll", "det", "col", "nf"))
        return piped, serial, h2d_, d2h_, same, ser

    e2e64 = run_e2e(np.float64)
    e2e_value, e2e_serial, h2d, d2h, e2e_matches, ser = run_e2e(np.float32)
    clocks = sampler.stop() if rank == 0 else None      # sampled over the device-resident and the end-to-end timed regions
    stats = (ser["det"], ser["col"], ser["nf"])
    o_vels = ser["vels"]
    sim2.close()

    # ---- episode statistics: the only collective (NCCL all-reduce of a few numbers) ----
    det, col, nf = stats
    g = gather_episode_stats(local_episode_stats(det, col, nf, o_vels), device="cuda")
    episode = {"detections": int(g["detections"]), "robot_ped_collisions": int(g["robot_ped_collisions"]),
               "nonfinite_worlds": int(g["nonfinite_worlds"]), "worlds": int(g["worlds"]),
               "mean_speed": g["speed_sum"] / max(g["pedestrians"], 1.0)}

    extra = {}
    if args.extra and rank == 0:
        for prec in ("fp32_plain", "mixed", "fp64"):
            s2 = BatchedSimulation(W, N, precision=prec, device=local_rank)
            s2.set_state(sc.poses, sc.vels); s2.set_speeds(sc.vmag); s2.set_waypoints_fixed(sc.table, loop=True)
            s2.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
            s2.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
            s2.snapshot_save()
            st2 = torch.cuda.ExternalStream(s2.stream)
            best = 0.0
            for rep in range(3):
                s2.snapshot_restore()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(st2); s2.step_async(0.01, K); e1.record(st2); e1.synchronize()
                if rep:
                    best = max(best, W * N * K / (e0.elapsed_time(e1) * 1e-3))
            extra[f"value_{prec}"] = best
            s2.close()
        # launch-per-step mode (state round-trips HBM every simulation step)
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(200):
            sim.step_async(0.01, 1)
        e1.record(stream); e1.synchronize()
        extra["value_one_sim_step_per_launch"] = W * N * 200 / (e0.elapsed_time(e1) * 1e-3)

    if rank == 0:
        info = sim.launch_info()
        kernel_ms = ms / max(launches, 1)
        cap = ncu_capture(KERNEL_NAMES.get(info["kernel"], str(info["kernel"])), W, N, K, args.precision)
        achieved_tflops = float(W) * N * K * flops_per_update(N) / (kernel_ms * 1e-3) / 1e12
        peak = fp32_peak if fp32_peak and fp32_peak > 0 else 148 * 128 * 2 * 1.965e9 / 1e12
        cpu = None
        if not args.no_cpu_baseline:
            res = cpu_reference_run(args, 3, 1)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": {"fp32": "f32", "fp32_plain": "f32", "mixed": "f32/f64", "fp64": "f64"}[args.precision],
            "data": "synthetic", "config": config_dict(args, world_size),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "io_dtype": "float32 host arrays (pms_set/get_ped_state_f32, pms_get_detector_f32)",
                    "how": "two handles alternate in asynchronous-I/O mode: the copies of one step overlap the kernel of the "
                           "other; every step pays its own H2D and D2H", "value_serial_one_handle": e2e_serial,
                    "pipelined_results_equal_serial": e2e_matches,
                    "float64_io": {"value": e2e64[0], "value_serial_one_handle": e2e64[1], "h2d_bytes_per_step": int(e2e64[2]),
                                   "d2h_bytes_per_step": int(e2e64[3]), "pipelined_results_equal_serial": e2e64[4]}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak,
                         "traffic": cap["dram_bytes"] if cap else None,
                         "fma_pipe_pct": cap["fma_pipe_cycles_pct"] if cap else None,
                         "xu_pipe_pct": cap["xu_pipe_pct"] if cap else None,
                         "issue_slots_pct": cap["issue_slots_pct"] if cap else None,
                         "executed_pipe_source": (cap["source"] + " (ncu --set full of this launch; executed, not algorithmic: "
                                                  "every unordered pair is evaluated once)") if cap else None,
                         "peak_source": "FP32 FMA-chain microbenchmark measured in this run (pms_measure_fp32_tflops); "
                                        "MEASURED_PEAKS.json has no non-tensor FP32 figure",
                         "kernel": KERNEL_NAMES.get(info["kernel"], str(info["kernel"])),
                         "flops_per_pedestrian_update": flops_per_update(N), "kernel_ms": kernel_ms,
                         "note": "achieved = ALGORITHMIC flops (30 per directed pair, SURVEY.md 8d) / measured kernel time. "
                                 "hsfm_warp_kernel evaluates every UNORDERED pair once (f_ij = -f_ji), i.e. executes about "
                                 "half of the algorithmic flops, so frac can exceed what the FMA pipe executes; the binding "
                                 "unit is the XU (MUFU) pipe: one rsqrt + one ex2 per unordered pair (profiles/)",
                         "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum per launch, " + cap["source"]) if cap else None},
            "cpu_baseline": cpu, "clocks": clocks, "launch": info, "episode_stats": episode,
        }
        if other is not None:
            line[other["scaling"] + "_scaling"] = other
        line.update(extra)
        emit(line)
    sim.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
