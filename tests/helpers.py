"""Shared builders: the same scenario into the CPU oracle and into the CUDA path."""
import numpy as np

from oracle import oracle as orc


def cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def make_oracle(sc, det=(5.0, 120.0, 0), loop=True, robot=orc.ROBOT_UNICYCLE, visible=True, worlds=None):
    sel = slice(None) if worlds is None else worlds
    b = orc.OracleBatch(sc.poses[sel], vels=sc.vels[sel], vmag=sc.vmag[sel])
    b.set_waypoints_fixed(sc.table[sel], loop=loop)
    if robot != orc.ROBOT_NONE:
        b.set_robot(robot, sc.robot_pose[sel], control=sc.robot_control[sel], visible=visible)
    b.set_detector(det[0], np.deg2rad(det[1]), det[2])
    return b


def make_cuda(sc, precision, det=(5.0, 120.0, 0), loop=True, robot=2, visible=True, worlds=None):
    from pyminisim_b200.batched import BatchedSimulation
    sel = slice(None) if worlds is None else worlds
    poses = sc.poses[sel]
    sim = BatchedSimulation(poses.shape[0], poses.shape[1], precision=precision)
    sim.set_state(poses, sc.vels[sel])
    sim.set_speeds(sc.vmag[sel])
    sim.set_waypoints_fixed(sc.table[sel], loop=loop)
    if robot != 0:
        sim.set_robot(robot, sc.robot_pose[sel], control=sc.robot_control[sel], visible=visible)
    sim.set_detector(det[0], np.deg2rad(det[1]), det[2])
    return sim
