"""The drop-in surface: `pyminisim.*` import paths through the alias (reference pyminisim/core/__init__.py:1-10,
pedestrians/__init__.py:1-6, sensors/__init__.py:1-7), the reference's own ABCs as base classes when the reference is
importable, and the lazy array-backed state objects (core/_pedestrians_model.py:21-27: fresh copies on every access).
Import-time decisions are checked in fresh interpreters so that the other tests keep their module state."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def _run(code: str, extra_path=()):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, *extra_path]))
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_alias_serves_the_reference_import_paths():
    out = _run("""
        import pyminisim_b200.alias as alias
        alias.install_as_pyminisim()
        from pyminisim.core import Simulation, SimulationState, WorldState, AbstractSensor, AbstractPedestriansModel, PEDESTRIAN_RADIUS
        from pyminisim.pedestrians import (HSFMParams, HeadedSocialForceModelPolicy, RandomWaypointTracker, FixedWaypointTracker,
                                           ORCAParams, ORCAPedestriansModel, ReplayPedestriansPolicy, ETHPedestriansRecord)
        from pyminisim.sensors import (PedestrianDetectorNoise, PedestrianDetectorConfig, PedestrianDetectorReading, PedestrianDetector,
                                       OmniObstacleDetectorNoise, OmniObstacleDetectorConfig, OmniObstacleDetectorReading, OmniObstacleDetector,
                                       LidarSensorConfig, LidarSensorReading, LidarSensorNoise, LidarSensor,
                                       SemanticDetectorNoise, SemanticDetectorConfig, SemanticDetection, SemanticDetectorReading, SemanticDetector)
        from pyminisim.robot import UnicycleRobotModel, BicycleRobotModel
        from pyminisim.world_map import CirclesWorld, EmptyWorld, AABBWorld, AABBObject
        from pyminisim.util import wrap_angle, angle_linspace, point_to_segment_distance, closest_point_of_segment
        from pyminisim.visual import Renderer, CircleDrawing
        import pyminisim_b200.core as c, pyminisim_b200.pedestrians as p
        assert Simulation is c.Simulation and HeadedSocialForceModelPolicy is p.HeadedSocialForceModelPolicy
        assert abs(wrap_angle(3.5) - (3.5 - 2 * 3.141592653589793)) < 1e-15 and len(angle_linspace(0.0, 1.0, 4)) == 5
        r = Renderer(simulation=None, resolution=80.0, screen_size=(700, 700)); r.initialize(); r.render(); r.close()
        print("ok", c.USING_REFERENCE_ABCS)
    """)
    assert out.split()[-2:] == ["ok", "False"]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only present in the build container")
def test_models_subclass_the_reference_abcs_when_the_reference_is_importable():
    out = _run("""
        import pyminisim.core as ref                      # the real reference (its pedestrians sub-package needs rvo2, core does not)
        import pyminisim_b200.core as c, pyminisim_b200.pedestrians as p, pyminisim_b200.sensors as s
        import pyminisim_b200.robot as r, pyminisim_b200.world_map as w
        assert c.USING_REFERENCE_ABCS
        assert issubclass(p.HeadedSocialForceModelPolicy, ref.AbstractPedestriansModel)
        assert issubclass(p.ORCAPedestriansModel, ref.AbstractPedestriansModel)
        assert issubclass(p.HSFMState, ref.AbstractPedestriansModelState)
        assert issubclass(p.RandomWaypointTracker, ref.AbstractWaypointTracker)
        assert issubclass(s.PedestrianDetector, ref.AbstractSensor) and issubclass(s.LidarSensor, ref.AbstractSensor)
        assert issubclass(r.UnicycleRobotModel, ref.AbstractRobotMotionModel)
        assert issubclass(w.CirclesWorld, ref.AbstractWorldMap)
        assert c.WorldState is ref.WorldState and c.SimulationState is ref.SimulationState
        print("ok")
    """, extra_path=[REFERENCE])
    assert out.strip().endswith("ok")


def test_array_backed_state_is_lazy_and_returns_fresh_copies():
    from pyminisim_b200.pedestrians import FixedWaypointTrackerState, HSFMState
    rng = np.random.default_rng(0)
    poses, vels = rng.normal(size=(5, 3)), rng.normal(size=(5, 3))
    wp = FixedWaypointTrackerState({i: np.zeros(2) for i in range(5)})
    st = HSFMState((poses.copy(), vels.copy()), wp)
    assert st._dict is None                                       # nothing per-pedestrian was built
    a, b = st.poses, st.poses
    assert list(a.keys()) == list(range(5)) and a[3].shape == (3,)
    assert np.array_equal(a[3], poses[3]) and a[3] is not b[3]
    a[3][:] = 99.0                                                # callers may mutate what they get (ownership: copies)
    assert np.array_equal(st.poses[3], poses[3]) and np.array_equal(st.velocities[2], vels[2])
    assert np.array_equal(st.pose_array, poses) and st.waypoints is wp
    # the reference's signature still works (reset_to_state of user code), with the reference's assertions
    st2 = HSFMState({7: (poses[0], vels[0]), 9: (poses[1], vels[1])}, wp)
    assert list(st2.poses.keys()) == [7, 9] and np.array_equal(st2.velocities[9], vels[1])
    with pytest.raises(AssertionError):
        HSFMState({0: (np.zeros(2), np.zeros(3))}, wp)
    # code written against the base class reads _pedestrians
    assert np.array_equal(st._pedestrians[4][1], vels[4])


@pytest.mark.gpu
def test_kat3_through_the_alias(golden):
    """examples/basic.py:13-60 with its own import lines (`from pyminisim... import ...`), Simulation.seed(42), 1001 steps."""
    g = golden("random_tracker")
    np.save("/tmp/_kat3_init.npy", g["init_poses"])
    out = _run("""
        import numpy as np
        import pyminisim_b200.alias as alias
        alias.install_as_pyminisim()
        from pyminisim.core import Simulation
        from pyminisim.world_map import EmptyWorld, CirclesWorld
        from pyminisim.robot import UnicycleRobotModel
        from pyminisim.pedestrians import HeadedSocialForceModelPolicy, RandomWaypointTracker, FixedWaypointTracker
        from pyminisim.sensors import PedestrianDetectorNoise, PedestrianDetector, LidarSensor, LidarSensorNoise
        from pyminisim.visual import Renderer, CircleDrawing
        Simulation.seed(42)
        robot_model = UnicycleRobotModel(initial_pose=np.array([0., 0., 0.0]), initial_control=np.array([0.0, np.deg2rad(25.0)]))
        tracker = RandomWaypointTracker(world_size=(7.0, 7.0))
        pedestrians_model = HeadedSocialForceModelPolicy(n_pedestrians=2, waypoint_tracker=tracker,
                                                         pedestrian_linear_velocity_magnitude=np.array([1.5, 2.5]),
                                                         initial_poses=np.load("/tmp/_kat3_init.npy"))
        sensors = [PedestrianDetector(noise=None)]
        sim = Simulation(world_map=EmptyWorld(), robot_model=robot_model, pedestrians_model=pedestrians_model, sensors=sensors, rt_factor=None)
        renderer = Renderer(simulation=sim, resolution=80.0, screen_size=(700, 700))
        renderer.initialize()
        for _ in range(1001):
            renderer.render()
            sim.step()
        p = sim.current_state.world.pedestrians.poses
        print(repr([list(map(float, p[i])) for i in range(2)]))
    """)
    poses = np.array(eval(out.strip().splitlines()[-1]))
    k = list(g["horizons"]).index(1001) if 1001 in list(g["horizons"]) else None
    # SURVEY.md Appendix C KAT-3 (8-digit print of the live reference)
    kat = np.array([[0.90157788, -1.67980664, -3.42613717], [-0.0380459, 0.8001509, 4.96696964]])
    assert np.abs(poses - kat).max() < 5e-6, poses
    if k is not None:
        assert np.abs(poses - g["poses"][k]).max() < 1e-6
