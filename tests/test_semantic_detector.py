"""SemanticDetector (sensors/_semantic_object_detector.py) and AABBWorld.closest_semantic_objects (_aabb_world.py:166-186) against
the live reference (tests/golden/semantic.npz).  Host arithmetic over a handful of boxes; the GPU test runs it inside this
package's Simulation (device robot step + map veto, the reference's 10 Hz hold-time scheduler, the global numpy generator)."""
import numpy as np
import pytest

from helpers import cuda_available

OBJECTS = [((3.0, 1.0, 2.0, 1.5), "table", "t1"), ((-2.0, -4.0, 1.0, 3.0), "shelf", None), ((0.5, 3.0, 0.5, 0.5), None, None),
           ((-4.0, 2.0, 3.0, 0.2), "door", "d7"), ((9.0, 9.0, 1.0, 1.0), "far", "f")]


def _world():
    from pyminisim_b200.world_map import AABBObject, AABBWorld
    return AABBWorld([AABBObject(box, class_name=c, object_name=o) for box, c, o in OBJECTS])


def test_closest_semantic_objects_matches_reference(golden):
    g = golden("semantic")
    world = _world()
    rows = [[q, idx, *v[2]] for q, p in enumerate(g["query_points"]) for idx, v in world.closest_semantic_objects(p, 2.5).items()]
    assert np.array_equal(np.array(rows), g["query_rows"])
    near = world.closest_semantic_objects(np.array([2.0, 2.0]), 10.0)
    assert set(near) == {0, 1, 3, 4} and near[0][:2] == ("table", "t1") and near[1][:2] == ("shelf", None)   # anonymous boxes never show
    from pyminisim_b200.world_map import AABBObject, AABBWorld
    with pytest.raises(ValueError, match="class_name must be specified"):
        AABBWorld([AABBObject((0., 0., 1., 1.), object_name="x")])


def test_noise_and_config_contract():
    from pyminisim_b200.sensors import (SemanticDetection, SemanticDetector, SemanticDetectorConfig, SemanticDetectorNoise,
                                        SemanticDetectorReading)
    assert SemanticDetector().period == pytest.approx(0.1) and SemanticDetector.NAME == "semantic_detector"
    assert SemanticDetector(SemanticDetectorConfig(frequency=np.inf)).period == 0.
    with pytest.raises(AssertionError):
        SemanticDetectorConfig(max_dist=0.)
    with pytest.raises(AssertionError):
        SemanticDetectorNoise(drop_prob=1.5)
    reading = SemanticDetectorReading({k: SemanticDetection("c", None, np.array([float(k), 0.])) for k in range(200)})
    np.random.seed(4)
    out = SemanticDetectorNoise(drop_prob=0.25).noisify_reading(reading).detections
    np.random.seed(4)
    drawn = np.random.binomial(1, 0.75, 200).astype(bool)
    shift = np.random.multivariate_normal(np.zeros(2), np.diag([0.001, 0.001]), 200)
    assert sorted(out) == [k for k in range(200) if not drawn[k]]            # the reference keeps a detection where the draw is 0
    assert all(np.array_equal(out[k].position, np.array([float(k), 0.]) + shift[k]) for k in out)


@pytest.mark.gpu
@pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")
@pytest.mark.parametrize("tag", ["clean", "noisy"])
def test_inside_simulation_matches_reference(golden, tag):
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import SemanticDetector, SemanticDetectorConfig, SemanticDetectorNoise
    want = golden("semantic")[f"{tag}_rows"]
    Simulation.seed(11)
    rm = UnicycleRobotModel(initial_pose=np.array([0.0, 0.0, 0.5]), initial_control=np.array([1.5, 0.9]))
    noise = SemanticDetectorNoise(drop_prob=0.4) if tag == "noisy" else None
    sim = Simulation(world_map=_world(), robot_model=rm, pedestrians_model=None,
                     sensors=[SemanticDetector(SemanticDetectorConfig(max_dist=3.0, frequency=10.), noise=noise)], sim_dt=0.01, rt_factor=None)
    rows, last = [], None
    for k in range(1, 401):
        reading = sim.step().sensors["semantic_detector"].reading
        if reading is not last:
            last = reading
            rows += [[k, idx, *d.position] for idx, d in reading.detections.items()]
            rows.append([k, -1, 0., 0.])
    rows = np.array(rows)
    assert np.array_equal(rows[:, :2], want[:, :2])                           # firing steps (1, 11, 22, ...) and who is seen
    assert np.allclose(rows[:, 2:], want[:, 2:], rtol=0, atol=1e-12)          # the robot pose carries FMA contraction
