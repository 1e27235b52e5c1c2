"""float32 host interface, device-pointer state I/O and zero-copy result buffers of the C ABI (include/pms.h):
the device-resident path must equal the host path bit for bit."""
import numpy as np
import pytest

from helpers import cuda_available, make_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")]


def _scene(W=24, N=64):
    from pyminisim_b200.scenarios import make_bench_crowd
    return make_bench_crowd(W, N)


@pytest.mark.parametrize("precision", ["fp32", "fp32_plain", "mixed", "fp64"])
def test_float32_host_interface(precision):
    sc = _scene()
    a, b = make_cuda(sc, precision), make_cuda(sc, precision)
    p32, v32 = sc.poses.astype(np.float32), (sc.vels + 0.25).astype(np.float32)
    a.set_state(p32, v32)                                   # float32 interface
    b.set_state(p32.astype(np.float64), v32.astype(np.float64))   # the same values through the float64 interface
    for s in (a, b):
        s.step(0.01, 40)
    pa, va = a.get_state(dtype=np.float32)
    pb, vb = b.get_state()
    assert pa.dtype == np.float32 and pb.dtype == np.float64
    assert np.array_equal(pa, pb.astype(np.float32)) and np.array_equal(va, vb.astype(np.float32))   # correctly rounded hi + lo
    m32, d32 = a.get_detector_f32()
    m64, d64 = b.get_detector()
    assert np.array_equal(m32, m64) and np.array_equal(d32, d64.astype(np.float32))
    a.close(); b.close()


@pytest.mark.parametrize("precision,tdtype", [("fp32", "float32"), ("fp32", "float64"), ("fp64", "float64"), ("mixed", "float32")])
def test_device_pointer_path_equals_host_path(precision, tdtype):
    import torch
    from pyminisim_b200 import _capi as capi
    sc = _scene()
    dt = getattr(torch, tdtype)
    host, dev = make_cuda(sc, precision), make_cuda(sc, precision)
    np_dt = np.float32 if tdtype == "float32" else np.float64
    p0, v0 = sc.poses.astype(np_dt), (sc.vels + 0.25).astype(np_dt)
    host.set_state(p0, v0)
    tp, tv = torch.from_numpy(p0).cuda(), torch.from_numpy(v0).cuda()
    torch.cuda.synchronize()
    dev.set_state_torch(tp, tv)
    for s in (host, dev):
        s.step(0.01, 60)
    ph, vh = host.get_state(dtype=np_dt)
    op, ov = dev.get_state_torch(dtype=dt)
    dev.sync()
    assert np.array_equal(op.cpu().numpy(), ph) and np.array_equal(ov.cpu().numpy(), vh)     # bit for bit
    # zero-copy views of the result buffers
    W, N, words = dev.W, dev.N, dev.words
    mask, vals = host.get_detector()
    assert np.array_equal(dev.device_tensor(capi.BUF_DET_MASK, np.int32, (W, words)).cpu().numpy().view(np.uint32), mask)
    assert np.array_equal(dev.device_tensor(capi.BUF_DET_VALUES, np.float64, (W, N, 2)).cpu().numpy(), vals)
    assert np.array_equal(dev.device_tensor(capi.BUF_COLL_MASK, np.int32, (W, words)).cpu().numpy().view(np.uint32), host.get_collisions())
    rp, rv, _ = host.get_robot()
    assert np.array_equal(dev.device_tensor(capi.BUF_ROBOT_POSE, np.float64, (W, 3)).cpu().numpy(), rp)
    det, col, nf = host.get_stats()
    assert np.array_equal(dev.device_tensor(capi.BUF_DETECTIONS, np.int64, (W,)).cpu().numpy(), det)
    assert np.array_equal(dev.device_tensor(capi.BUF_WAYPOINT_INDEX, np.int32, (W, N)).cpu().numpy(), host.get_waypoints()[1])
    host.close(); dev.close()


def test_record_buffers_alias_the_records():
    from pyminisim_b200 import _capi as capi
    sc = _scene(8, 32)
    sim = make_cuda(sc, "fp32")
    assert sim.device_tensor(capi.BUF_REC_COLL_MASK, np.int32) is None            # not scheduled: no buffer
    sim.schedule_sensor(capi.SENSOR_DETECTOR, 0.02, values=True)
    sim.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
    sim.step(0.01, 30)
    dm, dv = sim.detector_records(values=True)
    cm = sim.collision_records()
    F = len(dm)
    assert F == len(sim.sensor_fires(capi.SENSOR_DETECTOR)[0]) and len(cm) == 30
    assert np.array_equal(sim.device_tensor(capi.BUF_REC_DET_MASK, np.int32, (F, sim.W, sim.words)).cpu().numpy().view(np.uint32), dm)
    assert np.array_equal(sim.device_tensor(capi.BUF_REC_DET_VALUES, np.float64, (F, sim.W, sim.N, 2)).cpu().numpy(), dv)
    assert np.array_equal(sim.device_tensor(capi.BUF_REC_COLL_MASK, np.int32, (30, sim.W, sim.words)).cpu().numpy().view(np.uint32), cm)
    sim.close()


def test_device_resident_rollout_loop():
    """An RL-style loop that never leaves the GPU: controls from a CUDA tensor, state and readings as CUDA tensors / aliases;
    equal to the host-array path bit for bit."""
    import torch
    from pyminisim_b200 import _capi as capi
    sc = _scene(16, 32)
    rng = np.random.default_rng(9)
    ctr = np.stack([np.concatenate([rng.uniform(0.2, 1.0, (16, 1)), rng.uniform(-0.5, 0.5, (16, 1))], axis=1) for _ in range(60)])
    host, dev = make_cuda(sc, "fp32"), make_cuda(sc, "fp32")
    tctr = torch.from_numpy(ctr).cuda()
    torch.cuda.synchronize()
    for k in range(0, 60, 20):
        host.set_robot_controls(ctr[k:k + 20]); host.step(0.01, 20)
        dev.set_robot_controls_torch(tctr[k:k + 20]); dev.step_async(0.01, 20)
    poses, vels = dev.get_state_torch(torch.float32)
    det = dev.device_tensor(capi.BUF_DET_MASK, np.int32, (dev.W, dev.words))
    rp = dev.device_tensor(capi.BUF_ROBOT_POSE, np.float64, (dev.W, 3))
    dev.sync()
    ph, vh = host.get_state(dtype=np.float32)
    assert np.array_equal(poses.cpu().numpy(), ph) and np.array_equal(vels.cpu().numpy(), vh)
    assert np.array_equal(det.cpu().numpy().view(np.uint32), host.get_detector()[0])
    assert np.array_equal(rp.cpu().numpy(), host.get_robot()[0])
    host.close(); dev.close()
