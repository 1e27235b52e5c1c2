"""Error of the CUDA path against the fp64 CPU oracle per precision mode and horizon (run on the GPU box).
Test/measurement tooling: imports the oracle as the checker only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_cuda, make_oracle  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    out = []
    for gen_name, gen in (("local-goals", make_bench_crowd), ("mirrored-goals", make_crowd)):
        for n, W in ((8, 16), (32, 16), (64, 16)):
            sc = gen(W, n)
            for prec in ("fp32", "fp32_plain", "mixed", "fp64"):
                o = make_oracle(sc)
                c = make_cuda(sc, prec)
                done = 0
                for h in (1, 100, 300, 1000):
                    o.rollout(0.01, h - done, n_threads=8)
                    c.step(0.01, h - done)
                    done = h
                    p, v = c.get_state()
                    fin = np.isfinite(o.poses).all(axis=(1, 2)) & (np.abs(o.vels[..., :2]).max(axis=(1, 2)) < 10)
                    # a world on the brink of the explicit-Euler blow-up may blow up on one side only: counted, not averaged
                    with np.errstate(invalid="ignore"):
                        cuda_ok = np.isfinite(p).all(axis=(1, 2)) & (np.abs(v[..., :2]).max(axis=(1, 2)) < 10)
                    diverged = int((fin & ~cuda_ok).sum())
                    fin &= cuda_ok
                    if not fin.any():
                        continue
                    perw = [rel(p[w], o.poses[w]) for w in np.nonzero(fin)[0]]
                    out.append(dict(scenario=gen_name, n=n, precision=prec, steps=h, worlds=int(fin.sum()), blown_up_on_gpu_only=diverged,
                                    pos_max=max(perw), pos_median=float(np.median(perw)),
                                    vel_max=rel(v[fin], o.vels[fin])))
                    print(json.dumps(out[-1]))
    return out


if __name__ == "__main__":
    main()
