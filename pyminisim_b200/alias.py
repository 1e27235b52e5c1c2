"""``import pyminisim`` -> the accelerated package, so scripts written against the reference run unmodified.

    import pyminisim_b200.alias as alias
    alias.install_as_pyminisim()
    from pyminisim.core import Simulation                 # pyminisim_b200.core.Simulation
    from pyminisim.pedestrians import HeadedSocialForceModelPolicy, RandomWaypointTracker
    from pyminisim.sensors import PedestrianDetector

or, for a whole script (e.g. the reference's examples/basic.py), ``python -m pyminisim_b200.alias script.py [args]``.

The import paths SURVEY.md 8b lists are served: ``pyminisim.core`` (core/__init__.py:1-10), ``pyminisim.pedestrians``
(pedestrians/__init__.py:1-6), ``pyminisim.sensors`` (sensors/__init__.py:1-7), ``pyminisim.robot``, ``pyminisim.world_map``,
``pyminisim.util``.  ``pyminisim.visual`` (the pygame renderer) is outside the accelerated path: the alias serves a headless
stand-in whose ``Renderer`` accepts the reference's calls and draws nothing, so that a script's step loop still runs.
If the real reference is importable it is NOT replaced unless ``force=True``.
"""
from __future__ import annotations

import importlib
import runpy
import sys
import types


def _visual_module() -> types.ModuleType:
    m = types.ModuleType("pyminisim.visual")
    m.__doc__ = "Headless stand-in of pyminisim.visual: the pygame renderer is outside the accelerated path."

    class _Drawing:
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs

    class Renderer:
        """Accepts the calls of visual/_renderer.py (initialize / render / draw / clear_drawings / close); draws nothing."""

        def __init__(self, simulation=None, resolution: float = 50.0, screen_size=(500, 500), **kwargs):
            self._sim, self._drawings = simulation, {}

        def initialize(self):
            return None

        def render(self):
            return None

        def draw(self, drawing_id, drawing):
            self._drawings[drawing_id] = drawing

        def clear_drawings(self, drawing_ids=None):
            if drawing_ids is None:
                self._drawings.clear()
            else:
                for k in drawing_ids:
                    self._drawings.pop(k, None)

        def close(self):
            return None

    m.Renderer = Renderer
    for name in ("CircleDrawing", "Covariance2dDrawing", "AbstractDrawing"):
        setattr(m, name, type(name, (_Drawing,), {}))
    return m


def install_as_pyminisim(force: bool = False) -> types.ModuleType:
    """Register ``pyminisim`` and its sub-modules in ``sys.modules`` as views of this package.  Returns the alias package."""
    existing = sys.modules.get("pyminisim")
    if existing is not None and getattr(existing, "__pms_b200_alias__", False):
        return existing
    if not force:
        try:                                                  # the real reference wins unless the caller insists
            real = importlib.import_module("pyminisim")
            if not getattr(real, "__pms_b200_alias__", False):
                raise RuntimeError("the reference package `pyminisim` is importable; pass force=True to shadow it")
        except ImportError:
            pass
    for name in [n for n in sys.modules if n == "pyminisim" or n.startswith("pyminisim.")]:
        del sys.modules[name]
    pkg = types.ModuleType("pyminisim")
    pkg.__pms_b200_alias__ = True
    pkg.__path__ = []                                          # a package: `import pyminisim.core` resolves through sys.modules
    pkg.__doc__ = "alias of pyminisim_b200 (pyminisim_b200.alias.install_as_pyminisim)"
    sys.modules["pyminisim"] = pkg
    # core first: with the alias flag up it defines its own ABCs instead of looking for the reference
    for sub in ("core", "pedestrians", "sensors", "robot", "world_map", "util"):
        mod = importlib.import_module(f"pyminisim_b200.{sub}")
        sys.modules[f"pyminisim.{sub}"] = mod
        setattr(pkg, sub, mod)
    for sub, mod in (("visual", _visual_module()),):
        sys.modules[f"pyminisim.{sub}"] = mod
        setattr(pkg, sub, mod)
    return pkg


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m pyminisim_b200.alias script.py [args ...]")
    install_as_pyminisim(force=True)
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
