"""Load the unmodified reference (jviquerat/lbm): oracle pinning, the reference's own driver loop
and app classes on the GPU path, and the Numba CPU baseline of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under lbm_b200/ imports this.

Where the sources come from: /root/reference in the build container; on the GPU box (no
/root/reference there) the git-ignored copy baseline/_ref/ made by tools/make_ref_copy.py, which
travels with the gpurun snapshot like the built .so files.  Either way the reference is imported
from a scratch COPY (Numba cache=True would write __pycache__ next to the sources, SURVEY.md
section 10.1) with a stub matplotlib (the only missing dependency).
"""
import os
import shutil
import sys
import tempfile
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference", os.path.join(_REPO, "baseline", "_ref"))
_state = {}


def reference_root():
    for r in _CANDIDATES:
        if os.path.isdir(os.path.join(r, "lbm", "src", "core")):
            return r
    return None


REFERENCE_ROOT = reference_root() or _CANDIDATES[0]


def available():
    return reference_root() is not None


def _stub_matplotlib():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_lbm_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    mpl._lbm_stub = True
    plt = types.ModuleType("matplotlib.pyplot")
    cm = types.ModuleType("matplotlib.cm")

    class _Dummy:
        def __getattr__(self, name):
            return lambda *a, **k: _Dummy()

        def __call__(self, *a, **k):
            return _Dummy()

    def _noop(*a, **k):
        return _Dummy()

    for name in ("clf", "imshow", "axis", "savefig", "close", "imsave", "figaspect", "plot",
                 "fill", "scatter", "xlim", "ylim", "gca", "figure", "contour", "streamplot",
                 "annotate", "cla"):
        setattr(plt, name, _noop)
    plt.subplots = lambda *a, **k: (_Dummy(), _Dummy())
    cm.ocean = _Dummy()
    mpl.pyplot = plt
    mpl.cm = cm
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.cm"] = cm


def load():
    """Returns a namespace with the reference's modules; idempotent."""
    if _state:
        return _state["ns"]
    if not available():
        raise RuntimeError("reference not present (looked in %s)" % ", ".join(_CANDIDATES))
    scratch = tempfile.mkdtemp(prefix="lbm_ref_")
    shutil.copytree(os.path.join(reference_root(), "lbm"), os.path.join(scratch, "lbm"),
                    ignore=shutil.ignore_patterns("save", "__pycache__"))
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(scratch, "numba_cache"))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    _stub_matplotlib()
    import PIL.Image  # noqa: F401  (shapes.py uses PIL.Image / PIL.ImageChops)
    import PIL.ImageChops  # noqa: F401
    sys.path.insert(0, scratch)
    cwd = os.getcwd()
    os.chdir(scratch)  # the reference writes ./results/<timestamp>/ relative to cwd
    try:
        import lbm.src.core.lattice as ref_lattice
        import lbm.src.core.nb as ref_nb
        import lbm.src.core.run as ref_run
        import lbm.src.app.app as ref_app
        import lbm.src.utils.shapes as ref_shapes
        import lbm.src.core.obstacle as ref_obstacle
        import lbm.src.utils.buff as ref_buff
    finally:
        os.chdir(cwd)
    # shape.generate_image trims a PNG it expects matplotlib to have written
    ref_shapes.shape.generate_image = lambda self, *a, **k: None
    ns = types.SimpleNamespace(scratch=scratch, lattice=ref_lattice, nb=ref_nb, run=ref_run,
                               app=ref_app, shapes=ref_shapes, obstacle=ref_obstacle,
                               buff=ref_buff)
    _state["ns"] = ns
    return ns


class in_scratch:
    """Context manager: run reference code with cwd = scratch (for ./results)."""

    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(load().scratch)

    def __exit__(self, *a):
        os.chdir(self.cwd)
