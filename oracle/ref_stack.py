"""ref_stack.py — run the reference's OWN Python (unmodified copies staged by oracle/build_ref.ship_python under oracle/_ref/py)
over either backend, in one process:

    load("ours")  nerf/network_ff.py + nerf/renderer.py + encoding.py + activation.py (+ sdf/)  over THIS repo's drop-in packages
                  (torch-ngp_b200/{gridencoder,ffmlp,shencoder,raymarching})
    load("ref")   the same callers over the reference's own wrapper packages (gridencoder/grid.py, ffmlp/ffmlp.py,
                  shencoder/sphere_harmonics.py, raymarching/raymarching.py) and its own CUDA extensions (oracle/_ref/*.so, bound
                  under the module names `_gridencoder`, `_ffmlp`, `_shencoder`, `_raymarching` those wrappers import)

TEST / MEASUREMENT INFRASTRUCTURE ONLY (tests/, bench_ref_cuda.py, bench.py's reference arms).  Nothing under torch-ngp_b200/ imports
this.  Both stacks use the same top-level module names, so each keeps its modules in a private table that is swapped into sys.modules
while the stack is `active()`; run model construction and calls inside that context (encoding.get_encoder imports lazily).

Third-party modules the reference imports at module level but never touches on this path (trimesh, mcubes, tensorboardX, lpips,
torch_ema, torchmetrics, imageio, matplotlib, tkinter via `from turtle import ...`) are replaced by inert stubs when they are not
installed; the reference sources themselves are not edited.
"""
import contextlib
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
PKG = os.path.join(ROOT, "torch-ngp_b200")
PY = os.path.join(_HERE, "_ref", "py")
CALLERS, WRAPPERS = os.path.join(PY, "callers"), os.path.join(PY, "wrappers")

_CALLER_ROOTS = ("nerf", "sdf", "encoding", "activation", "loss")
_PKG_ROOTS = ("gridencoder", "ffmlp", "shencoder", "raymarching", "freqencoder")
_EXT_ROOTS = ("_gridencoder", "_ffmlp", "_shencoder", "_raymarching")
_MAYBE_MISSING = ("turtle", "trimesh", "mcubes", "tensorboardX", "lpips", "torch_ema", "torchmetrics", "imageio", "matplotlib",
                  "pysdf", "cv2", "rich", "tqdm", "pandas", "packaging")


def available(backend="ours"):
    ok = os.path.exists(os.path.join(CALLERS, "nerf", "network_ff.py"))
    if backend == "ref":
        from oracle import ref_driver
        ok = ok and os.path.exists(os.path.join(WRAPPERS, "gridencoder", "grid.py")) and ref_driver.available()
    return ok


# ---- inert stand-ins for absent third-party modules ----------------------------------------------------------------------------
class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert()


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self):
        self.roots = set()

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_finder = _StubFinder()


def _install_stubs():
    if _finder not in sys.meta_path:
        sys.meta_path.insert(0, _finder)
    for name in _MAYBE_MISSING:
        if name in _finder.roots:
            continue
        present = sys.modules.get(name)
        if present is not None:
            if getattr(present, "__spec__", None) is not None or hasattr(present, "__file__"):
                continue                     # a real module
            for k in [k for k in sys.modules if k.split(".")[0] == name]:    # somebody's bare placeholder: replace it by an inert stub
                del sys.modules[k]
            _finder.roots.add(name)
            continue
        try:
            importlib.import_module(name)
        except Exception:
            for k in [k for k in sys.modules if k.split(".")[0] == name]:
                del sys.modules[k]
            _finder.roots.add(name)


# ---- the stacks ------------------------------------------------------------------------------------------------------------------
class Stack:
    def __init__(self, backend):
        assert backend in ("ours", "ref")
        self.backend = backend
        self._mods = {}
        # names private to this stack while it is active
        self._roots = _CALLER_ROOTS + ((_PKG_ROOTS + _EXT_ROOTS) if backend == "ref" else ())
        self._paths = [WRAPPERS, CALLERS] if backend == "ref" else [PKG, CALLERS]
        if backend == "ref":
            from oracle import ref_driver
            for n in ("gridencoder", "ffmlp", "shencoder", "raymarching"):
                self._mods["_" + n] = ref_driver.mod(n)       # what `import _gridencoder as _backend` (grid.py:9-12) resolves to

    def _mine(self, name):
        return name.split(".")[0] in self._roots

    @contextlib.contextmanager
    def active(self):
        _install_stubs()
        saved = {k: v for k, v in sys.modules.items() if self._mine(k)}
        for k in saved:
            del sys.modules[k]
        sys.modules.update(self._mods)
        saved_path = list(sys.path)
        sys.path[:0] = self._paths
        try:
            yield self
        finally:
            self._mods = {k: v for k, v in sys.modules.items() if self._mine(k)}
            for k in self._mods:
                del sys.modules[k]
            sys.modules.update(saved)
            sys.path[:] = saved_path

    def module(self, name):
        """Import (inside the stack) and return a module by its reference name, e.g. 'nerf.network_ff'."""
        with self.active():
            return importlib.import_module(name)

    def file_of(self, name):
        return getattr(self.module(name), "__file__", None)


_stacks = {}


def load(backend):
    if backend not in _stacks:
        if not available(backend):
            raise RuntimeError(f"reference Python stack '{backend}' not staged: run oracle/build_ref.py where /root/reference exists")
        _stacks[backend] = Stack(backend)
    return _stacks[backend]


def make_nerf(stack, **kw):
    """NeRFNetwork of nerf/network_ff.py (cuda_ray=True) constructed inside `stack`."""
    with stack.active():
        net = importlib.import_module("nerf.network_ff")
        return net.NeRFNetwork(cuda_ray=True, **kw)
