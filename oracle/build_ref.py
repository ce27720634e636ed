"""Build recipe for oracle/_ref: the UNMODIFIED reference CUDA extensions.

TEST INFRASTRUCTURE ONLY.  Compiles the reference's own sources *where they lie*
under /root/reference (gridencoder/src, shencoder/src, raymarching/src, ffmlp/src +
its vendored CUTLASS 2.8 headers) for sm_100 into oracle/_ref/<name>/_ref_<name>.so.
The only flag changed w.r.t. the reference's */backend.py is -std=c++14 -> -std=c++17
(current torch headers require C++17).  No reference source is copied into this repo;
oracle/_ref/ is git-ignored and only holds build outputs.

The resulting modules expose the reference's pybind tables
(gridencoder/src/bindings.cpp:6-8, ffmlp/src/bindings.cpp:6-10,
shencoder/src/bindings.cpp:6-7, raymarching/src/bindings.cpp:7-18) and are used
  * by tests/ (-m gpu) as the authoritative parity oracle on the GPU box,
  * by tests/golden/make_golden.py to generate committed golden vectors,
  * by bench.py as the "reference CUDA build" timing arm (reported beside ours).

ship_python() additionally stages the reference's own PYTHON on the path — the callers (encoding.py, activation.py, nerf/, sdf/)
under oracle/_ref/py/callers and the four wrapper packages under oracle/_ref/py/wrappers — byte for byte, so that the GPU box (which
has no /root/reference) can run the unmodified nerf/network_ff.py + nerf/renderer.py over either this repo's drop-in packages or the
reference's wrappers + extensions (oracle/ref_stack.py).  Like the .so files these copies are git-ignored build outputs of this
recipe, travel with the gpurun snapshot, and are never imported by anything under torch-ngp_b200/.

Usage:  python oracle/build_ref.py [gridencoder shencoder raymarching ffmlp]
"""
import os
import sys

REF = os.environ.get("NGP_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "gridencoder": (["gridencoder.cu", "bindings.cpp"], []),
    "shencoder": (["shencoder.cu", "bindings.cpp"], []),
    "raymarching": (["raymarching.cu", "bindings.cpp"], []),
    "ffmlp": (["ffmlp.cu", "bindings.cpp"],
              ["dependencies/cutlass/include", "dependencies/cutlass/tools/util/include"]),
}


def so_path(name):
    return os.path.join(OUT, name, f"_ref_{name}.so")


def build(names=None, verbose=False):
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree {REF} not present (it only exists in the build container)")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    names = names or list(EXTS)
    for name in names:
        srcs, incs = EXTS[name]
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        if os.path.exists(so_path(name)):
            continue
        load(
            name=f"_ref_{name}",
            sources=[os.path.join(REF, name, "src", s) for s in srcs],
            extra_include_paths=[os.path.join(REF, name, i) for i in incs],
            extra_cflags=["-O3", "-std=c++17"],
            extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__",
                               "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
                               "-lineinfo"],
            build_directory=bdir,
            verbose=verbose,
            is_python_module=False,
        )
        assert os.path.exists(so_path(name)), so_path(name)
    return [so_path(n) for n in names]


PY_CALLERS = ["encoding.py", "activation.py", "loss.py", "nerf/network_ff.py", "nerf/network.py", "nerf/renderer.py", "nerf/utils.py",
              "sdf/netowrk_ff.py", "sdf/netowrk.py"]
PY_WRAPPERS = ["gridencoder/__init__.py", "gridencoder/grid.py", "ffmlp/__init__.py", "ffmlp/ffmlp.py",
               "shencoder/__init__.py", "shencoder/sphere_harmonics.py", "raymarching/__init__.py", "raymarching/raymarching.py"]


def ship_python():
    """Copy the reference's Python files of the path (unmodified) into oracle/_ref/py/{callers,wrappers}; returns the file list."""
    import shutil
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree {REF} not present (it only exists in the build container)")
    done = []
    for sub, files in (("callers", PY_CALLERS), ("wrappers", PY_WRAPPERS)):
        for rel in files:
            src = os.path.join(REF, rel)
            if not os.path.exists(src):
                continue
            dst = os.path.join(OUT, "py", sub, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            done.append(dst)
    return done


if __name__ == "__main__":
    print(build(sys.argv[1:] or None, verbose=True))
    print(ship_python())
