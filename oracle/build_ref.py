"""ORACLE build recipe (test infrastructure): compile the reference's OWN CUDA sources, unmodified and
in place under /root/reference, for sm_100a into ``oracle/_ref/ref_pn2_ext*.so`` (git-ignored; it
travels to the GPU box with the snapshot).  Nothing is copied into the repo.

Sources: inference/grasp_proposal/network_models/models/pointnet2_utils/csrc/
         {sampling,ball_query,grouping,interpolate}_kernel.cu + main.cpp
Flags mirror the reference's setup.py (nvcc -O2, default -fmad=true) plus the arch.

The module only RUNS on a GPU (tests/test_ref_cuda_parity.py, tests/test_reference_dropin_gpu.py, bench.py reference_cuda).
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/inference/grasp_proposal/network_models/models/pointnet2_utils/csrc"
OUT = os.path.join(HERE, "_ref")
NAME = "ref_pn2_ext"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(force=False, verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/_ref: reference sources not present; using the prebuilt file if any")
        return so_path() if os.path.exists(so_path()) else None
    if os.path.exists(so_path()) and not force:
        return so_path()
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(HERE, "ref_shim"), "-I" + SRC]
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    defs = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    common = ["-O2", "-std=c++17", "-include", "THC/THC.h"] + inc + defs
    nvcc = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
            "-Xcompiler", "-fPIC", "-w"] + common
    units = ["sampling_kernel.cu", "ball_query_kernel.cu", "grouping_kernel.cu", "interpolate_kernel.cu", "main.cpp"]

    def compile_one(u):
        obj = os.path.join(OUT, u.rsplit(".", 1)[0] + ".o")
        if u.endswith(".cu"):
            cmd = nvcc + ["-c", os.path.join(SRC, u), "-o", obj]
        else:
            cmd = ["/usr/bin/g++", "-fPIC", "-w"] + common + ["-c", os.path.join(SRC, u), "-o", obj]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(5) as ex:
        objs = list(ex.map(compile_one, units))
    libdir = ce.library_paths()[0]
    link = ["/usr/bin/g++", "-shared", "-o", so_path()] + objs + [
        "-L" + libdir, "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
        "-ltorch_python", "-lcudart", "-Wl,-rpath," + libdir]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    if verbose:
        print("oracle/_ref: built", so_path())
    return so_path()


def load():
    """Import the compiled reference module (GPU box or here; calling it needs a GPU)."""
    import importlib.util
    import torch  # noqa: F401  (must be loaded first)
    p = so_path()
    if not os.path.exists(p):
        raise ImportError("oracle/_ref/ref_pn2_ext.so missing: run `python oracle/build_ref.py` in the build container")
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)
