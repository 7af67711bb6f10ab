"""Build libprb.so in-tree (``python -m pyrates_b200.build``).

The shared library is host C++ (NVRTC + lazily loaded CUDA driver API); the CUDA kernels themselves are
generated per model and JIT-compiled for sm_100a through it.  nvcc is used as the compiler driver so the
same command also builds any ahead-of-time ``.cu`` translation units placed in csrc/.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
OUT = os.path.join(HERE, "libprb.so")


def build(verbose: bool = True, force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cpp")) + glob.glob(os.path.join(HERE, "csrc", "*.cu")))
    deps = srcs + [os.path.join(os.path.dirname(HERE), "include", "prb.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.path.join(CUDA, "bin", "nvcc")
    cmd = [nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-I", os.path.join(CUDA, "include"), *srcs,
           "-L", os.path.join(CUDA, "lib64"), "-lnvrtc", "-ldl",
           "-Xlinker", f"-rpath={os.path.join(CUDA, 'lib64')}", "-cudart", "none", "-o", OUT]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
