"""Build the CUDA shared library in-tree (zkvm-brainfuck_b200/libbfgpu.so) for sm_100a."""
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("BFGPU_SO") or os.path.join(_DIR, "libbfgpu.so")  # BFGPU_SO: experiment builds (tools/)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    csrc = os.path.join(_DIR, "csrc")
    return [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [os.path.join(_DIR, "..", "include", "bfgpu.h")]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not (force or needs_build()):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("BFGPU_NVCC_EXTRA", "").split()  # experiment switches, e.g. -DNTT2_MONT_TW=1
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(_DIR, "csrc", "bfgpu.cu")]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
