"""In-tree build of the CUDA library (sm_100a only) and of the test infrastructure (oracle port, CPU warp emulator).

    python -m ribbit_b200.build          # everything
The built .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libribbit_scan.so")
EMU = os.path.join(ROOT, "tests", "emu", "libemu.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]
CUDA_SOURCES = ["kernels.cu", "motif_kernels.cu", "fasta_kernels.cu", "api.cu"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(ROOT, "include", "ribbit_scan.h"))
    return d


def build_cuda(force=False, verbose=False):
    if not force and not _newer(LIB, _deps()):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in CUDA_SOURCES]
    subprocess.run(cmd, check=True)
    return LIB


def build_emulator(force=False):
    src = os.path.join(ROOT, "tests", "emu", "emu_scan.cpp")
    if not force and not _newer(EMU, _deps() + [src]):
        return EMU
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + CSRC, src, "-o", EMU], check=True)
    return EMU


MERGE = os.path.join(ROOT, "tests", "emu", "libseedmerge.so")


def build_merge(force=False):
    """The host seed-list merges (ribbit_b200/host/seed_merge.cpp) as a library of their own, for the tests."""
    src = os.path.join(HERE, "host", "seed_merge.cpp")
    if not force and not _newer(MERGE, _deps() + [src, os.path.join(HERE, "host", "seed_merge.h")]):
        return MERGE
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", src, "-o", MERGE], check=True)
    return MERGE


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)


def build_host():
    """ribbit_b200/bin/ribbit_gpu: the reference's host code (unmodified, from /root/reference) + the GPU processSequence.
    Needs the reference sources, so it is only (re)built in the build container; the binary travels with the snapshot."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host")], check=True)


def build_all(verbose=False):
    build_cuda(verbose=verbose)
    build_emulator()
    build_merge()
    build_oracle()
    build_host()


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv)
