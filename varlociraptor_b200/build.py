"""Builds the CUDA engine in-tree: varlociraptor_b200/csrc/libvlr_engine.so (sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvlr_engine.so")
SOURCES = ["vlr_engine.cu"]
DEPS = ["vlr_engine.cu", "engine_core.cuh", "engine_wave.cuh", "engine_resident.cuh", "engine_sets.cuh", "engine_types.cuh", "contamination.cuh", "scenario_prep.h", os.path.join("..", "..", "include", "vlr_engine.h")]
# -fmad=false: the reference (Rust) and the oracle never contract a*b+c; grid abscissae (linspace, observable
# limits) must round identically or adaptive grids diverge. The hot loop uses explicit fma() where fusion is wanted.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "-cudart", "shared"]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        extra = os.environ.get("VLR_NVCC_EXTRA", "").split()  # tuning experiments (-D...), never set in the product build
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
