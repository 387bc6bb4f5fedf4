"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmaple_b200.so")
SOURCES = ["maple_b200.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "maple_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # no FMA contraction: a*b+c rounds twice, like the CPython reference
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f)))


PTXAS_LOG = os.path.join(HERE, "ptxas.log")


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compiles the library; the ptxas resource report (registers, spills per kernel) is kept next to it in ptxas.log --
    the search kernel's register allocation is touchy and tests/test_build_resources.py watches it."""
    if force or _stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(PTXAS_LOG, "w") as f:
            f.write(r.stdout)
        if verbose or r.returncode != 0:
            print(r.stdout)
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, cmd)
    return LIB


def kernel_resources(log_path: str = PTXAS_LOG) -> dict:
    """{mangled kernel name: {"registers": r, "spill_stores": b, "spill_loads": b}} from the last build's ptxas report."""
    import re
    out, cur = {}, None
    if not os.path.isfile(log_path):
        return out
    for line in open(log_path):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = out.setdefault(m.group(1), {})
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "spill_stores" not in cur:
            cur["spill_stores"], cur["spill_loads"] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["registers"] = int(m.group(1))
            cur = None
    return out


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
