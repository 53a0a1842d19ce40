"""Build the CUDA extension in-tree: cuvite_b200/libcuvite_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the
gpurun snapshot.  `python -m cuvite_b200.build` rebuilds when sources changed.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcuvite_b200.so")
STAMP = OUT + ".stamp"
SOURCES = ["engine.cu", "vite_file.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest() -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return OUT
    if not os.path.exists(NVCC):
        if os.path.exists(OUT):
            if not (os.path.exists(STAMP) and open(STAMP).read() == dig):
                sys.stderr.write("cuvite_b200.build: WARNING: no nvcc here and libcuvite_b200.so does not match the "
                                 "current sources (stale binary in use)\n")
            return OUT  # GPU box without nvcc: use the shipped binary
        raise RuntimeError("nvcc not found and no prebuilt libcuvite_b200.so")
    cmd = [NVCC, *FLAGS, *[os.path.join(CSRC, s) for s in SOURCES], "-o", OUT]
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's $CXX lacks libgomp.spec; let nvcc find /usr/bin/g++
    env.pop("CC", None)
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
