"""Builds libgstim.so (host C++ + sm_100a CUDA) in-tree with nvcc.

    python -m stim_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU, so this also serves as the CPU-side build check.
The shared library lands next to this file (stim_b200/libgstim.so); it is git-ignored but is shipped to
the GPU box by gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgstim.so")
SOURCES = ["circuit.cc", "lowering.cc", "response.cc", "writers.cc", "tableau_ref.cc", "interp.cu", "kernels.cu", "sparse.cu", "api.cu", "dem.cu", "m2d.cu", "flipsim.cu"]
HEADERS = ["circuit.h", "lowering.h", "response.h", "sparse.cuh", "hostpipe.h", "writers.h", "tableau_ref.h", "kernels.cuh", "program.h", "log2_table.h", "log2_q26_table.h", "../../include/gstim.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "_obj", src + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if src.endswith(".cu") and verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {src} failed ---\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("libgstim build failed")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
