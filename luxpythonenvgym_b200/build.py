"""Builds the CUDA extension in-tree: luxpythonenvgym_b200/_lux_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lux_b200.so")
SOURCES = ["lux_step.cu", "lux_obs.cu", "lux_env.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lux_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


MAPGEN_OUT = os.path.join(HERE, "_lux_mapgen.so")


def build_mapgen(force=False):
    """Host-only native map generator (g++; no CUDA)."""
    src = os.path.join(CSRC, "lux_mapgen.cpp")
    hdr = os.path.join(HERE, "..", "include", "lux_b200.h")
    if not force and os.path.exists(MAPGEN_OUT) and os.path.getmtime(MAPGEN_OUT) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return MAPGEN_OUT
    cmd = ["g++", "-O3", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", MAPGEN_OUT, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed for lux_mapgen.cpp")
    return MAPGEN_OUT


def build(force=False, verbose=False):
    build_mapgen(force)
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "_build_log.txt"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(OUT)
