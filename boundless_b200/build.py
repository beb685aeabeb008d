"""Build libb200zkp.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

    python -m boundless_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libb200zkp.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SOURCES = ["ntt.cu", "hash.cu", "stark.cu", "halops.cu", "verify.cu", "compat.cu", "tables.cpp", "prover.cpp", "capi.cpp", "planner.cpp"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-ccbin", HOSTCXX,
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I", CSRC, "-I", os.path.join(HERE, "..", "include")]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", h) for h in ("b200zkp.h", "b200_risc0_sys_compat.h")]


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    inc = os.path.join(CSRC, "constants.inc")
    if not os.path.exists(inc):
        subprocess.check_call([sys.executable, os.path.join(HERE, "tools", "gen_constants.py"), inc])
    deps = _deps()
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        if force or _stale(obj, deps):
            cmd = [NVCC] + FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr
    with ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(run, jobs))
    if verbose:
        for o in outs:
            print(o)
    objs = [os.path.join(OBJ, s.rsplit(".", 1)[0] + ".o") for s in SOURCES]
    if force or jobs or _stale(SO, objs):
        run([NVCC, "-shared", "-ccbin", HOSTCXX, "-o", SO] + objs + ["-cudart", "static", "-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
