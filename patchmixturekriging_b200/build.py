"""Build libpmk_b200.so (sm_100a only) in-tree with nvcc.  `python -m patchmixturekriging_b200.build`."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PMK_VARIANT=name builds an experiment (different PMK_NVCC_EXTRA) next to the product as libpmk_b200_<name>.so;
# load it with PMK_LIB=<path> (see _lib.py).  The product library is always libpmk_b200.so.
VARIANT = os.environ.get("PMK_VARIANT", "")
LIB = os.path.join(HERE, f"libpmk_b200_{VARIANT}.so" if VARIANT else "libpmk_b200.so")
BUILD_DIR = os.path.join(HERE, f"build_{VARIANT}" if VARIANT else "build")
SOURCES = ["pmk_api.cu", "pmk_fit.cu", "pmk_tree.cu", "pmk_query.cu", "pmk_query_d1.cu", "pmk_query_d2.cu", "pmk_query_d3.cu", "pmk_gram.cu", "pmk_invert.cu", "pmk_partition.cu", "pmk_multi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("PMK_NVCC_EXTRA", "").split()     # e.g. PMK_NVCC_EXTRA=-DPMK_PROFILE_CYCLES for the phase counters
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-Xptxas", "-v"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "pmk.h"))
    hdrs.append(os.path.abspath(__file__))
    objs, jobs = [], []
    os.makedirs(BUILD_DIR, exist_ok=True)
    only = os.environ.get("PMK_VARIANT_ONLY", "").split()   # variant builds: recompile just these, link the rest from build/
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        if VARIANT and only and src not in only:
            objs.append(os.path.join(HERE, "build", src.replace(".cu", ".o")))
            continue
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC, *FLAGS, *EXTRA, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            log = os.path.join(BUILD_DIR, os.path.basename(cmd[-1]) + ".log")
            with open(log, "w") as f:
                f.write(r.stdout + r.stderr)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed: {' '.join(cmd)}")
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-pthread", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
