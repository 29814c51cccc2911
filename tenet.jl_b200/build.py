"""Build libtnb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["csrc/context.cu", "csrc/api.cu", "csrc/comm.cu", "csrc/multi.cu", "csrc/linalg.cu", "csrc/kernels_generic.cu", "csrc/kernels_c64_tc.cu", "csrc/kernels_c128_dmma.cu", "csrc/kernels_stem.cu", "csrc/planner.cpp"]
OUT = os.path.join(HERE, "libtnb200.so")
STAMP = OUT + ".stamp"      # digest of the sources the library was built from (travels with the .so)
LOCK = OUT + ".lock"
FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17", "--cudart", "static"]


def _deps():
    csrc = os.path.join(HERE, "csrc")
    deps = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cpp", ".h", ".inc")))
    deps.append(os.path.join(HERE, "..", "include", "tnb200.h"))
    return [d for d in deps if os.path.exists(d)]


def source_digest():
    """sha256 over flags + every source/header: the staleness test (mtimes do not survive a repo snapshot)."""
    h = hashlib.sha256(" ".join(FLAGS + SRC).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    # one builder at a time (bench.py runs one process per GPU, all calling build()); the others wait on the lock
    # and then find the stamp up to date.  The library is written under a temporary name and renamed atomically.
    with open(LOCK, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return OUT
            digest = source_digest()
            tmp = OUT + f".tmp{os.getpid()}"
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SRC + ["-ldl"]
            r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libtnb200.so")
            if verbose:
                sys.stderr.write(r.stderr)
            os.replace(tmp, OUT)
            with open(STAMP + ".tmp", "w") as f:
                f.write(digest + "\n")
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
