"""In-tree build of libb200t1.so (the C-ABI shared library of include/b200_t1.h) with nvcc for sm_100a.

The `.so` is git-ignored but lives in the tree (booster_gym_b200/libb200t1.so) so that it travels to the GPU box with
the repo snapshot.  There is no JIT and no fallback: if the library is missing on a machine without nvcc the package
raises at import of `_lib`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200t1.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
          "-diag-suppress", "550,177"] + os.environ.get("B200_NVCC_EXTRA", "").split()

# translation unit -> extra flags
UNITS = {
    "common.cu": [],
    "physics_kernels.cu": [],                   # FP32-pipe bound: keep FMA contraction
    "env_kernels.cu": ["--fmad=false"],         # rounds like the reference's separate torch ops (bit-exact masks)
    "learner_kernels.cu": [],
}


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: libb200t1.so cannot be built on this machine")
    return nv


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "include", "b200_t1.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


STAMP = LIB + ".srchash"


def source_hash():
    """content hash of everything the library is built from (+ the flags): file times do not survive a repo snapshot, contents do"""
    import hashlib

    h = hashlib.sha256()
    for path in sorted(_deps()):
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def up_to_date():
    """True if libb200t1.so exists and was built from exactly the sources in the tree"""
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == source_hash()
    except OSError:
        return False


def build_locked(verbose=False):
    """build() under an inter-process file lock (torchrun ranks of a fresh checkout all arrive here at once); the library is
    linked into a temporary file and renamed into place, so nobody can dlopen a half-written file"""
    import fcntl

    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    with open(os.path.join(ROOT, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if up_to_date():
                return LIB
            return build(force=False, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def build(force=False, verbose=False, ptxas_info=False):
    """Compile every CUDA translation unit for sm_100a and link libb200t1.so. Returns the library path."""
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]
    deps = _deps()
    if not force and not _stale(LIB, deps) and up_to_date():
        return LIB
    nv = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    procs = []
    for u in units:
        src = os.path.join(CSRC, u)
        # (the flags are part of the object's name: an object built with other flags, e.g. a -D measurement switch passed through
        # B200_NVCC_EXTRA, is never linked into a library whose stamp claims the current flags)
        import hashlib

        tag = hashlib.sha256(" ".join(ARCH + COMMON + UNITS[u]).encode()).hexdigest()[:10]
        obj = os.path.join(OBJ_DIR, u.replace(".cu", "." + tag + ".o"))
        objs.append(obj)
        if not force and not _stale(obj, deps):
            continue
        cmd = [nv] + ARCH + COMMON + UNITS[u] + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((u, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for u, p in procs:
        out, _ = p.communicate()
        if ptxas_info or verbose or p.returncode != 0:
            sys.stdout.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {u}")
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nv] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", tmp] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    with open(STAMP + ".tmp", "w") as f:
        f.write(source_hash())
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
