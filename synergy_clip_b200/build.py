"""Build the sclip shared library in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

    python -m synergy_clip_b200.build            # incremental
    python -m synergy_clip_b200.build --force

The product is ``synergy_clip_b200/lib/libsclip.so`` exporting the C ABI of ``include/sclip.h``.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsclip.so")
SOURCES = ("sclip_api.cu", "sclip_tc.cu", "sclip_simt.cu")
HEADERS = ("common.cuh", "ptx.cuh", os.path.join("..", "..", "include", "sclip.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sclip library cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(nvcc: str, src: str, obj: str) -> str:
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return res.stderr


def _up_to_date(stamp: str, digest: str) -> bool:
    try:
        with open(stamp) as f:
            return os.path.exists(LIB) and f.read().strip() == digest
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Incremental build.  Safe when several ranks (mp.spawn / torchrun) import the package at once on a fresh checkout:
    one process holds an exclusive file lock while it compiles into a private temporary directory, the library and its
    digest stamp appear by atomic rename, and the others re-check the stamp once they get the lock."""
    import fcntl
    import tempfile

    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libsclip.digest")
    digest = _digest()
    if not force and _up_to_date(stamp, digest):
        return LIB
    with open(os.path.join(LIBDIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(stamp, digest):  # another process built it while this one waited
                return LIB
            nvcc = _nvcc()
            with tempfile.TemporaryDirectory(dir=LIBDIR, prefix=".build-") as tmp:
                objs = [os.path.join(tmp, s.replace(".cu", ".o")) for s in SOURCES]
                with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
                    logs = list(pool.map(lambda so: _compile(nvcc, *so), zip(SOURCES, objs)))
                if verbose:
                    print("\n".join(logs))
                tmp_lib = os.path.join(tmp, "libsclip.so")
                link = [nvcc, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC"]
                res = subprocess.run(link, capture_output=True, text=True)
                if res.returncode != 0:
                    raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
                tmp_log = os.path.join(tmp, "ptxas.log")
                with open(tmp_log, "w") as f:
                    f.write("\n".join(logs))
                tmp_stamp = os.path.join(tmp, "libsclip.digest")
                with open(tmp_stamp, "w") as f:
                    f.write(digest)
                if os.path.exists(stamp):
                    os.remove(stamp)  # never a new library under an old stamp or the reverse
                os.replace(tmp_log, os.path.join(LIBDIR, "ptxas.log"))
                os.replace(tmp_lib, LIB)
                os.replace(tmp_stamp, stamp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
