"""In-tree build of libcaptra_ops.so (sm_100a only; nvcc cross-compiles without a GPU).

    python -m captra_b200.build [--force] [--verbose]

Every .cu under captra_b200/csrc is compiled to an object in captra_b200/_build/ (in parallel)
and linked into captra_b200/libcaptra_ops.so, which travels to the GPU box with the snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# A/B experiments: CAPTRA_EXTRA_NVCC_FLAGS="-DX" CAPTRA_LIB_OUT=captra_b200/libcaptra_ops_x.so python -m captra_b200.build
# builds a variant beside the product library (select it at run time with CAPTRA_LIB_PATH).
_VARIANT = os.environ.get("CAPTRA_LIB_OUT")
OBJ = os.path.join(PKG, "_build" if not _VARIANT else "_build_" + os.path.basename(_VARIANT).replace(".so", ""))
LIB = os.path.abspath(_VARIANT) if _VARIANT else os.path.join(PKG, "libcaptra_ops.so")

NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
] + os.environ.get("CAPTRA_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_stamp():
    h = hashlib.sha1()
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "captra_ops.h"))
    for p in hdrs:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, obj, verbose):
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = _deps_stamp()
    stamp_file = os.path.join(OBJ, "stamp")
    old = open(stamp_file).read() if os.path.exists(stamp_file) else ""
    if old != stamp:
        force = True
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, log in ex.map(lambda a: _compile(a[0], a[1], verbose), jobs):
                if verbose or rc != 0:
                    sys.stderr.write(log)
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s" % src)
    need_link = bool(jobs) or not os.path.exists(LIB) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
