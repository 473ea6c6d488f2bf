"""INTEGRATION.md route B, built for real: the reference's OWN C++ wrappers (network/models/pointnet_lib/src/
pointnet2_api.cpp, ball_query.cpp, group_points.cpp, interpolate.cpp, sampling.cpp) compiled against
include/captra_ops.h and linked with captra_b200/libcaptra_ops.so instead of the reference's *_gpu.cu kernels.

TEST INFRASTRUCTURE (called by oracle/Makefile `routeb`; output oracle/_ref/routeb/pointnet2_cuda.so, git-ignored).
The only edits, applied to temporary copies that are deleted after the compile, are the ones INTEGRATION.md lists:
  *.cpp     `#include <THC/THC.h>` -> `<ATen/cuda/CUDAContext.h>`, `extern THCState *state;` dropped (THC is gone
            from current torch; the variable is unused, ball_query.cpp:8)
  *_gpu.h   the `void ..._kernel_launcher...(...)` declarations replaced by `#include "captra_ops.h"`
            (same names and argument order, extern "C", int status)
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference/network/models/pointnet_lib/src")
OUT = os.path.join(HERE, "_ref", "routeb")


def main():
    if not os.path.isdir(REF):
        print("reference sources not present (%s); keeping oracle/_ref/routeb if any" % REF)
        return 0
    lib = os.path.join(ROOT, "captra_b200", "libcaptra_ops.so")
    if not os.path.exists(lib):
        print("libcaptra_ops.so not built yet; skipping route B")
        return 0
    from torch.utils import cpp_extension as ce
    import torch
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="routeb_")
    try:
        for f in glob.glob(os.path.join(REF, "*.cpp")) + glob.glob(os.path.join(REF, "*.h")):
            text = open(f).read()
            if f.endswith(".cpp"):
                text = text.replace("#include <THC/THC.h>", "#include <ATen/cuda/CUDAContext.h>")
                text = re.sub(r"^extern THCState \*state;\s*$", "", text, flags=re.M)
            elif f.endswith("_gpu.h"):
                text, n = re.subn(r"^void\s+\w*kernel_launcher\w*\s*\([^;]*\);", "", text, flags=re.M)
                assert n >= 1, f
                text, n = re.subn(r"^(#define\s+\w+_H\w*\s*)$", r'\1\n#include "captra_ops.h"', text, count=1, flags=re.M)
                assert n == 1, f
            open(os.path.join(tmp, os.path.basename(f)), "w").write(text)
        srcs = sorted(glob.glob(os.path.join(tmp, "*.cpp")))
        try:
            inc = ce.include_paths(device_type="cuda")
        except TypeError:
            inc = ce.include_paths(cuda=True)
        inc += [sysconfig.get_paths()["include"], os.path.join(ROOT, "include"), tmp, "/usr/local/cuda/include"]
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        out = os.path.join(OUT, "pointnet2_cuda.so")
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w", "-DTORCH_EXTENSION_NAME=pointnet2_cuda",
               "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
        cmd += ["-I" + p for p in inc] + srcs
        cmd += ["-L" + os.path.dirname(lib), "-lcaptra_ops", "-Wl,-rpath," + os.path.dirname(lib),
                "-Wl,-rpath,$ORIGIN/../../../captra_b200",
                "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-Wl,-rpath," + tlib,
                "-L/usr/local/cuda/lib64", "-lcudart", "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
            return 1
        print("built", os.path.relpath(out, ROOT))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
