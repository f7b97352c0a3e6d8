#!/usr/bin/env python
"""Builds experimental variants of libhc_b200.so (extra -D flags) into haploconduct_b200/lib/variants/
for A/B timing on the GPU box: HC_B200_LIB=<variant.so> python bench.py ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haploconduct_b200 import build as B  # noqa: E402

VARIANTS = {"base": [], "noprefetch": ["-DHC_NO_PREFETCH"], "noswz": ["-DHC_NO_SWZ"]}
VARIANTS.update({k: v for k, v in (a.split("=", 1) for a in sys.argv[1:] if "=" in a)} and
                {a.split("=", 1)[0]: a.split("=", 1)[1].split(",") for a in sys.argv[1:] if "=" in a})
out = os.path.join(B.LIBDIR, "variants")
os.makedirs(out, exist_ok=True)
for name, flags in VARIANTS.items():
    objs = []
    for src in B.CPP_SOURCES:
        o = os.path.join(out, name + "_" + src + ".o")
        subprocess.check_call([B.HOST_CXX, "-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", *flags, "-c", os.path.join(B.CSRC, src), "-o", o])
        objs.append(o)
    for src in B.CU_SOURCES:
        o = os.path.join(out, name + "_" + src + ".o")
        subprocess.check_call([B.NVCC, "-ccbin", B.HOST_CXX, "-O3", "-std=c++14", "-lineinfo", *B.ARCH, "-Xcompiler", "-fPIC,-fopenmp,-O2",
                               *flags, "-c", os.path.join(B.CSRC, src), "-o", o])
        objs.append(o)
    so = os.path.join(out, "libhc_b200_%s.so" % name)
    subprocess.check_call([B.NVCC, "-ccbin", B.HOST_CXX, "-shared", *B.ARCH, "-Xcompiler", "-fPIC,-fopenmp", "-o", so, *objs, "-lgomp"])
    for o in objs:
        os.remove(o)
    print(so)
