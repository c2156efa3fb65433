#!/usr/bin/env python3
"""Build experimental variants of liblokib200.so (DC field, smartActivation only: -DLK_BENCH_ONLY) for A/B timing on the GPU box.
usage: python tools/build_variants.py name1:"-DFLAG ..." name2:"..."   -> build/variants/lib_<name>.so ; select with LOKIB200_LIB"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loki_mc_b200._capi import HOST_SOURCES, NVCC_FLAGS  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)


def build(spec):
    name, _, flags = spec.partition(":")
    flags = flags.split()
    base = [f for f in NVCC_FLAGS]
    if "-fmad=true" in flags:
        base = [f for f in base if f != "-fmad=false"]
    cmd = ["/usr/local/cuda/bin/nvcc"] + base + ["-DLK_BENCH_ONLY"] + flags + ["-o", os.path.join(OUT, "lib_%s.so" % name),
                                                                                os.path.join(ROOT, "loki_mc_b200", "csrc", "lokib200.cu")] + \
          [os.path.join(ROOT, "loki_mc_b200", "host", f) for f in HOST_SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return name, r.returncode, r.stdout[-400:]


with ThreadPoolExecutor(max_workers=6) as ex:
    for name, rc, out in ex.map(build, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED\n" + out)
