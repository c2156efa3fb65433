#!/usr/bin/env python3
"""Time bench.py once per library variant (build/variants/lib_*.so) on the GPU box and print one line per variant."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))
for lib in libs:
    env = dict(os.environ, LOKIB200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-extras", "--steps", "30", "--relax", "30"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        d = json.loads(r.stdout.strip().split("\n")[-1])
        print("%-28s kernel %.3f ms  step %.3f ms  e2e %.3e ev/s  mean_e %.4f real %.4f" % (os.path.basename(lib), d["roofline"]["kernel_ms"], d["ms_per_step"],
                                                                                       d["e2e"]["value"], d["config"]["mean_energy_eV"], d["config"]["real_fraction"]), flush=True)
    except Exception:
        print(os.path.basename(lib), "FAILED", r.stdout[-500:], flush=True)
