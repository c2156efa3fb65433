#!/usr/bin/env python3
"""k_histogram (lokib200_sample_histograms) on a relaxed 1e7-electron N2 ensemble: the launch ncu captures for profiles/r2_hist_* (and CUDA-event timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import golden_io as gio
import loki_mc_b200 as lk
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
g = gio.load("n2_aniso")
eng = lk.Engine(g, n, seed=3)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
mx = eng.init_ensemble(2.41 / (1.5 * 1.38064852e-23 * 300 / 1.6021766208e-19))
eng.build_tables(2 * mx)
nu = eng.check_nu_trial(mx, eng.table_info()["nu_max_last"], horizon=11.0)
t = 0.0
for _ in range(40):
    nu = eng.check_nu_trial(mx, nu, horizon=11.0); t += 1 / nu; r = eng.advance(nu, t, True); mx = max(r[34], r[35])
eng.set_histogram_grid(1.2 * mx)
for _ in range(3):
    eng.sample_histograms(-1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    eng.sample_histograms(-1)
e1.record(); torch.cuda.synchronize()
print("k_histogram: %.1f us per launch for %d electrons (%.0f GB/s of 24 B per electron)" % (e0.elapsed_time(e1) / 20 * 1e3, n, n * 24 / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9))
eng.close()
