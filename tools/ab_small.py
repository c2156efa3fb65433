#!/usr/bin/env python3
"""Blocking advance loop (what the job driver does) of the one-electron-per-thread kernel at small ensemble sizes, for A/B runs of builds
(LOKIB200_LIB selects the library)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as gio
import loki_mc_b200 as lk

def run(n, model="o2_sdcs", steps=1500):
    g = gio.load(model)
    eng = lk.Engine(g, n, seed=3)
    mx = eng.init_ensemble(100.0)
    eng.build_tables(2 * mx)
    nu = eng.check_nu_trial(mx, eng.table_info()["nu_max_last"], horizon=11.0)
    t = 0.0
    for _ in range(200):
        nu = eng.check_nu_trial(mx, nu, horizon=11.0); t += 1 / nu; r = eng.advance(nu, t, True); mx = max(r[34], r[35])
    eng.kernel_time_ms()
    t0 = time.perf_counter(); ev = 0
    for _ in range(steps):
        nu = eng.check_nu_trial(mx, nu, horizon=11.0); t += 1 / nu; r = eng.advance(nu, t, True); mx = max(r[34], r[35]); ev += r[0] + r[1]
    dt = time.perf_counter() - t0
    ms, k = eng.kernel_time_ms()
    eng.close()
    return dt / steps * 1e6, ev / dt, ms * 1e3

for model in ("o2_sdcs", "n2_aniso"):
    for n in (20000, 100000):
        us, rate, k1 = run(n, model)
        print("%s %s n=%d: %.1f us per interval (K1 %.1f us), %.3g events/s" % (os.path.basename(lk.lib_path()), model, n, us, k1, rate), flush=True)
