"""K1 time of the two kernel forms (one electron per thread / streaming pool) against the ensemble size: where should the engine switch?
Device-resident intervals of the N2 bench workload and of the O2 process set, K1 timed by the engine's CUDA events."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for model in ("n2_aniso", "o2_sdcs"):
    for n in (5e4, 1e5, 2.5e5, 5e5, 1e6):
        row = []
        for form in ("thread", "stream"):
            env = dict(os.environ, LOKIB200_KERNEL=form)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--model", model, "--electrons", str(n), "--no-extras", "--no-cpu-baseline", "--steps", "100", "--relax", "40"],
                                 env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
            j = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
            row.append((j["roofline"]["kernel_ms"], j["ms_per_step"], j["value"], j["e2e"]["value"]))
        print("%-9s n=%8d  thread: K1 %7.1f us step %7.1f us %.3g ev/s (e2e %.3g) | stream: K1 %7.1f us step %7.1f us %.3g ev/s (e2e %.3g)" %
              (model, int(n), 1e3 * row[0][0], 1e3 * row[0][1], row[0][2], row[0][3], 1e3 * row[1][0], 1e3 * row[1][1], row[1][2], row[1][3]), flush=True)
