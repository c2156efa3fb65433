"""A/B of time-to-3-sigma on default_setup.in: env settings given as 'NAME=V,NAME=V' groups, interleaved repetitions, min / median wall per group."""
import os, sys, json, subprocess, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
groups = sys.argv[2:] if len(sys.argv) > 2 else ["LOKIB200_CONCURRENT_JOBS=1", "LOKIB200_CONCURRENT_JOBS=5"]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
code = "import bench, json; r = bench.run_time_to_3sigma('default', with_reference=False); print(json.dumps(dict(value=r['value'], ok=r['within_3sigma'], worst=r['worst_deviation_sigma'])))"
res = {g: [] for g in groups}
for rep in range(reps):
    for g in groups:
        env = dict(os.environ)
        for kv in g.split(","):
            if kv and kv != "-":
                k, v = kv.split("="); env[k] = v
        out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(g, "FAILED", out.stderr[-400:], flush=True); continue
        r = json.loads(line[-1]); res[g].append(r["value"])
        print("%-70s rep %d: %.3f s  within 3 sigma %s (worst %.2f)" % (g, rep, r["value"], r["ok"], r["worst"]), flush=True)
for g in groups:
    if res[g]:
        print("%-70s min %.3f s  median %.3f s  (%d runs)" % (g, min(res[g]), statistics.median(res[g]), len(res[g])))
