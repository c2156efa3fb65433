"""How much does a B200 gain from running the independent jobs of a parameter sweep side by side?  k copies of the configs[0] job (O2, 2e4
electrons, reference cadence: one blocking advance per sampling interval), each on its own engine, stream and host thread, against the same
k jobs one after the other.  ctypes releases the GIL during the calls, so Python threads are real host threads here."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_io as gio
import loki_mc_b200 as lk

def one(g, n, seed, points, out, i):
    eng = lk.Engine(g, n, seed=seed)
    job = lk.Job([eng], n_integration_points=points, n_integrated_ss_times=0.0)
    t0 = time.perf_counter()
    r = job.solve()
    out[i] = (time.perf_counter() - t0, r["n_sync_points"], r["total_collisions"] + r["null_collisions"], r["averaged_mean_energy"])
    job.close(); eng.close()

def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "o2_sdcs"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000
    points = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
    g = gio.load(model)
    out = {}
    one(g, n, 1, 500, out, 0)       # warm-up: context, module load
    ks = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else (1, 2, 3, 5, 8, 10)
    for k in ks:
        out = {}
        t0 = time.perf_counter()
        for i in range(k):
            one(g, n, 100 + i, points, out, i)
        seq = time.perf_counter() - t0
        iv_seq = sum(v[1] for v in out.values())
        out = {}
        th = [threading.Thread(target=one, args=(g, n, 100 + i, points, out, i)) for i in range(k)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        par = time.perf_counter() - t0
        iv = sum(v[1] for v in out.values())
        print("%s n=%d k=%2d jobs: one after the other %.3f s (%.1f us per interval), side by side %.3f s (%.1f us per interval of the set, %.1f us per job-interval): x%.2f"
              % (model, n, k, seq, 1e6 * seq / iv_seq, par, 1e6 * par / iv, 1e6 * par / (iv / k), seq / par), flush=True)

main()
