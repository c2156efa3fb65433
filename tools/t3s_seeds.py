import os, sys, json
sys.path.insert(0, os.getcwd())
import bench
for fast in (False, True):
    for seed in (1, 2, 3, 4, 5):
        os.environ["LOKIB200_SEED"] = str(seed)
        with bench.StdoutToStderr():
            d = bench.run_time_to_3sigma("default", with_reference=False, fast_mode=fast)
        print("fast" if fast else "ref ", seed, "%.2f s" % d["value"], "%.2f sigma" % d["worst_deviation_sigma"], d["worst_parameter"], flush=True)
