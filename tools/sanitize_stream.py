#!/usr/bin/env python3
"""Drive the streaming advance kernel (forced with LOKIB200_KERNEL=stream) over the golden models for a few intervals, small enough to sit
under `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_stream.py` on a GPU box.  racecheck reports read/write hazards
between the load and store lines of the flight phase: lanes without an electron read slot 0, whose owner may be writing it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["LOKIB200_KERNEL"] = "stream"
import numpy as np  # noqa: E402
import golden_io as gio  # noqa: E402
import test_gpu_parity as T  # noqa: E402

models = sys.argv[1:] or (["n2_aniso", "air", "arhe_true", "reid_ac", "reid_b", "reid_ecr", "reid_acb", "ls_att_aniso", "reid_true_aniso", "o2_sdcs", "n2_true_acb"]
                          + gio.FIELD_GT_MODELS)
for name in models:
    g = gio.load(name)
    n = 160_077
    rng = np.random.default_rng(5)
    hot = name in ("n2_aniso", "air", "arhe_true", "ls_att_aniso", "o2_sdcs", "n2_true_acb")
    s0 = T._start_state(g, n, rng, 1e-2, 40.0 if hot else 5.0)
    eng = T._engine(g, n, seed=77, first_electron_id=3)
    eng.build_tables(100.0 if hot else 12.0)
    nu = eng.table_info()["nu_max_last"]
    eng.set_ensemble(s0, 0.0)
    for it in range(1, 3):
        r = eng.advance(nu, it / nu, sample=True)
    print(name, "ok: real %d null %d born %d attached %d" % tuple(int(x) for x in r[:4]), flush=True)
    eng.close()
