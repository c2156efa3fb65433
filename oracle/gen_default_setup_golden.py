#!/usr/bin/env python3
"""Ensemble-level golden values of BASELINE.json configs[0] -- Code/Input/default_setup.in verbatim (GUI off), 5 jobs at E/N = 1, 5, 10, 50,
100 Td, 2e4 electrons -- from replicas of the UNMODIFIED reference (oracle/_ref/lokimc).  TEST / BASELINE INFRASTRUCTURE ONLY; runs in the
build container.  Writes tests/golden/ensemble_default_setup.json = {"jobs": {"<folder>": {"mean", "std", "reported_relstd", "replicas"}}, "wall": [...]}.
usage: python oracle/gen_default_setup_golden.py [replicas]"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_ensemble_golden as ge   # noqa: E402
import run_reference as rr  # noqa: E402

SETUP = "/root/reference/Code/Input/default_setup.in"


def setup_text():
    text = open(SETUP).read()
    return text.replace("gui: \n  isOn: true", "gui: \n  isOn: false")


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    text = setup_text()
    assert "isOn: false" in text.split("gui:")[1][:40]
    jobs, walls = {}, []
    for r in range(reps):
        res = rr.run(text, "swarm_O2_short", timeout=7200)
        walls.append(res["wall"])
        for job in res["jobs"]:
            d = {k: job["swarm"].get(k) for k in ge.KEYS}
            d.update({k + "/relstd": job["swarm"].get(k + "/relstd") for k in ge.KEYS if k + "/relstd" in job["swarm"]})
            d["real"] = job["details"]["total number of real collisions"]; d["null"] = job["details"]["total number of null collisions"]
            d["elapsed"] = job["details"]["Elapsed time"]
            jobs.setdefault(job["folder"], []).append(d)
        print("replica", r, "wall %.1f s" % res["wall"], {f: v[-1][ge.KEYS[0]] for f, v in jobs.items()}, flush=True)
    out = {}
    for folder, runs in jobs.items():
        keys = [k for k in ge.KEYS if all(x.get(k) is not None for x in runs)]
        out[folder] = dict(mean={k: float(np.mean([x[k] for x in runs])) for k in keys},
                           std={k: float(np.std([x[k] for x in runs], ddof=1)) if reps > 1 else 0.0 for k in keys},
                           reported_relstd={k: float(np.mean([x.get(k + "/relstd", 0.0) or 0.0 for x in runs])) for k in keys}, replicas=runs)
    with open(os.path.join(HERE, "..", "tests", "golden", "ensemble_default_setup.json"), "w") as f:
        json.dump(dict(setup="Code/Input/default_setup.in (gui.isOn: false)", threads=res["threads"], wall=walls, jobs=out, setup_text=text), f, indent=1)


if __name__ == "__main__":
    main()
