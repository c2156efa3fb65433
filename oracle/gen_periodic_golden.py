#!/usr/bin/env python3
"""Phase-resolved golden values of BASELINE.json configs[3] (N2, AC electric field + DC magnetic field, gasTemperatureEffect true) from replicas of
the UNMODIFIED reference (oracle/_ref/lokimc): MCTemporalInfo_periodic.txt (mean energy and flux velocity per phase, Output.h:756-782) and the
swarm parameters.  TEST INFRASTRUCTURE ONLY; runs in the build container.  Writes tests/golden/ensemble_n2_true_acb.json.
usage: python oracle/gen_periodic_golden.py [replicas]"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg   # noqa: E402
import gen_ensemble_golden as ge   # noqa: E402
import run_reference as rr  # noqa: E402

NAME, NEL, NPTS = "n2_true_acb", 20000, 4000


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    kw = dict(gg.MODELS[NAME][0]); kw["nelec"] = NEL
    # the steady-state criterion of the reference can take arbitrarily long for AC fields at low noise: cap it
    text = gg.setup_text(**kw).replace("nIntegrationPoints: 1E3", "nIntegrationPoints: %d\n    maxCollisionsBeforeSteadyState: 2E3" % NPTS)
    text = text.replace("output:\n  isOn: false", "output:\n  isOn: true\n  folder: ens_%s\n  dataFiles:\n    - swarmParameters\n    - MCSimDetails\n    - MCTemporalInfo_periodic" % NAME)
    runs, per = [], []
    for r in range(reps):
        res = rr.run(text, "ens_" + NAME, keep=True, timeout=7200)
        job = res["jobs"][0]
        d = {k: job["swarm"].get(k) for k in ge.KEYS}
        d.update({k + "/relstd": job["swarm"].get(k + "/relstd") for k in ge.KEYS if k + "/relstd" in job["swarm"]})
        d["real"] = job["details"]["total number of real collisions"]; d["null"] = job["details"]["total number of null collisions"]
        d["elapsed"] = job["details"]["Elapsed time"]
        runs.append(d)
        outdir = os.path.join(rr.REFDIR, "Output", "ens_" + NAME)
        rows = [[float(x) for x in ln.split()] for ln in open(os.path.join(outdir, "MCTemporalInfo_periodic.txt")).read().split("\n")[1:] if ln.strip()]
        per.append(np.array(rows)[:, :7])    # phase, phase time, E/N, mean energy, flux v_x, v_y, v_z
        shutil.rmtree(outdir, ignore_errors=True)
        print(NAME, r, "mean energy %.5f, wall %.1f s" % (d[ge.KEYS[0]], res["wall"]), flush=True)
    per = np.array(per)
    keys = [k for k in ge.KEYS if all(x.get(k) is not None for x in runs)]
    out = dict(model=NAME, n_electrons=NEL, n_integration_points=NPTS, threads=res["threads"], replicas=runs, setup_text=text,
               mean={k: float(np.mean([x[k] for x in runs])) for k in keys}, std={k: float(np.std([x[k] for x in runs], ddof=1)) for k in keys},
               reported_relstd={k: float(np.mean([x.get(k + "/relstd", 0.0) or 0.0 for x in runs])) for k in keys},
               periodic=dict(columns=["phase", "phase_time", "E/N", "mean_energy", "flux_vx", "flux_vy", "flux_vz"], mean=per.mean(axis=0).tolist(),
                             std=per.std(axis=0, ddof=1).tolist()))
    with open(os.path.join(gg.GOLD, "ensemble_%s.json" % NAME), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
