#!/usr/bin/env python3
"""Ensemble-level golden values from the UNMODIFIED reference (oracle/_ref/lokimc): swarm parameters, collision statistics and
their run-to-run scatter for the setups of tests/golden/<model>.npz.  TEST INFRASTRUCTURE ONLY; runs in the build container.

Writes tests/golden/ensemble_<model>.json = {"replicas": [...], "mean": {...}, "std": {...}, "reported_relstd": {...}}.
sigma_eff for the 3-sigma ensemble parity tests = max(reported Rel. std, replica scatter)  (SURVEY.md section 8(c)).
usage: python oracle/gen_ensemble_golden.py [model ...]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg   # noqa: E402
import run_reference as rr  # noqa: E402

CASES = {  # model: (nElectrons, nIntegrationPoints, nIntegratedSSTimes, replicas)
    "reid_dc": (20000, 3000, 2, 4),
    "reid_acb": (20000, 3000, 2, 3),
    "reid_true_aniso": (20000, 3000, 2, 3),
    "n2_aniso": (20000, 3000, 2, 4),
    "o2_sdcs": (20000, 3000, 2, 3),
    "arhe": (20000, 3000, 2, 3),
    "air": (20000, 3000, 2, 3),
    "ls_f05": (20000, 3000, 2, 4),
    "ls_att_aniso": (20000, 3000, 2, 3),
}
KEYS = ["Energy parameters/Mean energy", "Flux parameters/v_z", "Flux parameters/v_x", "Bulk parameters/v_z", "Bulk parameters/v_x",
        "Flux parameters/Reduced transverse diffusion coefficient", "Flux parameters/Reduced longitudinal diffusion coefficient",
        "Bulk parameters/Reduced transverse diffusion coefficient", "Bulk parameters/Reduced longitudinal diffusion coefficient",
        "Parameters obtained from the EEDF/Ionization coefficient", "Parameters obtained from the EEDF/Attachment coefficient"]


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        nel, npts, nss, reps = CASES[name]
        kw = dict(gg.MODELS[name][0]); kw["nelec"] = nel
        text = gg.setup_text(**kw).replace("nIntegrationPoints: 1E3", "nIntegrationPoints: %d\n    nIntegratedSSTimes: %g" % (npts, nss))
        text = text.replace("output:\n  isOn: false", "output:\n  isOn: true\n  folder: ens_%s\n  dataFiles:\n    - swarmParameters\n    - rateCoefficients\n    - MCSimDetails" % name)
        runs = []
        for r in range(reps):
            res = rr.run(text, "ens_" + name)
            job = res["jobs"][0]
            d = {k: job["swarm"].get(k) for k in KEYS}
            d.update({k + "/relstd": job["swarm"].get(k + "/relstd") for k in KEYS if k + "/relstd" in job["swarm"]})
            d["real"] = job["details"]["total number of real collisions"]; d["null"] = job["details"]["total number of null collisions"]
            d["elapsed"] = job["details"]["Elapsed time"]; d["steady_state_time"] = job["details"]["steady-state time"]
            d["final_time"] = job["details"]["final simulation time"]
            runs.append(d)
            print(name, r, {k.split("/")[-1]: d[k] for k in KEYS[:4]}, "events/s %.3g" % ((d["real"] + d["null"]) / d["elapsed"]), flush=True)
        mean = {k: float(np.mean([x[k] for x in runs])) for k in KEYS}
        std = {k: float(np.std([x[k] for x in runs], ddof=1)) for k in KEYS}
        rep = {k: float(np.mean([x.get(k + "/relstd", 0.0) or 0.0 for x in runs])) for k in KEYS}
        with open(os.path.join(gg.GOLD, "ensemble_%s.json" % name), "w") as f:
            json.dump(dict(model=name, n_electrons=nel, n_integration_points=npts, n_integrated_ss_times=nss, threads=res["threads"],
                           replicas=runs, mean=mean, std=std, reported_relstd=rep, setup_text=text), f, indent=1)


if __name__ == "__main__":
    main()
