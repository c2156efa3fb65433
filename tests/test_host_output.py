"""Output side of the host (lokib200_report_* / lokib200_output_*): the raw state of a finished reference job (dumped by
oracle/_ref/harness `solve`, oracle/gen_output_golden.py) goes through this project's post-processing and writers; every file
must have the reference's layout byte for byte, and every number must agree with the reference's to 1e-11 relative (the two
codes add the same terms in a different order in a few reductions).  CPU only."""
import os
import re
import tarfile
import tempfile

import numpy as np
import pytest

import loki_mc_b200 as lk

HERE = os.path.dirname(os.path.abspath(__file__))
FIX_INPUT = os.path.join(HERE, "fixtures", "Input")
GOLD = os.path.join(HERE, "golden")
NUM = re.compile(r"[-+]?\d+\.\d+e[-+]\d+")


@pytest.fixture(scope="module", autouse=True)
def _library():
    if not os.path.exists(lk.lib_path()):
        lk.build()


def read_raw(path, n_jobs_hint=None):
    """layout written by oracle/harness.cpp `solve`"""
    a = np.fromfile(path, dtype=np.float64)
    pos = [0]

    def take(n=1):
        v = a[pos[0]:pos[0] + n]; pos[0] += n
        return v if n > 1 else float(v[0])
    P, nE, nC, nR, nA, nPh, nS, sym = (int(take()) for _ in range(8))
    d = {"n_electrons": take()}
    r = lk.SolveResults()
    r.averaged_mean_energy, r.averaged_mean_energy_error = take(), take()
    for name, k in [("flux_drift_velocity", 3), ("flux_drift_velocity_error", 3), ("flux_diffusion", 9), ("flux_diffusion_error", 9),
                    ("bulk_drift_velocity", 3), ("bulk_drift_velocity_error", 3), ("bulk_diffusion", 9), ("bulk_diffusion_error", 9)]:
        v = take(k)
        for i in range(k):
            getattr(r, name)[i] = v[i]
    r.power_gain_field, r.power_growth, r.power_balance_rel_error = take(), take(), take()
    r.time, r.steady_state_time, r.total_integrated_time, r.trial_collision_frequency, r.max_eedf_energy, r.elapsed_seconds = (take() for _ in range(6))
    r.total_collisions, r.null_collisions, r.collisions_at_ss, r.null_collisions_at_ss = (take() for _ in range(4))
    r.n_sampling_points, r.n_integration_points = int(take()), int(take())
    d["results"] = r
    d["evdf_max_speed"] = take()
    for k in ("rate_coeffs", "power_gain", "power_loss", "counts"):
        d[k] = take(P).copy()
    d["eeh"] = take(nE).copy()
    if sym:
        d["eah"] = take(nE * nC).copy(); d["evh"] = take(nR * nA).copy()
    if nPh:
        d["eeh_periodic"] = take(nPh * nE).copy()
    d["n_samples"] = nS
    d["times"] = take(nS).copy(); d["mean_energy"] = take(nS).copy()
    d["mean_pos"] = take(3 * nS).copy(); d["mean_vel"] = take(3 * nS).copy(); d["pos_cov"] = take(9 * nS).copy()
    if nPh:
        d["points_per_phase"] = take(nPh).copy(); d["mean_energy_periodic"] = take(nPh).copy()
        d["flux_velocity_periodic"] = take(3 * nPh).copy(); d["bulk_velocity_periodic"] = take(3 * nPh).copy()
        d["flux_diffusion_periodic"] = take(9 * nPh).copy(); d["bulk_diffusion_periodic"] = take(9 * nPh).copy()
    assert pos[0] == len(a), "raw dump layout mismatch"
    return d


def compare_file(ours, ref, rel=1e-11):
    """identical text once the numbers are masked; numbers equal to `rel`.  Returns (n_numbers, n_not_identical)."""
    a, b = open(ours).read(), open(ref).read()
    a_lines, b_lines = a.split("\n"), b.split("\n")
    assert len(a_lines) == len(b_lines), "%s: %d lines, reference has %d" % (ours, len(a_lines), len(b_lines))
    n = diff = 0
    for ln, (x, y) in enumerate(zip(a_lines, b_lines)):
        if "Elapsed time" in y:
            continue
        assert NUM.sub("#", x) == NUM.sub("#", y), "%s line %d layout:\n%r\n%r" % (os.path.basename(ours), ln + 1, x, y)
        for u, v in zip(NUM.findall(x), NUM.findall(y)):
            n += 1
            if u != v:
                diff += 1
                fu, fv = float(u), float(v)
                digits = len(v.split("e")[0].split(".")[1])
                tol = max(rel, 2.0 * 10.0 ** (-digits)) * max(abs(fv), 1e-300)
                assert abs(fu - fv) <= tol, "%s line %d: %s vs reference %s" % (os.path.basename(ours), ln + 1, u, v)
    return n, diff


@pytest.mark.parametrize("setup,folder", [("setup_out_dc", "fx_dc"), ("setup_out_ac", "fx_ac")])
def test_output_files_match_the_reference(setup, folder):
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(GOLD, "output_%s.tgz" % folder)) as t:
            t.extractall(os.path.join(tmp, "ref"), filter="data")
        ref = os.path.join(tmp, "ref", folder)
        s = lk.Setup(FIX_INPUT, "fx/%s.in" % setup)
        out = lk.Output(s, os.path.join(tmp, "ours"))
        assert out.folder == os.path.join(tmp, "ours", folder)
        for job in range(s.n_jobs):
            rep = lk.Report(s, job, data=read_raw(os.path.join(ref, "job%d.raw.bin" % job)))
            out.write(rep)
        total = differing = files = 0
        for dirpath, _, names in os.walk(ref):
            for name in names:
                if name.endswith(".raw.bin"):
                    continue
                mine = os.path.join(out.folder, os.path.relpath(os.path.join(dirpath, name), ref))
                assert os.path.exists(mine), "missing output file " + mine
                n, d = compare_file(mine, os.path.join(dirpath, name))
                total += n; differing += d; files += 1
        ours = sum(len(n) for _, _, n in os.walk(out.folder))
        assert ours == files, "extra files written"
        assert files >= 10 and total > 5000
        assert differing <= 0.02 * total, "%d of %d numbers differ in the last printed digit" % (differing, total)


def test_report_values_and_errors():
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(GOLD, "output_fx_dc.tgz")) as t:
            t.extractall(tmp, filter="data")
        s = lk.Setup(FIX_INPUT, "fx/setup_out_dc.in")
        d = read_raw(os.path.join(tmp, "fx_dc", "job1.raw.bin"))
        rep = lk.Report(s, 1, data=d)
        assert rep.swarm("meanEnergy") == d["results"].averaged_mean_energy and rep.swarm("Te") == pytest.approx(2 / 3 * rep.swarm("meanEnergy"), rel=1e-15)
        e = rep.eedf()
        step = e["energy"][1] - e["energy"][0]
        assert np.sum(e["eedf"] * np.sqrt(e["energy"])) * step == pytest.approx(1.0, rel=1e-12)      # the EEDF is normalised
        assert rep.power("field") == d["results"].power_gain_field
        parts = sum(rep.power(k) for k in ("field", "elasticNet", "inelastic", "superelastic", "eDensGrowth"))
        assert rep.power("balance") == pytest.approx(parts, rel=1e-12)
        assert sum(rep.power("vibrationalIne", g) for g in ("XY", "Z", "X", "Y")) == pytest.approx(rep.power("vibrationalIne"), rel=1e-12)
        rates = rep.rates()
        assert [r["description"] for r in rates][-1] == "e+XY(X)->e+XY(X),Effective" and len(rep.rates(extra=True)) == 2
        with pytest.raises(KeyError):
            rep.swarm("noSuchParameter")
        bad = dict(d); bad["eeh"] = d["eeh"][:-1]
        with pytest.raises(lk.LokiB200Error):
            lk.Report(s, 7, data=d)
        bad = dict(d); bad["results"] = d["results"]; bad["n_samples"] = 3
        with pytest.raises(lk.LokiB200Error, match="fewer samples"):
            lk.Report(s, 1, data=bad)
