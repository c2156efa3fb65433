"""End to end on the GPU: setup file in, output folder out (lokib200_run_setup, the reference executable's main loop), compared with
the folder the reference wrote for the same setup (tests/golden/output_*.tgz): same files, same layout, swarm results within the
statistical errors the two runs report."""
import json
import os
import re
import subprocess
import sys
import tarfile
import tempfile

import pytest

import loki_mc_b200 as lk
from test_host_output import NUM

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from run_reference import parse_swarm as rr_parse  # noqa: E402  (a parser of the reference's file format; no reference code runs)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX_INPUT = os.path.join(HERE, "fixtures", "Input")
GOLD = os.path.join(HERE, "golden")


@pytest.mark.timeout(300)
@pytest.mark.parametrize("setup,folder,n_electrons", [("setup_out_dc", "fx_dc", 20000), ("setup_out_ac", "fx_ac", 20000)])
def test_run_setup_writes_the_reference_folder(setup, folder, n_electrons):
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(GOLD, "output_%s.tgz" % folder)) as t:
            t.extractall(os.path.join(tmp, "ref"), filter="data")
        ref = os.path.join(tmp, "ref", folder)
        text = open(os.path.join(FIX_INPUT, "fx", setup + ".in")).read().replace("nElectrons: 400", "nElectrons: %d" % n_electrons)
        path = os.path.join(tmp, "job.in")
        with open(path, "w") as f:
            f.write(text)
        summary = lk.run_setup(FIX_INPUT, path, os.path.join(tmp, "out"), verbose=False)
        out = os.path.join(tmp, "out", folder)
        n_files = 0
        for dirpath, _, names in os.walk(ref):
            for name in names:
                if name.endswith(".raw.bin"):
                    continue
                mine = os.path.join(out, os.path.relpath(os.path.join(dirpath, name), ref))
                assert os.path.exists(mine), "missing " + mine
                n_files += 1
                if name in ("MCTemporalInfo.txt", "setup.txt"):          # row count depends on the run; setup.txt differs in nElectrons
                    assert open(mine).readline() == open(os.path.join(dirpath, name)).readline()
                    continue
                def mask(text):   # numbers -> '#'; a sign eats one space of a left-justified column, so runs of spaces are collapsed
                    return [re.sub(r" +", " ", NUM.sub("#", x)) for x in text.split("\n")]
                a, b = mask(open(mine).read()), mask(open(os.path.join(dirpath, name)).read())
                assert len(a) == len(b), name
                for x, y in zip(a, b):
                    if "Elapsed" in y or "number of integration points" in y:
                        continue
                    assert re.sub(r"[-+]?(nan|inf)", "#", x) == re.sub(r"[-+]?(nan|inf)", "#", y), "%s:\n%r\n%r" % (name, x, y)
        assert n_files >= 10 and summary.n_jobs == (2 if folder == "fx_dc" else 1)
        gold = json.load(open(os.path.join(GOLD, "ensemble_fixture.json")))
        subs = [d for d in sorted(os.listdir(ref)) if os.path.isdir(os.path.join(ref, d))] or [""]
        for sub in subs:
            mine = rr_parse(os.path.join(out, sub, "swarmParameters.txt"))
            g = gold["jobs"]["%s/%s" % (setup, sub or ".")]
            checked = 0
            for key, mean in g["mean"].items():
                if key.endswith("v_x") or key.endswith("v_z") and not key.endswith("v_z'") and folder == "fx_ac" or mean == 0:
                    continue   # components that vanish by symmetry (pure noise); the AC run is compared in the frame of the field (v_z')
                # sigma: the reference's reported relative std, its replica scatter (4-6 replicas, hence 4 sigma) and this run's reported std; floor 0.4 %
                rel = max(g["reported_relstd"].get(key, 0.0), g["std"][key] / abs(mean), mine.get(key + "/relstd", 0.0), 4e-3)
                assert abs(mine[key] - mean) <= 4 * rel * abs(mean), "%s %s: %g vs reference %g (4 sigma = %.2f%%)" % (sub, key, mine[key], mean, 400 * rel)
                checked += 1
            assert checked >= 6
            # internal consistency of our own files
            pb = [l for l in open(os.path.join(out, sub, "powerBalance.txt")) if "Relative Power Balance" in l][0]
            assert float(NUM.findall(pb)[0]) < 2.0, pb


@pytest.mark.timeout(600)
def test_command_line_front_end():
    exe = os.path.join(os.path.dirname(lk.lib_path()), "lokimc_b200")
    if not os.path.exists(exe):
        lk.build()
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(FIX_INPUT, os.path.join(tmp, "Input"))
        text = open(os.path.join(FIX_INPUT, "fx", "setup_out_dc.in")).read().replace("[20,80]", "40").replace("nElectrons: 400", "nElectrons: 5000")
        with open(os.path.join(tmp, "cli.in"), "w") as f:
            f.write(text)
        r = subprocess.run([exe, os.path.join(tmp, "cli.in"), "1"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout
        assert "Finished!" in r.stdout and "Elapsed time is" in r.stdout
        assert sorted(os.listdir(os.path.join(tmp, "Output", "fx_dc"))) == sorted(
            ["MCSimDetails.txt", "MCTemporalInfo.txt", "eedf.txt", "evdf.txt", "powerBalance.txt", "rateCoefficients.txt", "rateCoefficientsMC.txt", "setup.txt",
             "swarmParameters.txt"])
        # an invalid setup ends with the reference's message, a non-zero status and errorLog.txt
        with open(os.path.join(tmp, "bad.in"), "w") as f:
            f.write(text.replace("nIntegrationPoints: 500", "nIntegrationPoints: 10"))
        r = subprocess.run([exe, os.path.join(tmp, "bad.in")], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        assert r.returncode != 0 and "Program stopped due to the following error" in r.stdout
        assert "Value should be a single integer >= 500" in open(os.path.join(tmp, "errorLog.txt")).read()
        r = subprocess.run([sys.executable, "-m", "loki_mc_b200", os.path.join(tmp, "bad.in")], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           timeout=300, env=dict(os.environ, PYTHONPATH=os.path.dirname(HERE)))
        assert r.returncode == 1 and "single integer >= 500" in r.stdout


@pytest.mark.timeout(300)
def test_jobs_side_by_side_write_what_the_sequential_loop_writes(monkeypatch):
    """host/run.cpp solves the jobs of a small-ensemble sweep side by side (one engine, stream and host thread each) and writes their reports in
    job order: same files, same row order in the look-up tables, same numbers within the run-to-run scatter as the sequential loop
    (LOKIB200_CONCURRENT_JOBS=1); bit-identical files where no electron is born or lost (no atomically ordered lists)."""
    text = open(os.path.join(FIX_INPUT, "fx", "setup_out_dc.in")).read().replace("[20,80]", "[20,40,60,80]").replace("nElectrons: 400", "nElectrons: 5000")
    trees = []
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "job.in")
        with open(path, "w") as f:
            f.write(text)
        for mode, k in (("seq", "1"), ("side", "4")):
            monkeypatch.setenv("LOKIB200_CONCURRENT_JOBS", k)
            summary = lk.run_setup(FIX_INPUT, path, os.path.join(tmp, mode), verbose=False)
            assert summary.n_jobs == 4
            files = {}
            for dirpath, _, names in os.walk(os.path.join(tmp, mode)):
                for name in names:
                    files[os.path.relpath(os.path.join(dirpath, name), os.path.join(tmp, mode))] = open(os.path.join(dirpath, name), errors="replace").read()
            trees.append(files)
    a, b = trees
    assert sorted(a) == sorted(b) and len(a) >= 4 * 9
    for name in a:
        la, lb = a[name].split("\n"), b[name].split("\n")
        assert len(la) == len(lb) or "MCTemporalInfo" in name, name
        if "lookUpTable" in name:       # one row per job, in job order: the first column is the swept reduced field
            col = lambda lines: [float(NUM.findall(x)[0]) for x in lines if NUM.findall(x) and not re.search(r"[A-Za-z]{3}", x)]
            assert col(la) == col(lb) == sorted(col(la)) and len(col(la)) == 4, name
        if name.endswith("swarmParameters.txt"):     # (the fixture attaches electrons: population control makes the two runs two realisations)
            checked = 0
            for x, y in zip(la, lb):
                if "Reduced mobility coefficient" in x or "Mean energy" in x:
                    p, q = float(NUM.findall(x)[0]), float(NUM.findall(y)[0])
                    assert abs(p - q) <= 0.15 * abs(q), (name, x, y)
                    checked += 1
            assert checked >= 2, name
