"""Input side of the host (include/lokib200_host.h): setup files, LXCat files and property databases -> flattened process set and
engine configuration, against what the UNMODIFIED reference builds from the same files (goldens: oracle/gen_input_golden.py,
oracle/gen_golden.py).  CPU only."""
import json
import os
import tempfile

import numpy as np
import pytest

import loki_mc_b200 as lk

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
FIX_INPUT = os.path.join(HERE, "fixtures", "Input")
REF_INPUT = os.path.join(ROOT, "oracle", "_ref", "Input")

PROCESS_KEYS = ["p_type", "p_elastic", "p_superelastic", "p_angular", "p_ap0", "p_ap1", "p_swf", "p_emin", "p_emax", "p_reldens", "p_mass", "p_redmass",
                "p_eloss", "p_thstd", "p_w", "gas_first", "gas_last", "gas_fraction", "xs_offset", "xs_energy", "xs_value"]
GOLDEN_MODELS = ["reid_dc", "reid_ac", "reid_b", "reid_ecr", "reid_acb", "reid_true_aniso", "o2_sdcs", "n2_aniso", "n2_true_acb", "arhe", "arhe_true", "air",
                 "ls_f05", "ls_att_aniso"]


@pytest.fixture(scope="module", autouse=True)
def _library():
    if not os.path.exists(lk.lib_path()):
        lk.build()


def _scalars(g):
    return dict(zip([str(x) for x in g["scalar_names"]], g["scalar_values"]))


def _check_processes(setup, g):
    d = setup.processes()
    for k in PROCESS_KEYS:
        a, b = np.asarray(d[k]), g[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a, b), "%s differs from the reference (max rel %.3g)" % (k, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    assert d["descriptions"] == [str(x) for x in g["descriptions"]]
    return d


def _check_config(setup, g, job=0):
    sc = _scalars(g)
    cfg = setup.config(job)
    assert cfg.gas_density == sc["totalGasDensity"] and cfg.gas_temperature == sc["gasTemperature"]
    assert cfg.gas_temperature_effect == int(sc["gasTemperatureEffect"]) and cfg.ionization_sharing == int(sc["energySharingIonizType"])
    assert cfg.energy_sharing_factor == sc["energySharingFactor"]
    assert cfg.is_cylindrically_symmetric == int(sc["isCylindricallySymmetric"])
    assert [cfg.electric_field[i] for i in range(3)] == list(g["electricField"])
    assert cfg.excitation_omega == sc["excitationFrequencyRadians"] and cfg.cyclotron_omega == sc["cyclotronFrequency"]
    assert cfg.n_interp_points == int(sc["nInterpPoints"])
    assert setup.controls().energy_max_elastic == sc["energyMaxElastic"]
    assert setup.controls().initial_temp_ratio == sc["initialElecTempOverGasTemp"]
    return cfg


@pytest.mark.parametrize("name", ["setup_a", "setup_b"])
def test_fixture_setups_flatten_like_the_reference(name):
    g = np.load(os.path.join(GOLD, "input_%s.npz" % name))
    s = lk.Setup(FIX_INPUT, "fx/%s.in" % name)
    d = _check_processes(s, g)
    cfg = _check_config(s, g)
    # numericsMC keys as the BoltzmannMC constructor stores them (harness 'controls')
    (nip, nss, tabs, errs, over, e_me, e_fd, e_fD, e_bd, e_bD, e_pb, cmin, cmax, cafter, sync, tratio, nE, nC, nR, nA, nPh, nI, nEl) = g["controls"]
    c = s.controls()
    assert (c.n_integration_points, c.n_integrated_ss_times, c.integrated_absolute_time) == (nip, nss, tabs)
    assert (c.errors_to_be_checked, c.sync_over_sampling) == (int(errs), int(over))
    assert (c.rel_err_mean_energy, c.rel_err_flux_drift, c.rel_err_flux_diff, c.rel_err_bulk_drift, c.rel_err_bulk_diff, c.rel_err_power_balance) == \
           (e_me, e_fd, e_fD, e_bd, e_bD, e_pb)
    assert (c.min_collisions_before_ss, c.max_collisions_before_ss, c.max_collisions_after_ss) == (cmin, cmax, cafter)
    assert (c.sync_factor, c.initial_temp_ratio) == (sync, tratio)
    assert (cfg.n_energy_cells, cfg.n_cos_cells, cfg.n_radial_cells, cfg.n_axial_cells, cfg.n_phases, cfg.n_interp_points, cfg.n_electrons) == \
           (int(nE), int(nC), int(nR), int(nA), int(nPh), int(nI), int(nEl))
    assert d["p_elastic"].sum() >= 2


def test_fixture_job_sweeps():
    a = lk.Setup(FIX_INPUT, "fx/setup_a.in")
    assert a.n_jobs == 5 and a.variable_condition == "reducedElecField"
    want = [float.fromhex(x) for x in json.load(open(os.path.join(GOLD, "expressions.json")))["vector"]["logspace(0,2,5)"]]
    assert [a.job_value(j) for j in range(5)] == want
    e0, e4 = a.config(0).electric_field[2], a.config(4).electric_field[2]
    assert e0 < 0 and e4 == pytest.approx(100 * e0, rel=1e-15)
    b = lk.Setup(FIX_INPUT, "fx/setup_b.in")
    assert b.n_jobs == 3 and b.variable_condition == "reducedMagField"
    assert [b.job_value(j) for j in range(3)] == [100.0, 200.0, 400.0]
    assert b.config(2).cyclotron_omega == pytest.approx(4 * b.config(0).cyclotron_omega, rel=1e-15)
    assert b.value("electronKinetics.numericsMC.relError.powerBalance") == "1E-3" and b.value("no.such.key") == ""
    assert "ionizationOperatorType: oneTakesAll" in b.dump()


@pytest.mark.parametrize("name", GOLDEN_MODELS)
def test_reference_input_files_flatten_like_the_reference(name):
    """The reference's own LXCat files / Databases (present only where oracle/_ref was built from /root/reference)."""
    if not os.path.isdir(os.path.join(REF_INPUT, "Databases")):
        pytest.skip("oracle/_ref/Input not built here")
    g = np.load(os.path.join(GOLD, name + ".npz"))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, name + ".in")
        with open(path, "w") as f:
            f.write(str(g["setup_text"]))
        if name.startswith("ls_"):   # the Lucas-Saelee LXCat files are generated by oracle/gen_golden.py into Input/_gen
            if not os.path.exists(os.path.join(REF_INPUT, "_gen", name + "_LXCat.txt")):
                pytest.skip("generated LXCat file missing")
        s = lk.Setup(REF_INPUT, path)
        _check_processes(s, g)
        _check_config(s, g)


def test_expressions_match_the_reference_parser():
    gold = json.load(open(os.path.join(GOLD, "expressions.json")))
    for text, want in gold["scalar"].items():
        assert lk.eval_expression(text) == float.fromhex(want), text
    for text, want in gold["vector"].items():
        assert list(lk.eval_vector_expression(text)) == [float.fromhex(x) for x in want], text


@pytest.mark.parametrize("text", ["5!", "foo(2)", "(1+2", "1+", "2*x", "", "1:2:3:4"])
def test_malformed_expressions_are_errors(text):
    with pytest.raises(lk.LokiB200Error):
        lk.eval_vector_expression(text) if ":" in text else lk.eval_expression(text)


def _write(tmp, name, text):
    with open(os.path.join(tmp, name), "w") as f:
        f.write(text)
    return os.path.join(tmp, name)


def test_setup_errors_are_reported_not_fatal():
    base = open(os.path.join(FIX_INPUT, "fx", "setup_a.in")).read()
    with tempfile.TemporaryDirectory() as tmp:
        with pytest.raises(lk.LokiB200Error, match="could not be opened"):
            lk.Setup(FIX_INPUT, os.path.join(tmp, "missing.in"))
        with pytest.raises(lk.LokiB200Error, match="Gas fractions are not properly normalized"):
            lk.Setup(FIX_INPUT, _write(tmp, "a.in", base.replace("- Z = 1-0.75", "- Z = 0.3")))
        with pytest.raises(lk.LokiB200Error, match=r"Electronic/ionic distribution XY\(\*\) is not properly normalized"):
            lk.Setup(FIX_INPUT, _write(tmp, "b.in", base.replace("- XY(A3) = 0.02", "- XY(A3) = 0.03")))
        with pytest.raises(lk.LokiB200Error, match="Mass of gas Z not found"):
            lk.Setup(FIX_INPUT, _write(tmp, "c.in", base.replace("mass: fx/masses.txt", "mass:\n      - XY = 1e-26")))
        with pytest.raises(lk.LokiB200Error, match="boltzmannMC"):
            lk.Setup(FIX_INPUT, _write(tmp, "d.in", base.replace("eedfType: boltzmannMC", "eedfType: boltzmann")))
        with pytest.raises(lk.LokiB200Error, match="Could not parse line"):
            lk.Setup(FIX_INPUT, _write(tmp, "e.in", base.replace("  gasTemperature: 350", "  gasTemperature 350 K now")))
        with pytest.raises(lk.LokiB200Error, match="nIntegrationPoints''.\nValue should be a single integer >= 500"):
            lk.Setup(FIX_INPUT, _write(tmp, "g.in", base.replace("nIntegrationPoints: 1E3", "nIntegrationPoints: 100")))
        with pytest.raises(lk.LokiB200Error, match="''gasTemperatureEffect'' field not found in the ''electronKinetics>numericsMC'' section"):
            lk.Setup(FIX_INPUT, _write(tmp, "h.in", base.replace("    gasTemperatureEffect: smartActivation\n", "")))
        with pytest.raises(lk.LokiB200Error, match="ionizationOperatorType''.\nValue should be either"):
            lk.Setup(FIX_INPUT, _write(tmp, "i.in", base.replace("ionizationOperatorType: usingSDCS", "ionizationOperatorType: sharing")))
        # a state that becomes a target through a '<->' collision needs its own Elastic: gas Z has no Effective to derive it from
        zfile = open(os.path.join(FIX_INPUT, "fx", "Z_LXCat.txt")).read().replace("e + Z(1S0) -> e + Z(3P2)", "e + Z(1S0) <-> e + Z(3P2)")
        os.makedirs(os.path.join(tmp, "fx2"))
        _write(tmp, "fx2/Z_LXCat.txt", zfile)
        text = base.replace("- fx/Z_LXCat.txt", "- %s" % os.path.relpath(os.path.join(tmp, "fx2", "Z_LXCat.txt"), FIX_INPUT)) \
                   .replace("- Z(1S0) = 1\n    population", "- Z(1S0) = 1\n      - Z(3P2) = 5\n    population") \
                   .replace("      - Z(1S0) = 1\n  numericsMC", "      - Z(1S0) = 0.5\n      - Z(3P2) = 0.5\n  numericsMC")
        with pytest.raises(lk.LokiB200Error, match="does not have an ''Effective'' collision defined"):
            lk.Setup(FIX_INPUT, _write(tmp, "f.in", text))
