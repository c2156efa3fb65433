"""Pins the CPU oracle (oracle/lokioracle.c) against golden vectors produced by the UNMODIFIED reference
(oracle/harness.cpp driving liblokiref.so built from /root/reference; generator: oracle/gen_golden.py).

Tolerances: tables are pure (+,*,/) arithmetic in the same order -> bit-exact.  Per-electron events go through libm
(sin/cos/log/pow/tan/atan) in both; the oracle is compiled by gcc with -ffp-contract=off like the reference, so the
agreement is expected far below the 1e-12 relative parity bar of BASELINE.json:north_star.
"""
import numpy as np
import pytest

import golden_io as gio
from oracle import lokioracle as lo

RTOL = 1e-12


def vec_close(a, b, rtol=RTOL, floor=0.0):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    scale = max(np.linalg.norm(b), floor)
    return np.linalg.norm(a - b) <= rtol * scale


@pytest.fixture(scope="module", params=gio.MODELS)
def gm(request):
    g = gio.load(request.param)
    return g, lo.Model(g)


def test_philox_known_answers():
    # Random123 v1.14 kat_vectors, philox4x32 10 rounds
    assert lo.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert lo.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert lo.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [lo.stream_uniform(7, 3, 1, j) for j in range(64)]
    assert all(0.0 < x < 1.0 for x in u) and len(set(u)) == 64


def test_tables_bit_exact(gm):
    g, m = gm
    t = m.build_tables(float(g["tab_maxE"]))
    assert t["nE"] == int(g["tab_nE"])
    assert t["dE"] == float(g["tab_dE"])
    rows = g["tab_rows"]
    assert np.array_equal(t["sigma"][rows], g["tab_sigma_rows"])
    assert np.array_equal(t["cum"][rows], g["tab_cum_rows"])
    assert np.array_equal(t["nu_tot"], g["tab_nu_tot"])
    assert np.array_equal(t["nu_max"], g["tab_nu_max"])
    chk = np.array([t["cum"].sum(), (t["cum"] * t["cum"]).sum(), t["sigma"].sum()])
    assert np.allclose(chk, g["tab_cum_checksum"], rtol=1e-13, atol=0)
    assert not np.isnan(t["cum"]).any()


def test_max_accel_energy(gm):
    g, m = gm
    for (e0, dt), want in zip(g["maxaccel_in"], g["maxaccel_out"]):
        assert abs(m.max_accel_energy(e0, dt) - want) <= 1e-14 * abs(want)


def test_events_match_reference(gm):
    g, m = gm
    m.build_tables(float(g["tab_maxE"]))
    I, O = gio.EV_IN, gio.EV_OUT
    n_real = 0
    for ein, eout in zip(g["ev_in"], g["ev_out"]):
        v = ein[I["v"]]
        st = np.concatenate([ein[I["r"]], v, [gio.energy_eV(v), ein[I["t_e"]], ein[I["t_cf"]], ein[I["nu_e"]]]])
        chosen, st2, out, used = m.event(ein[I["nu_trial"]], ein[I["t_sync"]], st, ein[I["draws"]])
        assert chosen == int(eout[O["chosen"]])
        assert used == int(eout[O["draws_used"]])
        assert vec_close(st2[0:3], eout[O["r"]])
        assert vec_close(st2[3:6], eout[O["v"]])
        assert abs(st2[6] - eout[O["eps"]]) <= RTOL * abs(eout[O["eps"]])
        assert abs(st2[7] - eout[O["t_e"]]) <= RTOL * abs(eout[O["t_e"]]) + 1e-30
        assert abs(st2[8] - eout[O["t_cf"]]) <= RTOL * abs(eout[O["t_cf"]])
        assert st2[9] == eout[O["nu_e"]]
        eps_scale = max(abs(eout[O["eps"]]), gio.energy_eV(v))
        assert abs(out[2] - eout[O["gain_field"]]) <= RTOL * eps_scale
        if chosen >= 0:
            n_real += 1
            assert abs(out[0] - eout[O["dE"]]) <= RTOL * eps_scale
            assert abs(out[1] - eout[O["dE_rel"]]) <= 1e-11
            if g["p_type"][chosen] == 1:
                assert vec_close(out[3:6], eout[O["ej_r"]])
                assert vec_close(out[6:9], eout[O["ej_v"]])
                assert abs(out[9] - eout[O["ej_eps"]]) <= RTOL * abs(eout[O["ej_eps"]])
    assert n_real > 100


def test_moments_and_histograms(gm):
    g, m = gm
    ens = g["ens"]
    mom = lo.moments(ens)
    want = g["moments"]
    assert abs(mom[0] - want[0]) <= 1e-13 * want[0]
    assert abs(mom[1] - want[1]) <= 1e-13 * want[1]
    assert np.allclose(mom[2:8], want[2:8], rtol=1e-12, atol=0)
    # covariances are differences of O(1e-6) terms: compare on the scale of the raw second moments
    assert np.allclose(mom[8:17], want[8:17], rtol=0, atol=1e-12 * np.abs(want[8:17]).max())
    assert np.allclose(mom[17:26], want[17:26], rtol=0, atol=1e-12 * np.abs(want[17:26]).max())
    hdr = g["hist_hdr"]
    ne, nc, nr, na = (int(x) for x in hdr[:4])
    cyl = bool(g["cond"]["is_cylindrically_symmetric"])
    eeh, eah, evh = lo.histograms(ens[3:6], float(hdr[4]), ne, nc, nr, na, cyl)
    assert np.array_equal(eeh, g["eeh"])
    assert eeh.sum() <= ens.shape[1]
    if cyl:
        want_eah = np.zeros((ne, nc)); want_eah[tuple(g["eah_idx"].T)] = g["eah_val"]
        want_evh = np.zeros((nr, na)); want_evh[tuple(g["evh_idx"].T)] = g["evh_val"]
        assert np.array_equal(eah, want_eah)
        assert np.array_equal(evh, want_evh)
