"""Multi-GPU exchange of the path: the shards of one job combine their per-interval result vectors with one grouped NCCL
all-reduce inside liblokib200.so (include/lokib200.h, "multi-GPU").  Shards are keyed by global electron id, so a sharded ensemble IS the
single-engine ensemble: integer-valued entries of the combined vector must be exact, sums equal up to the order of addition.
The two-device cases are skipped on a one-GPU box (the driver's scaling run and `gpurun --gpus 2` exercise them)."""
import numpy as np
import pytest

import golden_io as gio
import test_gpu_parity as T

pytestmark = pytest.mark.gpu


def _n_devices():
    import loki_mc_b200 as lk
    return lk.lib().lokib200_device_count()


def _run(engines, nu, intervals, combine):
    out = []
    for it in range(1, intervals + 1):
        for e in engines:
            e.advance_device(nu, it / nu, True, None)
        out.append(combine(engines))
    return out


def test_single_rank_communicator_is_identity():
    """a communicator of one rank: the blocking advance goes through the all-reduce path and must return what the plain engine returns"""
    import loki_mc_b200 as lk
    g = gio.load("reid_acb")
    n = 50_000
    s0 = T._start_state(g, n, np.random.default_rng(3), 1e-2, 5.0)
    res = []
    for with_comm in (False, True):
        eng = T._engine(g, n, seed=11)
        if with_comm:
            eng.comm_init_rank(lk.comm_unique_id(), 0, 1)
            assert eng.comm_size() == 1
        eng.build_tables(12.0)
        nu = eng.table_info()["nu_max_last"]
        eng.set_ensemble(s0, 0.0)
        res.append([eng.advance(nu, it / nu, sample=True) for it in range(1, 4)])
        eng.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("name", ["reid_acb", "air"])
def test_two_engines_share_a_communicator(name, p2p, monkeypatch):
    """both transports of the exchange: the peer-memory kernel (k_exchange) and the grouped NCCL all-reduce (LOKIB200_P2P=0)"""
    monkeypatch.setenv("LOKIB200_P2P", p2p)
    import loki_mc_b200 as lk
    R = lk.R
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    g = gio.load(name)
    n = 400_000
    hot = name == "air"
    s0 = T._start_state(g, n, np.random.default_rng(5), 1e-2, 40.0 if hot else 5.0)
    maxE = 100.0 if hot else 12.0
    # reference: the whole ensemble on one engine
    one = T._engine(g, n, seed=77)
    one.build_tables(maxE); nu = one.table_info()["nu_max_last"]; one.set_ensemble(s0, 0.0)
    want = [one.advance(nu, it / nu, sample=True) for it in range(1, 4)]
    one.close()
    # two shards on two devices, one communicator, combined on the devices
    h = n // 2
    engs = [T._engine(g, h, seed=77, device=d, first_electron_id=d * h) for d in range(2)]
    lk.comm_init_all(engs)
    assert all(e.comm_size() == 2 for e in engs)
    if p2p == "0":
        assert all(e.comm_transport() == "nccl" for e in engs)
    print("transport:", engs[0].comm_transport())
    for d, e in enumerate(engs):
        e.build_tables(maxE); e.set_ensemble(s0[:, d * h:(d + 1) * h], 0.0)

    def combine(es):
        lk.allreduce_results(es)
        a, b = es[0].read_result(), es[1].read_result()
        assert np.array_equal(a, b)          # every rank holds the same combined vector
        return a
    got = _run(engs, nu, 3, combine)
    P = engs[0].P
    exact = want[0][R.N_BORN] + want[0][R.N_ATTACHED] == 0
    for it, (a, b) in enumerate(zip(want, got)):
        if exact or it == 0:                 # population control is shard-local: from the second interval on the ensembles differ statistically
            for j in (R.N_REAL, R.N_NULL, R.N_BORN, R.N_ATTACHED, R.N_SAMPLED):
                assert a[j] == b[j], (it, j)
            assert np.array_equal(a[R.HEADER:R.HEADER + P], b[R.HEADER:R.HEADER + P])
            assert abs(a[R.MAX_EPS_SEEN] - b[R.MAX_EPS_SEEN]) <= 1e-12 * a[R.MAX_EPS_SEEN]
        assert b[R.N_SAMPLED] == n and b[R.OVERFLOW] == 0
        assert abs(a[R.SUM_EPS] - b[R.SUM_EPS]) <= (1e-11 if exact else 5e-3) * a[R.SUM_EPS]
    for e in engs:
        e.close()


def test_two_engine_job_equals_one_engine_job():
    """the job driver (lokib200_job_solve) over two engines with a communicator reproduces the one-engine job: same ensemble, same decisions"""
    import loki_mc_b200 as lk
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    g = gio.load("reid_dc")
    n = 40_000
    out = []
    for shards in (1, 2):
        h = n // shards
        engs = [lk.Engine(g, h, seed=5, device=d, first_electron_id=d * h) for d in range(shards)]
        if shards > 1:
            lk.comm_init_all(engs)
        job = lk.Job(engs, n_integration_points=600, n_integrated_ss_times=0.0)
        r = job.solve()
        hist = job.histograms()
        out.append((r, hist))
        job.close()
        for e in engs:
            e.close()
    a, b = out[0][0], out[1][0]
    assert a["n_sync_points"] == b["n_sync_points"] and a["n_integration_points"] == b["n_integration_points"]
    assert a["total_collisions"] == b["total_collisions"] and a["null_collisions"] == b["null_collisions"]
    assert abs(a["averaged_mean_energy"] / b["averaged_mean_energy"] - 1) < 1e-9
    assert np.allclose(a["flux_drift_velocity"], b["flux_drift_velocity"], rtol=1e-8, atol=1e-9 * abs(a["flux_drift_velocity"][2]))
    assert np.array_equal(out[0][1]["eeh"], out[1][1]["eeh"])      # histograms: combined once per job, counts are integers
    assert np.array_equal(out[0][1]["eah"], out[1][1]["eah"])


@pytest.mark.timeout(600)
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_processes_exchange_over_ipc(p2p, tmp_path):
    """one engine per process (torchrun): mailboxes mapped through CUDA IPC, the exchange kernel a node of the graph-launched blocking interval;
    both ranks must return the SAME bits, equal to the one-engine ensemble (counters exact, sums up to the order of addition)"""
    import os
    import subprocess
    import sys
    import loki_mc_b200 as lk
    R = lk.R
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    name, n = "reid_acb", 200_000
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, LOKIB200_P2P=p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(here, "_two_rank_worker.py"), str(tmp_path), name, str(n)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-3000:]
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    transport = open(tmp_path / "rank0.txt").read()
    assert transport == open(tmp_path / "rank1.txt").read() and transport in (("nccl",) if p2p == "0" else ("peer-memory", "nccl"))
    print("transport:", transport)
    assert np.array_equal(a, b)
    g = gio.load(name)
    s0 = T._start_state(g, n, np.random.default_rng(5), 1e-2, 5.0)
    one = T._engine(g, n, seed=77)
    one.build_tables(12.0); nu = one.table_info()["nu_max_last"]; one.set_ensemble(s0, 0.0)
    want = [one.advance(nu, it / nu, sample=(it != 2)) for it in range(1, 6)]
    one.close()
    P = (a.shape[1] - R.HEADER) // 3
    for it, (w, got) in enumerate(zip(want, a)):
        for j in (R.N_REAL, R.N_NULL, R.N_BORN, R.N_ATTACHED, R.N_SAMPLED):
            assert w[j] == got[j], (it, j)
        assert np.array_equal(w[R.HEADER:R.HEADER + P], got[R.HEADER:R.HEADER + P])
        if w[R.N_SAMPLED] > 0:
            assert abs(w[R.SUM_EPS] - got[R.SUM_EPS]) <= 1e-11 * w[R.SUM_EPS]
        assert got[R.OVERFLOW] == 0
