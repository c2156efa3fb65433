"""N > 1 host-side logic on CPU (gloo, world_size 2): the rule by which the shards' result vectors combine per sampling interval (what
lokib200_comm_allreduce_results does over NCCL inside the engine: three all-reduces, SUM / MAX / SUM), and the
sharding of an ensemble by global electron id.  The CPU oracle stands in for the engine (it defines the same draw streams), so
these tests pin the property the multi-GPU path relies on: shards keyed by global electron id reproduce the single-process
trajectories, integer-valued outputs add up exactly and the ensemble sums agree to rounding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SUM_COUNT, HEADER = 34, 37   # LOKIB200_R_SUM_COUNT, LOKIB200_R_HEADER


def _worker(rank, world, port, name, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from oracle import lokioracle as lo
    g = gio.load(name)
    P = len(g["p_type"]); L = HEADER + 3 * P
    m = lo.Model(g); t = m.build_tables(10.0); nu = float(t["nu_max"][-1])
    rng = np.random.default_rng(5)
    eps = np.exp(rng.uniform(np.log(1e-2), np.log(3.0), n))
    d = rng.normal(size=(3, n)); d /= np.linalg.norm(d, axis=0)
    s0 = np.vstack([rng.normal(size=(3, n)) * 1e-3, d * np.sqrt(2 * eps * gio.QE / gio.ME), np.full((1, n), -123456789.0), np.zeros((1, n))])
    lo_i, hi_i = rank * n // world, (rank + 1) * n // world
    ens = lo.Ensemble(m, hi_i - lo_i, 77, lo_i); ens.set(s0[:, lo_i:hi_i])
    total = None
    for it in range(1, 4):
        r = ens.advance(nu, it / nu, it, population_control=0)
        st = ens.get(); mom_n = st.shape[1]
        vec = np.zeros(L)
        vec[0], vec[1], vec[4] = r["n_real"], r["n_null"], r["field"]
        e = gio.energy_eV(st[3:6].T)
        vec[6] = e.sum(); vec[7:10] = st[0:3].sum(1); vec[10:13] = st[3:6].sum(1); vec[31] = mom_n
        vec[34] = e.max()
        vec[HEADER:HEADER + P] = r["counts"]; vec[HEADER + P:HEADER + 2 * P] = r["gain"]; vec[HEADER + 2 * P:] = r["loss"]
        vec[36] = 1.0 if (rank == 1 and it == 2) else 0.0       # an overflow flag raised on one rank must reach every rank (MAX entry)
        out = bench.allreduce_rule(dist, torch.from_numpy(vec), SUM_COUNT, HEADER)
        assert out[36] == (1.0 if it == 2 else 0.0)
        out[36] = 0.0
        total = out.clone() if total is None else total + out
    if rank == 0:
        q.put((total.numpy(), out.numpy()))
    full = np.zeros((8, n))
    full[:, lo_i:hi_i] = ens.get()
    ft = torch.from_numpy(full); dist.all_reduce(ft)
    if rank == 0:
        q.put(ft.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["reid_dc", "reid_acb"])
def test_two_ranks_reproduce_one(name):
    from oracle import lokioracle as lo
    n, world = 6000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    total2, last2 = q.get(timeout=120)
    state2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run of the whole ensemble
    g = gio.load(name)
    P = len(g["p_type"])
    m = lo.Model(g); t = m.build_tables(10.0); nu = float(t["nu_max"][-1])
    rng = np.random.default_rng(5)
    eps = np.exp(rng.uniform(np.log(1e-2), np.log(3.0), n))
    d = rng.normal(size=(3, n)); d /= np.linalg.norm(d, axis=0)
    s0 = np.vstack([rng.normal(size=(3, n)) * 1e-3, d * np.sqrt(2 * eps * gio.QE / gio.ME), np.full((1, n), -123456789.0), np.zeros((1, n))])
    ens = lo.Ensemble(m, n, 77, 0); ens.set(s0)
    real = null = 0; counts = np.zeros(P)
    for it in range(1, 4):
        r = ens.advance(nu, it / nu, it, population_control=0)
        real += r["n_real"]; null += r["n_null"]; counts += r["counts"]
    st = ens.get()
    assert np.array_equal(state2, st)                       # shards keyed by global id == the single ensemble, bit for bit
    assert total2[0] == real and total2[1] == null
    assert np.array_equal(total2[HEADER:HEADER + P], counts)
    e = gio.energy_eV(st[3:6].T)
    assert last2[31] == n
    assert abs(last2[6] - e.sum()) <= 1e-12 * e.sum()
    assert last2[34] == e.max()                              # MAX entries combine with max, not sum
    assert np.allclose(last2[7:10], st[0:3].sum(1), rtol=1e-10, atol=1e-18)
