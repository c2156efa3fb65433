"""worker of test_gpu_comm.py::test_two_processes_exchange_over_ipc: one engine per process (torchrun), the 128-byte communicator id travels over
gloo, the blocking interval ends with the in-engine exchange; every rank writes what the call returned"""
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import golden_io as gio  # noqa: E402
import loki_mc_b200 as lk  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def main():
    out_dir, name, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    g = gio.load(name)
    s0 = T._start_state(g, n, np.random.default_rng(5), 1e-2, 5.0)
    h = n // world
    eng = lk.Engine(g, h, seed=77, device=rank, first_electron_id=rank * h)
    box = [lk.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.comm_init_rank(box[0], rank, world)
    eng.build_tables(12.0)
    nu = eng.table_info()["nu_max_last"]
    eng.set_ensemble(s0[:, rank * h:(rank + 1) * h], 0.0)
    res = [eng.advance(nu, it / nu, sample=(it != 2)) for it in range(1, 6)]       # both graphs of the engine, exchange node included
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(res))
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write(eng.comm_transport())
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


main()
