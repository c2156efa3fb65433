#!/usr/bin/env python3
"""bench.py -- collision events/s (real + null, counted as the reference counts them, BoltzmannMC.C:1308-1320) of the electron
Monte Carlo hot path.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation on the host cores
  python bench.py --time-to-3sigma                         second metric: one setup file end to end, checked against the reference's replicas

Workload (BASELINE.json configs[1]): N2, DC field, E/N = 100 Td point of the sweep, anisotropic scattering (Born-dipole
rotational, Surendra excitation, momentum-conserving ionization), 1e7 electrons per GPU, reference cadence: one synchronisation
interval of 1/nu_trial per step (synchronizationTimeXMaxCollisionFrequency = 1, i.e. on average one trial event per electron per
step) with the ensemble sums of calculateMeanDataForSwarmParams sampled every step.  The process set comes from
tests/golden/n2_aniso.npz (flattened by the unmodified reference from its own LXCat input, see oracle/gen_golden.py); the
ensemble is synthetic: a Maxwellian at the steady-state mean energy, relaxed for --relax intervals before the warm-up.

A "step" = one synchronisation interval over the whole ensemble.  `value` = events of all ranks / device time with the state
resident in HBM and no host round trip per step; `e2e` = the same through the blocking C-ABI call lokib200_advance_to_sync,
i.e. with the per-step host <-> device traffic a real driver has (scalars in, result vector out) and the host-side trial-
frequency logic in the loop.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STATE_BYTES_PER_EVENT = 128.0   # 8 FP64 state words read + written once per trial event at sync factor 1
MEAN_ENERGY_EV = {"n2_aniso": 2.41, "reid_dc": 0.269, "air": 2.74, "arhe": 10.2, "o2_sdcs": 3.43}
KB_OVER_QE = 1.38064852e-23 / 1.6021766208e-19


def load_model(name):
    import golden_io as gio
    return gio.load(name)


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons of one GPU through NVML every 10 ms while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.bits, self._stop_evt = index, [], None, 0, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self._stop_evt.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self._stop_evt.wait(0.01)
        except Exception as ex:   # fall back to one nvidia-smi sample
            self.err = str(ex)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = [n for bit, n in names.items() if self.bits & bit]
        return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None, sm_max_mhz=self.mx, reasons=reasons, samples=len(self.sm))


def combine_results(dist, world, src, gathered, out, sum_count, header):
    """ONE collective per sampling interval: all-gather of the per-GPU result vectors (2.6 KB each), then SUM / MAX locally
    (include/lokib200.h: entries [SUM_COUNT, HEADER) combine with max, all others with sum).  Leaves the combined vector in `out`."""
    import torch
    L = out.numel()
    if world > 1:
        dist.all_gather_into_tensor(gathered, src)
        g2 = gathered.view(world, L)
        torch.sum(g2, dim=0, out=out)
        out[sum_count:header] = g2[:, sum_count:header].max(dim=0).values
    elif src is not out:
        out.copy_(src)
    return out


def combine_ring(dist, world, ring, gathered, sum_count, header):
    """Batched form of combine_results for the device-resident loop: `ring` holds the result vectors of K consecutive intervals
    ([K, L]); ONE all-gather moves all of them (K x 2.6 KB per GPU) and the same SUM / MAX rule is applied per interval.
    Returns the combined [K, L] tensor."""
    import torch
    if world == 1:
        return ring
    dist.all_gather_into_tensor(gathered, ring.reshape(-1))
    g3 = gathered.view(world, ring.shape[0], ring.shape[1])
    comb = g3.sum(dim=0)
    comb[:, sum_count:header] = g3[:, :, sum_count:header].max(dim=0).values
    return comb


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline_port(model_name, mean_energy, n_cpu=200_000, intervals=12):
    """the oracle (a port of the reference algorithm, OpenMP over electrons like BoltzmannMC.C:636) timed on the host cores on a
    bounded sample of the same workload: n_cpu electrons, `intervals` synchronisation intervals after 3 warm-up intervals"""
    from oracle import lokioracle as lo
    g = load_model(model_name)
    m = lo.Model(g)
    ens = lo.Ensemble(m, n_cpu, 12345, 0)
    ratio = mean_energy / (1.5 * KB_OVER_QE * g["cond"]["gas_temperature"])
    ens.init(ratio)
    mx = ens.max_energy()
    t = m.build_tables(2.0 * mx)
    nu = float(t["nu_max"][-1])
    tnow, ev = 0.0, 0
    for it in range(1, 4):
        tnow += 1.0 / nu
        ens.advance(nu, tnow, it, population_control=1)
    t0 = time.perf_counter()
    for it in range(4, 4 + intervals):
        tnow += 1.0 / nu
        r = ens.advance(nu, tnow, it, population_control=1)
        ev += r["n_real"] + r["n_null"]
    dt = time.perf_counter() - t0
    return dict(value=ev / dt, unit="events/s", cores=os.cpu_count(), kind="port",
                sample="%d electrons x %d sync intervals (%.3g events) of the same model, oracle/lokioracle.c with OpenMP on all host cores" % (n_cpu, intervals, ev))


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = load_model(args.model)
    line = dict(metric="collision_events_per_sec", unit="events/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference")
    from oracle import run_reference as rr
    n_ref = args.ref_electrons
    if rr.available():
        text = str(g["setup_text"])
        text = text.replace("nElectrons: 1000", "nElectrons: %d" % n_ref).replace("nIntegrationPoints: 1E3", "nIntegrationPoints: %d" % args.ref_points)
        text = text.replace("output:\n  isOn: false", "output:\n  isOn: true\n  folder: bench_ref\n  dataFiles:\n    - swarmParameters\n    - MCSimDetails")
        res = rr.run(text, "bench_ref")
        d = res["jobs"][0]["details"]
        events = d["total number of real collisions"] + d["total number of null collisions"]
        elapsed = d["Elapsed time"]
        n_int = max(1.0, events / n_ref)   # ~ one trial event per electron per synchronisation interval
        kind, cores = "reference", res["threads"]
        sample = "unmodified lokimc (oracle/_ref, g++ -O2 -fopenmp) on the %s setup with nElectrons=%d, nIntegrationPoints=%d: whole job incl. relaxation, %.3g events in %.1f s" % (
            args.model, n_ref, args.ref_points, events, elapsed)
        mean_e = res["jobs"][0]["swarm"].get("Energy parameters/Mean energy")
    else:
        from oracle import lokioracle as lo
        m = lo.Model(g)
        r = m.solve(n_ref, 1, args.ref_points)
        events, elapsed, n_int = r[29] + r[30], r[34], max(1.0, r[33])
        kind, cores = "port", os.cpu_count()
        sample = "oracle port lo_solve on the %s model, nElectrons=%d, nIntegrationPoints=%d" % (args.model, n_ref, args.ref_points)
        mean_e = r[0]
    value = events / elapsed
    line.update(value=value, ms_per_step=1e3 * elapsed / n_int,
                config=dict(workload="N2 DC 100 Td anisotropic (configs[1] point), reference CPU path", process_set=args.model, electrons=n_ref,
                            sync_factor=1.0, intervals=n_int, mean_energy_eV=mean_e),
                cpu_baseline=dict(value=value, unit="events/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit="events/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def run_time_to_3sigma(args):
    """BASELINE.json's second metric: wall time from job start until the run's own stop criterion is met with every swarm parameter within
    3 sigma_eff of the reference.  Setup: tests/fixtures/Input/fx/setup_out_dc.in (two gases, anisotropic models, two E/N
    jobs) at 2e4 electrons, through the setup-file entry point (lokib200_run_setup: parse -> solve -> post-process -> write the output
    folder).  sigma_eff per parameter = max(reference's reported std, scatter of its 10 replicas, this run's reported std), from
    tests/golden/ensemble_fixture.json.  Beside it: the unmodified reference on the same setup on the host cores, when oracle/_ref is there."""
    import shutil
    import tempfile
    import loki_mc_b200 as lk
    from oracle import run_reference as rr
    fix = os.path.join(ROOT, "tests", "fixtures", "Input")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ensemble_fixture.json")))["jobs"]
    n = 20000
    text = open(os.path.join(fix, "fx", "setup_out_dc.in")).read().replace("nElectrons: 400", "nElectrons: %d" % n)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "job.in")
    with open(path, "w") as f:
        f.write(text)
    lk.run_setup(fix, path, os.path.join(tmp, "warm"), verbose=False)          # untimed: context creation, module load
    t0 = time.perf_counter()
    lk.run_setup(fix, path, os.path.join(tmp, "out"), verbose=False)
    wall = time.perf_counter() - t0
    worst, checked, events = 0.0, 0, 0.0
    for sub in sorted(os.listdir(os.path.join(tmp, "out", "fx_dc"))):
        d = os.path.join(tmp, "out", "fx_dc", sub)
        if not os.path.isdir(d):
            continue
        mine, g = rr.parse_swarm(os.path.join(d, "swarmParameters.txt")), gold["setup_out_dc/" + sub]
        det = rr.parse_details(os.path.join(d, "MCSimDetails.txt"))
        events += det["total number of real collisions"] + det["total number of null collisions"]
        for key, mean in g["mean"].items():
            if key.endswith("v_x") or mean == 0 or key not in mine:
                continue                                                         # components that vanish by symmetry are pure noise
            rel = max(g["reported_relstd"].get(key, 0.0), g["std"][key] / abs(mean), mine.get(key + "/relstd", 0.0), 4e-3)
            worst = max(worst, abs(mine[key] - mean) / (rel * abs(mean)))
            checked += 1
    line = dict(metric="time_to_3sigma_s", value=wall, unit="s", higher_is_better=False, n_gpus=1, data="synthetic fixture gases (tests/fixtures/Input/fx)",
                config=dict(workload="fx/setup_out_dc.in: 2 jobs (E/N = 20, 80 Td), 2e4 electrons, reference stop criterion", electrons=n),
                within_3sigma=bool(worst <= 3.0), worst_deviation_sigma=worst, parameters_checked=checked, events=events, events_per_s=events / wall)
    if rr.available():
        dst = os.path.join(rr.REFDIR, "Input", "fx")
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(fix, "fx"), dst)
        res = rr.run(text, "fx_dc")
        line["reference"] = dict(value=res["wall"], unit="s", cores=res["threads"], kind="reference",
                                 sample="unmodified lokimc (oracle/_ref) on the same setup file and electron count, all host threads")
    shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="n2_aniso")
    ap.add_argument("--electrons", type=float, default=1e7, help="electrons per GPU")
    ap.add_argument("--relax", type=int, default=60, help="untimed relaxation intervals before the warm-up")
    ap.add_argument("--ref-electrons", type=int, default=200_000)
    ap.add_argument("--ref-points", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=16, help="device leg: intervals whose result vectors share one collective")
    ap.add_argument("--time-to-3sigma", action="store_true", help="second metric of BASELINE.json: setup file in, swarm parameters out, on one GPU")
    args = ap.parse_args()
    if args.time_to_3sigma:
        return run_time_to_3sigma(args)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import loki_mc_b200 as lk
    R = lk.R
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = load_model(args.model)
    n = int(args.electrons)
    eng = lk.Engine(g, n, seed=0x4C6F4B49, device=local, first_electron_id=rank * n)
    # everything (engine kernels, NCCL all-reduces, timing events) runs on ONE explicit non-default torch stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    P, L = eng.P, lk.result_len(eng.P)
    mean_e = MEAN_ENERGY_EV.get(args.model, 1.0)
    ratio = mean_e / (1.5 * KB_OVER_QE * g["cond"]["gas_temperature"])
    mx = eng.init_ensemble(ratio)
    eng.build_tables(2.0 * mx)
    nu = eng.check_nu_trial(mx, eng.table_info()["nu_max_last"], horizon=11.0)

    d_res = torch.zeros(L, dtype=torch.float64, device="cuda")
    d_buf = [torch.zeros(L, dtype=torch.float64, device="cuda") for _ in range(2)]
    d_gather = torch.zeros(world * L, dtype=torch.float64, device="cuda")
    d_events = torch.zeros(1, dtype=torch.float64, device="cuda")
    comm_stream = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def combine(src):
        return combine_results(dist, world, src, d_gather, d_res, R.SUM_COUNT, R.HEADER)

    def host_step(t):
        """the blocking C-ABI call with the host-side trial-frequency logic a driver runs every interval"""
        nonlocal nu, mx
        nu = eng.check_nu_trial(mx, nu, horizon=11.0)
        t += 1.0 / nu
        if world > 1:   # the result vector stays on the device until the ranks have combined it: one collective, one device -> host read per interval
            eng.advance_device(nu, t, True, d_buf[0].data_ptr()); combine(d_buf[0]); res = d_res.cpu().numpy()
        else:
            res = eng.advance(nu, t, sample=True)
        mx = max(res[R.MAX_EPS], res[R.MAX_EPS_SEEN])
        return t, res

    # ---- relaxation + warm-up (untimed) ----
    t = eng.time
    for _ in range(args.relax + args.warmup):
        t, res = host_step(t)
    mean_energy_now = res[R.SUM_EPS] / res[R.N_SAMPLED]

    # ---- timed leg 1: e2e through the blocking C-ABI call ----
    sampler = ClockSampler(local); sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_e2e = 0.0
    e0.record()
    for _ in range(args.steps):
        t, res = host_step(t)
        ev_e2e += res[R.N_REAL] + res[R.N_NULL]
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- timed leg 2: device-resident (no host round trip inside the region) ----
    nu = eng.check_nu_trial(mx, nu, horizon=float(args.steps) + 11.0)   # one bound for the whole region
    eng.kernel_time_ms()                                                 # reset the per-kernel event log
    launches0 = eng.launch_count()
    d_events.zero_()
    K = max(1, min(args.batch, args.steps))               # intervals per collective (result vectors wait in a device ring)
    rings = [torch.zeros(K, L, dtype=torch.float64, device="cuda") for _ in range(2)]
    d_gather_ring = torch.zeros(world * K * L, dtype=torch.float64, device="cuda")
    with torch.cuda.stream(comm_stream):                 # untimed: first use of the combine kernels (module load) and of the collective
        comb = combine_ring(dist, world, rings[0], d_gather_ring, R.SUM_COUNT, R.HEADER)
        d_events.add_(comb[:, R.N_REAL].sum() + comb[:, R.N_NULL].sum())
        d_res.copy_(comb[0])
    comm_stream.synchronize()
    d_events.zero_()
    barrier()
    e0.record()
    free_evt = [None, None]
    for i in range(args.steps):
        t += 1.0 / nu
        slot, which = i % K, (i // K) & 1
        ring = rings[which]
        if slot == 0 and free_evt[which] is not None:
            stream.wait_event(free_evt[which])          # the collective that read this ring two batches ago is done
        eng.advance_device(nu, t, True, ring[slot].data_ptr())
        if slot == K - 1 or i == args.steps - 1:
            ready = torch.cuda.Event(); ready.record(stream)
            with torch.cuda.stream(comm_stream):        # one collective per K intervals, on a side stream
                comm_stream.wait_event(ready)
                comb = combine_ring(dist, world, ring, d_gather_ring, R.SUM_COUNT, R.HEADER)[:slot + 1]
                d_events.add_(comb[:, R.N_REAL].sum() + comb[:, R.N_NULL].sum())
                d_res.copy_(comb[slot])
                free_evt[which] = torch.cuda.Event(); free_evt[which].record(comm_stream)
    stream.wait_stream(comm_stream)
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop()
    adv_ms, adv_n = eng.kernel_time_ms()
    fp64_peak = eng.measure_fp64_peak() if rank == 0 else None
    launches = eng.launch_count() - launches0
    ev_dev = float(d_events.item())   # already the sum over ranks when world > 1

    tm = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(tm[0]), float(tm[1])
    final = d_res.cpu().numpy()
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        traffic, fp64 = None, None
        try:   # DRAM bytes and FP64 operations of one K1 launch from the committed ncu captures (same workload only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if tj["electrons"] == n and args.model == "n2_aniso":
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                # second ceiling of SURVEY.md 8(d): FP64 pipe.  flop per event from ncu's SASS counters (DFMA = 2), peak = DFMA chain measured live
                flop_per_event = tj["fp64_flop"] / tj["events"]
                fp64 = dict(flop_per_event=flop_per_event, peak=fp64_peak, unit="TFLOP/s", peak_source="measured live (lokib200_measure_fp64_peak: 16 DFMA chains per thread)")
        except Exception:
            pass
        ev_per_launch_rank = ev_dev / world / args.steps
        achieved = STATE_BYTES_PER_EVENT * ev_per_launch_rank / (adv_ms * 1e-3) / 1e9 if adv_ms > 0 else None
        line = dict(
            metric="collision_events_per_sec", value=ev_dev / (ms_dev * 1e-3), unit="events/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload="N2 DC E/N=100 Td anisotropic scattering, 1e7 electrons per GPU, reference cadence (sync factor 1, ensemble sums every interval) [BASELINE.json configs[1]]",
                        process_set=args.model, electrons_per_gpu=n, processes=P, sync_factor=1.0, relax_intervals=args.relax,
                        mean_energy_eV=mean_energy_now, nu_trial=nu, table_mib=round(eng.table_info()["nE"] * ((P + 15) // 16 * 16) * 8 * 3 / 2 ** 20, 1),   # cumulative table (8 B / entry) + its row-pair form (16 B / entry)
                        l2_policy="state 640 MB per GPU >> 126 MB L2: every step streams it from HBM", intervals_per_collective=K, real_fraction=float(final[R.N_REAL] / (final[R.N_REAL] + final[R.N_NULL]))),
            e2e=dict(value=ev_e2e / (ms_e2e * 1e-3), unit="events/s", h2d_bytes_per_step=16 * world,
                     d2h_bytes_per_step=8 * L * world, ms_per_step=ms_e2e / args.steps),
            gpu_launches=int(launches * world), clocks=clocks,
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=(achieved / peak) if achieved else None, traffic=traffic,
                          kernel="k_advance", kernel_ms=adv_ms, kernel_launches=adv_n, bytes_per_event=STATE_BYTES_PER_EVENT, peak_source=peak_src,
                          kernel_share_of_step=adv_ms * args.steps / ms_dev if ms_dev > 0 else None))   # (the engine times at most its first 256 launches per leg)
        if fp64 is not None and adv_ms > 0:
            fp64["achieved"] = fp64["flop_per_event"] * ev_per_launch_rank / (adv_ms * 1e-3) / 1e12
            fp64["frac"] = fp64["achieved"] / fp64["peak"]
            line["roofline"]["fp64"] = fp64
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_port(args.model, mean_e)
            except Exception as ex:   # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = dict(value=None, unit="events/s", cores=os.cpu_count(), kind="port", sample="failed: %s" % ex)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
