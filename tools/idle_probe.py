"""Why is K1 ~100 us slower inside the blocking call than in the device-resident loop?  Same kernels, same arguments; the only difference is
the host round trip between intervals.  Variants of the device-resident loop over 1e7 N2 electrons, K1 timed by the engine's own CUDA events."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import golden_io as gio
import loki_mc_b200 as lk

g = gio.load("n2_aniso")
n = 10_000_000
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
eng = lk.Engine(g, n, seed=7)
eng.set_stream(stream.cuda_stream)
eng.build_tables(60.0)
eng.init_ensemble(0.01)
nu = eng.table_info()["nu_max_last"]
t = 0.0
for _ in range(60):
    t += 1.0 / nu; eng.advance(nu, t, sample=True)
filler = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")       # 256 MB: one pass ~ 80 us

def run(tag, between):
    global t
    eng.kernel_time_ms()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(100):
        t += 1.0 / nu
        eng.advance_device(nu, t, True, None)
        between()
    e1.record(); torch.cuda.synchronize()
    ms, k = eng.kernel_time_ms()
    print("%-62s K1 %.4f ms   step %.4f ms" % (tag, ms, e0.elapsed_time(e1) / 100), flush=True)

def sync(): torch.cuda.synchronize()
def sync_sleep():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    while time.perf_counter() - t0 < 300e-6: pass
def filler_then_sync_early():
    # the host waits for the interval (event), while a filler kernel keeps the GPU busy until after the next K1 is queued
    ev = torch.cuda.Event(); ev.record(); filler.add_(1.0); ev.synchronize()
def filler_only(): filler.add_(1.0)

run("back to back", lambda: None)
run("stream sync after every interval", sync)
run("stream sync + 300 us host pause", sync_sleep)
run("filler kernel (256 MB add) after every interval, no sync", filler_only)
run("host waits for the interval, filler keeps the GPU busy meanwhile", filler_then_sync_early)
run("back to back (again)", lambda: None)
eng.close()
