#!/usr/bin/env python3
"""DRAM traffic and FP64 operation counts of ONE launch of the advance kernel on the bench workload, from ncu, stamped with the build id of the
CUDA sources (bench.py ignores a capture whose build id differs from the sources it runs).  Run on a GPU box:
    python tools/capture_traffic.py            -> profiles/r2_traffic.json (+ the ncu CSV next to it)
A number printed by bench.py under ncu is never a bench value; only the counters are kept."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active"]
out_csv = os.path.join(ROOT, "profiles", "r2_traffic_ncu.csv")
cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:k_advance_stream", "--launch-skip", "80", "-c", "3", "--csv", "--log-file", out_csv,
       sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-extras", "--no-cpu-baseline"]
r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
line = json.loads([l for l in r.stdout.split("\n") if l.startswith("{")][-1])
rows = [x for x in csv.reader(open(out_csv)) if len(x) > 14 and x[0].isdigit()]
per = {}
for x in rows:
    per.setdefault(x[0], {})[x[12]] = float(x[14].replace(",", ""))
ids = sorted(per, key=int)
avg = {k: sum(per[i][k] for i in ids) / len(ids) for k in METRICS}
unit = {x[12]: x[13] for x in rows}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd = avg["dram__bytes_read.sum"] * scale.get(unit["dram__bytes_read.sum"], 1.0)
wr = avg["dram__bytes_write.sum"] * scale.get(unit["dram__bytes_write.sum"], 1.0)
dfma, dadd, dmul = (avg["smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k] for k in ("dfma", "dadd", "dmul"))
n = line["config"]["electrons_per_gpu"]
events = n * 1.0   # one trial event per electron per interval at sync factor 1 (the bench reports the exact count; the mean is n to 1e-3)
t_unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit["gpu__time_duration.sum"]]
out = dict(kernel=line["roofline"]["kernel"], model=line["config"]["process_set"], electrons=n, events=events, build=bench.build_id(), launches_averaged=len(ids),
           dram_bytes_read=rd, dram_bytes_write=wr, fp64_flop=2 * dfma + dadd + dmul, fp64_ops=dict(dfma=dfma, dadd=dadd, dmul=dmul),
           warp_instructions=avg["smsp__inst_executed.sum"], ncu_duration_ms=avg["gpu__time_duration.sum"] * t_unit,
           fp64_pipe_pct_active=avg["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], issue_active_pct=avg["smsp__issue_active.avg.pct_of_peak_sustained_active"],
           lanes_per_instruction=avg["smsp__thread_inst_executed_per_inst_executed.ratio"], warps_active_pct=avg["sm__warps_active.avg.pct_of_peak_sustained_active"],
           source="tools/capture_traffic.py: ncu --metrics ... --clock-control none, %d launches of the relaxed bench workload averaged; CSV in profiles/r2_traffic_ncu.csv" % len(ids))
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out))
