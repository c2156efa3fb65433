import os, sys, time, tempfile, shutil
sys.path.insert(0, os.getcwd())
import loki_mc_b200 as lk
inp = os.path.join(os.getcwd(), "oracle", "_ref", "Input")
text = open(os.path.join(inp, "default_setup.in")).read().replace("gui: \n  isOn: true", "gui: \n  isOn: false")
tmp = tempfile.mkdtemp()
p = os.path.join(tmp, "job.in"); open(p, "w").write(text)
w = os.path.join(tmp, "warm.in"); open(w, "w").write(text.replace("nIntegrationPoints: 1E4", "nIntegrationPoints: 500").replace("[1,5,10,50,100]", "[10]"))
lk.run_setup(inp, w, os.path.join(tmp, "warm"), verbose=False)
t0 = time.perf_counter()
lk.run_setup(inp, p, os.path.join(tmp, "out"), verbose=True)
print("wall", time.perf_counter() - t0)
for sub in sorted(os.listdir(os.path.join(tmp, "out", "swarm_O2_short"))):
    f = os.path.join(tmp, "out", "swarm_O2_short", sub, "MCSimDetails.txt")
    if os.path.exists(f): print(sub, open(f).read())
