"""Run the UNMODIFIED reference binary (oracle/_ref/lokimc, built by oracle/Makefile from /root/reference) on a setup text and
parse its own output files (Headers/Output.h: swarmParameters.txt :258-341, MCSimDetails.txt :784-823, rateCoefficientsMC.txt).

TEST / BASELINE INFRASTRUCTURE ONLY.  Works wherever oracle/_ref/ travelled (it is git-ignored but not gpurun-ignored)."""
import os
import re
import shutil
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REFDIR, "lokimc")) and os.path.isdir(os.path.join(REFDIR, "Input"))


_num = r"[-+]?\d\.\d+e[-+]\d+"


def parse_swarm(path):
    out = {}
    section = ""
    with open(path) as f:
        for line in f:
            m = re.match(r"\*+ (.*?) \*+", line.strip())
            if m:
                section = m.group(1)
                continue
            m = re.match(r"\s*(.*?) = (%s) \((.*?)\)\s*(?:;\s*Rel\. std:\s*(%s)%%)?" % (_num, _num), line)
            if m:
                key = (section + "/" + m.group(1).strip()) if section else m.group(1).strip()
                out[key] = float(m.group(2))
                if m.group(4):
                    out[key + "/relstd"] = float(m.group(4)) / 100.0
                continue
            m = re.match(r"\s*\| (v_[xyz]'?) \|\s+=?\s*\| (%s)\s*\|.*?\| (%s|-?nan|inf)\s*%% \|" % (_num, _num), line)
            if m:
                key = section + "/" + m.group(1)
                out[key] = float(m.group(2))
                try:
                    out[key + "/relstd"] = float(m.group(3)) / 100.0
                except ValueError:
                    pass
    return out


def parse_details(path):
    out = {}
    with open(path) as f:
        for line in f:
            m = re.match(r"\s*(.*?):\s+(%s|\d+)" % _num, line)
            if m:
                out[m.group(1).strip()] = float(m.group(2))
    return out


def run(setup_text, name, threads=None, timeout=3600, keep=False):
    """setup_text must contain `output:\n  isOn: true\n  folder: <name>` with dataFiles swarmParameters + MCSimDetails."""
    if not available():
        raise RuntimeError("oracle/_ref/lokimc is not built (needs /root/reference; run `make -C oracle ref`)")
    gen = os.path.join(REFDIR, "Input", "_gen")
    os.makedirs(gen, exist_ok=True)
    with open(os.path.join(gen, name + ".in"), "w") as f:
        f.write(setup_text)
    outdir = os.path.join(REFDIR, "Output", name)
    shutil.rmtree(outdir, ignore_errors=True)
    os.makedirs(os.path.join(REFDIR, "Output"), exist_ok=True)
    threads = threads or os.cpu_count()
    t0 = time.time()
    r = subprocess.run([os.path.join(REFDIR, "lokimc"), "_gen/%s.in" % name, str(threads)], cwd=REFDIR, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("lokimc failed:\n" + r.stdout[-2000:])
    jobs = []
    subdirs = sorted(d for d in os.listdir(outdir) if os.path.isdir(os.path.join(outdir, d)))
    for d in (subdirs or ["."]):
        p = os.path.join(outdir, d)
        job = dict(folder=d, swarm=parse_swarm(os.path.join(p, "swarmParameters.txt")), details=parse_details(os.path.join(p, "MCSimDetails.txt")))
        eedf = os.path.join(p, "eedf.txt")
        if os.path.exists(eedf):
            rows = [[float(x) for x in ln.split()] for ln in open(eedf).read().split("\n")[1:] if ln.strip()]
            job["eedf"] = rows
        jobs.append(job)
    if not keep:
        shutil.rmtree(outdir, ignore_errors=True)
    return dict(jobs=jobs, wall=wall, threads=threads)
