#!/usr/bin/env python3
"""Golden output files for the post-processing and the text writers, from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

For the fixture setups tests/fixtures/Input/fx/setup_out_*.in the reference (oracle/_ref/harness `solve`) runs every job on a
deterministic generator, its own Output class (Headers/Output.h) writes the files, and the harness dumps the raw BoltzmannMC state
the sinks read.  Committed as tests/golden/output_<folder>.tgz: the files + job<k>.raw.bin.  The test feeds the raw state to this
project's report (post-processing + writers) and compares the files.

usage: python oracle/gen_output_golden.py
"""
import os
import shutil
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

FIX = os.path.join(HERE, "..", "tests", "fixtures", "Input", "fx")
CASES = {"setup_out_dc": ("fx_dc", 20240611), "setup_out_ac": ("fx_ac", 777)}


def main():
    dst = os.path.join(gg.REFDIR, "Input", "fx")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(FIX, dst)
    gen = os.path.join(gg.REFDIR, "Input", "_gen")
    os.makedirs(gen, exist_ok=True)
    for setup, (folder, seed) in CASES.items():
        shutil.copy(os.path.join(FIX, setup + ".in"), os.path.join(gen, "output_%s.in" % setup))
        outdir = os.path.join(gg.REFDIR, "Output", folder)
        if os.path.isdir(outdir):
            shutil.rmtree(outdir)
        prefix = gg.run_harness("output_" + setup, ["solve %d" % seed])
        k = 0
        while os.path.exists("%s.job%d.raw.bin" % (prefix, k)):
            shutil.copy("%s.job%d.raw.bin" % (prefix, k), os.path.join(outdir, "job%d.raw.bin" % k))
            k += 1
        gold = os.path.join(gg.GOLD, "output_%s.tgz" % folder)
        with tarfile.open(gold, "w:gz") as t:
            t.add(outdir, arcname=folder)
        print(setup, "->", gold, ":", k, "jobs;", open(prefix + ".out.txt").read().strip())


if __name__ == "__main__":
    main()
