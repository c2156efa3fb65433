#!/usr/bin/env python3
"""Golden vectors for the INPUT side (setup file + LXCat + property databases -> flattened process set), from the
UNMODIFIED reference (oracle/_ref/harness).  TEST INFRASTRUCTURE ONLY; runs only in the build container.

  tests/golden/input_<setup>.npz   what the reference builds from tests/fixtures/Input/fx/<setup>.in
                                   (Setup.h:229-551 -> BoltzmannMC::allocateEvaluateVariablesFirstTime, BoltzmannMC.C:29-271)
  tests/golden/expressions.json    Parse::str2value / Parse::evalVectorExpress answers (Parse.C:610-755)

usage: python oracle/gen_input_golden.py
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

FIX = os.path.join(HERE, "..", "tests", "fixtures", "Input", "fx")
SETUPS = ["setup_a", "setup_b"]
EXPRESSIONS = ["1", "133.32", "1E3", "2.5e-4", ".5", "1e+2", "1-0.75", "3*14.007*1.660539040e-27", "(14.007+2*15.999)*1.660539040e-27",
               "2*pi*800E6*9.10938356e-31/(1.6021766208e-19*133.32/(1.38064852e-23*300))*1E27", "1E20*1.38064852e-23*300", "2^10", "2^3^2", "-2^2",
               "7%3", "sqrt(2)", "exp(-1.5)", "log(10)", "log10(1e5)", "sin(pi/6)", "cos(PI)", "tan(0.3)", "asin(0.5)", "acos(0.5)",
               "atan(2)", "abs(-3.5)", "sign(-2)", "e", "e^2", "pi*e", "1/3", "10/4*2", "1 + 2 * 3", " ( 1 + 2 ) * 3 ", "2*(3+(4-1))^2", "1e-3*1e3",
               "0.1+0.2", "1.0e-20/3", "4.000000*1.660539040e-27", "1000*9.10938356e-31", "0.9*0.2", "exp(log(7))", "-(2+3)", "2*-3", "6/-2",
               "12345678901234567890", "1.2345678901234567", "100/7", "factorial(4)", "SQRT(16)", "Pi"]
VECTORS = ["linspace(0,10,5)", "logspace(0,2,5)", "logspace(-1,3,9)", "1:5", "0:2.5:10", "1:0.1:2", "[100,200,400]", "[1,2*3,sqrt(16)]", "12", "linspace(1,1,1)",
           "linspace(0,1,4)", "3:3"]


def main():
    dst = os.path.join(gg.REFDIR, "Input", "fx")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(FIX, dst)
    gen = os.path.join(gg.REFDIR, "Input", "_gen")
    os.makedirs(gen, exist_ok=True)
    for s in SETUPS:
        shutil.copy(os.path.join(FIX, s + ".in"), os.path.join(gen, "input_%s.in" % s))
        prefix = gg.run_harness("input_" + s, ["model", "controls"])
        m = gg.parse_model(prefix + ".model.txt")
        out = {k: v for k, v in m.items() if k != "scalars"}
        sc = m["scalars"]
        out["scalar_names"] = np.array(sorted(k for k in sc if np.ndim(sc[k]) == 0))
        out["scalar_values"] = np.array([sc[k] for k in out["scalar_names"]])
        out["electricField"] = sc["electricField"]
        ctl = [ln for ln in open(prefix + ".out.txt") if ln.startswith("controls")][0].split()[1:]
        out["controls"] = np.array([float(x) for x in ctl])
        np.savez_compressed(os.path.join(gg.GOLD, "input_%s.npz" % s), **out)
        print(s, "P =", len(out["p_type"]), "gases =", len(out["gas_first"]))
    prefix = gg.run_harness("input_setup_a", ["expr " + e for e in EXPRESSIONS] + ["vexpr " + e for e in VECTORS])
    lines = [ln.split() for ln in open(prefix + ".out.txt")]
    ex = [float(t[1]).hex() for t in lines if t[0] == "expr"]
    vx = [[float(x).hex() for x in t[1:]] for t in lines if t[0] == "vexpr"]
    assert len(ex) == len(EXPRESSIONS) and len(vx) == len(VECTORS)
    with open(os.path.join(gg.GOLD, "expressions.json"), "w") as f:
        json.dump({"generator": "oracle/gen_input_golden.py", "scalar": dict(zip(EXPRESSIONS, ex)), "vector": dict(zip(VECTORS, vx))}, f, indent=1)
    print("expressions:", len(ex), "scalar,", len(vx), "vector")


if __name__ == "__main__":
    main()
