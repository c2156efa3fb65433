#!/usr/bin/env python3
"""Writes the synthetic input set tests/fixtures/Input/fx/ (LXCat files, property databases, setup files).

The gases 'XY' (molecule with an Effective cross section, vibrational / rotational manifolds, superelastics, dissociation,
ionization, attachment, a momentum-transfer/integral pair) and 'Z' (atom) are invented for the tests: analytic shapes,
no physical meaning, nothing taken from a database.  The files are committed; re-run only to change them, then re-run
oracle/gen_input_golden.py so the reference's answers (tests/golden/input_*.npz) follow.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "Input", "fx")


def block(target, kind, param, comment, pts, extra_comment=True):
    s = "SPECIES: e / %s\nPROCESS: E + %s, %s\nPARAM.:  %s\nCOMMENT: [%s] synthetic test data\n" % (target, target, kind.split(",")[0], param, comment)
    if extra_comment:
        s += "COMMENT: second comment line, ignored by the parser.\n"
    s += "UPDATED: 2024-01-01 00:00:00\nCOLUMNS: Energy (eV) | Cross section (m2)\n-----------------------------\n"
    s += "".join(" %.6e\t%.6e\n" % (e, v) for e, v in pts)
    s += "-----------------------------\n\n"
    return s


def ramp(th, slope, emax=500.0, n=40, peak=None):
    e = th + (emax - th) * (np.linspace(0, 1, n) ** 2.5)
    v = slope * (e - th) / (1.0 + ((e - th) / (peak or 4 * th + 5)) ** 2)
    v[0] = 0.0
    return list(zip(e, v * 1e-20))


def xy_file():
    g = np.concatenate([[0.0], np.logspace(-3, np.log10(800.0), 90)])
    out = ["Synthetic LXCat-format file for the input tests (gas XY).\n\n"]
    eff = [(e, (6.0 / (1 + e / 3.0) + 0.8 * np.sqrt(e) / (1 + (e / 40.0) ** 1.5) + 8.0) * 1e-20) for e in g]
    out.append(block("XY", "Effective", "m/M = 1.9e-05, complete set", "e + XY(X) -> e + XY(X), Effective", eff))
    for j0, j1, th in [(0, 2, 0.0015), (1, 3, 0.0025), (2, 4, 0.0035)]:
        out.append(block("XY", "Excitation", "E = %.6f eV, complete set" % th,
                         "e + XY(X,v=0,J=%d) <-> e + XY(X,v=0,J=%d), Rotational" % (j0, j1), ramp(th, 30.0, emax=50.0, n=25, peak=0.2)))
    for v1, th, sl in [(1, 0.25, 0.9), (2, 0.49, 0.3)]:
        out.append(block("XY", "Excitation", "E = %.6f eV, complete set" % th,
                         "e + XY(X,v=0) <-> e + XY(X,v=%d), Vibrational" % v1, ramp(th, sl, emax=100.0, n=30, peak=1.5)))
    out.append(block("XY", "Excitation", "E = 0.240000 eV, complete set", "e + XY(X,v=1) -> e + XY(X,v=2), Vibrational", ramp(0.24, 0.5, emax=100.0, n=20, peak=1.0)))
    out.append(block("XY", "Excitation", "E = 6.100000 eV, complete set", "e + XY(X) <-> e + XY(A3), Excitation", ramp(6.1, 0.05)))
    # the same collision given twice: integral first, then its momentum-transfer partner (merged by Collision::add)
    out.append(block("XY", "Excitation", "E = 8.400000 eV, complete set", "e + XY(X) -> e + XY(B1), Excitation, integral", ramp(8.4, 0.08)))
    out.append(block("XY", "Excitation", "E = 8.400000 eV, complete set", "e + XY(X) -> e + XY(B1), Excitation, momentum-transfer", ramp(8.4, 0.05)))
    out.append(block("XY", "Excitation", "E = 9.700000 eV, complete set", "e + XY(X) -> e + X(gnd) + X(gnd), Excitation", ramp(9.7, 0.03), extra_comment=False))
    out.append(block("XY", "Excitation", "E = 11.30000 eV, complete set", "e + XY(X) -> e + X(gnd)+Y(1D), Excitation", ramp(11.3, 0.02)))
    out.append(block("XY", "Ionization", "E = 14.20000 eV, complete set", "e + XY(X) -> e + e + XY(+,X), Ionization", ramp(14.2, 0.04, peak=90.0)))
    att = [(e, 2e-3 * np.exp(-((e - 6.5) / 1.2) ** 2) * 1e-20) for e in np.linspace(3.0, 11.0, 33)]
    out.append(block("XY", "Attachment", "E = 0 eV, complete set", "e + XY(X) -> X(-,gnd) + Y(3P), Attachment", att))
    return "".join(out)


def z_file():
    g = np.concatenate([[0.0], np.logspace(-2, 3, 60)])
    out = ["Synthetic LXCat-format file for the input tests (gas Z).\n\n"]
    el = [(e, (3.0 + 5.0 * e / (1 + (e / 12.0) ** 2)) * 1e-20) for e in g[1:]]   # starts above 0: the (0, sigma[0]) point is inserted (Collision.C:45-57)
    out.append(block("Z", "Elastic", "m/M = 2.5e-05, complete set", "e + Z(1S0) -> e + Z(1S0), Elastic", el))
    out.append(block("Z", "Excitation", "E = 10.50000 eV, complete set", "e + Z(1S0) -> e + Z(3P2), Excitation", ramp(10.5, 0.02)))
    out.append(block("Z", "Excitation", "E = 12.00000 eV, complete set", "e + Z(1S0) -> e + Z(1P1), Excitation", ramp(12.0, 0.06)))
    out.append(block("Z", "Ionization", "E = 15.00000 eV, complete set", "e + Z(1S0) -> e + e + Z(+,2P), Ionization", ramp(15.0, 0.05, peak=100.0)))
    return "".join(out)


MASSES = """% masses used by the input tests (kg); expressions are evaluated by the setup parser
XY    (12.5+17.25)*1.660539040e-27   % comment after the value
Z     40*1.660539040e-27
X     12.5*1.660539040e-27
Y     17.25*1.660539040e-27
QQ    1.0e-26
"""
CONSTANTS = {
    "harmonicFrequencies.txt": "% rad/s\nXY  3.1e14\n",
    "anharmonicFrequencies.txt": "% rad/s\nXY  2.2e12\n",
    "rotationalConstants.txt": "% eV\nXY  2.5e-4\n",
    "OPB.txt": "% eV\nXY  13.0\n",
}
XY_ENERGIES = "% energies of the electronic levels (eV)\nXY(X)    0\nXY(A3)   6.1\nXY(B1)   8.4\n"
EFF_POP = "% populations used to take the Elastic cross section out of the Effective one\nXY(X)  1\nXY(X,v=0)  0.9\nXY(X,v=1)  0.1\nXY(X,v=0,J=0) 0.9*0.2\nXY(X,v=0,J=1) 0.9*0.3\nXY(X,v=0,J=2) 0.9*0.5\n"
ANISO_FILE = "% angular models kept in a file\ngroup ; XY ; Rotational ; bornDipole\nsingle;e+XY(X)->e+XY(B1),Excitation;forward  % the integral cross section is given\n"

SETUP_A = """% setup A: mixture, property files + functions, Effective -> Elastic with Boltzmann populations at 300 K
workingConditions:
  gasPressure: 133.32*2
  gasTemperature: 350
  reducedElecField: logspace(0,2,5)
  reducedMagField: 0
  elecFieldAngle: 180
  excitationFrequency: 0
electronKinetics:
  isOn: true
  eedfType: boltzmannMC
  ionizationOperatorType: usingSDCS
  LXCatFiles:
    - fx/XY_LXCat.txt
    - fx/Z_LXCat.txt
  gasProperties:
    mass: fx/masses.txt
    harmonicFrequency: fx/harmonicFrequencies.txt
    anharmonicFrequency: fx/anharmonicFrequencies.txt
    rotationalConstant: fx/rotationalConstants.txt
    OPBParameter: fx/OPB.txt
    fraction:
      - XY = 0.75
      - Z = 1-0.75
  stateProperties:
    energy:
      - fx/XY_energies.txt
      - XY(X,v=*) = harmonicOscillatorEnergy
      - XY(X,v=0,J=*) = rigidRotorEnergy
      - Z(3P2) = 10.5
    statisticalWeight:
      - XY(X) = 1
      - XY(A3) = 3
      - XY(X,v=*) = 1.0
      - XY(X,v=0,J=*) = rotationalDegeneracy
      - Z(1S0) = 1
      - Z(3P2) = 5
    population:
      - XY(X) = 0.98
      - XY(A3) = 0.02
      - XY(X,v=*) = boltzmannPopulation@gasTemperature
      - XY(X,v=0,J=*) = boltzmannPopulation@gasTemperature
      - Z(1S0) = 1
  numericsMC:
    nElectrons: 1000
    gasTemperatureEffect: smartActivation
    nIntegrationPoints: 1E3
gui:
  isOn: false
output:
  isOn: false
"""
SETUP_B = """% setup B: anisotropic models (file + inline), prescribed Effective populations, non-equilibrium populations, AC+B field, all numericsMC keys
workingConditions:
  gasPressure: 1000
  gasTemperature: 300
  electronTemperature: 1.5
  reducedElecField: 60
  reducedMagField: [100,200,400]
  elecFieldAngle: 30
  excitationFrequency: 2.45E9
electronKinetics:
  isOn: true
  eedfType: boltzmannMC
  ionizationOperatorType: oneTakesAll
  LXCatFiles:
    - fx/XY_LXCat.txt
    - fx/Z_LXCat.txt
  effectiveCrossSectionPopulations:
    - fx/XY_effPop.txt
  gasProperties:
    mass: fx/masses.txt
    harmonicFrequency: fx/harmonicFrequencies.txt
    anharmonicFrequency: fx/anharmonicFrequencies.txt
    rotationalConstant: fx/rotationalConstants.txt
    fraction:
      - XY = 0.4
      - Z = 0.6
  stateProperties:
    energy:
      - fx/XY_energies.txt
      - XY(X,v=*) = morseOscillatorEnergy
      - XY(X,v=0,J=*) = rigidRotorEnergy
    statisticalWeight:
      - XY(X) = 1
      - XY(A3) = 3
      - XY(X,v=*) = 1.0
      - XY(X,v=0,J=*) = rotationalDegeneracy_N2
      - Z(1S0) = 1
      - Z(3P2) = 5
    population:
      - XY(X) = 1.0
      - XY(X,v=*) = treanorPopulation@gasTemperature,3000
      - XY(X,v=0,J=*) = boltzmannPopulation@500
      - Z(1S0) = 1
  anisotropicScattering:
    isOn: true
    angleNumber: 1500
    collisions:
      - fx/aniso.txt
      - group;XY;Vibrational;surendra
      - group;Z;Elastic;coulombScreen;0,4.5
      - group;Z;Ionization;coulombScreen;1,25
      - group;XY;Ionization;momentumConservationIonization
      - single;e+Z(1S0)->e+Z(3P2),Excitation;surendra
  numericsMC:
    nElectrons: 2E3
    gasTemperatureEffect: true
    initialElecTempOverGasTemp: 5
    minCollisionsBeforeSteadyState: 10
    maxCollisionsBeforeSteadyState: 2E3
    maxCollisionsAfterSteadyState: 1E4
    nEnergyCells: 500
    nCosAngleCells: 40
    nRadialVelocityCells: 60
    nAxialVelocityCells: 80
    nIntegrationPhases: 24
    nInterpPoints: 5E3
    nIntegrationPoints: 1E3
    nIntegratedSSTimes: 3
    integratedAbsoluteTime: 1E-7
    synchronizationTimeXMaxCollisionFrequency: 2
    synchronizationOverSampling: 3
    relError:
      meanEnergy: 1E-2
      fluxDriftVelocity: 2E-2
      bulkDriftVelocity: 3E-2
      fluxDiffusionCoeffs: 4E-2
      bulkDiffusionCoeffs: 5E-2
      powerBalance: 1E-3
gui:
  isOn: false
output:
  isOn: false
"""


def extra_file():
    out = ["Synthetic LXCat-format file with 'extra' cross sections (rate coefficients only).\n\n"]
    out.append(block("XY", "Excitation", "E = 7.500000 eV, complete set", "e + XY(X) -> e + XY(C3), Excitation", ramp(7.5, 0.011)))
    out.append(block("XY", "Excitation", "E = 3.200000 eV, complete set", "e + XY(A3) <-> e + XY(D1), Excitation", ramp(3.2, 0.02, peak=20.0)))
    return "".join(out)


OUTPUT_ALL = """output:
  isOn: true
  folder: %s
  dataFiles:
    - eedf
    - evdf
    - swarmParameters
    - rateCoefficients
    - powerBalance
    - MCSimDetails
    - MCTemporalInfo
    - MCTemporalInfo_periodic
    - lookUpTable
"""
# small jobs whose files are compared with the reference's Output byte layout (oracle/gen_output_golden.py)
SETUP_OUT_DC = SETUP_A.replace("logspace(0,2,5)", "[20,80]").replace("nElectrons: 1000", "nElectrons: 400") \
    .replace("    nIntegrationPoints: 1E3\n", "    nIntegrationPoints: 500\n    maxCollisionsBeforeSteadyState: 600\n    nEnergyCells: 40\n    nCosAngleCells: 10\n    nRadialVelocityCells: 12\n    nAxialVelocityCells: 12\n") \
    .replace("  LXCatFiles:\n", "  LXCatFilesExtra:\n    - fx/XY_extra_LXCat.txt\n  LXCatFiles:\n") \
    .replace("      - XY(A3) = 3\n", "      - XY(A3) = 3\n      - XY(D1) = 1\n") \
    .replace("output:\n  isOn: false\n", OUTPUT_ALL % "fx_dc").replace("% setup A:", "% output test (DC, two jobs), from setup A:")
SETUP_OUT_AC = SETUP_B.replace("[100,200,400]", "0").replace("nElectrons: 2E3", "nElectrons: 400") \
    .replace("    nEnergyCells: 500\n    nCosAngleCells: 40\n    nRadialVelocityCells: 60\n    nAxialVelocityCells: 80\n    nIntegrationPhases: 24\n",
             "    nEnergyCells: 30\n    nCosAngleCells: 8\n    nRadialVelocityCells: 10\n    nAxialVelocityCells: 10\n    nIntegrationPhases: 6\n") \
    .replace("    nIntegrationPoints: 1E3\n    nIntegratedSSTimes: 3\n    integratedAbsoluteTime: 1E-7\n", "    nIntegrationPoints: 500\n") \
    .replace("    minCollisionsBeforeSteadyState: 10\n    maxCollisionsBeforeSteadyState: 2E3\n    maxCollisionsAfterSteadyState: 1E4\n", "    maxCollisionsBeforeSteadyState: 1500\n") \
    .replace("    relError:\n      meanEnergy: 1E-2\n      fluxDriftVelocity: 2E-2\n      bulkDriftVelocity: 3E-2\n      fluxDiffusionCoeffs: 4E-2\n      bulkDiffusionCoeffs: 5E-2\n      powerBalance: 1E-3\n", "") \
    .replace("output:\n  isOn: false\n", OUTPUT_ALL % "fx_ac").replace("% setup B:", "% output test (AC + B, elecFieldAngle 30), from setup B:")


def main():
    os.makedirs(OUT, exist_ok=True)
    files = {"XY_LXCat.txt": xy_file(), "Z_LXCat.txt": z_file(), "masses.txt": MASSES, "XY_energies.txt": XY_ENERGIES, "XY_effPop.txt": EFF_POP,
             "aniso.txt": ANISO_FILE, "setup_a.in": SETUP_A, "setup_b.in": SETUP_B,
             "XY_extra_LXCat.txt": extra_file(), "setup_out_dc.in": SETUP_OUT_DC, "setup_out_ac.in": SETUP_OUT_AC}
    files.update(CONSTANTS)
    for name, text in files.items():
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
    print("wrote", len(files), "files to", OUT)


if __name__ == "__main__":
    main()
