"""Loader for tests/golden/*.npz (generated from the unmodified reference by oracle/gen_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["reid_dc", "reid_ac", "reid_b", "reid_ecr", "reid_acb", "reid_true_aniso", "o2_sdcs", "n2_aniso", "n2_true_acb",
          "arhe", "arhe_true", "air", "ls_f05", "ls_att_aniso"]
# the four non-DC field branches x {gasTemperatureEffect true, smartActivation}: with the models above, all 15 instantiations
# (5 field cases x 3 thermal modes) of the advance kernels are covered
FIELD_GT_MODELS = ["reid_%s_%s" % (f, t) for f in ("ac", "b", "ecr", "acb") for t in ("true", "smart")]
MODELS += FIELD_GT_MODELS


def load(name):
    """Returns a dict: process SoA ('p_*', 'gas_*', 'xs_*'), 'cond' (job conditions) and the golden arrays."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    sc = dict(zip([str(s) for s in g["scalar_names"]], g["scalar_values"]))
    g["scalars"] = sc
    g["cond"] = dict(
        gas_temperature_effect=int(sc["gasTemperatureEffect"]), ionization_sharing=int(sc["energySharingIonizType"]),
        energy_sharing_factor=float(sc["energySharingFactor"]), gas_density=float(sc["totalGasDensity"]),
        gas_temperature=float(sc["gasTemperature"]), electric_field=np.array(g["electricField"], dtype=np.float64),
        excitation_omega=float(sc["excitationFrequencyRadians"]), cyclotron_omega=float(sc["cyclotronFrequency"]),
        n_interp_points=int(sc["nInterpPoints"]), is_cylindrically_symmetric=int(sc["isCylindricallySymmetric"]),
        energy_max_elastic=float(sc["energyMaxElastic"]))
    g["name"] = name
    return g


# columns of ev_in / ev_out (oracle/harness.cpp 'event')
EV_IN = dict(nu_trial=0, t_e=1, r=slice(2, 5), v=slice(5, 8), t_cf=8, nu_e=9, t_sync=10, draws=slice(11, None))
EV_OUT = dict(chosen=0, r=slice(1, 4), v=slice(4, 7), eps=7, t_e=8, t_cf=9, nu_e=10, dE=11, dE_rel=12, gain_field=13,
              ej_r=slice(14, 17), ej_v=slice(17, 20), ej_eps=20, draws_used=21)
ME, QE = 9.10938356e-31, 1.6021766208e-19


def energy_eV(v):
    v = np.asarray(v)
    return 0.5 * ME * ((v[..., 0] ** 2 + v[..., 1] ** 2) + v[..., 2] ** 2) / QE
