"""ctypes binding of oracle/liblokioracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ND = -123456789.0

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_up = C.POINTER(C.c_uint64)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liblokioracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.lo_model_create.restype = C.c_void_p
        L.lo_model_create.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_ip] + [c_dp] * 11 + [c_ip, c_ip, c_dp, c_lp, c_dp, c_dp]
        L.lo_model_set_conditions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_dp, C.c_double, C.c_double, C.c_int]
        L.lo_model_destroy.argtypes = [C.c_void_p]
        L.lo_build_tables.argtypes = [C.c_void_p, C.c_double]
        L.lo_table_size.argtypes = [C.c_void_p]
        L.lo_table_step.argtypes = [C.c_void_p]; L.lo_table_step.restype = C.c_double
        for f in ("lo_table_sigma", "lo_table_cum", "lo_table_nu_tot", "lo_table_nu_max"):
            getattr(L, f).argtypes = [C.c_void_p]; getattr(L, f).restype = c_dp
        L.lo_max_accel_energy.argtypes = [C.c_void_p, C.c_double, C.c_double]; L.lo_max_accel_energy.restype = C.c_double
        L.lo_event_injected.argtypes = [C.c_void_p, C.c_double, C.c_double, c_dp, c_dp, C.c_int, c_dp, C.POINTER(C.c_int)]
        L.lo_moments.argtypes = [C.c_int64] + [c_dp] * 7
        L.lo_histograms.argtypes = [C.c_int64, c_dp, c_dp, c_dp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.lo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.lo_stream_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]; L.lo_stream_uniform.restype = C.c_double
        L.lo_ensemble_create.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64]; L.lo_ensemble_create.restype = C.c_void_p
        L.lo_ensemble_destroy.argtypes = [C.c_void_p]
        L.lo_ensemble_init.argtypes = [C.c_void_p, C.c_double]
        L.lo_ensemble_set.argtypes = [C.c_void_p, c_dp]
        L.lo_ensemble_get.argtypes = [C.c_void_p, c_dp]
        L.lo_ensemble_max_energy.argtypes = [C.c_void_p]; L.lo_ensemble_max_energy.restype = C.c_double
        L.lo_ensemble_time.argtypes = [C.c_void_p]; L.lo_ensemble_time.restype = C.c_double
        L.lo_ensemble_advance.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint32, C.c_int, c_up, c_up, c_dp, c_dp, c_dp]
        L.lo_solve.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, c_dp, c_dp]
        _LIB = L
    return _LIB


class Model:
    """Process set + job conditions (see tests/golden_io.py for the dict layout)."""

    def __init__(self, g):
        L = lib()
        self.g = g
        self.P = int(len(g["p_type"])); self.nG = int(len(g["gas_first"]))
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [i32(g["p_type"]), i32(g["p_superelastic"]), i32(g["p_angular"]), f64(g["p_ap0"]), f64(g["p_ap1"]), f64(g["p_swf"]),
                      f64(g["p_emin"]), f64(g["p_emax"]), f64(g["p_reldens"]), f64(g["p_mass"]), f64(g["p_redmass"]), f64(g["p_eloss"]),
                      f64(g["p_thstd"]), f64(g["p_w"]), i32(g["gas_first"]), i32(g["gas_last"]), f64(g["gas_fraction"]),
                      np.ascontiguousarray(g["xs_offset"], dtype=np.int64), f64(g["xs_energy"]), f64(g["xs_value"])]
        k = self._keep
        args = [k[0].ctypes.data_as(c_ip), k[1].ctypes.data_as(c_ip), k[2].ctypes.data_as(c_ip)] + [_dp(a) for a in k[3:14]] + \
               [k[14].ctypes.data_as(c_ip), k[15].ctypes.data_as(c_ip), _dp(k[16]), k[17].ctypes.data_as(c_lp), _dp(k[18]), _dp(k[19])]
        self.h = L.lo_model_create(self.P, self.nG, *args)
        c = g["cond"]
        E = f64(c["electric_field"])
        L.lo_model_set_conditions(self.h, int(c["gas_temperature_effect"]), int(c["ionization_sharing"]), float(c["energy_sharing_factor"]),
                                  float(c["gas_density"]), float(c["gas_temperature"]), _dp(E), float(c["excitation_omega"]),
                                  float(c["cyclotron_omega"]), int(c.get("n_interp_points", 10000)))

    def __del__(self):
        try:
            lib().lo_model_destroy(self.h)
        except Exception:
            pass

    def build_tables(self, maxE):
        L = lib()
        L.lo_build_tables(self.h, float(maxE))
        nE = L.lo_table_size(self.h); P = self.P
        arr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy()
        return dict(nE=nE, dE=L.lo_table_step(self.h), sigma=arr(L.lo_table_sigma(self.h), nE * P).reshape(nE, P),
                    cum=arr(L.lo_table_cum(self.h), nE * P).reshape(nE, P), nu_tot=arr(L.lo_table_nu_tot(self.h), nE),
                    nu_max=arr(L.lo_table_nu_max(self.h), nE))

    def max_accel_energy(self, e0, dt):
        return lib().lo_max_accel_energy(self.h, float(e0), float(dt))

    def event(self, nu_trial, t_sync, state10, draws):
        st = np.array(state10, dtype=np.float64)
        d = np.ascontiguousarray(draws, dtype=np.float64)
        out = np.zeros(10); used = C.c_int(0)
        chosen = lib().lo_event_injected(self.h, float(nu_trial), float(t_sync), _dp(st), _dp(d), len(d), _dp(out), C.byref(used))
        return chosen, st, out, used.value

    def solve(self, n, seed, n_integration_points, n_ss_times=0.0, sync_factor=1.0, temp_ratio=0.01, max_intervals=0):
        ctrl = np.array([n_integration_points, n_ss_times, sync_factor, temp_ratio, max_intervals], dtype=np.float64)
        res = np.zeros(35)
        lib().lo_solve(self.h, int(n), int(seed), _dp(ctrl), _dp(res))
        return res


class Ensemble:
    def __init__(self, model, n, seed, id_offset=0):
        self.model = model; self.n = int(n)
        self.h = lib().lo_ensemble_create(model.h, self.n, int(seed), int(id_offset))

    def __del__(self):
        try:
            lib().lo_ensemble_destroy(self.h)
        except Exception:
            pass

    def init(self, temp_ratio=0.01):
        lib().lo_ensemble_init(self.h, float(temp_ratio))

    def set(self, soa8):
        a = np.ascontiguousarray(soa8, dtype=np.float64); assert a.shape == (8, self.n)
        lib().lo_ensemble_set(self.h, _dp(a))

    def get(self):
        a = np.zeros((8, self.n)); lib().lo_ensemble_get(self.h, _dp(a)); return a

    def max_energy(self):
        return lib().lo_ensemble_max_energy(self.h)

    def advance(self, nu_trial, t_sync, interval, population_control=0):
        P = self.model.P
        counters = np.zeros(4, dtype=np.uint64); counts = np.zeros(P, dtype=np.uint64)
        gain = np.zeros(P); loss = np.zeros(P); scal = np.zeros(2)
        lib().lo_ensemble_advance(self.h, float(nu_trial), float(t_sync), int(interval), int(population_control),
                                  counters.ctypes.data_as(c_up), counts.ctypes.data_as(c_up), _dp(gain), _dp(loss), _dp(scal))
        return dict(n_real=int(counters[0]), n_null=int(counters[1]), n_born=int(counters[2]), n_attached=int(counters[3]),
                    counts=counts, gain=gain, loss=loss, field=scal[0], growth=scal[1])


def moments(soa):
    a = np.ascontiguousarray(soa[:6], dtype=np.float64); n = a.shape[1]
    out = np.zeros(26)
    lib().lo_moments(n, *[_dp(a[i]) for i in range(6)], _dp(out))
    return out


def histograms(v3, max_eedf_energy, n_energy=1000, n_cos=100, n_radial=200, n_axial=200, cylindrical=True):
    a = np.ascontiguousarray(v3, dtype=np.float64); n = a.shape[1]
    eeh = np.zeros(n_energy); eah = np.zeros((n_energy, n_cos)); evh = np.zeros((n_radial, n_axial))
    lib().lo_histograms(n, _dp(a[0]), _dp(a[1]), _dp(a[2]), float(max_eedf_energy), n_energy, n_cos, n_radial, n_axial, int(cylindrical),
                        _dp(eeh), _dp(eah), _dp(evh))
    return eeh, eah, evh


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().lo_philox4x32_10(c, k, o)
    return list(o)


def stream_uniform(seed, eid, interval, j):
    return lib().lo_stream_uniform(int(seed), int(eid), int(interval), int(j))
