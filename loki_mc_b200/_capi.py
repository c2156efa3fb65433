"""ctypes binding of include/lokib200.h (liblokib200.so).  No compute happens here."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc")
ND = -123456789.0

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


class LokiB200Error(RuntimeError):
    pass


class R:
    """indices of the result vector (LOKIB200_R_* in include/lokib200.h)"""
    N_REAL, N_NULL, N_BORN, N_ATTACHED, GAIN_FIELD, GROWTH, SUM_EPS, SUM_R, SUM_V, SUM_RR, SUM_RV = 0, 1, 2, 3, 4, 5, 6, 7, 10, 13, 22
    N_SAMPLED, N_TABLE_CLAMPED, N_NU_EXCEEDED, SUM_COUNT, MAX_EPS, MAX_EPS_SEEN, OVERFLOW, HEADER = 31, 32, 33, 34, 34, 35, 36, 37


def result_len(P):
    return R.HEADER + 3 * P


def lib_path():
    # LOKIB200_LIB selects an experimental build variant (tools/build_variants.py); the default is the product library
    return os.environ.get("LOKIB200_LIB") or os.path.join(PKG, "liblokib200.so")


HOST_SOURCES = ("boltzmann_mc.cpp", "setup_input.cpp", "report.cpp", "host_capi.cpp", "run.cpp")
HOST_HEADERS = ("setup_input.h", "report.h")


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA engine (csrc/lokib200.cu, sm_100a) and the host sources into loki_mc_b200/liblokib200.so (in-tree, so it
    travels to the GPU box).  Objects are cached under loki_mc_b200/build/ and rebuilt when a source or header is newer."""
    out = lib_path()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    public = [os.path.join(ROOT, "include", f) for f in ("lokib200.h", "lokib200_host.h")]
    cu = os.path.join(SRC, "lokib200.cu")
    cu_deps = [cu] + [os.path.join(SRC, f) for f in ("lk_stream.cuh", "lk_tile.cuh", "lk_kernels.cuh", "lk_physics.cuh")] + public
    host_deps = [os.path.join(PKG, "host", f) for f in HOST_HEADERS] + public
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    all_sources = cu_deps + host_deps + [os.path.join(PKG, "host", f) for f in HOST_SOURCES + ("lokimc_main.cpp",)]
    if not force and not _stale(out, all_sources) and os.path.exists(os.path.join(PKG, "lokimc_b200")):
        return out   # the object cache does not travel to the GPU box; a library newer than every source is up to date
    objs = []
    jobs = [(cu, os.path.join(objdir, "lokib200.o"), cu_deps)] + \
           [(os.path.join(PKG, "host", f), os.path.join(objdir, f[:-4] + ".o"), [os.path.join(PKG, "host", f)] + host_deps) for f in HOST_SOURCES]
    relinked = False
    for src, obj, deps in jobs:
        if force or _stale(obj, deps):
            subprocess.check_call([nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose and src == cu else []) + ["-c", "-o", obj, src])
            relinked = True
        objs.append(obj)
    if relinked or _stale(out, objs):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
    exe, main_src = os.path.join(PKG, "lokimc_b200"), os.path.join(PKG, "host", "lokimc_main.cpp")
    if out == os.path.join(PKG, "liblokib200.so") and (force or _stale(exe, [main_src, out] + public)):   # the command-line front end
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, main_src, "-L" + PKG, "-llokib200", "-Wl,-rpath,$ORIGIN"])
    return out


class Config(C.Structure):
    _fields_ = [("n_electrons", C.c_int64), ("seed", C.c_uint64), ("first_electron_id", C.c_uint64), ("device", C.c_int32),
                ("gas_temperature_effect", C.c_int32), ("ionization_sharing", C.c_int32), ("is_cylindrically_symmetric", C.c_int32),
                ("energy_sharing_factor", C.c_double), ("gas_density", C.c_double), ("gas_temperature", C.c_double),
                ("electric_field", C.c_double * 3), ("excitation_omega", C.c_double), ("cyclotron_omega", C.c_double),
                ("n_interp_points", C.c_int32), ("n_energy_cells", C.c_int32), ("n_cos_cells", C.c_int32), ("n_radial_cells", C.c_int32),
                ("n_axial_cells", C.c_int32), ("n_phases", C.c_int32), ("reserved", C.c_int32)]


c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)


class ProcessSoA(C.Structure):
    _fields_ = [("n_processes", C.c_int32), ("n_gases", C.c_int32), ("type", c_ip), ("is_superelastic", c_ip), ("angular_model", c_ip),
                ("angular_p0", c_dp), ("angular_p1", c_dp), ("superelastic_weight_factor", c_dp), ("energy_min", c_dp), ("energy_max", c_dp),
                ("rel_density", c_dp), ("target_mass", c_dp), ("reduced_mass", c_dp), ("energy_loss", c_dp), ("thermal_std", c_dp),
                ("w_parameter", c_dp), ("gas_first", c_ip), ("gas_last", c_ip), ("gas_fraction", c_dp), ("xs_offset", c_lp),
                ("xs_energy", c_dp), ("xs_value", c_dp)]


class SolveControls(C.Structure):
    _fields_ = [("n_integration_points", C.c_double), ("n_integrated_ss_times", C.c_double), ("integrated_absolute_time", C.c_double),
                ("errors_to_be_checked", C.c_int32), ("sync_over_sampling", C.c_int32),
                ("rel_err_mean_energy", C.c_double), ("rel_err_flux_drift", C.c_double), ("rel_err_flux_diff", C.c_double),
                ("rel_err_bulk_drift", C.c_double), ("rel_err_bulk_diff", C.c_double), ("rel_err_power_balance", C.c_double),
                ("min_collisions_before_ss", C.c_double), ("max_collisions_before_ss", C.c_double), ("max_collisions_after_ss", C.c_double),
                ("sync_factor", C.c_double), ("initial_temp_ratio", C.c_double), ("energy_max_elastic", C.c_double), ("max_intervals", C.c_int64),
                ("status_display", C.c_int32), ("fast_mode", C.c_int32), ("status_values", C.c_double * 4)]


class SolveResults(C.Structure):
    _fields_ = [("averaged_mean_energy", C.c_double), ("averaged_mean_energy_error", C.c_double),
                ("flux_drift_velocity", C.c_double * 3), ("flux_drift_velocity_error", C.c_double * 3),
                ("flux_diffusion", C.c_double * 9), ("flux_diffusion_error", C.c_double * 9),
                ("bulk_drift_velocity", C.c_double * 3), ("bulk_drift_velocity_error", C.c_double * 3),
                ("bulk_diffusion", C.c_double * 9), ("bulk_diffusion_error", C.c_double * 9),
                ("power_gain_field", C.c_double), ("power_growth", C.c_double), ("power_balance_rel_error", C.c_double),
                ("time", C.c_double), ("steady_state_time", C.c_double), ("total_integrated_time", C.c_double),
                ("trial_collision_frequency", C.c_double), ("max_eedf_energy", C.c_double), ("elapsed_seconds", C.c_double),
                ("total_collisions", C.c_double), ("null_collisions", C.c_double), ("collisions_at_ss", C.c_double), ("null_collisions_at_ss", C.c_double),
                ("n_sampling_points", C.c_int64), ("n_integration_points", C.c_int64), ("n_sync_points", C.c_int64), ("n_table_rebuilds", C.c_int64),
                ("good_statistical_errors", C.c_int32), ("stopped_by_max_collisions", C.c_int32),
                ("n_nu_exceeded", C.c_double), ("n_table_clamped", C.c_double), ("events_per_second", C.c_double)]


class JobData(C.Structure):
    _fields_ = [("results", C.POINTER(SolveResults)), ("n_electrons", C.c_double), ("evdf_max_speed", C.c_double),
                ("rate_coeffs", c_dp), ("power_gain", c_dp), ("power_loss", c_dp), ("counts", c_dp),
                ("eeh", c_dp), ("eah", c_dp), ("evh", c_dp), ("eeh_periodic", c_dp), ("n_samples", C.c_int64),
                ("times", c_dp), ("mean_energy", c_dp), ("mean_pos", c_dp), ("mean_vel", c_dp), ("pos_cov", c_dp),
                ("points_per_phase", c_dp), ("mean_energy_periodic", c_dp), ("flux_velocity_periodic", c_dp), ("bulk_velocity_periodic", c_dp),
                ("flux_diffusion_periodic", c_dp), ("bulk_diffusion_periodic", c_dp)]


class RunSummary(C.Structure):
    _fields_ = [("n_jobs", C.c_int32), ("last_mean_energy", C.c_double), ("total_collisions", C.c_double), ("device_seconds", C.c_double),
                ("elapsed_seconds", C.c_double)]


ELECTRON_DTYPE = np.dtype([("r", "f8", 3), ("v", "f8", 3), ("energy", "f8"), ("t", "f8"), ("t_cf", "f8"), ("nu_e", "f8")])
EVENT_DTYPE = np.dtype([("chosen", "i4"), ("draws_used", "i4"), ("dE", "f8"), ("dE_rel", "f8"), ("gain_field", "f8"), ("ej_r", "f8", 3),
                        ("ej_v", "f8", 3), ("ej_energy", "f8")])

_LIB = None
# every symbol include/lokib200.h declares (tests check they are all exported)
SYMBOLS = ["lokib200_abi_version", "lokib200_device_count", "lokib200_create", "lokib200_destroy", "lokib200_last_error", "lokib200_set_stream",
           "lokib200_set_processes", "lokib200_build_tables", "lokib200_upload_tables", "lokib200_get_tables", "lokib200_table_info",
           "lokib200_nu_max_at", "lokib200_init_ensemble", "lokib200_set_ensemble", "lokib200_get_ensemble", "lokib200_time",
           "lokib200_advance_to_sync", "lokib200_advance_to_sync_device", "lokib200_set_histogram_grid", "lokib200_sample_histograms",
           "lokib200_fetch_histograms", "lokib200_step_injected", "lokib200_max_accel_energy", "lokib200_check_nu_trial",
           "lokib200_launch_count", "lokib200_kernel_time_ms", "lokib200_measure_fp64_peak", "lokib200_get_config", "lokib200_process_count", "lokib200_get_rel_densities",
           "lokib200_sample_moments", "lokib200_regrid_energy_histograms", "lokib200_read_result", "lokib200_job_create", "lokib200_job_solve", "lokib200_job_results",
           "lokib200_job_process_outputs", "lokib200_job_time_series", "lokib200_job_histograms", "lokib200_job_periodic",
           "lokib200_job_periodic_diffusion", "lokib200_job_conditions", "lokib200_job_evdf_max_speed", "lokib200_job_last_error", "lokib200_job_destroy",
           "lokib200_sample_moments_device", "lokib200_comm_unique_id", "lokib200_comm_init_rank", "lokib200_comm_init_all", "lokib200_comm_destroy",
           "lokib200_comm_size", "lokib200_comm_transport", "lokib200_comm_allreduce_results", "lokib200_comm_allreduce_histograms", "lokib200_kernel_form", "lokib200_device_hbm_gbs", "lokib200_set_fast_mode"]
# every symbol include/lokib200_host.h declares
HOST_SYMBOLS = ["lokib200_setup_load", "lokib200_setup_destroy", "lokib200_setup_last_error", "lokib200_setup_job_count", "lokib200_setup_job_value",
                "lokib200_setup_variable_condition", "lokib200_setup_processes", "lokib200_setup_config", "lokib200_setup_controls",
                "lokib200_setup_process_description", "lokib200_setup_process_is_elastic", "lokib200_setup_energy_max_elastic",
                "lokib200_setup_value", "lokib200_setup_dump", "lokib200_setup_warning_count", "lokib200_setup_warning",
                "lokib200_eval_expression", "lokib200_eval_vector_expression", "lokib200_report_create", "lokib200_report_from_job",
                "lokib200_report_destroy", "lokib200_report_last_error", "lokib200_report_swarm", "lokib200_report_power", "lokib200_report_energy_cells",
                "lokib200_report_eedf", "lokib200_report_rate_count", "lokib200_report_rate", "lokib200_output_create", "lokib200_output_write",
                "lokib200_output_folder", "lokib200_output_last_error", "lokib200_output_destroy", "lokib200_run_setup", "lokib200_run_last_error"]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LokiB200Error("liblokib200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.lokib200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.lokib200_destroy.argtypes = [vp]; L.lokib200_destroy.restype = None
    L.lokib200_last_error.argtypes = [vp]; L.lokib200_last_error.restype = C.c_char_p
    L.lokib200_set_stream.argtypes = [vp, vp]
    L.lokib200_set_processes.argtypes = [vp, C.POINTER(ProcessSoA)]
    L.lokib200_build_tables.argtypes = [vp, C.c_double]
    L.lokib200_upload_tables.argtypes = [vp, c_dp, c_dp, c_dp, C.c_int32, C.c_double]
    L.lokib200_get_tables.argtypes = [vp, c_dp, c_dp, c_dp]
    L.lokib200_table_info.argtypes = [vp, c_ip, c_dp, c_dp, c_dp]
    L.lokib200_nu_max_at.argtypes = [vp, C.c_int32]; L.lokib200_nu_max_at.restype = C.c_double
    L.lokib200_init_ensemble.argtypes = [vp, C.c_double, c_dp]
    L.lokib200_set_ensemble.argtypes = [vp, c_dp, C.c_double]
    L.lokib200_get_ensemble.argtypes = [vp, c_dp]
    L.lokib200_time.argtypes = [vp]; L.lokib200_time.restype = C.c_double
    L.lokib200_advance_to_sync.argtypes = [vp, C.c_double, C.c_double, C.c_int32, c_dp]
    L.lokib200_advance_to_sync_device.argtypes = [vp, C.c_double, C.c_double, C.c_int32, vp]
    L.lokib200_set_histogram_grid.argtypes = [vp, C.c_double]
    L.lokib200_sample_histograms.argtypes = [vp, C.c_int32]
    L.lokib200_fetch_histograms.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.lokib200_step_injected.argtypes = [vp, C.c_int32, vp, C.c_double, c_dp, c_dp, C.c_int32, vp, vp]
    L.lokib200_max_accel_energy.argtypes = [vp, C.c_double, C.c_double]; L.lokib200_max_accel_energy.restype = C.c_double
    L.lokib200_check_nu_trial.argtypes = [vp, C.c_double, C.c_double, C.c_double, c_dp]
    L.lokib200_launch_count.argtypes = [vp]; L.lokib200_launch_count.restype = C.c_int64
    L.lokib200_kernel_form.argtypes = [vp]; L.lokib200_kernel_form.restype = C.c_int32
    L.lokib200_device_hbm_gbs.argtypes = [vp]; L.lokib200_device_hbm_gbs.restype = C.c_double
    L.lokib200_set_fast_mode.argtypes = [vp, C.c_int32]
    L.lokib200_kernel_time_ms.argtypes = [vp, c_dp, c_lp]
    L.lokib200_measure_fp64_peak.argtypes = [vp, c_dp]
    L.lokib200_sample_moments.argtypes = [vp, c_dp]
    L.lokib200_sample_moments_device.argtypes = [vp]
    L.lokib200_comm_unique_id.argtypes = [vp]
    L.lokib200_comm_init_rank.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.lokib200_comm_init_all.argtypes = [C.POINTER(vp), C.c_int32]
    L.lokib200_comm_destroy.argtypes = [vp]
    L.lokib200_comm_size.argtypes = [vp]; L.lokib200_comm_size.restype = C.c_int32
    L.lokib200_comm_transport.argtypes = [vp]; L.lokib200_comm_transport.restype = C.c_char_p
    L.lokib200_comm_allreduce_results.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(vp)]
    L.lokib200_comm_allreduce_histograms.argtypes = [C.POINTER(vp), C.c_int32]
    L.lokib200_regrid_energy_histograms.argtypes = [vp, C.c_double]
    L.lokib200_read_result.argtypes = [vp, c_dp]
    L.lokib200_get_config.argtypes = [vp, C.POINTER(Config)]
    L.lokib200_process_count.argtypes = [vp]
    L.lokib200_get_rel_densities.argtypes = [vp, c_dp]
    L.lokib200_job_create.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(SolveControls), C.POINTER(vp)]
    L.lokib200_job_solve.argtypes = [vp, C.POINTER(SolveResults)]
    L.lokib200_job_results.argtypes = [vp, C.POINTER(SolveResults)]
    L.lokib200_job_process_outputs.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.lokib200_job_time_series.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp]; L.lokib200_job_time_series.restype = C.c_int64
    L.lokib200_job_histograms.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.lokib200_job_periodic.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.lokib200_job_periodic_diffusion.argtypes = [vp, c_dp, c_dp]
    L.lokib200_job_conditions.argtypes = [vp, C.POINTER(Config), c_ip]
    L.lokib200_job_evdf_max_speed.argtypes = [vp]; L.lokib200_job_evdf_max_speed.restype = C.c_double
    L.lokib200_job_last_error.argtypes = [vp]; L.lokib200_job_last_error.restype = C.c_char_p
    L.lokib200_job_destroy.argtypes = [vp]; L.lokib200_job_destroy.restype = None
    _bind_host(L)
    _LIB = L
    return L


def _bind_host(L):
    vp = C.c_void_p
    L.lokib200_setup_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.lokib200_setup_destroy.argtypes = [vp]; L.lokib200_setup_destroy.restype = None
    L.lokib200_setup_last_error.argtypes = [vp]; L.lokib200_setup_last_error.restype = C.c_char_p
    L.lokib200_setup_job_count.argtypes = [vp]
    L.lokib200_setup_job_value.argtypes = [vp, C.c_int32]; L.lokib200_setup_job_value.restype = C.c_double
    L.lokib200_setup_variable_condition.argtypes = [vp]; L.lokib200_setup_variable_condition.restype = C.c_char_p
    L.lokib200_setup_processes.argtypes = [vp, C.POINTER(ProcessSoA)]
    L.lokib200_setup_config.argtypes = [vp, C.c_int32, C.POINTER(Config)]
    L.lokib200_setup_controls.argtypes = [vp, C.POINTER(SolveControls)]
    L.lokib200_setup_process_description.argtypes = [vp, C.c_int32]; L.lokib200_setup_process_description.restype = C.c_char_p
    L.lokib200_setup_process_is_elastic.argtypes = [vp, C.c_int32]
    L.lokib200_setup_energy_max_elastic.argtypes = [vp]; L.lokib200_setup_energy_max_elastic.restype = C.c_double
    L.lokib200_setup_value.argtypes = [vp, C.c_char_p]; L.lokib200_setup_value.restype = C.c_char_p
    L.lokib200_setup_dump.argtypes = [vp, C.c_char_p, C.c_int64]; L.lokib200_setup_dump.restype = C.c_int64
    L.lokib200_setup_warning_count.argtypes = [vp]
    L.lokib200_setup_warning.argtypes = [vp, C.c_int32]; L.lokib200_setup_warning.restype = C.c_char_p
    L.lokib200_eval_expression.argtypes = [C.c_char_p, c_ip]; L.lokib200_eval_expression.restype = C.c_double
    L.lokib200_eval_vector_expression.argtypes = [C.c_char_p, c_dp, C.c_int64, c_ip]; L.lokib200_eval_vector_expression.restype = C.c_int64
    L.lokib200_report_create.argtypes = [vp, C.c_int32, C.POINTER(JobData), C.POINTER(vp)]
    L.lokib200_report_from_job.argtypes = [vp, C.c_int32, vp, C.POINTER(vp)]
    L.lokib200_report_destroy.argtypes = [vp]; L.lokib200_report_destroy.restype = None
    L.lokib200_report_last_error.argtypes = [vp]; L.lokib200_report_last_error.restype = C.c_char_p
    L.lokib200_report_swarm.argtypes = [vp, C.c_char_p, c_ip]; L.lokib200_report_swarm.restype = C.c_double
    L.lokib200_report_power.argtypes = [vp, C.c_char_p, C.c_char_p, c_ip]; L.lokib200_report_power.restype = C.c_double
    L.lokib200_report_energy_cells.argtypes = [vp]
    L.lokib200_report_eedf.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.lokib200_report_rate_count.argtypes = [vp, C.c_int32]
    L.lokib200_report_rate.argtypes = [vp, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(C.c_char_p)]
    L.lokib200_output_create.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.lokib200_output_write.argtypes = [vp, vp]
    L.lokib200_output_folder.argtypes = [vp]; L.lokib200_output_folder.restype = C.c_char_p
    L.lokib200_output_last_error.argtypes = [vp]; L.lokib200_output_last_error.restype = C.c_char_p
    L.lokib200_output_destroy.argtypes = [vp]; L.lokib200_output_destroy.restype = None
    L.lokib200_run_setup.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(RunSummary)]
    L.lokib200_run_last_error.argtypes = []; L.lokib200_run_last_error.restype = C.c_char_p


def _dp(a):
    return a.ctypes.data_as(c_dp)


class Engine:
    """One job on one GPU.  `model` is a dict with the process SoA ('p_*', 'gas_*', 'xs_*') and 'cond' (job conditions),
    the same layout tests/golden_io.load returns."""

    def __init__(self, model, n_electrons, seed=0x4C6F4B49, device=0, first_electron_id=0, n_phases=100, cells=(1000, 100, 200, 200)):
        L = lib()
        self.L = L
        c = model["cond"]
        cfg = Config()
        cfg.n_electrons = int(n_electrons); cfg.seed = int(seed); cfg.first_electron_id = int(first_electron_id); cfg.device = int(device)
        cfg.gas_temperature_effect = int(c["gas_temperature_effect"]); cfg.ionization_sharing = int(c["ionization_sharing"])
        cfg.is_cylindrically_symmetric = int(c.get("is_cylindrically_symmetric", 1))
        cfg.energy_sharing_factor = float(c["energy_sharing_factor"]); cfg.gas_density = float(c["gas_density"])
        cfg.gas_temperature = float(c["gas_temperature"])
        for i in range(3):
            cfg.electric_field[i] = float(c["electric_field"][i])
        cfg.excitation_omega = float(c["excitation_omega"]); cfg.cyclotron_omega = float(c["cyclotron_omega"])
        cfg.n_interp_points = int(c.get("n_interp_points", 10000))
        cfg.n_energy_cells, cfg.n_cos_cells, cfg.n_radial_cells, cfg.n_axial_cells = (int(x) for x in cells)
        cfg.n_phases = int(n_phases)
        self.cfg = cfg
        self.n = int(n_electrons)
        self.energy_max_elastic = float(c.get("energy_max_elastic", 1e100))
        h = C.c_void_p()
        rc = L.lokib200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise LokiB200Error("lokib200_create failed (%d): %s" % (rc, L.lokib200_last_error(None).decode()))
        self.h = h
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.P = int(len(model["p_type"])); self.nG = int(len(model["gas_first"]))
        k = dict(type=i32(model["p_type"]), sup=i32(model["p_superelastic"]), ang=i32(model["p_angular"]), ap0=f64(model["p_ap0"]),
                 ap1=f64(model["p_ap1"]), swf=f64(model["p_swf"]), emin=f64(model["p_emin"]), emax=f64(model["p_emax"]),
                 rd=f64(model["p_reldens"]), mass=f64(model["p_mass"]), mu=f64(model["p_redmass"]), el=f64(model["p_eloss"]),
                 th=f64(model["p_thstd"]), w=f64(model["p_w"]), gf=i32(model["gas_first"]), gl=i32(model["gas_last"]),
                 gfr=f64(model["gas_fraction"]), xo=np.ascontiguousarray(model["xs_offset"], dtype=np.int64), xe=f64(model["xs_energy"]),
                 xv=f64(model["xs_value"]))
        self._keep = k
        ip = lambda a: a.ctypes.data_as(c_ip)
        soa = ProcessSoA(self.P, self.nG, ip(k["type"]), ip(k["sup"]), ip(k["ang"]), _dp(k["ap0"]), _dp(k["ap1"]), _dp(k["swf"]), _dp(k["emin"]),
                         _dp(k["emax"]), _dp(k["rd"]), _dp(k["mass"]), _dp(k["mu"]), _dp(k["el"]), _dp(k["th"]), _dp(k["w"]), ip(k["gf"]),
                         ip(k["gl"]), _dp(k["gfr"]), k["xo"].ctypes.data_as(c_lp), _dp(k["xe"]), _dp(k["xv"]))
        self._check(L.lokib200_set_processes(self.h, C.byref(soa)))

    def _check(self, rc):
        if rc != 0:
            raise LokiB200Error("lokib200 error %d: %s" % (rc, self.L.lokib200_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.lokib200_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- tables ---
    def build_tables(self, max_energy):
        self._check(self.L.lokib200_build_tables(self.h, float(max_energy)))

    def upload_tables(self, cum, nu_tot, nu_max, dE):
        cum = np.ascontiguousarray(cum, dtype=np.float64); nE = cum.shape[0]
        self._check(self.L.lokib200_upload_tables(self.h, _dp(cum), _dp(np.ascontiguousarray(nu_tot)), _dp(np.ascontiguousarray(nu_max)), nE, float(dE)))

    def table_info(self):
        nE = C.c_int32(); dE = C.c_double(); mE = C.c_double(); nm = C.c_double()
        self._check(self.L.lokib200_table_info(self.h, C.byref(nE), C.byref(dE), C.byref(mE), C.byref(nm)))
        return dict(nE=nE.value, dE=dE.value, max_energy=mE.value, nu_max_last=nm.value)

    def get_tables(self):
        nE = self.table_info()["nE"]
        cum = np.zeros((nE, self.P)); nt = np.zeros(nE); nm = np.zeros(nE)
        self._check(self.L.lokib200_get_tables(self.h, _dp(cum), _dp(nt), _dp(nm)))
        return cum, nt, nm

    # --- ensemble ---
    def init_ensemble(self, temp_ratio=0.01):
        mx = C.c_double()
        self._check(self.L.lokib200_init_ensemble(self.h, float(temp_ratio), C.byref(mx)))
        return mx.value

    def set_ensemble(self, soa8, time=0.0):
        a = np.ascontiguousarray(soa8, dtype=np.float64); assert a.shape == (8, self.n)
        self._check(self.L.lokib200_set_ensemble(self.h, _dp(a), float(time)))

    def get_ensemble(self):
        a = np.zeros((8, self.n))
        self._check(self.L.lokib200_get_ensemble(self.h, _dp(a)))
        return a

    @property
    def time(self):
        return self.L.lokib200_time(self.h)

    # --- hot path ---
    def advance(self, nu_trial, t_sync, sample=True):
        res = np.zeros(result_len(self.P))
        self._check(self.L.lokib200_advance_to_sync(self.h, float(nu_trial), float(t_sync), int(bool(sample)), _dp(res)))
        return res

    def advance_device(self, nu_trial, t_sync, sample, d_result_ptr):
        self._check(self.L.lokib200_advance_to_sync_device(self.h, float(nu_trial), float(t_sync), int(bool(sample)), C.c_void_p(d_result_ptr)))

    def set_stream(self, stream_ptr):
        self._check(self.L.lokib200_set_stream(self.h, C.c_void_p(stream_ptr)))

    def read_result(self):
        res = np.zeros(result_len(self.P))
        self._check(self.L.lokib200_read_result(self.h, _dp(res)))
        return res

    def sample_moments(self):
        res = np.zeros(result_len(self.P))
        self._check(self.L.lokib200_sample_moments(self.h, _dp(res)))
        return res

    # --- multi-GPU: shards of one job share a communicator (include/lokib200.h, "multi-GPU") ---
    def comm_init_rank(self, unique_id, rank, n_ranks):
        """one engine per process: `unique_id` = the 128 bytes of comm_unique_id() of rank 0, broadcast by the launcher"""
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._check(self.L.lokib200_comm_init_rank(self.h, C.cast(buf, C.c_void_p), int(rank), int(n_ranks)))

    def comm_size(self):
        return int(self.L.lokib200_comm_size(self.h))

    def comm_transport(self):
        """"none", "nccl" or "peer-memory" (lokib200_comm_transport)"""
        return self.L.lokib200_comm_transport(self.h).decode()

    def comm_destroy(self):
        self._check(self.L.lokib200_comm_destroy(self.h))

    def allreduce_results(self, d_result_ptr=None):
        """combine the result vector of the last advance_device / sample with all other ranks (asynchronous, on the engine's stream)"""
        arr = (C.c_void_p * 1)(self.h)
        ptrs = (C.c_void_p * 1)(C.c_void_p(d_result_ptr)) if d_result_ptr else None
        self._check(self.L.lokib200_comm_allreduce_results(arr, 1, ptrs))

    def step_injected(self, electrons, nu_trial, t_sync, draws):
        e = np.ascontiguousarray(electrons, dtype=ELECTRON_DTYPE); n = len(e)
        d = np.ascontiguousarray(draws, dtype=np.float64); assert d.shape[0] == n
        ts = np.ascontiguousarray(t_sync, dtype=np.float64)
        out = np.zeros(n, dtype=ELECTRON_DTYPE); ev = np.zeros(n, dtype=EVENT_DTYPE)
        self._check(self.L.lokib200_step_injected(self.h, n, e.ctypes.data_as(C.c_void_p), float(nu_trial), _dp(ts), _dp(d), d.shape[1],
                                                  out.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p)))
        return out, ev

    # --- histograms ---
    def set_histogram_grid(self, max_eedf_energy):
        self._check(self.L.lokib200_set_histogram_grid(self.h, float(max_eedf_energy)))

    def sample_histograms(self, phase_index=-1):
        self._check(self.L.lokib200_sample_histograms(self.h, int(phase_index)))

    def fetch_histograms(self, periodic=False):
        c = self.cfg
        eeh = np.zeros(c.n_energy_cells); eah = np.zeros((c.n_energy_cells, c.n_cos_cells)); evh = np.zeros((c.n_radial_cells, c.n_axial_cells))
        per = np.zeros((c.n_phases, c.n_energy_cells)) if periodic else None
        self._check(self.L.lokib200_fetch_histograms(self.h, _dp(eeh), _dp(eah), _dp(evh), _dp(per) if periodic else None))
        return (eeh, eah, evh, per) if periodic else (eeh, eah, evh)

    # --- scalar helpers ---
    def max_accel_energy(self, e0, dt):
        return self.L.lokib200_max_accel_energy(self.h, float(e0), float(dt))

    def check_nu_trial(self, max_energy, nu_trial, horizon=10.0):
        nu = C.c_double(float(nu_trial))
        self._check(self.L.lokib200_check_nu_trial(self.h, float(max_energy), float(horizon), self.energy_max_elastic, C.byref(nu)))
        return nu.value

    def set_fast_mode(self, on=True):
        """per-energy-band trial collision frequencies (not a reference feature; see include/lokib200.h)"""
        self._check(self.L.lokib200_set_fast_mode(self.h, int(bool(on))))

    def launch_count(self):
        return int(self.L.lokib200_launch_count(self.h))

    def kernel_form(self):
        """'k_advance' (one electron per thread) or 'k_advance_stream' (streaming pool): the advance kernel this engine launches"""
        return "k_advance_stream" if self.L.lokib200_kernel_form(self.h) == 1 else "k_advance"

    def kernel_time_ms(self):
        ms = C.c_double(); n = C.c_int64()
        self._check(self.L.lokib200_kernel_time_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self.L.lokib200_measure_fp64_peak(self.h, C.byref(v)))
        return v.value


COMM_ID_BYTES = 128


def comm_unique_id():
    buf = (C.c_char * COMM_ID_BYTES)()
    rc = lib().lokib200_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise LokiB200Error("lokib200_comm_unique_id failed (%d): %s" % (rc, lib().lokib200_last_error(None).decode()))
    return bytes(buf)


def comm_init_all(engines):
    """all engines live in this process, one per device: one communicator over them (ncclCommInitAll)"""
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rc = lib().lokib200_comm_init_all(arr, len(engines))
    if rc != 0:
        raise LokiB200Error("lokib200_comm_init_all failed (%d): %s" % (rc, lib().lokib200_last_error(engines[0].h).decode()))


def allreduce_results(engines):
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rc = lib().lokib200_comm_allreduce_results(arr, len(engines), None)
    if rc != 0:
        raise LokiB200Error("lokib200_comm_allreduce_results failed (%d): %s" % (rc, lib().lokib200_last_error(engines[0].h).decode()))


class Job:
    """One Monte Carlo job (BoltzmannMC::evaluateEEDF restated in loki_mc_b200/host/boltzmann_mc.cpp) over one or more engines."""

    def __init__(self, engines, n_integration_points=1000, n_integrated_ss_times=0.0, sync_factor=1.0, initial_temp_ratio=0.01,
                 max_intervals=0, **kw):
        self.L = lib()
        self.engines = list(engines)
        c = SolveControls()
        c.n_integration_points = float(n_integration_points); c.n_integrated_ss_times = float(n_integrated_ss_times)
        c.sync_over_sampling = int(kw.get("sync_over_sampling", 1)); c.sync_factor = float(sync_factor)
        c.initial_temp_ratio = float(initial_temp_ratio); c.energy_max_elastic = float(self.engines[0].energy_max_elastic)
        c.max_intervals = int(max_intervals)
        c.status_display = int(kw.get("status_display", 0)); c.fast_mode = int(kw.get("fast_mode", 0))
        for k in ("min_collisions_before_ss", "max_collisions_before_ss", "max_collisions_after_ss"):
            if k in kw:
                setattr(c, k, float(kw[k]))
        for k in ("rel_err_mean_energy", "rel_err_flux_drift", "rel_err_flux_diff", "rel_err_bulk_drift", "rel_err_bulk_diff", "rel_err_power_balance"):
            if k in kw:
                setattr(c, k, float(kw[k])); c.errors_to_be_checked = 1
        arr = (C.c_void_p * len(self.engines))(*[e.h for e in self.engines])
        h = C.c_void_p()
        rc = self.L.lokib200_job_create(arr, len(self.engines), C.byref(c), C.byref(h))
        if rc != 0:
            raise LokiB200Error("lokib200_job_create failed (%d)" % rc)
        self.h = h
        self.P = self.engines[0].P

    def solve(self):
        r = SolveResults()
        rc = self.L.lokib200_job_solve(self.h, C.byref(r))
        if rc != 0:
            raise LokiB200Error("lokib200_job_solve failed (%d): %s" % (rc, self.L.lokib200_job_last_error(self.h).decode()))
        out = {}
        for name, _ in SolveResults._fields_:
            v = getattr(r, name)
            out[name] = np.array(list(v)) if hasattr(v, "__len__") else v
        return out

    def process_outputs(self):
        a = [np.zeros(self.P) for _ in range(4)]
        self.L.lokib200_job_process_outputs(self.h, *[_dp(x) for x in a])
        return dict(rate_coeffs=a[0], power_gain=a[1], power_loss=a[2], counts=a[3])

    def time_series(self):
        n = self.L.lokib200_job_time_series(self.h, None, None, None, None, None)
        t = np.zeros(n); me = np.zeros(n); mp = np.zeros((n, 3)); mv = np.zeros((n, 3)); pc = np.zeros((n, 9))
        self.L.lokib200_job_time_series(self.h, _dp(t), _dp(me), _dp(mp), _dp(mv), _dp(pc))
        return dict(times=t, mean_energy=me, mean_position=mp, mean_velocity=mv, position_covariance=pc)

    def periodic(self):
        """phase-resolved averages of an AC job (lokib200_job_periodic)"""
        nph = self.engines[0].cfg.n_phases
        pts = np.zeros(nph); me = np.zeros(nph); fv = np.zeros((nph, 3)); bv = np.zeros((nph, 3))
        self.L.lokib200_job_periodic(self.h, _dp(pts), _dp(me), _dp(fv), _dp(bv))
        return dict(points_per_phase=pts, mean_energy=me, flux_velocity=fv, bulk_velocity=bv)

    def histograms(self):
        c = self.engines[0].cfg
        eeh = np.zeros(c.n_energy_cells); eah = np.zeros((c.n_energy_cells, c.n_cos_cells)); evh = np.zeros((c.n_radial_cells, c.n_axial_cells))
        per = np.zeros((c.n_phases, c.n_energy_cells))
        rc = self.L.lokib200_job_histograms(self.h, _dp(eeh), _dp(eah), _dp(evh), _dp(per))
        if rc != 0:
            raise LokiB200Error("no histograms (steady state not reached)")
        return dict(eeh=eeh, eah=eah, evh=evh, eeh_periodic=per)

    def close(self):
        if getattr(self, "h", None):
            self.L.lokib200_job_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Setup:
    """A parsed LoKI-MC setup file (include/lokib200_host.h): flattened process set + per-job engine configuration.
    Host-only; works without a GPU."""

    def __init__(self, input_dir, setup_file):
        L = lib()
        self.L = L
        h = C.c_void_p()
        rc = L.lokib200_setup_load(os.fsencode(input_dir), os.fsencode(setup_file), C.byref(h))
        if rc != 0:
            raise LokiB200Error(L.lokib200_setup_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.lokib200_setup_destroy(self.h); self.h = None

    __del__ = close

    @property
    def n_jobs(self):
        return int(self.L.lokib200_setup_job_count(self.h))

    def job_value(self, job):
        return float(self.L.lokib200_setup_job_value(self.h, job))

    @property
    def variable_condition(self):
        return self.L.lokib200_setup_variable_condition(self.h).decode()

    def value(self, key):
        return self.L.lokib200_setup_value(self.h, key.encode()).decode()

    def dump(self):
        n = self.L.lokib200_setup_dump(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.lokib200_setup_dump(self.h, buf, n + 1)
        return buf.value.decode()

    @property
    def warnings(self):
        return [self.L.lokib200_setup_warning(self.h, i).decode() for i in range(self.L.lokib200_setup_warning_count(self.h))]

    def config(self, job=0):
        cfg = Config()
        if self.L.lokib200_setup_config(self.h, job, C.byref(cfg)) != 0:
            raise LokiB200Error(self.L.lokib200_setup_last_error(self.h).decode())
        return cfg

    def controls(self):
        c = SolveControls()
        if self.L.lokib200_setup_controls(self.h, C.byref(c)) != 0:
            raise LokiB200Error(self.L.lokib200_setup_last_error(self.h).decode())
        return c

    def processes(self):
        """the flattened process set as the dict layout of tests/golden_io (copies)"""
        p = ProcessSoA()
        self.L.lokib200_setup_processes(self.h, C.byref(p))
        P, G = p.n_processes, p.n_gases
        def arr(ptr, n):
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy()
        off = arr(p.xs_offset, P + 1)
        d = {"p_type": arr(p.type, P), "p_superelastic": arr(p.is_superelastic, P), "p_angular": arr(p.angular_model, P),
             "p_ap0": arr(p.angular_p0, P), "p_ap1": arr(p.angular_p1, P), "p_swf": arr(p.superelastic_weight_factor, P),
             "p_emin": arr(p.energy_min, P), "p_emax": arr(p.energy_max, P), "p_reldens": arr(p.rel_density, P), "p_mass": arr(p.target_mass, P),
             "p_redmass": arr(p.reduced_mass, P), "p_eloss": arr(p.energy_loss, P), "p_thstd": arr(p.thermal_std, P), "p_w": arr(p.w_parameter, P),
             "gas_first": arr(p.gas_first, G), "gas_last": arr(p.gas_last, G), "gas_fraction": arr(p.gas_fraction, G), "xs_offset": off,
             "xs_energy": arr(p.xs_energy, int(off[-1])), "xs_value": arr(p.xs_value, int(off[-1]))}
        d["p_elastic"] = np.array([self.L.lokib200_setup_process_is_elastic(self.h, k) for k in range(P)], dtype=np.int32)
        d["descriptions"] = [self.L.lokib200_setup_process_description(self.h, k).decode() for k in range(P)]
        d["energy_max_elastic"] = float(self.L.lokib200_setup_energy_max_elastic(self.h))
        return d


def eval_expression(expr):
    ok = C.c_int32(0)
    v = lib().lokib200_eval_expression(expr.encode(), C.byref(ok))
    if not ok.value:
        raise LokiB200Error(lib().lokib200_setup_last_error(None).decode())
    return float(v)


def eval_vector_expression(expr):
    ok = C.c_int32(0)
    L = lib()
    n = L.lokib200_eval_vector_expression(expr.encode(), None, 0, C.byref(ok))
    if not ok.value:
        raise LokiB200Error(L.lokib200_setup_last_error(None).decode())
    out = np.zeros(max(int(n), 1))
    L.lokib200_eval_vector_expression(expr.encode(), _dp(out), n, C.byref(ok))
    return out[:n]


class Report:
    """Post-processing of one finished job (include/lokib200_host.h lokib200_report_*): distributions, power balance, rate
    coefficients and swarm parameters.  Build it from a solved Job, or from raw arrays (`data`: dict with the JobData fields)."""

    def __init__(self, setup, job_index, job=None, data=None):
        L = lib()
        self.L = L
        self.setup = setup
        h = C.c_void_p()
        if job is not None:
            rc = L.lokib200_report_from_job(setup.h, int(job_index), job.h, C.byref(h))
        else:
            jd = JobData()
            self._keep = []
            for name, ctype in JobData._fields_:
                v = data.get(name)
                if name == "results":
                    self._keep.append(v); jd.results = C.pointer(v)
                elif ctype is c_dp:
                    if v is not None:
                        a = np.ascontiguousarray(v, dtype=np.float64); self._keep.append(a); setattr(jd, name, _dp(a))
                else:
                    setattr(jd, name, v)
            rc = L.lokib200_report_create(setup.h, int(job_index), C.byref(jd), C.byref(h))
        if rc != 0:
            raise LokiB200Error(L.lokib200_report_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.lokib200_report_destroy(self.h); self.h = None

    __del__ = close

    def swarm(self, name):
        ok = C.c_int32(0)
        v = self.L.lokib200_report_swarm(self.h, name.encode(), C.byref(ok))
        if not ok.value:
            raise KeyError(name)
        return float(v)

    def power(self, name, gas=None):
        ok = C.c_int32(0)
        v = self.L.lokib200_report_power(self.h, name.encode(), gas.encode() if gas else None, C.byref(ok))
        if not ok.value:
            raise KeyError(name)
        return float(v)

    def eedf(self):
        n = self.L.lokib200_report_energy_cells(self.h)
        a = [np.zeros(n) for _ in range(4)]
        self.L.lokib200_report_eedf(self.h, *[_dp(x) for x in a])
        return dict(energy=a[0], eedf=a[1], first_anisotropy=a[2], second_anisotropy=a[3])

    def rates(self, extra=False):
        out = []
        for i in range(self.L.lokib200_report_rate_count(self.h, int(extra))):
            cid = C.c_int32(); v = [C.c_double() for _ in range(4)]; d = C.c_char_p()
            self.L.lokib200_report_rate(self.h, int(extra), i, C.byref(cid), *[C.byref(x) for x in v], C.byref(d))
            out.append(dict(id=cid.value, ine=v[0].value, sup=v[1].value, ine_mc=v[2].value, sup_mc=v[3].value, description=(d.value or b"").decode()))
        return out


class Output:
    """The reference's output folder (Headers/Output.h): setup.txt at creation, the selected data files per written job."""

    def __init__(self, setup, output_root):
        L = lib()
        self.L = L
        h = C.c_void_p()
        if L.lokib200_output_create(setup.h, os.fsencode(output_root), C.byref(h)) != 0:
            raise LokiB200Error(L.lokib200_output_last_error(None).decode())
        self.h = h

    @property
    def folder(self):
        return self.L.lokib200_output_folder(self.h).decode()

    def write(self, report):
        if self.L.lokib200_output_write(self.h, report.h) != 0:
            raise LokiB200Error(self.L.lokib200_output_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.lokib200_output_destroy(self.h); self.h = None

    __del__ = close


def run_setup(input_dir, setup_file, output_root, n_devices=1, first_device=0, verbose=True):
    """The reference executable's main loop on the GPU engine (include/lokib200_host.h lokib200_run_setup)."""
    L = lib()
    summary = RunSummary()
    rc = L.lokib200_run_setup(os.fsencode(input_dir), os.fsencode(setup_file), os.fsencode(output_root), int(n_devices), int(first_device), int(bool(verbose)),
                              C.byref(summary))
    if rc != 0:
        raise LokiB200Error(L.lokib200_run_last_error().decode())
    return summary
