"""loki_mc_b200 -- B200-native engine for LoKI-MC's electron Monte Carlo hot path.

The product is liblokib200.so (CUDA kernels + C ABI, include/lokib200.h).  This package is the thin Python binding the tests
and bench.py use; it never falls back to a CPU implementation: importing works without a GPU, creating an Engine does not.
"""
from ._capi import Engine, Job, Setup, Report, Output, SolveResults, run_setup, eval_expression, eval_vector_expression, LokiB200Error, build, lib, lib_path, result_len, R, comm_unique_id, comm_init_all, allreduce_results  # noqa: F401
