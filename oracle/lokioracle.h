/* oracle/lokioracle.h -- TEST INFRASTRUCTURE ONLY (never imported, linked or executed by the product path).
 *
 * Plain-C CPU restatement of LoKI-MC's electron Monte Carlo hot path, written from the algorithm in
 * /root/reference/Code/LoKI-MC/Sources/BoltzmannMC.C (cited per function in lokioracle.c).  It is the checker the
 * tests, __graft_entry__.smoke() and bench.py's cpu_baseline leg compare the CUDA path against.
 * PINNED: tests/test_oracle_golden.py checks every function here against the tests/golden npz files, which were produced by
 * the unmodified reference itself (oracle/harness.cpp + oracle/gen_golden.py).
 */
#ifndef LOKIORACLE_H
#define LOKIORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LO_NON_DEF (-123456789.0)      /* Constant.h:23 */
#define LO_NULL_COLLISION (-1)         /* GeneralDefinitions.h:39 */
#define LO_PARTIAL_FLIGHT (-2)         /* GeneralDefinitions.h:40 */

typedef struct lo_model lo_model;
typedef struct lo_ensemble lo_ensemble;

/* ---- model: process set (BoltzmannMC.C:29-271) + job constants (BoltzmannMC.C:428-489) ---- */
lo_model* lo_model_create(int n_processes, int n_gases,
                          const int32_t* type, const int32_t* is_superelastic, const int32_t* angular_model,
                          const double* angular_p0, const double* angular_p1, const double* superelastic_weight_factor,
                          const double* energy_min, const double* energy_max, const double* rel_density,
                          const double* target_mass, const double* reduced_mass, const double* energy_loss,
                          const double* thermal_std, const double* w_parameter,
                          const int32_t* gas_first, const int32_t* gas_last, const double* gas_fraction,
                          const int64_t* xs_offset, const double* xs_energy, const double* xs_value);
void lo_model_set_conditions(lo_model* m, int gas_temperature_effect, int ionization_sharing, double energy_sharing_factor,
                             double gas_density, double gas_temperature, const double electric_field[3],
                             double excitation_omega, double cyclotron_omega, int n_interp_points);
void lo_model_destroy(lo_model* m);

/* interpolateCrossSections (BoltzmannMC.C:561-615); tables owned by the model */
void lo_build_tables(lo_model* m, double max_energy);
int lo_table_size(const lo_model* m);
double lo_table_step(const lo_model* m);
const double* lo_table_sigma(const lo_model* m);   /* [nE][P] */
const double* lo_table_cum(const lo_model* m);     /* [nE][P] */
const double* lo_table_nu_tot(const lo_model* m);  /* [nE] */
const double* lo_table_nu_max(const lo_model* m);  /* [nE] */

/* maximizationAccelerationEnergy (BoltzmannMC.C:765-802) */
double lo_max_accel_energy(const lo_model* m, double initial_energy, double dt);

/* One pass of the per-electron loop body (BoltzmannMC.C:637-681) with injected draws.
 * state: [x y z vx vy vz eps t_e t_cf nu_e] in/out.  out: [dE_coll, dE_coll/eps_inc, dE_field, ej_x ej_y ej_z ej_vx ej_vy ej_vz ej_eps].
 * returns the chosen process id (>=0), LO_NULL_COLLISION or LO_PARTIAL_FLIGHT; *draws_used = number of uniforms consumed. */
int lo_event_injected(const lo_model* m, double nu_trial, double t_sync, double* state, const double* draws, int n_draws,
                      double* out, int* draws_used);

/* calculateMeanDataForSwarmParams (BoltzmannMC.C:1410-1454): out[26] = mean eps, max eps, <r>[3], <v>[3], cov(r,r)[9], cov(r,v)[9] */
void lo_moments(int64_t n, const double* x, const double* y, const double* z, const double* vx, const double* vy, const double* vz, double* out);
/* histogramCount / histogram2DCount (MathFunctions.C:61-81,106-127) on the grids of BoltzmannMC.C:1862-1883; arrays are += */
void lo_histograms(int64_t n, const double* vx, const double* vy, const double* vz, double max_eedf_energy,
                   int n_energy, int n_cos, int n_radial, int n_axial, int cylindrical, double* eeh, double* eah, double* evh);

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 known-answer vectors are checked in the tests) */
void lo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the uniform stream both the oracle and the CUDA path define: draw j of electron `id` in interval `interval` */
double lo_stream_uniform(uint64_t seed, uint64_t id, uint32_t interval, uint32_t j);

/* ---- ensemble with counter-based draws: reference semantics of electronDynamicsUntilSynchronization (BoltzmannMC.C:617-688)
 *      + nonParallelCollisionTasks (:1282-1408), OpenMP over electrons like the reference (:636) ---- */
lo_ensemble* lo_ensemble_create(const lo_model* m, int64_t n, uint64_t seed, uint64_t id_offset);
void lo_ensemble_destroy(lo_ensemble* e);
void lo_ensemble_init(lo_ensemble* e, double initial_temp_over_gas_temp);          /* BoltzmannMC.C:491-508 */
void lo_ensemble_set(lo_ensemble* e, const double* soa8);                          /* x y z vx vy vz t_cf nu_e, each [n] */
void lo_ensemble_get(const lo_ensemble* e, double* soa8);
double lo_ensemble_max_energy(const lo_ensemble* e);
/* advance every electron to t_sync with trial frequency nu_trial.  counters: [0] real, [1] null; per_process (may be NULL):
 * counts[P], gain[P], loss[P]; scal: [0] field gain, [1] growth.  interval = index used in the draw counter.
 * population_control: 0 -> births/attachments are not applied (returns their count in counters[2], [3]); 1 -> reference semantics. */
void lo_ensemble_advance(lo_ensemble* e, double nu_trial, double t_sync, uint32_t interval, int population_control,
                         uint64_t* counters, uint64_t* counts, double* gain, double* loss, double* scal);
double lo_ensemble_time(const lo_ensemble* e);

/* whole job: evaluateEEDF (BoltzmannMC.C:299-426) with the reference's nu_trial / table / steady-state / stop logic.
 * ctrl: [0] nIntegrationPoints [1] nIntegratedSSTimes [2] sync factor [3] initialElecTempOverGasTemp [4] max sync intervals (0 = none)
 * res:  [0] averaged mean energy [1] its error [2..4] flux drift v [5..7] error [8..16] flux D (3x3) [17..25] bulk D (3x3)
 *       [26..28] bulk drift v [29] real collisions [30] null collisions [31] steady-state time [32] total time [33] sync intervals
 *       [34] elapsed seconds in the MC loop */
void lo_solve(const lo_model* m, int64_t n, uint64_t seed, const double* ctrl, double* res);

#ifdef __cplusplus
}
#endif
#endif
