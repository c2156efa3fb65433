/* include/lokib200.h -- C ABI of the B200-native electron Monte Carlo engine (liblokib200.so).
 *
 * This is the drop-in boundary for LoKI-MC's data-parallel hot path.  The reference (IST-Lisbon/LoKI-MC v1.1.0) has no
 * plugin/FFI interface; the cut is made INSIDE BoltzmannMC::evaluateEEDF(), file Code/LoKI-MC/Sources/BoltzmannMC.C
 * ("BMC.C" below; "BMC.h" = Headers/BoltzmannMC.h).  Each entry point names the reference code it replaces.
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative lokib200_status and never
 * calls exit() (the reference's Message::error convention, Sources/Message.C:6-16, is applied by the host from
 * lokib200_last_error()).  The engine owns all device memory; the caller owns every pointer it passes and may free it on
 * return.  One engine = one job on one GPU; call from one host thread.  Units are SI except energies (eV), as in the reference.
 */
#ifndef LOKIB200_H
#define LOKIB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOKIB200_ABI_VERSION 2
#define LOKIB200_NON_DEF (-123456789.0)    /* Headers/Constant.h:23 */
#define LOKIB200_NULL_COLLISION (-1)       /* Headers/GeneralDefinitions.h:39 */
#define LOKIB200_PARTIAL_FLIGHT (-2)       /* Headers/GeneralDefinitions.h:40 */

typedef enum lokib200_status {
  LOKIB200_OK = 0,
  LOKIB200_ERR_INVALID = -1,      /* bad argument / call order */
  LOKIB200_ERR_CUDA = -2,         /* CUDA runtime failure (message in last_error) */
  LOKIB200_ERR_NO_DEVICE = -3,    /* no usable sm_100 device: there is NO CPU fallback */
  LOKIB200_ERR_OVERFLOW = -4,     /* birth list overflow inside one interval */
  LOKIB200_ERR_ENERGY_RANGE = -5  /* electron energy beyond the elastic cross-section range (BMC.C:1427-1430) */
} lokib200_status;

typedef struct lokib200_engine lokib200_engine;

/* Job constants: what BoltzmannMC's constructor (BMC.h:244-380) and evaluateNonConstantVariables (BMC.C:428-489) hold. */
typedef struct lokib200_config {
  int64_t n_electrons;              /* electrons of THIS shard (BMC.h:51) */
  uint64_t seed;                    /* Philox key; draw streams are keyed by the GLOBAL electron id */
  uint64_t first_electron_id;       /* global id of local electron 0 (shard offset); 0 on a single GPU */
  int32_t device;                   /* CUDA device ordinal */
  int32_t gas_temperature_effect;   /* 0 false, 1 true, 2 smartActivation (GeneralDefinitions.h:49-51) */
  int32_t ionization_sharing;       /* 0 equalSharing, 1 oneTakesAll, 2 usingSDCS, 3 randomUniform (GeneralDefinitions.h:43-46) */
  int32_t is_cylindrically_symmetric; /* WorkingConditions.h:164-172: EVDF / angular histograms only when E || z */
  double energy_sharing_factor;     /* BMC.h:53 */
  double gas_density;               /* totalGasDensity [m^-3] (BMC.C:436) */
  double gas_temperature;           /* [K] */
  double electric_field[3];         /* [V/m], BMC.C:465-481 (AC amplitude already multiplied by sqrt(2)) */
  double excitation_omega;          /* excitationFrequencyRadians [rad/s] (BMC.C:477) */
  double cyclotron_omega;           /* cyclotronFrequency [rad/s] (BMC.C:485) */
  int32_t n_interp_points;          /* interpolCrossSectionSize (BMC.h:104), default 10000 */
  int32_t n_energy_cells;           /* BMC.h:59  (1000) */
  int32_t n_cos_cells;              /* BMC.h:60  (100)  */
  int32_t n_radial_cells;           /* BMC.h:62  (200)  */
  int32_t n_axial_cells;            /* BMC.h:61  (200)  */
  int32_t n_phases;                 /* nIntegrationPhases (BMC.h:75), used when excitation_omega != 0 */
  int32_t reserved;
} lokib200_config;

/* Flattened process set: the per-process arrays BoltzmannMC::allocateEvaluateVariablesFirstTime fills (BMC.C:89-270).
 * Inelastic and superelastic directions are separate processes; a superelastic follows its inelastic (BMC.C:209-266). */
typedef struct lokib200_process_soa {
  int32_t n_processes;                       /* nProcesses (BMC.C:33-48) */
  int32_t n_gases;                           /* nGases */
  const int32_t* type;                       /* processTypes: 0 conservative, 1 ionization, 2 attachment */
  const int32_t* is_superelastic;            /* isSuperElastic */
  const int32_t* angular_model;              /* 0 isotropic 1 forward 2 bornDipole 3 surendra 4 coulombScreen 5 momentumConservationIonization
                                                (Headers/AngularScatteringFunctions.h:76-114) */
  const double* angular_p0;                  /* angularScatteringParams[k][0] (coulombScreen energy option) */
  const double* angular_p1;                  /* angularScatteringParams[k][1] (coulombScreen screening energy [eV]) */
  const double* superelastic_weight_factor;  /* superElasticStatWeightFactors */
  const double* energy_min;                  /* energyMinLimits [eV] */
  const double* energy_max;                  /* energyMaxLimits [eV] */
  const double* rel_density;                 /* relDensities (BMC.C:446-452) */
  const double* target_mass;                 /* targetMasses [kg] */
  const double* reduced_mass;                /* reducedMasses [kg] */
  const double* energy_loss;                 /* energyLosses [eV] (negative for superelastics, BMC.C:247) */
  const double* thermal_std;                 /* thermalStdDeviations sqrt(kB Tg / M) [m/s] */
  const double* w_parameter;                 /* wParameters [eV] (BMC.C:173-183) */
  const int32_t* gas_first;                  /* firstProcessIndexPerGas [n_gases] */
  const int32_t* gas_last;                   /* lastProcessIndexPerGas  [n_gases] */
  const double* gas_fraction;                /* gasFractions [n_gases] */
  const int64_t* xs_offset;                  /* [n_processes+1]: process k owns raw points [xs_offset[k], xs_offset[k+1]) */
  const double* xs_energy;                   /* crossSectionEnergies, concatenated [eV] */
  const double* xs_value;                    /* crossSectionValues, concatenated [m^2] */
} lokib200_process_soa;

/* What one advance returns to the host: everything nonParallelCollisionTasks (BMC.C:1282-1408) and
 * calculateMeanDataForSwarmParams (BMC.C:1410-1454) accumulate, as SUMS over this shard (so that shards combine by addition;
 * see INTEGRATION.md for the NCCL all-reduce layout).  Stored as doubles; integer-valued entries are exact below 2^53. */
enum {
  LOKIB200_R_N_REAL = 0,       /* totalCollisionCounter increment */
  LOKIB200_R_N_NULL = 1,       /* nullCollisionCounter increment */
  LOKIB200_R_N_BORN = 2,       /* ionization events (ejected electrons) */
  LOKIB200_R_N_ATTACHED = 3,   /* attachment events */
  LOKIB200_R_GAIN_FIELD = 4,   /* energyGainField increment [eV] */
  LOKIB200_R_GROWTH = 5,       /* energyGrowth increment [eV] */
  LOKIB200_R_SUM_EPS = 6,      /* sum of energies at t_sync [eV] */
  LOKIB200_R_SUM_R = 7,        /* 3: sum r */
  LOKIB200_R_SUM_V = 10,       /* 3: sum v */
  LOKIB200_R_SUM_RR = 13,      /* 9: sum r r^T (row-major xx xy xz yx ...) */
  LOKIB200_R_SUM_RV = 22,      /* 9: sum r v^T */
  LOKIB200_R_N_SAMPLED = 31,   /* electrons in the sums (= n_electrons) */
  LOKIB200_R_N_TABLE_CLAMPED = 32, /* collisions whose energy index hit the last table row (BMC.C:955,1038 clamp) */
  LOKIB200_R_N_NU_EXCEEDED = 33,   /* collisions that found nu_tot(eps) > nu_e (trial frequency too small) */
  LOKIB200_R_SUM_COUNT = 34,   /* --- entries below combine with MAX, not SUM --- */
  LOKIB200_R_MAX_EPS = 34,     /* max energy at t_sync [eV] (BMC.C:721, :1426) */
  LOKIB200_R_MAX_EPS_SEEN = 35,/* max energy seen at any collision point inside the interval */
  LOKIB200_R_OVERFLOW = 36,    /* != 0: a birth / death / pending list overflowed inside the interval (the sums are then wrong); MAX-combined,
                                  so the flag survives the all-reduce and every rank sees it */
  LOKIB200_R_HEADER = 37       /* followed by counts[P], gain[P], loss[P] (collisionCounters, energyGain/LossProcesses) : SUM */
};
#define LOKIB200_RESULT_LEN(P) (LOKIB200_R_HEADER + 3 * (P))

/* one electron for the injected-draw parity entry: the per-electron slice of BMC.h:148-153,133-134 */
typedef struct lokib200_electron {
  double r[3], v[3];
  double energy;    /* electronEnergies [eV] */
  double t;         /* electronTimes [s] */
  double t_cf;      /* collisionFreeTimes [s]; LOKIB200_NON_DEF = draw a new one */
  double nu_e;      /* trialCollisionFrequenciesEachElectron [1/s] */
} lokib200_electron;

/* scratch the parallel pass hands to the serial pass in the reference (BMC.h:226-231) */
typedef struct lokib200_event_out {
  int32_t chosen;   /* chosenProcessIDs */
  int32_t draws_used;
  double dE;        /* electronEnergyChanges [eV] */
  double dE_rel;    /* electronEnergyChangesOverIncidEnergies */
  double gain_field;/* energyGainsField [eV] */
  double ej_r[3], ej_v[3], ej_energy;   /* ejectedElectron{Positions,Velocities,Energies} */
} lokib200_event_out;

int lokib200_abi_version(void);
int lokib200_device_count(void);

/* --- life cycle (replaces: BoltzmannMC ctor allocations BMC.C:51-87; the reference never frees) --- */
int lokib200_create(const lokib200_config* cfg, lokib200_engine** out);
void lokib200_destroy(lokib200_engine* h);
const char* lokib200_last_error(const lokib200_engine* h);   /* h may be NULL: error of the last failed create */
/* run on an existing stream (cudaStream_t as void*), e.g. torch's current stream; NULL -> engine-owned stream */
int lokib200_set_stream(lokib200_engine* h, void* cuda_stream);

/* --- tables (replaces: allocateEvaluateVariablesFirstTime BMC.C:89-270 hand-off, interpolateCrossSections BMC.C:561-615) --- */
int lokib200_set_processes(lokib200_engine* h, const lokib200_process_soa* p);
/* read back what the engine holds (used by the host driver, which sees the engine only through this ABI) */
int lokib200_get_config(const lokib200_engine* h, lokib200_config* cfg);
int lokib200_process_count(const lokib200_engine* h);
int lokib200_get_rel_densities(const lokib200_engine* h, double* rel_density);
/* host flattening + upload: uniform grid of n_interp_points energies up to max_energy, sigma x relDensity, row cumsum, nu_tot, running max */
int lokib200_build_tables(lokib200_engine* h, double max_energy);
/* upload tables built elsewhere (e.g. dumped from the reference object); cum is row-major [nE][P] */
int lokib200_upload_tables(lokib200_engine* h, const double* cum, const double* nu_tot, const double* nu_max, int32_t nE, double dE);
int lokib200_get_tables(lokib200_engine* h, double* cum, double* nu_tot, double* nu_max);   /* device -> host, any may be NULL */
int lokib200_table_info(const lokib200_engine* h, int32_t* nE, double* dE, double* max_energy, double* nu_max_last);
double lokib200_nu_max_at(const lokib200_engine* h, int32_t index);   /* maxCollisionFrequencies[index] (BMC.C:755) */

/* --- ensemble state (replaces: evaluateNonConstantVariables BMC.C:491-516) --- */
int lokib200_init_ensemble(lokib200_engine* h, double initial_temp_over_gas_temp, double* max_energy);
/* soa8 = x,y,z,vx,vy,vz,t_cf,nu_e each [n_electrons] (host memory).  time = common clock of the ensemble. */
int lokib200_set_ensemble(lokib200_engine* h, const double* soa8, double time);
int lokib200_get_ensemble(lokib200_engine* h, double* soa8);
double lokib200_time(const lokib200_engine* h);

/* --- the hot path (replaces: electronDynamicsUntilSynchronization BMC.C:617-688 incl. accelerateElectron :804-905,
 *     performCollision :907-1113, conservative/ionization/attachmentCollision :1115-1280, nonParallelCollisionTasks :1282-1408,
 *     and, when `sample` != 0, the ensemble sums of calculateMeanDataForSwarmParams :1410-1454) ---
 * Advances every electron from the engine's time to t_sync with trial frequency nu_trial, applies birth/death population
 * control at t_sync, and writes LOKIB200_RESULT_LEN(P) doubles to `result` (host memory; may be NULL).
 * Blocking.  The whole interval (advance kernel ... copy of the vector into pinned memory, and the peer-memory exchange of a
 * communicator) is ONE CUDA graph launch, captured on the first call (LOKIB200_GRAPH=0: individual launches; identical results). */
int lokib200_advance_to_sync(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* result);
/* same, asynchronous: the result stays in device memory (`d_result`, >= LOKIB200_RESULT_LEN(P) doubles, caller-owned device
 * pointer, e.g. a torch tensor that is then all-reduced over NCCL); no host synchronisation */
int lokib200_advance_to_sync_device(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* d_result);
/* blocking read of the result of the last lokib200_advance_to_sync_device(..., d_result = NULL) call into host memory; returns
 * LOKIB200_ERR_OVERFLOW when the vector carries the overflow flag */
int lokib200_read_result(lokib200_engine* h, double* result);

/* --- multi-GPU: the one exchange of the path ---
 * The ensemble shards by global electron id (lokib200_config.first_electron_id); tables are replicated; per sampling interval the
 * result vectors of all shards are combined (SUM entries + MAX entries) in place in device memory, on the engines' streams, so that every
 * rank holds the same combined vector and takes the same trial-frequency / table decisions.  Transport: when every GPU of the communicator
 * can address the others (NVLink / NVSwitch peer access; CUDA IPC between processes), ONE kernel per GPU pushes its ~2.6 KB vector into a
 * mailbox in every peer's memory and adds the n slots in rank order (k_exchange: no ring, all ranks hold the same bits); otherwise ONE
 * grouped NCCL all-reduce.  LOKIB200_P2P=0 forces NCCL.  Histograms are combined once per job with NCCL.  NCCL is loaded at run time
 * (libnccl.so.2; the copy a host process such as PyTorch has already loaded is reused) and also carries the 64-byte IPC handles once.
 * The reference has no distributed backend (it is one OpenMP process, BMC.C:636). */
#define LOKIB200_COMM_ID_BYTES 128
int lokib200_comm_unique_id(void* id128);                                  /* ncclGetUniqueId: call on one rank, broadcast the 128 bytes */
/* one engine per process (torchrun / mpirun): collective call on all ranks */
int lokib200_comm_init_rank(lokib200_engine* h, const void* id128, int32_t rank, int32_t n_ranks);
/* all n engines live in this process, one per device (what `lokimc_b200 SETUP` with LOKIB200_GPUS=n does): ncclCommInitAll */
int lokib200_comm_init_all(lokib200_engine* const* engines, int32_t n);
int lokib200_comm_destroy(lokib200_engine* h);
int32_t lokib200_comm_size(const lokib200_engine* h);                      /* 1 without a communicator */
const char* lokib200_comm_transport(const lokib200_engine* h);             /* "none", "nccl" or "peer-memory": how result vectors are combined */
/* all-reduce the result vectors the last advance / sample left on the device; `engines` are the LOCAL members of the communicator
 * (all of them after lokib200_comm_init_all, the single one after lokib200_comm_init_rank); d_results[i] may name a caller-owned
 * device vector per engine (NULL = the engine's own, which lokib200_read_result then returns).  Asynchronous on the engines' streams. */
int lokib200_comm_allreduce_results(lokib200_engine* const* engines, int32_t n, double* const* d_results);
/* all-reduce (SUM) the accumulated histogram counts; afterwards every rank's lokib200_fetch_histograms returns the global counts */
int lokib200_comm_allreduce_histograms(lokib200_engine* const* engines, int32_t n);

/* --- distributions (replaces: getTimeDependDistributions BMC.C:1492-1572, histogramCount / histogram2DCount MathFunctions.C:61-127) ---
 * grids are those of checkSteadyState (BMC.C:1862-1883): energy [0,max_eedf_energy], cos in [-1,1], v_r in [0,v_max], v_z in [-v_max,v_max] */
int lokib200_set_histogram_grid(lokib200_engine* h, double max_eedf_energy);   /* also zeroes the accumulators */
/* counts the current ensemble (phase_index < 0: no phase-resolved EEDF).  On an engine that advances by graph launches the pass is queued into
 * the front of the next interval (the ensemble does not change in between); lokib200_fetch_histograms, a regrid, a new grid or ensemble, or
 * lokib200_advance_to_sync_device run it first, so a caller never observes the difference. */
int lokib200_sample_histograms(lokib200_engine* h, int32_t phase_index);
/* accumulated counts as doubles (eehSum [nE], eahSum [nE][nCos], evhSum [nR][nA], eehSum_periodic [nPhases][nE]); any may be NULL */
int lokib200_fetch_histograms(lokib200_engine* h, double* eeh, double* eah, double* evh, double* eeh_periodic);

/* --- injected-draw parity entry: n independent electrons, one pass of the loop body BMC.C:637-681 each, executed by the SAME
 *     device functions as lokib200_advance_to_sync.  draws is [n][n_draws] --- */
int lokib200_step_injected(lokib200_engine* h, int32_t n, const lokib200_electron* in, double nu_trial, const double* t_sync,
                           const double* draws, int32_t n_draws, lokib200_electron* out, lokib200_event_out* ev);

/* ensemble sums of the CURRENT state without advancing (the t = 0 sample of evaluateEEDF, BMC.C:310); fills the SUM_EPS..N_SAMPLED
 * and MAX_EPS entries of `result`, zeroes the rest */
int lokib200_sample_moments(lokib200_engine* h, double* result);
/* same, asynchronous: the vector stays on the device (engine-owned), to be combined with lokib200_comm_allreduce_results and read
 * with lokib200_read_result */
int lokib200_sample_moments_device(lokib200_engine* h);
/* getTimeDependDistributions' regrid (BMC.C:1497-1548): new energy grid [0,new_max_eedf_energy] for the EEDF / angular histograms
 * (the velocity grid is kept, as in the reference); the device accumulators restart from zero and the caller carries the
 * remapped old counts */
int lokib200_regrid_energy_histograms(lokib200_engine* h, double new_max_eedf_energy);

/* --- helpers that mirror host-side scalar logic of the path --- */
/* maximizationAccelerationEnergy (BMC.C:765-802) */
double lokib200_max_accel_energy(const lokib200_engine* h, double initial_energy, double dt);
/* checkMaxCollisionFrequency (BMC.C:716-763): may rebuild tables; returns the (possibly raised) trial frequency through *nu_trial */
int lokib200_check_nu_trial(lokib200_engine* h, double max_energy, double horizon, double energy_max_elastic, double* nu_trial);
/* which form of the advance kernel this engine launches: 0 = one electron per thread (k_advance, small ensembles), 1 = streaming pool in
 * shared memory (k_advance_stream, large ensembles); the choice is made in lokib200_set_processes from n_electrons and the SM count */
int32_t lokib200_kernel_form(const lokib200_engine* h);
/* Fast mode (not a reference feature; off by default = the reference's single trial collision frequency, BMC.C:716-763).  When on, every free
 * time is drawn against the trial frequency of the electron's energy band (half octaves; the maximum of nu_tot over all energies reachable within
 * a look-ahead time, from the same tables and the same acceleration bound) and a flight that outlasts the look-ahead time is cut and redrawn.
 * Same physics, far fewer null collisions; the event counters then count fewer (null) events per unit of physical time. */
int lokib200_set_fast_mode(lokib200_engine* h, int32_t on);
/* nominal HBM bandwidth of the engine's device [GB/s] from its memory clock and bus width (status display: fraction of the roofline) */
double lokib200_device_hbm_gbs(const lokib200_engine* h);
/* number of kernels launched by this engine so far (bench.py's gpu_launches) */
int64_t lokib200_launch_count(const lokib200_engine* h);
/* average device time [ms] of the advance kernel over the launches since the last call (CUDA events on the engine's stream) */
int lokib200_kernel_time_ms(lokib200_engine* h, double* advance_ms, int64_t* launches);
/* measurement aid ("the builder must measure the DFMA peak on the box"): runs a register-resident chain of
 * independent DFMAs on every SM of the engine's device and returns the sustained rate in TFLOP/s (one DFMA = 2 flop) */
int lokib200_measure_fp64_peak(lokib200_engine* h, double* tflops);


/* ------------------------------------------------------------------------------------------------------------------------
 * Host driver of one job: BoltzmannMC::evaluateEEDF (BMC.C:299-426) restated on top of the entry points above, in C++
 * (loki_mc_b200/host/boltzmann_mc.cpp).  It keeps the reference's control flow: checkMaxCollisionFrequency before every
 * interval, sampling every `sync_over_sampling` intervals, checkSteadyState (:1787-1893), checkStatisticalErrors (:1743-1767),
 * the stop criteria (:320-325) and the time averages (:1484-1741).  Several engines (one per GPU, disjoint electron-id ranges)
 * may be passed: when they share a communicator (lokib200_comm_init_all / _init_rank) their per-interval result vectors are combined
 * by one NCCL all-reduce per sampling interval, otherwise they are read back one by one and summed on the host.
 * ------------------------------------------------------------------------------------------------------------------------ */
typedef struct lokib200_solve_controls {   /* numericsMC keys, BMC.h:262-365 (values already multiplied by nElectrons where the ctor does) */
  double n_integration_points;           /* requiredIntegrationPoints */
  double n_integrated_ss_times;          /* requiredIntegratedSSTimes */
  double integrated_absolute_time;       /* requiredIntegratedAbsoluteTime */
  int32_t errors_to_be_checked;          /* numericsMC.relError present */
  int32_t sync_over_sampling;            /* synchronizationOverSampling (>= 1) */
  double rel_err_mean_energy, rel_err_flux_drift, rel_err_flux_diff, rel_err_bulk_drift, rel_err_bulk_diff, rel_err_power_balance;
  double min_collisions_before_ss, max_collisions_before_ss, max_collisions_after_ss;
  double sync_factor;                    /* synchronizationTimeXMaxCollisionFrequency */
  double initial_temp_ratio;             /* initialElecTempOverGasTemp */
  double energy_max_elastic;             /* energyMaxElastic (BMC.C:158) */
  int64_t max_intervals;                 /* safety stop (0 = none); not a reference key */
  int32_t status_display;                /* gui.terminalDisp contains MCStatus (BMC.h:368-374): print the status table of dispInfo (BMC.C:1895-1948) */
  int32_t fast_mode;                     /* 0 = the reference's single trial collision frequency (default); 1 = per-energy-band trial frequencies (numericsMC.fastMode, not a reference key) */
  double status_values[4];               /* header of the status table: E/N [Td], excitation frequency [Hz], field angle [degrees], B/N [Hx] */
} lokib200_solve_controls;

typedef struct lokib200_solve_results {   /* the public members the reference's sinks read (Output.h:123-147, 794-820) */
  double averaged_mean_energy, averaged_mean_energy_error;
  double flux_drift_velocity[3], flux_drift_velocity_error[3];
  double flux_diffusion[9], flux_diffusion_error[9];
  double bulk_drift_velocity[3], bulk_drift_velocity_error[3];
  double bulk_diffusion[9], bulk_diffusion_error[9];
  double power_gain_field, power_growth, power_balance_rel_error;   /* averagedPowerGainField, averagedPowerGrowth (eV m3 / s per electron) */
  double time, steady_state_time, total_integrated_time, trial_collision_frequency, max_eedf_energy, elapsed_seconds;
  double total_collisions, null_collisions, collisions_at_ss, null_collisions_at_ss;
  int64_t n_sampling_points, n_integration_points, n_sync_points, n_table_rebuilds;
  int32_t good_statistical_errors, stopped_by_max_collisions;
  double n_nu_exceeded;     /* collisions that found nu_tot(eps) above the trial frequency they were drawn with (LOKIB200_R_N_NU_EXCEEDED, whole job): the driver raises the trial frequency when it sees one (BMC.C:758-761) */
  double n_table_clamped;   /* collisions beyond the last table row (LOKIB200_R_N_TABLE_CLAMPED, whole job): the driver rebuilds the tables */
  double events_per_second; /* (real + null collisions) / elapsed_seconds */
} lokib200_solve_results;

typedef struct lokib200_job lokib200_job;
int lokib200_job_create(lokib200_engine* const* engines, int32_t n_engines, const lokib200_solve_controls* c, lokib200_job** out);
int lokib200_job_solve(lokib200_job* j, lokib200_solve_results* res);
/* the results of the last successful lokib200_job_solve (error if the job has not been solved) */
int lokib200_job_results(const lokib200_job* j, lokib200_solve_results* res);
/* per-process outputs: averagedRateCoeffs, averagedPowerGainProcesses, averagedPowerLossProcesses, collisionCounters (after steady state) */
int lokib200_job_process_outputs(const lokib200_job* j, double* rate_coeffs, double* power_gain, double* power_loss, double* counts);
/* time series (MCTemporalInfo): sampling times, mean energies, mean positions[3], mean velocities[3], position covariances[9] per sample; any may be NULL */
int64_t lokib200_job_time_series(const lokib200_job* j, double* times, double* mean_energy, double* mean_pos, double* mean_vel, double* pos_cov);
/* accumulated histograms incl. the carry of earlier grids: eehSum[nE], eahSum[nE][nCos], evhSum[nR][nA], eehSum_periodic[nPh][nE] */
int lokib200_job_histograms(lokib200_job* j, double* eeh, double* eah, double* evh, double* eeh_periodic);
/* phase-resolved sums (AC field): nIntegrationPointsPerPhase[nPh], meanEnergies_periodic[nPh], fluxVelocities_periodic[nPh][3],
 * bulkVelocities_periodic[nPh][3] (already divided by the points per phase, BMC.C:1728-1734) */
int lokib200_job_periodic(const lokib200_job* j, double* points_per_phase, double* mean_energy, double* flux_velocity, double* bulk_velocity);
/* fluxDiffusionCoeffs_periodic[nPh][9], bulkDiffusionCoeffs_periodic[nPh][9] (BMC.C:1479-1480, divided by the points per phase) */
int lokib200_job_periodic_diffusion(const lokib200_job* j, double* flux_diffusion, double* bulk_diffusion);
/* upper node of the velocity-histogram grids, sqrt(2 e maxEedfEnergy / m) at the moment the steady state was found (BMC.C:1877-1884) */
double lokib200_job_evdf_max_speed(const lokib200_job* j);
/* job constants: the summed-over-engines configuration (n_electrons = all shards) and the number of processes */
int lokib200_job_conditions(const lokib200_job* j, lokib200_config* cfg, int32_t* n_processes);
const char* lokib200_job_last_error(const lokib200_job* j);
void lokib200_job_destroy(lokib200_job* j);

#ifdef __cplusplus
}
#endif
#endif
