/* lokib200_host.h -- C ABI of the host-side rows either side of the hot path: setup/LXCat input, swarm-parameter
 * post-processing and the reference's text outputs.  Pure host code (no CUDA calls); lives in the same liblokib200.so.
 *
 * Reference interfaces replaced (Code/LoKI-MC):
 *   lokib200_setup_*      Headers/Setup.h:64-227 (Setup ctor + initializeEedf), Sources/Parse.C:24-125 (setupFile, LXCatFiles),
 *                         Sources/FieldInfo.C, Headers/WorkingConditions.h:56-178, Sources/BoltzmannMC.C:29-271 (process flattening)
 *   lokib200_eval_*       Sources/Parse.C:610-755 (evalVectorExpress, str2value)
 *   lokib200_report_*     Sources/BoltzmannMC.C:1574-1604 (getTimeAverageDistributions), :1727-1741 (getAveragedPeriodicParams), :1950-2063 (evaluatePower),
 *                         :2065-2271 (evaluateSwarmParameters), :2273-2384 (evaluateRateCoeff), Sources/Collision.C:355-396, Sources/Grid.C:46-55
 *   lokib200_run_setup    Sources/lokimc.C:14-95 (main, LoKISimulation), Headers/Setup.h:92-227, :948-957
 *   lokib200_output_*     Headers/Output.h:46-822 (folder layout, file names, headers, printf formats)
 */
#ifndef LOKIB200_HOST_H
#define LOKIB200_HOST_H
#include "lokib200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lokib200_setup lokib200_setup;

/* Parse `<input_dir>/<setup_file>`, the LXCat files and Databases it names, build the gas/state/collision ontology and flatten it.
 * Returns 0 or LOKIB200_ERR_INVALID; the message of a failed load is returned by lokib200_setup_last_error(NULL). */
int lokib200_setup_load(const char* input_dir, const char* setup_file, lokib200_setup** out);
void lokib200_setup_destroy(lokib200_setup* s);
const char* lokib200_setup_last_error(const lokib200_setup* s);

int32_t lokib200_setup_job_count(const lokib200_setup* s);                       /* WorkingConditions::nJobs */
double lokib200_setup_job_value(const lokib200_setup* s, int32_t job);             /* value of the swept working condition */
const char* lokib200_setup_variable_condition(const lokib200_setup* s);
/* the arrays stay owned by `s` (valid until destroy) */
int lokib200_setup_processes(const lokib200_setup* s, lokib200_process_soa* out);
int lokib200_setup_config(const lokib200_setup* s, int32_t job, lokib200_config* out);
int lokib200_setup_controls(const lokib200_setup* s, lokib200_solve_controls* out);
const char* lokib200_setup_process_description(const lokib200_setup* s, int32_t k);  /* Collision::description(), e.g. "e+N2(X)->e+N2(X),Elastic" */
int32_t lokib200_setup_process_is_elastic(const lokib200_setup* s, int32_t k);
double lokib200_setup_energy_max_elastic(const lokib200_setup* s);
const char* lokib200_setup_value(const lokib200_setup* s, const char* dotted_key);   /* FieldInfo::getFieldValue; "" when absent */
/* FieldInfo::printSetupInfo text (what Output writes to setup.txt); returns the length, copies at most cap-1 bytes */
int64_t lokib200_setup_dump(const lokib200_setup* s, char* buf, int64_t cap);
int32_t lokib200_setup_warning_count(const lokib200_setup* s);
const char* lokib200_setup_warning(const lokib200_setup* s, int32_t i);

/* Parse::str2value / evalVectorExpress; *ok = 0 on a malformed expression */
double lokib200_eval_expression(const char* expr, int32_t* ok);
int64_t lokib200_eval_vector_expression(const char* expr, double* out, int64_t cap, int32_t* ok);

/* ---- post-processing and text outputs of one finished job ---- */

/* What the reference's sinks read from a finished BoltzmannMC: exactly what lokib200_job_* return.  Arrays are row-major. */
typedef struct lokib200_job_data {
  const lokib200_solve_results* results;
  double n_electrons;                 /* all shards */
  double evdf_max_speed;              /* lokib200_job_evdf_max_speed */
  const double *rate_coeffs, *power_gain, *power_loss, *counts;          /* [P], lokib200_job_process_outputs */
  const double *eeh, *eah, *evh, *eeh_periodic;                          /* lokib200_job_histograms; eah/evh may be NULL when E is not along z */
  int64_t n_samples;
  const double *times, *mean_energy, *mean_pos, *mean_vel, *pos_cov;     /* lokib200_job_time_series */
  const double *points_per_phase, *mean_energy_periodic, *flux_velocity_periodic, *bulk_velocity_periodic;   /* from the job's periodic sums; AC field only */
  const double *flux_diffusion_periodic, *bulk_diffusion_periodic;       /* lokib200_job_periodic_diffusion */
} lokib200_job_data;

typedef struct lokib200_report lokib200_report;
typedef struct lokib200_output lokib200_output;

/* distributions (eedf, anisotropies, evdf), power balance, rate coefficients and swarm parameters of job `job` of the setup */
int lokib200_report_create(const lokib200_setup* s, int32_t job, const lokib200_job_data* d, lokib200_report** out);
/* the same, pulling the data from a solved lokib200_job */
int lokib200_report_from_job(const lokib200_setup* s, int32_t job, lokib200_job* j, lokib200_report** out);
void lokib200_report_destroy(lokib200_report* r);
const char* lokib200_report_last_error(const lokib200_report* r);            /* NULL: error of the last failed create */
/* swarmParam[name] (BMC.C:2065-2271: "fluxRedMobCoeff", "meanEnergy", "Te", "totalIonRateCoeff", ...); *found = 0 if absent */
double lokib200_report_swarm(const lokib200_report* r, const char* name, int32_t* found);
/* power.Map[name] or, with gas != NULL, power.gasesMap[gas][name] (BMC.C:1950-2063) */
double lokib200_report_power(const lokib200_report* r, const char* name, const char* gas, int32_t* found);
int32_t lokib200_report_energy_cells(const lokib200_report* r);
/* energyGrid->cell, eedf, first and second anisotropies, each [n_energy_cells]; any may be NULL */
int lokib200_report_eedf(const lokib200_report* r, double* energy, double* eedf, double* first_anisotropy, double* second_anisotropy);
/* rateCoeffAll (extra = 0) / rateCoeffExtra (extra = 1) */
int32_t lokib200_report_rate_count(const lokib200_report* r, int32_t extra);
int lokib200_report_rate(const lokib200_report* r, int32_t extra, int32_t i, int32_t* coll_id, double* ine, double* sup, double* ine_mc, double* sup_mc,
                         const char** description);

/* Output (Headers/Output.h): creates <output_root>/<output.folder>, writes setup.txt; no-op object when output.isOn is false */
int lokib200_output_create(const lokib200_setup* s, const char* output_root, lokib200_output** out);
/* Output::electronKineticsSolution: the data files selected in output.dataFiles for this job (jobs must be written in order) */
int lokib200_output_write(lokib200_output* o, const lokib200_report* r);
const char* lokib200_output_folder(const lokib200_output* o);
const char* lokib200_output_last_error(const lokib200_output* o);
void lokib200_output_destroy(lokib200_output* o);

/* ---- whole simulation: what the reference's executable does (Sources/lokimc.C) ---- */
typedef struct lokib200_run_summary {
  int32_t n_jobs;
  double last_mean_energy;      /* eV, of the last job */
  double total_collisions;      /* real + null, all jobs */
  double device_seconds;        /* sum of the jobs' solve times */
  double elapsed_seconds;       /* wall clock of the call */
} lokib200_run_summary;
/* Parse <input_dir>/<setup_file>, run every job on n_devices GPUs (devices first_device .. first_device + n_devices - 1, the ensemble sharded by
 * global electron id), post-process and write the output folder under <output_root> (if output.isOn).  verbose != 0 prints what the reference prints
 * to the terminal.  Returns 0 or a negative status; the message is lokib200_run_last_error(). */
int lokib200_run_setup(const char* input_dir, const char* setup_file, const char* output_root, int32_t n_devices, int32_t first_device, int32_t verbose,
                       lokib200_run_summary* summary);
const char* lokib200_run_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
