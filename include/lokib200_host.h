/* lokib200_host.h -- C ABI of the host-side rows either side of the hot path: setup/LXCat input, swarm-parameter
 * post-processing and the reference's text outputs.  Pure host code (no CUDA calls); lives in the same liblokib200.so.
 *
 * Reference interfaces replaced (Code/LoKI-MC):
 *   lokib200_setup_*      Headers/Setup.h:64-227 (Setup ctor + initializeEedf), Sources/Parse.C:24-125 (setupFile, LXCatFiles),
 *                         Sources/FieldInfo.C, Headers/WorkingConditions.h:56-178, Sources/BoltzmannMC.C:29-271 (process flattening)
 *   lokib200_eval_*       Sources/Parse.C:610-755 (evalVectorExpress, str2value)
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

#ifdef __cplusplus
}
#endif
#endif
