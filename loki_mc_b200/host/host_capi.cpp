// host_capi.cpp -- extern "C" surface of the host-side input rows (include/lokib200_host.h).
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/lokib200_host.h"
#include "setup_input.h"

struct lokib200_setup {
  std::unique_ptr<lokihost::SetupInput> in;
  std::string error, scratch;
};

static thread_local std::string g_setup_error;

extern "C" {

int lokib200_setup_load(const char* input_dir, const char* setup_file, lokib200_setup** out) {
  if (!input_dir || !setup_file || !out) { g_setup_error = "lokib200_setup_load: null argument"; return LOKIB200_ERR_INVALID; }
  *out = nullptr;
  try {
    auto s = std::make_unique<lokib200_setup>();
    s->in = std::make_unique<lokihost::SetupInput>(input_dir, setup_file);
    *out = s.release();
    return LOKIB200_OK;
  } catch (const std::exception& e) { g_setup_error = e.what(); return LOKIB200_ERR_INVALID; }
}
void lokib200_setup_destroy(lokib200_setup* s) { delete s; }
const char* lokib200_setup_last_error(const lokib200_setup* s) { return s ? s->error.c_str() : g_setup_error.c_str(); }

int32_t lokib200_setup_job_count(const lokib200_setup* s) { return s ? s->in->nJobs() : 0; }
double lokib200_setup_job_value(const lokib200_setup* s, int32_t job) { return (s && job >= 0 && job < s->in->nJobs()) ? s->in->jobValue(job) : 0.0; }
const char* lokib200_setup_variable_condition(const lokib200_setup* s) { return s ? s->in->wc.variableCondition.c_str() : ""; }

int lokib200_setup_processes(const lokib200_setup* s, lokib200_process_soa* out) {
  if (!s || !out) return LOKIB200_ERR_INVALID;
  *out = s->in->processes.soa();
  return LOKIB200_OK;
}
int lokib200_setup_config(const lokib200_setup* s, int32_t job, lokib200_config* out) {
  if (!s || !out || job < 0 || job >= s->in->nJobs()) return LOKIB200_ERR_INVALID;
  try { *out = s->in->config(job); } catch (const std::exception& e) { const_cast<lokib200_setup*>(s)->error = e.what(); return LOKIB200_ERR_INVALID; }
  return LOKIB200_OK;
}
int lokib200_setup_controls(const lokib200_setup* s, lokib200_solve_controls* out) {
  if (!s || !out) return LOKIB200_ERR_INVALID;
  try { *out = s->in->controls(); } catch (const std::exception& e) { const_cast<lokib200_setup*>(s)->error = e.what(); return LOKIB200_ERR_INVALID; }
  return LOKIB200_OK;
}
const char* lokib200_setup_process_description(const lokib200_setup* s, int32_t k) {
  return (s && k >= 0 && k < static_cast<int32_t>(s->in->processes.descriptions.size())) ? s->in->processes.descriptions[k].c_str() : "";
}
int32_t lokib200_setup_process_is_elastic(const lokib200_setup* s, int32_t k) {
  return (s && k >= 0 && k < static_cast<int32_t>(s->in->processes.isElastic.size())) ? s->in->processes.isElastic[k] : 0;
}
double lokib200_setup_energy_max_elastic(const lokib200_setup* s) { return s ? s->in->processes.energyMaxElastic : 0.0; }
const char* lokib200_setup_value(const lokib200_setup* s, const char* key) {
  if (!s || !key) return "";
  const_cast<lokib200_setup*>(s)->scratch = s->in->tree->value(key);
  return s->scratch.c_str();
}
int64_t lokib200_setup_dump(const lokib200_setup* s, char* buf, int64_t cap) {
  if (!s) return 0;
  const std::string d = s->in->tree->dump();
  if (buf && cap > 0) { const size_t n = std::min(static_cast<size_t>(cap - 1), d.size()); std::memcpy(buf, d.data(), n); buf[n] = 0; }
  return static_cast<int64_t>(d.size());
}
int32_t lokib200_setup_warning_count(const lokib200_setup* s) { return s ? static_cast<int32_t>(s->in->mixture->warnings.size()) : 0; }
const char* lokib200_setup_warning(const lokib200_setup* s, int32_t i) {
  return (s && i >= 0 && i < static_cast<int32_t>(s->in->mixture->warnings.size())) ? s->in->mixture->warnings[i].c_str() : "";
}

double lokib200_eval_expression(const char* expr, int32_t* ok) {
  try { const double v = lokihost::evalExpression(expr ? expr : ""); if (ok) *ok = 1; return v; }
  catch (const std::exception& e) { g_setup_error = e.what(); if (ok) *ok = 0; return 0.0; }
}
int64_t lokib200_eval_vector_expression(const char* expr, double* out, int64_t cap, int32_t* ok) {
  try {
    const auto v = lokihost::evalVectorExpression(expr ? expr : "");
    if (ok) *ok = 1;
    for (int64_t i = 0; i < static_cast<int64_t>(v.size()) && i < cap; ++i) out[i] = v[i];
    return static_cast<int64_t>(v.size());
  } catch (const std::exception& e) { g_setup_error = e.what(); if (ok) *ok = 0; return 0; }
}

}  // extern "C"
