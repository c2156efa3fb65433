// host_capi.cpp -- extern "C" surface of the host-side input rows (include/lokib200_host.h).
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/lokib200_host.h"
#include "report.h"
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

// ---------------------------------------------------------------- report / output ----------------------------------------------------------------
static lokihost::JobData copyJobData(const lokihost::SetupInput& in, int job, const lokib200_job_data* d) {
  if (!d || !d->results) throw lokihost::SetupError("null job data");
  const lokib200_config cfg = in.config(job);
  const size_t P = in.processes.type.size(), nE = cfg.n_energy_cells, nC = cfg.n_cos_cells, nR = cfg.n_radial_cells, nA = cfg.n_axial_cells;
  const size_t nPh = cfg.excitation_omega != 0 ? cfg.n_phases : 0, n = static_cast<size_t>(d->n_samples);
  auto vec = [](const double* p, size_t k) { return p ? std::vector<double>(p, p + k) : std::vector<double>(); };
  lokihost::JobData o;
  o.res = *d->results; o.nElectrons = d->n_electrons; o.evdfMaxSpeed = d->evdf_max_speed;
  o.rateCoeffsMC = vec(d->rate_coeffs, P); o.powerGain = vec(d->power_gain, P); o.powerLoss = vec(d->power_loss, P); o.counts = vec(d->counts, P);
  o.eehSum = vec(d->eeh, nE); o.eahSum = vec(d->eah, nE * nC); o.evhSum = vec(d->evh, nR * nA); o.eehSumPeriodic = vec(d->eeh_periodic, nPh * nE);
  o.samplingTimes = vec(d->times, n); o.meanEnergies = vec(d->mean_energy, n); o.meanPositions = vec(d->mean_pos, 3 * n); o.meanVelocities = vec(d->mean_vel, 3 * n);
  o.positionCovariances = vec(d->pos_cov, 9 * n);
  o.pointsPerPhase = vec(d->points_per_phase, nPh); o.meanEnergiesPeriodic = vec(d->mean_energy_periodic, nPh);
  o.fluxVelocitiesPeriodic = vec(d->flux_velocity_periodic, 3 * nPh); o.bulkVelocitiesPeriodic = vec(d->bulk_velocity_periodic, 3 * nPh);
  o.fluxDiffusionPeriodic = vec(d->flux_diffusion_periodic, 9 * nPh); o.bulkDiffusionPeriodic = vec(d->bulk_diffusion_periodic, 9 * nPh);
  if (static_cast<int64_t>(n) < o.res.n_sampling_points) throw lokihost::SetupError("job data: fewer samples than results->n_sampling_points");
  if (nPh && (o.pointsPerPhase.empty() || o.fluxDiffusionPeriodic.empty() || o.meanEnergiesPeriodic.empty() || o.fluxVelocitiesPeriodic.empty() || o.bulkVelocitiesPeriodic.empty() ||
              o.bulkDiffusionPeriodic.empty()))
    throw lokihost::SetupError("job data: the phase-resolved arrays are required when the field is AC");
  return o;
}

struct lokib200_report {
  std::unique_ptr<lokihost::Report> rep;
  std::string error;
};
struct lokib200_output {
  std::unique_ptr<lokihost::OutputWriter> out;
  std::string error;
};

extern "C" {

int lokib200_report_create(const lokib200_setup* s, int32_t job, const lokib200_job_data* d, lokib200_report** out) {
  if (!s || !out || job < 0 || job >= s->in->nJobs()) { g_setup_error = "lokib200_report_create: invalid argument"; return LOKIB200_ERR_INVALID; }
  *out = nullptr;
  try {
    auto r = std::make_unique<lokib200_report>();
    r->rep = std::make_unique<lokihost::Report>(*s->in, job, copyJobData(*s->in, job, d));
    *out = r.release();
    return LOKIB200_OK;
  } catch (const std::exception& e) { g_setup_error = e.what(); return LOKIB200_ERR_INVALID; }
}

int lokib200_report_from_job(const lokib200_setup* s, int32_t job, lokib200_job* j, lokib200_report** out) {
  if (!s || !j || !out || job < 0 || job >= s->in->nJobs()) { g_setup_error = "lokib200_report_from_job: invalid argument"; return LOKIB200_ERR_INVALID; }
  *out = nullptr;
  lokib200_config cfg; int32_t P = 0;
  lokib200_job_conditions(j, &cfg, &P);
  if (static_cast<size_t>(P) != s->in->processes.type.size()) { g_setup_error = "lokib200_report_from_job: the job was not built from this setup"; return LOKIB200_ERR_INVALID; }
  lokib200_solve_results res;
  // the averaged results were stored by lokib200_job_solve; re-read them without advancing (solve on a finished job returns them)
  if (lokib200_job_results(j, &res)) { g_setup_error = std::string("lokib200_report_from_job: ") + lokib200_job_last_error(j); return LOKIB200_ERR_INVALID; }
  const size_t nE = cfg.n_energy_cells, nC = cfg.n_cos_cells, nR = cfg.n_radial_cells, nA = cfg.n_axial_cells, nPh = cfg.excitation_omega != 0 ? cfg.n_phases : 0;
  std::vector<double> rate(P), gain(P), loss(P), counts(P), eeh(nE), eah(nE * nC), evh(nR * nA), per(std::max<size_t>(nPh * nE, 1));
  lokib200_job_process_outputs(j, rate.data(), gain.data(), loss.data(), counts.data());
  if (lokib200_job_histograms(j, eeh.data(), eah.data(), evh.data(), per.data())) { g_setup_error = std::string("lokib200_report_from_job: ") + lokib200_job_last_error(j); return LOKIB200_ERR_INVALID; }
  const int64_t n = lokib200_job_time_series(j, nullptr, nullptr, nullptr, nullptr, nullptr);
  std::vector<double> t(n), me(n), mp(3 * n), mv(3 * n), pc(9 * n);
  lokib200_job_time_series(j, t.data(), me.data(), mp.data(), mv.data(), pc.data());
  std::vector<double> pts(nPh), mep(nPh), fv(3 * nPh), bv(3 * nPh), fd(9 * nPh), bd(9 * nPh);
  if (nPh) { lokib200_job_periodic(j, pts.data(), mep.data(), fv.data(), bv.data()); lokib200_job_periodic_diffusion(j, fd.data(), bd.data()); }
  lokib200_job_data d{};
  d.results = &res; d.n_electrons = static_cast<double>(cfg.n_electrons); d.evdf_max_speed = lokib200_job_evdf_max_speed(j);
  d.rate_coeffs = rate.data(); d.power_gain = gain.data(); d.power_loss = loss.data(); d.counts = counts.data();
  d.eeh = eeh.data(); d.eah = eah.data(); d.evh = evh.data(); d.eeh_periodic = nPh ? per.data() : nullptr;
  d.n_samples = n; d.times = t.data(); d.mean_energy = me.data(); d.mean_pos = mp.data(); d.mean_vel = mv.data(); d.pos_cov = pc.data();
  if (nPh) { d.points_per_phase = pts.data(); d.mean_energy_periodic = mep.data(); d.flux_velocity_periodic = fv.data(); d.bulk_velocity_periodic = bv.data();
             d.flux_diffusion_periodic = fd.data(); d.bulk_diffusion_periodic = bd.data(); }
  return lokib200_report_create(s, job, &d, out);
}

void lokib200_report_destroy(lokib200_report* r) { delete r; }
const char* lokib200_report_last_error(const lokib200_report* r) { return r ? r->error.c_str() : g_setup_error.c_str(); }

double lokib200_report_swarm(const lokib200_report* r, const char* name, int32_t* found) {
  if (found) *found = 0;
  if (!r || !name) return 0.0;
  auto it = r->rep->swarm.find(name);
  if (it == r->rep->swarm.end()) return 0.0;
  if (found) *found = 1;
  return it->second;
}
double lokib200_report_power(const lokib200_report* r, const char* name, const char* gas, int32_t* found) {
  if (found) *found = 0;
  if (!r || !name) return 0.0;
  const std::map<std::string, double>* m = &r->rep->power;
  if (gas) { auto g = r->rep->powerByGas.find(gas); if (g == r->rep->powerByGas.end()) return 0.0; m = &g->second; }
  auto it = m->find(name);
  if (it == m->end()) return 0.0;
  if (found) *found = 1;
  return it->second;
}
int32_t lokib200_report_energy_cells(const lokib200_report* r) { return r ? r->rep->nE : 0; }
int lokib200_report_eedf(const lokib200_report* r, double* energy, double* eedf, double* a1, double* a2) {
  if (!r) return LOKIB200_ERR_INVALID;
  const auto& q = *r->rep;
  if (energy) std::copy(q.energyCell.begin(), q.energyCell.end(), energy);
  if (eedf) std::copy(q.eedf.begin(), q.eedf.end(), eedf);
  if (a1) std::copy(q.efadf.begin(), q.efadf.end(), a1);
  if (a2) std::copy(q.esadf.begin(), q.esadf.end(), a2);
  return LOKIB200_OK;
}
int32_t lokib200_report_rate_count(const lokib200_report* r, int32_t extra) { return r ? static_cast<int32_t>((extra ? r->rep->rateExtra : r->rep->rateAll).size()) : 0; }
int lokib200_report_rate(const lokib200_report* r, int32_t extra, int32_t i, int32_t* id, double* ine, double* sup, double* ineMC, double* supMC, const char** desc) {
  if (!r) return LOKIB200_ERR_INVALID;
  const auto& v = extra ? r->rep->rateExtra : r->rep->rateAll;
  if (i < 0 || i >= static_cast<int32_t>(v.size())) return LOKIB200_ERR_INVALID;
  if (id) *id = v[i].collID;
  if (ine) *ine = v[i].ineRate;
  if (sup) *sup = v[i].supRate;
  if (ineMC) *ineMC = v[i].ineRateMC;
  if (supMC) *supMC = v[i].supRateMC;
  if (desc) *desc = v[i].description.c_str();
  return LOKIB200_OK;
}

int lokib200_output_create(const lokib200_setup* s, const char* root, lokib200_output** out) {
  if (!s || !root || !out) { g_setup_error = "lokib200_output_create: null argument"; return LOKIB200_ERR_INVALID; }
  *out = nullptr;
  try {
    auto o = std::make_unique<lokib200_output>();
    o->out = std::make_unique<lokihost::OutputWriter>(*s->in, root);
    *out = o.release();
    return LOKIB200_OK;
  } catch (const std::exception& e) { g_setup_error = e.what(); return LOKIB200_ERR_INVALID; }
}
int lokib200_output_write(lokib200_output* o, const lokib200_report* r) {
  if (!o || !r) return LOKIB200_ERR_INVALID;
  try { o->out->write(*r->rep); return LOKIB200_OK; } catch (const std::exception& e) { o->error = e.what(); return LOKIB200_ERR_INVALID; }
}
const char* lokib200_output_folder(const lokib200_output* o) { return o ? o->out->folder.c_str() : ""; }
const char* lokib200_output_last_error(const lokib200_output* o) { return o ? o->error.c_str() : g_setup_error.c_str(); }
void lokib200_output_destroy(lokib200_output* o) { delete o; }

}  // extern "C"
