// run.cpp -- lokib200_run_setup: the reference's main loop (Sources/lokimc.C:14-42 LoKISimulation, Headers/Setup.h:92-227 initializeSimulation /
// nextJob, :948-957 finishSimulation) over this library: parse the setup, and for every job build the engines, solve, post-process, write.
//
// The jobs of a setup (one per value of the swept working condition) are independent: BoltzmannMC::evaluateNonConstantVariables starts every one
// from a fresh Maxwellian ensemble (BMC.C:491-508).  A job of the reference's usual size (1e4-1e5 electrons) occupies a fraction of a B200 and is
// bound by the event chain of its slowest warp, so small jobs are solved SIDE BY SIDE, each on its own engine, stream and host thread (and, with
// several devices, each on one device instead of sharded); their reports are written in job order, so the output folder is what the sequential
// loop writes.  Large ensembles keep the sequential, sharded form.  LOKIB200_CONCURRENT_JOBS=k overrides the choice (1 = the sequential loop).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lokib200_host.h"
#include "report.h"
#include "setup_input.h"

namespace {

thread_local std::string g_run_error;

struct Engines {
  std::vector<lokib200_engine*> e;
  ~Engines() { for (auto* p : e) lokib200_destroy(p); }
};
struct JobHandle {
  lokib200_job* j = nullptr;
  ~JobHandle() { if (j) lokib200_job_destroy(j); }
};

// what one solved job hands to the (sequential, ordered) post-processing and writing stage
struct Solved {
  lokihost::JobData data;
  lokib200_solve_results res{};
  std::string messages;     // warnings of this job, printed when its turn comes
  std::string error;
  bool done = false;
};

std::string warningText(const std::string& body) { return "\033[1;33mPay attention to the following warning:\n" + body + "\n\033[0m"; }

std::string format(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return buf;
}

// one job on devices [first_device, first_device + n_devices): engines, communicator, solve, everything the report needs
double seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const bool g_profile = std::getenv("LOKIB200_PROFILE") != nullptr;   // host-side time split on stderr

void solveJob(const lokihost::SetupInput& in, const lokib200_process_soa& soa, int job, int n_devices, int first_device, bool verbose, bool status, Solved& out) {
  const double t_begin = seconds();
  lokib200_config cfg = in.config(job);
  const int64_t total = cfg.n_electrons, per = total / n_devices;
  if (per < 1) throw lokihost::SetupError("numericsMC.nElectrons is smaller than the number of devices");
  Engines eng;
  uint64_t first = 0;
  for (int g = 0; g < n_devices; ++g) {
    lokib200_config c = cfg;
    c.device = first_device + g;
    c.n_electrons = per + (g == 0 ? total - per * n_devices : 0);   // shards by global electron id; the remainder goes to the first
    c.first_electron_id = first;
    first += static_cast<uint64_t>(c.n_electrons);
    lokib200_engine* h = nullptr;
    if (lokib200_create(&c, &h)) throw lokihost::SetupError(std::string("engine: ") + lokib200_last_error(nullptr));
    eng.e.push_back(h);
    if (lokib200_set_processes(h, &soa)) throw lokihost::SetupError(std::string("engine: ") + lokib200_last_error(h));
  }
  // shards of one job share a communicator: per sampling interval their result vectors are combined by one grouped NCCL all-reduce on the
  // devices.  Without a usable NCCL the driver reads the engines one by one and adds on the host.
  if (n_devices > 1 && lokib200_comm_init_all(eng.e.data(), n_devices) != 0 && verbose)
    out.messages += warningText(format("no NCCL communicator (%s): the per-interval sums of the %d GPUs are combined on the host", lokib200_last_error(eng.e[0]), n_devices));
  lokib200_solve_controls ctl = in.controls();
  if (!status) ctl.status_display = 0;
  {   // header of the status table (BMC.C:1914-1917): the working conditions of this job
    auto pick = [&](const std::vector<double>& a, const char* name) { return a[(in.wc.variableCondition == name) ? static_cast<size_t>(job) : 0]; };
    ctl.status_values[0] = pick(in.wc.reducedElecFieldArray, "reducedElecField"); ctl.status_values[1] = pick(in.wc.excitationFrequencyArray, "excitationFrequency");
    ctl.status_values[2] = pick(in.wc.elecFieldAngleArray, "elecFieldAngle"); ctl.status_values[3] = pick(in.wc.reducedMagFieldArray, "reducedMagField");
  }
  JobHandle jh;
  if (lokib200_job_create(eng.e.data(), n_devices, &ctl, &jh.j)) throw lokihost::SetupError("could not create the job");
  lokib200_solve_results& res = out.res;
  const double t_created = seconds();
  if (lokib200_job_solve(jh.j, &res)) throw lokihost::SetupError(lokib200_job_last_error(jh.j));
  const double t_solved = seconds();
  if (res.n_nu_exceeded > 1e-4 * (res.total_collisions + res.null_collisions) || res.n_table_clamped > 0)
    out.messages += warningText(format("%g collisions met a total collision frequency above their trial frequency and %g an energy beyond the cross-section tables inside a "
                                       "synchronisation interval (of %g events); the trial frequency was raised / the tables were rebuilt for the following intervals. [%s = %g]",
                                       res.n_nu_exceeded, res.n_table_clamped, res.total_collisions + res.null_collisions, in.wc.variableCondition.c_str(), in.jobValue(job)));
  if (res.stopped_by_max_collisions)   // BMC.C:398-415
    out.messages += warningText(format("Monte Carlo simulation ended after reaching ''maxCollisionsAfterSteadyState'' indicated in the setup file. [%s = %g]",
                                       in.wc.variableCondition.c_str(), in.jobValue(job)));
  int32_t P = 0;
  lokib200_config jc;
  lokib200_job_conditions(jh.j, &jc, &P);
  const size_t nE = jc.n_energy_cells, nC = jc.n_cos_cells, nR = jc.n_radial_cells, nA = jc.n_axial_cells, nPh = jc.excitation_omega != 0 ? jc.n_phases : 0;
  lokihost::JobData& d = out.data;
  d.res = res; d.nElectrons = static_cast<double>(jc.n_electrons); d.evdfMaxSpeed = lokib200_job_evdf_max_speed(jh.j);
  d.rateCoeffsMC.resize(P); d.powerGain.resize(P); d.powerLoss.resize(P); d.counts.resize(P);
  lokib200_job_process_outputs(jh.j, d.rateCoeffsMC.data(), d.powerGain.data(), d.powerLoss.data(), d.counts.data());
  d.eehSum.resize(nE); d.eahSum.resize(nE * nC); d.evhSum.resize(nR * nA); d.eehSumPeriodic.resize(nPh * nE);
  if (lokib200_job_histograms(jh.j, d.eehSum.data(), d.eahSum.data(), d.evhSum.data(), nPh ? d.eehSumPeriodic.data() : nullptr))
    throw lokihost::SetupError(std::string("histograms: ") + lokib200_job_last_error(jh.j));
  if (!jc.is_cylindrically_symmetric) { d.eahSum.clear(); d.evhSum.clear(); }
  const int64_t n = lokib200_job_time_series(jh.j, nullptr, nullptr, nullptr, nullptr, nullptr);
  d.samplingTimes.resize(n); d.meanEnergies.resize(n); d.meanPositions.resize(3 * n); d.meanVelocities.resize(3 * n); d.positionCovariances.resize(9 * n);
  lokib200_job_time_series(jh.j, d.samplingTimes.data(), d.meanEnergies.data(), d.meanPositions.data(), d.meanVelocities.data(), d.positionCovariances.data());
  if (nPh) {
    d.pointsPerPhase.resize(nPh); d.meanEnergiesPeriodic.resize(nPh); d.fluxVelocitiesPeriodic.resize(3 * nPh); d.bulkVelocitiesPeriodic.resize(3 * nPh);
    d.fluxDiffusionPeriodic.resize(9 * nPh); d.bulkDiffusionPeriodic.resize(9 * nPh);
    lokib200_job_periodic(jh.j, d.pointsPerPhase.data(), d.meanEnergiesPeriodic.data(), d.fluxVelocitiesPeriodic.data(), d.bulkVelocitiesPeriodic.data());
    lokib200_job_periodic_diffusion(jh.j, d.fluxDiffusionPeriodic.data(), d.bulkDiffusionPeriodic.data());
  }
  const double t_fetched = seconds();
  if (jh.j) { lokib200_job_destroy(jh.j); jh.j = nullptr; }
  const double t_job_gone = seconds();
  for (auto*& p : eng.e) { lokib200_destroy(p); p = nullptr; }
  if (g_profile)
    std::fprintf(stderr, "lokib200 job %d: engines + job created in %.3f s, solved in %.3f s (%lld sync points, %lld table builds), results fetched in %.3f s, job released in %.3f s, engines in %.3f s\n",
                 job, t_created - t_begin, t_solved - t_created, static_cast<long long>(res.n_sync_points), static_cast<long long>(res.n_table_rebuilds), t_fetched - t_solved,
                 t_job_gone - t_fetched, seconds() - t_job_gone);
}

// an ensemble this small leaves most of a B200 idle: it runs the one-electron-per-thread form of K1 (the engine switches to the streaming pool,
// which fills the SMs' shared memory, at 96 x 1024 electrons), a few hundred warps on 148 SMs x 64 warp slots
constexpr int64_t SIDE_BY_SIDE_MAX_ELECTRONS = 96 * 1024 - 1;
constexpr int SIDE_BY_SIDE_PER_DEVICE = 8;

}  // namespace

extern "C" {

const char* lokib200_run_last_error(void) { return g_run_error.c_str(); }

int lokib200_run_setup(const char* input_dir, const char* setup_file, const char* output_root, int32_t n_devices, int32_t first_device, int32_t verbose,
                       lokib200_run_summary* summary) {
  const auto start = std::chrono::high_resolution_clock::now();
  if (!input_dir || !setup_file || !output_root || n_devices < 1) { g_run_error = "lokib200_run_setup: invalid argument"; return LOKIB200_ERR_INVALID; }
  if (summary) *summary = lokib200_run_summary{};
  try {
    lokihost::SetupInput in(input_dir, setup_file);
    for (const auto& w : in.mixture->warnings) std::printf("%s", warningText(w).c_str());
    if (verbose) {
      const std::string on = in.tree->value("gui.isOn");
      bool show = false;
      if (on == "true" || on == "True" || on == "1") for (const auto& o : in.tree->childNames("gui.terminalDisp")) show = show || o == "setup";
      if (show) std::printf("%s", in.tree->dump().c_str());   // FieldInfo::printSetupInfo
      std::printf("Starting simulation...\n");
    }
    if (g_profile) std::fprintf(stderr, "lokib200 run: setup parsed in %.3f s\n", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count());
    lokihost::OutputWriter out(in, output_root);
    const lokib200_process_soa soa = in.processes.soa();
    const int nJobs = in.nJobs();

    // how many jobs at a time: small ensembles side by side (one device each), otherwise one job over all devices.  The status table of
    // gui.terminalDisp: MCStatus is a live display of ONE job, so it keeps the sequential loop.
    const bool status = verbose && in.controls().status_display != 0;
    int workers = 1;
    if (nJobs > 1 && !status && in.config(0).n_electrons <= SIDE_BY_SIDE_MAX_ELECTRONS) {
      // every worker spins in its stream wait: beyond about a third of the host's hardware threads they starve each other (measured: 8 workers on
      // a 16-thread box take twice as long per interval as 5)
      const int by_host = std::max(1, static_cast<int>(std::thread::hardware_concurrency()) / 3);
      workers = std::max(1, std::min({nJobs, SIDE_BY_SIDE_PER_DEVICE * n_devices, by_host}));
    }
    if (const char* env = std::getenv("LOKIB200_CONCURRENT_JOBS")) { const int k = std::atoi(env); if (k >= 1) workers = std::min(nJobs, k); }
    const bool side_by_side = workers > 1;

    std::vector<Solved> solved(static_cast<size_t>(nJobs));
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> next{0};
    std::atomic<bool> stop{false};
    const char* const skipped = "not run: an earlier job failed";
    auto work = [&](int w) {
      for (;;) {
        const int job = next.fetch_add(1);
        if (job >= nJobs) return;
        Solved& s = solved[static_cast<size_t>(job)];
        if (stop.load()) s.error = skipped;   // drain the queue so that nobody waits for a job that will never run
        else {
          try { solveJob(in, soa, job, 1, first_device + w % n_devices, verbose != 0, false, s); }
          catch (const std::exception& e) { s.error = e.what(); if (s.error.empty()) s.error = "job failed"; stop.store(true); }
        }
        { std::lock_guard<std::mutex> lk(mu); s.done = true; }
        cv.notify_all();
      }
    };
    std::vector<std::thread> pool;
    if (side_by_side) for (int w = 0; w < workers; ++w) pool.emplace_back(work, w);
    struct Joiner { std::vector<std::thread>& p; std::atomic<bool>& stop; ~Joiner() { stop.store(true); for (auto& t : p) if (t.joinable()) t.join(); } } joiner{pool, stop};

    for (int job = 0; job < nJobs; ++job) {
      const double t_iter = seconds();
      Solved& s = solved[static_cast<size_t>(job)];
      if (side_by_side) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return s.done; });
        if (s.error == skipped) for (auto& o : solved) if (o.done && !o.error.empty() && o.error != skipped) throw lokihost::SetupError(o.error);
      } else solveJob(in, soa, job, n_devices, first_device, verbose != 0, status, s);
      if (!s.error.empty()) throw lokihost::SetupError(s.error);
      std::printf("%s", s.messages.c_str());
      const lokib200_solve_results res = s.res;
      const double t_rep = seconds();
      lokihost::Report rep(in, job, std::move(s.data));
      out.write(rep);
      if (g_profile) std::fprintf(stderr, "lokib200 job %d: report + output files in %.3f s, %.3f s since this job's turn began\n", job, seconds() - t_rep, seconds() - t_iter);
      s = Solved{};   // release the time series and histograms of this job
      if (verbose)
        std::printf("job %d/%d  %s = %g : mean energy %.6e eV (rel. err %.2e), %lld integration points, %.3e collisions, power balance %.2e, %.2f s, %.3e events/s\n", job + 1, nJobs,
                    in.wc.variableCondition.c_str(), in.jobValue(job), res.averaged_mean_energy, res.averaged_mean_energy_error / res.averaged_mean_energy,
                    static_cast<long long>(res.n_integration_points), res.total_collisions, res.power_balance_rel_error, res.elapsed_seconds, res.events_per_second);
      if (summary) {
        summary->n_jobs = job + 1; summary->last_mean_energy = res.averaged_mean_energy; summary->total_collisions += res.total_collisions + res.null_collisions;
        summary->device_seconds += res.elapsed_seconds;
      }
    }
    const double elapsed = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
    if (summary) summary->elapsed_seconds = elapsed;
    if (verbose) std::printf("Finished!\nElapsed time is %g seconds.\n", elapsed);   // Setup.h:948-957
    return LOKIB200_OK;
  } catch (const std::exception& e) { g_run_error = e.what(); return LOKIB200_ERR_INVALID; }
}

}  // extern "C"
