// boltzmann_mc.cpp -- host driver of one Monte Carlo job on top of the C ABI (include/lokib200.h).
//
// Mirrors the control flow of the reference's BoltzmannMC::evaluateEEDF (Code/LoKI-MC/Sources/BoltzmannMC.C, "BMC.C"):
// member and method names follow the reference so that the two can be read side by side; the code is written against the
// engine's result vector (ensemble SUMS per interval) instead of per-electron arrays, which is what lets several engines
// (GPUs) be combined by plain addition.  Only the C ABI is used: this file never touches CUDA.
#include "../../include/lokib200.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr double NON_DEF = LOKIB200_NON_DEF;
constexpr double QE = 1.6021766208e-19, ME = 9.10938356e-31;
constexpr double TWO_PI = 2.0 * 3.14159265358979323846;

using Vec3 = std::array<double, 3>;
using Mat9 = std::array<double, 9>;

// MathFunctions::statisticalError (MathFunctions.C:211-231): batch-means error of the average with nBins groups
template <class Get>
double statisticalError(int64_t first, int64_t n, int nBins, Get get) {
  if (n == 0) return 0;
  if (n < nBins) return NON_DEF;
  const int64_t nPoints = n / nBins;
  double mean = 0;
  for (int64_t i = 0; i < n; ++i) mean += get(first + i);
  mean /= static_cast<double>(n);
  double sum = 0;
  for (int b = 0; b < nBins; ++b) {
    double bm = 0;
    for (int64_t i = 0; i < nPoints; ++i) bm += get(first + b * nPoints + i);
    bm /= static_cast<double>(nPoints);
    sum += (bm - mean) * (bm - mean);
  }
  return std::sqrt(sum) / nBins;
}
template <class Get>
double meanOf(int64_t first, int64_t n, Get get) {
  double acc = 0;
  for (int64_t i = 0; i < n; ++i) acc += get(first + i);
  return n ? acc / static_cast<double>(n) : 0.0;
}

}  // namespace

struct lokib200_job {
  std::vector<lokib200_engine*> engines;
  lokib200_solve_controls ctl{};
  lokib200_config cfg{};
  std::string err;
  int P = 0, L = 0;
  double nElectrons = 0, totalGasDensity = 0;
  std::vector<double> relDensities;

  // --- state named after BMC.h ---
  double time = 0, steadyStateTime = NON_DEF, totalIntegratedTime = 0, trialCollisionFrequency = 0;
  double maxElecEnergy = 0, maxEedfEnergy = 0;
  int64_t nSamplingPoints = 0, nIntegrationPoints = 0, nSynchronizationPoints = 0, firstIntegrationIndex = 0, nTableRebuilds = 0;
  double totalCollisionCounter = 0, nullCollisionCounter = 0, collisionCounterAtSS = 0, nullCollisionCounterAtSS = 0, collisionCounterAfterSS = 0;
  double energyGainField = 0, energyGrowth = 0;
  std::vector<double> collisionCounters, energyGainProcesses, energyLossProcesses;
  std::vector<double> samplingTimes, meanEnergies;
  std::vector<Vec3> meanPositions, meanVelocities, bulkVelocities;
  std::vector<Mat9> positionCovariances, fluxDiffusionCoeffs, bulkDiffusionCoeffs;
  bool goodStatisticalErrors = false, stoppedByMaxCollisions = false;
  // time averages
  double averagedMeanEnergy = NON_DEF, averagedMeanEnergyError = NON_DEF, powerBalanceRelError = NON_DEF;
  Vec3 averagedFluxDriftVelocity{}, averagedFluxDriftVelocityError{}, averagedBulkDriftVelocity{}, averagedBulkDriftVelocityError{};
  Mat9 averagedFluxDiffusionCoeffs{}, averagedFluxDiffusionCoeffsError{}, averagedBulkDiffusionCoeffs{}, averagedBulkDiffusionCoeffsError{};
  std::vector<double> averagedRateCoeffs, averagedPowerGainProcesses, averagedPowerLossProcesses;
  double averagedPowerGainField = 0, averagedPowerGrowth = 0;
  // periodic (AC) accumulators, BMC.C:524-533, :1468-1481
  int nPhases = 0;
  double integrationPhaseStep = 0;
  std::vector<double> nIntegrationPointsPerPhase, meanEnergies_periodic;
  std::vector<Vec3> fluxVelocities_periodic, bulkVelocities_periodic;
  std::vector<Mat9> fluxDiffusionCoeffs_periodic, bulkDiffusionCoeffs_periodic;
  // histogram carry of earlier energy grids (BMC.C:1497-1548) and final sums
  bool histGridSet = false;
  double evdfMaxSpeed = 0;
  bool solved = false;
  std::vector<double> carryEeh, carryEah, carryEehPeriodic;
  double elapsed = 0;
  std::vector<double> res, tmp;
  double nuExceededTotal = 0, tableClampedTotal = 0;
  std::chrono::high_resolution_clock::time_point solveStart;
  bool firstStatus = true;
  int64_t nPointsBetweenStatErrorsCheck = 128;

  int fail(const std::string& m, int code = LOKIB200_ERR_INVALID) { err = m; return code; }
  int engineFail(lokib200_engine* e, int rc) { err = std::string("engine: ") + lokib200_last_error(e); return rc; }

  // ---- engine fan-out: the only "exchange" of the path is the sum / max of the per-interval result vectors ----
  void combine(bool first) {
    if (first) { res = tmp; return; }
    for (int j = 0; j < L; ++j) {
      if (j >= LOKIB200_R_SUM_COUNT && j < LOKIB200_R_HEADER) res[j] = std::max(res[j], tmp[j]);
      else res[j] += tmp[j];
    }
  }
  int buildTables(double maxEnergy) {
    for (auto* e : engines) { int rc = lokib200_build_tables(e, maxEnergy); if (rc) return engineFail(e, rc); }
    ++nTableRebuilds;
    return 0;
  }
  // checkMaxCollisionFrequency (BMC.C:716-763); horizon in units of 1/nu_trial
  int checkMaxCollisionFrequency(double horizon) {
    double nu = trialCollisionFrequency;
    double before = 0;
    lokib200_table_info(engines[0], nullptr, nullptr, &before, nullptr);
    for (size_t i = 0; i < engines.size(); ++i) {
      double nui = trialCollisionFrequency;
      int rc = lokib200_check_nu_trial(engines[i], maxElecEnergyNow, horizon, ctl.energy_max_elastic, &nui);
      if (rc) return engineFail(engines[i], rc);
      if (i == 0) nu = nui;   // identical inputs -> identical answers on every engine
    }
    double after = 0;
    lokib200_table_info(engines[0], nullptr, nullptr, &after, nullptr);
    if (after != before) ++nTableRebuilds;
    trialCollisionFrequency = nu;
    return 0;
  }
  double maxElecEnergyNow = 0;   // electronEnergies.maxCoeff() of the current ensemble (BMC.C:721)

  // With a communicator (lokib200_comm_init_all for the engines of this process, lokib200_comm_init_rank for one engine per process) the
  // result vectors are combined on the devices by one grouped NCCL all-reduce per interval and ONE vector is read back; every rank
  // of a multi-process job then takes identical decisions from identical numbers.  Without one, the engines are read one by one.
  bool useComm = false;
  int collect() {
    if (useComm) {
      int rc = lokib200_comm_allreduce_results(engines.data(), static_cast<int32_t>(engines.size()), nullptr);
      if (rc) return engineFail(engines[0], rc);
      if ((rc = lokib200_read_result(engines[0], res.data()))) return engineFail(engines[0], rc);
      return 0;
    }
    for (size_t i = 0; i < engines.size(); ++i) {
      int rc = lokib200_read_result(engines[i], tmp.data());
      if (rc) return engineFail(engines[i], rc);
      combine(i == 0);
    }
    return 0;
  }
  int advance(double tSync, bool sample) {
    if (engines.size() == 1 && !useComm) {   // the blocking call: for a small ensemble the whole interval is one CUDA graph launch (csrc/lokib200.cu, advance_graph)
      int rc = lokib200_advance_to_sync(engines[0], trialCollisionFrequency, tSync, sample ? 1 : 0, res.data());
      return rc ? engineFail(engines[0], rc) : 0;
    }
    // (the launches are asynchronous: all GPUs run concurrently; collect() waits for them)
    for (auto* e : engines) { int rc = lokib200_advance_to_sync_device(e, trialCollisionFrequency, tSync, sample ? 1 : 0, nullptr); if (rc) return engineFail(e, rc); }
    return collect();
  }
  int sampleNow() {
    for (auto* e : engines) { int rc = lokib200_sample_moments_device(e); if (rc) return engineFail(e, rc); }
    return collect();
  }

  // nonParallelCollisionTasks' accumulations (BMC.C:1303-1328) from the combined result vector
  void accumulateTallies() {
    totalCollisionCounter += res[LOKIB200_R_N_REAL];
    nullCollisionCounter += res[LOKIB200_R_N_NULL];
    energyGainField += res[LOKIB200_R_GAIN_FIELD];
    energyGrowth += res[LOKIB200_R_GROWTH];
    for (int k = 0; k < P; ++k) {
      collisionCounters[k] += res[LOKIB200_R_HEADER + k];
      energyGainProcesses[k] += res[LOKIB200_R_HEADER + P + k];
      energyLossProcesses[k] += res[LOKIB200_R_HEADER + 2 * P + k];
    }
  }

  // calculateMeanDataForSwarmParams (BMC.C:1410-1482) from the ensemble sums
  int calculateMeanDataForSwarmParams() {
    const double n = res[LOKIB200_R_N_SAMPLED];
    maxElecEnergyNow = res[LOKIB200_R_MAX_EPS];
    maxElecEnergy = std::fmax(maxElecEnergy, maxElecEnergyNow);                       // :1426
    if (maxElecEnergy > ctl.energy_max_elastic)                                        // :1427-1430
      return fail("The energy of the electrons reached " + std::to_string(maxElecEnergy) + " eV, while at least one of the elastic cross sections are defined only until " +
                  std::to_string(ctl.energy_max_elastic) + " eV. Please reduce the electric field or change the cross sections.", LOKIB200_ERR_ENERGY_RANGE);
    Vec3 mr, mv; Mat9 cov, fd;
    const double me = res[LOKIB200_R_SUM_EPS] / n;
    for (int a = 0; a < 3; ++a) { mr[a] = res[LOKIB200_R_SUM_R + a] / n; mv[a] = res[LOKIB200_R_SUM_V + a] / n; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      cov[3 * a + b] = res[LOKIB200_R_SUM_RR + 3 * a + b] / n - mr[a] * mr[b];       // :1445
      fd[3 * a + b] = res[LOKIB200_R_SUM_RV + 3 * a + b] / n - mr[a] * mv[b];        // :1446
    }
    samplingTimes.push_back(time); meanEnergies.push_back(me); meanPositions.push_back(mr); meanVelocities.push_back(mv);
    positionCovariances.push_back(cov); fluxDiffusionCoeffs.push_back(fd);
    const size_t i = samplingTimes.size() - 1;
    Vec3 bv; Mat9 bd;
    if (i != 0) {                                                                      // :1456-1459
      const double dt = samplingTimes[i] - samplingTimes[i - 1];
      for (int a = 0; a < 3; ++a) bv[a] = (mr[a] - meanPositions[i - 1][a]) / dt;
      for (int a = 0; a < 9; ++a) bd[a] = 0.5 * (cov[a] - positionCovariances[i - 1][a]) / dt;
    } else { bv = mv; bd = fd; }
    bulkVelocities.push_back(bv); bulkDiffusionCoeffs.push_back(bd);
    if (cfg.excitation_omega != 0 && steadyStateTime != NON_DEF) {                    // :1468-1481
      const int ph = phaseIndex();
      nIntegrationPointsPerPhase[ph] += 1.0; meanEnergies_periodic[ph] += me;
      for (int a = 0; a < 3; ++a) { fluxVelocities_periodic[ph][a] += mv[a]; bulkVelocities_periodic[ph][a] += bv[a]; }
      for (int a = 0; a < 9; ++a) { fluxDiffusionCoeffs_periodic[ph][a] += fd[a]; bulkDiffusionCoeffs_periodic[ph][a] += bd[a]; }
    }
    return 0;
  }
  int phaseIndex() const {                                                             // :1470-1472
    const double phase = std::fmod(cfg.excitation_omega * time, TWO_PI);
    return static_cast<int>(std::fmin(std::floor(phase / integrationPhaseStep), nPhases - 1.0));
  }

  // getTimeDependDistributions (BMC.C:1492-1572)
  int getTimeDependDistributions() {
    if (maxElecEnergy > maxEedfEnergy) {                                               // regrid :1497-1548
      const int nE = cfg.n_energy_cells, nC = cfg.n_cos_cells;
      std::vector<double> oldEeh(nE), oldEah(static_cast<size_t>(nE) * nC), oldPer(static_cast<size_t>(nPhases) * nE);
      int rc = fetchHistograms(oldEeh.data(), oldEah.data(), nullptr, oldPer.data());
      if (rc) return rc;
      const double oldStep = (0.0 + 1 * (maxEedfEnergy - 0.0) / static_cast<double>(nE)) - 0.0;
      maxEedfEnergy = 1.2 * maxElecEnergy;
      const double newStep = (0.0 + 1 * (maxEedfEnergy - 0.0) / static_cast<double>(nE)) - 0.0;
      std::fill(carryEeh.begin(), carryEeh.end(), 0.0); std::fill(carryEah.begin(), carryEah.end(), 0.0);
      std::fill(carryEehPeriodic.begin(), carryEehPeriodic.end(), 0.0);
      for (int pos = 0; pos < nE; ++pos) {
        const double left = 0.0 + pos * (oldStep * nE - 0.0) / static_cast<double>(nE), right = 0.0 + (pos + 1) * (oldStep * nE - 0.0) / static_cast<double>(nE);
        const int newLeft = static_cast<int>(left / newStep), newRight = static_cast<int>(right / newStep);
        auto spread = [&](double frac, int dst) {
          carryEeh[dst] += frac * oldEeh[pos];
          for (int c = 0; c < nC; ++c) carryEah[static_cast<size_t>(dst) * nC + c] += frac * oldEah[static_cast<size_t>(pos) * nC + c];
          // the reference adds the periodic columns nIntegrationPhases times (BMC.C:1523-1541); the physically
          // meaningful remap (once) is used here -- the difference only shows after a regrid in AC runs
          for (int p = 0; p < nPhases; ++p) carryEehPeriodic[static_cast<size_t>(p) * nE + dst] += frac * oldPer[static_cast<size_t>(p) * nE + pos];
        };
        if (newLeft == newRight) spread(1.0, newLeft);
        else {
          const double newRightNode = 0.0 + newRight * (maxEedfEnergy - 0.0) / static_cast<double>(nE);
          const double leftFraction = (newRightNode - left) / oldStep;
          spread(leftFraction, newLeft); spread(1.0 - leftFraction, newRight);
        }
      }
      for (auto* e : engines) { rc = lokib200_regrid_energy_histograms(e, maxEedfEnergy); if (rc) return engineFail(e, rc); }
    }
    const int ph = (cfg.excitation_omega != 0) ? phaseIndex() : -1;                    // :1554-1560
    for (auto* e : engines) { int rc = lokib200_sample_histograms(e, ph); if (rc) return engineFail(e, rc); }
    return 0;
  }
  // device counts of all engines + the carry of earlier grids
  int fetchHistograms(double* eeh, double* eah, double* evh, double* per) {
    const size_t nE = cfg.n_energy_cells, nEa = nE * cfg.n_cos_cells, nEv = static_cast<size_t>(cfg.n_radial_cells) * cfg.n_axial_cells, nPer = nE * nPhases;
    std::vector<double> a(eeh ? nE : 0), b(eah ? nEa : 0), c(evh ? nEv : 0), d(per ? nPer : 0);
    if (eeh) std::copy(carryEeh.begin(), carryEeh.end(), eeh);
    if (eah) std::copy(carryEah.begin(), carryEah.end(), eah);
    if (evh) std::fill(evh, evh + nEv, 0.0);
    if (per) std::copy(carryEehPeriodic.begin(), carryEehPeriodic.end(), per);
    if (useComm) {   // the counts of all shards, combined on the devices; engine 0 returns them
      int rc = lokib200_comm_allreduce_histograms(engines.data(), static_cast<int32_t>(engines.size()));
      if (rc) return engineFail(engines[0], rc);
    }
    for (auto* e : engines) {
      if (useComm && e != engines[0]) break;
      int rc = lokib200_fetch_histograms(e, eeh ? a.data() : nullptr, eah ? b.data() : nullptr, evh ? c.data() : nullptr, per ? d.data() : nullptr);
      if (rc) return engineFail(e, rc);
      if (eeh) for (size_t i = 0; i < nE; ++i) eeh[i] += a[i];
      if (eah) for (size_t i = 0; i < nEa; ++i) eah[i] += b[i];
      if (evh) for (size_t i = 0; i < nEv; ++i) evh[i] += c[i];
      if (per) for (size_t i = 0; i < nPer; ++i) per[i] += d[i];
    }
    return 0;
  }

  // checkSteadyState (BMC.C:1787-1893)
  int checkSteadyState() {
    const double t1 = 0.5 * time, t2 = 0.75 * time;
    double averageFirstPart = 0, averageSecondPart = 0, relStdSecondPart = 0;
    int64_t nFirst = 0, nSecond = 0;
    for (int64_t i = 0; i < nSamplingPoints; ++i) {
      const double st = samplingTimes[i];
      if (st >= t1 && st <= t2) { averageFirstPart += meanEnergies[i]; ++nFirst; }
      else if (st > t2) { const double v = meanEnergies[i]; averageSecondPart += v; relStdSecondPart += v * v; ++nSecond; }
    }
    averageFirstPart /= static_cast<double>(nFirst); averageSecondPart /= static_cast<double>(nSecond);
    relStdSecondPart = std::sqrt((relStdSecondPart / static_cast<double>(nSecond) - averageSecondPart * averageSecondPart) / static_cast<double>(nSecond));
    if ((averageFirstPart >= averageSecondPart && relStdSecondPart < 0.01) || totalCollisionCounter >= ctl.max_collisions_before_ss) {   // :1815
      firstIntegrationIndex = nSamplingPoints - 1; nIntegrationPoints = 1; steadyStateTime = time;
      collisionCounterAtSS = totalCollisionCounter; nullCollisionCounterAtSS = nullCollisionCounter;
      energyGainField = 0; energyGrowth = 0;                                            // :1850-1856
      std::fill(collisionCounters.begin(), collisionCounters.end(), 0.0);
      std::fill(energyGainProcesses.begin(), energyGainProcesses.end(), 0.0); std::fill(energyLossProcesses.begin(), energyLossProcesses.end(), 0.0);
      totalIntegratedTime = 0;
      maxEedfEnergy = 1.2 * maxElecEnergy;                                              // :1862
      for (auto* e : engines) { int rc = lokib200_set_histogram_grid(e, maxEedfEnergy); if (rc) return engineFail(e, rc); }
      histGridSet = true;
      evdfMaxSpeed = std::sqrt(2.0 * maxEedfEnergy * 1.6021766208e-19 / 9.10938356e-31);   // :1877, fixed from here on
      return getTimeDependDistributions();                                              // :1891
    }
    return 0;
  }

  // getTimeAverage* (BMC.C:1484-1490, 1607-1668)
  void timeAverages() {
    const int64_t f = firstIntegrationIndex, n = nIntegrationPoints;
    averagedMeanEnergy = meanOf(f, n, [&](int64_t i) { return meanEnergies[i]; });
    averagedMeanEnergyError = statisticalError(f, n, 50, [&](int64_t i) { return meanEnergies[i]; });
    for (int a = 0; a < 3; ++a) {
      averagedFluxDriftVelocity[a] = meanOf(f, n, [&](int64_t i) { return meanVelocities[i][a]; });
      averagedFluxDriftVelocityError[a] = statisticalError(f, n, 50, [&](int64_t i) { return meanVelocities[i][a]; });
      averagedBulkDriftVelocity[a] = meanOf(f, n, [&](int64_t i) { return bulkVelocities[i][a]; });
      averagedBulkDriftVelocityError[a] = statisticalError(f, n, 50, [&](int64_t i) { return bulkVelocities[i][a]; });
    }
    for (int a = 0; a < 9; ++a) {
      averagedFluxDiffusionCoeffs[a] = meanOf(f, n, [&](int64_t i) { return fluxDiffusionCoeffs[i][a]; });
      averagedFluxDiffusionCoeffsError[a] = statisticalError(f, n, 50, [&](int64_t i) { return fluxDiffusionCoeffs[i][a]; });
      averagedBulkDiffusionCoeffs[a] = meanOf(f, n, [&](int64_t i) { return bulkDiffusionCoeffs[i][a]; });
      averagedBulkDiffusionCoeffsError[a] = statisticalError(f, n, 50, [&](int64_t i) { return bulkDiffusionCoeffs[i][a]; });
    }
  }
  void checkPowerBalance() {                                                           // BMC.C:1769-1785
    double totalGain = energyGainField, totalLoss = 0;
    if (energyGrowth > 0) totalGain += energyGrowth; else totalLoss += energyGrowth;
    for (int k = 0; k < P; ++k) { totalGain += energyGainProcesses[k]; totalLoss += energyLossProcesses[k]; }
    powerBalanceRelError = std::fabs(totalGain + totalLoss) / totalGain;
  }
  void checkStatisticalErrors() {                                                      // BMC.C:1743-1767
    timeAverages();
    checkPowerBalance();
    auto relAbs = [](const Vec3& v, const Vec3& e) {
      return (std::fabs(v[0]) * e[0] + std::fabs(v[1]) * e[1] + std::fabs(v[2]) * e[2]) / (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    };
    const auto& fD = averagedFluxDiffusionCoeffs; const auto& fE = averagedFluxDiffusionCoeffsError;
    const auto& bD = averagedBulkDiffusionCoeffs; const auto& bE = averagedBulkDiffusionCoeffsError;
    if (averagedMeanEnergyError / averagedMeanEnergy <= ctl.rel_err_mean_energy && relAbs(averagedFluxDriftVelocity, averagedFluxDriftVelocityError) <= ctl.rel_err_flux_drift &&
        fE[0] / fD[0] <= ctl.rel_err_flux_diff && fE[4] / fD[4] <= ctl.rel_err_flux_diff && fE[8] / fD[8] <= ctl.rel_err_flux_diff &&
        relAbs(averagedBulkDriftVelocity, averagedBulkDriftVelocityError) <= ctl.rel_err_bulk_drift && bE[0] / bD[0] <= ctl.rel_err_bulk_diff &&
        bE[4] / bD[4] <= ctl.rel_err_bulk_diff && bE[8] / bD[8] <= ctl.rel_err_bulk_diff && powerBalanceRelError <= ctl.rel_err_power_balance)
      goodStatisticalErrors = true;
  }

  // dispInfo (BMC.C:1895-1948): the reference's status table, plus two lines of the engine's own (events per second and the fraction of the
  // device's nominal HBM bandwidth that 128 bytes per event at this rate amount to)
  void dispInfo() {
    if (firstStatus) { firstStatus = false; std::printf("\n*********Simulation status*********\n\n"); }
    else { std::printf(" \n"); for (int i = 0; i < 32; ++i) std::printf("\033[1A\033[2K"); }
    std::printf("E/N: %g Td\nexcitFreq: %g Hz\nE-field angle: %g degrees\nB/N: %g Hx\n\n", ctl.status_values[0], ctl.status_values[1], ctl.status_values[2], ctl.status_values[3]);
    std::printf("Number of real collisions: %g\nNumber of null collisions: %g\n\n", totalCollisionCounter, nullCollisionCounter);
    const double secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - solveStart).count();
    const double rate = secs > 0 ? (totalCollisionCounter + nullCollisionCounter) / secs : 0.0;
    const double hbm = lokib200_device_hbm_gbs(engines[0]) * static_cast<double>(engines.size());
    std::printf("Collision events per second: %.4g\nHBM roofline fraction (128 B per event at sync factor %g): %.3f\n\n", rate, ctl.sync_factor,
                hbm > 0 ? rate * 128.0 / ctl.sync_factor / (hbm * 1e9) : 0.0);
    std::printf("Current time: %g s\n", time);
    int blank = 18;
    if (steadyStateTime == NON_DEF) { std::printf("Mean energy: %g\n", meanEnergies.empty() ? 0.0 : meanEnergies.back()); blank = 19; }
    else {
      std::printf("Steady-state time: %g s\n\n", steadyStateTime);
      if (nIntegrationPoints >= 3 * nPointsBetweenStatErrorsCheck && averagedMeanEnergy != NON_DEF) {
        blank = 0;
        const auto& fv = averagedFluxDriftVelocity; const auto& fe = averagedFluxDriftVelocityError;
        const auto& bv = averagedBulkDriftVelocity; const auto& be = averagedBulkDriftVelocityError;
        const auto& fD = averagedFluxDiffusionCoeffs; const auto& fE = averagedFluxDiffusionCoeffsError;
        const auto& bD = averagedBulkDiffusionCoeffs; const auto& bE = averagedBulkDiffusionCoeffsError;
        std::printf("Number of integration points: %g\n\nMean energy [eV]: %g\nRelative error: %g\n\n", static_cast<double>(nIntegrationPoints), averagedMeanEnergy,
                    averagedMeanEnergyError / averagedMeanEnergy);
        std::printf("Flux drift velocity [m/s]: %g %g %g\nRelative error: %g %g %g\n\n", fv[0], fv[1], fv[2], fe[0] / std::fabs(fv[0]), fe[1] / std::fabs(fv[1]), fe[2] / std::fabs(fv[2]));
        std::printf("Bulk drift velocity [m/s]: %g %g %g\nRelative error: %g %g %g\n\n", bv[0], bv[1], bv[2], be[0] / std::fabs(bv[0]), be[1] / std::fabs(bv[1]), be[2] / std::fabs(bv[2]));
        std::printf("Flux diffusion coefficients [m^2 s^-1]: %g %g %g\nRelative error: %g %g %g\n\n", fD[0], fD[4], fD[8], fE[0] / fD[0], fE[4] / fD[4], fE[8] / fD[8]);
        std::printf("Bulk diffusion coefficients [m^2 s^-1]: %g %g %g\nRelative error: %g %g %g\n\n", bD[0], bD[4], bD[8], bE[0] / bD[0], bE[4] / bD[4], bE[8] / bD[8]);
        std::printf("Power balance relative error: %g\n", powerBalanceRelError);
      }
    }
    for (int i = 0; i < blank; ++i) std::printf("\n");
    std::fflush(stdout);
  }

  // evaluateEEDF (BMC.C:299-426)
  int evaluateEEDF() {
    const auto start = std::chrono::high_resolution_clock::now();
    solveStart = start;
    // ---- evaluateNonConstantVariables (BMC.C:428-559) ----
    time = 0; steadyStateTime = NON_DEF;
    for (auto* e : engines) lokib200_set_fast_mode(e, ctl.fast_mode);
    double maxInit = 0;
    for (auto* e : engines) {
      double mx = 0;
      int rc = lokib200_init_ensemble(e, ctl.initial_temp_ratio, &mx);
      if (rc) return engineFail(e, rc);
      maxInit = std::max(maxInit, mx);
    }
    maxElecEnergyNow = maxInit; maxElecEnergy = maxInit;                               // :537
    int rc = buildTables(2.0 * maxInit);                                               // :512
    if (rc) return rc;
    double nuLast = 0;
    lokib200_table_info(engines[0], nullptr, nullptr, nullptr, &nuLast);
    trialCollisionFrequency = nuLast;                                                  // :515
    // ---- t = 0 sample (:306-310) ----
    nSamplingPoints = 1; nSynchronizationPoints = 1; collisionCounterAfterSS = 0; nIntegrationPoints = 0; totalIntegratedTime = 0;
    if ((rc = sampleNow())) return rc;
    nElectrons = res[LOKIB200_R_N_SAMPLED];   // all shards of all ranks
    if ((rc = calculateMeanDataForSwarmParams())) return rc;
    if (ctl.status_display) dispInfo();                                                // :313-315
    const int over = std::max(1, ctl.sync_over_sampling);
    // ---- main loop (:320-384) ----
    while ((!goodStatisticalErrors && ctl.errors_to_be_checked) || static_cast<double>(nIntegrationPoints) < ctl.n_integration_points ||
           totalIntegratedTime / steadyStateTime < ctl.n_integrated_ss_times || totalIntegratedTime < ctl.integrated_absolute_time) {
      if (collisionCounterAfterSS >= ctl.max_collisions_after_ss && nIntegrationPoints > 100) { stoppedByMaxCollisions = true; break; }
      if (ctl.max_intervals > 0 && nSynchronizationPoints > ctl.max_intervals) break;
      // electronDynamicsUntilSynchronization (:617-688): one launch advances the whole interval, so the energy bound that the
      // reference re-derives before every micro-pass (:634) is taken once over interval + 10 mean free times
      if ((rc = checkMaxCollisionFrequency(ctl.sync_factor + 10.0))) return rc;
      const double nextSynchronizTime = time + ctl.sync_factor / trialCollisionFrequency;   // :622-623
      ++nSynchronizationPoints;
      const bool sample = (nSynchronizationPoints % over == 0);
      if ((rc = advance(nextSynchronizTime, sample))) return rc;
      time = nextSynchronizTime;
      accumulateTallies();
      maxElecEnergyNow = std::max(res[LOKIB200_R_MAX_EPS], res[LOKIB200_R_MAX_EPS_SEEN]);
      // The reference re-checks the trial frequency before every micro-pass (BMC.C:634); here one bound serves a whole interval, and these two
      // counters are what tells that it was not enough: some electron met nu_tot(eps) above its trial frequency, or an energy beyond the
      // tables.  A few stragglers are normal while the ensemble heats up (an electron keeps the frequency its free time was drawn with until
      // its next event, also in the reference: BMC.C:650-655); more than one event in a thousand means the bound itself failed: raise the
      // frequency as checkMaxCollisionFrequency does (:758-761) / let the next check rebuild the tables from the energy seen.
      nuExceededTotal += res[LOKIB200_R_N_NU_EXCEEDED]; tableClampedTotal += res[LOKIB200_R_N_TABLE_CLAMPED];
      if (res[LOKIB200_R_N_NU_EXCEEDED] > 1e-3 * (res[LOKIB200_R_N_REAL] + res[LOKIB200_R_N_NULL])) trialCollisionFrequency *= 1.1;
      if (res[LOKIB200_R_N_TABLE_CLAMPED] > 0) { double top = 0; lokib200_table_info(engines[0], nullptr, nullptr, &top, nullptr); maxElecEnergyNow = std::max(maxElecEnergyNow, top); }
      if (!sample) continue;
      ++nSamplingPoints;                                                               // :334-341
      if ((rc = calculateMeanDataForSwarmParams())) return rc;
      maxElecEnergyNow = std::max(maxElecEnergyNow, res[LOKIB200_R_MAX_EPS_SEEN]);
      const int dec1 = static_cast<int>(std::fmax(std::log(static_cast<double>(nSamplingPoints)) / std::log(2.0) - 11, 6));
      const int64_t nPointsBetweenSteadyStateCheck = static_cast<int64_t>(std::pow(2, dec1));   // :344-345
      if (steadyStateTime != NON_DEF) {                                                // :348-360
        ++nIntegrationPoints;
        collisionCounterAfterSS = totalCollisionCounter - collisionCounterAtSS;
        totalIntegratedTime = time - steadyStateTime;
        if ((rc = getTimeDependDistributions())) return rc;
      } else if (nSamplingPoints >= 100 && nSamplingPoints % nPointsBetweenSteadyStateCheck == 0 && totalCollisionCounter > ctl.min_collisions_before_ss) {
        if ((rc = checkSteadyState())) return rc;                                      // :363-365
      }
      const int dec2 = static_cast<int>(std::fmax(std::log(static_cast<double>(std::max<int64_t>(nIntegrationPoints, 1))) / std::log(2.0) - 9, 7));
      nPointsBetweenStatErrorsCheck = static_cast<int64_t>(std::pow(2, dec2));                  // :369-370
      if (steadyStateTime != NON_DEF && nIntegrationPoints > 200 && nIntegrationPoints % nPointsBetweenStatErrorsCheck == 0) checkStatisticalErrors();
      if (ctl.status_display && nSamplingPoints % nPointsBetweenSteadyStateCheck == 0) dispInfo();   // :377-381
    }
    // ---- time averages (:387-396) ----
    totalIntegratedTime = time - steadyStateTime;
    timeAverages();
    for (int k = 0; k < P; ++k) {                                                      // getTimeAverageRateCoeffs :1630-1637, PowerBalance :1643-1648
      averagedRateCoeffs[k] = (relDensities[k] != 0) ? collisionCounters[k] / (totalIntegratedTime * nElectrons * relDensities[k] * totalGasDensity) : 0.0;
      averagedPowerGainProcesses[k] = energyGainProcesses[k] / (nElectrons * totalIntegratedTime * totalGasDensity);
      averagedPowerLossProcesses[k] = energyLossProcesses[k] / (nElectrons * totalIntegratedTime * totalGasDensity);
    }
    averagedPowerGainField = energyGainField / (nElectrons * totalIntegratedTime * totalGasDensity);
    averagedPowerGrowth = energyGrowth / (nElectrons * totalIntegratedTime * totalGasDensity);
    if (cfg.excitation_omega != 0) {                                                   // getAveragedPeriodicParams :1728-1734
      for (int p = 0; p < nPhases; ++p) {
        const double c = nIntegrationPointsPerPhase[p];
        meanEnergies_periodic[p] /= c;
        for (int a = 0; a < 3; ++a) { fluxVelocities_periodic[p][a] /= c; bulkVelocities_periodic[p][a] /= c; }
        for (int a = 0; a < 9; ++a) { fluxDiffusionCoeffs_periodic[p][a] /= c; bulkDiffusionCoeffs_periodic[p][a] /= c; }
      }
    }
    checkStatisticalErrors();                                                          // :418
    if (ctl.status_display) dispInfo();                                                // :421-423
    elapsed = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
    return 0;
  }
};

extern "C" {

int lokib200_job_create(lokib200_engine* const* engines, int32_t n_engines, const lokib200_solve_controls* c, lokib200_job** out) {
  if (!engines || n_engines <= 0 || !c || !out) return LOKIB200_ERR_INVALID;
  auto* j = new lokib200_job();
  j->engines.assign(engines, engines + n_engines);
  j->ctl = *c;
  if (j->ctl.sync_factor <= 0) j->ctl.sync_factor = 1.0;
  if (j->ctl.initial_temp_ratio <= 0) j->ctl.initial_temp_ratio = 0.01;
  if (j->ctl.energy_max_elastic <= 0) j->ctl.energy_max_elastic = 1e100;
  if (j->ctl.max_collisions_before_ss <= 0) j->ctl.max_collisions_before_ss = 1e100;
  if (j->ctl.max_collisions_after_ss <= 0) j->ctl.max_collisions_after_ss = 1e100;
  for (double* r : {&j->ctl.rel_err_mean_energy, &j->ctl.rel_err_flux_drift, &j->ctl.rel_err_flux_diff, &j->ctl.rel_err_bulk_drift, &j->ctl.rel_err_bulk_diff,
                    &j->ctl.rel_err_power_balance})
    if (*r <= 0) *r = 1e100;
  lokib200_get_config(engines[0], &j->cfg);
  j->P = lokib200_process_count(engines[0]);
  if (j->P <= 0) { delete j; return LOKIB200_ERR_INVALID; }
  j->L = LOKIB200_RESULT_LEN(j->P);
  j->nElectrons = 0;
  for (auto* e : j->engines) { lokib200_config ce; lokib200_get_config(e, &ce); j->nElectrons += static_cast<double>(ce.n_electrons); }
  j->useComm = lokib200_comm_size(engines[0]) > 1;
  for (auto* e : j->engines) if ((lokib200_comm_size(e) > 1) != j->useComm) { delete j; return LOKIB200_ERR_INVALID; }
  j->totalGasDensity = j->cfg.gas_density;
  j->relDensities.resize(j->P);
  lokib200_get_rel_densities(engines[0], j->relDensities.data());
  j->collisionCounters.assign(j->P, 0.0); j->energyGainProcesses.assign(j->P, 0.0); j->energyLossProcesses.assign(j->P, 0.0);
  j->averagedRateCoeffs.assign(j->P, 0.0); j->averagedPowerGainProcesses.assign(j->P, 0.0); j->averagedPowerLossProcesses.assign(j->P, 0.0);
  j->res.assign(j->L, 0.0); j->tmp.assign(j->L, 0.0);
  j->nPhases = j->cfg.n_phases;
  j->integrationPhaseStep = TWO_PI / j->nPhases;                                        // BMC.C:525
  j->nIntegrationPointsPerPhase.assign(j->nPhases, 0.0); j->meanEnergies_periodic.assign(j->nPhases, 0.0);
  j->fluxVelocities_periodic.assign(j->nPhases, Vec3{}); j->bulkVelocities_periodic.assign(j->nPhases, Vec3{});
  j->fluxDiffusionCoeffs_periodic.assign(j->nPhases, Mat9{}); j->bulkDiffusionCoeffs_periodic.assign(j->nPhases, Mat9{});
  j->carryEeh.assign(j->cfg.n_energy_cells, 0.0);
  j->carryEah.assign(static_cast<size_t>(j->cfg.n_energy_cells) * j->cfg.n_cos_cells, 0.0);
  j->carryEehPeriodic.assign(static_cast<size_t>(j->nPhases) * j->cfg.n_energy_cells, 0.0);
  *out = j;
  return 0;
}

static void fillResults(const lokib200_job* j, lokib200_solve_results* r) {
  {
    std::memset(r, 0, sizeof(*r));
    r->averaged_mean_energy = j->averagedMeanEnergy; r->averaged_mean_energy_error = j->averagedMeanEnergyError;
    for (int a = 0; a < 3; ++a) {
      r->flux_drift_velocity[a] = j->averagedFluxDriftVelocity[a]; r->flux_drift_velocity_error[a] = j->averagedFluxDriftVelocityError[a];
      r->bulk_drift_velocity[a] = j->averagedBulkDriftVelocity[a]; r->bulk_drift_velocity_error[a] = j->averagedBulkDriftVelocityError[a];
    }
    for (int a = 0; a < 9; ++a) {
      r->flux_diffusion[a] = j->averagedFluxDiffusionCoeffs[a]; r->flux_diffusion_error[a] = j->averagedFluxDiffusionCoeffsError[a];
      r->bulk_diffusion[a] = j->averagedBulkDiffusionCoeffs[a]; r->bulk_diffusion_error[a] = j->averagedBulkDiffusionCoeffsError[a];
    }
    r->power_gain_field = j->averagedPowerGainField; r->power_growth = j->averagedPowerGrowth; r->power_balance_rel_error = j->powerBalanceRelError;
    r->time = j->time; r->steady_state_time = j->steadyStateTime; r->total_integrated_time = j->totalIntegratedTime;
    r->trial_collision_frequency = j->trialCollisionFrequency; r->max_eedf_energy = j->maxEedfEnergy; r->elapsed_seconds = j->elapsed;
    r->total_collisions = j->totalCollisionCounter; r->null_collisions = j->nullCollisionCounter; r->collisions_at_ss = j->collisionCounterAtSS;
    r->null_collisions_at_ss = j->nullCollisionCounterAtSS;
    r->n_sampling_points = j->nSamplingPoints; r->n_integration_points = j->nIntegrationPoints; r->n_sync_points = j->nSynchronizationPoints;
    r->n_table_rebuilds = j->nTableRebuilds;
    r->good_statistical_errors = j->goodStatisticalErrors; r->stopped_by_max_collisions = j->stoppedByMaxCollisions;
    r->n_nu_exceeded = j->nuExceededTotal; r->n_table_clamped = j->tableClampedTotal;
    r->events_per_second = j->elapsed > 0 ? (j->totalCollisionCounter + j->nullCollisionCounter) / j->elapsed : 0.0;
  }
}

int lokib200_job_solve(lokib200_job* j, lokib200_solve_results* r) {
  if (!j) return LOKIB200_ERR_INVALID;
  const int rc = j->evaluateEEDF();
  j->solved = (rc == 0);
  if (r) fillResults(j, r);
  return rc;
}

int lokib200_job_process_outputs(const lokib200_job* j, double* rate, double* gain, double* loss, double* counts) {
  if (!j) return LOKIB200_ERR_INVALID;
  if (rate) std::copy(j->averagedRateCoeffs.begin(), j->averagedRateCoeffs.end(), rate);
  if (gain) std::copy(j->averagedPowerGainProcesses.begin(), j->averagedPowerGainProcesses.end(), gain);
  if (loss) std::copy(j->averagedPowerLossProcesses.begin(), j->averagedPowerLossProcesses.end(), loss);
  if (counts) std::copy(j->collisionCounters.begin(), j->collisionCounters.end(), counts);
  return 0;
}

int64_t lokib200_job_time_series(const lokib200_job* j, double* times, double* me, double* mp, double* mv, double* pc) {
  if (!j) return 0;
  const int64_t n = static_cast<int64_t>(j->samplingTimes.size());
  for (int64_t i = 0; i < n; ++i) {
    if (times) times[i] = j->samplingTimes[i];
    if (me) me[i] = j->meanEnergies[i];
    if (mp) for (int a = 0; a < 3; ++a) mp[3 * i + a] = j->meanPositions[i][a];
    if (mv) for (int a = 0; a < 3; ++a) mv[3 * i + a] = j->meanVelocities[i][a];
    if (pc) for (int a = 0; a < 9; ++a) pc[9 * i + a] = j->positionCovariances[i][a];
  }
  return n;
}

int lokib200_job_histograms(lokib200_job* j, double* eeh, double* eah, double* evh, double* per) {
  if (!j || !j->histGridSet) return LOKIB200_ERR_INVALID;
  return j->fetchHistograms(eeh, eah, evh, per);
}

int lokib200_job_periodic(const lokib200_job* j, double* pts, double* me, double* fv, double* bv) {
  if (!j) return LOKIB200_ERR_INVALID;
  for (int p = 0; p < j->nPhases; ++p) {
    if (pts) pts[p] = j->nIntegrationPointsPerPhase[p];
    if (me) me[p] = j->meanEnergies_periodic[p];
    if (fv) for (int a = 0; a < 3; ++a) fv[3 * p + a] = j->fluxVelocities_periodic[p][a];
    if (bv) for (int a = 0; a < 3; ++a) bv[3 * p + a] = j->bulkVelocities_periodic[p][a];
  }
  return 0;
}

int lokib200_job_periodic_diffusion(const lokib200_job* j, double* fd, double* bd) {
  if (!j) return LOKIB200_ERR_INVALID;
  for (int p = 0; p < j->nPhases; ++p) for (int a = 0; a < 9; ++a) {
    if (fd) fd[9 * p + a] = j->fluxDiffusionCoeffs_periodic[p][a];
    if (bd) bd[9 * p + a] = j->bulkDiffusionCoeffs_periodic[p][a];
  }
  return 0;
}

int lokib200_job_results(const lokib200_job* j, lokib200_solve_results* r) {
  if (!j || !r || !j->solved) return LOKIB200_ERR_INVALID;
  fillResults(j, r);
  return 0;
}

double lokib200_job_evdf_max_speed(const lokib200_job* j) { return j ? j->evdfMaxSpeed : 0.0; }

int lokib200_job_conditions(const lokib200_job* j, lokib200_config* cfg, int32_t* n_processes) {
  if (!j) return LOKIB200_ERR_INVALID;
  if (cfg) { *cfg = j->cfg; cfg->n_electrons = static_cast<int64_t>(j->nElectrons); }
  if (n_processes) *n_processes = j->P;
  return 0;
}

const char* lokib200_job_last_error(const lokib200_job* j) { return j ? j->err.c_str() : "null job"; }
void lokib200_job_destroy(lokib200_job* j) { delete j; }

}  // extern "C"
