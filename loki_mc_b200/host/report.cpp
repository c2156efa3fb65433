// report.cpp -- see report.h.  "BMC.C" = Sources/BoltzmannMC.C, "Output.h" = Headers/Output.h of /root/reference/Code/LoKI-MC.
#include "report.h"

#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstring>

namespace lokihost {

namespace {

constexpr double NON_DEF = LOKIB200_NON_DEF;
constexpr double QE = 1.6021766208e-19, ME = 9.10938356e-31;
constexpr double PI = 3.14159265358979323846;

// Eigen::ArrayXd::LinSpaced(n, low, high) (Eigen 3.4 linspaced_op)
std::vector<double> linSpaced(int n, double low, double high) {
  std::vector<double> v(static_cast<size_t>(n));
  if (n == 1) { v[0] = high; return v; }
  const double step = (high - low) / static_cast<double>(n - 1);
  const bool flip = std::fabs(high) < std::fabs(low);
  for (int i = 0; i < n; ++i)
    v[i] = flip ? (i == 0 ? low : high - static_cast<double>(n - 1 - i) * step) : (i == n - 1 ? high : low + static_cast<double>(i) * step);
  return v;
}

// Eigen's vectorised linear reduction of a contiguous double array (SSE2 packets of 2, two packet accumulators)
double packetSum(const double* v, size_t n) {
  if (n == 0) return 0.0;
  const size_t aligned = (n / 2) * 2, aligned2 = (n / 4) * 4;
  if (aligned == 0) return v[0];
  double a0 = v[0], a1 = v[1];
  if (aligned > 2) {
    double b0 = v[2], b1 = v[3];
    for (size_t i = 4; i < aligned2; i += 4) { a0 += v[i]; a1 += v[i + 1]; b0 += v[i + 2]; b1 += v[i + 3]; }
    a0 += b0; a1 += b1;
    if (aligned > aligned2) { a0 += v[aligned2]; a1 += v[aligned2 + 1]; }
  }
  double r = a0 + a1;
  for (size_t i = aligned; i < n; ++i) r += v[i];
  return r;
}
double packetSum(const std::vector<double>& v) { return packetSum(v.data(), v.size()); }

void mkdirs(const std::string& path) {
  for (size_t i = 1; i <= path.size(); ++i)
    if (i == path.size() || path[i] == '/') ::mkdir(path.substr(0, i).c_str(), 0777);
}

struct File {
  FILE* f;
  File(const std::string& name, const char* mode) : f(std::fopen(name.c_str(), mode)) {
    if (!f) throw SetupError("The file '" + name + "' could not be created. Please check the correspondent directory.");
  }
  ~File() { if (f) std::fclose(f); }
  operator FILE*() const { return f; }
};

void rotate(const double R[9], const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
void rotate2(const double R[9], const double M[9], double out[9]) {   // R M R^T
  double t[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3 * i + j] = R[3 * i] * M[j] + R[3 * i + 1] * M[3 + j] + R[3 * i + 2] * M[6 + j];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[3 * i + j] = t[3 * i] * R[3 * j] + t[3 * i + 1] * R[3 * j + 1] + t[3 * i + 2] * R[3 * j + 2];
}

}  // namespace

// ------------------------------------------------------------------ post-processing ------------------------------------------------------------------
Report::Report(const SetupInput& input, int jobIndex, JobData data) : in(input), job(jobIndex), d(std::move(data)) {
  cfg = in.config(job);
  auto pick = [&](const std::vector<double>& a, const char* name) { return a[(in.wc.variableCondition == name) ? static_cast<size_t>(job) : 0]; };
  reducedElecField = pick(in.wc.reducedElecFieldArray, "reducedElecField"); reducedMagField = pick(in.wc.reducedMagFieldArray, "reducedMagField");
  elecFieldAngle = pick(in.wc.elecFieldAngleArray, "elecFieldAngle"); excitationFrequency = pick(in.wc.excitationFrequencyArray, "excitationFrequency");
  nE = cfg.n_energy_cells; nCos = cfg.n_cos_cells; nR = cfg.n_radial_cells; nA = cfg.n_axial_cells;
  nPh = (cfg.excitation_omega != 0) ? cfg.n_phases : 0;
  const size_t P = in.processes.type.size();
  if (d.rateCoeffsMC.size() != P || d.powerGain.size() != P || d.powerLoss.size() != P) throw SetupError("report: per-process arrays do not match the process set");
  if (d.eehSum.size() != static_cast<size_t>(nE)) throw SetupError("report: energy histogram size does not match numericsMC.nEnergyCells");
  if (cfg.is_cylindrically_symmetric && (d.eahSum.size() != static_cast<size_t>(nE) * nCos || d.evhSum.size() != static_cast<size_t>(nR) * nA))
    throw SetupError("report: angular / velocity histogram sizes do not match the setup");
  if (nPh && d.eehSumPeriodic.size() != static_cast<size_t>(nPh) * nE) throw SetupError("report: periodic histogram size does not match the setup");
  distributions();
  adjustCrossSections();
  evaluatePower();
  evaluateRateCoeffs();
  evaluateSwarm();
}

void Report::distributions() {   // BMC.C:1574-1604, :1727-1741; grids of :1862-1884 and Grid.C:46-55
  const double maxEnergy = d.res.max_eedf_energy;
  energyStep = maxEnergy / nE;
  energyNode = linSpaced(nE + 1, 0.0, maxEnergy);
  energyCell = linSpaced(nE, energyStep / 2.0, maxEnergy - energyStep / 2.0);
  const double eedfStep = energyNode[1];
  const double normalizer = packetSum(d.eehSum);
  eedf.assign(nE, 0.0); efadf.assign(nE, 0.0); esadf.assign(nE, 0.0);
  for (int i = 0; i < nE; ++i) eedf[i] = d.eehSum[i] / (std::sqrt(energyCell[i]) * normalizer * eedfStep);
  if (cfg.is_cylindrically_symmetric) {
    const auto cosNodes = linSpaced(nCos + 1, -1.0, 1.0);
    const double cosStep = cosNodes[1] + 1.0;
    cosCells = linSpaced(nCos, -1.0 + cosStep / 2.0, 1.0 - cosStep / 2.0);
    eadf.assign(static_cast<size_t>(nE) * nCos, 0.0);
    for (int i = 0; i < nE; ++i) {
      const double den = std::sqrt(energyCell[i]) * normalizer * eedfStep * cosStep;
      double s1 = 0, s2 = 0;
      for (int j = 0; j < nCos; ++j) {
        const double a = 2.0 * d.eahSum[static_cast<size_t>(i) * nCos + j] / den;
        eadf[static_cast<size_t>(i) * nCos + j] = a;
        s1 += a * cosCells[j];
        s2 += a * (0.5 * (3.0 * (cosCells[j] * cosCells[j]) - 1));
      }
      efadf[i] = (1 + 0.5) * s1 * cosStep;
      esadf[i] = (2 + 0.5) * s2 * cosStep;
    }
    const double maxSpeed = d.evdfMaxSpeed;
    const auto rNodes = linSpaced(nR + 1, 0.0, maxSpeed);
    const double rStep = rNodes[1];
    radialCells = linSpaced(nR, rStep / 2.0, maxSpeed - rStep / 2.0);
    const auto aNodes = linSpaced(nA + 1, -maxSpeed, maxSpeed);
    const double aStep = aNodes[1] + maxSpeed;
    axialCells = linSpaced(nA, -maxSpeed + aStep / 2.0, maxSpeed - aStep / 2.0);
    double total = 0;
    for (int i = 0; i < nR; ++i) {
      double row = 0;
      for (int j = 0; j < nA; ++j) row += d.evhSum[static_cast<size_t>(i) * nA + j];
      total += row * aStep * 2.0 * PI * radialCells[i] * rStep;
    }
    evdf.resize(d.evhSum.size());
    for (size_t k = 0; k < evdf.size(); ++k) evdf[k] = d.evhSum[k] / total;
  }
  if (nPh) {
    const double phaseStep = 2.0 * PI / nPh;
    integrationPhases = linSpaced(nPh, 0.5 * phaseStep, 2.0 * PI - 0.5 * phaseStep);
    eedfPeriodic.assign(static_cast<size_t>(nPh) * nE, 0.0);
    for (int p = 0; p < nPh; ++p) {
      double norm = 0;
      for (int i = 0; i < nE; ++i) norm += d.eehSumPeriodic[static_cast<size_t>(p) * nE + i];
      for (int i = 0; i < nE; ++i) eedfPeriodic[static_cast<size_t>(p) * nE + i] = d.eehSumPeriodic[static_cast<size_t>(p) * nE + i] / (std::sqrt(energyCell[i]) * norm * eedfStep);
    }
  }
}

void Report::adjustCrossSections() {   // Collision::adjustCrossSection / reAdjustCrossSection, Collision.C:201-230, :398-418
  const auto& cols = in.mixture->collisions;
  xs_.assign(cols.size(), GridXS{});
  for (const auto& cp : cols) {
    const Collision& c = *cp;
    if (c.rawIntegral.empty()) continue;
    if (c.type == "Elastic" && energyNode.back() > c.rawIntegral.e.back())
      throw SetupError("''" + c.description() + "'' cross section data is not available for the maximum energy of the simulation (" + std::to_string(energyNode.back()) +
                       " eV).\nSimulation is not reliable under these conditions.");
    GridXS& g = xs_[c.id];
    g.integral = interpolatedCrossSection(c, false, energyNode);
    g.momTransf = (c.angularType == "isotropic" || c.rawMomTransf.empty()) ? g.integral : interpolatedCrossSection(c, true, energyNode);
  }
}

void Report::evaluateRate(const Collision& c, const std::vector<double>& f) {   // Collision::evaluateRateCoeff, Collision.C:355-396
  GridXS& g = xs_[c.id];
  const double factor = std::sqrt(2.0 * QE / ME);
  const int lmin = static_cast<int>(std::floor(c.threshold / energyStep));
  const int partial = static_cast<int>(g.integral.size()) - 1 - lmin;
  if (partial <= 0) { g.ine = 0; if (c.isReverse) g.sup = 0; return; }
  std::vector<double> ine(partial), sup(partial);
  for (int j = 0; j < partial; ++j) {
    const double aux = (g.integral[lmin + j] + g.integral[lmin + j + 1]) / 2.0 * energyCell[lmin + j];
    ine[j] = aux * f[lmin + j];
    sup[j] = aux * f[j];
  }
  g.ine = factor * packetSum(ine) * energyStep;
  if (c.isReverse) {
    if (c.target->statisticalWeight == NON_DEF) throw SetupError("Unable to find '" + c.target->name + "' statistical weight for the evaluation of superelastic rate coefficient of " + c.description() + "\n");
    if (c.products[0]->statisticalWeight == NON_DEF) throw SetupError("Unable to find '" + c.products[0]->name + "' statistical weight for the evaluation of superelastic rate coefficient of " + c.description() + "\n");
    g.sup = factor * (c.target->statisticalWeight / c.products[0]->statisticalWeight) * packetSum(sup) * energyStep;
  }
}

void Report::evaluatePower() {   // BMC.C:1950-2063
  static const char* keys[] = {"field", "elasticNet", "elasticGain", "elasticLoss", "carNet", "carGain", "carLoss", "excitationIne", "excitationSup", "excitationNet",
                               "vibrationalIne", "vibrationalSup", "vibrationalNet", "rotationalIne", "rotationalSup", "rotationalNet", "ionizationIne", "attachmentIne",
                               "inelastic", "superelastic", "eDensGrowth", "electronElectron"};
  for (const char* k : keys) power[k] = 0;
  const ProcessSet& ps = in.processes;
  const size_t P = ps.type.size();
  power["field"] = d.res.power_gain_field;
  for (size_t i = 0; i < P; ++i)
    if (ps.collisionOf[i]->type == "Elastic") { power["elasticGain"] += d.powerGain[i]; power["elasticLoss"] += d.powerLoss[i]; }
  power["elasticNet"] = power["elasticGain"] + power["elasticLoss"];
  for (const auto& gas : in.mixture->gases) {
    std::map<std::string, double> gp;
    for (const char* k : {"excitationIne", "excitationSup", "excitationNet", "vibrationalIne", "vibrationalSup", "vibrationalNet", "rotationalIne", "rotationalSup",
                          "rotationalNet", "ionizationIne", "attachmentIne"}) gp[k] = 0;
    for (size_t i = 0; i < P; ++i) {
      const Collision* c = ps.collisionOf[i];
      if (c->target->gas != gas.get()) continue;
      const double net = d.powerGain[i] + d.powerLoss[i];
      const bool sup = ps.isSuperelastic[i] != 0;
      if (c->type == "Excitation") gp[sup ? "excitationSup" : "excitationIne"] += net;
      else if (c->type == "Vibrational") gp[sup ? "vibrationalSup" : "vibrationalIne"] += net;
      else if (c->type == "Rotational") gp[sup ? "rotationalSup" : "rotationalIne"] += net;
      else if (c->type == "Ionization") gp["ionizationIne"] += net;
      else if (c->type == "Attachment") gp["attachmentIne"] += net;
    }
    gp["excitationNet"] = gp["excitationIne"] + gp["excitationSup"];
    gp["vibrationalNet"] = gp["vibrationalIne"] + gp["vibrationalSup"];
    gp["rotationalNet"] = gp["rotationalIne"] + gp["rotationalSup"];
    gp["inelastic"] = gp["excitationIne"] + gp["vibrationalIne"] + gp["rotationalIne"] + gp["ionizationIne"] + gp["attachmentIne"];
    gp["superelastic"] = gp["excitationSup"] + gp["vibrationalSup"] + gp["rotationalSup"];
    for (const char* k : {"excitationIne", "excitationSup", "vibrationalIne", "vibrationalSup", "rotationalIne", "rotationalSup", "ionizationIne", "attachmentIne"}) power[k] += gp[k];
    powerByGas[gas->name] = gp;
  }
  power["excitationNet"] = power["excitationIne"] + power["excitationSup"];
  power["vibrationalNet"] = power["vibrationalIne"] + power["vibrationalSup"];
  power["rotationalNet"] = power["rotationalIne"] + power["rotationalSup"];
  power["inelastic"] = power["excitationIne"] + power["vibrationalIne"] + power["rotationalIne"] + power["ionizationIne"] + power["attachmentIne"];
  power["superelastic"] = power["excitationSup"] + power["vibrationalSup"] + power["rotationalSup"];
  power["eDensGrowth"] = d.res.power_growth;
  power["balance"] = power["field"] + power["elasticNet"] + power["inelastic"] + power["superelastic"] + power["eDensGrowth"];
  double totalGain = 0;
  for (const char* k : {"field", "elasticGain", "elasticLoss", "excitationSup", "excitationIne", "vibrationalSup", "vibrationalIne", "rotationalSup", "rotationalIne", "eDensGrowth"})
    if (power[k] > 0) totalGain += power[k];
  power["relativeBalance"] = std::fabs(power["balance"]) / totalGain;
  power["reference"] = totalGain;
}

void Report::evaluateRateCoeffs() {   // BMC.C:2273-2384
  const ProcessSet& ps = in.processes;
  const int P = static_cast<int>(ps.type.size());
  auto pass = [&](const std::vector<double>& f, bool withMC, std::vector<RateCoeff>& all, std::vector<RateCoeff>& extra) {
    for (int i = 0; i < P; ++i) {
      RateCoeff rc;
      if (!ps.isSuperelastic[i]) {
        const Collision& c = *ps.collisionOf[i];
        rc.collID = c.id;
        evaluateRate(c, f);
        rc.ineRate = xs_[c.id].ine; rc.supRate = xs_[c.id].sup;
        if (withMC) rc.ineRateMC = d.rateCoeffsMC[i];
        if (rc.supRate != NON_DEF) { ++i; if (withMC) rc.supRateMC = d.rateCoeffsMC[i]; }
        rc.description = c.description();
      }
      all.push_back(rc);
    }
    for (const auto& gas : in.mixture->gases) {
      RateCoeff rc;
      for (const Collision* c : gas->collisions) {
        if (c->type != "Effective") continue;
        rc.collID = c->id; evaluateRate(*c, f); rc.ineRate = xs_[c->id].ine; rc.supRate = xs_[c->id].sup; rc.description = c->description();
        all.push_back(rc);
      }
      for (const Collision* c : gas->collisionsExtra) {
        rc.collID = c->id; evaluateRate(*c, f); rc.ineRate = xs_[c->id].ine; rc.supRate = xs_[c->id].sup; rc.description = c->description();
        extra.push_back(rc);
      }
    }
  };
  for (int p = 0; p < nPh; ++p) {
    std::vector<double> f(eedfPeriodic.begin() + static_cast<size_t>(p) * nE, eedfPeriodic.begin() + static_cast<size_t>(p + 1) * nE);
    rateAllPeriodic.emplace_back(); rateExtraPeriodic.emplace_back();
    pass(f, false, rateAllPeriodic.back(), rateExtraPeriodic.back());
  }
  pass(eedf, true, rateAll, rateExtra);
}

void Report::evaluateSwarm() {   // BMC.C:2065-2271
  const double N = cfg.gas_density;
  const double reducedElecFieldSI = reducedElecField * 1e-21;
  const double rotAngle = -elecFieldAngle / 180.0 * PI;
  const double R[9] = {std::cos(rotAngle), 0, std::sin(rotAngle), 0, 1, 0, -std::sin(rotAngle), 0, std::cos(rotAngle)};
  double A[9];
  for (int i = 0; i < 9; ++i) A[i] = std::fabs(R[i]);
  rotate(R, d.res.flux_drift_velocity, rotFluxV); rotate(A, d.res.flux_drift_velocity_error, rotFluxVErr);
  rotate2(R, d.res.flux_diffusion, rotFluxD); rotate2(A, d.res.flux_diffusion_error, rotFluxDErr);
  swarm["fluxRedTransvDiffCoeff"] = N * (rotFluxD[0] + rotFluxD[4]) / 2.0;
  swarm["fluxRedTransvDiffCoeffError"] = N * (rotFluxDErr[0] + rotFluxDErr[4]) / 2.0;
  swarm["fluxRedLongDiffCoeff"] = N * rotFluxD[8];
  swarm["fluxRedLongDiffCoeffError"] = N * rotFluxDErr[8];
  swarm["totalIonRateCoeff"] = 0; swarm["totalAttRateCoeff"] = 0;
  for (const auto& gas : in.mixture->gases) for (const Collision* c : gas->collisions) {
    if (c->type == "Ionization") swarm["totalIonRateCoeff"] += c->target->density * xs_[c->id].ine;
    else if (c->type == "Attachment") swarm["totalAttRateCoeff"] += c->target->density * xs_[c->id].ine;
  }
  if (excitationFrequency == 0) {
    const double v = std::fabs(rotFluxV[2]);
    swarm["fluxRedMobCoeff"] = v / reducedElecFieldSI;
    swarm["fluxRedMobCoeffError"] = std::fabs(rotFluxVErr[2]) / reducedElecFieldSI;
    swarm["fluxRedTownsendCoeff"] = swarm["totalIonRateCoeff"] / v;
    swarm["fluxRedAttCoeff"] = swarm["totalAttRateCoeff"] / v;
    swarm["fluxCharacEnergy"] = swarm["fluxRedTransvDiffCoeff"] / swarm["fluxRedMobCoeff"];
    swarm["fluxCharacEnergyError"] = swarm["fluxRedTransvDiffCoeffError"] / swarm["fluxRedMobCoeff"] +
                                     swarm["fluxRedMobCoeffError"] * swarm["fluxRedTransvDiffCoeff"] / std::pow(swarm["fluxRedMobCoeff"], 2);
  }
  rotate(R, d.res.bulk_drift_velocity, rotBulkV); rotate(A, d.res.bulk_drift_velocity_error, rotBulkVErr);
  rotate2(R, d.res.bulk_diffusion, rotBulkD); rotate2(A, d.res.bulk_diffusion_error, rotBulkDErr);
  swarm["bulkRedTransvDiffCoeff"] = N * (rotBulkD[0] + rotBulkD[4]) / 2.0;
  swarm["bulkRedTransvDiffCoeffError"] = N * (rotBulkDErr[0] + rotBulkDErr[4]) / 2.0;
  swarm["bulkRedLongDiffCoeff"] = N * rotBulkD[8];
  swarm["bulkRedLongDiffCoeffError"] = N * rotBulkDErr[8];
  if (excitationFrequency == 0) {
    const double v = std::fabs(rotBulkV[2]);
    swarm["bulkRedMobCoeff"] = v / reducedElecFieldSI;
    swarm["bulkRedMobCoeffError"] = std::fabs(rotBulkVErr[2]) / reducedElecFieldSI;
    swarm["bulkRedTownsendCoeff"] = swarm["totalIonRateCoeff"] / v;
    swarm["bulkRedAttCoeff"] = swarm["totalAttRateCoeff"] / v;
    swarm["bulkCharacEnergy"] = swarm["bulkRedTransvDiffCoeff"] / swarm["bulkRedMobCoeff"];
    swarm["bulkCharacEnergyError"] = swarm["bulkRedTransvDiffCoeffError"] / swarm["bulkRedMobCoeff"] +
                                     swarm["bulkRedMobCoeffError"] * swarm["bulkRedTransvDiffCoeff"] / std::pow(swarm["bulkRedMobCoeff"], 2);
    swarm["effSSTAverageVelocity"] = 0.5 * v + std::sqrt(0.25 * v * v - N * rotBulkD[8] * (swarm["totalIonRateCoeff"] - swarm["totalAttRateCoeff"]));
    swarm["effSSTRedTownsendCoeff"] = swarm["totalIonRateCoeff"] / swarm["effSSTAverageVelocity"];
    swarm["effSSTRedAttCoeff"] = swarm["totalAttRateCoeff"] / swarm["effSSTAverageVelocity"];
  }
  swarm["meanEnergy"] = d.res.averaged_mean_energy; swarm["meanEnergyError"] = d.res.averaged_mean_energy_error;
  swarm["Te"] = 2.0 / 3.0 * swarm["meanEnergy"]; swarm["TeError"] = 2.0 / 3.0 * swarm["meanEnergyError"];

  // two-term expressions on the EEDF (:2182-2270)
  const double factor = std::sqrt(2.0 * QE / ME) / 3.0;
  const int Nn = nE + 1, Nc = nE;
  std::vector<double> mt(Nn, 0.0), er(Nn, 0.0), erClassic(Nn, 0.0);
  for (const auto& gas : in.mixture->gases) {
    if (gas->collisions.empty()) continue;
    const double massRatio = ME / gas->get("mass");
    for (const Collision* c : gas->collisions) {
      if (c->type == "Effective") continue;
      const GridXS& g = xs_[c->id];
      const double dens = c->target->density;
      for (int i = 0; i < Nn; ++i) mt[i] += g.momTransf[i] * dens;
      if (c->type == "Elastic") { for (int i = 0; i < Nn; ++i) er[i] += 2.0 * massRatio * g.integral[i] * dens; }
      else {
        for (int i = 0; i < Nn; ++i) {
          const double w = g.integral[i] * dens;
          er[i] += (i == 0) ? 0.0 : w * c->threshold / energyNode[i];
          erClassic[i] += w;
        }
      }
      if (c->isReverse) {
        const double pd = c->products[0]->density;
        const auto supI = superElasticCrossSection(*c, false, energyNode), supM = superElasticCrossSection(*c, true, energyNode);
        for (int i = 0; i < Nn; ++i) {
          const double w = supI[i] * pd;
          mt[i] += supM[i] * pd;
          er[i] += (i == 0) ? 0.0 : w * c->threshold / energyNode[i];
          erClassic[i] += w;
        }
      }
    }
  }
  auto cellIntegral = [&](const std::vector<double>& nodeXS) {
    std::vector<double> t(Nc);
    for (int i = 0; i < Nc; ++i) t[i] = eedf[i] * energyCell[i] * ((nodeXS[i + 1] + nodeXS[i]) * 0.5);
    return N * 3.0 * factor * energyStep * packetSum(t);
  };
  swarm["momTransfFreq"] = cellIntegral(mt);
  swarm["energyRelaxFreq"] = cellIntegral(er);
  swarm["energyRelaxFreqClassic"] = cellIntegral(erClassic);
  const double growth = swarm["totalIonRateCoeff"] - swarm["totalAttRateCoeff"];
  std::vector<double> mtAux = mt, freqAux(Nn);
  for (int i = 1; i < Nn; ++i) mtAux[i] += 1.0 / (3.0 * factor * std::sqrt(energyNode[i])) * growth;
  for (int i = 0; i < Nn; ++i) freqAux[i] = N * mtAux[i] * 3.0 * factor * std::sqrt(energyNode[i]);
  std::vector<double> t(Nc);
  for (int i = 0; i < Nc; ++i) t[i] = energyCell[i] * energyCell[i] * eedf[i] / (mtAux[i] + mtAux[i + 1]);
  swarm["redDiffCoeffEnergy_eedf"] = 2.0 * factor * energyStep * packetSum(t);
  for (int i = 0; i < Nc; ++i) t[i] = energyCell[i] * eedf[i] / (mtAux[i] + mtAux[i + 1]);
  swarm["redDiffCoeff_eedf"] = 2.0 * factor * energyStep * packetSum(t);
  const int M = Nn - 2;   // interior nodes 1 .. Nn-2
  std::vector<double> u(M > 0 ? M : 0);
  for (int i = 0; i < M; ++i) u[i] = energyNode[i + 1] * energyNode[i + 1] * (eedf[i + 1] - eedf[i]) / mtAux[i + 1];
  swarm["redMobCoeffEnergy_eedf"] = -factor * packetSum(u);
  for (int i = 0; i < M; ++i) u[i] = energyNode[i + 1] * (eedf[i + 1] - eedf[i]) / mtAux[i + 1];
  swarm["redMobCoeff_DC_eedf"] = -factor * packetSum(u);
  swarm["characEnergy_eedf"] = swarm["redDiffCoeff_eedf"] / swarm["redMobCoeff_DC_eedf"];
  // reduced mobility matrix (AC / magnetised), :2252-2270
  const double w = cfg.excitation_omega, wc = cfg.cyclotron_omega;
  double sxxr = 0, sxxi = 0, sxyr = 0, sxyi = 0, szzr = 0, szzi = 0;
  for (int i = 0; i < M; ++i) {
    const double aux1 = -2.0 / 3.0 * QE / ME * std::pow(energyNode[i + 1], 1.5) * (eedf[i + 1] - eedf[i]);
    const double nu = freqAux[i + 1], nu2 = nu * nu;
    const double aux2 = (nu2 + std::pow(w - wc, 2)) * (nu2 + std::pow(w + wc, 2)), aux3 = nu2 + w * w;
    sxxr += nu * (nu2 + w * w + wc * wc) / aux2 * aux1;
    sxxi += w * (nu2 + w * w - wc * wc) / aux2 * aux1;
    sxyr += (nu2 - w * w + wc * wc) / aux2 * aux1;
    sxyi += nu / aux2 * aux1;
    szzr += nu / aux3 * aux1;
    szzi += w / aux3 * aux1;
  }
  swarm["redMobCoeff_xx_real_eedf"] = N * sxxr; swarm["redMobCoeff_xx_imag_eedf"] = -N * sxxi;
  swarm["redMobCoeff_xy_real_eedf"] = -N * wc * sxyr; swarm["redMobCoeff_xy_imag_eedf"] = N * 2.0 * w * wc * sxyi;
  swarm["redMobCoeff_zz_real_eedf"] = N * szzr; swarm["redMobCoeff_zz_imag_eedf"] = -N * szzi;
}

// ------------------------------------------------------------------ text sinks (Output.h) ------------------------------------------------------------------
OutputWriter::OutputWriter(const SetupInput& in, const std::string& outputRoot) : in_(in) {   // Output.h:46-115
  const SetupTree& t = *in.tree;
  const std::string on = t.value("output.isOn");
  enabled = (on == "true" || on == "True" || on == "1");
  if (!enabled) return;
  folder = outputRoot + "/" + (t.has("output.folder") ? t.value("output.folder") : std::string("lokib200_output"));
  mkdirs(folder);
  for (const auto& f : t.childNames("output.dataFiles")) {
    if (f == "eedf") eedf_ = true; else if (f == "evdf") evdf_ = true; else if (f == "powerBalance") power_ = true;
    else if (f == "swarmParameters") swarm_ = true; else if (f == "rateCoefficients") rates_ = true;
    else if (f == "lookUpTable" && in.nJobs() > 1) lookUp_ = true; else if (f == "MCTemporalInfo") temporal_ = true;
    else if (f == "MCTemporalInfo_periodic") temporalPeriodic_ = true; else if (f == "MCSimDetails") details_ = true;
  }
  File f(folder + "/setup.txt", "w");
  const std::string dump = t.dump();
  std::fwrite(dump.data(), 1, dump.size(), f);
}

std::string OutputWriter::subFolder(const Report& r) {   // Output.h:151-178
  if (in_.nJobs() <= 1) return "";
  char cond[100];
  std::snprintf(cond, sizeof cond, "%g", in_.jobValue(r.job));
  const std::string sub = "/" + in_.wc.variableCondition + "_" + cond;
  mkdirs(folder + sub);
  return sub;
}

void OutputWriter::write(const Report& r) {   // Output.h:117-149
  if (!enabled) return;
  ++currentJob_;
  const std::string dir = folder + subFolder(r);
  if (eedf_) saveEedf(r, dir);
  if (evdf_ && r.cfg.is_cylindrically_symmetric) saveEvdf(r, dir);
  if (power_) savePower(r, dir);
  if (swarm_) saveSwarm(r, dir);
  if (rates_) saveRateCoefficients(r, dir);
  if (lookUp_) saveLookUpTable(r);
  if (temporal_) saveMCTemporalInfo(r, dir);
  if (temporalPeriodic_) saveMCTemporalInfoPeriodic(r, dir);
  if (details_) saveMCSimDetails(r, dir);
}

void OutputWriter::saveEedf(const Report& r, const std::string& dir) {   // Output.h:180-221
  {
    File f(dir + "/eedf.txt", "w");
    if (!r.cfg.is_cylindrically_symmetric) {
      std::fprintf(f, "Energy(eV)           EEDF(eV^-(3/2))\n");
      for (int i = 0; i < r.nE; ++i) std::fprintf(f, "%-20.14e %-20.14e \n", r.energyCell[i], r.eedf[i]);
    } else {
      std::fprintf(f, "Energy(eV)           EEDF(eV^-(3/2))      First Anisotropy     Second Anisotropy    \n");
      for (int i = 0; i < r.nE; ++i) std::fprintf(f, "%-20.14e %-20.14e %-20.14e %-20.14e\n", r.energyCell[i], r.eedf[i], r.efadf[i], r.esadf[i]);
    }
  }
  if (r.excitationFrequency != 0) {
    File f(dir + "/eedf_periodic.txt", "w");
    for (int p = 0; p < r.nPh; ++p) {
      std::fprintf(f, "%-20.14e ", r.integrationPhases[p]);
      for (int i = 0; i < r.nE; ++i) std::fprintf(f, "%-20.14e ", r.eedfPeriodic[static_cast<size_t>(p) * r.nE + i]);
      std::fprintf(f, "\n");
    }
  }
}

void OutputWriter::saveEvdf(const Report& r, const std::string& dir) {   // Output.h:223-237 (the reference's loop assumes nRadial == nAxial)
  File f(dir + "/evdf.txt", "w");
  std::fprintf(f, "v_r(m/s)           v_z(m/s)           EVDF(m-3s-3)       \n");
  for (int i = 0; i < r.nR; ++i)
    for (int j = 0; j < r.nA; ++j) std::fprintf(f, "%-18.10e %-18.10e %-18.10e\n", r.radialCells[i], r.axialCells[j], r.evdf[static_cast<size_t>(i) * r.nA + j]);
}

namespace {
void velocityBlock(FILE* f, const double v[3], const double err[3], bool rotated) {
  double rel[3];
  for (int i = 0; i < 3; ++i) rel[i] = err[i] / std::fabs(v[i]) * 100.0;
  if (!rotated) {
    std::fprintf(f, "                                   | v_x |   | %-15.8e |                       | %-9.3e%% |\n", v[0], rel[0]);
    std::fprintf(f, "                                   | v_y | = | %-15.8e |  (m/s)    ; Rel. std: | %-9.3e%% |\n", v[1], rel[1]);
    std::fprintf(f, "                                   | v_z |   | %-15.8e |                       | %-9.3e%% |\n\n", v[2], rel[2]);
  } else {
    std::fprintf(f, "                                  | v_x' |   | %-15.8e |                       | %-9.3e%% |\n", v[0], rel[0]);
    std::fprintf(f, "                                  | v_y' | = | %-15.8e |  (m/s)    ; Rel. std: | %-9.3e%% |\n", v[1], rel[1]);
    std::fprintf(f, "                                  | v_z' |   | %-15.8e |                       | %-9.3e%% |\n\n", v[2], rel[2]);
  }
}
void diffusionBlock(FILE* f, const double D[9], const double err[9], double N, bool rotated) {
  double nd[9], rel[9];
  for (int i = 0; i < 9; ++i) { nd[i] = D[i] * N; rel[i] = err[i] / std::fabs(D[i]) * 100.0; }
  const char* l0 = rotated ? "                | ND_x'x' ND_x'y' ND_x'z'|   " : "                      | ND_xx ND_xy ND_xz|   ";
  const char* l1 = rotated ? "                | ND_y'x' ND_y'y' ND_y'z'| = " : "                      | ND_yx ND_yy ND_yz| = ";
  const char* l2 = rotated ? "                | ND_z'x' ND_z'y' ND_z'z'|   " : "                      | ND_zx ND_zy ND_zz|   ";
  std::fprintf(f, "%s| %-15.8e %-15.8e %-15.8e |                      | %-9.3e%% %-9.3e%% %-9.3e%% |\n", l0, nd[0], nd[1], nd[2], rel[0], rel[1], rel[2]);
  std::fprintf(f, "%s| %-15.8e %-15.8e %-15.8e | (1/(ms)) ; Rel. std: | %-9.3e%% %-9.3e%% %-9.3e%% |\n", l1, nd[3], nd[4], nd[5], rel[3], rel[4], rel[5]);
  std::fprintf(f, "%s| %-15.8e %-15.8e %-15.8e |                      | %-9.3e%% %-9.3e%% %-9.3e%% |\n\n", l2, nd[6], nd[7], nd[8], rel[6], rel[7], rel[8]);
}
double at(const std::map<std::string, double>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0.0 : it->second; }
std::string banner(int n, const std::string& text) { return std::string(n, '*') + " " + text + " " + std::string(n, '*'); }
}  // namespace

void OutputWriter::saveSwarm(const Report& r, const std::string& dir) {   // Output.h:239-345
  File f(dir + "/swarmParameters.txt", "w");
  const auto& s = r.swarm;
  const double N = r.cfg.gas_density;
  std::fprintf(f, "                    Reduced electric field = %#.14e (Td)\n", r.reducedElecField);
  std::fprintf(f, "                      Electric field angle = %#.14e (Degrees)\n", r.elecFieldAngle);
  std::fprintf(f, "                      Excitation frequency = %#.14e (Hz)\n", r.excitationFrequency);
  std::fprintf(f, "                    Reduced magnetic field = %#.14e (Hx)\n", r.reducedMagField);
  for (int pass = 0; pass < 2; ++pass) {
    const bool flux = pass == 0;
    const char* pre = flux ? "flux" : "bulk";
    auto key = [&](const char* k) { return at(s, (std::string(pre) + k).c_str()); };
    std::fprintf(f, "\n%s\n\n", banner(35, flux ? "Flux parameters" : "Bulk parameters").c_str());
    velocityBlock(f, flux ? r.d.res.flux_drift_velocity : r.d.res.bulk_drift_velocity, flux ? r.d.res.flux_drift_velocity_error : r.d.res.bulk_drift_velocity_error, false);
    diffusionBlock(f, flux ? r.d.res.flux_diffusion : r.d.res.bulk_diffusion, flux ? r.d.res.flux_diffusion_error : r.d.res.bulk_diffusion_error, N, false);
    std::fprintf(f, "Parameters after rotation to a ref. frame (x'y'z'), where E-field is along z'\n\n");
    velocityBlock(f, flux ? r.rotFluxV : r.rotBulkV, flux ? r.rotFluxVErr : r.rotBulkVErr, true);
    diffusionBlock(f, flux ? r.rotFluxD : r.rotBulkD, flux ? r.rotFluxDErr : r.rotBulkDErr, N, true);
    std::fprintf(f, "  Reduced transverse diffusion coefficient = %#.14e (1/(ms)) ; Rel. std: %-9.3e%%\n", key("RedTransvDiffCoeff"), key("RedTransvDiffCoeffError") / key("RedTransvDiffCoeff") * 100.0);
    std::fprintf(f, "Reduced longitudinal diffusion coefficient = %#.14e (1/(ms)) ; Rel. std: %-9.3e%%\n", key("RedLongDiffCoeff"), key("RedLongDiffCoeffError") / key("RedLongDiffCoeff") * 100.0);
    if (r.reducedMagField == 0 && r.excitationFrequency == 0) {
      std::fprintf(f, "              Reduced mobility coefficient = %#.14e (1/(msV)); Rel. std: %-9.3e%%\n", key("RedMobCoeff"), key("RedMobCoeffError") / key("RedMobCoeff") * 100.0);
      std::fprintf(f, "                     Characteristic energy = %#.14e (eV)     ; Rel. std: %-9.3e%%\n", key("CharacEnergy"), key("CharacEnergyError") / key("CharacEnergy") * 100.0);
    }
    if (r.excitationFrequency == 0) {
      std::fprintf(f, "              Reduced Townsend coefficient = %#.14e (m2)\n", key("RedTownsendCoeff"));
      std::fprintf(f, "            Reduced attachment coefficient = %#.14e (m2)\n", key("RedAttCoeff"));
    }
  }
  std::fprintf(f, "\n%s\n\n", banner(15, "Effective SST parameters deduced from the TOF simulation").c_str());
  std::fprintf(f, "                     SST averaged velocity = %#.14e (m/s)\n", at(s, "effSSTAverageVelocity"));
  std::fprintf(f, "          SST reduced Townsend coefficient = %#.14e (m/s)\n", at(s, "effSSTRedTownsendCoeff"));
  std::fprintf(f, "        SST reduced attachment coefficient = %#.14e (m/s)\n", at(s, "effSSTRedAttCoeff"));
  std::fprintf(f, "\n%s\n\n", banner(35, "Energy parameters").c_str());
  std::fprintf(f, "                               Mean energy = %#.14e (eV) ; Rel. std: %-9.3e%%\n", at(s, "meanEnergy"), at(s, "meanEnergyError") / at(s, "meanEnergy") * 100.0);
  std::fprintf(f, "                      Electron temperature = %#.14e (eV) ; Rel. std: %-9.3e%%\n", at(s, "Te"), at(s, "TeError") / at(s, "Te") * 100.0);
  std::fprintf(f, "\n%s\n\n", banner(27, "Parameters obtained from the EEDF").c_str());
  std::fprintf(f, "                    Ionization coefficient = %#.14e (m-3)\n", at(s, "totalIonRateCoeff"));
  std::fprintf(f, "                    Attachment coefficient = %#.14e (m-3)\n", at(s, "totalAttRateCoeff"));
  std::fprintf(f, "               Momentum-transfer frequency = %#.14e (s-1)\n", at(s, "momTransfFreq"));
  std::fprintf(f, "               Energy-relaxation frequency = %#.14e (s-1)\n", at(s, "energyRelaxFreq"));
  std::fprintf(f, "      Reduced energy diffusion coefficient = %#.14e (eV/(ms))\n", at(s, "redDiffCoeffEnergy_eedf"));
  std::fprintf(f, "                   Reduced energy mobility = %#.14e (eV/(msV))\n", at(s, "redMobCoeffEnergy_eedf"));
  std::fprintf(f, "             Reduced diffusion coefficient = %#.14e (1/(ms))\n", at(s, "redDiffCoeff_eedf"));
  std::fprintf(f, "        DC non-magnetized reduced mobility = %#.14e (1/(msV))\n", at(s, "redMobCoeff_DC_eedf"));
  std::fprintf(f, "                     Characteristic energy = %#.14e (eV)\n\n", at(s, "characEnergy_eedf"));
}

void OutputWriter::savePower(const Report& r, const std::string& dir) {   // Output.h:347-417
  File f(dir + "/powerBalance.txt", "w");
  const auto& p = r.power;
  const std::string rule(73, '-');
  std::fprintf(f, "                               Field = %#+.14e (eVm3/s)\n", at(p, "field"));
  std::fprintf(f, "           Elastic collisions (gain) = %#+.14e (eVm3/s)\n", at(p, "elasticGain"));
  std::fprintf(f, "           Elastic collisions (loss) = %#+.14e (eVm3/s)\n", at(p, "elasticLoss"));
  std::fprintf(f, "                          CAR (gain) = %#+.14e (eVm3/s)\n", at(p, "carGain"));
  std::fprintf(f, "                          CAR (loss) = %#+.14e (eVm3/s)\n", at(p, "carLoss"));
  std::fprintf(f, "     Excitation inelastic collisions = %#+.14e (eVm3/s)\n", at(p, "excitationIne"));
  std::fprintf(f, "  Excitation superelastic collisions = %#+.14e (eVm3/s)\n", at(p, "excitationSup"));
  std::fprintf(f, "    Vibrational inelastic collisions = %#+.14e (eVm3/s)\n", at(p, "vibrationalIne"));
  std::fprintf(f, " Vibrational superelastic collisions = %#+.14e (eVm3/s)\n", at(p, "vibrationalSup"));
  std::fprintf(f, "     Rotational inelastic collisions = %#+.14e (eVm3/s)\n", at(p, "rotationalIne"));
  std::fprintf(f, "  Rotational superelastic collisions = %#+.14e (eVm3/s)\n", at(p, "rotationalSup"));
  std::fprintf(f, "               Ionization collisions = %#+.14e (eVm3/s)\n", at(p, "ionizationIne"));
  std::fprintf(f, "               Attachment collisions = %#+.14e (eVm3/s)\n", at(p, "attachmentIne"));
  std::fprintf(f, "             Electron density growth = %#+.14e (eVm3/s) +\n", at(p, "eDensGrowth"));
  std::fprintf(f, " %s\n", rule.c_str());
  std::fprintf(f, "                       Power Balance = %#+.14e (eVm3/s)\n", at(p, "balance"));
  std::fprintf(f, "              Relative Power Balance = %#.14e%%\n\n", at(p, "relativeBalance") * 100.0);
  std::fprintf(f, "           Elastic collisions (gain) = %#+.14e (eVm3/s)\n", at(p, "elasticGain"));
  std::fprintf(f, "           Elastic collisions (loss) = %#+.14e (eVm3/s) +\n", at(p, "elasticLoss"));
  std::fprintf(f, " %s\n", rule.c_str());
  std::fprintf(f, "            Elastic collisions (net) = %#+.14e (eVm3/s)\n\n", at(p, "elasticNet"));
  std::fprintf(f, "                          CAR (gain) = %#+.14e (eVm3/s)\n", at(p, "carGain"));
  std::fprintf(f, "                          CAR (gain) = %#+.14e (eVm3/s) +\n", at(p, "carLoss"));
  std::fprintf(f, " %s\n", rule.c_str());
  std::fprintf(f, "                           CAR (net) = %#+.14e (eVm3/s)\n\n", at(p, "carNet"));
  auto triple = [&](const std::map<std::string, double>& m, bool lastBlank) {
    std::fprintf(f, "     Excitation inelastic collisions = %#+.14e (eVm3/s)\n", at(m, "excitationIne"));
    std::fprintf(f, "  Excitation superelastic collisions = %#+.14e (eVm3/s) +\n", at(m, "excitationSup"));
    std::fprintf(f, " %s\n", rule.c_str());
    std::fprintf(f, "         Excitation collisions (net) = %#+.14e (eVm3/s)\n\n", at(m, "excitationNet"));
    std::fprintf(f, "    Vibrational inelastic collisions = %#+.14e (eVm3/s)\n", at(m, "vibrationalIne"));
    std::fprintf(f, " Vibrational superelastic collisions = %#+.14e (eVm3/s) +\n", at(m, "vibrationalSup"));
    std::fprintf(f, " %s\n", rule.c_str());
    std::fprintf(f, "        Vibrational collisions (net) = %#+.14e (eVm3/s)\n\n", at(m, "vibrationalNet"));
    std::fprintf(f, "     Rotational inelastic collisions = %#+.14e (eVm3/s)\n", at(m, "rotationalIne"));
    std::fprintf(f, "  Rotational superelastic collisions = %#+.14e (eVm3/s) +\n", at(m, "rotationalSup"));
    std::fprintf(f, " %s\n", rule.c_str());
    std::fprintf(f, lastBlank ? "         Rotational collisions (net) = %#+.14e (eVm3/s)\n\n" : "         Rotational collisions (net) = %#+.14e (eVm3/s)\n", at(m, "rotationalNet"));
  };
  triple(p, false);
  for (const auto& kv : r.powerByGas) {
    const std::string& gas = kv.first;
    std::fprintf(f, "\n%s\n\n", (std::string(37, '*') + " " + gas + " " + std::string(39 - gas.size(), '*')).c_str());
    triple(kv.second, true);
    std::fprintf(f, "               Ionization collisions = %#+.14e (eVm3/s)\n", at(kv.second, "ionizationIne"));
    std::fprintf(f, "               Attachment collisions = %#+.14e (eVm3/s)\n", at(kv.second, "attachmentIne"));
  }
}

namespace {
void rateLine(FILE* f, const RateCoeff& rc, bool mc) {
  const double ine = mc ? rc.ineRateMC : rc.ineRate, sup = mc ? rc.supRateMC : rc.supRate;
  if (rc.supRate == NON_DEF) std::fprintf(f, "%4d %20.14e (N/A)                %s\n", rc.collID + 1, ine, rc.description.c_str());
  else std::fprintf(f, "%4d %20.14e %20.14e %s\n", rc.collID + 1, ine, sup, rc.description.c_str());
}
// the commented header of the rate-coefficient tables (Output.h:486-513, :612-637); returns the column header
std::string rateTableHeader(FILE* f, const Report& r, std::string header) {
  std::fprintf(f, "%s\n# %-76s #\n", std::string(80, '#').c_str(), "ID   Description");
  auto add = [&](const std::vector<RateCoeff>& v) {
    for (const auto& rc : v) {
      const int id = rc.collID + 1;
      std::fprintf(f, "# %-4d %-71s #\n", id, rc.description.c_str());
      std::string s = "R" + std::to_string(id) + "_ine(m3/s)";
      header += s + std::string(22 - s.size(), ' ');
      if (rc.supRate != NON_DEF) { s = "R" + std::to_string(id) + "_sup(m3/s)"; header += s + std::string(22 - s.size(), ' '); }
    }
  };
  add(r.rateAll);
  std::fprintf(f, "#%s#\n# %-76s #\n#%s#\n# %-76s #\n", std::string(78, ' ').c_str(), "*** Extra rate coefficients ***", std::string(78, '#').c_str(), "ID   Description");
  add(r.rateExtra);
  std::fprintf(f, "%s\n\n%s\n", std::string(80, '#').c_str(), header.c_str());
  return header;
}
void rateRow(FILE* f, const std::vector<RateCoeff>& all, const std::vector<RateCoeff>& extra) {
  for (const auto* v : {&all, &extra}) for (const auto& rc : *v) {
    std::fprintf(f, "%-21.14e ", rc.ineRate);
    if (rc.supRate != NON_DEF) std::fprintf(f, "%-21.14e ", rc.supRate);
  }
  std::fprintf(f, "\n");
}
}  // namespace

void OutputWriter::saveRateCoefficients(const Report& r, const std::string& dir) {   // Output.h:419-538
  {
    File f(dir + "/rateCoefficients.txt", "w");
    std::fprintf(f, " ID  Ine.R.Coeff.(m3/s)   Sup.R.Coeff.(m3/s)   Description\n");
    for (const auto& rc : r.rateAll) rateLine(f, rc, false);
    if (!r.rateExtra.empty()) {
      std::fprintf(f, "\n%s\n* Extra Rate Coefficients *\n%s\n\n", std::string(27, '*').c_str(), std::string(27, '*').c_str());
      for (const auto& rc : r.rateExtra) rateLine(f, rc, false);
    }
  }
  {
    File f(dir + "/rateCoefficientsMC.txt", "w");
    std::fprintf(f, " ID  Ine.R.Coeff.(m3/s)   Sup.R.Coeff.(m3/s)   Description\n");
    for (const auto& rc : r.rateAll) if (rc.description.find("Effective") == std::string::npos) rateLine(f, rc, true);
  }
  const double w = r.excitationFrequency * 2.0 * PI;
  if (w != 0) {
    File f(dir + "/rateCoefficients_periodic.txt", "w");
    std::string header;
    for (const char* v : {"Phase(rad)", "Phase(s)", "E/N(Td)"}) header += v + std::string(22 - std::strlen(v), ' ');
    rateTableHeader(f, r, header);
    for (int p = 0; p < r.nPh; ++p) {
      const double ph = r.integrationPhases[p];
      std::fprintf(f, "%-21.14e %-21.14e %-21.14e ", ph, ph / w, std::sqrt(2) * r.reducedElecField * std::cos(ph));
      rateRow(f, r.rateAllPeriodic[p], r.rateExtraPeriodic[p]);
    }
  }
}

void OutputWriter::saveLookUpTable(const Report& r) {   // Output.h:540-727
  const std::string n1 = folder + "/lookUpTableSwarm.txt", n2 = folder + "/lookUpTablePower.txt", n3 = folder + "/lookUpTableRateCoeff.txt",
                    n4 = folder + "/lookUpTableEedf.txt", n5 = folder + "/lookUpTableEGrid.txt";
  const std::string& vc = in_.wc.variableCondition;
  if (!lookUpInitialized_) {
    const std::string varCond = vc == "reducedElecField" ? "RedElecField(Td)      " : vc == "reducedMagField" ? "RedMagField(Hx)       " :
                                vc == "elecFieldAngle" ? "elecFieldAngle(degr)  " : vc == "excitationFrequency" ? "ExcitationFreq(Hz)    " : "";
    File f1(n1, "w"), f2(n2, "w"), f3(n3, "w"), f4(n4, "w"), f5(n5, "w");
    static const char* swarmVars[] = {"FluxND_xx(1/(ms))", "FluxND_xy(1/(ms))", "FluxND_xz(1/(ms))", "FluxND_yx(1/(ms))", "FluxND_yy(1/(ms))", "FluxND_yz(1/(ms))",
        "FluxND_zx(1/(ms))", "FluxND_zy(1/(ms))", "FluxND_zz(1/(ms))", "FluxV_x(m/s)", "FluxV_y(m/s)", "FluxV_z(m/s)", "BulkND_xx(1/(ms))", "BulkND_xy(1/(ms))",
        "BulkND_xz(1/(ms))", "BulkND_yx(1/(ms))", "BulkND_yy(1/(ms))", "BulkND_yz(1/(ms))", "BulkND_zx(1/(ms))", "BulkND_zy(1/(ms))", "BulkND_zz(1/(ms))", "BulkV_x(m/s)",
        "BulkV_y(m/s)", "BulkV_z(m/s)", "SSTVelocity(m/s)", "SSTRedTow(m2)", "SSTRedAtt(m2)", "MeanE(eV)", "EleTemp(eV)", "IonCoeff(m3/s)", "AttCoeff(m3/s)",
        "MomTransfFreq(1/s)", "EnergyRelFreq(1/s)", "RedDiffE_eedf(eV/(ms))", "RedMobE_eedf(eV/(msV))", "RedDiff_eedf(1/(ms))", "RedMob_DC_eedf(1/(msV))",
        "CharacEnergy_eedf(eV)", "trialCollFreq(1/s)"};
    std::string h = varCond;
    for (const char* v : swarmVars) h += v + std::string(24 - std::strlen(v), ' ');
    std::fprintf(f1, "%s\n", h.c_str());
    static const char* powerVars[] = {"PowerField(eVm3/s)", "PwrElaGain(eVm3/s)", "PwrElaLoss(eVm3/s)", "PwrElaNet(eVm3/s)", "PwrEleGain(eVm3/s)", "PwrEleLoss(eVm3/s)",
        "PwrEleNet(eVm3/s)", "PwrVibGain(eVm3/s)", "PwrVibLoss(eVm3/s)", "PwrVibNet(eVm3/s)", "PwrRotGain(eVm3/s)", "PwrRotLoss(eVm3/s)", "PwrRotNet(eVm3/s)",
        "PwrIon(eVm3/s)", "PwrAtt(eVm3/s)", "PwrGrowth(eVm3/s)", "PwrBalance(eVm3/s)", "RelPwrBalance"};
    h = varCond;
    for (const char* v : powerVars) h += v + std::string(22 - std::strlen(v), ' ');
    std::fprintf(f2, "%s\n", h.c_str());
    rateTableHeader(f3, r, varCond);
    lookUpInitialized_ = true;
  }
  File f1(n1, "a"), f2(n2, "a"), f3(n3, "a"), f4(n4, "a"), f5(n5, "a");
  const double cond = in_.jobValue(r.job);
  const double N = r.cfg.gas_density;
  const auto& s = r.swarm; const auto& p = r.power;
  std::fprintf(f1, "%-21.14e ", cond);
  for (int i = 0; i < 9; ++i) std::fprintf(f1, "%-23.14e ", r.d.res.flux_diffusion[i] * N);
  for (int i = 0; i < 3; ++i) std::fprintf(f1, "%-23.14e ", r.d.res.flux_drift_velocity[i]);
  for (int i = 0; i < 9; ++i) std::fprintf(f1, "%-23.14e ", r.d.res.bulk_diffusion[i] * N);
  for (int i = 0; i < 3; ++i) std::fprintf(f1, "%-23.14e ", r.d.res.bulk_drift_velocity[i]);
  for (const char* k : {"effSSTAverageVelocity", "effSSTRedTownsendCoeff", "effSSTRedAttCoeff", "meanEnergy", "Te", "totalIonRateCoeff", "totalAttRateCoeff", "momTransfFreq",
                        "energyRelaxFreq", "redDiffCoeffEnergy_eedf", "redMobCoeffEnergy_eedf", "redDiffCoeff_eedf", "redMobCoeff_DC_eedf", "characEnergy_eedf"})
    std::fprintf(f1, "%-23.14e ", at(s, k));
  std::fprintf(f1, "%-23.14e \n", r.d.res.trial_collision_frequency);
  std::fprintf(f2, "%-21.14e ", cond);
  for (const char* k : {"field", "elasticGain", "elasticLoss", "elasticNet", "excitationSup", "excitationIne", "excitationNet", "vibrationalSup", "vibrationalIne",
                        "vibrationalNet", "rotationalSup", "rotationalIne", "rotationalNet", "ionizationIne", "attachmentIne", "eDensGrowth", "balance"})
    std::fprintf(f2, "%-21.14e ", at(p, k));
  std::fprintf(f2, "%19.14e%%\n", at(p, "relativeBalance") * 100.0);
  std::fprintf(f3, "%-21.14e ", cond);
  rateRow(f3, r.rateAll, r.rateExtra);
  std::fprintf(f4, "%-20.14e ", cond); std::fprintf(f5, "%-20.14e ", cond);
  for (int i = 0; i < r.nE; ++i) { std::fprintf(f4, "%-20.14e ", r.eedf[i]); std::fprintf(f5, "%-20.14e ", r.energyCell[i]); }
  std::fprintf(f4, "\n"); std::fprintf(f5, "\n");
}

void OutputWriter::saveMCTemporalInfo(const Report& r, const std::string& dir) {   // Output.h:730-752
  File f(dir + "/MCTemporalInfo.txt", "w");
  std::string header;
  for (const char* v : {"Time(s)", "MeanEnergy(eV)", "xPos(m)", "yPos(m)", "zPos(m)", "xSqWidth(m2)", "ySqWidth(m2)", "zSqWidth(m2)", "xVel(m/s)", "yVel(m/s)", "zVel(m/s)"})
    header += v + std::string(20 - std::strlen(v), ' ');
  std::fprintf(f, "%s\n", header.c_str());
  const auto& d = r.d;
  for (int64_t i = 0; i < d.res.n_sampling_points; ++i)
    std::fprintf(f, "%-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e %-19.10e \n", d.samplingTimes[i], d.meanEnergies[i],
                 d.meanPositions[3 * i], d.meanPositions[3 * i + 1], d.meanPositions[3 * i + 2], d.positionCovariances[9 * i], d.positionCovariances[9 * i + 4],
                 d.positionCovariances[9 * i + 8], d.meanVelocities[3 * i], d.meanVelocities[3 * i + 1], d.meanVelocities[3 * i + 2]);
}

void OutputWriter::saveMCTemporalInfoPeriodic(const Report& r, const std::string& dir) {   // Output.h:754-782
  const double w = r.excitationFrequency * 2.0 * PI;
  if (w == 0) return;
  File f(dir + "/MCTemporalInfo_periodic.txt", "w");
  std::string header;
  for (const char* v : {"Phase(rad)", "Phase(s)", "E/N(Td)", "MeanEnergy(eV)", "FluxV_x(m/s)", "FluxV_y(m/s)", "FluxV_z(m/s)", "BulkV_x(m/s)", "BulkV_y(m/s)", "BulkV_z(m/s)",
                        "FluxND_xx(1/(ms))", "FluxND_xy(1/(ms))", "FluxND_xz(1/(ms))", "FluxND_yx(1/(ms))", "FluxND_yy(1/(ms))", "FluxND_yz(1/(ms))", "FluxND_zx(1/(ms))",
                        "FluxND_zy(1/(ms))", "FluxND_zz(1/(ms))", "BulkND_xx(1/(ms))", "BulkND_xy(1/(ms))", "BulkND_xz(1/(ms))", "BulkND_yx(1/(ms))", "BulkND_yy(1/(ms))",
                        "BulkND_yz(1/(ms))", "BulkND_zx(1/(ms))", "BulkND_zy(1/(ms))", "BulkND_zz(1/(ms))"})
    header += v + std::string(20 - std::strlen(v), ' ');
  std::fprintf(f, "%s\n", header.c_str());
  const double N = in_.wc.gasDensity;
  const auto& d = r.d;
  for (int p = 0; p < r.nPh; ++p) {
    const double ph = r.integrationPhases[p];
    std::fprintf(f, "%-19.10e %-19.10e %-19.10e %-19.10e", ph, ph / w, std::sqrt(2) * r.reducedElecField * std::cos(ph), d.meanEnergiesPeriodic[p]);
    for (int a = 0; a < 3; ++a) std::fprintf(f, " %-19.10e", d.fluxVelocitiesPeriodic[3 * p + a]);
    for (int a = 0; a < 3; ++a) std::fprintf(f, " %-19.10e", d.bulkVelocitiesPeriodic[3 * p + a]);
    for (int a = 0; a < 9; ++a) std::fprintf(f, " %-19.10e", N * d.fluxDiffusionPeriodic[9 * p + a]);
    for (int a = 0; a < 9; ++a) std::fprintf(f, " %-19.10e", N * d.bulkDiffusionPeriodic[9 * p + a]);
    std::fprintf(f, "\n");
  }
}

void OutputWriter::saveMCSimDetails(const Report& r, const std::string& dir) {   // Output.h:784-822
  File f(dir + "/MCSimDetails.txt", "w");
  const auto& q = r.d.res;
  const std::string rule(82, '-');
  std::fprintf(f, "                          number of electrons: %e\n", r.d.nElectrons);
  std::fprintf(f, "                      trialCollisionFrequency: %e s-1\n\n", q.trial_collision_frequency);
  std::fprintf(f, "                        final simulation time: %e s\n", q.time);
  std::fprintf(f, "                            steady-state time: %e s\n", q.steady_state_time);
  std::fprintf(f, "                 number of integration points: %d\n\n", static_cast<int>(q.n_integration_points));
  std::fprintf(f, "********************************** Collisions ************************************\n\n");
  std::fprintf(f, "number of real collisions before steady-state: %e\n", q.collisions_at_ss);
  std::fprintf(f, " number of real collisions after steady-state: %e\n", q.total_collisions - q.collisions_at_ss);
  std::fprintf(f, "%s\n", rule.c_str());
  std::fprintf(f, "              total number of real collisions: %e\n\n", q.total_collisions);
  std::fprintf(f, "number of null collisions before steady-state: %e\n", q.null_collisions_at_ss);
  std::fprintf(f, " number of null collisions after steady-state: %e\n", q.null_collisions - q.null_collisions_at_ss);
  std::fprintf(f, "%s\n", rule.c_str());
  std::fprintf(f, "              total number of null collisions: %e\n\n", q.null_collisions);
  std::fprintf(f, "                  fraction of real collisions: %e%%\n", q.total_collisions / (q.total_collisions + q.null_collisions) * 100.0);
  std::fprintf(f, "                  fraction of null collisions: %e%%\n", q.null_collisions / (q.total_collisions + q.null_collisions) * 100.0);
  std::fprintf(f, "%s\n", rule.c_str());
  std::fprintf(f, "                                        total: %e%%\n\n", 100.0);
  std::fprintf(f, "************************** Simulation relative errors ****************************\n\n");
  std::fprintf(f, "                                  Mean energy: %.4e\n", q.averaged_mean_energy_error / q.averaged_mean_energy);
  std::fprintf(f, "                          Flux drift velocity: %.4e %.4e %.4e\n", q.flux_drift_velocity_error[0] / std::fabs(q.flux_drift_velocity[0]),
               q.flux_drift_velocity_error[1] / std::fabs(q.flux_drift_velocity[1]), q.flux_drift_velocity_error[2] / std::fabs(q.flux_drift_velocity[2]));
  std::fprintf(f, "                  Flux diffusion coefficients: %.4e %.4e %.4e\n", q.flux_diffusion_error[0] / std::fabs(q.flux_diffusion[0]),
               q.flux_diffusion_error[4] / std::fabs(q.flux_diffusion[4]), q.flux_diffusion_error[8] / std::fabs(q.flux_diffusion[8]));
  std::fprintf(f, "                                Power balance: %.4e\n", q.power_balance_rel_error);
  std::fprintf(f, "**********************************************************************************\n\n");
  std::fprintf(f, "                                 Elapsed time: %e s\n", q.elapsed_seconds);
}

}  // namespace lokihost
