// oracle/harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference (liblokiref.so, built from /root/reference by oracle/Makefile) with INJECTED
// random draws, to generate golden vectors for the hot path and to dump the reference's flattened process set
// and tables.  The executable defines MathFunctions::unitUniformRand itself; ELF symbol interposition makes the
// -fPIC library (including its unitNormalRand3) use this definition instead of its mt19937_64 one
// (reference: Code/LoKI-MC/Sources/MathFunctions.C:31-43,54-59).
//
// usage (cwd must contain Input/):  harness <setup.in (relative to Input/)> <command file> <output prefix>
// commands, one per line:
//   model                       -> <prefix>.model.txt   (process SoA, raw cross sections, working conditions)
//   tables <maxE>               -> <prefix>.tables.bin  (sigma[nE][P], cum[nE][P], nu_tot[nE], nu_max[nE]); also sets
//                                  trialCollisionFrequency = maxCollisionFrequencies[nE-1] (BoltzmannMC.C:515)
//   event <nu_trial> <t_e> <x y z> <vx vy vz> <t_cf|ND> <nu_e> <t_sync> <n> <d0..dn-1>
//                               -> one pass of the per-electron loop body, BoltzmannMC.C:637-681, on electron 0
//   expr <text> / vexpr <text>  -> Parse::str2value / Parse::evalVectorExpress (Parse.C:610-755)
//   controls                    -> the numericsMC keys as stored by the BoltzmannMC constructor
//   solve <seed>                -> runs every job of the setup (BoltzmannMC::solve, Setup::nextJob) on a deterministic generator; dumps the raw
//                                  state the sinks read (<prefix>.job<k>.raw.bin) and lets the reference's Output write its files
//   maxaccel <eps> <dt>         -> maximizationAccelerationEnergy (BoltzmannMC.C:765-802)
//   moments <file.bin>          -> calculateMeanDataForSwarmParams (BoltzmannMC.C:1410-1482) on N=(nElectrons) electrons
//                                  read from file (x[N] y[N] z[N] vx[N] vy[N] vz[N] doubles)
//   hists <file.bin> <maxElecEnergy> -> forces the steady-state grid set-up (BoltzmannMC.C:1862-1891), i.e. one
//                                  getTimeDependDistributions sample; writes <prefix>.hists.bin (eeh, eah, evh)
// every result is printed to <prefix>.out.txt with %.17g.

#include "LoKI-MC/Headers/Parse.h"
#include "LoKI-MC/Headers/Setup.h"
#include "LoKI-MC/Headers/BoltzmannMC.h"
#include "LoKI-MC/Headers/PrescribedEedf.h"
#include "LoKI-MC/Headers/FieldInfo.h"
#include "LoKI-MC/Headers/Message.h"
#include "LoKI-MC/Headers/MathFunctions.h"
#include "LoKI-MC/Headers/GeneralDefinitions.h"
#include <omp.h>
#include <cstdio>
#include <deque>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

static std::deque<double> g_draws;
static long g_used = 0;
static unsigned long long g_prng = 0;

// interposes the library's definition
double MathFunctions::unitUniformRand(bool includeZero, bool includeOne) {
  (void)includeZero; (void)includeOne;
  ++g_used;
  if (g_draws.empty() && g_prng) {   // 'solve': a deterministic generator (xorshift64*), open interval
    g_prng ^= g_prng >> 12; g_prng ^= g_prng << 25; g_prng ^= g_prng >> 27;
    return ((double)((g_prng * 0x2545F4914F6CDD1DULL) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  if (g_draws.empty()) return 0.5;
  double d = g_draws.front();
  g_draws.pop_front();
  return d;
}

static void w17(FILE* f, const char* name, double v) { std::fprintf(f, "%s %.17g\n", name, v); }

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: harness setup cmdfile outprefix\n"); return 2; }
  std::string setupFile = argv[1], cmdFile = argv[2], prefix = argv[3];
  omp_set_num_threads(1);
  Eigen::setNbThreads(1);
  Parse::setupFile(setupFile);
  Setup<BoltzmannMC> setup(setupFile);
  setup.initializeSimulation();
  BoltzmannMC* ek = setup.electronKinetics;
  ek->evaluateNonConstantVariables();  // consumes 4*N draws (0.5 each)
  const int N = (int)ek->nElectrons, P = ek->nProcesses, nE = ek->interpolCrossSectionSize;

  FILE* out = std::fopen((prefix + ".out.txt").c_str(), "w");
  std::ifstream cmds(cmdFile);
  std::string line;
  while (std::getline(cmds, line)) {
    std::istringstream is(line);
    std::string cmd;
    if (!(is >> cmd) || cmd[0] == '#') continue;

    if (cmd == "model") {
      FILE* f = std::fopen((prefix + ".model.txt").c_str(), "w");
      std::fprintf(f, "nProcesses %d\nnGases %d\nnInterpPoints %d\n", P, ek->nGases, nE);
      w17(f, "totalGasDensity", ek->totalGasDensity);
      w17(f, "gasTemperature", ek->gasTemperature);
      w17(f, "gasEnergy", ek->gasEnergy);
      std::fprintf(f, "gasTemperatureEffect %d\nenergySharingIonizType %d\n", ek->gasTemperatureEffect, ek->energySharingIonizType);
      w17(f, "energySharingFactor", ek->energySharingFactor);
      w17(f, "energyMaxElastic", ek->energyMaxElastic);
      w17(f, "reducedElecField", ek->reducedElecField);
      w17(f, "elecFieldAngle", ek->elecFieldAngle);
      w17(f, "excitationFrequency", ek->excitationFrequency);
      w17(f, "excitationFrequencyRadians", ek->excitationFrequencyRadians);
      w17(f, "reducedMagField", ek->reducedMagField);
      w17(f, "cyclotronFrequency", ek->cyclotronFrequency);
      std::fprintf(f, "electricField %.17g %.17g %.17g\n", ek->electricField[0], ek->electricField[1], ek->electricField[2]);
      std::fprintf(f, "accelerationElecField %.17g %.17g %.17g\n", ek->accelerationElecField[0], ek->accelerationElecField[1], ek->accelerationElecField[2]);
      std::fprintf(f, "isCylindricallySymmetric %d\n", (int)ek->isCylindricallySymmetric);
      w17(f, "initialElecTempOverGasTemp", ek->initialElecTempOverGasTemp);
      for (int g = 0; g < ek->nGases; ++g)
        std::fprintf(f, "gas %d first %d last %d fraction %.17g\n", g, (int)ek->firstProcessIndexPerGas[g], (int)ek->lastProcessIndexPerGas[g], ek->gasFractions[g]);
      for (int k = 0; k < P; ++k) {
        int ang = -1;  // 0 isotropic 1 forward 2 bornDipole 3 surendra 4 coulombScreen 5 momentumConservationIonization
        const std::string& a = ek->realCollisionPointers[k]->angularScatteringType;
        if (a == "isotropic") ang = 0; else if (a == "forward") ang = 1; else if (a == "bornDipole") ang = 2;
        else if (a == "surendra") ang = 3; else if (a == "coulombScreen") ang = 4; else if (a == "momentumConservationIonization") ang = 5;
        double p0 = 0, p1 = 0;
        if (ek->angularScatteringParams[k].size() >= 2) { p0 = ek->angularScatteringParams[k][0]; p1 = ek->angularScatteringParams[k][1]; }
        std::fprintf(f, "process %d type %d elastic %d superelastic %d ionization %d gasid %d angular %d momcons %d", k, ek->processTypes[k],
                     (ek->processTypes[k] == 0 && !ek->isSuperElastic[k] && ek->realCollisionPointers[k]->type == "Elastic") ? 1 : 0,
                     (int)ek->isSuperElastic[k], (int)ek->isIonization[k], ek->targetGasIDs[k], ang, (int)ek->isMomentumConservationIonizationScattering[k]);
        std::fprintf(f, " swf %.17g emin %.17g emax %.17g reldens %.17g mass %.17g redmass %.17g eloss %.17g thstd %.17g w %.17g ap0 %.17g ap1 %.17g",
                     ek->superElasticStatWeightFactors[k], ek->energyMinLimits[k], ek->energyMaxLimits[k], ek->relDensities[k], ek->targetMasses[k],
                     ek->reducedMasses[k], ek->energyLosses[k], ek->thermalStdDeviations[k], ek->wParameters[k], p0, p1);
        int n = (int)ek->crossSectionEnergies[k].size();
        std::fprintf(f, " npts %d\n", n);
        std::fprintf(f, "desc %s\n", ek->realCollisionPointers[k]->description().c_str());
        for (int i = 0; i < n; ++i) std::fprintf(f, "%.17g %.17g\n", ek->crossSectionEnergies[k][i], ek->crossSectionValues[k][i]);
      }
      std::fclose(f);
      std::fprintf(out, "model ok\n");
    }
    else if (cmd == "tables") {
      double maxE; is >> maxE;
      ek->interpolateCrossSections(maxE);
      ek->trialCollisionFrequency = ek->maxCollisionFrequencies[nE - 1];
      ek->trialCollisionFrequenciesEachElectron.fill(ek->trialCollisionFrequency);
      FILE* f = std::fopen((prefix + ".tables.bin").c_str(), "wb");
      double hdr[4] = {(double)nE, (double)P, ek->crossSectionEnergyStep, ek->maxInterpolatedEnergy};
      std::fwrite(hdr, 8, 4, f);
      for (int i = 0; i < nE; ++i) std::fwrite(ek->interpolCrossSectionsXrelDens[i], 8, P, f);
      for (int i = 0; i < nE; ++i) std::fwrite(ek->cumulSumInterpolCrossSectionsXrelDens[i], 8, P, f);
      std::fwrite(ek->totalCollisionFrequencies, 8, nE, f);
      std::fwrite(ek->maxCollisionFrequencies, 8, nE, f);
      std::fclose(f);
      std::fprintf(out, "tables dE %.17g nu_trial %.17g\n", ek->crossSectionEnergyStep, ek->trialCollisionFrequency);
    }
    else if (cmd == "event") {
      // restates the per-electron body of BoltzmannMC.C:637-681 around the reference's own accelerateElectron / performCollision
      double nuTrial, te, tcf, nue, tsync; Eigen::Array3d r, v; std::string tcfs; int n;
      is >> nuTrial >> te >> r[0] >> r[1] >> r[2] >> v[0] >> v[1] >> v[2] >> tcfs >> nue >> tsync >> n;
      tcf = (tcfs == "ND") ? Constant::NON_DEF : std::stod(tcfs);
      g_draws.clear();
      for (int i = 0; i < n; ++i) { double d; is >> d; g_draws.push_back(d); }
      g_used = 0;
      const int id = 0;
      ek->trialCollisionFrequency = nuTrial;
      ek->trialCollisionFrequenciesEachElectron[id] = nue;
      double eps = 0.5 * Constant::electronMass * v.matrix().squaredNorm() / Constant::electronCharge;
      ek->electronEnergyChanges[id] = 0; ek->electronEnergyChangesOverIncidEnergies[id] = 0;
      ek->ejectedElectronEnergies[id] = 0; ek->ejectedElectronVelocities.row(id).setZero(); ek->ejectedElectronPositions.row(id).setZero();
      if (tcf == Constant::NON_DEF) {
        tcf = -std::log(MathFunctions::unitUniformRand(false, false)) / ek->trialCollisionFrequency;
        ek->trialCollisionFrequenciesEachElectron[id] = ek->trialCollisionFrequency;
      }
      if (te + tcf > tsync) {
        double dt = tsync - te;
        ek->accelerateElectron(id, te, dt, r, v, eps);
        te = tsync; tcf -= dt;
        ek->chosenProcessIDs[id] = GeneralDefinitions::partialFreeFlightID;
      } else {
        ek->accelerateElectron(id, te, tcf, r, v, eps);
        te += tcf;
        ek->performCollision(id, r, v, eps);
        tcf = -std::log(MathFunctions::unitUniformRand(false, false)) / ek->trialCollisionFrequency;
        ek->trialCollisionFrequenciesEachElectron[id] = ek->trialCollisionFrequency;
      }
      std::fprintf(out, "event %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %ld\n",
                   ek->chosenProcessIDs[id], r[0], r[1], r[2], v[0], v[1], v[2], eps, te, tcf, ek->trialCollisionFrequenciesEachElectron[id],
                   ek->electronEnergyChanges[id], ek->electronEnergyChangesOverIncidEnergies[id], ek->energyGainsField[id],
                   ek->ejectedElectronPositions(id, 0), ek->ejectedElectronPositions(id, 1), ek->ejectedElectronPositions(id, 2),
                   ek->ejectedElectronVelocities(id, 0), ek->ejectedElectronVelocities(id, 1), ek->ejectedElectronVelocities(id, 2),
                   ek->ejectedElectronEnergies[id], g_used);
    }
    else if (cmd == "expr" || cmd == "vexpr") {   // Parse::str2value / evalVectorExpress on the rest of the line
      std::string rest; std::getline(is, rest);
      rest.erase(0, rest.find_first_not_of(' '));
      if (cmd == "expr") std::fprintf(out, "expr %.17g\n", Parse::str2value(rest));
      else { std::fprintf(out, "vexpr"); for (double v : Parse::evalVectorExpress(rest)) std::fprintf(out, " %.17g", v); std::fprintf(out, "\n"); }
    }
    else if (cmd == "controls") {   // numericsMC keys as the BoltzmannMC ctor stores them (BoltzmannMC.h:262-365)
      std::fprintf(out, "controls %.17g %.17g %.17g %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d %d %d %d %d %.17g\n",
                   ek->requiredIntegrationPoints, ek->requiredIntegratedSSTimes, ek->requiredIntegratedAbsoluteTime, (int)ek->errorsToBeChecked,
                   ek->synchronizationOverSampling, ek->requiredMeanEnergyRelError, ek->requiredFluxDriftVelocityRelError,
                   ek->requiredFluxDiffusionCoeffsRelError, ek->requiredBulkDriftVelocityRelError, ek->requiredBulkDiffusionCoeffsRelError,
                   ek->requiredPowerBalanceRelError, ek->minCollisionsBeforeSteadyState, ek->maxCollisionsBeforeSteadyState,
                   ek->maxCollisionsAfterSteadyState, ek->synchronizationTimeXMaxCollisionFrequency, ek->initialElecTempOverGasTemp,
                   (int)ek->nEnergyCells, (int)ek->nCosAngleCells, (int)ek->nRadialVelocityCells, (int)ek->nAxialVelocityCells,
                   (int)ek->nIntegrationPhases, (int)ek->interpolCrossSectionSize, (double)ek->nElectrons);
    }
    else if (cmd == "solve") {   // full jobs with a deterministic generator; raw BoltzmannMC state -> <prefix>.job<k>.raw.bin, files via Output
      unsigned long long seed; is >> seed;
      for (int job = 0; setup.currentJobID < setup.numberOfJobs; ++job) {
        g_prng = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(job + 1);
        g_draws.clear();
        ek->solve();
        FILE* f = std::fopen((prefix + ".job" + std::to_string(job) + ".raw.bin").c_str(), "wb");
        auto put = [&](double v) { std::fwrite(&v, sizeof(double), 1, f); };
        auto putv = [&](const Eigen::ArrayXd& a) { for (int i = 0; i < a.size(); ++i) put(a[i]); };
        auto putm = [&](const Eigen::ArrayXXd& a, int rows) { for (int i = 0; i < rows; ++i) for (int c = 0; c < a.cols(); ++c) put(a(i, c)); };
        auto putm3 = [&](const Eigen::Matrix3d& a) { for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) put(a(i, c)); };
        const bool ac = ek->excitationFrequencyRadians != 0;
        const int nS = ek->nSamplingPoints, nPh = ac ? ek->nIntegrationPhases : 0;
        put(P); put(ek->nEnergyCells); put(ek->nCosAngleCells); put(ek->nRadialVelocityCells); put(ek->nAxialVelocityCells); put(nPh); put(nS);
        put(ek->isCylindricallySymmetric); put(ek->nElectrons);
        put(ek->averagedMeanEnergy); put(ek->averagedMeanEnergyError);
        putv(ek->averagedFluxDriftVelocity); putv(ek->averagedFluxDriftVelocityError); putm3(ek->averagedFluxDiffusionCoeffs); putm3(ek->averagedFluxDiffusionCoeffsError);
        putv(ek->averagedBulkDriftVelocity); putv(ek->averagedBulkDriftVelocityError); putm3(ek->averagedBulkDiffusionCoeffs); putm3(ek->averagedBulkDiffusionCoeffsError);
        put(ek->averagedPowerGainField); put(ek->averagedPowerGrowth); put(ek->powerBalanceRelError);
        put(ek->time); put(ek->steadyStateTime); put(ek->totalIntegratedTime); put(ek->trialCollisionFrequency); put(ek->maxEedfEnergy); put(ek->elapsedTime);
        put((double)ek->totalCollisionCounter); put((double)ek->nullCollisionCounter); put((double)ek->collisionCounterAtSS); put((double)ek->nullCollisionCounterAtSS);
        put(ek->nSamplingPoints); put(ek->nIntegrationPoints);
        put(ek->radialVelocityNodes[ek->nRadialVelocityCells]);   // maxSpeed of the velocity grid (fixed at steady state)
        for (int k = 0; k < P; ++k) put(ek->averagedRateCoeffs[k]);
        for (int k = 0; k < P; ++k) put(ek->averagedPowerGainProcesses[k]);
        for (int k = 0; k < P; ++k) put(ek->averagedPowerLossProcesses[k]);
        for (int k = 0; k < P; ++k) put((double)ek->collisionCounters[k]);
        putv(ek->eehSum);
        if (ek->isCylindricallySymmetric) { putm(ek->eahSum, ek->nEnergyCells); putm(ek->evhSum, ek->nRadialVelocityCells); }
        if (ac) putm(ek->eehSum_periodic, nPh);
        for (int i = 0; i < nS; ++i) put(ek->samplingTimes[i]);
        for (int i = 0; i < nS; ++i) put(ek->meanEnergies[i]);
        putm(ek->meanPositions, nS); putm(ek->meanVelocities, nS); putm(ek->positionCovariances, nS);
        if (ac) {
          putv(ek->nIntegrationPointsPerPhase); putv(ek->meanEnergies_periodic); putm(ek->fluxVelocities_periodic, nPh); putm(ek->bulkVelocities_periodic, nPh);
          putm(ek->fluxDiffusionCoeffs_periodic, nPh); putm(ek->bulkDiffusionCoeffs_periodic, nPh);
        }
        std::fclose(f);
        std::fprintf(out, "solve job %d meanEnergy %.17g samples %d\n", job, ek->averagedMeanEnergy, nS);
        setup.nextJob();
      }
      setup.finishSimulation();
    }
    else if (cmd == "maxaccel") {
      double e0, dt; is >> e0 >> dt;
      std::fprintf(out, "maxaccel %.17g\n", ek->maximizationAccelerationEnergy(e0, dt));
    }
    else if (cmd == "moments" || cmd == "hists") {
      std::string fn; is >> fn;
      std::vector<double> buf(6 * (size_t)N);
      FILE* f = std::fopen(fn.c_str(), "rb");
      if (!f || std::fread(buf.data(), 8, buf.size(), f) != buf.size()) { std::fprintf(stderr, "cannot read %s\n", fn.c_str()); return 3; }
      std::fclose(f);
      for (int i = 0; i < N; ++i) {
        for (int c = 0; c < 3; ++c) { ek->electronPositions(i, c) = buf[c * (size_t)N + i]; ek->electronVelocities(i, c) = buf[(3 + c) * (size_t)N + i]; }
      }
      ek->electronEnergies = 0.5 * Constant::electronMass * ek->electronVelocities.matrix().rowwise().squaredNorm() / Constant::electronCharge;
      if (cmd == "moments") {
        ek->maxElecEnergy = 0;
        ek->nSamplingPoints = 1; ek->currentSamplingIndex = 0; ek->samplingTimes[0] = 0;
        ek->calculateMeanDataForSwarmParams();
        std::fprintf(out, "moments %.17g %.17g", ek->meanEnergies[0], ek->maxElecEnergy);
        for (int c = 0; c < 3; ++c) std::fprintf(out, " %.17g", ek->meanPositions(0, c));
        for (int c = 0; c < 3; ++c) std::fprintf(out, " %.17g", ek->meanVelocities(0, c));
        for (int c = 0; c < 9; ++c) std::fprintf(out, " %.17g", ek->positionCovariances(0, c));
        for (int c = 0; c < 9; ++c) std::fprintf(out, " %.17g", ek->fluxDiffusionCoeffs(0, c));
        std::fprintf(out, "\n");
      } else {
        double maxElec; is >> maxElec;
        ek->maxElecEnergy = maxElec;
        ek->time = 1e-6; ek->nSamplingPoints = 1; ek->currentSamplingIndex = 0; ek->samplingTimes[0] = 1e-6; ek->meanEnergies[0] = 1;
        ek->maxCollisionsBeforeSteadyState = 0;  // forces the steady-state branch (BoltzmannMC.C:1815)
        ek->checkSteadyState();                  // grid set-up :1862-1889 then one getTimeDependDistributions sample :1891
        FILE* g = std::fopen((prefix + ".hists.bin").c_str(), "wb");
        int ne = (int)ek->nEnergyCells, nc = (int)ek->nCosAngleCells, nr = (int)ek->nRadialVelocityCells, na = (int)ek->nAxialVelocityCells;
        double hdr[8] = {(double)ne, (double)nc, (double)nr, (double)na, ek->maxEedfEnergy, ek->eedfEnergyStep, ek->radialVelocityStep, ek->axialVelocityStep};
        std::fwrite(hdr, 8, 8, g);
        std::fwrite(ek->eehSum.data(), 8, ne, g);
        for (int i = 0; i < ne; ++i) for (int j = 0; j < nc; ++j) { double x = ek->isCylindricallySymmetric ? ek->eahSum(i, j) : 0; std::fwrite(&x, 8, 1, g); }
        for (int i = 0; i < nr; ++i) for (int j = 0; j < na; ++j) { double x = ek->isCylindricallySymmetric ? ek->evhSum(i, j) : 0; std::fwrite(&x, 8, 1, g); }
        std::fclose(g);
        std::fprintf(out, "hists maxEedfEnergy %.17g\n", ek->maxEedfEnergy);
      }
    }
    else {
      std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
      return 2;
    }
  }
  std::fclose(out);
  return 0;
}
