// report.h -- output side of the host: what LoKI-MC does with a finished Monte Carlo job.
//
// Behavioural restatement of BoltzmannMC::getTimeAverageDistributions / getAveragedPeriodicParams (Sources/BoltzmannMC.C:1574-1604,
// 1727-1741), evaluatePower (:1950-2063), evaluateSwarmParameters (:2065-2271), evaluateRateCoeff (:2273-2384),
// Collision::evaluateRateCoeff (Sources/Collision.C:355-396), Grid::updateMaxValue (Sources/Grid.C:46-55) and of the text sinks in
// Headers/Output.h (file names, headers, column layout and printf formats are the reference's).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "setup_input.h"

namespace lokihost {

// Everything the sinks read from a finished BoltzmannMC: exactly what lokib200_job_* return (include/lokib200.h).
struct JobData {
  lokib200_solve_results res{};
  double nElectrons = 0;
  double evdfMaxSpeed = 0;   // upper node of the velocity grids (fixed when the steady state is found, BMC.C:1877-1884)
  std::vector<double> rateCoeffsMC, powerGain, powerLoss, counts;                         // [P]
  std::vector<double> eehSum, eahSum, evhSum, eehSumPeriodic;                             // [nE], [nE][nCos], [nR][nA], [nPh][nE]
  std::vector<double> samplingTimes, meanEnergies, meanPositions, meanVelocities, positionCovariances;   // n, n, 3n, 3n, 9n
  std::vector<double> pointsPerPhase, meanEnergiesPeriodic, fluxVelocitiesPeriodic, bulkVelocitiesPeriodic, fluxDiffusionPeriodic, bulkDiffusionPeriodic;
};

struct RateCoeff {   // GeneralDefinitions::RateCoeffStruct
  int collID = -1;
  double ineRate = LOKIB200_NON_DEF, supRate = LOKIB200_NON_DEF, ineRateMC = LOKIB200_NON_DEF, supRateMC = LOKIB200_NON_DEF;
  std::string description;
};

class Report {
 public:
  Report(const SetupInput& in, int job, JobData data);
  const SetupInput& in;
  int job;
  JobData d;
  lokib200_config cfg;
  double reducedElecField, reducedMagField, elecFieldAngle, excitationFrequency;

  // grids and distributions
  int nE, nCos, nR, nA, nPh;
  double energyStep = 0;
  std::vector<double> energyNode, energyCell, cosCells, radialCells, axialCells, integrationPhases;
  std::vector<double> eedf, efadf, esadf, eadf, evdf, eedfPeriodic;
  // power balance (GeneralDefinitions::PowerStruct)
  std::map<std::string, double> power;
  std::map<std::string, std::map<std::string, double>> powerByGas;
  // rate coefficients
  std::vector<RateCoeff> rateAll, rateExtra;
  std::vector<std::vector<RateCoeff>> rateAllPeriodic, rateExtraPeriodic;
  // swarm parameters
  std::map<std::string, double> swarm;
  double rotFluxV[3], rotFluxVErr[3], rotBulkV[3], rotBulkVErr[3], rotFluxD[9], rotFluxDErr[9], rotBulkD[9], rotBulkDErr[9];

 private:
  struct GridXS { std::vector<double> integral, momTransf; double ine = LOKIB200_NON_DEF, sup = LOKIB200_NON_DEF; };
  std::vector<GridXS> xs_;   // per Collision::id, adjusted to the energy grid (Collision::adjustCrossSection)
  void distributions();
  void adjustCrossSections();
  void evaluateRate(const Collision& c, const std::vector<double>& f);
  void evaluatePower();
  void evaluateRateCoeffs();
  void evaluateSwarm();
};

// Output.h: one object per setup file; write() is Output::electronKineticsSolution for one finished job.
class OutputWriter {
 public:
  OutputWriter(const SetupInput& in, const std::string& outputRoot);   // creates <outputRoot>/<output.folder> and setup.txt
  void write(const Report& r);
  std::string folder;
  bool enabled = false;

 private:
  const SetupInput& in_;
  bool eedf_ = false, evdf_ = false, power_ = false, swarm_ = false, rates_ = false, lookUp_ = false, temporal_ = false, temporalPeriodic_ = false, details_ = false;
  bool lookUpInitialized_ = false;
  int currentJob_ = -1;
  std::string subFolder(const Report& r);
  void saveEedf(const Report& r, const std::string& dir);
  void saveEvdf(const Report& r, const std::string& dir);
  void saveSwarm(const Report& r, const std::string& dir);
  void savePower(const Report& r, const std::string& dir);
  void saveRateCoefficients(const Report& r, const std::string& dir);
  void saveLookUpTable(const Report& r);
  void saveMCTemporalInfo(const Report& r, const std::string& dir);
  void saveMCTemporalInfoPeriodic(const Report& r, const std::string& dir);
  void saveMCSimDetails(const Report& r, const std::string& dir);
};

}  // namespace lokihost
