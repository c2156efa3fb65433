// setup_input.h -- input side of the host: LoKI-MC setup files, LXCat cross sections and Databases -> flattened process set.
//
// Behavioural restatement of the reference's input path (Code/LoKI-MC: Sources/Parse.C, Sources/FieldInfo.C,
// Headers/Setup.h:229-551, Headers/WorkingConditions.h, Headers/{Gas,State}.h, Sources/Eedf{Gas,State}.C,
// Headers/{Gas,State}PropertyFunctions.h, Sources/Collision.C, Headers/AngularDistributionFunctions.h) and of the process
// flattening in BoltzmannMC::allocateEvaluateVariablesFirstTime / evaluateNonConstantVariables (Sources/BoltzmannMC.C:29-271,
// 428-489).  File formats and semantics are the reference's; the code structure is this project's own.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lokib200.h"

namespace lokihost {

struct SetupError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- arithmetic expressions of the setup files (Parse::str2value -> External/MathParser) ----
double evalExpression(const std::string& expr);
std::vector<double> evalVectorExpression(const std::string& expr);   // linspace / logspace / a:b / a:s:b / [a,b,c] (Parse.C:610-672)

// ---- setup tree (FieldInfo) ----
struct SetupNode {
  std::string name, value;
  bool isEnumeration = false;
  int spaces = 0, level = 0, line = 0;
  SetupNode* parent = nullptr;
  std::vector<SetupNode*> children;
};

class SetupTree {
 public:
  SetupTree(const std::string& inputDir, const std::string& text);
  static std::string readFile(const std::string& path);
  const SetupNode* find(const std::string& dottedPath) const;                     // FieldInfo::getField
  bool has(const std::string& p) const { return find(p) != nullptr; }
  std::string value(const std::string& p) const;                                   // "" when absent
  double number(const std::string& p) const;                                       // getFieldNumericValue
  std::vector<std::string> childNames(const std::string& p) const;                 // getFieldChildNames
  std::map<std::string, std::string> map(const std::string& p) const;              // getFieldMap (files expanded)
  std::map<std::string, double> numericMap(const std::string& p) const;            // getFieldNumericMap
  const std::string& inputDir() const { return inputDir_; }
  std::string dump() const;                                                        // FieldInfo::printSetupInfo / saveSetupInfo layout

 private:
  std::string inputDir_;
  std::vector<std::unique_ptr<SetupNode>> nodes_;
};

// ---- working conditions (WorkingConditions.h:56-178) ----
struct WorkingConditions {
  double gasPressure = 0, gasTemperature = 0, gasDensity = 0, electronTemperature = 0;
  std::vector<double> reducedElecFieldArray, reducedMagFieldArray, elecFieldAngleArray, excitationFrequencyArray;
  std::string variableCondition;
  bool isCylindricallySymmetric = true;
  int nJobs() const;
  double get(const std::string& name) const;   // WorkingConditions::getValue for property-function arguments
};

// ---- gas / state / collision ontology ----
struct Gas; struct State; struct Collision;

struct Gas {
  int id = -1;
  std::string name;
  std::map<std::string, double> prop;   // mass, fraction, harmonicFrequency, ... (absent = NON_DEF; fraction defaults to 0)
  std::vector<State*> states;
  std::vector<Collision*> collisions, collisionsExtra;
  std::vector<double> effectivePopulations;
  double get(const std::string& p) const { auto it = prop.find(p); return it == prop.end() ? (p == "fraction" ? 0.0 : LOKIB200_NON_DEF) : it->second; }
};

struct State {
  int id = -1;
  std::string type, ionCharg, eleLevel, vibLevel, rotLevel, name;
  Gas* gas = nullptr;
  State* parent = nullptr;
  std::vector<State*> siblings, children;
  double energy = LOKIB200_NON_DEF, statisticalWeight = LOKIB200_NON_DEF, population = 0, density = 0;
  bool isTarget = false;
  std::vector<Collision*> collisions, collisionsExtra;
};

struct CrossSection { std::vector<double> e, v; bool empty() const { return e.empty(); } };

struct Collision {
  int id = -1;
  std::string type;
  State* target = nullptr;
  std::vector<State*> products;
  std::vector<double> productStoi;
  bool isExtra = false, isReverse = false;
  double threshold = 0;
  CrossSection rawIntegral, rawMomTransf;
  std::string angularType = "isotropic";
  std::vector<double> angularParams;
  std::string description() const;
};

// Collision::interpolatedCrossSection / superElasticCrossSection (Collision.C:232-353): linear interpolation of the raw curve, 0 outside;
// Klein-Rosseland relation for the reverse direction
std::vector<double> interpolatedCrossSection(const Collision& c, bool momTransf, const std::vector<double>& energies);
std::vector<double> superElasticCrossSection(const Collision& c, bool momTransf, const std::vector<double>& energies);

struct ProcessSet {   // what BoltzmannMC holds per process (BMC.h:86-118), ready for lokib200_set_processes
  std::vector<int32_t> type, isSuperelastic, isElastic, angularModel, gasFirst, gasLast;
  std::vector<double> ap0, ap1, swf, emin, emax, relDensity, targetMass, reducedMass, energyLoss, thermalStd, wParameter, gasFraction;
  std::vector<int64_t> xsOffset;
  std::vector<double> xsEnergy, xsValue;
  std::vector<std::string> descriptions;
  std::vector<const Collision*> collisionOf;
  double energyMaxElastic = 1e100;
  lokib200_process_soa soa() const;
};

class Mixture {
 public:
  Mixture(const SetupTree& tree, const WorkingConditions& wc);
  std::vector<std::unique_ptr<Gas>> gases;
  std::vector<std::unique_ptr<State>> states;
  std::vector<std::unique_ptr<Collision>> collisions;   // index == Collision::id
  std::vector<std::string> warnings;
  ProcessSet flatten(double gasTemperature) const;      // BMC.C:29-271 + relDensities/thermalStd of :438-455

 private:
  Gas* addGas(const std::string& name);
  State* addState(Gas* g, const std::string& ion, const std::string& ele, const std::string& vib, const std::string& rot);
  std::vector<State*> findStates(const std::string& gas, const std::string& ion, const std::string& ele, const std::string& vib, const std::string& rot) const;
  Collision* addCollision(const std::string& type, State* target, const std::vector<State*>& products, const std::vector<double>& stoi, bool isReverse,
                          double threshold, const CrossSection& integral, const CrossSection& momTransf, bool isExtra);
  void loadLXCat(const SetupTree& tree, const std::string& key, bool isExtra);
  void gasProperties(const SetupTree& tree, const WorkingConditions& wc);
  void stateProperties(const SetupTree& tree, const WorkingConditions& wc);
  void assignAngularScattering(const SetupTree& tree);
  void checkPopulationNorms(const Gas* g) const;
  void checkElasticCollisions(Gas* g);
  CrossSection elasticFromEffective(Gas* g);
};

// one setup file -> everything the engine needs, per job
class SetupInput {
 public:
  SetupInput(const std::string& inputDir, const std::string& setupFile);
  std::unique_ptr<SetupTree> tree;
  WorkingConditions wc;
  std::unique_ptr<Mixture> mixture;
  ProcessSet processes;
  int nJobs() const { return wc.nJobs(); }
  lokib200_config config(int job) const;            // BMC ctor keys + evaluateNonConstantVariables fields (BMC.C:461-489)
  lokib200_solve_controls controls() const;         // numericsMC keys (BMC.h:262-365)
  double jobValue(int job) const;                   // value of the swept condition for this job
};

}  // namespace lokihost
