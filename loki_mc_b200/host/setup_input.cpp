// setup_input.cpp -- see setup_input.h.  Citations: "Parse.C", "FieldInfo.C", "Setup.h", "Collision.C", "EedfGas.C", "EedfState.C",
// "State.h", "StatePropertyFunctions.h" (SPF.h), "GasPropertyFunctions.h" (GPF.h), "AngularDistributionFunctions.h" (ADF.h),
// "WorkingConditions.h" (WC.h), "BoltzmannMC.C" (BMC.C) under /root/reference/Code/LoKI-MC/{Sources,Headers}.
#include "setup_input.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

namespace lokihost {

namespace {

constexpr double NON_DEF = LOKIB200_NON_DEF;
constexpr double KB = 1.38064852e-23, QE = 1.6021766208e-19, ME = 9.10938356e-31, PLANCK = 6.626070040e-34, AVOGADRO = 6.02214076e23;
constexpr double PI = 3.14159265358979323846;
const double KB_EV = KB / QE;                               // Constant::boltzmannInEV
const double HBAR_EV = PLANCK / (2.0 * PI * QE);            // Constant::planckReducedInEV

std::string stripComment(const std::string& s) { const size_t i = s.find('%'); return i == std::string::npos ? s : s.substr(0, i); }

std::vector<std::string> splitAny(const std::string& s, const char* delims) {   // strtok semantics: runs of delimiters, no empty tokens
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && std::strchr(delims, s[i])) ++i;
    size_t j = i;
    while (j < s.size() && !std::strchr(delims, s[j])) ++j;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}
std::vector<std::string> splitSpaces(const std::string& s) { return splitAny(s, " \r\n\v\f\t"); }
std::string removeSpaces(const std::string& s) { std::string o; for (auto& t : splitSpaces(s)) o += t; return o; }
std::string upper(std::string s) { for (auto& c : s) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c))); return s; }
bool iequals(const std::string& a, const std::string& b) { return upper(a) == upper(b); }

// ------------------------------------------------------------------ expression evaluator ------------------------------------------------------------------
// Grammar and precedence of External/MathParser/parser.cpp (levels 4-10); the result goes through "%.16g" and back exactly like
// Parse::str2value does (parser.cpp:96, Parse.C:738).
class Expr {
 public:
  explicit Expr(const std::string& s) : s_(s) {}
  double run() {
    next();
    const double v = addsub();
    if (!tok_.empty()) throw SetupError("Could not parse the mathematical expression: " + s_);
    return v;
  }

 private:
  std::string s_, tok_;
  size_t p_ = 0;
  enum Kind { END, DELIM, NUMBER, NAME } kind_ = END;
  static bool isDelim(char c) { return std::strchr("&|<>=+/*%^!", c) != nullptr; }
  void next() {
    tok_.clear();
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t')) ++p_;
    if (p_ >= s_.size()) { kind_ = END; return; }
    const char c = s_[p_];
    if (c == '-' || c == '(' || c == ')') { kind_ = DELIM; tok_ = std::string(1, c); ++p_; return; }
    if (isDelim(c)) { kind_ = DELIM; while (p_ < s_.size() && isDelim(s_[p_])) tok_ += s_[p_++]; return; }
    if (std::isdigit(static_cast<unsigned char>(c)) || c == '.') {
      kind_ = NUMBER;
      while (p_ < s_.size() && (std::isdigit(static_cast<unsigned char>(s_[p_])) || s_[p_] == '.')) tok_ += s_[p_++];
      if (p_ < s_.size() && std::toupper(static_cast<unsigned char>(s_[p_])) == 'E') {
        tok_ += s_[p_++];
        if (p_ < s_.size() && (s_[p_] == '+' || s_[p_] == '-')) tok_ += s_[p_++];
        while (p_ < s_.size() && std::isdigit(static_cast<unsigned char>(s_[p_]))) tok_ += s_[p_++];
      }
      return;
    }
    if (std::isalpha(static_cast<unsigned char>(c)) || c == '_') {
      kind_ = NAME;
      while (p_ < s_.size() && (std::isalnum(static_cast<unsigned char>(s_[p_])) || s_[p_] == '_')) tok_ += s_[p_++];
      return;
    }
    throw SetupError("Could not parse the mathematical expression: " + s_);
  }
  double addsub() {
    double v = muldiv();
    while (kind_ == DELIM && (tok_ == "+" || tok_ == "-")) { const std::string op = tok_; next(); const double r = muldiv(); v = (op == "+") ? v + r : v - r; }
    return v;
  }
  double muldiv() {
    double v = power();
    while (kind_ == DELIM && (tok_ == "*" || tok_ == "/" || tok_ == "%")) {
      const std::string op = tok_; next(); const double r = power();
      if (op == "*") v = v * r; else if (op == "/") v = v / r; else v = static_cast<double>(static_cast<int>(v) % static_cast<int>(r));
    }
    return v;
  }
  double power() {
    double v = factorial();
    while (kind_ == DELIM && tok_ == "^") { next(); v = std::pow(v, factorial()); }
    return v;
  }
  double factorial() {
    double v = unary();
    while (kind_ == DELIM && tok_ == "!") { next(); double f = 1; for (int i = 2; i <= static_cast<int>(v); ++i) f *= i; v = f; }
    return v;
  }
  double unary() {
    if (kind_ == DELIM && tok_ == "-") { next(); return -function(); }
    return function();
  }
  double function() {
    if (kind_ == NAME) {
      const std::string name = upper(tok_);
      size_t q = p_;
      while (q < s_.size() && (s_[q] == ' ' || s_[q] == '\t')) ++q;
      if (q < s_.size() && s_[q] == '(') {
        next();
        const double a = primary();
        if (name == "ABS") return std::fabs(a);
        if (name == "EXP") return std::exp(a);
        if (name == "SIGN") return (a > 0) - (a < 0);
        if (name == "SQRT") return std::sqrt(a);
        if (name == "LOG") return std::log(a);
        if (name == "LOG10") return std::log10(a);
        if (name == "SIN") return std::sin(a);
        if (name == "COS") return std::cos(a);
        if (name == "TAN") return std::tan(a);
        if (name == "ASIN") return std::asin(a);
        if (name == "ACOS") return std::acos(a);
        if (name == "ATAN") return std::atan(a);
        if (name == "FACTORIAL") { double f = 1; for (int i = 2; i <= static_cast<int>(a); ++i) f *= i; return f; }
        throw SetupError("Could not parse the mathematical expression: " + s_);
      }
    }
    return primary();
  }
  double primary() {
    if (kind_ == DELIM && tok_ == "(") {
      next();
      const double v = addsub();
      if (!(kind_ == DELIM && tok_ == ")")) throw SetupError("Could not parse the mathematical expression: " + s_);
      next();
      return v;
    }
    if (kind_ == NUMBER) { const double v = std::strtod(tok_.c_str(), nullptr); next(); return v; }
    if (kind_ == NAME) {
      const std::string name = upper(tok_);
      next();
      if (name == "E") return 2.7182818284590452353602874713527;
      if (name == "PI" || name == "M_PI") return 3.1415926535897932384626433832795;
    }
    throw SetupError("Could not parse the mathematical expression: " + s_);
  }
};

}  // namespace

double evalExpression(const std::string& expr) {
  // Parse::mathExpIsValid (Parse.C:757-795): balanced parentheses, only known constants / functions
  if (std::count(expr.begin(), expr.end(), '(') != std::count(expr.begin(), expr.end(), ')')) throw SetupError("Could not parse the mathematical expression: " + expr);
  static const char* known[] = {"E", "PI", "M_PI", "ABS", "EXP", "SIGN", "SQRT", "LOG", "LOG10", "SIN", "COS", "TAN", "ASIN", "ACOS", "ATAN", "FACTORIAL"};
  for (const auto& part : splitAny(expr, "()*+-/ .^0123456789%")) {
    bool ok = false;
    for (const char* k : known) ok = ok || iequals(part, k);
    if (!ok) throw SetupError("Could not parse the mathematical expression: " + expr + ". Please check the input files.");
  }
  const double v = Expr(expr).run();
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.16g", v);              // parser.cpp:96 -> sscanf in Parse.C:738
  const double r = std::strtod(buf, nullptr);
  if (std::isnan(r)) throw SetupError("Could not parse the mathematical expression: " + expr + ". Please check the input files.");
  return r;
}

static std::vector<double> linSpace(double start, double end, double num) {   // Parse.C:683-701
  std::vector<double> v;
  if (num == 0) return v;
  if (num == 1) { v.push_back(start); return v; }
  const double delta = (end - start) / (num - 1);
  for (double i = 0; i < num - 1; ++i) v.push_back(start + delta * i);
  v.push_back(end);
  return v;
}

std::vector<double> evalVectorExpression(const std::string& e) {   // Parse.C:610-672
  auto params = [&](const char* what) {
    const size_t a = e.find('('), b = e.find(')');
    const auto toks = splitAny(e.substr(a + 1, b - a - 1), ",");
    if (toks.size() != 3) throw SetupError(std::string("Invalid use of '") + what + "' in the following expression: " + e);
    return std::vector<double>{evalExpression(toks[0]), evalExpression(toks[1]), evalExpression(toks[2])};
  };
  if (e.find("linspace") != std::string::npos) { const auto p = params("linspace"); return linSpace(p[0], p[1], p[2]); }
  if (e.find("logspace") != std::string::npos) { const auto p = params("logspace"); auto v = linSpace(p[0], p[1], p[2]); for (auto& x : v) x = std::pow(10, x); return v; }
  const auto colons = std::count(e.begin(), e.end(), ':');
  std::vector<double> v;
  if (colons == 1) {
    const auto t = splitAny(e, ":");
    const int ini = static_cast<int>(evalExpression(t[0])), fin = static_cast<int>(evalExpression(t[1]));
    if (ini > fin) throw SetupError("Error while using 'evalVectorExpress': " + e);
    for (int i = ini; i <= fin; ++i) v.push_back(i);
    return v;
  }
  if (colons == 2) {
    const auto t = splitAny(e, ":");
    const double ini = evalExpression(t[0]), step = evalExpression(t[1]), fin = evalExpression(t[2]);
    if (ini > fin || step <= 0) throw SetupError("Error while using 'evalVectorExpress': " + e);
    for (double i = ini; i <= fin; i += step) v.push_back(i);
    return v;
  }
  if (e.find('[') != std::string::npos && e.find(']') != std::string::npos) { for (const auto& t : splitAny(e, "[],")) v.push_back(evalExpression(t)); return v; }
  return {evalExpression(e)};
}

// ------------------------------------------------------------------ setup tree ------------------------------------------------------------------
std::string SetupTree::readFile(const std::string& path) {
  std::ifstream f(path);
  if (!f.is_open()) throw SetupError("The file '" + path + "' could not be opened.");
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}

SetupTree::SetupTree(const std::string& inputDir, const std::string& text) : inputDir_(inputDir) {
  std::istringstream in(text);
  std::string raw;
  int lineNumber = 0;
  while (std::getline(in, raw)) {
    const std::string line = stripComment(raw);
    const auto tokens = splitSpaces(line);
    if (tokens.empty()) continue;                                   // Parse.C:27-29
    auto node = std::make_unique<SetupNode>();
    node->line = lineNumber++;
    auto bad = [&]() { return SetupError("Could not parse line " + std::to_string(node->line) + " of the setup file:\n" + line); };
    if (tokens[0] == "-") {                                         // FieldInfo.C:33-50
      node->isEnumeration = true;
      if (tokens.size() == 2) node->name = tokens[1];
      else if (tokens.size() == 4) { node->name = tokens[1]; node->value = tokens[3]; }
      else throw bad();
    } else {                                                        // FieldInfo.C:51-71
      const size_t colon = tokens[0].find(':');
      if (colon == std::string::npos || tokens.size() > 2) throw bad();
      node->name = tokens[0].substr(0, colon);
      if (tokens.size() == 2) node->value = tokens[1];
    }
    for (char c : line) { if (std::isspace(static_cast<unsigned char>(c))) ++node->spaces; else break; }
    for (int k = static_cast<int>(nodes_.size()) - 1; k >= 0; --k) { // FieldInfo.C:88-105
      SetupNode* cand = nodes_[k].get();
      if (node->spaces > cand->spaces && cand->value.empty() && !cand->isEnumeration) { node->parent = cand; node->level = cand->level + 1; cand->children.push_back(node.get()); break; }
    }
    nodes_.push_back(std::move(node));
  }
}

const SetupNode* SetupTree::find(const std::string& path) const {   // FieldInfo.C:333-366
  const auto names = splitAny(path, ".");
  const SetupNode* cur = nullptr;
  for (const auto& n : nodes_) if (n->level == 0 && n->name == names[0]) cur = n.get();   // last level-0 match wins
  for (size_t i = 1; i < names.size() && cur; ++i) {
    const SetupNode* nxt = nullptr;
    for (const SetupNode* c : cur->children) if (c->name == names[i]) { nxt = c; break; }
    cur = nxt;
  }
  return cur;
}
std::string SetupTree::value(const std::string& p) const { const SetupNode* n = find(p); return n ? n->value : std::string(); }
double SetupTree::number(const std::string& p) const { const SetupNode* n = find(p); return n ? evalExpression(n->value) : 0.0; }
std::vector<std::string> SetupTree::childNames(const std::string& p) const {
  std::vector<std::string> out;
  const SetupNode* n = find(p);
  if (!n) return out;
  if (n->value.empty()) for (const SetupNode* c : n->children) out.push_back(c->name); else out.push_back(n->value);
  return out;
}
static void mergePropertyFile(const std::string& path, std::map<std::string, std::string>& m) {   // Parse::modifyPropertyMap, Parse.C:493-520
  std::istringstream in(SetupTree::readFile(path));
  std::string line;
  while (std::getline(in, line)) {
    const auto t = splitSpaces(stripComment(line));
    if (t.empty()) continue;
    if (t.size() != 2) throw SetupError("Error in the parsing of the following property file: " + path + ". Check the following line: \n" + line);
    m[t[0]] = t[1];
  }
}
std::map<std::string, std::string> SetupTree::map(const std::string& p) const {   // FieldInfo.C:218-248
  std::map<std::string, std::string> m;
  const SetupNode* n = find(p);
  if (!n) return m;
  if (n->value.empty()) {
    for (const SetupNode* c : n->children) { if (c->value.empty()) mergePropertyFile(inputDir_ + "/" + c->name, m); else m[c->name] = c->value; }
  } else mergePropertyFile(inputDir_ + "/" + n->value, m);
  return m;
}
std::map<std::string, double> SetupTree::numericMap(const std::string& p) const {
  std::map<std::string, double> m;
  for (const auto& kv : map(p)) m[kv.first] = evalExpression(kv.second);
  return m;
}
std::string SetupTree::dump() const {   // FieldInfo.C:108-129
  std::string out;
  for (const auto& n : nodes_) {
    out += std::string(2 * static_cast<size_t>(n->level), ' ');
    if (n->isEnumeration) out += n->value.empty() ? "- " + n->name : "- " + n->name + " = " + n->value;
    else out += n->name + ": " + n->value;
    out += "\n";
  }
  return out;
}

// ------------------------------------------------------------------ working conditions ------------------------------------------------------------------
int WorkingConditions::nJobs() const {
  if (variableCondition == "reducedElecField") return static_cast<int>(reducedElecFieldArray.size());
  if (variableCondition == "reducedMagField") return static_cast<int>(reducedMagFieldArray.size());
  if (variableCondition == "elecFieldAngle") return static_cast<int>(elecFieldAngleArray.size());
  if (variableCondition == "excitationFrequency") return static_cast<int>(excitationFrequencyArray.size());
  return 1;
}
double WorkingConditions::get(const std::string& n) const {
  if (n == "gasPressure") return gasPressure;
  if (n == "gasTemperature") return gasTemperature;
  if (n == "gasDensity") return gasDensity;
  if (n == "electronTemperature") return electronTemperature;
  if (n == "reducedElecField") return reducedElecFieldArray.empty() ? 0 : reducedElecFieldArray[0];
  if (n == "reducedMagField") return reducedMagFieldArray.empty() ? 0 : reducedMagFieldArray[0];
  if (n == "elecFieldAngle") return elecFieldAngleArray.empty() ? 0 : elecFieldAngleArray[0];
  if (n == "excitationFrequency") return excitationFrequencyArray.empty() ? 0 : excitationFrequencyArray[0];
  throw SetupError("working condition '" + n + "' cannot be used as a property-function argument");
}

static WorkingConditions makeWorkingConditions(const SetupTree& t) {   // WC.h:56-142
  WorkingConditions w;
  const auto m = t.map("workingConditions");
  auto need = [&](const char* k) { auto it = m.find(k); if (it == m.end()) throw SetupError(std::string("workingConditions.") + k + " is missing in the setup file"); return it->second; };
  w.gasPressure = evalExpression(need("gasPressure"));
  w.gasTemperature = evalExpression(need("gasTemperature"));
  w.gasDensity = w.gasPressure / (KB * w.gasTemperature);
  if (m.count("electronTemperature")) w.electronTemperature = evalVectorExpression(m.at("electronTemperature"))[0];
  if (m.count("reducedElecField")) w.reducedElecFieldArray = evalVectorExpression(m.at("reducedElecField"));
  if (m.count("reducedMagField")) w.reducedMagFieldArray = evalVectorExpression(m.at("reducedMagField"));
  if (m.count("elecFieldAngle")) w.elecFieldAngleArray = evalVectorExpression(m.at("elecFieldAngle"));
  if (m.count("excitationFrequency")) w.excitationFrequencyArray = evalVectorExpression(m.at("excitationFrequency"));
  if (w.reducedElecFieldArray.empty() || w.reducedMagFieldArray.empty() || w.elecFieldAngleArray.empty() || w.excitationFrequencyArray.empty())
    throw SetupError("Error in the configuration of the working conditions. When choosing 'boltzmannMC' eedfType, the 'reducedElecField', 'reducedMagField', 'elecFieldAngle' and 'excitationFrequency' must be defined.");
  const int multi = (w.reducedElecFieldArray.size() > 1) + (w.reducedMagFieldArray.size() > 1) + (w.elecFieldAngleArray.size() > 1) + (w.excitationFrequencyArray.size() > 1);
  if (multi > 1) throw SetupError("Error in the configuration of the working conditions. Only one of 'reducedElecField', 'reducedMagField', 'elecFieldAngle' and 'excitationFrequency' may have multiple values.");
  if (w.reducedMagFieldArray.size() > 1) w.variableCondition = "reducedMagField";
  else if (w.elecFieldAngleArray.size() > 1) w.variableCondition = "elecFieldAngle";
  else if (w.excitationFrequencyArray.size() > 1) w.variableCondition = "excitationFrequency";
  else w.variableCondition = "reducedElecField";
  if (w.gasPressure < 0 || w.gasTemperature < 0) throw SetupError("Error in the configuration of the working conditions. The working conditions must be all non-negative.");
  for (double x : w.reducedElecFieldArray) if (x < 0) throw SetupError("Error in the configuration of the working conditions. 'reducedElecField' must be non-negative.");
  for (double x : w.reducedMagFieldArray) if (x < 0) throw SetupError("Error in the configuration of the working conditions. 'reducedMagField' must be non-negative.");
  for (double x : w.excitationFrequencyArray) if (x < 0) throw SetupError("Error in the configuration of the working conditions. 'excitationFrequency' must be non-negative.");
  w.isCylindricallySymmetric = true;                                  // WC.h:164-172
  for (double a : w.elecFieldAngleArray) if (a != 180) w.isCylindricallySymmetric = false;
  return w;
}

// ------------------------------------------------------------------ LXCat ------------------------------------------------------------------
namespace {

struct RawState { std::string gas, ion, ele, vib, rot; };

RawState parseStateName(const std::string& stateName) {   // Parse::getRawState, Parse.C:403-470
  RawState r;
  const auto tok = splitAny(stateName, "()");
  if (tok.size() < 2) throw SetupError("The states must have the electronic state defined between parentheses!\nPlease check '" + stateName + "'");
  r.gas = tok[0];
  const auto f = splitAny(tok[1], ",");
  auto level = [](const std::string& s) { return splitAny(s, "=")[1]; };
  auto isVib = [](const std::string& s) { return s.find("v=") != std::string::npos || s.find("w=") != std::string::npos; };
  switch (f.size()) {
    case 1: r.ele = f[0]; break;
    case 2:
      if (isVib(f[1])) { r.ele = f[0]; r.vib = level(f[1]); } else { r.ion = f[0]; r.ele = f[1]; }
      break;
    case 3:
      if (f[2].find("J=") != std::string::npos) { r.ele = f[0]; if (isVib(f[1])) r.vib = level(f[1]); r.rot = level(f[2]); }
      else { r.ion = f[0]; r.ele = f[1]; r.vib = f[2]; }
      break;
    case 4:
      r.ion = f[0]; r.ele = f[1];
      if (isVib(f[2])) r.vib = level(f[2]);
      r.rot = level(f[3]);
      break;
    default: throw SetupError("Error! Check the state '" + stateName + "'.");
  }
  return r;
}

struct LXCatEntry {
  std::string type;
  bool isReverse = false;
  double threshold = 0;
  std::vector<RawState> target, products;
  std::vector<double> productStoi;
  CrossSection integral, momTransf;
};

double leadingCoefficient(std::string& tok) {   // Parse::getStoiCoeff, Parse.C:376-400
  size_t n = 0;
  while (n < tok.size() && !std::isalpha(static_cast<unsigned char>(tok[n]))) ++n;
  if (n == 0) return 1;
  const double c = evalExpression(tok.substr(0, n));
  tok.erase(0, n);
  return c;
}

// one side of "A + B(...) + ..." -> states (+ electron count); '+' inside parentheses belongs to the state name (Parse.C:192-266)
void parseSide(const std::string& side, std::vector<RawState>& states, std::vector<double>& stoi) {
  std::string cur;
  bool inPar = false;
  for (size_t i = 0; i < side.size(); ++i) {
    const char c = side[i];
    const bool last = (i == side.size() - 1);
    if ((c == '+' && !inPar) || last) {
      inPar = false;
      if (last) cur += c;
      const double k = leadingCoefficient(cur);
      if (!(cur == "e" || cur == "E")) { states.push_back(parseStateName(cur)); stoi.push_back(k); }
      cur.clear();
    } else { if (c == '(') inPar = true; else if (c == ')') inPar = false; cur += c; }
  }
}

LXCatEntry parseEntry(const std::string& descriptionLine, double threshold, const CrossSection& xs) {   // Parse::addLXCatEntry, Parse.C:127-295
  LXCatEntry e;
  const std::string d = removeSpaces(descriptionLine);
  const size_t b = d.find_first_of('['), en = d.find_last_of(']');
  auto parts = splitAny(d.substr(b + 1, en - b - 1), ",");
  std::string last = parts.back(), subType;
  if (iequals(last, "momentum-transfer") || iequals(last, "integral")) { subType = upper(last) == "INTEGRAL" ? "integral" : "momentum-transfer"; parts.pop_back(); last = parts.back(); }
  else subType = (last == "Elastic" || last == "Effective") ? "momentum-transfer" : "integral";
  static const char* types[] = {"Elastic", "Effective", "Attachment", "Ionization", "Excitation", "Vibrational", "Rotational"};
  if (std::none_of(std::begin(types), std::end(types), [&](const char* t) { return last == t; }))
    throw SetupError("Error in the parsing of one LXCat file. Invalid type of collision in the following line: \n" + descriptionLine);
  e.type = last;
  parts.pop_back();
  std::string coll;
  for (const auto& p : parts) coll += "," + p;
  coll.erase(0, 1);
  std::string left, right;
  size_t sep;
  if ((sep = coll.find("<->")) != std::string::npos) { e.isReverse = true; left = coll.substr(0, sep); right = coll.substr(sep + 3); }
  else if ((sep = coll.find("->")) != std::string::npos) { left = coll.substr(0, sep); right = coll.substr(sep + 2); }
  else throw SetupError("Error in the parsing of one LXCat file. Check the direction of the collision presented in the following line: \n" + descriptionLine);
  std::vector<double> targetStoi;
  parseSide(left, e.target, targetStoi);
  if (e.target.empty()) throw SetupError("Could not find a target in the collision presented in the following LXCat line:\n" + descriptionLine);
  std::vector<RawState> prod; std::vector<double> pst;
  parseSide(right, prod, pst);
  for (size_t i = 0; i < prod.size(); ++i) {   // Parse::removeDuplicatedStates
    bool merged = false;
    for (size_t j = 0; j < e.products.size() && !merged; ++j) {
      const RawState& a = prod[i]; const RawState& q = e.products[j];
      if (a.gas == q.gas && a.ion == q.ion && a.ele == q.ele && a.vib == q.vib && a.rot == q.rot) { e.productStoi[j] += pst[i]; merged = true; }
    }
    if (!merged) { e.products.push_back(prod[i]); e.productStoi.push_back(pst[i]); }
  }
  e.threshold = threshold;
  (subType == "momentum-transfer" ? e.momTransf : e.integral) = xs;
  return e;
}

std::vector<LXCatEntry> parseLXCatFiles(const std::string& inputDir, const std::vector<std::string>& files) {   // Parse::LXCatFiles, Parse.C:43-125
  std::vector<LXCatEntry> out;
  for (const auto& fn : files) {
    std::istringstream in(SetupTree::readFile(inputDir + "/" + fn));
    std::string line;
    while (std::getline(in, line)) {
      if (line.empty() || line.find("PROCESS:") == std::string::npos) continue;
      const std::string processLine = line;
      double threshold = 0;
      std::getline(in, line);
      const size_t b = line.find("E ="), e = line.find("eV");
      if (b != std::string::npos && e != std::string::npos) threshold = evalExpression(line.substr(b + 3, e - b - 3));
      else if (processLine.find("Elastic") == std::string::npos && processLine.find("Effective") == std::string::npos)
        throw SetupError("Error! Could not find a threshold for the following process at file '" + fn + "':\n" + processLine);
      std::string description;
      std::getline(in, description);
      while (std::getline(in, line)) if (line.find("-----") != std::string::npos) break;
      CrossSection xs;
      while (std::getline(in, line) && line.find("-----") == std::string::npos) {
        const auto t = splitSpaces(line);
        if (t.empty()) continue;
        if (t.size() != 2) throw SetupError("Error when reading the cross section values of the LXCat collision described in the following line:\n" + description + "Check the LXCat file '" + fn + "'");
        xs.e.push_back(evalExpression(t[0])); xs.v.push_back(evalExpression(t[1]));
      }
      if (xs.e.empty()) throw SetupError("Error when reading the cross section values of the LXCat collision described in the following line:\n" + description);
      out.push_back(parseEntry(description, threshold, xs));
    }
  }
  return out;
}

// GSL linear interpolation semantics (gsl_interp_linear + bsearch)
double linInterp(const std::vector<double>& x, const std::vector<double>& y, double xv) {
  size_t lo = 0, hi = x.size() - 1;
  while (hi > lo + 1) { const size_t mid = (hi + lo) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
  const double dx = x[lo + 1] - x[lo];
  return y[lo] + (xv - x[lo]) / dx * (y[lo + 1] - y[lo]);
}

// Collision::interpolatedCrossSection (Collision.C:232-289) at the points `en`
std::vector<double> interpolatedImpl(const Collision& c, bool momTransf, const std::vector<double>& en) {
  std::vector<double> out(en.size(), 0.0);
  int minIndex = -1;
  if (c.type == "Effective" || c.type == "Elastic") minIndex = 0;
  else { for (size_t i = 0; i < en.size(); ++i) if (en[i] > c.threshold) { minIndex = static_cast<int>(i); break; } if (minIndex == -1) return out; }
  const CrossSection& xs = momTransf ? c.rawMomTransf : c.rawIntegral;
  const double first = xs.e.front(), lastE = xs.e.back();
  for (size_t i = static_cast<size_t>(minIndex); i < en.size(); ++i) out[i] = (en[i] >= first && en[i] <= lastE) ? linInterp(xs.e, xs.v, en[i]) : 0.0;
  return out;
}

// Collision::superElasticCrossSection (Collision.C:291-353), Klein-Rosseland
std::vector<double> superElasticImpl(const Collision& c, bool momTransf, const std::vector<double>& en) {
  if (c.target->statisticalWeight == NON_DEF) throw SetupError("The statistical weight of the state '" + c.target->name + "' is not defined.\nThe super elastic cross section of the collision '" + c.description() + " cannot be evaluated.\n");
  if (c.products[0]->statisticalWeight == NON_DEF) throw SetupError("The statistical weight of the state '" + c.products[0]->name + "' is not defined.\nThe super elastic cross section of the collision '" + c.description() + " cannot be evaluated.\n");
  std::vector<double> out(en.size(), 0.0);
  size_t minIndex = 0;
  if (en[0] == 0) { if (en.size() == 1) return out; minIndex = 1; }
  std::vector<double> shifted(en.size());
  for (size_t i = 0; i < en.size(); ++i) shifted[i] = en[i] + c.threshold;
  auto interp = interpolatedImpl(c, momTransf, shifted);
  const double ratio = c.target->statisticalWeight / c.products[0]->statisticalWeight;
  for (auto& v : interp) v = v * ratio;
  for (size_t i = minIndex; i < en.size(); ++i) out[i] = (1.0 + c.threshold / en[i]) * interp[i];
  return out;
}

// I(eps, theta) of ADF.h:32-112
double angularDistribution(const Collision& c, double energy, double angle) {
  const std::string& t = c.angularType;
  const double iso = 1.0 / (4.0 * PI);
  if (t == "isotropic" || t == "momentumConservationIonization") return iso;
  if (t == "forward") return angle == 0 ? 1.0 : 0.0;
  if (t == "bornDipole") {
    const double after = energy - c.threshold;
    if (after <= 0) return iso;
    const double dk2 = after + energy - 2.0 * std::sqrt(energy * after) * std::cos(angle);
    const double se = std::sqrt(energy), sa = std::sqrt(after);
    return 1.0 / (4.0 * PI) * se * sa / dk2 / std::log((sa + se) / std::sqrt(c.threshold));
  }
  if (t == "surendra") {
    if (energy == 0) return iso;
    return energy / (4.0 * PI * (1.0 + energy * std::pow(std::sin(angle / 2.0), 2)) * std::log(1.0 + energy));
  }
  if (t == "coulombScreen") {
    const double eps = (c.angularParams[0] == 0) ? energy : energy - c.threshold;
    if (eps <= 0) return iso;
    const double s = c.angularParams[1] / eps;
    return (s * (s + 1.0)) / PI / std::pow(2.0 * s + 1.0 - std::cos(angle), 2);
  }
  throw SetupError("The angular distribution function '" + t + "' is not defined.");
}

int angularModelId(const std::string& t) {
  static const char* names[] = {"isotropic", "forward", "bornDipole", "surendra", "coulombScreen", "momentumConservationIonization"};
  for (int i = 0; i < 6; ++i) if (t == names[i]) return i;
  throw SetupError("The angular scattering function '" + t + "' is not defined.");
}

}  // namespace

std::vector<double> interpolatedCrossSection(const Collision& c, bool momTransf, const std::vector<double>& energies) { return interpolatedImpl(c, momTransf, energies); }
std::vector<double> superElasticCrossSection(const Collision& c, bool momTransf, const std::vector<double>& energies) {
  if (!c.isReverse) throw SetupError("Collision '" + c.description() + " is not defined as bidirectional.\n");
  return superElasticImpl(c, momTransf, energies);
}

// ------------------------------------------------------------------ ontology ------------------------------------------------------------------
std::string Collision::description() const {   // Collision.C:88-125
  std::string s = "e+" + target->name + (isReverse ? "<->" : "->");
  if (type == "Ionization") s += "e+e+"; else if (type != "Attachment") s += "e+";
  for (size_t i = 0; i < products.size(); ++i) {
    if (productStoi[i] > 1) s += std::to_string(static_cast<int>(productStoi[i]));
    s += products[i]->name;
    if (i + 1 < products.size()) s += "+";
  }
  return s + "," + type;
}

Gas* Mixture::addGas(const std::string& name) {
  for (auto& g : gases) if (g->name == name) return g.get();
  auto g = std::make_unique<Gas>();
  g->id = static_cast<int>(gases.size()); g->name = name;
  gases.push_back(std::move(g));
  return gases.back().get();
}

std::vector<State*> Mixture::findStates(const std::string& gas, const std::string& ion, const std::string& ele, const std::string& vib, const std::string& rot) const {   // State.h findPointer
  std::vector<State*> out;
  if (ele == "*") {
    for (auto& s : states) if (s->gas->name == gas && s->type == "ele") { out.push_back(s.get()); for (State* q : s->siblings) out.push_back(q); break; }
  } else if (vib == "*") {
    for (auto& s : states) if (s->gas->name == gas && s->eleLevel == ele && s->type == "ele") { out = s->children; break; }
  } else if (rot == "*") {
    for (auto& s : states) if (s->gas->name == gas && s->eleLevel == ele && s->vibLevel == vib && s->type == "vib") { out = s->children; break; }
  } else {
    for (auto& s : states) if (s->gas->name == gas && s->ionCharg == ion && s->eleLevel == ele && s->vibLevel == vib && s->rotLevel == rot) { out.push_back(s.get()); break; }
  }
  return out;
}

State* Mixture::addState(Gas* g, const std::string& ion, const std::string& ele, const std::string& vib, const std::string& rot) {   // State.h add + EedfState.C ctor
  const auto found = findStates(g->name, ion, ele, vib, rot);
  if (!found.empty()) return found[0];
  auto sp = std::make_unique<State>();
  State* s = sp.get();
  s->id = static_cast<int>(states.size()); s->gas = g; s->ionCharg = ion; s->eleLevel = ele; s->vibLevel = vib; s->rotLevel = rot;
  s->type = !ion.empty() ? "ion" : !rot.empty() ? "rot" : !vib.empty() ? "vib" : "ele";
  for (State* o : g->states) {   // EedfState::addFamily
    if (s->type == "rot") {
      if (o->type == "vib" && o->eleLevel == ele && o->vibLevel == vib) { s->parent = o; o->children.push_back(s); }
      else if (o->type == "rot" && o->eleLevel == ele && o->vibLevel == vib) { s->siblings.push_back(o); o->siblings.push_back(s); }
    } else if (s->type == "vib") {
      if (o->type == "ele" && o->eleLevel == ele) { s->parent = o; o->children.push_back(s); }
      else if (o->type == "vib" && o->eleLevel == ele) { s->siblings.push_back(o); o->siblings.push_back(s); }
      else if (o->type == "rot" && o->eleLevel == ele && o->vibLevel == vib) { s->children.push_back(o); o->parent = s; }
    } else if (s->type == "ele") {
      if (o->type == "ele") { s->siblings.push_back(o); o->siblings.push_back(s); }
      else if (o->type == "vib" && o->eleLevel == ele) { s->children.push_back(o); o->parent = s; }
    } else if (o->type == "ion") { s->siblings.push_back(o); o->siblings.push_back(s); }
  }
  s->name = g->name + "(" + (ion.empty() ? "" : ion + ",") + ele + (vib.empty() ? "" : ",v=" + vib + (rot.empty() ? "" : ",J=" + rot)) + ")";   // State::evaluateName
  g->states.push_back(s);
  states.push_back(std::move(sp));
  return s;
}

Collision* Mixture::addCollision(const std::string& type, State* target, const std::vector<State*>& products, const std::vector<double>& stoi, bool isReverse,
                                 double threshold, const CrossSection& integral, const CrossSection& momTransf, bool isExtra) {   // Collision::add / find / ctor
  for (Collision* c : target->collisions) {
    if (c->threshold != threshold || c->type != type || c->isReverse != isReverse || c->products.size() != products.size()) continue;
    bool same = true;
    for (size_t j = 0; j < products.size() && same; ++j) {
      bool hit = false;
      for (size_t k = 0; k < products.size(); ++k) if (c->products[j] == products[k] && c->productStoi[j] == stoi[k]) { hit = true; break; }
      same = hit;
    }
    if (!same) continue;
    if (c->rawIntegral.empty() && !integral.empty()) c->rawIntegral = integral;
    else if (c->rawMomTransf.empty() && !momTransf.empty()) c->rawMomTransf = momTransf;
    else warnings.push_back("Warning! Avoiding duplicated electron impact collision:\n\t " + c->description());
    return c;
  }
  auto cp = std::make_unique<Collision>();
  Collision* c = cp.get();
  c->id = static_cast<int>(collisions.size()); c->type = type; c->target = target; c->products = products; c->productStoi = stoi;
  c->isExtra = isExtra; c->isReverse = isReverse; c->threshold = threshold; c->rawIntegral = integral; c->rawMomTransf = momTransf;
  for (CrossSection* xs : {&c->rawIntegral, &c->rawMomTransf}) {   // Collision.C:33-57
    if (xs->empty()) continue;
    for (size_t i = 0; i + 1 < xs->e.size(); ++i)
      if (xs->e[i] >= xs->e[i + 1]) { collisions.push_back(std::move(cp)); throw SetupError("The energy column of the cross section corresponding to the collision shown below is not strictly increasing.\n" + c->description() + "\nPlease fix this in the LXCat files."); }
    if ((type == "Elastic" || type == "Effective") && xs->e[0] != 0) { xs->e.insert(xs->e.begin(), 0.0); xs->v.insert(xs->v.begin(), xs->v[0]); }
  }
  if (isExtra) { target->gas->collisionsExtra.push_back(c); target->collisionsExtra.push_back(c); }
  else { target->gas->collisions.push_back(c); target->collisions.push_back(c); target->isTarget = true; }
  if (isReverse) {
    if (products.size() == 1 && stoi[0] == 1) { if (isExtra) products[0]->collisionsExtra.push_back(c); else { products[0]->collisions.push_back(c); products[0]->isTarget = true; } }
    else { collisions.push_back(std::move(cp)); throw SetupError("Error while creating collision '" + c->description() + "'. Klein-Rosseland microreversibility relation valid only for binary collisions.\n"); }
  }
  if ((type == "Effective" || type == "Elastic") && target->type != "ele") { collisions.push_back(std::move(cp)); throw SetupError("Found ''" + type + "'' collision with ''" + target->name + "'' as target. " + type + " collisions are only allowed for electronic states. Please check LXCat files.\n"); }
  collisions.push_back(std::move(cp));
  return c;
}

void Mixture::loadLXCat(const SetupTree& tree, const std::string& key, bool isExtra) {   // Setup.h:292-349
  for (const auto& e : parseLXCatFiles(tree.inputDir(), tree.childNames(key))) {
    Gas* g = addGas(e.target[0].gas);
    State* target = addState(g, e.target[0].ion, e.target[0].ele, e.target[0].vib, e.target[0].rot);
    std::vector<State*> prods;
    for (const auto& p : e.products) prods.push_back(addState(addGas(p.gas), p.ion, p.ele, p.vib, p.rot));
    addCollision(e.type, target, prods, e.productStoi, e.isReverse, e.threshold, e.integral, e.momTransf, isExtra);
  }
}

static std::pair<std::vector<double>, std::vector<std::string>> functionArguments(const std::string& prop, const WorkingConditions& wc, const SetupTree& tree) {   // Setup.h:556-571
  std::vector<double> num; std::vector<std::string> str;
  if (prop.find('@') == std::string::npos) return {num, str};
  str = splitAny(splitAny(prop, "@")[1], ",");
  const auto wcm = tree.map("workingConditions");
  for (const auto& a : str) num.push_back(wcm.count(a) ? wc.get(a) : evalExpression(a));
  return {num, str};
}

void Mixture::gasProperties(const SetupTree& tree, const WorkingConditions& wc) {   // Setup.h:460-511, GPF.h
  static const char* props[] = {"mass", "harmonicFrequency", "anharmonicFrequency", "rotationalConstant", "lennardJonesDistance", "lennardJonesDepth",
                                "electricDipolarMoment", "electricQuadrupoleMoment", "polarizability", "fraction", "heatCapacity", "thermalConductivity", "OPBParameter"};
  for (const char* p : props) {
    for (const auto& kv : tree.map(std::string("electronKinetics.gasProperties.") + p)) {
      Gas* g = nullptr;
      for (auto& q : gases) if (q->name == kv.first) g = q.get();
      if (!g) continue;
      const auto args = functionArguments(kv.second, wc, tree);
      const std::string fn = splitAny(kv.second, "@")[0];
      const double T = wc.gasTemperature;
      if (fn == "nitrogenHeatCapacity" || fn == "oxygenHeatCapacity") {
        const bool n2 = fn[0] == 'n';
        double c = (n2 ? 29.1 + 2494.2 / (553.4 * std::sqrt(PI / 2.0)) * std::exp(-2.0 * std::pow((T - 1047.4) / 553.4, 2))
                       : 28.8 + 6456.2 / (788.3 * std::sqrt(PI / 2.0)) * std::exp(-2.0 * std::pow((T - 1006.9) / 788.3, 2)));
        if (args.first.size() != 1) throw SetupError("Wrong number of arguments when evaluating " + fn + " function.");
        if (args.first[0] != 1) c += -AVOGADRO * KB;
        g->prop[p] = c / QE;
      } else if (fn == "nitrogenThermalConductivity") g->prop[p] = (1.717 + 0.084 * T - 1.948e-5 * std::pow(T, 2)) * 1e-3 / QE;
      else if (fn == "oxygenThermalConductivity") g->prop[p] = (1.056 + 0.087 * T - 8.912e-6 * std::pow(T, 2)) * 1e-3 / QE;
      else if (kv.second.find('@') != std::string::npos) throw SetupError("Error! Trying to use the property function '" + fn + "' which is not defined in the code.");
      else g->prop[p] = evalExpression(kv.second);
    }
  }
  double norm = 0;   // Gas::checkFractionNorm
  for (auto& g : gases) norm += g->get("fraction");
  if (std::fabs(norm - 1) > 10 * std::numeric_limits<double>::epsilon())
    throw SetupError("Gas fractions are not properly normalized (Error = " + std::to_string(norm - 1) + "). Please, check input file.\n");
}

void Mixture::stateProperties(const SetupTree& tree, const WorkingConditions& wc) {   // Setup.h:513-551, SPF.h
  static const char* props[] = {"energy", "statisticalWeight", "reducedDiffCoeff", "reducedMobility", "population"};
  for (const char* pc : props) {
    const std::string p = pc;
    for (const auto& kv : tree.map("electronKinetics.stateProperties." + p)) {
      const RawState rs = parseStateName(kv.first);
      std::vector<State*> sel = findStates(rs.gas, rs.ion, rs.ele, rs.vib, rs.rot);
      if (sel.empty()) continue;
      const auto args = functionArguments(kv.second, wc, tree);
      const std::string fn = splitAny(kv.second, "@")[0];
      auto wrongProp = [&](const char* want) { if (p != want) throw SetupError("Trying to use " + fn + " function to set up property " + p + ". \nCheck input file"); };
      auto temperature = [&](size_t i) {   // SPF.h:173-183
        if (args.second.size() <= i) throw SetupError("Wrong number of arguments when evaluating " + fn + " function. \nCheck input file");
        if (args.second[i] == "gasTemperature") return wc.gasTemperature;
        if (args.second[i] == "electronTemperature") return wc.electronTemperature * QE / KB;
        return args.first[i];
      };
      auto needEnergyAndWeight = [&]() {
        for (State* s : sel) {
          if (s->energy == NON_DEF) throw SetupError("Unable to find " + s->name + " energy for the evaluation of " + fn + " function. \nCheck input file");
          if (s->statisticalWeight == NON_DEF) throw SetupError("Unable to find " + s->name + " statistical weight for the evaluation of " + fn + " function. \nCheck input file");
        }
      };
      auto boltzmann = [&](double T, int cutoffKind, double cutoff) {   // SPF.h:172-266
        needEnergyAndWeight();
        double norm = 0, ground = 1E20;
        for (State* s : sel) ground = std::fmin(ground, s->energy);
        for (State* s : sel) {
          const bool cut = (cutoffKind == 1 && std::stod(s->vibLevel) > cutoff) || (cutoffKind == 2 && std::stod(s->rotLevel) > cutoff);
          if (cut) s->population = 0;
          else { s->population = s->statisticalWeight * std::exp(-(s->energy - ground) / (KB_EV * T)); norm += s->population; }
        }
        for (State* s : sel) s->population /= norm;
      };
      if (fn == "rotationalDegeneracy") { wrongProp("statisticalWeight"); for (State* s : sel) { const double J = std::stod(s->rotLevel); s->statisticalWeight = 2.0 * J + 1.0; } }
      else if (fn == "rotationalDegeneracy_H2") { wrongProp("statisticalWeight"); for (State* s : sel) { const double J = std::stod(s->rotLevel); s->statisticalWeight = (2.0 - std::pow(-1, J)) * (2.0 * J + 1.0); } }
      else if (fn == "rotationalDegeneracy_N2") { wrongProp("statisticalWeight"); for (State* s : sel) { const double J = std::stod(s->rotLevel); s->statisticalWeight = 3.0 * (1.0 + 0.5 * (1.0 + std::pow(-1, J))) * (2.0 * J + 1.0); } }
      else if (fn == "rotationalDegeneracy_NO") { wrongProp("statisticalWeight"); for (State* s : sel) { const double J = std::stod(s->rotLevel); s->statisticalWeight = 3.0 * (2.0 * J + 1.0); } }
      else if (fn == "rotationalDegeneracy_H2O") {
        wrongProp("statisticalWeight");
        for (State* s : sel) {
          const double J = std::stod(s->rotLevel.substr(0, 1)), Ka = std::stod(s->rotLevel.substr(1, 1)), Kc = std::stod(s->rotLevel.substr(2, 1));
          double w = 2.0 * J + 1;
          if (std::fmod(std::fabs(Ka - Kc), 2) != 0) w *= 3.0;
          s->statisticalWeight = w;
        }
      }
      else if (fn == "harmonicOscillatorEnergy") { wrongProp("energy"); for (State* s : sel) { const double v = std::stod(s->vibLevel); s->energy = HBAR_EV * s->gas->get("harmonicFrequency") * (v + 0.5); } }
      else if (fn == "morseOscillatorEnergy") { wrongProp("energy"); for (State* s : sel) { const double v = std::stod(s->vibLevel); s->energy = HBAR_EV * (s->gas->get("harmonicFrequency") * (v + 0.5) - s->gas->get("anharmonicFrequency") * std::pow((v + 0.5), 2)); } }
      else if (fn == "rigidRotorEnergy") { wrongProp("energy"); for (State* s : sel) { const double J = std::stod(s->rotLevel); s->energy = s->gas->get("rotationalConstant") * J * (J + 1.0); } }
      else if (fn == "rigidRotorEnergy_NO") {
        wrongProp("energy");
        const bool half = sel[0]->eleLevel == "X_1/2";
        const double B = half ? 0.2073e-3 : 0.2133e-3, om = half ? 0.5 : 1.5;
        for (State* s : sel) { const double J = std::stod(s->rotLevel); s->energy = B * (J * (J + 1.0) - om * om); }
      }
      else if (fn == "boltzmannPopulation") { wrongProp("population"); boltzmann(temperature(0), 0, 0); }
      else if (fn == "boltzmannPopulationVibrationalCutoff") { wrongProp("population"); boltzmann(temperature(0), 1, args.first.at(1)); }
      else if (fn == "boltzmannPopulationRotationalCutoff") { wrongProp("population"); boltzmann(temperature(0), 2, args.first.at(1)); }
      else if (fn == "treanorPopulation" || fn == "treanorGordietsPopulation") {   // SPF.h:269-409
        wrongProp("population");
        needEnergyAndWeight();
        const double T0 = temperature(0), T1 = temperature(1);
        double ground = 0, first = 0; bool hasG = false, hasF = false;
        for (State* s : sel) { if (s->vibLevel == "0") { ground = s->energy; hasG = true; } else if (s->vibLevel == "1") { first = s->energy; hasF = true; } }
        if (!hasG || !hasF) throw SetupError("Unable to find groundEnergy or firstEnergy to populate state" + sel[0]->name + " and its siblings with function " + fn + ". \nCheck input file.");
        double norm = 0;
        for (State* s : sel) {
          const double v = std::stod(s->vibLevel);
          s->population = s->statisticalWeight * std::exp(-1.0 / KB_EV * (v * (first - ground) * (1.0 / T1 - 1.0 / T0) + (s->energy - ground) / T0));
          norm += s->population;
        }
        for (State* s : sel) s->population /= norm;
        if (fn == "treanorGordietsPopulation") {
          const double vLimit = std::floor(0.5 * (1.0 + (first - ground) * T0 / (HBAR_EV * sel[0]->gas->get("anharmonicFrequency") * T1)));
          size_t iLim = 0;
          for (size_t i = 0; i < sel.size(); ++i) if (std::stod(sel[i]->vibLevel) == vLimit) iLim = i;
          std::vector<double> tg(sel.size());
          double n2 = 0;
          for (size_t i = 0; i < sel.size(); ++i) { const double v = std::stod(sel[i]->vibLevel); tg[i] = (v <= vLimit) ? sel[i]->population : sel[iLim]->population * vLimit / v; n2 += tg[i]; }
          for (size_t i = 0; i < sel.size(); ++i) sel[i]->population = tg[i] / n2;
        }
      }
      else if (fn == "generalizedTemperatureDependentCoeff") { /* transport of heavy species: not used by the electron kinetics */ }
      else if (kv.second.find('@') != std::string::npos) throw SetupError("Error! Trying to use the property function '" + fn + "' which is not defined in the code.");
      else {   // SPF.h constantValue
        const double v = evalExpression(kv.second);
        for (State* s : sel) { if (p == "energy") s->energy = v; else if (p == "statisticalWeight") s->statisticalWeight = v; else if (p == "population") s->population = v; }
      }
    }
  }
}

void Mixture::assignAngularScattering(const SetupTree& tree) {   // Setup.h:351-459, Collision.C:127-199
  double angleNumber = 0;
  if (tree.has("electronKinetics.anisotropicScattering") && (tree.value("electronKinetics.anisotropicScattering.isOn") == "true" || tree.value("electronKinetics.anisotropicScattering.isOn") == "True" ||
                                                               tree.value("electronKinetics.anisotropicScattering.isOn") == "1")) {
    angleNumber = tree.number("electronKinetics.anisotropicScattering.angleNumber");
    std::vector<std::string> lines;
    for (const auto& s : tree.childNames("electronKinetics.anisotropicScattering.collisions")) {
      if (s.find(';') != std::string::npos) lines.push_back(s);
      else { std::istringstream in(SetupTree::readFile(tree.inputDir() + "/" + s)); std::string l; while (std::getline(in, l)) { const std::string c = removeSpaces(stripComment(l)); if (!c.empty()) lines.push_back(c); } }
    }
    for (const auto& line : lines) {
      const auto f = splitAny(line, ";");
      bool found = false;
      auto params = [&](size_t idx) { std::vector<double> v; if (f.size() > idx) for (const auto& t : splitAny(f[idx], ",")) v.push_back(evalExpression(t)); return v; };
      if (f[0] == "group") {
        if (f.size() < 4) throw SetupError("Error while reading the following setup line of electronKinetics.anisotropicScattering.collisions:\n" + line);
        for (auto& g : gases) if (g->name == f[1]) for (Collision* c : g->collisions) if (c->type == f[2]) { c->angularType = f[3]; c->angularParams = params(4); found = true; }
      } else if (f[0] == "single") {
        if (f.size() < 3) throw SetupError("Error while reading the following setup line of electronKinetics.anisotropicScattering.collisions:\n" + line);
        for (auto& c : collisions) if (c->description() == f[1]) { c->angularType = f[2]; c->angularParams = params(3); found = true; break; }
      } else throw SetupError("Error while reading the following setup line of electronKinetics.anisotropicScattering.collisions:\n" + line);
      if (!found) throw SetupError("Error while reading the following setup line of electronKinetics.anisotropicScattering.collisions:\n" + line + "\nThe collision (group or single) was not found.");
    }
  }
  // angles of the trapezoidal integration (Collision.C:130-132): Eigen::ArrayXd::LinSpaced(n, 0, pi)
  const int n = static_cast<int>(angleNumber);
  std::vector<double> angles, aux;
  if (n > 1) {
    const double step = (PI - 0.0) / static_cast<double>(n - 1);
    for (int i = 0; i < n; ++i) { const double a = (i == n - 1) ? PI : 0.0 + static_cast<double>(i) * step; angles.push_back(a); aux.push_back(std::cos(a) * std::sin(a)); }
  }
  const double angleStep = PI / (angleNumber - 1.0);
  for (auto& g : gases) for (Collision* c : g->collisions) {
    const int nParams = (c->angularType == "coulombScreen") ? 2 : 0;
    angularModelId(c->angularType);
    if (static_cast<int>(c->angularParams.size()) != nParams) throw SetupError("The angular distribution function '" + c->angularType + "' requires exactly " + std::to_string(nParams) + " parameters, and not " + std::to_string(c->angularParams.size()));
    const bool hasI = !c->rawIntegral.empty(), hasM = !c->rawMomTransf.empty();
    if (!(hasI && hasM)) {
      if (c->angularType == "forward") {
        if (!hasI) throw SetupError("When choosing the ''forward'' angularScatteringType for the collision\n" + c->description() + "\nthe integralCrossSection must be defined in the LXCat files, and not the momentum-transfer!");
        c->rawMomTransf.e = c->rawIntegral.e; c->rawMomTransf.v.assign(c->rawIntegral.e.size(), 0.0);
      } else if (c->angularType == "isotropic") { if (hasI) c->rawMomTransf = c->rawIntegral; else c->rawIntegral = c->rawMomTransf; }
      else {
        if (n < 2) throw SetupError("anisotropicScattering.angleNumber must be at least 2");
        CrossSection& have = hasI ? c->rawIntegral : c->rawMomTransf;
        CrossSection& want = hasI ? c->rawMomTransf : c->rawIntegral;
        want.e = have.e; want.v.assign(have.e.size(), 0.0);
        for (size_t i = 0; i < have.e.size(); ++i) {
          const double energy = have.e[i];
          double integ = aux[0] * angularDistribution(*c, energy, angles[0]);
          for (int j = 1; j < n - 1; ++j) integ += 2.0 * aux[j] * angularDistribution(*c, energy, angles[j]);
          integ += aux[n - 1] * angularDistribution(*c, energy, angles[n - 1]);
          integ *= PI * angleStep;
          want.v[i] = hasI ? have.v[i] * (1.0 - integ) : have.v[i] / (1.0 - integ);
        }
      }
    }
    if (c->angularType != "isotropic" && (c->type == "Effective" || c->type == "Attachment"))
      throw SetupError("Error in the following collision:\n" + c->description() + "\n'" + c->type + "' collisions cannot have a user-prescribed angular scattering model.");
  }
}

CrossSection Mixture::elasticFromEffective(Gas* g) {   // EedfGas.C:171-301
  Collision* eff = nullptr;
  for (Collision* c : g->collisions) if (c->type == "Effective") { eff = c; break; }
  if (!eff) throw SetupError("Gas ''" + g->name + "'' does not have an ''Effective'' collision defined, so ''Elastic'' collisions cannot be evaluated from it.\nPlease, check the corresponding LXCat file.");
  CrossSection el = eff->rawMomTransf;
  int maxID = g->states[0]->id;
  for (State* s : g->states) maxID = std::max(maxID, s->id);
  auto& pop = g->effectivePopulations;
  if (pop.empty()) {
    pop.assign(static_cast<size_t>(maxID) + 1, 0.0);
    State* ground = eff->target;
    pop[ground->id] = 1;
    State* vibGround = nullptr;
    double norm = 0;
    auto need = [&](State* s) {
      if (s->energy == NON_DEF) throw SetupError("Unable to find " + s->name + " energy for the evaluation of ''Elastic'' cross section of " + s->gas->name + ".\nCheck input file");
      if (s->statisticalWeight == NON_DEF) throw SetupError("Unable to find " + s->name + " statistical weight for the evaluation of ''Elastic'' cross section of " + s->gas->name + ".\nCheck input file");
    };
    if (!ground->children.empty()) {
      vibGround = ground->children[0];
      for (State* s : ground->children) {
        need(s);
        if (s->energy < vibGround->energy) vibGround = s;
        pop[s->id] = s->statisticalWeight * std::exp(-s->energy / (KB_EV * 300.0));
        norm += pop[s->id];
      }
    }
    for (State* s : ground->children) pop[s->id] = pop[s->id] / norm;
    if (vibGround && !vibGround->children.empty()) {
      norm = 0;
      for (State* s : vibGround->children) { need(s); pop[s->id] = s->statisticalWeight * std::exp(-s->energy / (KB_EV * 300.0)); norm += pop[s->id]; }
      for (State* s : vibGround->children) pop[s->id] = pop[vibGround->id] * pop[s->id] / norm;
    }
  }
  while (pop.size() < static_cast<size_t>(maxID) + 1) pop.push_back(0.0);
  for (Collision* c : g->collisions) {
    if (c->type == "Effective" || c->type == "Elastic") continue;
    const auto xs = interpolatedImpl(*c, true, el.e);
    for (size_t i = 0; i < el.e.size(); ++i) el.v[i] -= pop[c->target->id] * xs[i];
    if (c->isReverse) {
      const auto sup = superElasticImpl(*c, true, el.e);
      for (size_t i = 0; i < el.e.size(); ++i) el.v[i] -= pop[c->products[0]->id] * sup[i];
    }
  }
  bool negative = false;
  for (auto& v : el.v) if (v < 0) { negative = true; v = 0; }
  if (negative) warnings.push_back("Negative values obtained when evaluating an Elastic cross section from an Effective one (" + g->name + ").\nNegative values have been clipped to 0 and unreliable results may be obtained.\nPlease, carefully check inputs and outputs of your simulation");
  return el;
}

void Mixture::checkPopulationNorms(const Gas* g) const {   // EedfGas.C:29-99: the populations of sibling TARGET states add to one
  if (g->get("fraction") == 0) return;
  const double tol = 10 * std::numeric_limits<double>::epsilon();
  double gasNorm = 0;
  bool eleToCheck = true, ionToCheck = true;
  for (const State* st : g->states) {
    if (st->type == "ele" && eleToCheck) {
      std::vector<const State*> eles(st->siblings.begin(), st->siblings.end());
      eles.insert(eles.begin(), st);
      for (const State* e : eles) {
        if (!(e->isTarget && e->population != 0)) continue;
        gasNorm = gasNorm + e->population;
        if (e->children.empty()) continue;
        double vibNorm = 0;
        bool anyVib = false;
        for (const State* v : e->children) {
          if (!(v->isTarget && v->population != 0)) continue;
          anyVib = true;
          vibNorm = vibNorm + v->population;
          if (v->children.empty()) continue;
          double rotNorm = 0;
          for (const State* r : v->children) if (r->isTarget && r->population != 0) rotNorm = rotNorm + r->population;
          if (std::fabs(rotNorm - 1) > tol)
            throw SetupError("Rotational distribution " + v->name.substr(0, v->name.size() - 1) + ",J=*) is not properly normalized. (Error = " + std::to_string(rotNorm - 1) + ")\n");
        }
        if (anyVib && std::fabs(vibNorm - 1) > tol)
          throw SetupError("Vibrational distribution " + e->name.substr(0, e->name.size() - 1) + ",v=*) is not properly normalized. (Error = " + std::to_string(vibNorm - 1) + ")\n");
      }
      eleToCheck = false;
    }
    if (st->type == "ion" && ionToCheck) {
      if (st->population != 0) gasNorm = gasNorm + st->population;
      for (const State* i : st->siblings) if (i->population != 0) gasNorm = gasNorm + i->population;
      ionToCheck = false;
    }
  }
  if (std::fabs(gasNorm - 1) > tol)
    throw SetupError("Electronic/ionic distribution " + g->name + "(*) is not properly normalized. (Error = " + std::to_string(gasNorm - 1) + ")\n");
}

void Mixture::checkElasticCollisions(Gas* g) {   // EedfGas.C:123-169
  for (State* st : g->states) {
    if (st->type != "ele") continue;
    std::vector<State*> eleStates = st->siblings;
    eleStates.insert(eleStates.begin(), st);
    for (State* e : eleStates) {
      if (!e->isTarget) continue;
      bool hasElastic = false;
      for (Collision* c : e->collisions) if (c->type == "Elastic") { hasElastic = true; break; }
      if (hasElastic) continue;
      const CrossSection raw = elasticFromEffective(g);
      addCollision("Elastic", e, {e}, {1.0}, false, 0.0, raw, raw, false);
    }
    break;
  }
}

Mixture::Mixture(const SetupTree& tree, const WorkingConditions& wc) {   // Setup.h:229-290
  loadLXCat(tree, "electronKinetics.LXCatFiles", false);
  if (tree.has("electronKinetics.LXCatFilesExtra")) loadLXCat(tree, "electronKinetics.LXCatFilesExtra", true);
  for (size_t i = 0; i < states.size(); ++i) if (states[i]->type == "rot" && !states[i]->parent) addState(states[i]->gas, states[i]->ionCharg, states[i]->eleLevel, states[i]->vibLevel, "");   // fixOrphanStates
  for (size_t i = 0; i < states.size(); ++i) if (states[i]->type == "vib" && !states[i]->parent) addState(states[i]->gas, states[i]->ionCharg, states[i]->eleLevel, "", "");
  gasProperties(tree, wc);
  stateProperties(tree, wc);
  for (auto& s : states) {   // State::evaluateDensity
    const double f = s->gas->get("fraction");
    if (s->type == "rot") s->density = s->population * s->parent->population * s->parent->parent->population * f;
    else if (s->type == "vib") s->density = s->population * s->parent->population * f;
    else s->density = s->population * f;
  }
  const auto effPop = tree.numericMap("electronKinetics.effectiveCrossSectionPopulations");
  if (!effPop.empty()) for (auto& g : gases) g->effectivePopulations.assign(g->states.size(), 0.0);
  for (const auto& kv : effPop) {
    const RawState rs = parseStateName(kv.first);
    for (State* s : findStates(rs.gas, rs.ion, rs.ele, rs.vib, rs.rot)) { auto& v = s->gas->effectivePopulations; if (static_cast<size_t>(s->id) >= v.size()) v.resize(static_cast<size_t>(s->id) + 1, 0.0); v[s->id] = kv.second; }
  }
  assignAngularScattering(tree);
  for (auto& g : gases) {
    if (g->collisions.empty()) continue;
    if (g->get("mass") == NON_DEF) throw SetupError("Mass of gas " + g->name + " not found.\nNeeded for the evaluation of the elastic collision operator (Boltzmann).\nCheck input file.");
    checkPopulationNorms(g.get());
    checkElasticCollisions(g.get());
  }
}

ProcessSet Mixture::flatten(double gasTemperature) const {   // BMC.C:29-271, :438-455
  ProcessSet ps;
  for (const auto& g : gases) {
    if (g->collisions.empty()) continue;
    ps.gasFirst.push_back(static_cast<int32_t>(ps.type.size()));
    ps.gasFraction.push_back(g->get("fraction"));
    for (const Collision* c : g->collisions) {
      if (c->type == "Effective") continue;
      const int32_t t = c->type == "Ionization" ? 1 : c->type == "Attachment" ? 2 : 0;
      std::vector<double> e = c->rawIntegral.e, v = c->rawIntegral.v;
      while (!e.empty() && e[0] < c->threshold) { e.erase(e.begin()); v.erase(v.begin()); }   // :139-142
      if (e.empty() || e[0] != c->threshold) { e.insert(e.begin(), c->threshold); v.insert(v.begin(), 0.0); }   // :144-147
      const double M = c->target->gas->get("mass");
      ps.type.push_back(t); ps.isSuperelastic.push_back(0); ps.isElastic.push_back(c->type == "Elastic");
      ps.angularModel.push_back(angularModelId(c->angularType));
      ps.ap0.push_back(c->angularParams.size() >= 2 ? c->angularParams[0] : 0.0); ps.ap1.push_back(c->angularParams.size() >= 2 ? c->angularParams[1] : 0.0);
      ps.swf.push_back(0.0); ps.emin.push_back(c->threshold); ps.emax.push_back(e.back());
      if (c->type == "Elastic") ps.energyMaxElastic = std::fmin(ps.energyMaxElastic, e.back());
      ps.relDensity.push_back(c->target->density);
      ps.targetMass.push_back(M); ps.reducedMass.push_back(ME * M / (ME + M)); ps.energyLoss.push_back(c->threshold);
      ps.thermalStd.push_back(std::sqrt(KB * gasTemperature / M));
      const double opb = c->target->gas->get("OPBParameter");
      ps.wParameter.push_back(c->type == "Ionization" ? (opb == NON_DEF ? c->threshold : opb) : 0.0);
      ps.xsOffset.push_back(static_cast<int64_t>(ps.xsEnergy.size()));
      ps.xsEnergy.insert(ps.xsEnergy.end(), e.begin(), e.end()); ps.xsValue.insert(ps.xsValue.end(), v.begin(), v.end());
      ps.descriptions.push_back(c->description()); ps.collisionOf.push_back(c);
      if (c->angularType == "momentumConservationIonization" && t != 1) throw   // BMC.C:195-198
        SetupError("Trying to assign the angularScatteringType 'momentumConservationIonization' to the process\n" + c->description() + "\nwhich is not 'Ionization'");
      if (c->isReverse) {   // :209-266
        const State* prod = c->products[0];
        const double Mp = prod->gas->get("mass");
        ps.type.push_back(0); ps.isSuperelastic.push_back(1); ps.isElastic.push_back(0);
        ps.angularModel.push_back(ps.angularModel.back()); ps.ap0.push_back(ps.ap0.back()); ps.ap1.push_back(ps.ap1.back());
        ps.swf.push_back(c->target->statisticalWeight / prod->statisticalWeight);
        ps.emin.push_back(0.0); ps.emax.push_back(e.back() - c->threshold);
        ps.relDensity.push_back(prod->density);
        ps.targetMass.push_back(Mp); ps.reducedMass.push_back(ME * Mp / (ME + Mp)); ps.energyLoss.push_back(-c->threshold);
        ps.thermalStd.push_back(std::sqrt(KB * gasTemperature / Mp));
        ps.wParameter.push_back(0.0);
        ps.xsOffset.push_back(static_cast<int64_t>(ps.xsEnergy.size()));
        ps.xsEnergy.push_back(0.0); ps.xsValue.push_back(0.0);
        ps.descriptions.push_back(c->description()); ps.collisionOf.push_back(c);
      }
    }
    ps.gasLast.push_back(static_cast<int32_t>(ps.type.size()) - 1);
  }
  ps.xsOffset.push_back(static_cast<int64_t>(ps.xsEnergy.size()));
  if (ps.type.empty()) throw SetupError("no electron collisions found in the LXCat files");
  return ps;
}

lokib200_process_soa ProcessSet::soa() const {
  lokib200_process_soa p{};
  p.n_processes = static_cast<int32_t>(type.size()); p.n_gases = static_cast<int32_t>(gasFirst.size());
  p.type = type.data(); p.is_superelastic = isSuperelastic.data(); p.angular_model = angularModel.data(); p.angular_p0 = ap0.data(); p.angular_p1 = ap1.data();
  p.superelastic_weight_factor = swf.data(); p.energy_min = emin.data(); p.energy_max = emax.data(); p.rel_density = relDensity.data();
  p.target_mass = targetMass.data(); p.reduced_mass = reducedMass.data(); p.energy_loss = energyLoss.data(); p.thermal_std = thermalStd.data();
  p.w_parameter = wParameter.data(); p.gas_first = gasFirst.data(); p.gas_last = gasLast.data(); p.gas_fraction = gasFraction.data();
  p.xs_offset = xsOffset.data(); p.xs_energy = xsEnergy.data(); p.xs_value = xsValue.data();
  return p;
}

// ------------------------------------------------------------------ setup self-diagnostic ------------------------------------------------------------------
// Setup::selfDiagnostic (Setup.h:576-945), the checks that apply to a boltzmannMC setup; messages are the reference's.
static void selfDiagnostic(const SetupTree& t) {
  auto missing = [](const std::string& field, const std::string& section) {
    return SetupError("Error found in the configuration of the setup file.\n''" + field + "'' field not found in the ''" + section +
                      "'' section of the setup file.\nPlease, fix the problem and run the code again.");
  };
  auto wrong = [](const std::string& field, const std::string& should, bool quoted = true) {
    const std::string q = quoted ? "''" : "";
    return SetupError("Error found in the configuration of the setup file.\nWrong value for the field " + q + field + q + ".\nValue should " + should +
                      ".\nPlease, fix the problem and run the code again.");
  };
  auto logical = [](const std::string& v) { return v == "true" || v == "True" || v == "1" || v == "0" || v == "false" || v == "False"; };
  auto isTrue = [](const std::string& v) { return v == "true" || v == "True" || v == "1"; };
  auto num = [&](const std::string& k) { return t.has(k) ? t.number(k) : 0.0; };
  const std::string ek = "electronKinetics";
  if (t.has(ek)) {
    if (!t.has(ek + ".isOn")) throw missing("isOn", ek);
    if (!logical(t.value(ek + ".isOn"))) throw wrong("electronKinetics>isOn", "be logical (''true'' or ''false'')", false);
    if (isTrue(t.value(ek + ".isOn"))) {
      if (!t.has(ek + ".eedfType")) throw missing("eedfType", ek);
      const std::string type = t.value(ek + ".eedfType");
      if (type != "boltzmannMC" && type != "prescribedEedf") throw wrong("electronKinetics>eedfType", "be: ''boltzmannMC'' or ''prescribedEedf''");
      if (type == "boltzmannMC") {
        if (!t.has(ek + ".ionizationOperatorType")) throw missing("ionizationOperatorType", ek);
        const std::string ion = t.value(ek + ".ionizationOperatorType");
        if (ion != "oneTakesAll" && ion != "equalSharing" && ion != "usingSDCS" && ion != "randomUniform")
          throw wrong("electronKinetics>ionizationOperatorType", "be either: ''oneTakesAll'', ''equalSharing'', ''usingSDCS'' or ''randomUniform''");
      }
      if (!t.has(ek + ".LXCatFiles")) throw missing("LXCatFiles", ek);
      if (!t.has(ek + ".gasProperties")) throw missing("gasProperties", ek);
      if (!t.has(ek + ".gasProperties.mass")) throw missing("mass", "electronKinetics>gasProperties");
      if (!t.has(ek + ".gasProperties.fraction")) throw missing("fraction", "electronKinetics>gasProperties");
      if (!t.has(ek + ".stateProperties")) throw missing("stateProperties", ek);
      if (!t.has(ek + ".stateProperties.population")) throw missing("population", "electronKinetics>stateProperties");
      if (type.find("boltzmannMC") != std::string::npos) {
        const std::string mc = ek + ".numericsMC";
        if (!t.has(mc)) throw missing("numericsMC", ek);
        if (!t.has(mc + ".nElectrons")) throw missing("nElectrons", "electronKinetics>numericsMC");
        if (!t.has(mc + ".gasTemperatureEffect")) throw missing("gasTemperatureEffect", "electronKinetics>numericsMC");
        const std::string gt = t.value(mc + ".gasTemperatureEffect");
        if (gt != "false" && gt != "true" && gt != "smartActivation") throw wrong("electronKinetics>numericsMC", "be ''false'', ''true'' or ''smartActivation''", false);
        for (const char* k : {"nEnergyCells", "nCosAngleCells", "nAxialVelocityCells", "nRadialVelocityCells", "nInterpPoints"})
          if ((t.has(mc + "." + k) && num(mc + "." + k) <= 0) || std::fmod(num(mc + "." + k), 1) != 0)
            throw wrong(std::string("electronKinetics>numericsMC>") + k, "be a single positive integer");
        if (t.has(mc + ".synchronizationTimeXMaxCollisionFrequency") && num(mc + ".synchronizationTimeXMaxCollisionFrequency") <= 0)
          throw wrong("electronKinetics>numericsMC>synchronizationTimeXMaxCollisionFrequency", "be a single positive number");
        if ((t.has(mc + ".synchronizationOverSampling") && num(mc + ".synchronizationOverSampling") < 1) || std::fmod(num(mc + ".synchronizationOverSampling"), 1) != 0)
          throw wrong("electronKinetics>numericsMC>synchronizationOverSampling", "be a single integer >= 1");
        if (t.has(mc + ".initialElecTempOverGasTemp") && num(mc + ".initialElecTempOverGasTemp") <= 0)
          throw wrong("electronKinetics>numericsMC>initialElecTempOverGasTemp", "be a single positive number");
        if (t.has(mc + ".minCollisionsBeforeSteadyState") && num(mc + ".minCollisionsBeforeSteadyState") < 0)
          throw wrong("electronKinetics>numericsMC>minCollisionsBeforeSteadyState", "be a single non-negative number");
        for (const char* k : {"maxCollisionsBeforeSteadyState", "maxCollisionsAfterSteadyState"})
          if (t.has(mc + "." + k) && num(mc + "." + k) <= 0) throw wrong(std::string("electronKinetics>numericsMC>") + k, "be a single positive number");
        if ((t.has(mc + ".nIntegrationPoints") && num(mc + ".nIntegrationPoints") < 500) || std::fmod(num(mc + ".nIntegrationPoints"), 1) != 0)
          throw wrong("electronKinetics>numericsMC>nIntegrationPoints", "be a single integer >= 500");
        if ((t.has(mc + ".nIntegrationPhases") && num(mc + ".nIntegrationPhases") < 1) || std::fmod(num(mc + ".nIntegrationPhases"), 1) != 0)
          throw wrong("electronKinetics>numericsMC>nIntegrationPhases", "be a single integer > 1");
        for (const char* k : {"nIntegratedSSTimes", "integratedAbsoluteTime"})
          if (t.has(mc + "." + k) && num(mc + "." + k) <= 0) throw wrong(std::string("electronKinetics>numericsMC>") + k, "be a single positive number");
        for (const char* k : {"meanEnergy", "fluxDriftVelocity", "fluxDiffusionCoeffs", "bulkDriftVelocity", "bulkDiffusionCoeffs", "powerBalance"})
          if (t.has(mc + ".relError." + k) && num(mc + ".relError." + k) <= 0) throw wrong(std::string("electronKinetics>numericsMC>relError>") + k, "be a single positive number");
      }
      const std::string an = ek + ".anisotropicScattering";
      if (t.has(an)) {
        if (!t.has(an + ".isOn")) throw missing("isOn", "electronKinetics>anisotropicScattering");
        if (!logical(t.value(an + ".isOn"))) throw wrong("electronKinetics>anisotropicScattering>isOn", "be logical (''true'' or ''false'')", false);
        if (isTrue(t.value(an + ".isOn"))) {
          if (!t.has(an + ".collisions")) throw missing("collisions", "electronKinetics>anisotropicScattering");
          if (!t.has(an + ".angleNumber")) throw missing("angleNumber", "electronKinetics>anisotropicScattering");
          if (num(an + ".angleNumber") <= 0 || std::fmod(num(an + ".angleNumber"), 1) != 0) throw wrong("electronKinetics>anisotropicScattering>angleNumber", "be a single positive integer");
        }
      }
    }
  }
  if (!isTrue(t.value(ek + ".isOn")))
    throw SetupError("Error found in the configuration of the setup file.\n''electronKinetics'' module is not activated.\nPlease, fix the problem and run the code again.");
  for (const char* sec : {"gui", "output"}) {
    const std::string s2 = sec;
    if (!t.has(s2)) continue;
    if (!t.has(s2 + ".isOn")) throw missing("isOn", s2);
    if (!logical(t.value(s2 + ".isOn"))) throw wrong(s2 + ">isOn", "be logical (''true'' or ''false'')", false);
  }
  if (t.has("output") && isTrue(t.value("output.isOn"))) {
    if (!t.has("output.dataFiles")) throw missing("dataFiles", "output");
    static const char* files[] = {"eedf", "evdf", "swarmParameters", "rateCoefficients", "powerBalance", "lookUpTable", "MCTemporalInfo", "MCTemporalInfo_periodic", "MCSimDetails"};
    for (const auto& f : t.childNames("output.dataFiles")) {
      bool ok = false;
      for (const char* k : files) ok = ok || f == k;
      if (!ok) throw SetupError("Error found in the configuration of the setup file.\nWrong value for the field ''gui>dataFiles''. Possible data files are: eedf, evdf, swarmParameters, "
                                "rateCoefficients, powerBalance, lookUpTable, MCTemporalInfo, MCTemporalInfo_periodic, MCSimDetails.\nPlease, fix the problem and run the code again.");
    }
  }
}

// ------------------------------------------------------------------ one setup file ------------------------------------------------------------------
SetupInput::SetupInput(const std::string& inputDir, const std::string& setupFile) {
  tree = std::make_unique<SetupTree>(inputDir, SetupTree::readFile(setupFile.rfind("/", 0) == 0 ? setupFile : inputDir + "/" + setupFile));   // an absolute path is taken as is
  selfDiagnostic(*tree);
  const std::string eedfType = tree->value("electronKinetics.eedfType");
  if (eedfType != "boltzmannMC") throw SetupError("Please choose a valid 'electronKinetics->eedfType' in the setup file: this build implements 'boltzmannMC' (found '" + eedfType + "').");
  for (const char* key : {"electronKinetics.ionizationOperatorType", "electronKinetics.LXCatFiles", "electronKinetics.numericsMC.nElectrons", "electronKinetics.numericsMC.gasTemperatureEffect"})
    if (!tree->has(key)) throw SetupError(std::string("The mandatory field '") + key + "' is missing in the setup file");
  wc = makeWorkingConditions(*tree);
  mixture = std::make_unique<Mixture>(*tree, wc);
  processes = mixture->flatten(wc.gasTemperature);
}

double SetupInput::jobValue(int job) const {
  if (wc.variableCondition == "reducedMagField") return wc.reducedMagFieldArray[job];
  if (wc.variableCondition == "elecFieldAngle") return wc.elecFieldAngleArray[job];
  if (wc.variableCondition == "excitationFrequency") return wc.excitationFrequencyArray[job];
  return wc.reducedElecFieldArray[job];
}

lokib200_config SetupInput::config(int job) const {   // BMC.h:262-289, BMC.C:431-489
  lokib200_config c{};
  auto pick = [&](const std::vector<double>& a, const char* name) { return a[(wc.variableCondition == name) ? static_cast<size_t>(job) : 0]; };
  const double EN = pick(wc.reducedElecFieldArray, "reducedElecField"), BN = pick(wc.reducedMagFieldArray, "reducedMagField");
  const double angle = pick(wc.elecFieldAngleArray, "elecFieldAngle"), freq = pick(wc.excitationFrequencyArray, "excitationFrequency");
  c.n_electrons = static_cast<int64_t>(tree->number("electronKinetics.numericsMC.nElectrons"));
  c.seed = 0x4C6F4B49ull + static_cast<uint64_t>(job);   // reproducible by default; LOKIB200_SEED selects another realisation
  if (const char* env = std::getenv("LOKIB200_SEED")) c.seed += 0x9E3779B97F4A7C15ull * std::strtoull(env, nullptr, 10);
  const std::string gt = tree->value("electronKinetics.numericsMC.gasTemperatureEffect");
  c.gas_temperature_effect = gt == "true" ? 1 : gt == "smartActivation" ? 2 : 0;
  const std::string ion = tree->value("electronKinetics.ionizationOperatorType");
  c.ionization_sharing = ion == "oneTakesAll" ? 1 : ion == "usingSDCS" ? 2 : ion == "randomUniform" ? 3 : 0;
  c.energy_sharing_factor = ion == "oneTakesAll" ? 0.0 : 0.5;
  c.is_cylindrically_symmetric = wc.isCylindricallySymmetric;
  c.gas_density = wc.gasDensity; c.gas_temperature = wc.gasTemperature;
  const double E = (EN * 1e-21) * wc.gasDensity;                         // WC.h:81, BMC.C:463
  if (angle == 180) { c.electric_field[0] = 0; c.electric_field[1] = 0; c.electric_field[2] = -E; }
  else { c.electric_field[0] = E * std::sin(angle / 180.0 * PI); c.electric_field[1] = 0; c.electric_field[2] = E * std::cos(angle / 180.0 * PI); }
  if (freq != 0) for (double& x : c.electric_field) x *= std::sqrt(2);
  c.excitation_omega = freq * 2.0 * PI;
  c.cyclotron_omega = QE * ((BN * 1e-27) * wc.gasDensity) / ME;
  auto opt = [&](const char* k, double dflt) { const std::string p = std::string("electronKinetics.numericsMC.") + k; return tree->has(p) ? tree->number(p) : dflt; };
  c.n_interp_points = static_cast<int32_t>(opt("nInterpPoints", 1e4));
  c.n_energy_cells = static_cast<int32_t>(opt("nEnergyCells", 1000)); c.n_cos_cells = static_cast<int32_t>(opt("nCosAngleCells", 100));
  c.n_radial_cells = static_cast<int32_t>(opt("nRadialVelocityCells", 200)); c.n_axial_cells = static_cast<int32_t>(opt("nAxialVelocityCells", 200));
  c.n_phases = static_cast<int32_t>(opt("nIntegrationPhases", 100));
  return c;
}

lokib200_solve_controls SetupInput::controls() const {   // BMC.h:291-365
  lokib200_solve_controls s{};
  const double n = tree->number("electronKinetics.numericsMC.nElectrons");
  auto has = [&](const char* k) { return tree->has(std::string("electronKinetics.numericsMC.") + k); };
  auto num = [&](const char* k) { return tree->number(std::string("electronKinetics.numericsMC.") + k); };
  s.n_integration_points = has("nIntegrationPoints") ? num("nIntegrationPoints") : 200;
  s.n_integrated_ss_times = has("nIntegratedSSTimes") ? num("nIntegratedSSTimes") : 0;
  s.integrated_absolute_time = has("integratedAbsoluteTime") ? num("integratedAbsoluteTime") : 0;
  s.errors_to_be_checked = has("relError");
  s.rel_err_mean_energy = has("relError.meanEnergy") ? num("relError.meanEnergy") : 1e100;
  s.rel_err_flux_drift = has("relError.fluxDriftVelocity") ? num("relError.fluxDriftVelocity") : 1e100;
  s.rel_err_bulk_drift = has("relError.bulkDriftVelocity") ? num("relError.bulkDriftVelocity") : 1e100;
  s.rel_err_flux_diff = has("relError.fluxDiffusionCoeffs") ? num("relError.fluxDiffusionCoeffs") : 1e100;
  s.rel_err_bulk_diff = has("relError.bulkDiffusionCoeffs") ? num("relError.bulkDiffusionCoeffs") : 1e100;
  s.rel_err_power_balance = has("relError.powerBalance") ? num("relError.powerBalance") : 1e100;
  s.min_collisions_before_ss = has("minCollisionsBeforeSteadyState") ? num("minCollisionsBeforeSteadyState") * n : 0;
  s.max_collisions_before_ss = has("maxCollisionsBeforeSteadyState") ? num("maxCollisionsBeforeSteadyState") * n : 1e100;
  s.max_collisions_after_ss = has("maxCollisionsAfterSteadyState") ? num("maxCollisionsAfterSteadyState") * n : 1e100;
  s.sync_factor = has("synchronizationTimeXMaxCollisionFrequency") ? num("synchronizationTimeXMaxCollisionFrequency") : 1;
  s.sync_over_sampling = has("synchronizationOverSampling") ? static_cast<int32_t>(num("synchronizationOverSampling")) : 1;
  s.initial_temp_ratio = has("initialElecTempOverGasTemp") ? num("initialElecTempOverGasTemp") : 0.01;
  s.energy_max_elastic = processes.energyMaxElastic;
  s.fast_mode = has("fastMode") ? (tree->value("electronKinetics.numericsMC.fastMode") == "true" ? 1 : 0) : 0;   // not a reference key: default = reference behaviour
  const std::string gui = tree->has("gui.isOn") ? tree->value("gui.isOn") : "false";
  if (gui == "true" || gui == "True" || gui == "1")
    for (const auto& o : tree->childNames("gui.terminalDisp")) if (o == "MCStatus") s.status_display = 1;   // BMC.h:368-374
  return s;
}

}  // namespace lokihost
