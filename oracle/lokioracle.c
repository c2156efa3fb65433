/* oracle/lokioracle.c -- TEST INFRASTRUCTURE ONLY.  See lokioracle.h.
 *
 * Plain-C restatement of the electron Monte Carlo hot path of LoKI-MC v1.1.0.  "BMC.C" below is
 * /root/reference/Code/LoKI-MC/Sources/BoltzmannMC.C, "Math.C" is Sources/MathFunctions.C, "ASF.h" is
 * Headers/AngularScatteringFunctions.h.  Arithmetic keeps the reference's operation order where cancellation matters so
 * that results agree with the reference far below the 1e-12 parity tolerance.
 * Compiled with -ffp-contract=off (the reference is built by g++ -O2 on x86-64, i.e. without FMA contraction).
 */
#include "lokioracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Constant.h:5-23 (CODATA-2014 values as used by the reference) */
static const double KB = 1.38064852e-23, QE = 1.6021766208e-19, ME = 9.10938356e-31;
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { T_CONSERVATIVE = 0, T_IONIZATION = 1, T_ATTACHMENT = 2 };                        /* GeneralDefinitions.h:34-36 */
enum { SH_EQUAL = 0, SH_ONE_TAKES_ALL = 1, SH_SDCS = 2, SH_UNIFORM = 3 };               /* GeneralDefinitions.h:43-46 */
enum { GT_FALSE = 0, GT_TRUE = 1, GT_SMART = 2 };                                       /* GeneralDefinitions.h:49-51 */
enum { A_ISOTROPIC = 0, A_FORWARD = 1, A_BORN_DIPOLE = 2, A_SURENDRA = 3, A_COULOMB = 4, A_MOMCONS_ION = 5 };

struct lo_model {
  int P, nG, nE;
  int32_t *type, *superel, *angular, *gas_first, *gas_last;
  double *ap0, *ap1, *swf, *emin, *emax, *reldens, *mass, *redmass, *eloss, *thstd, *wpar, *gas_fraction;
  int64_t* xs_off;
  double *xs_e, *xs_v;
  int gastemp, sharing;
  double sharing_factor, Ngas, Tg, gas_energy, E[3], aE[3], omega, Omega;
  double maxE, dE, *sigma, *cum, *nu_tot, *nu_max;
};

static void* dupmem(const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (p && bytes) memcpy(q, p, bytes); return q; }

lo_model* lo_model_create(int P, int nG, const int32_t* type, const int32_t* superel, const int32_t* angular, const double* ap0,
                          const double* ap1, const double* swf, const double* emin, const double* emax, const double* reldens,
                          const double* mass, const double* redmass, const double* eloss, const double* thstd, const double* wpar,
                          const int32_t* gas_first, const int32_t* gas_last, const double* gas_fraction, const int64_t* xs_off,
                          const double* xs_e, const double* xs_v) {
  lo_model* m = (lo_model*)calloc(1, sizeof(lo_model));
  m->P = P; m->nG = nG; m->nE = 10000;
#define DUPI(f) m->f = (int32_t*)dupmem(f, sizeof(int32_t) * (size_t)P)
#define DUPD(f) m->f = (double*)dupmem(f, sizeof(double) * (size_t)P)
  DUPI(type); DUPI(superel); DUPI(angular);
  DUPD(ap0); DUPD(ap1); DUPD(swf); DUPD(emin); DUPD(emax); DUPD(reldens); DUPD(mass); DUPD(redmass); DUPD(eloss); DUPD(thstd); DUPD(wpar);
  m->gas_first = (int32_t*)dupmem(gas_first, sizeof(int32_t) * (size_t)nG);
  m->gas_last = (int32_t*)dupmem(gas_last, sizeof(int32_t) * (size_t)nG);
  m->gas_fraction = (double*)dupmem(gas_fraction, sizeof(double) * (size_t)nG);
  m->xs_off = (int64_t*)dupmem(xs_off, sizeof(int64_t) * (size_t)(P + 1));
  m->xs_e = (double*)dupmem(xs_e, sizeof(double) * (size_t)xs_off[P]);
  m->xs_v = (double*)dupmem(xs_v, sizeof(double) * (size_t)xs_off[P]);
  m->maxE = LO_NON_DEF;
  return m;
}

/* job constants: BMC.C:431-489.  electric_field is the vector of BMC.C:465-481 (already x sqrt(2) when AC). */
void lo_model_set_conditions(lo_model* m, int gastemp, int sharing, double sharing_factor, double Ngas, double Tg, const double E[3],
                             double omega, double Omega, int nE) {
  m->gastemp = gastemp; m->sharing = sharing; m->sharing_factor = sharing_factor;
  m->Ngas = Ngas; m->Tg = Tg;
  m->gas_energy = 1.5 * KB * Tg / QE;                               /* BMC.C:433 */
  for (int c = 0; c < 3; ++c) { m->E[c] = E[c]; m->aE[c] = -QE / ME * E[c]; } /* BMC.C:489 */
  m->omega = omega; m->Omega = Omega;
  if (nE != m->nE) { m->nE = nE; free(m->sigma); free(m->cum); free(m->nu_tot); free(m->nu_max); m->sigma = m->cum = m->nu_tot = m->nu_max = NULL; m->maxE = LO_NON_DEF; }
}

void lo_model_destroy(lo_model* m) {
  if (!m) return;
  free(m->type); free(m->superel); free(m->angular); free(m->gas_first); free(m->gas_last); free(m->ap0); free(m->ap1); free(m->swf);
  free(m->emin); free(m->emax); free(m->reldens); free(m->mass); free(m->redmass); free(m->eloss); free(m->thstd); free(m->wpar);
  free(m->gas_fraction); free(m->xs_off); free(m->xs_e); free(m->xs_v); free(m->sigma); free(m->cum); free(m->nu_tot); free(m->nu_max);
  free(m);
}

/* GSL gsl_interp_linear semantics (interp/linear.c + gsl_interp_bsearch): i with x[i] <= xv < x[i+1], clamped to [0,n-2] */
static double lin_interp(const double* x, const double* y, int64_t n, double xv) {
  int64_t lo = 0, hi = n - 1;
  while (hi > lo + 1) { int64_t mid = (hi + lo) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
  double dx = x[lo + 1] - x[lo];
  return y[lo] + (xv - x[lo]) / dx * (y[lo + 1] - y[lo]);
}

/* BMC.C:561-615 */
void lo_build_tables(lo_model* m, double maxE) {
  if (maxE == m->maxE) return;                                      /* :567 */
  const int nE = m->nE, P = m->P;
  if (!m->sigma) {
    m->sigma = (double*)malloc(sizeof(double) * (size_t)nE * P); m->cum = (double*)malloc(sizeof(double) * (size_t)nE * P);
    m->nu_tot = (double*)malloc(sizeof(double) * nE); m->nu_max = (double*)malloc(sizeof(double) * nE);
  }
  m->maxE = maxE;
  m->dE = maxE / (double)(nE - 1);                                  /* :573 */
  double running_max = 0;
  for (int i = 0; i < nE; ++i) {
    double energy = i * m->dE, acc = 0;
    for (int k = 0; k < P; ++k) {
      double value = 0;
      if (m->superel[k]) {                                          /* Klein-Rosseland :584-595 */
        if (energy > m->emin[k] && energy <= m->emax[k]) {
          const int64_t o = m->xs_off[k - 1], n = m->xs_off[k] - o;
          value = m->swf[k] * (1.0 + m->emin[k - 1] / energy) * lin_interp(m->xs_e + o, m->xs_v + o, n, energy + m->emin[k - 1]) * m->reldens[k];
        }
      } else if (energy >= m->emin[k] && energy <= m->emax[k]) {    /* :597-603 */
        const int64_t o = m->xs_off[k], n = m->xs_off[k + 1] - o;
        value = lin_interp(m->xs_e + o, m->xs_v + o, n, energy) * m->reldens[k];
      }
      m->sigma[(size_t)i * P + k] = value;
      acc += value;
      m->cum[(size_t)i * P + k] = acc;
    }
    acc *= m->Ngas * sqrt(energy * 2.0 * QE / ME);                  /* :610 */
    m->nu_tot[i] = acc;
    running_max = fmax(acc, running_max);
    m->nu_max[i] = running_max;
  }
}
int lo_table_size(const lo_model* m) { return m->nE; }
double lo_table_step(const lo_model* m) { return m->dE; }
const double* lo_table_sigma(const lo_model* m) { return m->sigma; }
const double* lo_table_cum(const lo_model* m) { return m->cum; }
const double* lo_table_nu_tot(const lo_model* m) { return m->nu_tot; }
const double* lo_table_nu_max(const lo_model* m) { return m->nu_max; }

/* BMC.C:765-802 */
double lo_max_accel_energy(const lo_model* m, double e0, double dt) {
  const double e_me = QE / ME;
  const double Ex0 = fabs(m->E[0]), Ez0 = fabs(m->E[2]);
  const double Ex02 = Ex0 * Ex0, Ez02 = Ez0 * Ez0, E02 = Ex02 + Ez02, E0 = sqrt(E02);
  const double v0 = sqrt(e0 * QE * 2.0 / ME);
  const double w = m->omega, W = m->Omega;
  double gain = (E0 * v0 + 0.5 * e_me * E02 * dt) * dt;
  if (W == 0) {
    if (w != 0) gain = fmin(gain, 2.0 / w * (e_me * E02 / w + v0 * (Ex0 + Ez0)));
  } else if (w == 0) {
    gain = fmin(gain, 0.5 * e_me * Ez02 * dt * dt + (2.0 * e_me * Ex02 / W + 3.0 * v0 * Ex0) / W + v0 * dt * Ez0);
  } else if (fabs(w - W) / w < 1E-6) {
    const double W2 = W * W;
    gain = fmin(gain, 2.0 * e_me * Ez02 / W2 + e_me * Ex02 / (8.0 * W2) * (4.0 + W * dt * (2.0 + W * dt)) + (v0 * dt + v0 / W) * Ex0 + 2.0 * v0 * Ez0 / W);
  } else {
    const double w2 = w * w, W2 = W * W, d = w2 - W2;
    gain = fmin(gain, 2.0 * e_me * Ez02 / w2 + 0.5 * e_me * Ex02 / (d * d) * (5.0 * w2 + 8.0 * w * W + 5.0 * W * W) + 3.0 * v0 * Ex0 / fabs(w - W) + 2.0 * v0 * Ez0 / w);
  }
  return e0 + gain;
}

/* ------------------------------------------------------------------ random draws ------------------------------------------------------------------ */

void lo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t hi, uint32_t lo) {   /* (k + 0.5) 2^-52, k = top 52 bits (k + 0.5 is exact in a double): strictly inside (0,1) */
  const uint64_t k = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)k + 0.5) * (1.0 / 4503599627370496.0);
}

/* counter = (id_lo, id_hi, interval, j/2), key = (seed_lo, seed_hi): two uniforms per Philox call */
double lo_stream_uniform(uint64_t seed, uint64_t id, uint32_t interval, uint32_t j) {
  uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), interval, j >> 1}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  lo_philox4x32_10(ctr, key, o);
  return (j & 1) ? u53(o[3], o[2]) : u53(o[1], o[0]);
}

typedef struct {
  const double* inj; int n_inj;          /* injected mode when inj != NULL */
  uint64_t seed, id; uint32_t interval;  /* counter mode otherwise */
  uint32_t used;
} draws_t;

/* counter mode only: every free-time draw starts on an even draw index, so that the (t_cf, R) pair of a null event is ONE Philox call */
static inline void align_draws(draws_t* d) { if (!d->inj && (d->used & 1u)) ++d->used; }

static inline double draw(draws_t* d) {
  double u;
  if (d->inj) u = ((int)d->used < d->n_inj) ? d->inj[d->used] : 0.5;
  else u = lo_stream_uniform(d->seed, d->id, d->interval, d->used);
  ++d->used;
  return u;
}

/* ------------------------------------------------------------------ per-electron physics ------------------------------------------------------------------ */

static inline double kinetic_energy_eV(const double v[3]) { return 0.5 * ME * ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) / QE; }

/* Math.C:143-163 */
static void cart2sph(const double a[3], double* norm, double* sT, double* cT, double* sP, double* cP) {
  const double x = a[0], y = a[1], z = a[2], xy2 = x * x + y * y, nxy = sqrt(xy2);
  *norm = sqrt(xy2 + z * z);
  *sT = nxy / *norm; *cT = z / *norm;
  if (nxy != 0) { *sP = y / nxy; *cP = x / nxy; } else { *sP = 1; *cP = 0; }
}

/* Math.C:129-141 (Yousfi 1994) */
static void euler(double sC, double cC, double sE, double cE, double sT, double cT, double sP, double cP, double out[3]) {
  const double sCsE = sC * sE, aux = sC * cE * cT + cC * sT;
  out[0] = -sCsE * sP + aux * cP;
  out[1] = sCsE * cP + aux * sP;
  out[2] = -sC * cE * sT + cC * cT;
}

/* accelerateElectron, BMC.C:804-905.  Returns the field energy gain (:904). */
static double flight(const lo_model* m, double t_e, double dt, double r[3], double v[3], double* eps) {
  const double prev = *eps, w = m->omega, W = m->Omega;
  const double Ex = m->E[0], Ez = m->E[2];
  if (W == 0) {
    if (w == 0) {                                                     /* DC, B = 0 :812-815 */
      for (int c = 0; c < 3; ++c) { r[c] += v[c] * dt + (m->aE[c] * (0.5 * dt * dt)); v[c] += (m->aE[c] * dt); }
    } else {                                                          /* AC, B = 0 :816-831 */
      const double phi = w * t_e, sPhi = sin(phi), cPhi = cos(phi), wdt = w * dt, phase = wdt + phi, sPh = sin(phase), cPh = cos(phase);
      const double e_me_w = QE / (ME * w);
      const double aux1 = e_me_w / w * (cPh + wdt * sPhi - cPhi), aux2 = e_me_w * (sPhi - sPh);
      r[0] += v[0] * dt + Ex * aux1; r[1] += v[1] * dt; r[2] += v[2] * dt + Ez * aux1;
      v[0] += Ex * aux2; v[2] += Ez * aux2;
    }
  } else {
    const double vx0 = v[0], vy0 = v[1], vz0 = v[2];
    if (w == 0) {                                                     /* DC E + B || z :835-849 */
      const double Wdt = W * dt, s = sin(Wdt), c = cos(Wdt), s_W = s / W, aux2 = (c - 1.0) / W;
      const double vEx = QE * Ex / (ME * W), az = QE * Ez / ME;
      r[0] += vx0 * s_W + (vy0 + vEx) * aux2;
      r[1] += -vx0 * aux2 + vy0 * s_W + vEx * (s_W - dt);
      r[2] += vz0 * dt - 0.5 * az * dt * dt;
      v[0] = vx0 * c - vy0 * s - vEx * s;
      v[1] = vx0 * s + vy0 * c + vEx * (c - 1.0);
      v[2] = vz0 - az * dt;
    } else if (fabs(w - W) / w < 1E-6) {                              /* electron cyclotron resonance :850-872 */
      const double phi = W * t_e, sPhi = sin(phi), cPhi = cos(phi), Wdt = W * dt, s = sin(Wdt), c = cos(Wdt);
      const double phase = Wdt + phi, sPh = sin(phase), cPh = cos(phase), cOpp = cos(phi - Wdt);
      const double e_me_W = QE / (ME * W), vEx = e_me_W * Ex, vEz = e_me_W * Ez;
      const double s_W = s / W, cm1_W = (c - 1.0) / W, cPh_W = cPh / W;
      r[0] += vx0 * s_W + vy0 * cm1_W - 0.25 * vEx * (cPh_W - cOpp / W + 2.0 * dt * sPh);
      r[1] += -vx0 * cm1_W + vy0 * s_W + 0.5 * vEx * (dt * cPh_W - 2.0 * cm1_W * sPhi - cPhi * s_W);
      r[2] += vz0 * dt + vEz * (cPh_W - cPhi / W + dt * sPhi);
      v[0] = vx0 * c - vy0 * s - 0.5 * vEx * (Wdt * cPh + cPhi * s);
      v[1] = vx0 * s + vy0 * c - 0.5 * vEx * (Wdt * c * sPhi + (Wdt * cPhi - sPhi) * s);
      v[2] = vz0 + vEz * (sPhi - sPh);
    } else {                                                          /* general AC E + DC B :873-897 */
      const double phi = w * t_e, sPhi = sin(phi), cPhi = cos(phi), wdt = w * dt, swdt = sin(wdt), cwdt = cos(wdt);
      const double phase = wdt + phi, sPh = sin(phase), cPh = cos(phase), Wdt = W * dt, s = sin(Wdt), c = cos(Wdt);
      const double w2 = w * w, W2 = W * W, d = w2 - W2;
      const double vEx = QE / (ME * W) * Ex, vEz = QE / (ME * w) * Ez, vEx_d = vEx / d, WvEx_d = W * vEx_d;
      const double s_W = s / W, cm1_W = (c - 1.0) / W, wsPhi = w * sPhi;
      r[0] += vx0 * s_W + vy0 * cm1_W + WvEx_d * (cPhi * (cwdt - c) + (wsPhi * s_W - sPhi * swdt));
      r[1] += -vx0 * cm1_W + vy0 * s_W - vEx_d / w * ((W2 + w2 * c - w2) * sPhi + W2 * (w * cPhi * s_W - sPh));
      r[2] += vz0 * dt + vEz * ((cPh - cPhi) / w + dt * sPhi);
      v[0] = vx0 * c - vy0 * s + WvEx_d * w * (c * sPhi - sPh + cPhi / w * W * s);
      v[1] = vx0 * s + vy0 * c + vEx_d * W2 * (cPh - cPhi * c + wsPhi * s_W);
      v[2] = vz0 + vEz * (sPhi - sPh);
    }
  }
  *eps = kinetic_energy_eV(v);                                        /* :901 */
  return *eps - prev;                                                 /* :904 */
}

/* ASF.h:26-74; returns cos(chi) */
static double cos_chi(const lo_model* m, int k, double energy, double energy_after, draws_t* d) {
  switch (m->angular[k]) {
    case A_FORWARD: return 1;
    case A_BORN_DIPOLE: {
      const double sq = sqrt(energy_after) + sqrt(energy), ratio = m->eloss[k] / (sq * sq), r2 = ratio * ratio;
      return 1.0 + 2.0 * r2 / (1.0 - r2) * (1.0 - pow(r2, -draw(d)));
    }
    case A_SURENDRA: return (2.0 + energy - 2.0 * pow(1.0 + energy, draw(d))) / energy;
    case A_COULOMB: {
      const double e = (m->ap0[k] == 0) ? energy : energy_after, s = m->ap1[k] / e, R = draw(d);
      if (!(e > 0)) return 1.0 - 2.0 * R;   /* screening parameter -> inf (oneTakesAll ejects at eps = 0): the reference's expression is inf/inf = NaN there; its limit is isotropic */
      return (s + 1.0 - (2.0 * s + 1.0) * R) / (s + 1.0 - R);
    }
    default: return 1.0 - 2.0 * draw(d);                              /* isotropic */
  }
}

typedef struct { double dE, dE_rel, ej_r[3], ej_v[3], ej_eps; } coll_out;

static int thermal_branch(const lo_model* m, double eps) {
  return m->gastemp == GT_TRUE || (m->gastemp == GT_SMART && eps < 20.0 * m->gas_energy);    /* BMC.C:916, :1125 */
}

/* conservativeCollision, BMC.C:1115-1193.  Returns 0 when the event is relabelled as a null collision (:1137-1140, :1169-1172). */
static int conservative(const lo_model* m, int k, double v[3], const double V[3], double* eps, draws_t* d, coll_out* o) {
  const double M = m->mass[k], mu = m->redmass[k], loss = m->eloss[k], inc = *eps;
  double speed, sT, cT, sP, cP, dir[3];
  if (thermal_branch(m, inc)) {
    const double rel[3] = {v[0] - V[0], v[1] - V[1], v[2] - V[2]};
    cart2sph(rel, &speed, &sT, &cT, &sP, &cP);
    const double erel = 0.5 * mu * speed * speed / QE, eafter = erel - loss;
    if (eafter <= 0) return 0;
    const double cC = cos_chi(m, k, erel, eafter, d), sC = sqrt(1.0 - cC * cC);
    const double eta = 2.0 * M_PI * draw(d), sE = sin(eta), cE = cos(eta);
    const double after = sqrt(speed * speed - 2.0 / mu * loss * QE);
    euler(sC, cC, sE, cE, sT, cT, sP, cP, dir);
    for (int c = 0; c < 3; ++c) v[c] = M / (ME + M) * (after * dir[c]) + (ME * v[c] + M * V[c]) / (ME + M);   /* :1156-1157 */
  } else {
    cart2sph(v, &speed, &sT, &cT, &sP, &cP);
    const double eafter = inc - loss;
    if (eafter <= 0) return 0;
    const double cC = cos_chi(m, k, inc, eafter, d), sC = sqrt(1.0 - cC * cC);
    const double eta = 2.0 * M_PI * draw(d), sE = sin(eta), cE = cos(eta);
    const double after = sqrt((speed * speed - 2.0 / ME * loss * QE) * (1.0 - 2.0 * mu / (ME + M) * (1.0 - cC)));   /* :1184-1185 */
    euler(sC, cC, sE, cE, sT, cT, sP, cP, dir);
    for (int c = 0; c < 3; ++c) v[c] = after * dir[c];
  }
  *eps = kinetic_energy_eV(v);
  o->dE = *eps - inc; o->dE_rel = o->dE / inc;
  return 1;
}

/* ionizationCollision, BMC.C:1195-1272 */
static int ionization(const lo_model* m, int k, const double r[3], double v[3], double* eps, draws_t* d, coll_out* o) {
  const double inc = *eps, I = m->eloss[k];
  if (inc < I) return 0;                                              /* :1202-1205 */
  double speed, sT, cT, sP, cP;
  cart2sph(v, &speed, &sT, &cT, &sP, &cP);
  const double net = inc - I;
  double e_ej;
  if (m->sharing == SH_SDCS) e_ej = m->wpar[k] * tan(draw(d) * atan(net / (2.0 * m->wpar[k])));   /* :1215 */
  else if (m->sharing == SH_UNIFORM) e_ej = draw(d) * net;                                          /* :1218 */
  else e_ej = m->sharing_factor * net;                                                              /* :1221 */
  const double e_sc = net - e_ej;
  double sCs, cCs, sCe, cCe, sEs, cEs, sEe, cEe;
  if (m->angular[k] == A_MOMCONS_ION) {                               /* :1228-1241 (Boeuf 1982) */
    cCs = sqrt(e_sc / net); sCs = sqrt(1.0 - cCs * cCs);
    const double eta = 2.0 * M_PI * draw(d);
    sEs = sin(eta); cEs = cos(eta);
    cCe = sqrt(e_ej / net); sCe = sqrt(1.0 - cCe * cCe);
    sEe = -sEs; cEe = -cEs;
  } else {                                                            /* :1242-1253 */
    cCs = cos_chi(m, k, inc, e_sc, d); sCs = sqrt(1.0 - cCs * cCs);
    const double eta = 2.0 * M_PI * draw(d);
    sEs = sin(eta); cEs = cos(eta);
    cCe = cos_chi(m, k, inc, e_ej, d); sCe = sqrt(1.0 - cCe * cCe);
    const double eta2 = 2.0 * M_PI * draw(d);
    sEe = sin(eta2); cEe = cos(eta2);
  }
  double dir[3];
  euler(sCs, cCs, sEs, cEs, sT, cT, sP, cP, dir);
  const double vs = sqrt(2.0 * e_sc * QE / ME);
  for (int c = 0; c < 3; ++c) v[c] = vs * dir[c];
  euler(sCe, cCe, sEe, cEe, sT, cT, sP, cP, dir);
  const double ve = sqrt(2.0 * e_ej * QE / ME);
  for (int c = 0; c < 3; ++c) { o->ej_v[c] = ve * dir[c]; o->ej_r[c] = r[c]; }
  o->ej_eps = e_ej;
  *eps = e_sc;                                                        /* :1223: stored energy is net - ejected, not recomputed from v */
  o->dE = -I; o->dE_rel = -I / inc;
  return 1;
}

/* Math.C:54-59: three normals from four U(0,1] */
static void normal3(draws_t* d, double g[3]) {
  const double r1 = draw(d), r2 = draw(d), r3 = draw(d), r4 = draw(d);
  const double a1 = sqrt(-2.0 * log(r1)), a2 = 2.0 * M_PI * r2;
  g[0] = a1 * cos(a2); g[1] = a1 * sin(a2); g[2] = sqrt(-2.0 * log(r3)) * cos(2.0 * M_PI * r4);
}

/* performCollision, BMC.C:907-1113.  Returns the chosen process id or LO_NULL_COLLISION. */
static int collide(const lo_model* m, double nu_e, const double r[3], double v[3], double* eps, draws_t* d, coll_out* o) {
  const int nE = m->nE, P = m->P;
  double V[3] = {0, 0, 0};
  int chosen = LO_NULL_COLLISION;
  if (thermal_branch(m, *eps)) {                                      /* :916-1031 */
    double g[3];
    normal3(d, g);
    const double R = nu_e * draw(d) / m->Ngas;
    double prev = 0;
    for (int ig = 0; ig < m->nG && chosen == LO_NULL_COLLISION; ++ig) {
      if (m->gas_fraction[ig] == 0) continue;
      int left = m->gas_first[ig], right = m->gas_last[ig];
      for (int c = 0; c < 3; ++c) V[c] = g[c] * m->thstd[left];
      const double dx = v[0] - V[0], dy = v[1] - V[1], dz = v[2] - V[2];
      const double vrel = sqrt((dx * dx + dy * dy) + dz * dz);
      const double x = 0.5 * m->redmass[left] * vrel * vrel / QE / m->dE;
      const int i1 = (int)fmin(x, nE - 1), i2 = (int)fmin(i1 + 1, nE - 1);
      double w1 = (double)i2 - x;
      w1 = (w1 < 0) ? 0.0 : 1.0;                                      /* :959-966: nearest-lower row, no interpolation */
      const double w2 = 1.0 - w1;
      const double *c1 = m->cum + (size_t)i1 * P, *c2 = m->cum + (size_t)i2 * P, *s1 = m->sigma + (size_t)i1 * P, *s2 = m->sigma + (size_t)i2 * P;
      double ref = 0;
      if (left > 0) ref = w1 * c1[left - 1] + w2 * c2[left - 1];
      const double limit = prev + (w1 * c1[right] + w2 * c2[right] - ref) * vrel;
      if (R > limit) { prev = limit; continue; }
      while (left != right) {
        const int t = (left + right) / 2;
        const double tv = prev + (w1 * c1[t] + w2 * c2[t] - ref) * vrel;
        if (R < tv) right = t; else if (R > tv) left = t + 1; else { chosen = t; break; }
      }
      if (left == right) chosen = left;
      while (w1 * s1[chosen] == 0 && w2 * s2[chosen] == 0) --chosen;  /* :1013-1015 */
    }
    if (chosen == LO_NULL_COLLISION) return chosen;
  } else {                                                            /* cold-gas branch :1034-1097 */
    const double Rnu = nu_e * draw(d);
    const double x = *eps / m->dE;
    const int i1 = (int)fmin(x, nE - 1), i2 = (int)fmin(i1 + 1, nE - 1);
    double w1 = (double)i2 - x;
    if (w1 < 0) w1 = 0.0;
    const double w2 = 1.0 - w1;
    if (Rnu > w1 * m->nu_tot[i1] + w2 * m->nu_tot[i2]) return LO_NULL_COLLISION;
    const double *c1 = m->cum + (size_t)i1 * P, *c2 = m->cum + (size_t)i2 * P, *s1 = m->sigma + (size_t)i1 * P, *s2 = m->sigma + (size_t)i2 * P;
    const double R = Rnu / m->Ngas / sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    int left = 0, right = P - 1;
    while (left != right) {
      const int t = (left + right) / 2;
      const double tv = w1 * c1[t] + w2 * c2[t];
      if (R < tv) right = t; else if (R > tv) left = t + 1; else { chosen = t; break; }
    }
    if (left == right) chosen = left;
    while (w1 * s1[chosen] == 0 && w2 * s2[chosen] == 0) --chosen;    /* :1091-1093 */
  }
  int ok = 1;
  if (m->type[chosen] == T_CONSERVATIVE) ok = conservative(m, chosen, v, V, eps, d, o);
  else if (m->type[chosen] == T_IONIZATION) ok = ionization(m, chosen, r, v, eps, d, o);
  else { o->dE = -*eps; o->dE_rel = -1; }                             /* attachmentCollision :1274-1280 */
  return ok ? chosen : LO_NULL_COLLISION;
}

/* one pass of the loop body BMC.C:637-681 for one electron.  st = [x y z vx vy vz eps t_e t_cf nu_e] */
static int event(const lo_model* m, double nu_trial, double t_sync, double* st, draws_t* d, coll_out* o, double* gain_field) {
  double *r = st, *v = st + 3, *eps = st + 6, *t_e = st + 7, *t_cf = st + 8, *nu_e = st + 9;
  if (*t_cf == LO_NON_DEF) { align_draws(d); *t_cf = -log(draw(d)) / nu_trial; *nu_e = nu_trial; }   /* :650-655 */
  if (*t_e + *t_cf > t_sync) {                                        /* :657-663 */
    const double dt = t_sync - *t_e;
    *gain_field = flight(m, *t_e, dt, r, v, eps);
    *t_e = t_sync; *t_cf -= dt;
    return LO_PARTIAL_FLIGHT;
  }
  *gain_field = flight(m, *t_e, *t_cf, r, v, eps);                    /* :666-675 */
  *t_e += *t_cf;
  const int chosen = collide(m, *nu_e, r, v, eps, d, o);
  align_draws(d);
  *t_cf = -log(draw(d)) / nu_trial;
  *nu_e = nu_trial;
  return chosen;
}

int lo_event_injected(const lo_model* m, double nu_trial, double t_sync, double* state, const double* draws, int n_draws, double* out, int* used) {
  draws_t d; memset(&d, 0, sizeof d); d.inj = draws; d.n_inj = n_draws;
  coll_out o; memset(&o, 0, sizeof o);
  double gain = 0;
  const int chosen = event(m, nu_trial, t_sync, state, &d, &o, &gain);
  out[0] = o.dE; out[1] = o.dE_rel; out[2] = gain;
  for (int c = 0; c < 3; ++c) { out[3 + c] = o.ej_r[c]; out[6 + c] = o.ej_v[c]; }
  out[9] = o.ej_eps;
  if (used) *used = (int)d.used;
  return chosen;
}

/* ------------------------------------------------------------------ sampling ------------------------------------------------------------------ */

/* BMC.C:1426-1454 */
void lo_moments(int64_t n, const double* x, const double* y, const double* z, const double* vx, const double* vy, const double* vz, double* out) {
  const double* r[3] = {x, y, z}; const double* v[3] = {vx, vy, vz};
  double se = 0, maxe = 0, sr[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, rr[9] = {0}, rv[9] = {0};
  for (int64_t i = 0; i < n; ++i) {
    const double vi[3] = {vx[i], vy[i], vz[i]};
    const double e = kinetic_energy_eV(vi);
    se += e; if (e > maxe) maxe = e;
    for (int a = 0; a < 3; ++a) {
      sr[a] += r[a][i]; sv[a] += v[a][i];
      for (int b = 0; b < 3; ++b) { rr[3 * a + b] += r[a][i] * r[b][i]; rv[3 * a + b] += r[a][i] * v[b][i]; }
    }
  }
  out[0] = se / (double)n; out[1] = maxe;
  for (int a = 0; a < 3; ++a) { out[2 + a] = sr[a] / (double)n; out[5 + a] = sv[a] / (double)n; }
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
    out[8 + 3 * a + b] = rr[3 * a + b] / (double)n - out[2 + a] * out[2 + b];       /* :1445 */
    out[17 + 3 * a + b] = rv[3 * a + b] / (double)n - out[2 + a] * out[5 + b];      /* :1446 */
  }
}

/* grids: BMC.C:1862-1883 (LinSpaced nodes); counting: Math.C:61-81 (1-D, only the upper bound is checked) and :106-127 (2-D) */
void lo_histograms(int64_t n, const double* vx, const double* vy, const double* vz, double max_eedf_energy, int nEn, int nC, int nR, int nA,
                   int cylindrical, double* eeh, double* eah, double* evh) {
  /* Eigen LinSpaced(size, low, high)[i] = low + i*(high-low)/(size-1); step = nodes[1]-nodes[0] */
  const double e_step = (0.0 + 1 * (max_eedf_energy - 0.0) / (double)nEn) - 0.0;
  const double c_first = -1.0, c_step = (-1.0 + 1 * (1.0 - (-1.0)) / (double)nC) - c_first;
  const double max_speed = sqrt(2.0 * max_eedf_energy * QE / ME);
  const double r_step = (0.0 + 1 * (max_speed - 0.0) / (double)nR) - 0.0;
  const double a_first = -max_speed, a_step = (-max_speed + 1 * (max_speed - (-max_speed)) / (double)nA) - a_first;
  for (int64_t i = 0; i < n; ++i) {
    const double vi[3] = {vx[i], vy[i], vz[i]};
    const double e = kinetic_energy_eV(vi);
    const int ie = (int)((e - 0.0) / e_step);
    if (ie < nEn) eeh[ie] += 1.0;
    if (cylindrical) {
      const double cosang = vi[2] / sqrt((vi[0] * vi[0] + vi[1] * vi[1]) + vi[2] * vi[2]);
      const int ic = (int)((cosang - c_first) / c_step);
      if (ie < nEn && ie >= 0 && ic < nC && ic >= 0) eah[(size_t)ie * nC + ic] += 1.0;
      const double vr = sqrt(vi[0] * vi[0] + vi[1] * vi[1]);
      const int ir = (int)((vr - 0.0) / r_step), ia = (int)((vi[2] - a_first) / a_step);
      if (ir < nR && ir >= 0 && ia < nA && ia >= 0) evh[(size_t)ir * nA + ia] += 1.0;
    }
  }
}

/* ------------------------------------------------------------------ ensemble (reference semantics) ------------------------------------------------------------------ */

struct lo_ensemble {
  const lo_model* m;
  int64_t n;
  uint64_t seed, id_offset;
  double *x, *y, *z, *vx, *vy, *vz, *eps, *te, *tcf, *nue;
  /* per-electron scratch handed from the parallel pass to the serial pass (BMC.h:226-231) */
  int* chosen; uint32_t* ndraw;
  double *dE, *gain, *ejx, *ejy, *ejz, *ejvx, *ejvy, *ejvz, *ejeps;
  double time;
  uint64_t pc_draws;   /* population-control stream position */
};

lo_ensemble* lo_ensemble_create(const lo_model* m, int64_t n, uint64_t seed, uint64_t id_offset) {
  lo_ensemble* e = (lo_ensemble*)calloc(1, sizeof(lo_ensemble));
  e->m = m; e->n = n; e->seed = seed; e->id_offset = id_offset;
  double** arrs[] = {&e->x, &e->y, &e->z, &e->vx, &e->vy, &e->vz, &e->eps, &e->te, &e->tcf, &e->nue, &e->dE, &e->gain,
                     &e->ejx, &e->ejy, &e->ejz, &e->ejvx, &e->ejvy, &e->ejvz, &e->ejeps};
  for (size_t i = 0; i < sizeof(arrs) / sizeof(arrs[0]); ++i) *arrs[i] = (double*)calloc((size_t)n, sizeof(double));
  e->chosen = (int*)calloc((size_t)n, sizeof(int)); e->ndraw = (uint32_t*)calloc((size_t)n, sizeof(uint32_t));
  return e;
}
void lo_ensemble_destroy(lo_ensemble* e) {
  if (!e) return;
  double* arrs[] = {e->x, e->y, e->z, e->vx, e->vy, e->vz, e->eps, e->te, e->tcf, e->nue, e->dE, e->gain, e->ejx, e->ejy, e->ejz, e->ejvx, e->ejvy, e->ejvz, e->ejeps};
  for (size_t i = 0; i < sizeof(arrs) / sizeof(arrs[0]); ++i) free(arrs[i]);
  free(e->chosen); free(e->ndraw); free(e);
}

#define LO_INIT_INTERVAL 0xFFFFFFFFu   /* draw-counter "interval" reserved for the initial Maxwellian */
#define LO_PC_ID 0xFFFFFFFFFFFFull     /* stream id reserved for the serial population-control draws */

/* BMC.C:491-508 */
void lo_ensemble_init(lo_ensemble* e, double temp_ratio) {
  const double sd = sqrt(KB * temp_ratio * e->m->Tg / ME);
  for (int64_t i = 0; i < e->n; ++i) {
    draws_t d; memset(&d, 0, sizeof d); d.seed = e->seed; d.id = e->id_offset + (uint64_t)i; d.interval = LO_INIT_INTERVAL;
    double g[3]; normal3(&d, g);
    e->x[i] = e->y[i] = e->z[i] = 0;
    e->vx[i] = g[0] * sd; e->vy[i] = g[1] * sd; e->vz[i] = g[2] * sd;
    const double v[3] = {e->vx[i], e->vy[i], e->vz[i]};
    e->eps[i] = kinetic_energy_eV(v);
    e->te[i] = 0; e->tcf[i] = LO_NON_DEF; e->nue[i] = 0;
  }
  e->time = 0;
}
void lo_ensemble_set(lo_ensemble* e, const double* s) {
  const int64_t n = e->n;
  for (int64_t i = 0; i < n; ++i) {
    e->x[i] = s[i]; e->y[i] = s[n + i]; e->z[i] = s[2 * n + i]; e->vx[i] = s[3 * n + i]; e->vy[i] = s[4 * n + i]; e->vz[i] = s[5 * n + i];
    e->tcf[i] = s[6 * n + i]; e->nue[i] = s[7 * n + i];
    const double v[3] = {e->vx[i], e->vy[i], e->vz[i]};
    e->eps[i] = kinetic_energy_eV(v); e->te[i] = e->time;
  }
}
void lo_ensemble_get(const lo_ensemble* e, double* s) {
  const int64_t n = e->n;
  for (int64_t i = 0; i < n; ++i) {
    s[i] = e->x[i]; s[n + i] = e->y[i]; s[2 * n + i] = e->z[i]; s[3 * n + i] = e->vx[i]; s[4 * n + i] = e->vy[i]; s[5 * n + i] = e->vz[i];
    s[6 * n + i] = e->tcf[i]; s[7 * n + i] = e->nue[i];
  }
}
double lo_ensemble_max_energy(const lo_ensemble* e) { double mx = 0; for (int64_t i = 0; i < e->n; ++i) if (e->eps[i] > mx) mx = e->eps[i]; return mx; }
double lo_ensemble_time(const lo_ensemble* e) { return e->time; }

static double pc_uniform(lo_ensemble* e) { const uint64_t j = e->pc_draws++; return lo_stream_uniform(e->seed, LO_PC_ID, (uint32_t)(j >> 32), (uint32_t)j); }

typedef struct { uint64_t real, null, born, attached; uint64_t* counts; double *gain, *loss, field, growth; } tallies;

/* the omp-parallel pass BMC.C:636-681; returns 1 if any electron still has to be advanced */
static int parallel_pass(lo_ensemble* e, double nu_trial, double t_sync, uint32_t interval) {
  int any = 0;
  const lo_model* m = e->m;
#pragma omp parallel for schedule(static) reduction(| : any)
  for (int64_t i = 0; i < e->n; ++i) {
    if (e->te[i] == t_sync) { e->chosen[i] = (int)LO_NON_DEF; continue; }            /* :640-643 */
    double st[10] = {e->x[i], e->y[i], e->z[i], e->vx[i], e->vy[i], e->vz[i], e->eps[i], e->te[i], e->tcf[i], e->nue[i]};
    draws_t d; memset(&d, 0, sizeof d); d.seed = e->seed; d.id = e->id_offset + (uint64_t)i; d.interval = interval; d.used = e->ndraw[i];
    coll_out o; memset(&o, 0, sizeof o);
    double gain = 0;
    const int chosen = event(m, nu_trial, t_sync, st, &d, &o, &gain);
    if (chosen != LO_PARTIAL_FLIGHT) any = 1;
    e->x[i] = st[0]; e->y[i] = st[1]; e->z[i] = st[2]; e->vx[i] = st[3]; e->vy[i] = st[4]; e->vz[i] = st[5];
    e->eps[i] = st[6]; e->te[i] = st[7]; e->tcf[i] = st[8]; e->nue[i] = st[9];
    e->chosen[i] = chosen; e->ndraw[i] = d.used; e->dE[i] = o.dE; e->gain[i] = gain;
    e->ejx[i] = o.ej_r[0]; e->ejy[i] = o.ej_r[1]; e->ejz[i] = o.ej_r[2]; e->ejvx[i] = o.ej_v[0]; e->ejvy[i] = o.ej_v[1]; e->ejvz[i] = o.ej_v[2];
    e->ejeps[i] = o.ej_eps;
  }
  return any;
}

static void place_ejected(lo_ensemble* e, int64_t slot, int64_t parent, double nu_trial) {   /* BMC.C:1348-1353, :1389-1394 */
  e->x[slot] = e->ejx[parent]; e->y[slot] = e->ejy[parent]; e->z[slot] = e->ejz[parent];
  e->vx[slot] = e->ejvx[parent]; e->vy[slot] = e->ejvy[parent]; e->vz[slot] = e->ejvz[parent];
  e->eps[slot] = e->ejeps[parent]; e->te[slot] = e->te[parent]; e->tcf[slot] = LO_NON_DEF; e->nue[slot] = nu_trial;
}

/* nonParallelCollisionTasks, BMC.C:1282-1408 */
static void serial_pass(lo_ensemble* e, double nu_trial, int population_control, tallies* t) {
  const lo_model* m = e->m;
  const int64_t n = e->n;
  int64_t n_ej = 0, n_at = 0, cap = 16;
  int64_t *ej = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap), *at = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
  int64_t cap_at = cap;
  unsigned char* is_att = NULL;
  for (int64_t i = 0; i < n; ++i) {
    const int k = e->chosen[i];
    if (k == (int)LO_NON_DEF) continue;
    t->field += e->gain[i];                                           /* :1303 */
    if (k == LO_NULL_COLLISION) { ++t->null; continue; }
    if (k == LO_PARTIAL_FLIGHT) continue;
    ++t->real;
    if (t->counts) ++t->counts[k];
    if (t->gain) { if (e->dE[i] >= 0) t->gain[k] += e->dE[i]; else t->loss[k] += e->dE[i]; }
    if (m->type[k] == T_IONIZATION) { if (n_ej == cap) { cap *= 2; ej = (int64_t*)realloc(ej, sizeof(int64_t) * (size_t)cap); } ej[n_ej++] = i; }
    else if (m->type[k] == T_ATTACHMENT) { if (n_at == cap_at) { cap_at *= 2; at = (int64_t*)realloc(at, sizeof(int64_t) * (size_t)cap_at); } at[n_at++] = i; }
  }
  t->born += (uint64_t)n_ej; t->attached += (uint64_t)n_at;
  if (population_control && (n_ej || n_at)) {
    if (n_at) { is_att = (unsigned char*)calloc((size_t)n, 1); for (int64_t a = 0; a < n_at; ++a) is_att[at[a]] = 1; }
    for (int64_t a = 0; a < n_at; ++a) {                              /* :1344-1377 */
      const int64_t slot = at[a];
      if (n_ej > 0) { place_ejected(e, slot, ej[n_ej - 1], nu_trial); is_att[slot] = 0; --n_ej; }
      else {
        int64_t j = (int64_t)fmin(pc_uniform(e) * (double)n, (double)(n - 1));
        while (is_att[j]) j = (int64_t)fmin(pc_uniform(e) * (double)n, (double)(n - 1));
        t->growth += e->eps[j];
        e->x[slot] = e->x[j]; e->y[slot] = e->y[j]; e->z[slot] = e->z[j]; e->vx[slot] = e->vx[j]; e->vy[slot] = e->vy[j]; e->vz[slot] = e->vz[j];
        e->eps[slot] = e->eps[j]; e->te[slot] = e->te[j]; e->tcf[slot] = e->tcf[j]; e->nue[slot] = e->nue[j];
      }
    }
    while (n_ej > 0) {                                                /* :1380-1407 */
      int64_t j = (int64_t)fmin(pc_uniform(e) * (double)(n + n_ej), (double)(n + n_ej - 1));
      if (j < n) { t->growth -= e->eps[j]; place_ejected(e, j, ej[n_ej - 1], nu_trial); }
      else { j -= n; t->growth -= e->ejeps[ej[j]]; const int64_t tmp = ej[j]; ej[j] = ej[n_ej - 1]; ej[n_ej - 1] = tmp; }
      --n_ej;
    }
  }
  free(ej); free(at); free(is_att);
}

void lo_ensemble_advance(lo_ensemble* e, double nu_trial, double t_sync, uint32_t interval, int population_control, uint64_t* counters,
                         uint64_t* counts, double* gain, double* loss, double* scal) {
  tallies t; memset(&t, 0, sizeof t); t.counts = counts; t.gain = gain; t.loss = loss;
  memset(e->ndraw, 0, sizeof(uint32_t) * (size_t)e->n);
  int any = 1;
  while (any) {                                                       /* BMC.C:630-684 */
    any = parallel_pass(e, nu_trial, t_sync, interval);
    serial_pass(e, nu_trial, population_control, &t);
  }
  e->time = t_sync;
  if (counters) { counters[0] += t.real; counters[1] += t.null; counters[2] += t.born; counters[3] += t.attached; }
  if (scal) { scal[0] += t.field; scal[1] += t.growth; }
}

/* ------------------------------------------------------------------ whole job ------------------------------------------------------------------ */

/* Math.C:211-231 */
static double stat_error(const double* a, int64_t n, int stride, int nbins) {
  if (n == 0) return 0;
  const int64_t np = n / nbins;
  double mean = 0;
  for (int64_t i = 0; i < n; ++i) mean += a[i * stride];
  mean /= (double)n;
  double sum = 0;
  for (int b = 0; b < nbins; ++b) {
    double bm = 0;
    for (int64_t i = 0; i < np; ++i) bm += a[(b * np + i) * stride];
    bm /= (double)np;
    sum += (bm - mean) * (bm - mean);
  }
  return sqrt(sum) / nbins;
}

/* checkMaxCollisionFrequency, BMC.C:716-763 */
static void check_nu_trial(lo_model* m, const lo_ensemble* e, double* nu_trial) {
  const double max_before = lo_ensemble_max_energy(e);
  const double thermal = (m->gastemp == GT_TRUE || m->gastemp == GT_SMART) ? 10.0 * m->gas_energy : 0;
  double emax_el = 1E100;
  for (int k = 0; k < m->P; ++k) if (m->type[k] == T_CONSERVATIVE && !m->superel[k] && m->eloss[k] == 0) emax_el = fmin(emax_el, m->emax[k]);
  int updated = 1;
  while (updated) {
    updated = 0;
    const double maxE = lo_max_accel_energy(m, max_before, 10.0 / *nu_trial) + thermal;
    if (maxE > m->maxE || 2.5 * maxE < m->maxE) lo_build_tables(m, (2.0 * maxE < emax_el) ? 2.0 * maxE : emax_el);
    const int idx = (int)fmin(ceil(maxE / m->dE), m->nE - 1);
    if (*nu_trial < m->nu_max[idx]) { updated = 1; *nu_trial *= 1.1; }
  }
}

void lo_solve(const lo_model* model, int64_t n, uint64_t seed, const double* ctrl, double* res) {
  lo_model* m = (lo_model*)model;   /* tables are rebuilt in place, like the reference */
  const double need_points = ctrl[0], need_ss_times = ctrl[1], sync_factor = ctrl[2] > 0 ? ctrl[2] : 1.0, temp_ratio = ctrl[3] > 0 ? ctrl[3] : 0.01;
  const int64_t max_intervals = (int64_t)ctrl[4];
  lo_ensemble* e = lo_ensemble_create(m, n, seed, 0);
  lo_ensemble_init(e, temp_ratio);
  lo_build_tables(m, 2.0 * lo_ensemble_max_energy(e));               /* BMC.C:512 */
  double nu_trial = m->nu_max[m->nE - 1];                             /* :515 */
  for (int64_t i = 0; i < n; ++i) e->nue[i] = nu_trial;
  int64_t cap = 1024, ns = 0;
  double* ts = (double*)malloc(sizeof(double) * (size_t)cap);
  double* mom = (double*)malloc(sizeof(double) * 26 * (size_t)cap);  /* per sample: lo_moments output */
  double* bulkv = (double*)malloc(sizeof(double) * 3 * (size_t)cap);
  double* bulkd = (double*)malloc(sizeof(double) * 9 * (size_t)cap);
  tallies t; memset(&t, 0, sizeof t);
  double t_ss = LO_NON_DEF, integrated = 0;
  int64_t first_int = 0, n_int = 0, n_sync = 0;
  uint64_t real_at_ss = 0;
  (void)real_at_ss;
  const double wall0 = omp_get_wtime();
  for (;;) {
    /* sample (BMC.C:310, :341 -> :1410-1464) */
    if (ns == cap) { cap *= 2; ts = (double*)realloc(ts, sizeof(double) * (size_t)cap); mom = (double*)realloc(mom, sizeof(double) * 26 * (size_t)cap);
      bulkv = (double*)realloc(bulkv, sizeof(double) * 3 * (size_t)cap); bulkd = (double*)realloc(bulkd, sizeof(double) * 9 * (size_t)cap); }
    ts[ns] = e->time;
    lo_moments(n, e->x, e->y, e->z, e->vx, e->vy, e->vz, mom + 26 * ns);
    if (ns > 0) {
      const double dt = ts[ns] - ts[ns - 1];
      for (int c = 0; c < 3; ++c) bulkv[3 * ns + c] = (mom[26 * ns + 2 + c] - mom[26 * (ns - 1) + 2 + c]) / dt;
      for (int c = 0; c < 9; ++c) bulkd[9 * ns + c] = 0.5 * (mom[26 * ns + 8 + c] - mom[26 * (ns - 1) + 8 + c]) / dt;
    } else { for (int c = 0; c < 3; ++c) bulkv[c] = mom[5 + c]; for (int c = 0; c < 9; ++c) bulkd[c] = mom[17 + c]; }
    ++ns;
    if (t_ss != LO_NON_DEF) { ++n_int; integrated = e->time - t_ss; }
    else if (ns >= 100) {                                             /* steady-state check cadence :344-345, :363 */
      int dec = (int)fmax(log((double)ns) / log(2.0) - 11, 6);
      const int64_t every = (int64_t)pow(2, dec);
      if (ns % every == 0) {                                          /* checkSteadyState :1787-1815 */
        const double t1 = 0.5 * e->time, t2 = 0.75 * e->time;
        double a1 = 0, a2 = 0, s2 = 0; int64_t n1 = 0, n2 = 0;
        for (int64_t i = 0; i < ns; ++i) {
          if (ts[i] >= t1 && ts[i] <= t2) { a1 += mom[26 * i]; ++n1; }
          else if (ts[i] > t2) { a2 += mom[26 * i]; s2 += mom[26 * i] * mom[26 * i]; ++n2; }
        }
        a1 /= (double)n1; a2 /= (double)n2;
        const double rel = sqrt((s2 / (double)n2 - a2 * a2) / (double)n2);
        if (a1 >= a2 && rel < 0.01) { first_int = ns - 1; n_int = 1; t_ss = e->time; real_at_ss = t.real; integrated = 0; }
      }
    }
    /* stop criteria (BMC.C:320-321) */
    if (!((double)n_int < need_points || (t_ss == LO_NON_DEF) || integrated / t_ss < need_ss_times)) break;
    if (max_intervals > 0 && n_sync >= max_intervals) break;
    /* electronDynamicsUntilSynchronization (BMC.C:617-688) */
    check_nu_trial(m, e, &nu_trial);
    const double t_sync = e->time + sync_factor / nu_trial;
    ++n_sync;
    memset(e->ndraw, 0, sizeof(uint32_t) * (size_t)n);
    int any = 1;
    while (any) {
      check_nu_trial(m, e, &nu_trial);
      any = parallel_pass(e, nu_trial, t_sync, (uint32_t)n_sync);
      serial_pass(e, nu_trial, 1, &t);
    }
    e->time = t_sync;
  }
  const double wall = omp_get_wtime() - wall0;
  memset(res, 0, sizeof(double) * 35);
  if (n_int >= 50) {
    double acc = 0;
    for (int64_t i = 0; i < n_int; ++i) acc += mom[26 * (first_int + i)];
    res[0] = acc / (double)n_int;
    res[1] = stat_error(mom + 26 * first_int, n_int, 26, 50);
    for (int c = 0; c < 3; ++c) {
      acc = 0; for (int64_t i = 0; i < n_int; ++i) acc += mom[26 * (first_int + i) + 5 + c];
      res[2 + c] = acc / (double)n_int; res[5 + c] = stat_error(mom + 26 * first_int + 5 + c, n_int, 26, 50);
      acc = 0; for (int64_t i = 0; i < n_int; ++i) acc += bulkv[3 * (first_int + i) + c];
      res[26 + c] = acc / (double)n_int;
    }
    for (int c = 0; c < 9; ++c) {
      acc = 0; for (int64_t i = 0; i < n_int; ++i) acc += mom[26 * (first_int + i) + 17 + c];
      res[8 + c] = acc / (double)n_int;
      acc = 0; for (int64_t i = 0; i < n_int; ++i) acc += bulkd[9 * (first_int + i) + c];
      res[17 + c] = acc / (double)n_int;
    }
  }
  res[29] = (double)t.real; res[30] = (double)t.null; res[31] = t_ss; res[32] = e->time; res[33] = (double)n_sync; res[34] = wall;
  free(ts); free(mom); free(bulkv); free(bulkd);
  lo_ensemble_destroy(e);
}
