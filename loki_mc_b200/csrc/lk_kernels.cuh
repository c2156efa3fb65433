// lk_kernels.cuh -- sm_100a kernels of the electron Monte Carlo hot path.
//
//   k_init_ensemble   K0  initial Maxwellian                         (BMC.C:491-508)
//   k_advance         K1  advance every electron to t_sync           (BMC.C:617-688, 804-1280) + tallies of BMC.C:1293-1339
//                         (+ fused ensemble sums / histograms when no birth/death channel exists)
//   k_pc_*            K2  birth/death population control at t_sync   (BMC.C:1341-1407, batched)
//   k_sample          K3  ensemble sums + histograms as a separate pass (BMC.C:1410-1454, 1551-1571)
//   k_finalize            fixed-order reduction of the per-block partials into the result vector
//   k_step_injected       parity entry: one loop-body pass per thread with host-supplied draws
//
// One electron per thread; a warp stays converged at the end of every event round so that tallies use full-mask warp
// primitives.  Per-block partial results are combined in a fixed order (k_finalize) so that all integer-valued outputs and
// the ensemble sums are reproducible run to run.
#pragma once
#include "lk_physics.cuh"

namespace lk {

constexpr int ADV_THREADS = 256;
constexpr int ADV_WARPS = ADV_THREADS / 32;
constexpr int CHILD_STACK = 4;          // pending ejected electrons per thread inside one interval
constexpr unsigned FULL = 0xFFFFFFFFu;

// result vector layout: mirrors the LOKIB200_R_* enum of include/lokib200.h (static_asserted in lokib200.cu)
enum : int {
  R_N_REAL = 0, R_N_NULL = 1, R_N_BORN = 2, R_N_ATTACHED = 3, R_GAIN_FIELD = 4, R_GROWTH = 5, R_SUM_EPS = 6, R_SUM_R = 7, R_SUM_V = 10,
  R_SUM_RR = 13, R_SUM_RV = 22, R_N_SAMPLED = 31, R_N_TABLE_CLAMPED = 32, R_N_NU_EXCEEDED = 33, R_SUM_COUNT = 34, R_MAX_EPS = 34,
  R_MAX_EPS_SEEN = 35, R_OVERFLOW = 36, R_HEADER = 37
};
constexpr int N_SAMPLE_SUMS = 26;        // R_SUM_EPS .. R_N_SAMPLED

struct State { double *x, *y, *z, *vx, *vy, *vz, *tcf, *nue; };

enum : int { C_BIRTHS = 0, C_DEAD = 1, C_FREED = 2, C_PLACED = 3, C_OVERFLOW = 4, C_TERMS = 5, C_PENDING = 6, C_COUNT = 8 };

struct Lists {
  double* birth;              // [8][birth_cap]: x y z vx vy vz tcf nue of electrons born inside the interval, already at t_sync
  unsigned int* dead;         // [dead_cap] slots whose electron attached
  unsigned int* freed;        // [birth_cap] slots vacated by the surplus lottery
  unsigned int* claim;        // [n + birth_cap] lottery claims (0 = free)
  unsigned char* dead_flag;   // [n]
  double* growth_terms;       // [birth_cap + dead_cap] signed energyGrowth contributions
  unsigned int* counters;     // [C_COUNT]
  unsigned int birth_cap, dead_cap;
};

struct HistGrid {             // grids of BMC.C:1862-1883; steps computed on the host exactly as Eigen::LinSpaced does
  int enabled, cylindrical, nEn, nC, nR, nA, phase, pad;
  double e_step, c_first, c_step, r_step, a_first, a_step;
  unsigned long long *eeh, *eah, *evh, *eeh_phase;   // eeh_phase already points at the row of the current phase (or null)
  // shared-memory tiles of the two 2-D grids (k_sample): energy rows [0, ea_rows) x all cos cells, radial rows [0, ev_rows) x axial cells
  // [ev_a0, ev_a0 + ev_aw): where the bulk of a swarm lives.  0 rows = no tile (every count goes to global memory).
  int ea_rows, ev_rows, ev_a0, ev_aw;
  double inv_e, inv_c, inv_r, inv_a;   // reciprocals of the four steps (fast path of hist_bin)
};

// 16-bit counters, two per 32-bit word: a CTA flushes its tiles before any counter can reach 2^16 (HIST_PASS electrons per pass)
constexpr long long HIST_PASS = 65280;
__host__ __device__ inline size_t hist_tile_words(const HistGrid& h) {
  return (static_cast<size_t>(h.ea_rows) * h.nC + 1) / 2 + (static_cast<size_t>(h.ev_rows) * h.ev_aw + 1) / 2;
}

struct AdvArgs {
  long long n;
  unsigned long long first_id, seed;
  unsigned int interval, pad;
  double nu_trial, t0, t_sync;
};

// stream of an electron ejected by a parent with stream (c0, c1, k1) that had consumed `used` draws when the ionization was complete
__device__ __forceinline__ void child_stream(uint32_t c1, uint32_t k1, uint32_t used, uint32_t& cc1, uint32_t& ck1) {
  cc1 = c1 + ((used + 1u) << 8);
  ck1 = k1 + 0x632BE5ABu;
}

// ------------------------------------------------------------------ sampling helpers ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// ensemble sums of calculateMeanDataForSwarmParams (BMC.C:1432-1444) for one warp of electrons -> per-warp accumulators
__device__ __forceinline__ void sample_moments(bool alive, double x, double y, double z, double vx, double vy, double vz, double eps, double* warp_acc, int lane) {
  double val[N_SAMPLE_SUMS];
  const double a = alive ? 1.0 : 0.0;
  if (!alive) { x = y = z = vx = vy = vz = eps = 0.0; }
  val[0] = eps;
  val[1] = x; val[2] = y; val[3] = z;
  val[4] = vx; val[5] = vy; val[6] = vz;
  val[7] = x * x; val[8] = x * y; val[9] = x * z;
  val[10] = val[8]; val[11] = y * y; val[12] = y * z;
  val[13] = val[9]; val[14] = val[12]; val[15] = z * z;
  val[16] = x * vx; val[17] = x * vy; val[18] = x * vz;
  val[19] = y * vx; val[20] = y * vy; val[21] = y * vz;
  val[22] = z * vx; val[23] = z * vy; val[24] = z * vz;
  val[25] = a;
#pragma unroll
  for (int j = 0; j < N_SAMPLE_SUMS; ++j) {
    if (j == 10 || j == 13 || j == 14) continue;   // symmetric entries of r r^T are filled in by k_finalize
    const double s = warp_sum(val[j]);
    if (lane == 0) warp_acc[R_SUM_EPS + j] += s;
  }
}

// histogramCount / histogram2DCount (Math.C:61-81, :106-127) for one electron.  The EEDF row is privatised per CTA (s_eeh); the two 2-D grids
// are privatised where the swarm is dense (s_ea, s_ev: 16-bit counters packed in pairs, native 32-bit shared atomics) and counted straight
// in global memory elsewhere.  One 64-bit global atomic per electron and grid -- the first form of this pass -- serialised on the few
// hundred bins that hold most of the swarm: 0.38 ms for 1e7 electrons against 0.08 ms for reading them (profiles/r2_hist_*).
// bin index static_cast<int>((value - first) / step) as the reference computes it (Math.C:72, :118-119), without the FP64 division on the
// common path: (value - first) * (1 / step) differs from the quotient by a few ulp, so whenever it is not within 1e-7 of an integer both
// truncate to the same bin; otherwise (one value in ten million) the division decides.  Bit-exact, four divisions fewer per electron.
__device__ __forceinline__ int hist_bin(double value, double first, double step, double inv_step) {
  const double d = value - first, t = d * inv_step;
  if (fabs(t - rint(t)) < 1e-7) return static_cast<int>(d / step);
  return static_cast<int>(t);
}

__device__ __forceinline__ void sample_histograms(const HistGrid& h, double vx, double vy, double vz, double eps, unsigned int* s_eeh, unsigned int* s_ea, unsigned int* s_ev) {
  const int ie = hist_bin(eps, 0.0, h.e_step, h.inv_e);
  if (ie < h.nEn) atomicAdd(&s_eeh[ie], 1u);
  if (h.cylindrical) {
    const double vxy2 = vx * vx + vy * vy;
    const double cosang = vz / sqrt(vxy2 + vz * vz);
    const int ic = hist_bin(cosang, h.c_first, h.c_step, h.inv_c);
    if (ie < h.nEn && ie >= 0 && ic < h.nC && ic >= 0) {
      if (ie < h.ea_rows) { const int k = ie * h.nC + ic; atomicAdd(&s_ea[k >> 1], 1u << ((k & 1) * 16)); }
      else atomicAdd(&h.eah[static_cast<size_t>(ie) * h.nC + ic], 1ull);
    }
    const double vr = sqrt(vxy2);
    const int ir = hist_bin(vr, 0.0, h.r_step, h.inv_r), ia = hist_bin(vz, h.a_first, h.a_step, h.inv_a);
    if (ir < h.nR && ir >= 0 && ia < h.nA && ia >= 0) {
      const int ja = ia - h.ev_a0;
      if (ir < h.ev_rows && ja >= 0 && ja < h.ev_aw) { const int k = ir * h.ev_aw + ja; atomicAdd(&s_ev[k >> 1], 1u << ((k & 1) * 16)); }
      else atomicAdd(&h.evh[static_cast<size_t>(ir) * h.nA + ia], 1ull);
    }
  }
}

__device__ __forceinline__ void clear_histogram_tiles(const HistGrid& h, unsigned int* s_eeh, unsigned int* s_ea) {
  const int words = static_cast<int>(hist_tile_words(h));   // s_ev follows s_ea
  for (int b = threadIdx.x; b < h.nEn; b += blockDim.x) s_eeh[b] = 0;
  for (int b = threadIdx.x; b < words; b += blockDim.x) s_ea[b] = 0;
}

// CTA-private counts -> the global 64-bit accumulators (one atomic per non-empty bin and CTA)
__device__ __forceinline__ void flush_histograms(const HistGrid& h, const unsigned int* s_eeh, const unsigned int* s_ea, const unsigned int* s_ev) {
  for (int b = threadIdx.x; b < h.nEn; b += blockDim.x) {
    const unsigned int c = s_eeh[b];
    if (c) {
      atomicAdd(&h.eeh[b], static_cast<unsigned long long>(c));
      if (h.eeh_phase) atomicAdd(&h.eeh_phase[b], static_cast<unsigned long long>(c));
    }
  }
  if (!h.cylindrical) return;
  const int n_ea = h.ea_rows * h.nC, n_ev = h.ev_rows * h.ev_aw;
  for (int w = threadIdx.x; 2 * w < n_ea; w += blockDim.x) {
    const unsigned int c = s_ea[w];
    if (c & 0xFFFFu) atomicAdd(&h.eah[2 * w], static_cast<unsigned long long>(c & 0xFFFFu));
    if (c >> 16) atomicAdd(&h.eah[2 * w + 1], static_cast<unsigned long long>(c >> 16));
  }
  for (int w = threadIdx.x; 2 * w < n_ev; w += blockDim.x) {
    const unsigned int c = s_ev[w];
    const int k0 = 2 * w, k1 = 2 * w + 1;
    if (c & 0xFFFFu) atomicAdd(&h.evh[static_cast<size_t>(k0 / h.ev_aw) * h.nA + h.ev_a0 + k0 % h.ev_aw], static_cast<unsigned long long>(c & 0xFFFFu));
    if (c >> 16) atomicAdd(&h.evh[static_cast<size_t>(k1 / h.ev_aw) * h.nA + h.ev_a0 + k1 % h.ev_aw], static_cast<unsigned long long>(c >> 16));
  }
}

// per-block partials -> global, in a fixed layout [block][R_HEADER + 3P]
__device__ __forceinline__ void write_partials(double (*s_acc)[R_HEADER], const unsigned int* s_cnt, const double* s_gain, const double* s_loss,
                                               int P, double* partials) {
  const int len = R_HEADER + 3 * P;
  double* out = partials + static_cast<size_t>(blockIdx.x) * len;
  for (int j = threadIdx.x; j < R_HEADER; j += blockDim.x) {
    double v = s_acc[0][j];
    for (int w = 1; w < ADV_WARPS; ++w) v = (j >= R_SUM_COUNT) ? fmax(v, s_acc[w][j]) : v + s_acc[w][j];
    out[j] = v;
  }
  for (int k = threadIdx.x; k < P; k += blockDim.x) {
    out[R_HEADER + k] = s_cnt ? static_cast<double>(s_cnt[k]) : 0.0;
    out[R_HEADER + P + k] = s_gain ? s_gain[k] : 0.0;
    out[R_HEADER + 2 * P + k] = s_loss ? s_loss[k] : 0.0;
  }
}

// ------------------------------------------------------------------ K0 ------------------------------------------------------------------
// evaluateNonConstantVariables, ensemble part (BMC.C:491-508): r = 0, v ~ Maxwellian(sd) via unitNormalRand3 (Math.C:54-59), t_cf undefined
__global__ void k_identity_ids(unsigned long long* id, long long n, unsigned long long first_id) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    id[i] = first_id + static_cast<unsigned long long>(i);
}

__global__ void __launch_bounds__(256) k_init_ensemble(State s, long long n, unsigned long long first_id, unsigned long long seed, double sd,
                                                        unsigned long long* max_eps_bits) {
  double mx = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    PhiloxRng rng; rng.init(seed, first_id + static_cast<unsigned long long>(i), INIT_INTERVAL);
    const double r1 = rng.next(), r2 = rng.next(), r3 = rng.next(), r4 = rng.next();
    const double a1 = sqrt(-2.0 * log(r1)), a2 = 2.0 * PI * r2;
    const double vx = a1 * cos(a2) * sd, vy = a1 * sin(a2) * sd, vz = sqrt(-2.0 * log(r3)) * cos(2.0 * PI * r4) * sd;
    s.x[i] = 0; s.y[i] = 0; s.z[i] = 0; s.vx[i] = vx; s.vy[i] = vy; s.vz[i] = vz; s.tcf[i] = NON_DEF; s.nue[i] = 0;
    mx = fmax(mx, kinetic_eV(vx, vy, vz));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomicMax(max_eps_bits, static_cast<unsigned long long>(__double_as_longlong(mx)));   // eps >= 0: bit order == value order
}

// ------------------------------------------------------------------ K1 ------------------------------------------------------------------
template <int FIELD, int GT, bool SAMPLE>
__global__ void __launch_bounds__(ADV_THREADS, 2) k_advance(const Model m, const State s, const Lists L, const AdvArgs a, const HistGrid h,
                                                             double* __restrict__ partials) {
  extern __shared__ unsigned char smem_raw[];
  // dynamic smem: gain[P] | loss[P] | cnt[P] | eeh[nEn]
  double* s_gain = reinterpret_cast<double*>(smem_raw);
  double* s_loss = s_gain + m.P;
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_loss + m.P);
  unsigned int* s_eeh = s_cnt + m.P;                                    // [nEn], then the two 2-D tiles
  unsigned int* s_ea = s_eeh + h.nEn;
  unsigned int* s_ev = s_ea + (static_cast<size_t>(h.ea_rows) * h.nC + 1) / 2;
  __shared__ double s_acc[ADV_WARPS][R_HEADER];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < m.P; k += blockDim.x) { s_gain[k] = 0; s_loss[k] = 0; s_cnt[k] = 0; }
  if (SAMPLE && h.enabled) clear_histogram_tiles(h, s_eeh, s_ea);
  for (int j = threadIdx.x; j < ADV_WARPS * R_HEADER; j += blockDim.x) (&s_acc[0][0])[j] = 0;
  __syncthreads();

  unsigned int n_real = 0, n_null = 0, n_born = 0, n_att = 0, n_clamp = 0, n_nuex = 0;
  double gain_field = 0, max_end = 0, max_seen = 0;
  double stk[CHILD_STACK][8];

  const long long n_round = (a.n + 31) & ~31ll;
  for (long long i = blockIdx.x * static_cast<long long>(ADV_THREADS) + threadIdx.x; i < n_round; i += static_cast<long long>(gridDim.x) * ADV_THREADS) {
    bool active = i < a.n;
    Particle p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    PhiloxRng rng;
    rng.init(a.seed, 0, a.interval);
    if (active) {
      p.x = s.x[i]; p.y = s.y[i]; p.z = s.z[i]; p.vx = s.vx[i]; p.vy = s.vy[i]; p.vz = s.vz[i]; p.tcf = s.tcf[i]; p.nue = s.nue[i];
      p.t = a.t0; p.eps = kinetic_eV(p.vx, p.vy, p.vz);
      rng.init(a.seed, a.first_id + static_cast<unsigned long long>(i), a.interval);
    }
    bool is_child = false, alive = false;
    int sp = 0;
    Particle f = p;   // state of the slot's electron at t_sync (stored coalesced after the loop)

    while (__any_sync(FULL, active)) {
      int chosen = NOT_ADVANCED;
      double dE = 0;
      if (active) {
        EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
        chosen = event<FIELD, GT>(m, p, a.nu_trial, a.t_sync, rng, o);
        gain_field += o.gain_field;
        max_seen = fmax(max_seen, p.eps);
        n_clamp += o.table_clamped; n_nuex += o.nu_exceeded;
        bool done = false;
        if (chosen == PARTIAL_FLIGHT) {
          if (!is_child) { f = p; alive = true; max_end = fmax(max_end, p.eps); }
          else {
            const unsigned int idx = atomicAdd(&L.counters[C_BIRTHS], 1u);
            if (idx < L.birth_cap) {
              double* b = L.birth + idx; const size_t c = L.birth_cap;
              b[0] = p.x; b[c] = p.y; b[2 * c] = p.z; b[3 * c] = p.vx; b[4 * c] = p.vy; b[5 * c] = p.vz; b[6 * c] = p.tcf; b[7 * c] = p.nue;
            } else atomicExch(&L.counters[C_OVERFLOW], 1u);
          }
          done = true;
        } else if (chosen >= 0) {
          dE = o.dE;
          const int type = __ldg(&m.type[chosen]);
          if (type == T_IONIZATION) {                                // the ejected electron waits on the thread's stack (BMC.C:1346-1353 semantics:
            ++n_born;                                                 // parent's clock, undefined free time)
            if (sp < CHILD_STACK) {
              uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, o.used_mark, cc1, ck1);
              stk[sp][0] = o.ejx; stk[sp][1] = o.ejy; stk[sp][2] = o.ejz; stk[sp][3] = o.ejvx; stk[sp][4] = o.ejvy; stk[sp][5] = o.ejvz; stk[sp][6] = p.t;
              stk[sp][7] = __longlong_as_double(static_cast<long long>(static_cast<unsigned long long>(cc1) | (static_cast<unsigned long long>(ck1) << 32)));
              ++sp;
            } else atomicExch(&L.counters[C_OVERFLOW], 1u);
          } else if (type == T_ATTACHMENT) {
            ++n_att;
            if (!is_child) {
              const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
              if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(i); L.dead_flag[i] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
            }
            done = true;
          }
        }
        if (done) {
          if (sp > 0) {
            --sp;
            p.x = stk[sp][0]; p.y = stk[sp][1]; p.z = stk[sp][2]; p.vx = stk[sp][3]; p.vy = stk[sp][4]; p.vz = stk[sp][5]; p.t = stk[sp][6];
            const unsigned long long w = static_cast<unsigned long long>(__double_as_longlong(stk[sp][7]));
            rng.c1 = static_cast<uint32_t>(w); rng.k1 = static_cast<uint32_t>(w >> 32); rng.used = 0; rng.blk = 0xFFFFFFFFu;
            p.eps = kinetic_eV(p.vx, p.vy, p.vz); p.tcf = NON_DEF; p.nue = a.nu_trial;
            is_child = true;
          } else active = false;
        }
      }
      // ---- converged: tallies of nonParallelCollisionTasks (BMC.C:1303-1328) ----
      const bool real = chosen >= 0;
      n_real += real ? 1u : 0u;
      n_null += (chosen == NULL_COLLISION) ? 1u : 0u;
      const unsigned rm = __ballot_sync(FULL, real);
      if (real) {
        const unsigned peers = __match_any_sync(rm, chosen);
        const double g = (dE >= 0) ? dE : 0.0, l = (dE >= 0) ? 0.0 : dE;
        double gs = 0, ls = 0;
        for (unsigned rem = peers; rem; rem &= rem - 1) {
          const int src = __ffs(rem) - 1;
          gs += __shfl_sync(peers, g, src); ls += __shfl_sync(peers, l, src);
        }
        if (lane == __ffs(peers) - 1) {
          atomicAdd(&s_cnt[chosen], static_cast<unsigned int>(__popc(peers)));
          if (gs != 0) atomicAdd(&s_gain[chosen], gs);
          if (ls != 0) atomicAdd(&s_loss[chosen], ls);
        }
      }
    }

    if (alive) {   // coalesced write-back of the 8 state words
      s.x[i] = f.x; s.y[i] = f.y; s.z[i] = f.z; s.vx[i] = f.vx; s.vy[i] = f.vy; s.vz[i] = f.vz; s.tcf[i] = f.tcf; s.nue[i] = f.nue;
    }
    if (SAMPLE) {
      sample_moments(alive, f.x, f.y, f.z, f.vx, f.vy, f.vz, f.eps, s_acc[warp], lane);
      if (h.enabled && alive) sample_histograms(h, f.vx, f.vy, f.vz, f.eps, s_eeh, s_ea, s_ev);
    }
  }

  // thread tallies -> warp -> block partials
  {
    const double v0 = warp_sum(static_cast<double>(n_real)), v1 = warp_sum(static_cast<double>(n_null)), v2 = warp_sum(static_cast<double>(n_born)),
                 v3 = warp_sum(static_cast<double>(n_att)), v4 = warp_sum(gain_field), v5 = warp_sum(static_cast<double>(n_clamp)),
                 v6 = warp_sum(static_cast<double>(n_nuex)), m0 = warp_max(max_end), m1 = warp_max(max_seen);
    if (lane == 0) {
      double* acc = s_acc[warp];
      acc[R_N_REAL] += v0; acc[R_N_NULL] += v1; acc[R_N_BORN] += v2; acc[R_N_ATTACHED] += v3; acc[R_GAIN_FIELD] += v4;
      acc[R_N_TABLE_CLAMPED] += v5; acc[R_N_NU_EXCEEDED] += v6; acc[R_MAX_EPS] = m0; acc[R_MAX_EPS_SEEN] = m1;
    }
  }
  __syncthreads();
  write_partials(s_acc, s_cnt, s_gain, s_loss, m.P, partials);
  if (SAMPLE && h.enabled) flush_histograms(h, s_eeh, s_ea, s_ev);
}

// FP64 roofline denominator: 16 independent DFMA chains per thread, nothing but the FP64 pipe (bench.py: roofline.fp64)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
  double acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = seed + j * 1e-9 + threadIdx.x * 1e-12;
  const double m = seed, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __fma_rn(acc[j], m, c);
  }
  double t = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) t += acc[j];
  if (t == 123.456) out[blockIdx.x] = t;   // never true: keeps the chains alive
}

// ------------------------------------------------------------------ K3 ------------------------------------------------------------------
// ensemble sums + histograms as a separate pass (used when births/deaths can change the ensemble at t_sync)
__global__ void __launch_bounds__(ADV_THREADS) k_sample(const State s, long long n, const HistGrid h, int P, double* __restrict__ partials) {
  extern __shared__ unsigned char smem_raw[];
  unsigned int* s_eeh = reinterpret_cast<unsigned int*>(smem_raw);              // [nEn] EEDF row, then the tiles of the two 2-D grids
  unsigned int* s_ea = s_eeh + h.nEn;
  unsigned int* s_ev = s_ea + (static_cast<size_t>(h.ea_rows) * h.nC + 1) / 2;
  __shared__ double s_acc[ADV_WARPS][R_HEADER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (h.enabled) clear_histogram_tiles(h, s_eeh, s_ea);
  for (int j = threadIdx.x; j < ADV_WARPS * R_HEADER; j += blockDim.x) (&s_acc[0][0])[j] = 0;
  __syncthreads();
  double max_end = 0;
  double val[N_SAMPLE_SUMS];
#pragma unroll
  for (int j = 0; j < N_SAMPLE_SUMS; ++j) val[j] = 0;
  // two electrons per iteration: twelve 8-byte streaming loads in flight per thread (six were not enough to cover HBM latency at 16 warps/SM:
  // 4.5 TB/s in profiles/r1_v17_launches.csv).  The trip count is uniform over the CTA (the tiles are flushed at a barrier every HIST_PASS electrons).
  const long long stride = static_cast<long long>(gridDim.x) * ADV_THREADS;
  int since_flush = 0;
  for (long long base = blockIdx.x * static_cast<long long>(ADV_THREADS); base < n; base += 2 * stride) {
    const long long i0 = base + threadIdx.x, i2 = i0 + stride;
    const bool one = i0 < n, two = i2 < n;
    const long long i = one ? i0 : 0, j2 = two ? i2 : i;
    const double x = __ldcs(&s.x[i]), y = __ldcs(&s.y[i]), z = __ldcs(&s.z[i]), vx = __ldcs(&s.vx[i]), vy = __ldcs(&s.vy[i]), vz = __ldcs(&s.vz[i]);
    const double xb = __ldcs(&s.x[j2]), yb = __ldcs(&s.y[j2]), zb = __ldcs(&s.z[j2]), vxb = __ldcs(&s.vx[j2]), vyb = __ldcs(&s.vy[j2]), vzb = __ldcs(&s.vz[j2]);
    if (one) {
      const double eps = kinetic_eV(vx, vy, vz);
      max_end = fmax(max_end, eps);
      val[0] += eps; val[1] += x; val[2] += y; val[3] += z; val[4] += vx; val[5] += vy; val[6] += vz;
      val[7] += x * x; val[8] += x * y; val[9] += x * z; val[11] += y * y; val[12] += y * z; val[15] += z * z;
      val[16] += x * vx; val[17] += x * vy; val[18] += x * vz; val[19] += y * vx; val[20] += y * vy; val[21] += y * vz;
      val[22] += z * vx; val[23] += z * vy; val[24] += z * vz; val[25] += 1.0;
      if (h.enabled) sample_histograms(h, vx, vy, vz, eps, s_eeh, s_ea, s_ev);
    }
    if (two) {
      const double eps = kinetic_eV(vxb, vyb, vzb);
      max_end = fmax(max_end, eps);
      val[0] += eps; val[1] += xb; val[2] += yb; val[3] += zb; val[4] += vxb; val[5] += vyb; val[6] += vzb;
      val[7] += xb * xb; val[8] += xb * yb; val[9] += xb * zb; val[11] += yb * yb; val[12] += yb * zb; val[15] += zb * zb;
      val[16] += xb * vxb; val[17] += xb * vyb; val[18] += xb * vzb; val[19] += yb * vxb; val[20] += yb * vyb; val[21] += yb * vzb;
      val[22] += zb * vxb; val[23] += zb * vyb; val[24] += zb * vzb; val[25] += 1.0;
      if (h.enabled) sample_histograms(h, vxb, vyb, vzb, eps, s_eeh, s_ea, s_ev);
    }
    since_flush += 2 * ADV_THREADS;
    if (h.enabled && since_flush + 2 * ADV_THREADS > HIST_PASS) {   // before a 16-bit counter can wrap
      __syncthreads();
      flush_histograms(h, s_eeh, s_ea, s_ev);
      __syncthreads();
      clear_histogram_tiles(h, s_eeh, s_ea);
      __syncthreads();
      since_flush = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < N_SAMPLE_SUMS; ++j) {
    if (j == 10 || j == 13 || j == 14) continue;
    const double sum = warp_sum(val[j]);
    if (lane == 0) s_acc[warp][R_SUM_EPS + j] = sum;
  }
  const double m0 = warp_max(max_end);
  if (lane == 0) s_acc[warp][R_MAX_EPS] = m0;
  __syncthreads();
  write_partials(s_acc, nullptr, nullptr, nullptr, P, partials);
  if (h.enabled) flush_histograms(h, s_eeh, s_ea, s_ev);
}

// getTimeDependDistributions' counting pass (BMC.C:1551-1571) on its own: only the three velocity columns are read (24 B per electron), there
// are no running sums to keep in registers, so the kernel runs at full occupancy (the ensemble sums of the same sample come from k_sample
// without a grid, inside the advance).
constexpr int HIST_THREADS = 512;   // two CTAs per SM (their tiles take 79 KB each): 32 warps per SM
__global__ void __launch_bounds__(HIST_THREADS, 2) k_histogram(const State s, long long n, const HistGrid h) {
  extern __shared__ unsigned char smem_raw[];
  unsigned int* s_eeh = reinterpret_cast<unsigned int*>(smem_raw);              // [nEn] EEDF row, then the tiles of the two 2-D grids
  unsigned int* s_ea = s_eeh + h.nEn;
  unsigned int* s_ev = s_ea + (static_cast<size_t>(h.ea_rows) * h.nC + 1) / 2;
  clear_histogram_tiles(h, s_eeh, s_ea);
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * HIST_THREADS;
  int since_flush = 0;
  for (long long base = blockIdx.x * static_cast<long long>(HIST_THREADS); base < n; base += 4 * stride) {   // uniform trip count over the CTA
    double vx[4], vy[4], vz[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {   // twelve 8-byte streaming loads in flight per thread
      const long long i = base + threadIdx.x + u * stride;
      ok[u] = i < n;
      const long long j = ok[u] ? i : 0;
      vx[u] = __ldcs(&s.vx[j]); vy[u] = __ldcs(&s.vy[j]); vz[u] = __ldcs(&s.vz[j]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (ok[u]) sample_histograms(h, vx[u], vy[u], vz[u], kinetic_eV(vx[u], vy[u], vz[u]), s_eeh, s_ea, s_ev);
    since_flush += 4 * HIST_THREADS;
    if (since_flush + 4 * HIST_THREADS > HIST_PASS) {   // before a 16-bit counter can wrap
      __syncthreads();
      flush_histograms(h, s_eeh, s_ea, s_ev);
      __syncthreads();
      clear_histogram_tiles(h, s_eeh, s_ea);
      __syncthreads();
      since_flush = 0;
    }
  }
  __syncthreads();
  flush_histograms(h, s_eeh, s_ea, s_ev);
}

// ------------------------------------------------------------------ K2: population control ------------------------------------------------------------------
// Batched form of BMC.C:1341-1407 applied once at t_sync (every electron carries the same clock there):
//   (a) attached slots are refilled by ejected electrons, last born first (:1346-1359);
//   (b) attached slots left over get a copy of a uniformly drawn non-attached electron (:1361-1376);
//   (c) K surplus ejected electrons: a uniformly random K-subset of the (n + K) pool is removed, exactly the distribution the
//       reference's sequential draw-and-replace loop produces (:1380-1407); survivors move into the vacated slots.
__device__ __forceinline__ void copy_birth_to_slot(const State& s, const Lists& L, unsigned int slot, unsigned int b) {
  const double* src = L.birth + b; const size_t c = L.birth_cap;
  s.x[slot] = src[0]; s.y[slot] = src[c]; s.z[slot] = src[2 * c]; s.vx[slot] = src[3 * c]; s.vy[slot] = src[4 * c]; s.vz[slot] = src[5 * c];
  s.tcf[slot] = src[6 * c]; s.nue[slot] = src[7 * c];
}

// The five phases as device functions over a thread range [first, first + step, ...): the general path runs each as its own grid (a grid-wide
// dependency separates them), k_pc_small runs all of them in ONE CTA with barriers in between (small ensembles: the lists hold a few entries and
// five launches cost more than the work).
__device__ __forceinline__ void pc_fill(const State& s, const Lists& L, unsigned int first, unsigned int step) {
  const unsigned int nB = min(L.counters[C_BIRTHS], L.birth_cap), nD = min(L.counters[C_DEAD], L.dead_cap), nF = min(nB, nD);
  for (unsigned int j = first; j < nF; j += step) {
    const unsigned int slot = L.dead[j];
    copy_birth_to_slot(s, L, slot, nB - 1 - j);
    L.dead_flag[slot] = 0;
  }
}

__device__ __forceinline__ void pc_copy(const State& s, const Lists& L, long long n, unsigned long long first_id, unsigned long long seed, unsigned int interval,
                                        unsigned int first, unsigned int step) {
  const unsigned int nB = min(L.counters[C_BIRTHS], L.birth_cap), nD = min(L.counters[C_DEAD], L.dead_cap);
  for (unsigned int j = nB + first; j < nD; j += step) {
    const unsigned int slot = L.dead[j];
    PhiloxRng rng; rng.init(seed, first_id + slot, interval, 0); rng.c2 = interval; rng.c1 |= POPCTRL_INTERVAL_BIT;
    long long donor = 0;
    for (int it = 0; it < 4096; ++it) {                                   // BMC.C:1362-1365
      donor = static_cast<long long>(fmin(rng.next() * static_cast<double>(n), static_cast<double>(n - 1)));
      if (!L.dead_flag[donor]) break;
    }
    const double vx = s.vx[donor], vy = s.vy[donor], vz = s.vz[donor];
    L.growth_terms[atomicAdd(&L.counters[C_TERMS], 1u)] = kinetic_eV(vx, vy, vz);   // energyGrowth += eps_donor (:1367)
    s.x[slot] = s.x[donor]; s.y[slot] = s.y[donor]; s.z[slot] = s.z[donor]; s.vx[slot] = vx; s.vy[slot] = vy; s.vz[slot] = vz;
    s.tcf[slot] = s.tcf[donor]; s.nue[slot] = s.nue[donor];               // copies inherit t_cf and nu_e (:1373-1374)
  }
}

__device__ __forceinline__ void pc_lottery(const State& s, const Lists& L, long long n, unsigned long long first_id, unsigned long long seed, unsigned int interval,
                                           unsigned int first, unsigned int step) {
  const unsigned int nB = min(L.counters[C_BIRTHS], L.birth_cap), nD = min(L.counters[C_DEAD], L.dead_cap);
  if (nB <= nD) return;
  const unsigned int K = nB - nD;
  const double pool = static_cast<double>(n) + static_cast<double>(K);
  for (unsigned int b = first; b < K; b += step) {
    PhiloxRng rng; rng.init(seed, first_id + static_cast<unsigned long long>(n) + b, interval, 0); rng.c1 |= POPCTRL_INTERVAL_BIT;
    long long j = 0;
    for (int it = 0; it < 1 << 20; ++it) {                                // sampling without replacement by rejection
      j = static_cast<long long>(fmin(rng.next() * pool, pool - 1.0));
      if (atomicCAS(&L.claim[j], 0u, 1u) == 0u) break;
    }
    double eps;
    if (j < n) {                                                          // a member of the ensemble leaves (:1384-1395)
      L.freed[atomicAdd(&L.counters[C_FREED], 1u)] = static_cast<unsigned int>(j);
      eps = kinetic_eV(s.vx[j], s.vy[j], s.vz[j]);
    } else {                                                              // an ejected electron is dropped (:1397-1403)
      const size_t c = L.birth_cap; const double* src = L.birth + (j - n);
      eps = kinetic_eV(src[3 * c], src[4 * c], src[5 * c]);
    }
    L.growth_terms[atomicAdd(&L.counters[C_TERMS], 1u)] = -eps;
  }
}

__device__ __forceinline__ void pc_place(const State& s, const Lists& L, long long n, unsigned int first, unsigned int step) {
  const unsigned int nB = min(L.counters[C_BIRTHS], L.birth_cap), nD = min(L.counters[C_DEAD], L.dead_cap);
  if (nB <= nD) return;
  const unsigned int K = nB - nD;
  for (unsigned int b = first; b < K; b += step) {
    if (L.claim[n + b] == 0u) copy_birth_to_slot(s, L, L.freed[atomicAdd(&L.counters[C_PLACED], 1u)], b);
  }
}

// sums the growth terms in a fixed order into pc_result[0], clears claims / flags / counters for the next interval (one CTA)
__device__ __forceinline__ void pc_reset(const Lists& L, long long n, double* pc_result, double* red /* [blockDim.x] shared */) {
  const unsigned int nB = min(L.counters[C_BIRTHS], L.birth_cap), nD = min(L.counters[C_DEAD], L.dead_cap);
  const unsigned int K = (nB > nD) ? nB - nD : 0u, nFreed = L.counters[C_FREED], nTerms = L.counters[C_TERMS];
  for (unsigned int b = threadIdx.x; b < K; b += blockDim.x) L.claim[n + b] = 0u;
  for (unsigned int f = threadIdx.x; f < nFreed; f += blockDim.x) L.claim[L.freed[f]] = 0u;
  for (unsigned int j = threadIdx.x; j < nD; j += blockDim.x) L.dead_flag[L.dead[j]] = 0;
  double acc = 0;
  for (unsigned int t = threadIdx.x; t < nTerms; t += blockDim.x) acc += L.growth_terms[t];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (unsigned int t = 0; t < blockDim.x; ++t) tot += red[t];
    pc_result[0] = tot;
    pc_result[1] = static_cast<double>(L.counters[C_OVERFLOW]);
  }
  __syncthreads();
  if (threadIdx.x < C_COUNT) L.counters[threadIdx.x] = 0u;
}

__global__ void k_pc_fill(const State s, const Lists L) { pc_fill(s, L, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
__global__ void k_pc_copy(const State s, const Lists L, long long n, unsigned long long first_id, unsigned long long seed, unsigned int interval) {
  pc_copy(s, L, n, first_id, seed, interval, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void k_pc_lottery(const State s, const Lists L, long long n, unsigned long long first_id, unsigned long long seed, unsigned int interval) {
  pc_lottery(s, L, n, first_id, seed, interval, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void k_pc_place(const State s, const Lists L, long long n) { pc_place(s, L, n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
__global__ void k_pc_reset(const Lists L, long long n, double* pc_result) {
  __shared__ double red[256];
  pc_reset(L, n, pc_result, red);
}

// all five phases in one CTA: each phase reads what the previous one wrote (ensemble slots, claims, counters), so a CTA barrier - which also
// orders the CTA's global-memory accesses - stands where the general path has a kernel boundary
constexpr int PC_SMALL_THREADS = 256;
__global__ void __launch_bounds__(PC_SMALL_THREADS) k_pc_small(const State s, const Lists L, long long n, unsigned long long first_id, unsigned long long seed,
                                                               unsigned int interval, double* pc_result) {
  __shared__ double red[PC_SMALL_THREADS];
  pc_fill(s, L, threadIdx.x, PC_SMALL_THREADS);
  __syncthreads();
  pc_copy(s, L, n, first_id, seed, interval, threadIdx.x, PC_SMALL_THREADS);
  __syncthreads();
  pc_lottery(s, L, n, first_id, seed, interval, threadIdx.x, PC_SMALL_THREADS);
  __syncthreads();
  pc_place(s, L, n, threadIdx.x, PC_SMALL_THREADS);
  __syncthreads();
  pc_reset(L, n, pc_result, red);
}

// ------------------------------------------------------------------ finalize ------------------------------------------------------------------
// result[j] = fixed-order combination over blocks of the K1 partials (+ the partials of the births pass, + K3 partials for the
// sampled sums, + K2 growth)
__global__ void k_finalize(const double* __restrict__ adv_partials, int adv_blocks, const double* __restrict__ birth_partials, int birth_blocks,
                           const double* __restrict__ smp_partials, int smp_blocks, const double* __restrict__ pc_result, int P,
                           double* __restrict__ result) {
  const int len = R_HEADER + 3 * P;
  const int j = blockIdx.x;   // one warp per output entry: lanes stride over the blocks, then a fixed-shape shuffle tree
  if (j >= len) return;
  const int lane = threadIdx.x;
  int src = j;   // symmetric entries of sum r r^T are accumulated once (xy, xz, yz) and mirrored here
  if (j == R_SUM_RR + 3) src = R_SUM_RR + 1; else if (j == R_SUM_RR + 6) src = R_SUM_RR + 2; else if (j == R_SUM_RR + 7) src = R_SUM_RR + 5;
  const bool is_max = (j >= R_SUM_COUNT && j < R_HEADER);
  const bool from_sample = smp_partials && ((j >= R_SUM_EPS && j <= R_N_SAMPLED) || j == R_MAX_EPS);
  double v = 0;
  if (from_sample) {
    for (int b = lane; b < smp_blocks; b += 32) { const double t = smp_partials[static_cast<size_t>(b) * len + src]; v = is_max ? fmax(v, t) : v + t; }
  } else {
    for (int b = lane; b < adv_blocks; b += 32) { const double t = adv_partials[static_cast<size_t>(b) * len + src]; v = is_max ? fmax(v, t) : v + t; }
    if (birth_partials)
      for (int b = lane; b < birth_blocks; b += 32) { const double t = birth_partials[static_cast<size_t>(b) * len + src]; v = is_max ? fmax(v, t) : v + t; }
  }
  v = is_max ? warp_max(v) : warp_sum(v);
  if (j == R_GROWTH && pc_result) v = pc_result[0];
  if (j == R_OVERFLOW) v = pc_result ? pc_result[1] : 0.0;   // list overflow of this interval (k_pc_reset): travels with the vector, MAX-combined
  if (lane == 0) result[j] = v;
}

// ------------------------------------------------------------------ exchange over peer memory ------------------------------------------------------------------
// The one exchange of the path: the result vectors of the shards of a job are combined every sampling interval.  The payload is
// ~2.6 KB, so the cost is latency, not bandwidth: one CTA per GPU pushes its vector straight into a mailbox slot in EVERY peer's memory
// (NVLink stores), publishes a flag, waits for the flags of the others in its own mailbox and adds the n slots in rank order.  No ring, no
// second kernel, no host: one launch per interval and GPU, and because every rank adds in the same order all ranks hold the same bits.
// Mailbox of one rank: slot[2][n][stride] doubles, then flag[2][n] u64 (two parities: a rank can be at most one exchange ahead of another,
// because it cannot leave exchange k + 1 before every peer has entered it, i.e. has finished reading the slots of exchange k).
constexpr int EXCHANGE_MAX_RANKS = 16;
constexpr int EXCHANGE_THREADS = 256;
struct Mailboxes { double* box[EXCHANGE_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(EXCHANGE_THREADS) k_exchange(double* __restrict__ v, int len, int stride, const Mailboxes m, int me, int n, unsigned long long epoch,
                                                               long long timeout_cycles) {
  const int par = static_cast<int>(epoch & 1ull), tid = threadIdx.x;
  const size_t slot_me = (static_cast<size_t>(par) * n + me) * stride, flags_at = 2ull * n * stride;
  // 1. my vector into slot [par][me] of every rank, my own included
  for (int r = 0; r < n; ++r) {
    double* dst = m.box[r] + slot_me;
    for (int j = tid; j < len; j += EXCHANGE_THREADS) dst[j] = v[j];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish: flag [par][me] of every rank <- epoch
  if (tid < n) st_release_sys(reinterpret_cast<unsigned long long*>(m.box[tid] + flags_at) + par * n + me, epoch);
  // 3. wait for every rank's flag in my mailbox
  __shared__ int s_timeout;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < n) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(m.box[me] + flags_at) + par * n + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch) {
      if (clock64() - t0 > timeout_cycles) { s_timeout = 1; break; }   // a peer that never arrives must not hang the GPU
      __nanosleep(100);
    }
  }
  __syncthreads();
  // 4. the same sum on every rank: slots in rank order; [R_SUM_COUNT, R_HEADER) by max
  const double* mine = m.box[me] + static_cast<size_t>(par) * n * stride;
  for (int j = tid; j < len; j += EXCHANGE_THREADS) {
    double acc = __ldcg(mine + j);
    const bool is_max = (j >= R_SUM_COUNT && j < R_HEADER);
    for (int r = 1; r < n; ++r) { const double t = __ldcg(mine + static_cast<size_t>(r) * stride + j); acc = is_max ? fmax(acc, t) : acc + t; }
    v[j] = acc;
  }
  __syncthreads();
  if (tid == 0 && s_timeout) v[R_OVERFLOW] = 2.0;   // read by the host as an error (lokib200_read_result)
}

// ------------------------------------------------------------------ parity entry ------------------------------------------------------------------
struct ElectronIO { double r[3], v[3], energy, t, t_cf, nu_e; };
struct EventIO { int chosen, draws_used; double dE, dE_rel, gain_field, ej_r[3], ej_v[3], ej_energy; };

template <int FIELD, int GT>
__global__ void k_step_injected(const Model m, int n, const ElectronIO* __restrict__ in, double nu_trial, const double* __restrict__ t_sync,
                                const double* __restrict__ draws, int n_draws, ElectronIO* __restrict__ out, EventIO* __restrict__ ev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Particle p;
  p.x = in[i].r[0]; p.y = in[i].r[1]; p.z = in[i].r[2]; p.vx = in[i].v[0]; p.vy = in[i].v[1]; p.vz = in[i].v[2];
  p.eps = in[i].energy; p.t = in[i].t; p.tcf = in[i].t_cf; p.nue = in[i].nu_e;
  InjectedRng rng{draws + static_cast<size_t>(i) * n_draws, n_draws, 0};
  EventOut o; o.dE = 0; o.dE_rel = 0; o.gain_field = 0; o.ejx = o.ejy = o.ejz = o.ejvx = o.ejvy = o.ejvz = o.ejeps = 0; o.table_clamped = 0; o.nu_exceeded = 0;
  const int chosen = event<FIELD, GT>(m, p, nu_trial, t_sync[i], rng, o);
  out[i].r[0] = p.x; out[i].r[1] = p.y; out[i].r[2] = p.z; out[i].v[0] = p.vx; out[i].v[1] = p.vy; out[i].v[2] = p.vz;
  out[i].energy = p.eps; out[i].t = p.t; out[i].t_cf = p.tcf; out[i].nu_e = p.nue;
  ev[i].chosen = chosen; ev[i].draws_used = rng.used; ev[i].dE = o.dE; ev[i].dE_rel = o.dE_rel; ev[i].gain_field = o.gain_field;
  ev[i].ej_r[0] = o.ejx; ev[i].ej_r[1] = o.ejy; ev[i].ej_r[2] = o.ejz; ev[i].ej_v[0] = o.ejvx; ev[i].ej_v[1] = o.ejvy; ev[i].ej_v[2] = o.ejvz;
  ev[i].ej_energy = o.ejeps;
}

}  // namespace lk
