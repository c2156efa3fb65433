// lk_tile.cuh -- K1, tile form: the production advance kernel for large ensembles.
//
// Why: with one electron per thread the number of events per electron per interval is Poisson(1) and a real collision costs
// several times a null one, so a warp spends most of its issue slots predicated off (measured on B200: 6.2 of 32 lanes active,
// 133 KB of divergent code thrashing the instruction cache; profiles/r1_v1_*).  Here a CTA keeps a TILE of electrons resident
// in shared memory (SoA, 84 B per electron) and advances it in ROUNDS; every round has two phases that each run on compacted
// lists, so warps are full and all warps of the CTA execute the same short code:
//   phase A  free flight to min(t + t_cf, t_sync) + null-collision test          (BMC.C:650-667, 804-905, 1035-1053)
//   phase B  process selection + collision dynamics for the electrons that passed (BMC.C:916-1031, 1054-1280)
// Compaction is a block scan over per-slot flags (lists are in slot order, so the schedule -- and with it every tally -- is
// deterministic).  Draws are counter-based per electron, so the result is bit-identical to the one-thread-per-electron kernel.
// Electrons ejected by ionization go to a global pending list and are advanced to t_sync by k_advance_births.
#pragma once
#include "lk_kernels.cuh"

namespace lk {

constexpr int TILE = 1024;
constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_SPT = TILE / TILE_THREADS;   // slots owned by one thread in the scan
static_assert(TILE_SPT == 4, "the scan reads the 4 flags of a thread as one 32-bit word");

enum : unsigned char { FL_EMPTY = 0, FL_FLIGHT = 1, FL_REAL = 2, FL_DONE = 3, FL_DEAD = 4 };
enum : int { COL_X = 0, COL_Y, COL_Z, COL_VX, COL_VY, COL_VZ, COL_TCF, COL_NUE, COL_T, COL_AUX, N_COLS };

__host__ __device__ inline size_t tile_smem_bytes(int P, int nEn_hist) {
  size_t b = static_cast<size_t>(N_COLS) * TILE * 8;        // state columns
  b += static_cast<size_t>(TILE_WARPS) * R_HEADER * 8;      // per-warp accumulators
  b += static_cast<size_t>(P) * 16;                         // gain, loss
  b += static_cast<size_t>(TILE) * 4;                       // draw counters
  b += static_cast<size_t>(P) * 4;                          // counts
  b += static_cast<size_t>(nEn_hist) * 4;                   // energy histogram
  b += 64;                                                  // scan scratch
  b += static_cast<size_t>(TILE) * 2 * 2;                   // two lists
  b += TILE;                                                // flags
  return (b + 15) & ~static_cast<size_t>(15);
}

struct Pending {   // electrons born inside the interval, still at their birth time
  double* col;     // [9][cap]: x y z vx vy vz t, (c0 | c1 << 32), k1
  unsigned int cap;
};

__device__ __forceinline__ void push_pending(const Pending& pend, unsigned int* counters, const EventOut& o, double t, uint32_t c0, uint32_t cc1, uint32_t ck1) {
  const unsigned int idx = atomicAdd(&counters[C_PENDING], 1u);
  if (idx < pend.cap) {
    double* b = pend.col + idx; const size_t c = pend.cap;
    b[0] = o.ejx; b[c] = o.ejy; b[2 * c] = o.ejz; b[3 * c] = o.ejvx; b[4 * c] = o.ejvy; b[5 * c] = o.ejvz; b[6 * c] = t;
    b[7 * c] = __longlong_as_double(static_cast<long long>(static_cast<unsigned long long>(c0) | (static_cast<unsigned long long>(cc1) << 32)));
    b[8 * c] = __longlong_as_double(static_cast<long long>(ck1));
  } else atomicExch(&counters[C_OVERFLOW], 1u);
}

// warp-converged tally of one batch of collisions (BMC.C:1308-1328): counts exactly, gain/loss aggregated per process
__device__ __forceinline__ void tally_collisions(int chosen, double dE, unsigned int* s_cnt, double* s_gain, double* s_loss, int lane) {
  const bool real = chosen >= 0;
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

#ifdef LK_BUILD_TILE_KERNEL   // superseded by k_advance_stream (lk_stream.cuh); kept as the measured stepping stone of profiles/r1_v2_*
template <int FIELD, int GT, bool SAMPLE>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_advance_tile(const Model m, const State s, const Lists L, const Pending pend, const AdvArgs a,
                                                                   const HistGrid h, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* col = reinterpret_cast<double*>(smem_raw);                         // [N_COLS][TILE]
  double (*s_acc)[R_HEADER] = reinterpret_cast<double (*)[R_HEADER]>(col + N_COLS * TILE);
  double* s_gain = reinterpret_cast<double*>(s_acc) + TILE_WARPS * R_HEADER;
  double* s_loss = s_gain + m.P;
  unsigned int* s_used = reinterpret_cast<unsigned int*>(s_loss + m.P);      // [TILE]
  unsigned int* s_cnt = s_used + TILE;                                       // [P]
  unsigned int* s_eeh = s_cnt + m.P;                                         // [nEn] when sampling histograms
  const int n_hist = (SAMPLE && h.enabled) ? h.nEn : 0;
  unsigned int* s_scan = s_eeh + n_hist;                                     // [16]: warp totals [8], nF, nR
  unsigned short* listF = reinterpret_cast<unsigned short*>(s_scan + 16);    // [TILE]
  unsigned short* listR = listF + TILE;                                      // [TILE]
  unsigned char* flag = reinterpret_cast<unsigned char*>(listR + TILE);      // [TILE]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < m.P; k += TILE_THREADS) { s_gain[k] = 0; s_loss[k] = 0; s_cnt[k] = 0; }
  for (int b = tid; b < n_hist; b += TILE_THREADS) s_eeh[b] = 0;
  for (int j = tid; j < TILE_WARPS * R_HEADER; j += TILE_THREADS) (&s_acc[0][0])[j] = 0;

  unsigned int n_null = 0, n_born = 0, n_att = 0, n_clamp = 0, n_nuex = 0;
  double gain_field = 0, max_end = 0, max_seen = 0;
  const uint32_t k0 = static_cast<uint32_t>(a.seed), k1 = static_cast<uint32_t>(a.seed >> 32);
  const double* gcol[8] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.tcf, s.nue};

  const long long n_tiles = (a.n + TILE - 1) / TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TILE;
    const int cnt = static_cast<int>(min(static_cast<long long>(TILE), a.n - base));
    __syncthreads();   // the previous tile's shared memory is free
    // ---- load: coalesced global -> shared, one column at a time ----
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double* __restrict__ src = gcol[c] + base;
#pragma unroll
      for (int j = 0; j < TILE_SPT; ++j) { const int sl = j * TILE_THREADS + tid; if (sl < cnt) col[c * TILE + sl] = __ldcs(&src[sl]); }
    }
#pragma unroll
    for (int j = 0; j < TILE_SPT; ++j) {
      const int sl = j * TILE_THREADS + tid;
      col[COL_T * TILE + sl] = a.t0; s_used[sl] = 0;
      flag[sl] = (sl < cnt) ? FL_FLIGHT : FL_EMPTY;
      listF[sl] = static_cast<unsigned short>(sl);
    }
    int nF = cnt;
    __syncthreads();

    for (;;) {
      // ================= phase A: flight + null test on the compacted flight list =================
      for (int chunk = warp; chunk * 32 < nF; chunk += TILE_WARPS) {
        const int k = chunk * 32 + lane;
        if (k < nF) {
          const int sl = listF[k];
          if (flag[sl] != FL_DEAD) {
            Particle p;
            p.x = col[COL_X * TILE + sl]; p.y = col[COL_Y * TILE + sl]; p.z = col[COL_Z * TILE + sl];
            p.vx = col[COL_VX * TILE + sl]; p.vy = col[COL_VY * TILE + sl]; p.vz = col[COL_VZ * TILE + sl];
            p.tcf = col[COL_TCF * TILE + sl]; p.nue = col[COL_NUE * TILE + sl]; p.t = col[COL_T * TILE + sl];
            p.eps = kinetic_eV(p.vx, p.vy, p.vz);
            PhiloxRng rng;
            const unsigned long long id = a.first_id + static_cast<unsigned long long>(base + sl);
            rng.k0 = k0; rng.k1 = k1; rng.c0 = static_cast<uint32_t>(id); rng.c1 = static_cast<uint32_t>(id >> 32); rng.c2 = a.interval;
            rng.used = s_used[sl]; rng.blk = 0xFFFFFFFFu;
            if (p.tcf == NON_DEF) { rng.align(); p.tcf = -log(rng.next()) / a.nu_trial; p.nue = a.nu_trial; }   // BMC.C:650-655
            unsigned char outcome;
            if (p.t + p.tcf > a.t_sync) {                            // partial flight, BMC.C:657-663
              const double dt = a.t_sync - p.t;
              gain_field += flight<FIELD>(m, p, dt);
              p.t = a.t_sync; p.tcf -= dt;
              outcome = FL_DONE;
              max_end = fmax(max_end, p.eps);
            } else {                                                 // BMC.C:666-667
              gain_field += flight<FIELD>(m, p, p.tcf);
              p.t += p.tcf;
              if (thermal_branch<GT>(m, p.eps)) outcome = FL_REAL;   // the thermal-target branch draws its own numbers in phase B
              else {
                EventOut o; o.table_clamped = 0; o.nu_exceeded = 0;
                double Rnu;
                if (cold_null_test(m, p, rng, Rnu, o)) { outcome = FL_REAL; col[COL_AUX * TILE + sl] = Rnu; }
                else { outcome = FL_FLIGHT; p.tcf = NON_DEF; ++n_null; }
                n_clamp += o.table_clamped; n_nuex += o.nu_exceeded;
              }
            }
            max_seen = fmax(max_seen, p.eps);
            col[COL_X * TILE + sl] = p.x; col[COL_Y * TILE + sl] = p.y; col[COL_Z * TILE + sl] = p.z;
            col[COL_VX * TILE + sl] = p.vx; col[COL_VY * TILE + sl] = p.vy; col[COL_VZ * TILE + sl] = p.vz;
            col[COL_TCF * TILE + sl] = p.tcf; col[COL_NUE * TILE + sl] = p.nue; col[COL_T * TILE + sl] = p.t;
            s_used[sl] = rng.used;
            flag[sl] = outcome;
          }
        }
      }
      __syncthreads();
      // ================= one block scan builds both lists in slot order =================
      int nR, nFnext;
      {
        const unsigned int f4 = reinterpret_cast<const unsigned int*>(flag)[tid];
        unsigned int cF = 0, cR = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const unsigned int f = (f4 >> (8 * q)) & 0xFFu; cF += (f == FL_FLIGHT || f == FL_REAL); cR += (f == FL_REAL); }
        const unsigned int mine = cF | (cR << 16);
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        unsigned int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TILE_WARPS; ++w) { const unsigned int v = s_scan[w]; total += v; if (w < warp) before += v; }
        const unsigned int excl = before + incl - mine;
        unsigned int pF = excl & 0xFFFFu, pR = excl >> 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned int f = (f4 >> (8 * q)) & 0xFFu;
          if (f == FL_FLIGHT || f == FL_REAL) listF[pF++] = static_cast<unsigned short>(tid * 4 + q);
          if (f == FL_REAL) listR[pR++] = static_cast<unsigned short>(tid * 4 + q);
        }
        nFnext = static_cast<int>(total & 0xFFFFu); nR = static_cast<int>(total >> 16);
      }
      __syncthreads();
      if (nFnext == 0) break;
      // ================= phase B: collisions on the compacted real list =================
      for (int chunk = warp; chunk * 32 < nR; chunk += TILE_WARPS) {
        const int k = chunk * 32 + lane;
        int chosen = NOT_ADVANCED;
        double dE = 0;
        if (k < nR) {
          const int sl = listR[k];
          Particle p;
          p.x = col[COL_X * TILE + sl]; p.y = col[COL_Y * TILE + sl]; p.z = col[COL_Z * TILE + sl];
          p.vx = col[COL_VX * TILE + sl]; p.vy = col[COL_VY * TILE + sl]; p.vz = col[COL_VZ * TILE + sl];
          p.nue = col[COL_NUE * TILE + sl]; p.t = col[COL_T * TILE + sl]; p.tcf = NON_DEF;
          p.eps = kinetic_eV(p.vx, p.vy, p.vz);
          PhiloxRng rng;
          const unsigned long long id = a.first_id + static_cast<unsigned long long>(base + sl);
          rng.k0 = k0; rng.k1 = k1; rng.c0 = static_cast<uint32_t>(id); rng.c1 = static_cast<uint32_t>(id >> 32); rng.c2 = a.interval;
          rng.used = s_used[sl]; rng.blk = 0xFFFFFFFFu;
          EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
          if (thermal_branch<GT>(m, p.eps)) chosen = thermal_collide<GT>(m, p, rng, o);
          else chosen = cold_collide<GT>(m, p, col[COL_AUX * TILE + sl], rng, o);
          n_clamp += o.table_clamped;
          unsigned char outcome = FL_FLIGHT;
          if (chosen >= 0) {
            dE = o.dE;
            const int type = __ldg(&m.type[chosen]);
            if (type == T_IONIZATION) {
              ++n_born;
              uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, rng.used, cc1, ck1);
              push_pending(pend, L.counters, o, p.t, rng.c0, cc1, ck1);
            } else if (type == T_ATTACHMENT) {
              ++n_att; outcome = FL_DEAD;
              const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
              if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(base + sl); L.dead_flag[base + sl] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
            }
          } else ++n_null;                                           // aborted picks count as null collisions (BMC.C:1137-1140)
          max_seen = fmax(max_seen, p.eps);
          col[COL_VX * TILE + sl] = p.vx; col[COL_VY * TILE + sl] = p.vy; col[COL_VZ * TILE + sl] = p.vz;
          col[COL_TCF * TILE + sl] = NON_DEF;                        // the next free time is drawn at the start of the next phase A (same stream position)
          s_used[sl] = rng.used;
          flag[sl] = outcome;
        }
        tally_collisions(chosen, dE, s_cnt, s_gain, s_loss, lane);
      }
      __syncthreads();
      nF = nFnext;
    }

    // ---- tile done: every live electron sits at t_sync ----
    if (SAMPLE) {
      double val[N_SAMPLE_SUMS];
#pragma unroll
      for (int j = 0; j < N_SAMPLE_SUMS; ++j) val[j] = 0;
#pragma unroll
      for (int j = 0; j < TILE_SPT; ++j) {
        const int sl = j * TILE_THREADS + tid;
        if (flag[sl] == FL_DONE) {
          const double x = col[COL_X * TILE + sl], y = col[COL_Y * TILE + sl], z = col[COL_Z * TILE + sl];
          const double vx = col[COL_VX * TILE + sl], vy = col[COL_VY * TILE + sl], vz = col[COL_VZ * TILE + sl];
          const double eps = kinetic_eV(vx, vy, vz);
          val[0] += eps; val[1] += x; val[2] += y; val[3] += z; val[4] += vx; val[5] += vy; val[6] += vz;
          val[7] += x * x; val[8] += x * y; val[9] += x * z; val[11] += y * y; val[12] += y * z; val[15] += z * z;
          val[16] += x * vx; val[17] += x * vy; val[18] += x * vz; val[19] += y * vx; val[20] += y * vy; val[21] += y * vz;
          val[22] += z * vx; val[23] += z * vy; val[24] += z * vz; val[25] += 1.0;
          if (h.enabled) sample_histograms(h, vx, vy, vz, eps, s_eeh);
        }
      }
#pragma unroll
      for (int j = 0; j < N_SAMPLE_SUMS; ++j) {
        if (j == 10 || j == 13 || j == 14) continue;
        const double sum = warp_sum(val[j]);
        if (lane == 0) s_acc[warp][R_SUM_EPS + j] += sum;
      }
    }
    // ---- store: coalesced shared -> global ----
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double* __restrict__ dst = const_cast<double*>(gcol[c]) + base;
#pragma unroll
      for (int j = 0; j < TILE_SPT; ++j) { const int sl = j * TILE_THREADS + tid; if (sl < cnt) __stcs(&dst[sl], col[c * TILE + sl]); }
    }
  }

  {
    const double v1 = warp_sum(static_cast<double>(n_null)), v2 = warp_sum(static_cast<double>(n_born)), v3 = warp_sum(static_cast<double>(n_att)),
                 v4 = warp_sum(gain_field), v5 = warp_sum(static_cast<double>(n_clamp)), v6 = warp_sum(static_cast<double>(n_nuex)),
                 m0 = warp_max(max_end), m1 = warp_max(max_seen);
    if (lane == 0) {
      double* acc = s_acc[warp];
      acc[R_N_NULL] += v1; acc[R_N_BORN] += v2; acc[R_N_ATTACHED] += v3; acc[R_GAIN_FIELD] += v4;
      acc[R_N_TABLE_CLAMPED] += v5; acc[R_N_NU_EXCEEDED] += v6; acc[R_MAX_EPS] = m0; acc[R_MAX_EPS_SEEN] = m1;
    }
  }
  __syncthreads();
  if (tid == 0) {   // real collisions = sum of the per-process counts
    double nr = 0;
    for (int k = 0; k < m.P; ++k) nr += static_cast<double>(s_cnt[k]);
    s_acc[0][R_N_REAL] = nr;
  }
  __syncthreads();
  write_partials(s_acc, s_cnt, s_gain, s_loss, m.P, partials);
  if (SAMPLE && h.enabled) flush_energy_histogram(h, s_eeh);
}

#endif  // LK_BUILD_TILE_KERNEL

// Electrons ejected inside the interval: advance each from its birth time to t_sync (one per thread, they are few), append the
// survivor to the birth list K2 consumes; their own offspring wait on the thread's stack (BMC.C:1346-1353 semantics).
template <int FIELD, int GT>
__global__ void __launch_bounds__(128) k_advance_births(const Model m, const Lists L, const Pending pend, const AdvArgs a, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_gain = reinterpret_cast<double*>(smem_raw);
  double* s_loss = s_gain + m.P;
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_loss + m.P);
  __shared__ double s_acc[ADV_WARPS][R_HEADER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < m.P; k += blockDim.x) { s_gain[k] = 0; s_loss[k] = 0; s_cnt[k] = 0; }
  for (int j = threadIdx.x; j < ADV_WARPS * R_HEADER; j += blockDim.x) (&s_acc[0][0])[j] = 0;
  __syncthreads();
  unsigned int n_real = 0, n_null = 0, n_born = 0, n_att = 0, n_clamp = 0, n_nuex = 0;
  double gain_field = 0, max_seen = 0;
  double stk[CHILD_STACK][8];
  const unsigned int n_pend = min(L.counters[C_PENDING], pend.cap);
  const unsigned int n_round = (n_pend + 31u) & ~31u;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    bool active = i < n_pend;
    Particle p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    PhiloxRng rng; rng.init(a.seed, 0, a.interval);
    if (active) {
      const double* b = pend.col + i; const size_t c = pend.cap;
      p.x = b[0]; p.y = b[c]; p.z = b[2 * c]; p.vx = b[3 * c]; p.vy = b[4 * c]; p.vz = b[5 * c]; p.t = b[6 * c];
      const unsigned long long w = static_cast<unsigned long long>(__double_as_longlong(b[7 * c]));
      rng.c0 = static_cast<uint32_t>(w); rng.c1 = static_cast<uint32_t>(w >> 32);
      rng.k1 = static_cast<uint32_t>(static_cast<unsigned long long>(__double_as_longlong(b[8 * c])));
      p.eps = kinetic_eV(p.vx, p.vy, p.vz); p.tcf = NON_DEF; p.nue = a.nu_trial;
    }
    int sp = 0;
    while (__any_sync(FULL, active)) {
      int chosen = NOT_ADVANCED;
      double dE = 0;
      if (active) {
        EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
        chosen = event<FIELD, GT>(m, p, a.nu_trial, a.t_sync, rng, o);
        gain_field += o.gain_field; max_seen = fmax(max_seen, p.eps);
        n_clamp += o.table_clamped; n_nuex += o.nu_exceeded;
        bool done = false;
        if (chosen == PARTIAL_FLIGHT) {
          const unsigned int idx = atomicAdd(&L.counters[C_BIRTHS], 1u);
          if (idx < L.birth_cap) {
            double* b = L.birth + idx; const size_t c = L.birth_cap;
            b[0] = p.x; b[c] = p.y; b[2 * c] = p.z; b[3 * c] = p.vx; b[4 * c] = p.vy; b[5 * c] = p.vz; b[6 * c] = p.tcf; b[7 * c] = p.nue;
          } else atomicExch(&L.counters[C_OVERFLOW], 1u);
          done = true;
        } else if (chosen >= 0) {
          dE = o.dE;
          const int type = __ldg(&m.type[chosen]);
          if (type == T_IONIZATION) {
            ++n_born;
            if (sp < CHILD_STACK) {
              uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, o.used_mark, cc1, ck1);
              stk[sp][0] = o.ejx; stk[sp][1] = o.ejy; stk[sp][2] = o.ejz; stk[sp][3] = o.ejvx; stk[sp][4] = o.ejvy; stk[sp][5] = o.ejvz; stk[sp][6] = p.t;
              stk[sp][7] = __longlong_as_double(static_cast<long long>(static_cast<unsigned long long>(cc1) | (static_cast<unsigned long long>(ck1) << 32)));
              ++sp;
            } else atomicExch(&L.counters[C_OVERFLOW], 1u);
          } else if (type == T_ATTACHMENT) { ++n_att; done = true; }
        }
        if (done) {
          if (sp > 0) {
            --sp;
            p.x = stk[sp][0]; p.y = stk[sp][1]; p.z = stk[sp][2]; p.vx = stk[sp][3]; p.vy = stk[sp][4]; p.vz = stk[sp][5]; p.t = stk[sp][6];
            const unsigned long long w = static_cast<unsigned long long>(__double_as_longlong(stk[sp][7]));
            rng.c1 = static_cast<uint32_t>(w); rng.k1 = static_cast<uint32_t>(w >> 32); rng.used = 0; rng.blk = 0xFFFFFFFFu;
            p.eps = kinetic_eV(p.vx, p.vy, p.vz); p.tcf = NON_DEF; p.nue = a.nu_trial;
          } else active = false;
        }
      }
      n_real += (chosen >= 0) ? 1u : 0u;
      n_null += (chosen == NULL_COLLISION) ? 1u : 0u;
      tally_collisions(chosen, dE, s_cnt, s_gain, s_loss, lane);
    }
  }
  {
    const double v0 = warp_sum(static_cast<double>(n_real)), v1 = warp_sum(static_cast<double>(n_null)), v2 = warp_sum(static_cast<double>(n_born)),
                 v3 = warp_sum(static_cast<double>(n_att)), v4 = warp_sum(gain_field), v5 = warp_sum(static_cast<double>(n_clamp)),
                 v6 = warp_sum(static_cast<double>(n_nuex)), m1 = warp_max(max_seen);
    if (lane == 0) {
      double* acc = s_acc[warp];
      acc[R_N_REAL] += v0; acc[R_N_NULL] += v1; acc[R_N_BORN] += v2; acc[R_N_ATTACHED] += v3; acc[R_GAIN_FIELD] += v4;
      acc[R_N_TABLE_CLAMPED] += v5; acc[R_N_NU_EXCEEDED] += v6; acc[R_MAX_EPS_SEEN] = m1;
    }
  }
  __syncthreads();
  // (blockDim = 128: warps 4..7 of s_acc stay zero)
  write_partials(s_acc, s_cnt, s_gain, s_loss, m.P, partials);
}

}  // namespace lk
