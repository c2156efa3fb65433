// lk_stream.cuh -- K1, streaming-pool form: the production advance kernel for large ensembles.
//
// A CTA owns a contiguous range of the ensemble and keeps a POOL of electrons resident in shared memory.  Every round
//   (1) one block scan over the slot flags RETIRES the electrons that reached t_sync (written back in arrival order, so stores are
//       dense), REFILLS the freed slots from the CTA's input cursor (dense loads) and builds the two work lists in slot order;
//   (2) phase B runs the collisions of the electrons that passed the null test      (BMC.C:916-1031, 1054-1280);
//   (3) phase A runs one free flight + null-collision test for every active electron (BMC.C:650-667, 804-905, 1035-1053).
// Both phases run on compacted lists: warps are full, all warps of the CTA execute the same short code, and -- unlike a tile that
// is drained to empty (lk_tile.cuh: 8 of 12 resident warps wait at barriers, profiles/r1_v2_*) -- the pool stays full until the
// CTA's range is exhausted, so the Poisson tail is paid once per CTA, not once per tile.  Collisions wait in the pool until a
// whole CTA-iteration of them (256) is available, so phase B is balanced as well.
// Electrons therefore PERMUTE inside the CTA's range (in place: the write cursor never overtakes the read cursor); the `id`
// column travels with each electron and keys its counter-based draw stream, so the physics is bit-identical to the
// one-thread-per-electron kernel.  The schedule involves no atomics, hence it is deterministic.
#pragma once
#include "lk_tile.cuh"

namespace lk {

constexpr int POOL = 1024;
constexpr int STREAM_THREADS = 256;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
static_assert(POOL == 4 * STREAM_THREADS, "the scan reads the 4 flags of a thread as one 32-bit word");

enum : int { SC_X = 0, SC_Y, SC_Z, SC_VX, SC_VY, SC_VZ, SC_TCF, SC_NUE, SC_T, SC_AUX, SC_ID, SC_COLS };

__host__ __device__ inline size_t stream_smem_bytes(int P, int nEn_hist) {
  size_t b = static_cast<size_t>(SC_COLS) * POOL * 8;       // state columns (+ time, aux, id)
  b += static_cast<size_t>(STREAM_WARPS) * R_HEADER * 8;    // per-warp accumulators
  b += static_cast<size_t>(P) * 16;                         // gain, loss
  b += 16 * 8;                                              // scan scratch (64-bit warp totals)
  b += static_cast<size_t>(POOL) * 4;                       // draw counters
  b += static_cast<size_t>(P) * 4;                          // counts
  b += static_cast<size_t>(nEn_hist) * 4;                   // energy histogram
  b += static_cast<size_t>(POOL) * 2 * 2;                   // two lists
  b += POOL;                                                // flags
  return (b + 15) & ~static_cast<size_t>(15);
}

struct StateId { State s; unsigned long long* id; };

template <int FIELD, int GT, bool SAMPLE>
__global__ void __launch_bounds__(STREAM_THREADS, 2) k_advance_stream(const Model m, const StateId sid, const Lists L, const Pending pend, const AdvArgs a,
                                                                       const HistGrid h, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* col = reinterpret_cast<double*>(smem_raw);                                        // [SC_COLS][POOL]
  double (*s_acc)[R_HEADER] = reinterpret_cast<double (*)[R_HEADER]>(col + SC_COLS * POOL);
  double* s_gain = reinterpret_cast<double*>(s_acc) + STREAM_WARPS * R_HEADER;
  double* s_loss = s_gain + m.P;
  unsigned long long* s_scan = reinterpret_cast<unsigned long long*>(s_loss + m.P);         // [16]
  unsigned int* s_used = reinterpret_cast<unsigned int*>(s_scan + 16);                      // [POOL]
  unsigned int* s_cnt = s_used + POOL;                                                      // [P]
  unsigned int* s_eeh = s_cnt + m.P;
  const int n_hist = (SAMPLE && h.enabled) ? h.nEn : 0;
  unsigned short* listF = reinterpret_cast<unsigned short*>(s_eeh + n_hist);                // [POOL]
  unsigned short* listR = listF + POOL;                                                     // [POOL]
  unsigned char* flag = reinterpret_cast<unsigned char*>(listR + POOL);                     // [POOL]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < m.P; k += STREAM_THREADS) { s_gain[k] = 0; s_loss[k] = 0; s_cnt[k] = 0; }
  for (int b = tid; b < n_hist; b += STREAM_THREADS) s_eeh[b] = 0;
  for (int j = tid; j < STREAM_WARPS * R_HEADER; j += STREAM_THREADS) (&s_acc[0][0])[j] = 0;
  reinterpret_cast<unsigned int*>(flag)[tid] = 0u;   // all slots FL_EMPTY

  unsigned int n_null = 0, n_born = 0, n_att = 0, n_clamp = 0, n_nuex = 0;
  double gain_field = 0, max_end = 0, max_seen = 0;
  const uint32_t k0 = static_cast<uint32_t>(a.seed), k1 = static_cast<uint32_t>(a.seed >> 32);
  double* const gcol[8] = {sid.s.x, sid.s.y, sid.s.z, sid.s.vx, sid.s.vy, sid.s.vz, sid.s.tcf, sid.s.nue};
  double val[N_SAMPLE_SUMS];
  if (SAMPLE) {
#pragma unroll
    for (int j = 0; j < N_SAMPLE_SUMS; ++j) val[j] = 0;
  }

  // the CTA's range [lo, hi) of the ensemble; in/out cursors are CTA-uniform
  const long long chunk = (((a.n + gridDim.x - 1) / gridDim.x) + 31) & ~31ll;
  const long long lo = min(static_cast<long long>(blockIdx.x) * chunk, a.n), hi = min(lo + chunk, a.n);
  long long in_ptr = lo, out_ptr = lo;
  __syncthreads();

  for (;;) {
    // ================= (1) scan: retire, refill, build lists =================
    const unsigned int f4 = reinterpret_cast<const unsigned int*>(flag)[tid];
    unsigned int cFl = 0, cRe = 0, cRet = 0, cFree = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      cFl += (f == FL_FLIGHT); cRe += (f == FL_REAL); cRet += (f == FL_DONE || f == FL_DEAD); cFree += (f == FL_EMPTY || f == FL_DONE || f == FL_DEAD);
    }
    const unsigned long long mine = static_cast<unsigned long long>(cFl) | (static_cast<unsigned long long>(cRe) << 12) |
                                    (static_cast<unsigned long long>(cRet) << 24) | (static_cast<unsigned long long>(cFree) << 36);
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < STREAM_WARPS; ++w) { const unsigned long long v = s_scan[w]; total += v; if (w < warp) before += v; }
    const unsigned long long excl = before + incl - mine;
    const int nFl = static_cast<int>(total & 0xFFFu), nRe = static_cast<int>((total >> 12) & 0xFFFu), nRet = static_cast<int>((total >> 24) & 0xFFFu),
              nFree = static_cast<int>((total >> 36) & 0xFFFu);
    const int nRefill = static_cast<int>(min(static_cast<long long>(nFree), hi - in_ptr));
    // collisions are run in whole CTA-iterations; everything is flushed when the flights alone cannot keep the CTA busy
    const int nB = (nFl + nRefill < STREAM_THREADS) ? nRe : (nRe / STREAM_THREADS) * STREAM_THREADS;
    int eFl = static_cast<int>(excl & 0xFFFu), eRe = static_cast<int>((excl >> 12) & 0xFFFu), eRet = static_cast<int>((excl >> 24) & 0xFFFu),
        eFree = static_cast<int>((excl >> 36) & 0xFFFu);
    unsigned int new_f4 = f4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sl = tid * 4 + q;
      unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      if (f == FL_DONE || f == FL_DEAD) {                          // ---- retire: dense write-back in arrival order ----
        const long long pos = out_ptr + eRet; ++eRet;
#pragma unroll
        for (int c = 0; c < 8; ++c) __stcs(&gcol[c][pos], col[c * POOL + sl]);
        sid.id[pos] = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * POOL + sl]));
        if (f == FL_DEAD) {
          const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
          if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(pos); L.dead_flag[pos] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
        } else if (SAMPLE) {                                       // ensemble sums of BMC.C:1432-1444, accumulated as electrons retire
          const double x = col[SC_X * POOL + sl], y = col[SC_Y * POOL + sl], z = col[SC_Z * POOL + sl];
          const double vx = col[SC_VX * POOL + sl], vy = col[SC_VY * POOL + sl], vz = col[SC_VZ * POOL + sl];
          const double eps = kinetic_eV(vx, vy, vz);
          val[0] += eps; val[1] += x; val[2] += y; val[3] += z; val[4] += vx; val[5] += vy; val[6] += vz;
          val[7] += x * x; val[8] += x * y; val[9] += x * z; val[11] += y * y; val[12] += y * z; val[15] += z * z;
          val[16] += x * vx; val[17] += x * vy; val[18] += x * vz; val[19] += y * vx; val[20] += y * vy; val[21] += y * vz;
          val[22] += z * vx; val[23] += z * vy; val[24] += z * vz; val[25] += 1.0;
          if (h.enabled) sample_histograms(h, vx, vy, vz, eps, s_eeh);
        }
        f = FL_EMPTY;
      }
      if (f == FL_EMPTY) {                                         // ---- refill: dense loads from the input cursor ----
        if (eFree < nRefill) {
          const long long pos = in_ptr + eFree;
#pragma unroll
          for (int c = 0; c < 8; ++c) col[c * POOL + sl] = __ldcs(&gcol[c][pos]);
          col[SC_ID * POOL + sl] = __longlong_as_double(static_cast<long long>(sid.id[pos]));
          col[SC_T * POOL + sl] = a.t0; s_used[sl] = 0;
          f = FL_FLIGHT;
          listF[eFl + min(eFree, nRefill) + min(eRe, nB)] = static_cast<unsigned short>(sl);
        }
        ++eFree;
      } else if (f == FL_FLIGHT) {
        listF[eFl + min(eFree, nRefill) + min(eRe, nB)] = static_cast<unsigned short>(sl);
        ++eFl;
      } else if (f == FL_REAL) {
        if (eRe < nB) { listR[eRe] = static_cast<unsigned short>(sl); listF[eFl + min(eFree, nRefill) + eRe] = static_cast<unsigned short>(sl); }
        ++eRe;
      }
      new_f4 = (new_f4 & ~(0xFFu << (8 * q))) | (f << (8 * q));
    }
    reinterpret_cast<unsigned int*>(flag)[tid] = new_f4;
    in_ptr += nRefill; out_ptr += nRet;
    const int nF = nFl + nRefill + nB;
    __syncthreads();
    if (nF == 0) break;                                            // nothing in flight, nothing parked, nothing left to load

    // ================= (2) phase B: collisions on the compacted real list =================
    if (nB > 0) {
      for (int chunk_i = warp; chunk_i * 32 < nB; chunk_i += STREAM_WARPS) {
        const int k = chunk_i * 32 + lane;
        int chosen = NOT_ADVANCED;
        double dE = 0;
        if (k < nB) {
          const int sl = listR[k];
          Particle p;
          p.x = col[SC_X * POOL + sl]; p.y = col[SC_Y * POOL + sl]; p.z = col[SC_Z * POOL + sl];
          p.vx = col[SC_VX * POOL + sl]; p.vy = col[SC_VY * POOL + sl]; p.vz = col[SC_VZ * POOL + sl];
          p.nue = col[SC_NUE * POOL + sl]; p.t = col[SC_T * POOL + sl]; p.tcf = NON_DEF;
          p.eps = kinetic_eV(p.vx, p.vy, p.vz);
          PhiloxRng rng;
          const unsigned long long id = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * POOL + sl]));
          rng.k0 = k0; rng.k1 = k1; rng.c0 = static_cast<uint32_t>(id); rng.c1 = static_cast<uint32_t>(id >> 32); rng.c2 = a.interval;
          rng.used = s_used[sl]; rng.blk = 0xFFFFFFFFu;
          EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
          if (thermal_branch<GT>(m, p.eps)) chosen = thermal_collide<GT>(m, p, rng, o);
          else chosen = cold_collide<GT>(m, p, col[SC_AUX * POOL + sl], rng, o);
          n_clamp += o.table_clamped;
          unsigned char outcome = FL_FLIGHT;
          if (chosen >= 0) {
            dE = o.dE;
            const int type = __ldg(&m.type[chosen]);
            if (type == T_IONIZATION) {
              ++n_born;
              uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, rng.used, cc1, ck1);
              push_pending(pend, L.counters, o, p.t, rng.c0, cc1, ck1);
            } else if (type == T_ATTACHMENT) { ++n_att; outcome = FL_DEAD; }
          } else ++n_null;                                         // aborted picks count as null collisions (BMC.C:1137-1140)
          max_seen = fmax(max_seen, p.eps);
          col[SC_VX * POOL + sl] = p.vx; col[SC_VY * POOL + sl] = p.vy; col[SC_VZ * POOL + sl] = p.vz;
          col[SC_TCF * POOL + sl] = NON_DEF;                       // the next free time is drawn at the start of phase A (same stream position)
          s_used[sl] = rng.used;
          flag[sl] = outcome;
        }
        tally_collisions(chosen, dE, s_cnt, s_gain, s_loss, lane);
      }
      __syncthreads();
    }

    // ================= (3) phase A: flight + null test on the compacted flight list =================
    for (int chunk_i = warp; chunk_i * 32 < nF; chunk_i += STREAM_WARPS) {
      const int k = chunk_i * 32 + lane;
      if (k < nF) {
        const int sl = listF[k];
        if (flag[sl] == FL_FLIGHT) {                               // (an electron attached in phase B stays FL_DEAD and retires in the next scan)
          Particle p;
          p.x = col[SC_X * POOL + sl]; p.y = col[SC_Y * POOL + sl]; p.z = col[SC_Z * POOL + sl];
          p.vx = col[SC_VX * POOL + sl]; p.vy = col[SC_VY * POOL + sl]; p.vz = col[SC_VZ * POOL + sl];
          p.tcf = col[SC_TCF * POOL + sl]; p.nue = col[SC_NUE * POOL + sl]; p.t = col[SC_T * POOL + sl];
          p.eps = kinetic_eV(p.vx, p.vy, p.vz);
          PhiloxRng rng;
          const unsigned long long id = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * POOL + sl]));
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
              if (cold_null_test(m, p, rng, Rnu, o)) { outcome = FL_REAL; col[SC_AUX * POOL + sl] = Rnu; }
              else { outcome = FL_FLIGHT; p.tcf = NON_DEF; ++n_null; }
              n_clamp += o.table_clamped; n_nuex += o.nu_exceeded;
            }
          }
          max_seen = fmax(max_seen, p.eps);
          col[SC_X * POOL + sl] = p.x; col[SC_Y * POOL + sl] = p.y; col[SC_Z * POOL + sl] = p.z;
          col[SC_VX * POOL + sl] = p.vx; col[SC_VY * POOL + sl] = p.vy; col[SC_VZ * POOL + sl] = p.vz;
          col[SC_TCF * POOL + sl] = p.tcf; col[SC_NUE * POOL + sl] = p.nue; col[SC_T * POOL + sl] = p.t;
          s_used[sl] = rng.used;
          flag[sl] = outcome;
        }
      }
    }
    __syncthreads();
  }

  if (SAMPLE) {
#pragma unroll
    for (int j = 0; j < N_SAMPLE_SUMS; ++j) {
      if (j == 10 || j == 13 || j == 14) continue;
      const double sum = warp_sum(val[j]);
      if (lane == 0) s_acc[warp][R_SUM_EPS + j] = sum;
    }
  }
  {
    const double v1 = warp_sum(static_cast<double>(n_null)), v2 = warp_sum(static_cast<double>(n_born)), v3 = warp_sum(static_cast<double>(n_att)),
                 v4 = warp_sum(gain_field), v5 = warp_sum(static_cast<double>(n_clamp)), v6 = warp_sum(static_cast<double>(n_nuex)),
                 m0 = warp_max(max_end), m1 = warp_max(max_seen);
    if (lane == 0) {
      double* acc = s_acc[warp];
      acc[R_N_NULL] = v1; acc[R_N_BORN] = v2; acc[R_N_ATTACHED] = v3; acc[R_GAIN_FIELD] = v4;
      acc[R_N_TABLE_CLAMPED] = v5; acc[R_N_NU_EXCEEDED] = v6; acc[R_MAX_EPS] = m0; acc[R_MAX_EPS_SEEN] = m1;
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

}  // namespace lk
