// lk_stream.cuh -- K1, streaming-pool form: the production advance kernel for large ensembles.
//
// (Design history, all measured on B200 and kept under profiles/: one electron per thread -> 6 of 32 lanes active, I-cache thrash;
//  CTA tile drained to empty -> 8 of 12 warps waiting at barriers; this CTA-wide pool with refill; warp-private pools -> I-cache thrash.)
//
// A CTA owns a contiguous range of the ensemble and keeps a POOL of electrons resident in shared memory.  Every round
//   (1) one block scan over the slot flags RETIRES the electrons that reached t_sync (written back in arrival order, so stores are
//       dense), REFILLS the freed slots from the CTA's input cursor (dense loads) and builds the two work lists in slot order;
//   (2) phase B runs the collisions of the electrons that passed the null test      (BMC.C:916-1031, 1054-1280);
//   (3) phase A runs one free flight + null-collision test for every active electron (BMC.C:650-667, 804-905, 1035-1053).
// Both phases run on compacted lists, so warps are full, and all warps of the CTA execute the same code at the same time, which is
// what keeps the instruction cache effective (a variant with warp-private pools and no CTA barriers was measured 1.7x slower:
// 16 desynchronised warps thrash the I-cache, stall_no_instruction 7.0 per issue).  The pool stays full until the CTA's range is
// exhausted, so the Poisson tail is paid once per CTA, not once per tile.  Collisions wait in the pool until a whole CTA-iteration
// of them (256) is available.
// Electrons therefore PERMUTE inside the CTA's range (in place: the write cursor never overtakes the read cursor); the `id`
// column travels with each electron and keys its counter-based draw stream, so the physics is bit-identical to the
// one-thread-per-electron kernel.  The schedule involves no atomics, hence it is deterministic.
#pragma once
#include "lk_tile.cuh"

namespace lk {

constexpr int POOL = 1024;                       // electrons resident per CTA
constexpr int STREAM_THREADS = 256;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
static_assert(POOL == 4 * STREAM_THREADS, "the scan reads the 4 flags of a thread as one 32-bit word");

enum : unsigned char { FL_REALT = 5 };   // passed the flight, waits for the thermal-target collision branch (FL_REAL = cold-gas branch)
// SC_TCF doubles as the hand-over slot for nu_e * U of an electron that waits for its collision (it has no free time then)
enum : int { SC_X = 0, SC_Y, SC_Z, SC_VX, SC_VY, SC_VZ, SC_TCF, SC_NUE, SC_T, SC_ID, SC_COLS };

// energy tallies per process are kept in 2^-36 eV fixed point: integer adds commute, so the sums are bitwise reproducible and
// a warp can aggregate them with REDUX (resolution 1.5e-11 eV per event; the reference uses them only for the power balance)
constexpr double TALLY_SCALE = 68719476736.0;

__host__ __device__ inline size_t stream_smem_bytes(int P, int nEn_hist) {
  size_t b = static_cast<size_t>(SC_COLS) * POOL * 8;       // state columns (+ time, id)
  b += static_cast<size_t>(STREAM_WARPS) * R_HEADER * 8;    // per-warp accumulators
  b += static_cast<size_t>(P) * 16;                         // gain, loss (fixed point)
  b += 16 * 8;                                              // scan scratch (64-bit warp totals)
  b += static_cast<size_t>(POOL) * 4;                       // draw counters
  b += static_cast<size_t>(P) * 4;                          // counts
  b += static_cast<size_t>(nEn_hist) * 4;                   // energy histogram
  b += static_cast<size_t>(POOL) * 2 * 5;                   // five lists
  b += POOL;                                                // flags
  return (b + 15) & ~static_cast<size_t>(15);
}

struct StateId { State s; unsigned long long* id; };

// 8-byte asynchronous global -> shared copy (LDGSTS): the refill of a freed slot overlaps with the collision phase
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned int d = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// warp-converged tally of one batch of collisions (BMC.C:1308-1328).  Lanes that chose the same process are found with MATCH,
// their fixed-point energy changes are summed with REDUX in three 21-bit limbs, and one lane per process does the shared atomics.
__device__ __forceinline__ void tally_collisions_fx(int chosen, double dE, unsigned int* s_cnt, unsigned long long* s_gain, unsigned long long* s_loss, int lane) {
  const bool real = chosen >= 0;
  const unsigned rm = __ballot_sync(FULL, real);
  if (real) {
    const unsigned peers = __match_any_sync(rm, chosen);
    const long long q = __double2ll_rn(dE * TALLY_SCALE);
    const unsigned long long g = (q > 0) ? static_cast<unsigned long long>(q) : 0ull, l = (q < 0) ? static_cast<unsigned long long>(-q) : 0ull;
    const unsigned g0 = __reduce_add_sync(peers, static_cast<unsigned>(g & 0x1FFFFFu)), g1 = __reduce_add_sync(peers, static_cast<unsigned>((g >> 21) & 0x1FFFFFu)),
                   g2 = __reduce_add_sync(peers, static_cast<unsigned>(g >> 42));
    const unsigned l0 = __reduce_add_sync(peers, static_cast<unsigned>(l & 0x1FFFFFu)), l1 = __reduce_add_sync(peers, static_cast<unsigned>((l >> 21) & 0x1FFFFFu)),
                   l2 = __reduce_add_sync(peers, static_cast<unsigned>(l >> 42));
    if (lane == __ffs(peers) - 1) {
      atomicAdd(&s_cnt[chosen], static_cast<unsigned int>(__popc(peers)));
      const unsigned long long gs = static_cast<unsigned long long>(g0) + (static_cast<unsigned long long>(g1) << 21) + (static_cast<unsigned long long>(g2) << 42);
      const unsigned long long ls = static_cast<unsigned long long>(l0) + (static_cast<unsigned long long>(l1) << 21) + (static_cast<unsigned long long>(l2) << 42);
      if (gs) atomicAdd(&s_gain[chosen], gs);
      if (ls) atomicAdd(&s_loss[chosen], ls);
    }
  }
}

template <int FIELD, int GT, bool SAMPLE>
__global__ void __launch_bounds__(STREAM_THREADS, 2) k_advance_stream(const Model m, const StateId sid, const Lists L, const Pending pend, const AdvArgs a,
                                                                       const HistGrid h, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* col = reinterpret_cast<double*>(smem_raw);                                        // [SC_COLS][POOL]
  double (*s_acc)[R_HEADER] = reinterpret_cast<double (*)[R_HEADER]>(col + SC_COLS * POOL);
  unsigned long long* s_gain = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(s_acc) + STREAM_WARPS * R_HEADER);
  unsigned long long* s_loss = s_gain + m.P;
  unsigned long long* s_scan = s_loss + m.P;                                                // [16]
  unsigned int* s_used = reinterpret_cast<unsigned int*>(s_scan + 16);                      // [POOL]
  unsigned int* s_cnt = s_used + POOL;                                                      // [P]
  unsigned short* listF = reinterpret_cast<unsigned short*>(s_cnt + m.P);                   // [POOL] flights of this round
  unsigned short* listR = listF + POOL;                                                     // [POOL] collisions of this round: cold first, thermal after
  unsigned short* listO = listR + POOL;                                                     // [POOL] slots retiring this round (output order; bit 15: attached)
  unsigned short* listK = listO + POOL;                                                     // [POOL] input rank of a retiring slot refilled at once (0xFFFF: stays empty)
  unsigned short* listI = listK + POOL;                                                     // [POOL] refilled slots that were already empty (0xFFFF: see listK)
  unsigned char* flag = reinterpret_cast<unsigned char*>(listI + POOL);                     // [POOL]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < m.P; k += STREAM_THREADS) { s_gain[k] = 0; s_loss[k] = 0; s_cnt[k] = 0; }
  for (int j = tid; j < STREAM_WARPS * R_HEADER; j += STREAM_THREADS) (&s_acc[0][0])[j] = 0;
  reinterpret_cast<unsigned int*>(flag)[tid] = 0u;   // all slots FL_EMPTY

  unsigned int n_null = 0, n_born = 0, n_att = 0, n_clamp = 0, n_nuex = 0;
  double gain_field = 0, max_end = 0, max_seen = 0;
  const uint32_t k0 = static_cast<uint32_t>(a.seed), k1 = static_cast<uint32_t>(a.seed >> 32);
  double* const gcol[8] = {sid.s.x, sid.s.y, sid.s.z, sid.s.vx, sid.s.vy, sid.s.vz, sid.s.tcf, sid.s.nue};

  // the CTA's range [lo, hi) of the ensemble; in/out cursors are CTA-uniform
  const long long chunk = (((a.n + gridDim.x - 1) / gridDim.x) + 31) & ~31ll;
  const long long lo = min(static_cast<long long>(blockIdx.x) * chunk, a.n), hi = min(lo + chunk, a.n);
  long long in_ptr = lo, out_ptr = lo;
  __syncthreads();

  for (;;) {
    // ================= (1) block scan over the slot flags: build the lists of this round in slot order =================
    const unsigned int f4 = reinterpret_cast<const unsigned int*>(flag)[tid];
    unsigned int cFl = 0, cRc = 0, cRt = 0, cRet = 0, cFree = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      cFl += (f == FL_FLIGHT); cRc += (f == FL_REAL); cRt += (f == FL_REALT); cRet += (f == FL_DONE || f == FL_DEAD);
      cFree += (f == FL_EMPTY || f == FL_DONE || f == FL_DEAD);
    }
    const unsigned long long mine = static_cast<unsigned long long>(cFl) | (static_cast<unsigned long long>(cRc) << 12) | (static_cast<unsigned long long>(cRt) << 24) |
                                    (static_cast<unsigned long long>(cRet) << 36) | (static_cast<unsigned long long>(cFree) << 48);
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < STREAM_WARPS; ++w) { const unsigned long long v = s_scan[w]; total += v; if (w < warp) before += v; }
    const unsigned long long excl = before + incl - mine;
    const int nFl = static_cast<int>(total & 0xFFFu), nRc = static_cast<int>((total >> 12) & 0xFFFu), nRt = static_cast<int>((total >> 24) & 0xFFFu),
              nRet = static_cast<int>((total >> 36) & 0xFFFu), nFree = static_cast<int>((total >> 48) & 0xFFFu);
    const int nRefill = static_cast<int>(min(static_cast<long long>(nFree), hi - in_ptr));
    // collisions are run in whole CTA-iterations; everything parked is flushed when the flights alone cannot keep the CTA busy
    const bool flush = (nFl + nRefill < STREAM_THREADS);
    const int nBc = flush ? nRc : (nRc / STREAM_THREADS) * STREAM_THREADS, nBt = flush ? nRt : (nRt / STREAM_THREADS) * STREAM_THREADS;
    int eFl = static_cast<int>(excl & 0xFFFu), eRc = static_cast<int>((excl >> 12) & 0xFFFu), eRt = static_cast<int>((excl >> 24) & 0xFFFu),
        eRet = static_cast<int>((excl >> 36) & 0xFFFu), eFree = static_cast<int>((excl >> 48) & 0xFFFu);
    unsigned int new_f4 = f4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sl = tid * 4 + q;
      unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      const bool retiring = (f == FL_DONE || f == FL_DEAD);
      if (retiring) {   // the thread that writes a slot back also issues its refill (program order, no barrier between the two)
        listO[eRet] = static_cast<unsigned short>(sl | (f == FL_DEAD ? 0x8000 : 0));
        listK[eRet] = (eFree < nRefill) ? static_cast<unsigned short>(eFree) : static_cast<unsigned short>(0xFFFF);
        ++eRet; f = FL_EMPTY;
      }
      const int posF = eFl + min(eFree, nRefill) + min(eRc, nBc) + min(eRt, nBt);
      if (f == FL_EMPTY) {
        if (eFree < nRefill) { listI[eFree] = retiring ? static_cast<unsigned short>(0xFFFF) : static_cast<unsigned short>(sl); listF[posF] = static_cast<unsigned short>(sl); f = FL_FLIGHT; }
        ++eFree;
      } else if (f == FL_FLIGHT) { listF[posF] = static_cast<unsigned short>(sl); ++eFl; }
      else if (f == FL_REAL) { if (eRc < nBc) { listR[eRc] = static_cast<unsigned short>(sl); listF[posF] = static_cast<unsigned short>(sl); } ++eRc; }
      else if (f == FL_REALT) { if (eRt < nBt) { listR[nBc + eRt] = static_cast<unsigned short>(sl); listF[posF] = static_cast<unsigned short>(sl); } ++eRt; }
      new_f4 = (new_f4 & ~(0xFFu << (8 * q))) | (f << (8 * q));
    }
    reinterpret_cast<unsigned int*>(flag)[tid] = new_f4;
    const int nF = nFl + nRefill + nBc + nBt;
    __syncthreads();
    if (nF == 0 && nRet == 0) break;                               // nothing in flight, nothing parked, nothing to write back or load

    // ================= (2) retire + refill, dense: electron k of the list <-> global element cursor + k =================
    for (int k = tid; k < nRet; k += STREAM_THREADS) {
      const unsigned int e = listO[k];
      const int sl = static_cast<int>(e & 0x7FFFu);
      const long long pos = out_ptr + k;
#pragma unroll
      for (int c = 0; c < 8; ++c) __stcs(&gcol[c][pos], col[c * POOL + sl]);
      sid.id[pos] = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * POOL + sl]));
      if (e & 0x8000u) {                                             // attached: population control refills this position at t_sync
        const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
        if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(pos); L.dead_flag[pos] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
      }
      const unsigned int kin = listK[k];
      if (kin != 0xFFFFu) {                                          // refill the slot just written back
        const long long pin = in_ptr + kin;
#pragma unroll
        for (int c = 0; c < 8; ++c) cp_async8(&col[c * POOL + sl], &gcol[c][pin]);
        cp_async8(&col[SC_ID * POOL + sl], &sid.id[pin]);
        col[SC_T * POOL + sl] = a.t0; s_used[sl] = 0;
      }
    }
    for (int k = tid; k < nRefill; k += STREAM_THREADS) {          // refills of slots that were already empty (start and end of the range)
      const unsigned int sl = listI[k];
      if (sl == 0xFFFFu) continue;
      const long long pos = in_ptr + k;
#pragma unroll
      for (int c = 0; c < 8; ++c) cp_async8(&col[c * POOL + sl], &gcol[c][pos]);
      cp_async8(&col[SC_ID * POOL + sl], &sid.id[pos]);
      col[SC_T * POOL + sl] = a.t0; s_used[sl] = 0;
    }
    cp_async_commit();
    in_ptr += nRefill; out_ptr += nRet;
    {   // pull the next round's input lines towards L2 while this round computes
      const long long ahead = in_ptr + static_cast<long long>(tid) * 16;   // 16 doubles = one 128-byte line per thread and column
      if (ahead < hi && tid < 48) {
#pragma unroll
        for (int c = 0; c < 8; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(&gcol[c][ahead]));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&sid.id[ahead]));
      }
    }

    // ================= (3) phase B: collisions on the compacted lists (BMC.C:916-1031, 1054-1280), cold-gas branch first =================
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 0 ? (GT == GT_TRUE) : (GT == GT_FALSE)) continue;
      const int first = pass == 0 ? 0 : nBc, count = pass == 0 ? nBc : nBt;
      for (int chunk_i = warp; chunk_i * 32 < count; chunk_i += STREAM_WARPS) {
        const int k = chunk_i * 32 + lane;
        int chosen = NOT_ADVANCED;
        double dE = 0;
        if (k < count) {
          const int sl = listR[first + k];
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
          if (pass == 1) chosen = thermal_collide<GT>(m, p, rng, o);
          else chosen = cold_collide<GT>(m, p, col[SC_TCF * POOL + sl], rng, o);
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
        tally_collisions_fx(chosen, dE, s_cnt, s_gain, s_loss, lane);
      }
    }
    cp_async_wait_all();   // this thread's refills have landed; the barrier publishes everybody's (and phase B's results)
    __syncthreads();

    // ================= (4) phase A: one free flight + null-collision test per electron (BMC.C:650-667, 804-905, 1035-1053) =================
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
          // one convergent draw site: the free time (if needed) and the null-test uniform come from the same Philox block
          const unsigned long long id = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * POOL + sl]));
          const bool need_tcf = (p.tcf == NON_DEF);
          unsigned int used = s_used[sl];
          if (need_tcf) used = (used + 1u) & ~1u;                    // free-time draws start on an even index (PhiloxRng::align)
          uint32_t o4[4];
          philox4x32_10(static_cast<uint32_t>(id), static_cast<uint32_t>(id >> 32), a.interval, used >> 1, k0, k1, o4);
          const double u0 = u52(o4[1], o4[0]), u1 = u52(o4[3], o4[2]);
          const double drawn = -log(u0) / a.nu_trial;               // BMC.C:650-655 (computed by every lane, used by those that need it)
          if (need_tcf) { p.tcf = drawn; p.nue = a.nu_trial; ++used; }
          const double u_null = (used & 1u) ? u1 : u0;
          const bool partial = (p.t + p.tcf > a.t_sync);             // BMC.C:657
          const double dt = partial ? (a.t_sync - p.t) : p.tcf;
          gain_field += flight<FIELD>(m, p, dt);                     // one flight site for both outcomes (BMC.C:659, :666)
          unsigned char outcome;
          if (partial) { p.t = a.t_sync; p.tcf -= dt; outcome = FL_DONE; max_end = fmax(max_end, p.eps); }
          else {
            p.t += p.tcf;
            if (thermal_branch<GT>(m, p.eps)) { outcome = FL_REALT; p.tcf = NON_DEF; }   // the thermal-target branch draws its own numbers in phase B
            else {
              EventOut o; o.table_clamped = 0; o.nu_exceeded = 0;
              double Rnu;
              ++used;
              if (cold_null_test_u(m, p, u_null, Rnu, o)) { outcome = FL_REAL; p.tcf = Rnu; }
              else { outcome = FL_FLIGHT; p.tcf = NON_DEF; ++n_null; }
              n_clamp += o.table_clamped; n_nuex += o.nu_exceeded;
            }
          }
          max_seen = fmax(max_seen, p.eps);
          col[SC_X * POOL + sl] = p.x; col[SC_Y * POOL + sl] = p.y; col[SC_Z * POOL + sl] = p.z;
          col[SC_VX * POOL + sl] = p.vx; col[SC_VY * POOL + sl] = p.vy; col[SC_VZ * POOL + sl] = p.vz;
          col[SC_TCF * POOL + sl] = p.tcf; col[SC_NUE * POOL + sl] = p.nue; col[SC_T * POOL + sl] = p.t;
          s_used[sl] = used;
          flag[sl] = outcome;
        }
      }
    }
    __syncthreads();
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
  {   // partials: header from the warp accumulators, per-process tallies converted from fixed point
    const int len = R_HEADER + 3 * m.P;
    double* out = partials + static_cast<size_t>(blockIdx.x) * len;
    for (int j = tid; j < R_HEADER; j += STREAM_THREADS) {
      double v = s_acc[0][j];
      for (int w = 1; w < STREAM_WARPS; ++w) v = (j >= R_SUM_COUNT) ? fmax(v, s_acc[w][j]) : v + s_acc[w][j];
      out[j] = v;
    }
    for (int k = tid; k < m.P; k += STREAM_THREADS) {
      out[R_HEADER + k] = static_cast<double>(s_cnt[k]);
      out[R_HEADER + m.P + k] = static_cast<double>(s_gain[k]) / TALLY_SCALE;
      out[R_HEADER + 2 * m.P + k] = -static_cast<double>(s_loss[k]) / TALLY_SCALE;
    }
  }
}

}  // namespace lk
