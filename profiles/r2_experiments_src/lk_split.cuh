// lk_split.cuh -- K1 as a pipeline of two kinds of kernels: the production advance path for large ensembles (round 2).
//
// What one synchronisation interval asks of an electron is a chain  flight, flight, ..., flight  in which every flight ends either at
// t_sync (done), in a null collision (72 % of the events of N2 at 100 Td: the chain simply goes on) or in a real collision.  The two
// earlier shared-memory kernels (lk_stream.cuh, lk_lane.cuh) keep flights and collisions in ONE kernel; measured on B200 that costs, in
// this order: 128 registers per thread for everybody because the collision code needs them (16 warps per SM, issue slots 41 % busy),
// a shared-memory round trip plus a block scan around every single flight (39 % of all instructions), and barriers that exist only to
// keep 20 KB of flight code and 80 KB of collision code from thrashing the instruction cache (profiles/r1_v17_*, profiles/r2_lane_*).
// Here the two halves are separate kernels that hand electrons over through queues in device memory:
//
//   k_flight   (BMC.C:650-667, 804-905, 1035-1053)  one electron per LANE, kept in registers for as long as its events are null
//              collisions; no collision code, no barrier in the loop.  A lane whose electron reached t_sync writes it to the output
//              cursor of the warp's range; a lane whose electron passed the null test appends it to the warp's segment of the collision
//              queue; either way the lane takes the next input electron from a small cp.async-fed ring in shared memory, so all 32
//              lanes fly in every iteration.
//   k_collide  (BMC.C:916-1031, 1054-1280)  one queue entry per thread, full warps of cold-gas picks or of thermal-target picks;
//              process selection, scattering, birth/death bookkeeping and the per-process tallies of BMC.C:1308-1328; the post-collision
//              electron is written back in place and becomes input of the next k_flight round.
//   k_tail     after LK_SPLIT_ROUNDS rounds ~1 % of the ensemble is still on its way; one thread per electron finishes them with the
//              complete event loop (the queues of later rounds would be launches with nothing in them).
//
// Every warp of round 0 owns a contiguous range ("segment") of the ensemble, and everything that descends from it stays in that segment:
// finished electrons fill the range from the bottom, the ids of attached electrons from the top, queue entries of cold-gas picks sit at
// the bottom of the same index range of the queue arrays and thermal-target picks at its top.  All positions come from ballot ranks and
// per-segment counters, never from atomics, so the result is reproducible run to run; the `id` column keys each electron's counter-based
// draw stream, so the physics is bit-identical to the one-thread-per-electron kernel (tests/test_gpu_parity.py).
#pragma once
#include "lk_stream.cuh"

namespace lk {

#ifndef LK_SPLIT_ROUNDS
#define LK_SPLIT_ROUNDS 3
#endif
constexpr int SP_ROUNDS = LK_SPLIT_ROUNDS;      // flight/collide rounds before the tail kernel
constexpr int FL_THREADS = 256;
constexpr int FL_WARPS = FL_THREADS / 32;
constexpr int FL_CTAS_PER_SM = 4;
constexpr int FL_RING = 64;                     // staged input electrons per warp (two cp.async batches of 32)
enum : int { QC_X = 0, QC_Y, QC_Z, QC_VX, QC_VY, QC_VZ, QC_AUX, QC_NUE, QC_T, QC_ID, QC_COLS };   // QC_AUX: nu_e * U of a cold-gas pick, the free time otherwise
constexpr unsigned int QF_DEAD = 0x80000000u;   // flag in the `used` word of a queue entry: the electron attached in k_collide
constexpr int FL_NU_ROWS = 512;                 // rows of nu_tot staged per CTA (four CTAs per SM share 227 KB)

__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

struct Queue {          // SoA with the index space of the ensemble: entry p of segment s lives at seg_lo[s] + p (cold) or seg_lo[s] + seg_len[s] - 1 - p (thermal)
  double* col;          // [QC_COLS][n]
  unsigned int* used;   // [n] draws consumed so far (| QF_DEAD)
};

struct Segs {           // per-segment bookkeeping, all in device memory; a segment = the range of one warp of the round-0 flight kernel
  int nseg;
  long long seg_len;    // electrons per segment (a multiple of 32); segment s = [s * seg_len, min((s + 1) * seg_len, n))
  int* out_cur;         // [nseg] finished electrons written so far (they fill the segment's range from the bottom)
  int* dead_cur;        // [nseg] ids of attached electrons written so far (from the top)
  int* qc;              // [2][nseg] cold-gas entries in queue 0 / 1
  int* qt;              // [2][nseg] thermal-target entries
  int* offc;            // [2][nseg + 1] exclusive prefix sums of qc (k_seg_scan)
  int* offt;            // [2][nseg + 1]
};

// ------------------------------------------------------------------ k_flight ------------------------------------------------------------------
constexpr size_t FM_GF = 0;                                                           // [FL_THREADS] doubles: field gain per thread
constexpr size_t FM_TMAX = FM_GF + static_cast<size_t>(FL_THREADS) * 8;               // [2][FL_THREADS] doubles
constexpr size_t FM_FL = FM_TMAX + static_cast<size_t>(FL_THREADS) * 16;              // [FL_WARPS] u64 flights
constexpr size_t FM_MISC = FM_FL + static_cast<size_t>(FL_WARPS) * 8;                 // [8] u32
constexpr size_t FM_RING = FM_MISC + 32;                                              // per warp: [QC_COLS][FL_RING] doubles + [FL_RING] u32
constexpr size_t FW_BYTES = static_cast<size_t>(QC_COLS) * FL_RING * 8 + static_cast<size_t>(FL_RING) * 4;
constexpr size_t FM_NU = FM_RING + static_cast<size_t>(FL_WARPS) * FW_BYTES;          // [FL_NU_ROWS] doubles
constexpr size_t FM_BYTES = FM_NU + static_cast<size_t>(FL_NU_ROWS) * 8;
static_assert(FM_RING % 16 == 0 && FW_BYTES % 16 == 0 && FM_NU % 16 == 0, "alignment");

// header-only partials of the flight kernels
enum : int { FP_FLIGHTS = 0, FP_GAIN, FP_MAX_END, FP_MAX_SEEN, FP_CLAMP, FP_NUEX, FP_COUNT = 8 };

// ROUND0: input = the ensemble (8 state columns + id; clock t0, no draws used).  Otherwise input = queue `qin` (post-collision electrons of
// this warp's segment: cold part, then thermal part).  Real collisions go to `qout`.
template <int FIELD, int GT, bool ROUND0>
__global__ void __launch_bounds__(FL_THREADS, FL_CTAS_PER_SM) k_flight(const Model m, const StateId sid, const Lists L, const AdvArgs a, const Segs sg, const Queue qin,
                                                                        const Queue qout, const int par_in, double* __restrict__ fpart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_gf = reinterpret_cast<double*>(smem_raw + FM_GF);
  double* s_tmax = reinterpret_cast<double*>(smem_raw + FM_TMAX);
  unsigned long long* s_fl = reinterpret_cast<unsigned long long*>(smem_raw + FM_FL);
  unsigned int* s_misc = reinterpret_cast<unsigned int*>(smem_raw + FM_MISC);
  const double* s_nu = reinterpret_cast<const double*>(smem_raw + FM_NU);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* ring = reinterpret_cast<double*>(smem_raw + FM_RING + static_cast<size_t>(warp) * FW_BYTES);    // [QC_COLS][FL_RING]
  unsigned int* ring_used = reinterpret_cast<unsigned int*>(ring + QC_COLS * FL_RING);                    // [FL_RING]

  if (tid < 8) s_misc[tid] = 0;
  for (int j = tid; j < static_cast<int>(a.pad); j += FL_THREADS) reinterpret_cast<double*>(smem_raw + FM_NU)[j] = __ldg(&m.nu_tot[j]);
  __syncthreads();

  const int seg = blockIdx.x * FL_WARPS + warp;
  const bool has_seg = seg < sg.nseg;
  const long long lo = has_seg ? min(static_cast<long long>(seg) * sg.seg_len, a.n) : a.n;
  const int len = has_seg ? static_cast<int>(min(lo + sg.seg_len, a.n) - lo) : 0;
  double* const g0 = sid.s.x + lo;
  unsigned long long* const gid = sid.id + lo;
  const int par_out = par_in ^ 1;

  // warp-uniform cursors
  int out_cur = 0, dead_cur = 0, nC_in = 0, in_n = len;
  if (!ROUND0 && has_seg) { out_cur = sg.out_cur[seg]; dead_cur = sg.dead_cur[seg]; nC_in = sg.qc[par_in * sg.nseg + seg]; in_n = nC_in + sg.qt[par_in * sg.nseg + seg]; }
  int in_cur = 0;                 // next input element to stage
  int issued = 0, landed = 0, cons = 0;   // ring positions (free-running; slot = position & (FL_RING - 1))
  int qc = 0, qt = 0;             // entries appended to the output queue
  unsigned long long flights = 0ull;

  double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, tcf = 0, nue = 1, t = a.t0;
  unsigned long long id = 0;
  unsigned int used = 0;
  bool live = false;

  const int staged_rows = static_cast<int>(a.pad);
  const double rnu = recip_for_div(a.nu_trial);
  const unsigned lt = lanemask_lt();
  double gain_sum = 0, seen_max = 0, end_max = 0;

  auto stage = [&]() {   // issue the next batch of up to 32 input electrons into the ring
    const int cnt = min(32, in_n - in_cur);
    if (lane < cnt) {
      const int sl = (issued + lane) & (FL_RING - 1);
      const int i = in_cur + lane;
      if (ROUND0) {
        const double* const gq = g0 + i;
#pragma unroll
        for (int c = 0; c < 8; ++c) cp_async8(&ring[c * FL_RING + sl], gq + c * a.n);   // state column c = queue column c (QC_AUX = t_cf, QC_NUE)
        cp_async8(&ring[QC_ID * FL_RING + sl], &gid[i]);
        ring[QC_T * FL_RING + sl] = a.t0; ring_used[sl] = 0u;
      } else {
        const long long p = (i < nC_in) ? (lo + i) : (lo + len - 1 - (i - nC_in));   // cold part from the bottom, thermal part from the top
#pragma unroll
        for (int c = 0; c < QC_COLS; ++c) cp_async8(&ring[c * FL_RING + sl], qin.col + c * a.n + p);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned int>(__cvta_generic_to_shared(&ring_used[sl]))), "l"(qin.used + p) : "memory");
      }
    }
    cp_async_commit();
    issued += cnt; in_cur += cnt;
  };

  if (in_n > 0) stage();
  if (in_cur < in_n) stage();

#pragma unroll 1
  for (;;) {
    // ---- refill the empty lanes from the ring ----
    {
      const unsigned me = __ballot_sync(FULL, !live);
      if (me != 0u && cons < issued) {
        const int want = __popc(me);
        if (landed - cons < want && landed < issued) { cp_async_wait_all(); __syncwarp(); landed = issued; }
        const int avail = landed - cons, rk = __popc(me & lt);
        bool dead = false;
        if (!live && rk < avail) {
          const int sl = (cons + rk) & (FL_RING - 1);
          x = ring[QC_X * FL_RING + sl]; y = ring[QC_Y * FL_RING + sl]; z = ring[QC_Z * FL_RING + sl];
          vx = ring[QC_VX * FL_RING + sl]; vy = ring[QC_VY * FL_RING + sl]; vz = ring[QC_VZ * FL_RING + sl];
          tcf = ring[QC_AUX * FL_RING + sl]; nue = ring[QC_NUE * FL_RING + sl]; t = ring[QC_T * FL_RING + sl];
          id = static_cast<unsigned long long>(__double_as_longlong(ring[QC_ID * FL_RING + sl]));
          used = ring_used[sl];
          dead = !ROUND0 && (used & QF_DEAD);
          live = !dead;
        }
        cons += min(want, avail);
        if (!ROUND0) {   // an electron that attached in k_collide leaves through the top of the segment's range: population control refills that position at t_sync
          const unsigned md = __ballot_sync(FULL, dead);
          if (dead) {
            const int r = len - 1 - (dead_cur + __popc(md & lt));
            gid[r] = id;
            const long long pos = lo + r;
            const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
            if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(pos); L.dead_flag[pos] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
          }
          dead_cur += __popc(md);
        }
        __syncwarp();
        if (issued - cons <= 32 && in_cur < in_n) stage();   // keep one batch landed and one on its way
      }
    }
    const unsigned ml = __ballot_sync(FULL, live);
    if (ml == 0u) { if (cons >= issued && in_cur >= in_n) break; continue; }
    flights += static_cast<unsigned long long>(__popc(ml));

    // ---- one free flight + null-collision test (BMC.C:650-667, 804-905, 1035-1053), branch-free ----
    const bool need = live && (tcf == NON_DEF);
    unsigned int usedv = need ? ((used + 1u) & ~1u) : used;          // free-time draws start on an even index (PhiloxRng::align)
    uint32_t o4[4];
    philox4x32_10(static_cast<uint32_t>(id), static_cast<uint32_t>(id >> 32), a.interval, usedv >> 1, static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32), o4);
    const double u0 = u52(o4[1], o4[0]), u1 = u52(o4[3], o4[2]);
    if (__any_sync(FULL, need)) {
      const double drawn = div_by(-log_normal(u0), a.nu_trial, rnu);   // -log(u) / nu_trial, BMC.C:650-655
      if (need) { tcf = drawn; nue = a.nu_trial; ++usedv; }
    }
    Particle p;
    p.x = x; p.y = y; p.z = z; p.vx = vx; p.vy = vy; p.vz = vz; p.tcf = tcf; p.nue = nue; p.t = t;
    p.eps = kinetic_eV(vx, vy, vz);
    const double u_null = (usedv & 1u) ? u1 : u0;
    const bool partial = (p.t + p.tcf > a.t_sync);                     // BMC.C:657
    const double dt = partial ? (a.t_sync - p.t) : p.tcf;
    const double gain = flight<FIELD>(m, p, dt);                       // one flight site for both outcomes (BMC.C:659, :666)
    const bool thermal = thermal_branch<GT>(m, p.eps);
    double Rnu; bool clamped, exceeded;
    const bool real = stream_null_test(m, s_nu, staged_rows, p.eps, p.nue, u_null, Rnu, clamped, exceeded);
    const bool tested = !partial && !thermal;                          // the thermal-target branch draws its own numbers in k_collide
    if (live) {
      x = p.x; y = p.y; z = p.z; vx = p.vx; vy = p.vy; vz = p.vz;
      tcf = partial ? (p.tcf - dt) : (tested && real) ? Rnu : NON_DEF;
      t = partial ? a.t_sync : (p.t + p.tcf);
      used = usedv + (tested ? 1u : 0u);
      gain_sum += gain;
      seen_max = fmax(seen_max, p.eps);
      if (partial) end_max = fmax(end_max, p.eps);
      if (tested && (clamped || exceeded)) { if (clamped) atomicAdd(&s_misc[MC_CLAMP], 1u); if (exceeded) atomicAdd(&s_misc[MC_NUEX], 1u); }
    }
    // ---- done: dense stores at the output cursor of the segment ----
    const bool ret = live && partial;
    const bool toC = live && !partial && !thermal && real, toT = live && !partial && thermal;
    const unsigned mr = __ballot_sync(FULL, ret), mc = __ballot_sync(FULL, toC), mt = __ballot_sync(FULL, toT);
    if (ret) {
      const int r = out_cur + __popc(mr & lt);
      double* const gp = g0 + r;
      __stcs(gp, x); __stcs(gp + a.n, y); __stcs(gp + 2 * a.n, z); __stcs(gp + 3 * a.n, vx); __stcs(gp + 4 * a.n, vy);
      __stcs(gp + 5 * a.n, vz); __stcs(gp + 6 * a.n, tcf); __stcs(gp + 7 * a.n, nue);
      gid[r] = id;
    }
    out_cur += __popc(mr);
    // ---- real collision: append to the segment of the collision queue (cold-gas picks from the bottom, thermal-target picks from the top) ----
    if (toC || toT) {
      const long long pq = toC ? (lo + qc + __popc(mc & lt)) : (lo + len - 1 - (qt + __popc(mt & lt)));
      double* const qp = qout.col + pq;
      qp[QC_X * a.n] = x; qp[QC_Y * a.n] = y; qp[QC_Z * a.n] = z; qp[QC_VX * a.n] = vx; qp[QC_VY * a.n] = vy; qp[QC_VZ * a.n] = vz;
      qp[QC_AUX * a.n] = tcf; qp[QC_NUE * a.n] = nue; qp[QC_T * a.n] = t; qp[QC_ID * a.n] = __longlong_as_double(static_cast<long long>(id));
      qout.used[pq] = used;
    }
    qc += __popc(mc); qt += __popc(mt);
    live = live && !(ret || toC || toT);
  }

  // ---- per-segment state for the next stage, per-CTA header partials ----
  if (lane == 0 && has_seg) {
    sg.out_cur[seg] = out_cur; sg.dead_cur[seg] = dead_cur;
    sg.qc[par_out * sg.nseg + seg] = qc; sg.qt[par_out * sg.nseg + seg] = qt;
  }
  s_gf[tid] = gain_sum; s_tmax[tid] = end_max; s_tmax[FL_THREADS + tid] = seen_max;
  if (lane == 0) s_fl[warp] = flights;
  __syncthreads();
  if (tid == 0) {
    unsigned long long fl = 0;
    double gf = 0, m0 = 0, m1 = 0;
    for (int w = 0; w < FL_WARPS; ++w) {
      fl += s_fl[w];
      double ws = 0;
      for (int l = 0; l < 32; ++l) { ws += s_gf[w * 32 + l]; m0 = fmax(m0, s_tmax[w * 32 + l]); m1 = fmax(m1, s_tmax[FL_THREADS + w * 32 + l]); }
      gf += ws;
    }
    double* out = fpart + static_cast<size_t>(blockIdx.x) * FP_COUNT;
    out[FP_FLIGHTS] = static_cast<double>(fl); out[FP_GAIN] = gf; out[FP_MAX_END] = m0; out[FP_MAX_SEEN] = fmax(m0, m1);
    out[FP_CLAMP] = static_cast<double>(s_misc[MC_CLAMP]); out[FP_NUEX] = static_cast<double>(s_misc[MC_NUEX]);
    out[6] = 0; out[7] = 0;
  }
}

// ------------------------------------------------------------------ segment prefix sums ------------------------------------------------------------------
// offc / offt of queue `par`: exclusive prefix sums of the per-segment entry counts (one CTA; nseg is a few thousand)
__global__ void __launch_bounds__(1024) k_seg_scan(const Segs sg, const int par) {
  __shared__ int s_part[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* qc = sg.qc + par * sg.nseg;
  const int* qt = sg.qt + par * sg.nseg;
  int* offc = sg.offc + par * (sg.nseg + 1);
  int* offt = sg.offt + par * (sg.nseg + 1);
  int carry_c = 0, carry_t = 0;
  for (int base = 0; base < sg.nseg; base += 1024) {
    const int i = base + tid;
    const int vc = (i < sg.nseg) ? qc[i] : 0, vt = (i < sg.nseg) ? qt[i] : 0;
    int ic = vc, it = vt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(FULL, ic, o), b = __shfl_up_sync(FULL, it, o); if (lane >= o) { ic += a; it += b; } }
    if (lane == 31) { s_part[0][warp] = ic; s_part[1][warp] = it; }
    __syncthreads();
    int bc = 0, bt = 0, tc = 0, tt = 0;
    for (int w = 0; w < 32; ++w) { const int a = s_part[0][w], b = s_part[1][w]; if (w < warp) { bc += a; bt += b; } tc += a; tt += b; }
    if (i < sg.nseg) { offc[i] = carry_c + bc + ic - vc; offt[i] = carry_t + bt + it - vt; }
    carry_c += tc; carry_t += tt;
    __syncthreads();
  }
  if (tid == 0) { offc[sg.nseg] = carry_c; offt[sg.nseg] = carry_t; }
}

// entry index i of the concatenated per-segment lists -> (segment, position inside the segment's list): binary search on the prefix sums
__device__ __forceinline__ int seg_of(const int* __restrict__ off, int nseg, int i) {
  int lo = 0, hi = nseg;   // off[lo] <= i < off[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(&off[mid]) <= i) lo = mid; else hi = mid; }
  return lo;
}

// ------------------------------------------------------------------ k_collide ------------------------------------------------------------------
constexpr int CO_THREADS = 256;
// KIND 0: the cold-gas picks of queue `par` (BMC.C:1054-1097), KIND 1: its thermal-target picks (BMC.C:916-1031); then the dynamics of
// the chosen process (BMC.C:1101-1111, 1115-1280) and the tallies of BMC.C:1308-1328.  Results are written back in place.
template <int GT, int KIND>
__global__ void __launch_bounds__(CO_THREADS, 2) k_collide(const Model m, const Lists L, const Pending pend, const AdvArgs a, const Segs sg, const Queue q, const int par,
                                                            double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tally* s_tally = reinterpret_cast<Tally*>(smem_raw);                                     // [P]
  __shared__ unsigned int s_misc[8];
  __shared__ double s_seen[CO_THREADS];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int k = tid; k < m.P; k += CO_THREADS) { s_tally[k].gain = 0; s_tally[k].loss = 0; s_tally[k].cnt = 0; }
  if (tid < 8) s_misc[tid] = 0;
  __syncthreads();
  const int* __restrict__ off = (KIND == 0 ? sg.offc : sg.offt) + par * (sg.nseg + 1);
  const int total = off[sg.nseg];
  double seen = 0;
  const int total_r = (total + 31) & ~31;
#pragma unroll 1
  for (int i = blockIdx.x * CO_THREADS + tid; i < total_r; i += gridDim.x * CO_THREADS) {
    int chosen = NOT_ADVANCED;
    double dE = 0;
    if (i < total) {
      const int s = seg_of(off, sg.nseg, i);
      const int j = i - __ldg(&off[s]);
      const long long lo = static_cast<long long>(s) * sg.seg_len;
      const int len = static_cast<int>(min(lo + sg.seg_len, a.n) - lo);
      const long long pq = (KIND == 0) ? (lo + j) : (lo + len - 1 - j);
      double* const qp = q.col + pq;
      Particle p;
      p.x = qp[QC_X * a.n]; p.y = qp[QC_Y * a.n]; p.z = qp[QC_Z * a.n]; p.vx = qp[QC_VX * a.n]; p.vy = qp[QC_VY * a.n]; p.vz = qp[QC_VZ * a.n];
      p.nue = qp[QC_NUE * a.n]; p.t = qp[QC_T * a.n]; p.tcf = NON_DEF;
      p.eps = kinetic_eV(p.vx, p.vy, p.vz);
      const unsigned long long id = static_cast<unsigned long long>(__double_as_longlong(qp[QC_ID * a.n]));
      PhiloxRng rng;
      rng.k0 = static_cast<uint32_t>(a.seed); rng.k1 = static_cast<uint32_t>(a.seed >> 32);
      rng.c0 = static_cast<uint32_t>(id); rng.c1 = static_cast<uint32_t>(id >> 32); rng.c2 = a.interval;
      rng.used = q.used[pq]; rng.blk = 0xFFFFFFFFu;
      EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
      double Vx = 0, Vy = 0, Vz = 0;
      if (KIND == 1) chosen = thermal_select(m, p, rng, o, Vx, Vy, Vz);
      else chosen = cold_select(m, p, qp[QC_AUX * a.n]);
      if (chosen != NULL_COLLISION) chosen = collide_dynamics<GT>(m, chosen, p, Vx, Vy, Vz, rng, o);
      if (o.table_clamped) atomicAdd(&s_misc[MC_CLAMP], 1u);
      if (o.nu_exceeded) atomicAdd(&s_misc[MC_NUEX], 1u);
      unsigned int flags = 0u;
      if (chosen >= 0) {
        dE = o.dE;
        const int type = __ldg(&m.type[chosen]);
        if (type == T_IONIZATION) {
          atomicAdd(&s_misc[MC_BORN], 1u);
          uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, rng.used, cc1, ck1);
          push_pending(pend, L.counters, o, p.t, rng.c0, cc1, ck1);
        } else if (type == T_ATTACHMENT) { atomicAdd(&s_misc[MC_ATT], 1u); flags = QF_DEAD; }
      }                                                              // aborted picks count as null collisions (BMC.C:1137-1140)
      seen = fmax(seen, p.eps);
      qp[QC_VX * a.n] = p.vx; qp[QC_VY * a.n] = p.vy; qp[QC_VZ * a.n] = p.vz;
      qp[QC_AUX * a.n] = NON_DEF;                                    // the next free time is drawn at the start of the next flight (same stream position)
      q.used[pq] = rng.used | flags;
    }
    tally_collisions_fx(chosen, dE, s_tally, lane);
  }
  s_seen[tid] = seen;
  __syncthreads();
  {
    const int plen = R_HEADER + 3 * m.P;
    double* out = partials + static_cast<size_t>(blockIdx.x) * plen;
    if (tid == 0) {
      double nr = 0, mx = 0;
      for (int k = 0; k < m.P; ++k) nr += static_cast<double>(s_tally[k].cnt);
      for (int k = 0; k < CO_THREADS; ++k) mx = fmax(mx, s_seen[k]);
      for (int j = 0; j < R_HEADER; ++j) out[j] = 0;
      out[R_N_REAL] = nr; out[R_N_BORN] = static_cast<double>(s_misc[MC_BORN]); out[R_N_ATTACHED] = static_cast<double>(s_misc[MC_ATT]);
      out[R_N_TABLE_CLAMPED] = static_cast<double>(s_misc[MC_CLAMP]); out[R_N_NU_EXCEEDED] = static_cast<double>(s_misc[MC_NUEX]);
      out[R_MAX_EPS_SEEN] = mx;
    }
    for (int k = tid; k < m.P; k += CO_THREADS) {
      out[R_HEADER + k] = static_cast<double>(s_tally[k].cnt);
      out[R_HEADER + m.P + k] = static_cast<double>(s_tally[k].gain) / TALLY_SCALE;
      out[R_HEADER + 2 * m.P + k] = -static_cast<double>(s_tally[k].loss) / TALLY_SCALE;
    }
  }
}

// ------------------------------------------------------------------ k_tail ------------------------------------------------------------------
// The electrons still in queue `par` after the last collide round (about 1 % of the ensemble): one per thread, complete event loop to
// t_sync (BMC.C:637-681).  Entry j of segment s finishes at position out_cur[s] + j of the segment's range, whatever happens to it.
template <int FIELD, int GT>
__global__ void __launch_bounds__(CO_THREADS, 2) k_tail(const Model m, const StateId sid, const Lists L, const Pending pend, const AdvArgs a, const Segs sg, const Queue q,
                                                         const int par, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tally* s_tally = reinterpret_cast<Tally*>(smem_raw);
  __shared__ unsigned int s_misc[8];
  __shared__ double s_red[3][CO_THREADS];
  __shared__ unsigned long long s_flights[CO_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int k = tid; k < m.P; k += CO_THREADS) { s_tally[k].gain = 0; s_tally[k].loss = 0; s_tally[k].cnt = 0; }
  if (tid < 8) s_misc[tid] = 0;
  __syncthreads();
  const int* __restrict__ offc = sg.offc + par * (sg.nseg + 1);
  const int* __restrict__ offt = sg.offt + par * (sg.nseg + 1);
  const int totalC = offc[sg.nseg], total = totalC + offt[sg.nseg];
  const int total_r = (total + 31) & ~31;
  double gain_field = 0, max_end = 0, max_seen = 0;
  unsigned long long flights = 0ull;
#pragma unroll 1
  for (int i = blockIdx.x * CO_THREADS + tid; i < total_r; i += gridDim.x * CO_THREADS) {
    bool active = i < total;
    Particle p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    PhiloxRng rng; rng.init(a.seed, 0, a.interval);
    long long lo = 0, pos = 0;
    int len = 0;
    unsigned long long id = 0;
    bool dead = false;
    if (active) {
      const bool cold = i < totalC;
      const int ii = cold ? i : i - totalC;
      const int s = seg_of(cold ? offc : offt, sg.nseg, ii);
      const int j = ii - __ldg(&(cold ? offc : offt)[s]);
      lo = static_cast<long long>(s) * sg.seg_len;
      len = static_cast<int>(min(lo + sg.seg_len, a.n) - lo);
      const long long pq = cold ? (lo + j) : (lo + len - 1 - j);
      const int nc_s = __ldg(&offc[s + 1]) - __ldg(&offc[s]);
      pos = lo + sg.out_cur[s] + (cold ? j : nc_s + j);
      const double* const qp = q.col + pq;
      p.x = qp[QC_X * a.n]; p.y = qp[QC_Y * a.n]; p.z = qp[QC_Z * a.n]; p.vx = qp[QC_VX * a.n]; p.vy = qp[QC_VY * a.n]; p.vz = qp[QC_VZ * a.n];
      p.tcf = qp[QC_AUX * a.n]; p.nue = qp[QC_NUE * a.n]; p.t = qp[QC_T * a.n];
      p.eps = kinetic_eV(p.vx, p.vy, p.vz);
      id = static_cast<unsigned long long>(__double_as_longlong(qp[QC_ID * a.n]));
      const unsigned int w = q.used[pq];
      rng.init(a.seed, id, a.interval, w & ~QF_DEAD);
      dead = (w & QF_DEAD) != 0u;
      if (dead) active = false;
    }
    while (__any_sync(FULL, active)) {
      int chosen = NOT_ADVANCED;
      double dE = 0;
      if (active) {
        EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
        chosen = event<FIELD, GT>(m, p, a.nu_trial, a.t_sync, rng, o);
        ++flights;
        gain_field += o.gain_field;
        max_seen = fmax(max_seen, p.eps);
        if (o.table_clamped) atomicAdd(&s_misc[MC_CLAMP], 1u);
        if (o.nu_exceeded) atomicAdd(&s_misc[MC_NUEX], 1u);
        if (chosen == PARTIAL_FLIGHT) {
          max_end = fmax(max_end, p.eps);
          double* const gp = sid.s.x + pos;
          gp[0] = p.x; gp[a.n] = p.y; gp[2 * a.n] = p.z; gp[3 * a.n] = p.vx; gp[4 * a.n] = p.vy; gp[5 * a.n] = p.vz; gp[6 * a.n] = p.tcf; gp[7 * a.n] = p.nue;
          sid.id[pos] = id;
          active = false;
        } else if (chosen >= 0) {
          dE = o.dE;
          const int type = __ldg(&m.type[chosen]);
          if (type == T_IONIZATION) {
            atomicAdd(&s_misc[MC_BORN], 1u);
            uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, o.used_mark, cc1, ck1);
            push_pending(pend, L.counters, o, p.t, rng.c0, cc1, ck1);
          } else if (type == T_ATTACHMENT) { atomicAdd(&s_misc[MC_ATT], 1u); dead = true; active = false; }
        }
      }
      tally_collisions_fx(chosen, dE, s_tally, lane);
    }
    if (dead) {   // attached (here or in the last collide round): the position is refilled by population control at t_sync
      sid.id[pos] = id;
      const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
      if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(pos); L.dead_flag[pos] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
    }
  }
  s_red[0][tid] = gain_field; s_red[1][tid] = max_end; s_red[2][tid] = max_seen;
  {
    unsigned long long f = flights;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(FULL, f, o);
    if (lane == 0) s_flights[tid >> 5] = f;
  }
  __syncthreads();
  {
    const int plen = R_HEADER + 3 * m.P;
    double* out = partials + static_cast<size_t>(blockIdx.x) * plen;
    if (tid == 0) {
      double nr = 0, gf = 0, m0 = 0, m1 = 0;
      unsigned long long fl = 0;
      for (int k = 0; k < m.P; ++k) nr += static_cast<double>(s_tally[k].cnt);
      for (int k = 0; k < CO_THREADS; ++k) { gf += s_red[0][k]; m0 = fmax(m0, s_red[1][k]); m1 = fmax(m1, s_red[2][k]); }
      for (int w = 0; w < CO_THREADS / 32; ++w) fl += s_flights[w];
      for (int j = 0; j < R_HEADER; ++j) out[j] = 0;
      out[R_N_REAL] = nr; out[R_N_NULL] = static_cast<double>(fl);   // flights of the tail: k_finalize_split turns flights into events
      out[R_N_BORN] = static_cast<double>(s_misc[MC_BORN]); out[R_N_ATTACHED] = static_cast<double>(s_misc[MC_ATT]);
      out[R_N_TABLE_CLAMPED] = static_cast<double>(s_misc[MC_CLAMP]); out[R_N_NU_EXCEEDED] = static_cast<double>(s_misc[MC_NUEX]);
      out[R_GAIN_FIELD] = gf; out[R_MAX_EPS] = m0; out[R_MAX_EPS_SEEN] = fmax(m0, m1);
    }
    for (int k = tid; k < m.P; k += CO_THREADS) {
      out[R_HEADER + k] = static_cast<double>(s_tally[k].cnt);
      out[R_HEADER + m.P + k] = static_cast<double>(s_tally[k].gain) / TALLY_SCALE;
      out[R_HEADER + 2 * m.P + k] = -static_cast<double>(s_tally[k].loss) / TALLY_SCALE;
    }
  }
}

// ------------------------------------------------------------------ combining the pipeline's partials ------------------------------------------------------------------
// One block of `partials` ([R_HEADER + 3P]) that stands for the whole pipeline, so that k_finalize sees the same layout as from the
// single-kernel forms: header-only partials of the flight kernels + full partials of the collide / tail kernels, in a fixed order.
//   events = flights - electrons + attached (every electron flies once more than it has events, unless it attaches: BMC.C:1308-1320)
// Rows of skipped launches must be zero (the host clears cpart once per interval).
__global__ void __launch_bounds__(32) k_split_combine(const double* __restrict__ fpart, int f_rows, const double* __restrict__ cpart, int c_rows, int P, long long n,
                                                       double* __restrict__ out) {
  const int len = R_HEADER + 3 * P;
  const int j = blockIdx.x, lane = threadIdx.x;   // one warp per output entry, lanes stride over the rows, then a fixed-shape shuffle tree
  if (j >= len) return;
  const bool is_max = (j >= R_SUM_COUNT && j < R_HEADER);
  auto csum = [&](int col, bool mx) {
    double v = 0;
    for (int b = lane; b < c_rows; b += 32) { const double t = cpart[static_cast<size_t>(b) * len + col]; v = mx ? fmax(v, t) : v + t; }
    return mx ? warp_max(v) : warp_sum(v);
  };
  auto fsum = [&](int col, bool mx) {
    double v = 0;
    for (int b = lane; b < f_rows; b += 32) { const double t = fpart[static_cast<size_t>(b) * FP_COUNT + col]; v = mx ? fmax(v, t) : v + t; }
    return mx ? warp_max(v) : warp_sum(v);
  };
  double v = csum(j, is_max);
  if (j == R_GAIN_FIELD) v += fsum(FP_GAIN, false);
  if (j == R_MAX_EPS) v = fmax(v, fsum(FP_MAX_END, true));
  if (j == R_MAX_EPS_SEEN) v = fmax(v, fsum(FP_MAX_SEEN, true));
  if (j == R_N_TABLE_CLAMPED) v += fsum(FP_CLAMP, false);
  if (j == R_N_NU_EXCEEDED) v += fsum(FP_NUEX, false);
  if (j == R_N_NULL) {   // the R_N_NULL column of the tail's partials carries its flights (v); R_N_REAL and R_N_ATTACHED are complete
    const double real = csum(R_N_REAL, false), att = csum(R_N_ATTACHED, false), flights = fsum(FP_FLIGHTS, false) + v;
    v = (flights - static_cast<double>(n) + att) - real;
  }
  if (lane == 0) out[j] = v;
}

}  // namespace lk
