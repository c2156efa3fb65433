// lk_stream.cuh -- K1, streaming-pool form: the production advance kernel for large ensembles.
//
// (Design history, all measured on B200 and kept under profiles/: one electron per thread -> 6 of 32 lanes active, I-cache thrash;
//  CTA tile drained to empty -> 8 of 12 warps waiting at barriers; CTA-wide pool with refill, two barrier intervals per round;
//  warp-private pools -> I-cache thrash; this form.  DESIGN.md section 5 has the numbers.)
//
// A CTA owns a contiguous range of the ensemble and keeps a POOL of electrons resident in shared memory.  Every round
//   (1) one block scan over the slot flags builds four lists in slot order: continuing flights, collisions (cold-gas picks first,
//       thermal-target picks after), retiring slots, empty slots;                                                     [barrier]
//   (2) retire + refill by rank: thread t owns ranks t, t + 256, ...; it writes the r-th retiring slot to output element out + r
//       (dense stores) and loads input element in + r into the same slot with cp.async (dense loads);
//   (3) collisions of the electrons that passed the null test, in chunks of 32, chunk c on warp c % 8       (BMC.C:916-1031, 1054-1280);
//   (4) one free flight + null-collision test per electron, two electrons per lane                (BMC.C:650-667, 804-905, 1035-1053):
//       a warp flies the chunks it collided in (3), its share of the continuing chunks and - once its own copies have landed - the
//       electrons its threads refilled in (2).  Nothing a warp reads was written by another warp in this round.       [barrier]
// All warps of the CTA execute the same phase at about the same time, which is what keeps the instruction cache effective (a variant
// with warp-private pools and no CTA barriers was measured 1.7x slower: 16 desynchronised warps thrash the I-cache).  The pool stays
// full until the CTA's range is exhausted, so the Poisson tail is paid once per CTA, not once per tile.  Collisions wait in the pool
// until a whole CTA-iteration of them (256) is available.
// Electrons therefore PERMUTE inside the CTA's range (in place: the write cursor never overtakes the read cursor); the `id`
// column travels with each electron and keys its counter-based draw stream, so the physics is bit-identical to the
// one-thread-per-electron kernel.  The schedule involves no atomics, hence it is deterministic.
#pragma once
#include "lk_tile.cuh"

namespace lk {

constexpr int POOL = 1024;                       // electrons resident per CTA
// Every per-slot array has PSTRIDE entries: slot POOL is a DUMMY that nobody ever writes after the kernel's prologue.  Lanes of the flight
// phase that have no electron read it (their results are discarded), so no lane ever reads a slot that another warp may be writing:
// compute-sanitizer racecheck is clean without predicated loads (tools/sanitize_stream.py).
#ifdef LK_NO_DUMMY_SLOT
constexpr int PSTRIDE = POOL;
#else
constexpr int PSTRIDE = POOL + 8;
#endif
constexpr int STREAM_THREADS = 256;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
static_assert(POOL == 4 * STREAM_THREADS, "the scan reads the 4 flags of a thread as one 32-bit word");

enum : unsigned char { FL_REALT = 5 };   // passed the flight, waits for the thermal-target collision branch (FL_REAL = cold-gas branch)
// SC_TCF doubles as the hand-over slot for nu_e * U of an electron that waits for its collision (it has no free time then)
enum : int { SC_X = 0, SC_Y, SC_Z, SC_VX, SC_VY, SC_VZ, SC_TCF, SC_NUE, SC_T, SC_ID, SC_COLS };

// energy tallies per process are kept in 2^-36 eV fixed point: integer adds commute, so the sums are bitwise reproducible and
// a warp can aggregate them with REDUX (resolution 1.5e-11 eV per event; the reference uses them only for the power balance)
constexpr double TALLY_SCALE = 68719476736.0;

// Shared-memory layout: every array sits at a COMPILE-TIME offset from the start; the only part whose size depends on the job (one tally
// record per process) comes last.  Offsets that depend on P made ptxas 12.9 keep "base + 16 P" in a uniform register and then use that
// register as the plain base for the per-thread sums in the ECR and AC+B instantiations (found with compute-sanitizer racecheck; the
// thread-vs-stream test catches it as a wrong field gain) -- with constant offsets there is only one base to keep.
struct Tally { unsigned long long gain, loss; unsigned int cnt, pad; };   // fixed point 2^-36 eV, see TALLY_SCALE
constexpr size_t SM_COL = 0;                                                   // [SC_COLS][PSTRIDE] doubles
constexpr size_t SM_HDR = SM_COL + static_cast<size_t>(SC_COLS) * PSTRIDE * 8;    // [R_HEADER] doubles: result header of the CTA
constexpr size_t SM_GF = SM_HDR + static_cast<size_t>(R_HEADER) * 8;           // [STREAM_THREADS] doubles: per-thread field-gain sums
constexpr size_t SM_TMAX = SM_GF + static_cast<size_t>(STREAM_THREADS) * 8;    // [2][STREAM_THREADS] doubles: per-thread energy maxima
constexpr size_t SM_SCAN = SM_TMAX + static_cast<size_t>(STREAM_THREADS) * 16; // [16] u64: warp totals of the scan, range start, flight count
constexpr size_t SM_USED = SM_SCAN + 16 * 8;                                   // [PSTRIDE] u32: draw counters
constexpr size_t SM_LISTS = SM_USED + static_cast<size_t>(PSTRIDE) * 4;        // 4 x [POOL] u16
constexpr size_t SM_FLAG = SM_LISTS + static_cast<size_t>(POOL) * 2 * 4;       // [PSTRIDE] u8
constexpr size_t SM_MISC = SM_FLAG + PSTRIDE;                                  // [8] u32: rare-event counters
constexpr size_t SM_RS = SM_MISC + 32;                                         // [16] i32: CTA-uniform round state
// The first NU_STAGE_ROWS rows of nu_tot (the null test of every event interpolates two of them; they are the rows nearly every electron
// is in) are staged here when the tally records still fit behind them; otherwise the tally starts at SM_NU and the kernel reads nu_tot from
// global memory only.  Both starts are compile-time constants.
constexpr int NU_STAGE_ROWS = 1024;
constexpr size_t SM_NU = SM_RS + 64;                                           // [NU_STAGE_ROWS] doubles, optional
constexpr size_t SM_TALLY_STAGED = SM_NU + static_cast<size_t>(NU_STAGE_ROWS) * 8;   // [P] Tally when the rows are staged
constexpr size_t SM_TALLY_PLAIN = SM_NU;                                       // [P] Tally otherwise
static_assert(SM_NU % 8 == 0 && SM_SCAN % 8 == 0, "64-bit members need 8-byte offsets");
constexpr size_t STREAM_SMEM_BUDGET = (227u * 1024u) / 2u - 1024u;            // two CTAs per SM, 1 KB per CTA reserved by the driver

__host__ __device__ inline bool stream_stages_nu(int P) { return SM_TALLY_STAGED + static_cast<size_t>(P) * sizeof(Tally) + 16 <= STREAM_SMEM_BUDGET; }
__host__ __device__ inline size_t stream_smem_bytes(int P) {
  return ((stream_stages_nu(P) ? SM_TALLY_STAGED : SM_TALLY_PLAIN) + static_cast<size_t>(P) * sizeof(Tally) + 15) & ~static_cast<size_t>(15);
}

struct StateId { State s; unsigned long long* id; };   // the 8 columns of s are one allocation: column c starts at s.x + c * n

// 8-byte asynchronous global -> shared copy (LDGSTS): the refill of a freed slot overlaps with the collision phase
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned int d = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// warp-converged tally of one batch of collisions (BMC.C:1308-1328).  Lanes that chose the same process are found with MATCH,
// their fixed-point energy changes are summed with REDUX in three 21-bit limbs, and one lane per process does the shared atomics.
__device__ __forceinline__ void tally_collisions_fx(int chosen, double dE, Tally* s_tally, int lane) {
  const bool real = chosen >= 0;
  const unsigned rm = __ballot_sync(FULL, real);
  if (real) {
    const unsigned peers = __match_any_sync(rm, chosen);
    const long long q = __double2ll_rn(dE * TALLY_SCALE);
    const unsigned long long g = (q > 0) ? static_cast<unsigned long long>(q) : 0ull, l = (q < 0) ? static_cast<unsigned long long>(-q) : 0ull;
    // a side nobody in the warp contributes to (gains, in a batch of cold-gas collisions without superelastics) is skipped for the whole warp
    const bool any_gain = __any_sync(rm, g != 0), any_loss = __any_sync(rm, l != 0);
    unsigned g0 = 0, g1 = 0, g2 = 0, l0 = 0, l1 = 0, l2 = 0;
    if (any_gain) {
      g0 = __reduce_add_sync(peers, static_cast<unsigned>(g & 0x1FFFFFu)); g1 = __reduce_add_sync(peers, static_cast<unsigned>((g >> 21) & 0x1FFFFFu));
      g2 = __reduce_add_sync(peers, static_cast<unsigned>(g >> 42));
    }
    if (any_loss) {
      l0 = __reduce_add_sync(peers, static_cast<unsigned>(l & 0x1FFFFFu)); l1 = __reduce_add_sync(peers, static_cast<unsigned>((l >> 21) & 0x1FFFFFu));
      l2 = __reduce_add_sync(peers, static_cast<unsigned>(l >> 42));
    }
    if (lane == __ffs(peers) - 1) {
      Tally* t = s_tally + chosen;
      atomicAdd(&t->cnt, static_cast<unsigned int>(__popc(peers)));
      const unsigned long long gs = static_cast<unsigned long long>(g0) + (static_cast<unsigned long long>(g1) << 21) + (static_cast<unsigned long long>(g2) << 42);
      const unsigned long long ls = static_cast<unsigned long long>(l0) + (static_cast<unsigned long long>(l1) << 21) + (static_cast<unsigned long long>(l2) << 42);
      if (gs) atomicAdd(&t->gain, gs);
      if (ls) atomicAdd(&t->loss, ls);
    }
  }
}

// branch-free null test (BMC.C:1035-1053, like cold_null_test)
__device__ __forceinline__ bool stream_null_test(const Model& m, const double* s_nu, int staged_rows, double eps, double nue, double u, double& Rnu, bool& clamped, bool& exceeded) {
  Rnu = nue * u;
  int i1, i2; double w1, w2;
  cold_rows(m, eps, i1, i2, w1, w2);
  clamped = (i1 == m.nE - 1);
  // the staged rows are read from shared memory, the rest from global: one generic load per row, no branch (the block stays straight-line)
  const bool in_smem = i2 < staged_rows;
  const double* p1 = in_smem ? s_nu + i1 : m.nu_tot + i1;
  const double* p2 = in_smem ? s_nu + i2 : m.nu_tot + i2;
  const double nu_here = w1 * *p1 + w2 * *p2;
  exceeded = nu_here > nue;
  return !(Rnu > nu_here);                                         // BMC.C:1050
}

__device__ __forceinline__ int tid_now() { int t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }   // re-read, never spilled

// (Energy maxima are kept per THREAD in shared memory and reduced once at the end.  A warp-level REDUX.MAX folded into a shared slot was
// cheaper on paper, but ptxas 12.9 allocated its uniform destination register on top of the live shared-memory base in two of the
// fifteen instantiations of this kernel (ECR and AC+B, cold gas): tools/check_ur_clobber.py, tests/test_host_abi.py.)

struct Flyer {   // one electron in the flight phase (two per lane)
  Particle p;
  unsigned long long id;
  double gain;
  unsigned int used;
  int sl;
  unsigned char outcome;
  bool active, need, clamped, exceeded, virt;
};

enum : int { MC_BORN = 0, MC_ATT, MC_CLAMP, MC_NUEX, MC_VIRT, MC_COUNT };   // rare events: counted with shared atomics, not in registers
// CTA-uniform state of a round lives in shared memory and is re-read where it is used: with 128 registers per thread every value that stays
// live across the inlined collision code is a spill candidate, and local-memory spills miss the small L1 (profiles/r1_v10_*)
enum : int { RS_IN = 0, RS_OUT, RS_LEN, RS_NFL, RS_NBC, RS_NBT, RS_NRET, RS_NREFILL, RS_COUNT };   // (the 64-bit ones, range start and flight count, sit in s_scan[8], s_scan[9])

template <int FIELD, int GT, bool SAMPLE>
__global__ void __launch_bounds__(STREAM_THREADS, 2) k_advance_stream(const Model m, const StateId sid, const Lists L, const Pending pend, const AdvArgs a,
                                                                       const HistGrid h, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* col = reinterpret_cast<double*>(smem_raw + SM_COL);                               // [SC_COLS][POOL]
  double* s_hdr = reinterpret_cast<double*>(smem_raw + SM_HDR);                             // [R_HEADER]
  double* s_gf = reinterpret_cast<double*>(smem_raw + SM_GF);                               // [STREAM_THREADS] field gain, one slot per thread
  double* s_tmax = reinterpret_cast<double*>(smem_raw + SM_TMAX);                           // [2][STREAM_THREADS] max energy at t_sync / at any event
  unsigned long long* s_scan = reinterpret_cast<unsigned long long*>(smem_raw + SM_SCAN);   // [16]
  unsigned int* s_used = reinterpret_cast<unsigned int*>(smem_raw + SM_USED);               // [POOL]
  unsigned short* listF = reinterpret_cast<unsigned short*>(smem_raw + SM_LISTS);           // [POOL] continuing flights of this round
  unsigned short* listR = listF + POOL;                                                     // [POOL] collisions of this round: cold first, thermal after
  unsigned short* listO = listR + POOL;                                                     // [POOL] slots retiring this round (output order; bit 15: attached)
  unsigned short* listE = listO + POOL;                                                     // [POOL] empty slots, in slot order
  unsigned char* flag = smem_raw + SM_FLAG;                                                 // [POOL]
  unsigned int* s_misc = reinterpret_cast<unsigned int*>(smem_raw + SM_MISC);               // [MC_COUNT]
  volatile int* s_rs = reinterpret_cast<volatile int*>(smem_raw + SM_RS);                   // [RS_COUNT] round state
  const double* s_nu = reinterpret_cast<const double*>(smem_raw + SM_NU);                   // [a.pad] first rows of nu_tot (a.pad = 0: not staged)
  Tally* s_tally = reinterpret_cast<Tally*>(smem_raw + (a.pad ? SM_TALLY_STAGED : SM_TALLY_PLAIN));   // [P]

  {
  const int tid = tid_now();
  for (int k = tid; k < m.P; k += STREAM_THREADS) { s_tally[k].gain = 0; s_tally[k].loss = 0; s_tally[k].cnt = 0; }
  if (tid < R_HEADER) s_hdr[tid] = 0;
  s_gf[tid] = 0;
  s_tmax[tid] = 0; s_tmax[STREAM_THREADS + tid] = 0;
  if (tid < MC_COUNT) s_misc[tid] = 0;
  for (int j = tid; j < static_cast<int>(a.pad); j += STREAM_THREADS) reinterpret_cast<double*>(smem_raw + SM_NU)[j] = __ldg(&m.nu_tot[j]);
  reinterpret_cast<unsigned int*>(flag)[tid] = 0u;   // all slots FL_EMPTY
#ifndef LK_NO_DUMMY_SLOT
  if (tid < SC_COLS) col[tid * PSTRIDE + POOL] = 0.0;   // the dummy slot (never written again)
  if (tid == 0) { s_used[POOL] = 0u; flag[POOL] = FL_EMPTY; }
#endif

  // the CTA's range [lo, lo + len) of the ensemble; cursors are CTA-uniform offsets into it.  Column c of the state is sid.s.x + lo + c * a.n.
  if (tid == 0) {
    const long long chunk = (((a.n + gridDim.x - 1) / gridDim.x) + 31) & ~31ll;
    const long long lo = min(static_cast<long long>(blockIdx.x) * chunk, a.n);
    s_rs[RS_IN] = 0; s_rs[RS_OUT] = 0; s_rs[RS_LEN] = static_cast<int>(min(lo + chunk, a.n) - lo);
    *reinterpret_cast<volatile long long*>(&s_scan[8]) = lo;
    *reinterpret_cast<volatile unsigned long long*>(&s_scan[9]) = 0ull;
  }
  }
  __syncthreads();

#pragma unroll 1
  for (;;) {
    // ================= (1) block scan over the slot flags: build the lists of this round in slot order =================
    const int tid = tid_now(), lane = tid & 31, warp = tid >> 5;
    const unsigned int f4 = reinterpret_cast<const unsigned int*>(flag)[tid];
    unsigned int cFl = 0, cRc = 0, cRt = 0, cRet = 0, cEmp = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      cFl += (f == FL_FLIGHT); cRc += (f == FL_REAL); cRt += (f == FL_REALT); cRet += (f == FL_DONE || f == FL_DEAD); cEmp += (f == FL_EMPTY);
    }
    const unsigned long long mine = static_cast<unsigned long long>(cFl) | (static_cast<unsigned long long>(cRc) << 12) | (static_cast<unsigned long long>(cRt) << 24) |
                                    (static_cast<unsigned long long>(cRet) << 36) | (static_cast<unsigned long long>(cEmp) << 48);
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
              nRet = static_cast<int>((total >> 36) & 0xFFFu), nEmp = static_cast<int>((total >> 48) & 0xFFFu);
    const int in_off = s_rs[RS_IN], out_off = s_rs[RS_OUT];
    // refill rank r: input element in_off + r goes to the r-th retiring slot (r < nRet), then to the empty slots in slot order
    const int nRefill = min(nRet + nEmp, s_rs[RS_LEN] - in_off);
    // collisions are run in whole CTA-iterations; everything parked is flushed when the flights alone cannot keep the CTA busy
    const bool flush = (nFl + nRefill < STREAM_THREADS);
    const int nBc = flush ? nRc : (nRc / STREAM_THREADS) * STREAM_THREADS, nBt = flush ? nRt : (nRt / STREAM_THREADS) * STREAM_THREADS;
    int eFl = static_cast<int>(excl & 0xFFFu), eRc = static_cast<int>((excl >> 12) & 0xFFFu), eRt = static_cast<int>((excl >> 24) & 0xFFFu),
        eRet = static_cast<int>((excl >> 36) & 0xFFFu), eEmp = static_cast<int>((excl >> 48) & 0xFFFu);
    unsigned int new_f4 = f4;
    // branch-free: the four lists are contiguous ([F | R | O | E], POOL entries each), so every slot computes ONE destination index
    // (POOL * 4 = nowhere) and does one predicated store; the five kinds would otherwise be five divergent paths per slot
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sl = tid * 4 + q;
      const unsigned int f = (f4 >> (8 * q)) & 0xFFu;
      const bool isRet = (f == FL_DONE || f == FL_DEAD), isEmp = (f == FL_EMPTY), isFl = (f == FL_FLIGHT), isRc = (f == FL_REAL), isRt = (f == FL_REALT);
      const bool refilled = isRet ? (eRet < nRefill) : (isEmp && nRet + eEmp < nRefill);
      int dst = 4 * POOL;
      dst = isFl ? eFl : dst;
      dst = (isRc && eRc < nBc) ? POOL + eRc : dst;
      dst = (isRt && eRt < nBt) ? POOL + nBc + eRt : dst;
      dst = isRet ? 2 * POOL + eRet : dst;
      dst = isEmp ? 3 * POOL + eEmp : dst;
      const unsigned int val = static_cast<unsigned int>(sl) | ((f == FL_DEAD) ? 0x8000u : 0u);
      if (dst < 4 * POOL) listF[dst] = static_cast<unsigned short>(val);
      eFl += isFl; eRc += isRc; eRt += isRt; eRet += isRet; eEmp += isEmp;
      const unsigned int nf = (isRet || isEmp) ? (refilled ? static_cast<unsigned int>(FL_FLIGHT) : static_cast<unsigned int>(FL_EMPTY)) : f;
      new_f4 = (new_f4 & ~(0xFFu << (8 * q))) | (nf << (8 * q));
    }
    reinterpret_cast<unsigned int*>(flag)[tid] = new_f4;
    const int nF = nFl + nRefill + nBc + nBt;
    if (tid == 0) {   // every electron in F flies once this round unless it attaches in (3): sum of nF - electrons = non-partial flights = real + null events
      s_rs[RS_NFL] = nFl; s_rs[RS_NBC] = nBc; s_rs[RS_NBT] = nBt; s_rs[RS_NRET] = nRet; s_rs[RS_NREFILL] = nRefill;
      *reinterpret_cast<volatile unsigned long long*>(&s_scan[9]) += static_cast<unsigned long long>(nF);
    }
    __syncthreads();
    if (nF == 0 && nRet == 0) break;                               // nothing in flight, nothing parked, nothing to write back or load
    const long long lo = *reinterpret_cast<volatile long long*>(&s_scan[8]);
    double* const g0 = sid.s.x + lo;
    unsigned long long* const gid = sid.id + lo;

    // ================= (2) retire + refill by rank: thread t owns ranks t, t + 256, ... =================
    // rank r < nRet: the r-th retiring slot is written to output element out_off + r; rank r < nRefill: input element in_off + r is
    // loaded into that same slot (or, past nRet, into an empty one).  One thread does both for a slot, in program order.
    {
      const int nOwn = max(nRet, nRefill);
      for (int rb = warp * 32; rb < nOwn; rb += STREAM_THREADS) {
        const int r = rb + lane;
        double eps_end = 0;
        if (r < nOwn) {
          const unsigned int e = (r < nRet) ? listO[r] : listE[r - nRet];
          const int sl = static_cast<int>(e & 0x7FFFu);
          if (r < nRet) {
            double* const gp = g0 + (out_off + r);
            const double vx = col[SC_VX * PSTRIDE + sl], vy = col[SC_VY * PSTRIDE + sl], vz = col[SC_VZ * PSTRIDE + sl];
#pragma unroll
            for (int c = 0; c < 8; ++c) __stcs(gp + c * a.n, col[c * PSTRIDE + sl]);
            gid[out_off + r] = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * PSTRIDE + sl]));
            if (e & 0x8000u) {                                         // attached: population control refills this position at t_sync
              const long long pos = lo + out_off + r;
              const unsigned int idx = atomicAdd(&L.counters[C_DEAD], 1u);
              if (idx < L.dead_cap) { L.dead[idx] = static_cast<unsigned int>(pos); L.dead_flag[pos] = 1; } else atomicExch(&L.counters[C_OVERFLOW], 1u);
            } else eps_end = kinetic_eV(vx, vy, vz);                   // the energy at t_sync (same bits as the flight computed, BMC.C:901)
          }
          if (r < nRefill) {
            const double* const gq = g0 + (in_off + r);
#pragma unroll
            for (int c = 0; c < 8; ++c) cp_async8(&col[c * PSTRIDE + sl], gq + c * a.n);
            cp_async8(&col[SC_ID * PSTRIDE + sl], &gid[in_off + r]);
            col[SC_T * PSTRIDE + sl] = a.t0; s_used[sl] = 0;
          }
        }
        s_tmax[tid] = fmax(s_tmax[tid], eps_end);
      }
    }
    cp_async_commit();
    if (tid == 0) { s_rs[RS_IN] = in_off + nRefill; s_rs[RS_OUT] = out_off + nRet; }   // read again after the barrier that ends the round
    {   // pull the next round's input lines towards L2 while this round computes
      const int ahead = in_off + nRefill + tid * 16;   // 16 doubles = one 128-byte line per thread and column
      if (ahead < s_rs[RS_LEN] && tid < 48) {
#pragma unroll
        for (int c = 0; c < 8; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(g0 + c * a.n + ahead));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&gid[ahead]));
      }
    }

    // ================= (3) collisions on the compacted list (BMC.C:916-1031, 1054-1280): cold-gas entries first, thermal-target after =================
    // chunk c belongs to warp c % 8, which also flies its electrons in (4): no CTA barrier between the two.  Both kinds share ONE
    // instance of the collision dynamics (the process selection differs, the scattering code does not).
    {
      const int nBcr = s_rs[RS_NBC], nB = nBcr + s_rs[RS_NBT];
#pragma unroll 1
      for (int chunk_i = warp; chunk_i * 32 < nB; chunk_i += STREAM_WARPS) {
        const int k = chunk_i * 32 + lane;
        int chosen = NOT_ADVANCED;
        double dE = 0, seen = 0;
        if (k < nB) {
          const int sl = listR[k];
          Particle p;
          p.x = col[SC_X * PSTRIDE + sl]; p.y = col[SC_Y * PSTRIDE + sl]; p.z = col[SC_Z * PSTRIDE + sl];
          p.vx = col[SC_VX * PSTRIDE + sl]; p.vy = col[SC_VY * PSTRIDE + sl]; p.vz = col[SC_VZ * PSTRIDE + sl];
          p.nue = col[SC_NUE * PSTRIDE + sl]; p.t = col[SC_T * PSTRIDE + sl]; p.tcf = NON_DEF;
          p.eps = kinetic_eV(p.vx, p.vy, p.vz);
          PhiloxRng rng;
          const unsigned long long id = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * PSTRIDE + sl]));
          rng.k0 = static_cast<uint32_t>(a.seed); rng.k1 = static_cast<uint32_t>(a.seed >> 32);
          rng.c0 = static_cast<uint32_t>(id); rng.c1 = static_cast<uint32_t>(id >> 32); rng.c2 = a.interval;
          rng.used = s_used[sl]; rng.blk = 0xFFFFFFFFu;
          EventOut o; o.table_clamped = 0; o.nu_exceeded = 0; o.dE = 0;
          double Vx = 0, Vy = 0, Vz = 0;
          if (GT != GT_FALSE && (GT == GT_TRUE || k >= nBcr)) chosen = thermal_select(m, p, rng, o, Vx, Vy, Vz);
          else chosen = cold_select(m, p, col[SC_TCF * PSTRIDE + sl]);
          if (chosen != NULL_COLLISION) chosen = collide_dynamics<GT>(m, chosen, p, Vx, Vy, Vz, rng, o);
          if (o.table_clamped) atomicAdd(&s_misc[MC_CLAMP], 1u);
            if (o.nu_exceeded) atomicAdd(&s_misc[MC_NUEX], 1u);
          unsigned char outcome = FL_FLIGHT;
          if (chosen >= 0) {
            dE = o.dE;
            const int type = __ldg(&m.type[chosen]);
            if (type == T_IONIZATION) {
              atomicAdd(&s_misc[MC_BORN], 1u);
              uint32_t cc1, ck1; child_stream(rng.c1, rng.k1, rng.used, cc1, ck1);
              push_pending(pend, L.counters, o, p.t, rng.c0, cc1, ck1);
            } else if (type == T_ATTACHMENT) { atomicAdd(&s_misc[MC_ATT], 1u); outcome = FL_DEAD; }
          }                                                        // aborted picks count as null collisions (BMC.C:1137-1140): see s_scan[9]
          seen = p.eps;
          col[SC_VX * PSTRIDE + sl] = p.vx; col[SC_VY * PSTRIDE + sl] = p.vy; col[SC_VZ * PSTRIDE + sl] = p.vz;
          col[SC_TCF * PSTRIDE + sl] = NON_DEF;                       // the next free time is drawn at the start of the flight (same stream position)
          s_used[sl] = rng.used;
          flag[sl] = outcome;
        }
        tally_collisions_fx(chosen, dE, s_tally, lane);
        s_tmax[STREAM_THREADS + tid] = fmax(s_tmax[STREAM_THREADS + tid], seen);
      }
    }
    __syncwarp();   // the warp's own collision results are visible to all its lanes

    // ================= (4) one free flight + null-collision test per electron (BMC.C:650-667, 804-905, 1035-1053) =================
    // Two electrons per lane, advanced side by side through one branch-free block: with four warps per scheduler the dependent chains
    // of a single electron (ten Philox rounds, the logarithm, the interpolated table test) leave the issue slots idle two cycles out of
    // three; two independent chains in the same basic block let ptxas interleave them.
    // A warp's work items, in order: the chunks it collided in (3), its share of the continuing electrons (chunk c of listF belongs to
    // warp (c + rot) % 8), and -- after its own cp.async copies have landed -- the electrons it refilled in (2).  Nothing in this list
    // was written by another warp in this round, so the round needs no barrier before its end.
    {
      const int nFlr = s_rs[RS_NFL], nB = s_rs[RS_NBC] + s_rs[RS_NBT], nRetr = s_rs[RS_NRET], nRefr = s_rs[RS_NREFILL];
      const int kc = (nB + 31) >> 5, cc = (nFlr + 31) >> 5, rc = (nRefr + 31) >> 5;                // 32-wide chunks: collided, continuing, refilled
      const int nK = (kc - warp + STREAM_WARPS - 1) >> 3;                                          // this warp's collided chunks (warp, warp + 8, ...)
      const int wrot = (warp - kc - rc) & (STREAM_WARPS - 1);                                      // continuing chunk c -> warp (c + kc + rc) mod 8: the round-robin goes on where the owner-bound chunks ended
      const int nC = (cc - wrot + STREAM_WARPS - 1) >> 3;
      const int nR = (rc - warp + STREAM_WARPS - 1) >> 3;                                          // refill ranks [32 g, 32 g + 32), g = warp, warp + 8, ...: issued by these very threads
      const int nA = nK + nC, nItems = nA + nR;
      const int staged_rows = static_cast<int>(a.pad);
      const double rnu = recip_for_div(a.nu_trial);
      bool waited = false;
      double gain_round = 0, seen_round = 0;   // folded into the per-thread shared slots once per round
#pragma unroll 1
      for (int it = 0; 2 * it < nItems; ++it) {
        if (!waited && 2 * it + 1 >= nA) { cp_async_wait_all(); waited = true; }   // this pair holds a refilled chunk: this thread's copies have landed (only it reads them)
        Flyer e[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          Flyer& f = e[j];
#ifdef LK_NO_DUMMY_SLOT
          int sl = 0; bool act = false;
#else
          int sl = POOL; bool act = false;   // no electron: the dummy slot
#endif
          const int item = 2 * it + j;
          if (item < nK) { const int k = 32 * (warp + 8 * item) + lane; if (k < nB) { sl = listR[k]; act = true; } }
          else if (item < nA) { const int k = 32 * (wrot + 8 * (item - nK)) + lane; if (k < nFlr) { sl = listF[k]; act = true; } }
          else if (item < nItems) { const int r = 32 * (warp + 8 * (item - nA)) + lane; if (r < nRefr) { sl = (r < nRetr) ? (listO[r] & 0x7FFF) : listE[r - nRetr]; act = true; } }
          f.sl = sl;
          f.active = act && (flag[sl] == FL_FLIGHT);                // (an electron attached in (3) stays FL_DEAD and retires in the next scan)
          f.p.x = col[SC_X * PSTRIDE + sl]; f.p.y = col[SC_Y * PSTRIDE + sl]; f.p.z = col[SC_Z * PSTRIDE + sl];
          f.p.vx = col[SC_VX * PSTRIDE + sl]; f.p.vy = col[SC_VY * PSTRIDE + sl]; f.p.vz = col[SC_VZ * PSTRIDE + sl];
          f.p.tcf = col[SC_TCF * PSTRIDE + sl]; f.p.nue = col[SC_NUE * PSTRIDE + sl]; f.p.t = col[SC_T * PSTRIDE + sl];
          f.id = static_cast<unsigned long long>(__double_as_longlong(col[SC_ID * PSTRIDE + sl]));
          f.used = s_used[sl];
          f.need = f.active && (f.p.tcf == NON_DEF);
          if (f.need) f.used = (f.used + 1u) & ~1u;                 // free-time draws start on an even index (PhiloxRng::align)
        }
        // a warp of refilled electrons (they come with the rest of their previous free time) skips the logarithm: warp-uniform branch
        const bool any_draw = __any_sync(FULL, e[0].need || e[1].need);
        double u0[2], u1[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {   // one convergent draw site: the free time (if needed) and the null-test uniform come from the same Philox block
          uint32_t o4[4];
          philox4x32_10(static_cast<uint32_t>(e[j].id), static_cast<uint32_t>(e[j].id >> 32), a.interval, e[j].used >> 1, static_cast<uint32_t>(a.seed),
                        static_cast<uint32_t>(a.seed >> 32), o4);
          u0[j] = u52(o4[1], o4[0]); u1[j] = u52(o4[3], o4[2]);
        }
        if (any_draw) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            double nu = a.nu_trial, r = rnu, tau = 0;
            if (m.banded) {   // fast mode (CTA-uniform branch): trial frequency and look-ahead time of the electron's energy band, see draw_free_time
              const int b = energy_band(kinetic_eV(e[j].p.vx, e[j].p.vy, e[j].p.vz));
              nu = m.band_nu[b]; tau = m.band_tau[b]; r = recip_for_div(nu);
            }
            const double drawn = div_by(-log_normal(u0[j]), nu, r);   // -log(u) / nu_trial, BMC.C:650-655
            const bool cut = m.banded && drawn > tau;
            if (e[j].need) { e[j].p.tcf = cut ? tau : drawn; e[j].p.nue = cut ? -nu : nu; ++e[j].used; }
          }
        }
        double seen = 0;
        bool rare = false;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          Flyer& f = e[j];
          Particle& p = f.p;
          p.eps = kinetic_eV(p.vx, p.vy, p.vz);
          const double u_null = (f.used & 1u) ? u1[j] : u0[j];
          const bool partial = (p.t + p.tcf > a.t_sync);             // BMC.C:657
          const double dt = partial ? (a.t_sync - p.t) : p.tcf;
          f.gain = flight<FIELD>(m, p, dt);                          // one flight site for both outcomes (BMC.C:659, :666)
          const bool thermal = thermal_branch<GT>(m, p.eps);
          double Rnu; bool clamped, exceeded;
          const bool real = stream_null_test(m, s_nu, staged_rows, p.eps, p.nue, u_null, Rnu, clamped, exceeded);
          const bool virt = p.nue < 0;                               // fast mode: the flight was cut at the band's look-ahead time, nothing is tested
          const bool tested = !partial && !thermal && !virt;         // the thermal-target branch draws its own numbers in (3)
          f.outcome = partial ? FL_DONE : virt ? FL_FLIGHT : thermal ? FL_REALT : real ? FL_REAL : FL_FLIGHT;
          const double t_event = p.t + p.tcf;
          p.tcf = partial ? (p.tcf - dt) : (tested && real) ? Rnu : NON_DEF;
          p.t = partial ? a.t_sync : t_event;
          f.used += tested ? 1u : 0u;
          f.clamped = f.active && tested && clamped; f.exceeded = f.active && tested && exceeded;
          f.virt = f.active && virt && !partial;
          rare = rare || f.clamped || f.exceeded || f.virt;
          if (f.active) seen = fmax(seen, p.eps);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const Flyer& f = e[j];
          if (f.active) {
            const int sl = f.sl;
            gain_round += f.gain;
            col[SC_X * PSTRIDE + sl] = f.p.x; col[SC_Y * PSTRIDE + sl] = f.p.y; col[SC_Z * PSTRIDE + sl] = f.p.z;
            col[SC_VX * PSTRIDE + sl] = f.p.vx; col[SC_VY * PSTRIDE + sl] = f.p.vy; col[SC_VZ * PSTRIDE + sl] = f.p.vz;
            col[SC_TCF * PSTRIDE + sl] = f.p.tcf; col[SC_NUE * PSTRIDE + sl] = f.p.nue; col[SC_T * PSTRIDE + sl] = f.p.t;
            s_used[sl] = f.used;
            flag[sl] = f.outcome;
          }
        }
        if (rare) {
#pragma unroll
          for (int j = 0; j < 2; ++j) { if (e[j].clamped) atomicAdd(&s_misc[MC_CLAMP], 1u); if (e[j].exceeded) atomicAdd(&s_misc[MC_NUEX], 1u); if (e[j].virt) atomicAdd(&s_misc[MC_VIRT], 1u); }
        }
        seen_round = fmax(seen_round, seen);
      }
      s_gf[tid] += gain_round;
      s_tmax[STREAM_THREADS + tid] = fmax(s_tmax[STREAM_THREADS + tid], seen_round);
      if (!waited) cp_async_wait_all();
    }
    __syncthreads();
  }

  const int tid = tid_now();
  if (tid == 0) {   // header of the CTA: real collisions = sum of the per-process counts; rare-event counters; fixed-order sums
    double nr = 0;
    for (int k = 0; k < m.P; ++k) nr += static_cast<double>(s_tally[k].cnt);
    s_hdr[R_N_REAL] = nr;
    // events = non-partial flights = (flights flown) - (electrons of the range); null = events - real  (BMC.C:1308-1320)
    const unsigned long long flights = *reinterpret_cast<volatile unsigned long long*>(&s_scan[9]) - s_misc[MC_ATT];
    const unsigned long long partials_n = static_cast<unsigned long long>(s_rs[RS_LEN]) - s_misc[MC_ATT];
    s_hdr[R_N_NULL] = static_cast<double>(flights - partials_n) - nr - static_cast<double>(s_misc[MC_VIRT]);   // (cut flights of the fast mode are not trial events)
    s_hdr[R_N_BORN] = static_cast<double>(s_misc[MC_BORN]); s_hdr[R_N_ATTACHED] = static_cast<double>(s_misc[MC_ATT]);
    s_hdr[R_N_TABLE_CLAMPED] = static_cast<double>(s_misc[MC_CLAMP]); s_hdr[R_N_NU_EXCEEDED] = static_cast<double>(s_misc[MC_NUEX]);
    double gf = 0, m0 = 0, m1 = 0;
    for (int w = 0; w < STREAM_WARPS; ++w) {   // same order as a warp-shuffle tree over lanes followed by a sum over warps would not be needed: any fixed order is reproducible
      double ws = 0;
      for (int l = 0; l < 32; ++l) { ws += s_gf[w * 32 + l]; m0 = fmax(m0, s_tmax[w * 32 + l]); m1 = fmax(m1, s_tmax[STREAM_THREADS + w * 32 + l]); }
      gf += ws;
    }
    s_hdr[R_GAIN_FIELD] = gf; s_hdr[R_MAX_EPS] = m0; s_hdr[R_MAX_EPS_SEEN] = fmax(m0, m1);
  }
  __syncthreads();
  {   // partials: header, per-process tallies converted from fixed point
    const int plen = R_HEADER + 3 * m.P;
    double* out = partials + static_cast<size_t>(blockIdx.x) * plen;
    for (int j = tid; j < R_HEADER; j += STREAM_THREADS) out[j] = s_hdr[j];
    for (int k = tid; k < m.P; k += STREAM_THREADS) {
      out[R_HEADER + k] = static_cast<double>(s_tally[k].cnt);
      out[R_HEADER + m.P + k] = static_cast<double>(s_tally[k].gain) / TALLY_SCALE;
      out[R_HEADER + 2 * m.P + k] = -static_cast<double>(s_tally[k].loss) / TALLY_SCALE;
    }
  }
}

}  // namespace lk
