// lk_tile.cuh -- pieces shared by the pool kernels: slot flags, the pending list of electrons born inside an interval, and the
// births pass (k_advance_births).  (The first shared-memory form of K1, a CTA tile drained to empty, lived here; it was measured at
// 2.84 ms per interval against 1.31 ms for the streaming pool of lk_stream.cuh -- profiles/r1_v2_* -- and has been removed.)
#pragma once
#include "lk_kernels.cuh"

namespace lk {

enum : unsigned char { FL_EMPTY = 0, FL_FLIGHT = 1, FL_REAL = 2, FL_DONE = 3, FL_DEAD = 4 };
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
