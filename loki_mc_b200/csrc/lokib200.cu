// lokib200.cu -- engine and C ABI (include/lokib200.h) of the B200-native electron Monte Carlo hot path.
//
// Host side of the path: owns device memory, flattens cross sections into the device tables (restating
// BoltzmannMC::interpolateCrossSections, BMC.C:561-615), mirrors the scalar trial-frequency logic (BMC.C:716-802) and
// launches the kernels of lk_kernels.cuh.  There is NO CPU fallback: every entry point fails with LOKIB200_ERR_NO_DEVICE
// when no CUDA device is usable.
#include "../../include/lokib200.h"

#include <atomic>
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lk_stream.cuh"

using namespace lk;

static_assert(LOKIB200_R_N_REAL == R_N_REAL && LOKIB200_R_GROWTH == R_GROWTH && LOKIB200_R_SUM_EPS == R_SUM_EPS && LOKIB200_R_SUM_RR == R_SUM_RR &&
              LOKIB200_R_SUM_RV == R_SUM_RV && LOKIB200_R_N_SAMPLED == R_N_SAMPLED && LOKIB200_R_MAX_EPS == R_MAX_EPS &&
              LOKIB200_R_MAX_EPS_SEEN == R_MAX_EPS_SEEN && LOKIB200_R_HEADER == R_HEADER && LOKIB200_R_SUM_COUNT == R_SUM_COUNT &&
              LOKIB200_R_N_TABLE_CLAMPED == R_N_TABLE_CLAMPED && LOKIB200_R_N_NU_EXCEEDED == R_N_NU_EXCEEDED && LOKIB200_R_OVERFLOW == R_OVERFLOW,
              "result layout");
static_assert(sizeof(lokib200_electron) == sizeof(ElectronIO) && sizeof(lokib200_event_out) == sizeof(EventIO), "parity structs");

static thread_local std::string g_create_error;   // engines are driven from several host threads (concurrent jobs of a sweep)

struct lokib200_engine {
  lokib200_config cfg{};
  std::string err;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;

  // process set (host copies)
  int P = 0, nG = 0;
  bool have_processes = false, has_pc = false;
  std::vector<int> type, superel, angular, gas_first, gas_last;
  std::vector<double> ap0, ap1, swf, emin, emax, reldens, mass, redmass, eloss, thstd, wpar, gas_fraction, xs_e, xs_v;
  std::vector<int64_t> xs_off;
  // device copies
  int *d_type = nullptr, *d_angular = nullptr, *d_gas_first = nullptr, *d_gas_last = nullptr;
  double *d_ap0 = nullptr, *d_ap1 = nullptr, *d_mass = nullptr, *d_redmass = nullptr, *d_eloss = nullptr, *d_thstd = nullptr, *d_wpar = nullptr,
         *d_gas_fraction = nullptr;

  // tables
  int nE = 0, stride = 0;
  double dE = 0, maxE = LOKIB200_NON_DEF;
  bool have_tables = false;
  std::vector<double> h_cum, h_nu_tot, h_nu_max;   // host copies (h_cum padded to stride)
  double *d_cum = nullptr, *d_nu_tot = nullptr;
  double2* d_pair = nullptr;   // row-pair form of the cumulative table (see lk_physics.cuh)
  std::vector<double2> h_pair;
  // the three tables live in ONE allocation (row pairs | cumulative rows | nu_tot) so that a single L2 access-policy window can pin them
  unsigned char* d_tables = nullptr;
  size_t d_tables_cap = 0;   // bytes
  size_t l2_persist_max = 0, l2_window_max = 0;   // device limits (lokib200_create)
  const void* l2_window_base = nullptr; size_t l2_window_bytes = 0; cudaStream_t l2_window_stream = nullptr;   // the window currently set

  // ensemble
  State st{};
  double* d_state = nullptr;
  unsigned long long* d_id = nullptr;   // electron id of each slot (the stream kernel permutes electrons inside a CTA's range)
  bool permuted = false;
  double time = 0;
  uint32_t interval = 0;
  Lists lists{};
  Pending pend{};
  double *d_adv_part = nullptr, *d_birth_part = nullptr, *d_smp_part = nullptr, *d_result = nullptr, *d_pc_result = nullptr, *h_result = nullptr;
  unsigned long long* d_maxbits = nullptr;
  int adv_blocks = 0, tile_blocks = 0, birth_blocks = 0, smp_blocks = 0, part_len = 0;
  bool use_tile = false;   // a shared-memory kernel (stream or lane) instead of one electron per thread
  int last_adv_blocks = 0;

  // fast mode: trial frequency per energy band (lk_physics.cuh draw_free_time); the tables are rebuilt when the trial frequency or the cross-section tables change
  bool fast_mode = false;
  double band_nu[N_BANDS] = {}, band_tau[N_BANDS] = {}, band_for_nu = 0;
  uint64_t table_version = 0, band_for_version = ~0ull;

  // multi-GPU: communicator over the shards of one job (null = none)
  ncclComm_t comm = nullptr;
  int comm_size = 1, comm_rank = 0;
  bool comm_local_group = false;   // created by lokib200_comm_init_all: collectives must be issued for all local engines in one NCCL group
  // per-interval exchange over peer memory (k_exchange, lk_kernels.cuh): this rank's mailbox, the peers' mailboxes mapped into this process
  double* d_mail = nullptr;
  Mailboxes mail{};
  bool mail_ipc[EXCHANGE_MAX_RANKS] = {};   // box[r] was opened with cudaIpcOpenMemHandle
  int mail_stride = 0;
  unsigned long long exchange_epoch = 0;
  bool p2p = false;
  unsigned long long* d_hist_red = nullptr;   // all-reduced copy of the four histogram arrays (the accumulators keep this rank's own counts)
  bool hist_reduced = false;                  // d_hist_red holds the combined counts of the current accumulators

  // histograms
  HistGrid hist{};
  unsigned long long *d_eeh = nullptr, *d_eah = nullptr, *d_evh = nullptr, *d_eeh_per = nullptr;
  double max_eedf_energy = 0;

  // kernel timing
  // kernel timing: event pairs around the advance kernel of the first EV_CAP intervals after each lokib200_kernel_time_ms call (a job of 1e5
  // intervals must neither create 2e5 events nor pay two cudaEventRecord per interval)
  static constexpr size_t EV_CAP = 256;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  bool timing = true;

  // the blocking interval of a small ensemble as ONE CUDA graph launch (advance_graph below): index = sample flag
  struct IntervalGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t k1 = nullptr, births = nullptr, pc_copy = nullptr, pc_lottery = nullptr, pc_small = nullptr, hist = nullptr, exchange = nullptr;   // the nodes whose arguments change from interval to interval
    cudaKernelNodeParams k1_p{}, births_p{}, copy_p{}, lot_p{}, small_p{}, hist_p{}, exchange_p{};
    int kernels = 0;
  } ig[4];   // index = sample flag + 2 * (the deferred histogram pass of the previous sample comes first)
  bool graph_off = false;
  // lokib200_sample_histograms of a graph-driven engine is DEFERRED: the pass reads the ensemble, which nothing touches before the next interval,
  // so it becomes the first node of that interval's graph (one launch less per sample); every call that reads the histograms, changes their
  // grid or changes the ensemble by another route flushes it first (flush_pending_hist)
  bool hist_pending = false;
  int hist_pending_phase = -1;

  // host-side time split of the blocking interval (LOKIB200_PROFILE=1 prints it when the engine is destroyed)
  double prof_submit = 0, prof_wait = 0, prof_outside = 0, prof_last_return = 0;
  int64_t prof_calls = 0;
};

namespace {

void drop_graph(lokib200_engine::IntervalGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.graph) cudaGraphDestroy(g.graph);
  g = lokib200_engine::IntervalGraph{};
}

#define CK(call)                                                                                                         \
  do {                                                                                                                   \
    cudaError_t e_ = (call);                                                                                             \
    if (e_ != cudaSuccess) {                                                                                             \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                                       \
      return LOKIB200_ERR_CUDA;                                                                                          \
    }                                                                                                                    \
  } while (0)

int fail(lokib200_engine* h, int code, const std::string& msg) { if (h) h->err = msg; return code; }

template <class T>
int upload(lokib200_engine* h, T** dptr, const std::vector<T>& v) {
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  CK(cudaMalloc(dptr, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int field_case(const lokib200_config& c) {   // the branch structure of accelerateElectron (BMC.C:811-898)
  const double w = c.excitation_omega, W = c.cyclotron_omega;
  if (W == 0) return (w == 0) ? F_DC : F_AC;
  if (w == 0) return F_DCB;
  if (std::fabs(w - W) / w < 1E-6) return F_ECR;
  return F_ACB;
}

Model make_model(const lokib200_engine* h) {
  Model m{};
  const lokib200_config& c = h->cfg;
  m.P = h->P; m.stride = h->stride; m.nG = h->nG; m.nE = h->nE;
  m.sharing = c.ionization_sharing; m.sharing_factor = c.energy_sharing_factor;
  m.Ngas = c.gas_density; m.dE = h->dE; m.inv_dE = (h->dE > 0) ? 1.0 / h->dE : 0.0;
  m.smart_limit = 20.0 * (1.5 * KB * c.gas_temperature / QE);            // BMC.C:433, :916
  const double Ex = c.electric_field[0], Ez = c.electric_field[2], w = c.excitation_omega, W = c.cyclotron_omega;
  m.Ex = Ex; m.Ez = Ez; m.w = w; m.W = W;
  m.aEx = -QE / ME * c.electric_field[0]; m.aEy = -QE / ME * c.electric_field[1]; m.aEz = -QE / ME * c.electric_field[2];   // BMC.C:489
  if (w != 0) { m.ac_e_me_w = QE / (ME * w); m.ac_e_me_w_w = m.ac_e_me_w / w; }
  if (W != 0) {
    m.dcb_vEx = QE * Ex / (ME * W); m.dcb_az = QE * Ez / ME; m.dcb_half_az = 0.5 * m.dcb_az;
    const double e_me_W = QE / (ME * W);
    m.ecr_vEx = e_me_W * Ex; m.ecr_vEz = e_me_W * Ez;
    if (w != 0) {
      const double w2 = w * w, W2 = W * W, d = w2 - W2;
      const double vEx = QE / (ME * W) * Ex, vEx_d = vEx / d, WvEx_d = W * vEx_d;
      m.acb_vEz = QE / (ME * w) * Ez; m.acb_WvEx_d = WvEx_d; m.acb_vEx_d_w = vEx_d / w; m.acb_WvEx_d_w = WvEx_d * w; m.acb_vEx_d_W2 = vEx_d * W2;
      m.acb_w2 = w2; m.acb_W2 = W2;
    }
  }
  m.cum = h->d_cum; m.nu_tot = h->d_nu_tot;
  m.pair = h->d_pair;
  m.banded = h->fast_mode ? 1 : 0;
  if (h->fast_mode) { std::memcpy(m.band_nu, h->band_nu, sizeof(m.band_nu)); std::memcpy(m.band_tau, h->band_tau, sizeof(m.band_tau)); }
  m.type = h->d_type; m.angular = h->d_angular; m.ap0 = h->d_ap0; m.ap1 = h->d_ap1; m.mass = h->d_mass; m.redmass = h->d_redmass;
  m.eloss = h->d_eloss; m.thstd = h->d_thstd; m.wpar = h->d_wpar; m.gas_first = h->d_gas_first; m.gas_last = h->d_gas_last;
  m.gas_fraction = h->d_gas_fraction;
  return m;
}

// GSL gsl_interp_linear semantics (the reference evaluates cross sections with gsl_spline_eval, BMC.C:589,598)
double lin_interp(const double* x, const double* y, int64_t n, double xv) {
  int64_t lo = 0, hi = n - 1;
  while (hi > lo + 1) { const int64_t mid = (hi + lo) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
  const double dx = x[lo + 1] - x[lo];
  return y[lo] + (xv - x[lo]) / dx * (y[lo + 1] - y[lo]);
}

size_t adv_smem_bytes(const lokib200_engine* h, bool sample) {
  return static_cast<size_t>(h->P) * (8 + 8 + 4) + ((sample && h->hist.enabled) ? (static_cast<size_t>(h->hist.nEn) + hist_tile_words(h->hist)) * 4 : 0) + 16;
}

template <int F, int G, bool S>
int launch_advance_t(lokib200_engine* h, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  const size_t smem = adv_smem_bytes(h, S);
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_advance<F, G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  k_advance<F, G, S><<<h->adv_blocks, ADV_THREADS, smem, h->stream>>>(m, h->st, h->lists, a, hg, h->d_adv_part);
  return 0;
}
template <int F, int G>
int launch_advance_s(lokib200_engine* h, bool sample, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  return sample ? launch_advance_t<F, G, true>(h, m, a, hg) : launch_advance_t<F, G, false>(h, m, a, hg);
}
template <int F>
int launch_advance_g(lokib200_engine* h, int gt, bool sample, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  switch (gt) {
#ifndef LK_BENCH_ONLY
    case GT_FALSE: return launch_advance_s<F, GT_FALSE>(h, sample, m, a, hg);
    case GT_TRUE: return launch_advance_s<F, GT_TRUE>(h, sample, m, a, hg);
#endif
    default: return launch_advance_s<F, GT_SMART>(h, sample, m, a, hg);
  }
}
int launch_advance(lokib200_engine* h, bool sample, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  switch (field_case(h->cfg)) {
    case F_DC: return launch_advance_g<F_DC>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
#ifndef LK_BENCH_ONLY
    case F_AC: return launch_advance_g<F_AC>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
    case F_DCB: return launch_advance_g<F_DCB>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
    case F_ECR: return launch_advance_g<F_ECR>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
    case F_ACB: return launch_advance_g<F_ACB>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
#endif
    default: return launch_advance_g<F_DC>(h, h->cfg.gas_temperature_effect, sample, m, a, hg);
  }
}

template <int F, int G>
int launch_stream_t(lokib200_engine* h, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  const size_t smem = stream_smem_bytes(h->P);
  AdvArgs as = a;
  as.pad = stream_stages_nu(h->P) ? static_cast<unsigned int>(std::min(h->nE, NU_STAGE_ROWS)) : 0u;   // rows of nu_tot the kernel stages in shared memory
  // the limit belongs to the kernel and the device, not to the engine (two engines with different P share one instantiation): raise it to the
  // budget once per instantiation and device, never lower it
  static std::atomic<bool> attr_set[64];
  if (!attr_set[h->cfg.device & 63]) {
    CK(cudaFuncSetAttribute(k_advance_stream<F, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(STREAM_SMEM_BUDGET)));
    CK(cudaFuncSetAttribute(k_advance_stream<F, G, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[h->cfg.device & 63] = true;
  }
  const StateId sid{h->st, h->d_id};   // the kernel addresses column c as st.x + c * n (one allocation, lokib200_create)
  k_advance_stream<F, G, false><<<h->tile_blocks, STREAM_THREADS, smem, h->stream>>>(m, sid, h->lists, h->pend, as, hg, h->d_adv_part);
  return 0;
}
template <int F>
int launch_stream_g(lokib200_engine* h, int gt, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  switch (gt) {
#ifndef LK_BENCH_ONLY
    case GT_FALSE: return launch_stream_t<F, GT_FALSE>(h, m, a, hg);
    case GT_TRUE: return launch_stream_t<F, GT_TRUE>(h, m, a, hg);
#endif
    default: return launch_stream_t<F, GT_SMART>(h, m, a, hg);
  }
}
int launch_stream(lokib200_engine* h, const Model& m, const AdvArgs& a, const HistGrid& hg) {
  switch (field_case(h->cfg)) {
    case F_DC: return launch_stream_g<F_DC>(h, h->cfg.gas_temperature_effect, m, a, hg);
#ifndef LK_BENCH_ONLY
    case F_AC: return launch_stream_g<F_AC>(h, h->cfg.gas_temperature_effect, m, a, hg);
    case F_DCB: return launch_stream_g<F_DCB>(h, h->cfg.gas_temperature_effect, m, a, hg);
    case F_ECR: return launch_stream_g<F_ECR>(h, h->cfg.gas_temperature_effect, m, a, hg);
    case F_ACB: return launch_stream_g<F_ACB>(h, h->cfg.gas_temperature_effect, m, a, hg);
#endif
    default: return launch_stream_g<F_DC>(h, h->cfg.gas_temperature_effect, m, a, hg);
  }
}

template <int F>
void launch_births_g(lokib200_engine* h, int gt, const Model& m, const AdvArgs& a) {
  const size_t smem = static_cast<size_t>(h->P) * 20 + 16;
  switch (gt) {
#ifndef LK_BENCH_ONLY
    case GT_FALSE: k_advance_births<F, GT_FALSE><<<h->birth_blocks, 128, smem, h->stream>>>(m, h->lists, h->pend, a, h->d_birth_part); break;
    case GT_TRUE: k_advance_births<F, GT_TRUE><<<h->birth_blocks, 128, smem, h->stream>>>(m, h->lists, h->pend, a, h->d_birth_part); break;
#endif
    default: k_advance_births<F, GT_SMART><<<h->birth_blocks, 128, smem, h->stream>>>(m, h->lists, h->pend, a, h->d_birth_part); break;
  }
}
void launch_births(lokib200_engine* h, const Model& m, const AdvArgs& a) {
  const int gt = h->cfg.gas_temperature_effect;
  switch (field_case(h->cfg)) {
    case F_DC: launch_births_g<F_DC>(h, gt, m, a); break;
#ifndef LK_BENCH_ONLY
    case F_AC: launch_births_g<F_AC>(h, gt, m, a); break;
    case F_DCB: launch_births_g<F_DCB>(h, gt, m, a); break;
    case F_ECR: launch_births_g<F_ECR>(h, gt, m, a); break;
    case F_ACB: launch_births_g<F_ACB>(h, gt, m, a); break;
#endif
    default: launch_births_g<F_DC>(h, gt, m, a); break;
  }
}

template <int F>
void launch_injected_g(lokib200_engine* h, int gt, const Model& m, int n, const ElectronIO* in, double nu, const double* ts, const double* dr, int nd,
                       ElectronIO* out, EventIO* ev) {
  const int blocks = (n + 127) / 128;
  switch (gt) {
#ifndef LK_BENCH_ONLY
    case GT_FALSE: k_step_injected<F, GT_FALSE><<<blocks, 128, 0, h->stream>>>(m, n, in, nu, ts, dr, nd, out, ev); break;
    case GT_TRUE: k_step_injected<F, GT_TRUE><<<blocks, 128, 0, h->stream>>>(m, n, in, nu, ts, dr, nd, out, ev); break;
#endif
    default: k_step_injected<F, GT_SMART><<<blocks, 128, 0, h->stream>>>(m, n, in, nu, ts, dr, nd, out, ev); break;
  }
}

// NCCL is bound at run time: a host process that already carries a copy (PyTorch ships its own libnccl.so.2) must see ONE instance
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string why;
  bool ok = false;
};
NcclApi load_nccl() {
  NcclApi api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
  auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) api.why = std::string("libnccl misses ") + n; return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
  return api;
}
NcclApi* nccl_api() {
  static NcclApi api = load_nccl();   // initialised once, thread-safe
  return &api;
}
#define NK(call)                                                                                                         \
  do {                                                                                                                   \
    ncclResult_t r_ = (call);                                                                                            \
    if (r_ != ncclSuccess) {                                                                                             \
      h->err = std::string(#call) + ": " + nc->GetErrorString(r_);                                                       \
      return LOKIB200_ERR_CUDA;                                                                                          \
    }                                                                                                                    \
  } while (0)

int ensure_ready(lokib200_engine* h, bool need_tables) {
  if (!h) return LOKIB200_ERR_INVALID;
  if (!h->have_processes) return fail(h, LOKIB200_ERR_INVALID, "lokib200_set_processes has not been called");
  if (need_tables && !h->have_tables) return fail(h, LOKIB200_ERR_INVALID, "no tables: call lokib200_build_tables or lokib200_upload_tables first");
  return 0;
}

// The tables are the only data of K1 with reuse (every probe of the process search, ~25 MB for N2) while 1.4 GB of ensemble state stream
// through the same L2 per launch: an access-policy window on the engine's stream marks them persisting.  Measured: K1 1.319 -> 1.303 ms with
// the row-pair table alone.  LOKIB200_L2_PERSIST=0 switches it off.
void apply_l2_window(lokib200_engine* h) {
  const char* env = std::getenv("LOKIB200_L2_PERSIST");
  if ((env && env[0] == '0') || !h->d_tables || !h->stream || !h->l2_persist_max || !h->l2_window_max) return;
  const size_t window = std::min<size_t>({h->d_tables_cap, h->l2_persist_max, h->l2_window_max});
  size_t have = 0;
  if (cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize) == cudaSuccess && have < window) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, window);
  if (h->l2_window_base == h->d_tables && h->l2_window_bytes == window && h->l2_window_stream == h->stream) return;   // unchanged (a rebuild inside the same allocation)
  h->l2_window_base = h->d_tables; h->l2_window_bytes = window; h->l2_window_stream = h->stream;
  cudaStreamAttrValue attr{};
  attr.accessPolicyWindow.base_ptr = h->d_tables; attr.accessPolicyWindow.num_bytes = window; attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
  for (auto& g : h->ig) drop_graph(g);   // captured kernel nodes carry the window of the stream they were captured on
}

int push_tables(lokib200_engine* h) {   // h_cum (padded) / h_nu_tot -> device
  const size_t need = static_cast<size_t>(h->nE) * h->stride, need_nu = static_cast<size_t>(h->nE);
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  const size_t off_cum = up(need * sizeof(double2)), off_nu = off_cum + up(need * sizeof(double)), bytes = off_nu + up(need_nu * sizeof(double));
  if (bytes > h->d_tables_cap) {
    if (h->d_tables) cudaFree(h->d_tables);
    h->d_tables = nullptr; h->d_tables_cap = 0; h->d_pair = nullptr; h->d_cum = nullptr; h->d_nu_tot = nullptr;
    CK(cudaMalloc(&h->d_tables, bytes));
    h->d_tables_cap = bytes;
  }
  h->d_pair = reinterpret_cast<double2*>(h->d_tables);
  h->d_cum = reinterpret_cast<double*>(h->d_tables + off_cum);
  h->d_nu_tot = reinterpret_cast<double*>(h->d_tables + off_nu);
  // row-pair form, derived from the same doubles (no arithmetic: the kernels see identical table values)
  h->h_pair.resize(need);
  for (int i = 0; i < h->nE; ++i) {
    const double* r1 = h->h_cum.data() + static_cast<size_t>(i) * h->stride;
    const double* r2 = h->h_cum.data() + static_cast<size_t>(std::min(i + 1, h->nE - 1)) * h->stride;
    double2* pr = h->h_pair.data() + static_cast<size_t>(i) * h->stride;
    for (int k = 0; k < h->stride; ++k) pr[k] = make_double2(r1[k], r2[k]);
  }
  CK(cudaMemcpyAsync(h->d_pair, h->h_pair.data(), need * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_cum, h->h_cum.data(), need * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_nu_tot, h->h_nu_tot.data(), need_nu * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_tables = true;
  ++h->table_version;
  apply_l2_window(h);
  return 0;
}

}  // namespace

extern "C" {

static void build_bands(lokib200_engine* h, double nu_trial);
static int flush_pending_hist(lokib200_engine* h);
static void release_mailboxes(lokib200_engine* h);

int lokib200_abi_version(void) { return LOKIB200_ABI_VERSION; }

int lokib200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* lokib200_last_error(const lokib200_engine* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lokib200_create(const lokib200_config* cfg, lokib200_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return LOKIB200_ERR_INVALID; }
  *out = nullptr;
  if (cfg->n_electrons <= 0 || cfg->n_electrons >= (1ll << 32) - (1ll << 28)) { g_create_error = "n_electrons out of range (1 .. ~4.0e9 per GPU)"; return LOKIB200_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device || cfg->device < 0) {
    g_create_error = "no usable CUDA device (this engine has no CPU fallback)";
    return LOKIB200_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop{};
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return LOKIB200_ERR_CUDA; }
  if (prop.major < 10) { g_create_error = std::string("device '") + prop.name + "' is not sm_100-class; the kernels are built for sm_100a only"; return LOKIB200_ERR_NO_DEVICE; }
  auto* h = new lokib200_engine();
  h->cfg = *cfg;
  h->l2_persist_max = prop.persistingL2CacheMaxSize > 0 ? static_cast<size_t>(prop.persistingL2CacheMaxSize) : 0;
  h->l2_window_max = prop.accessPolicyMaxWindowSize > 0 ? static_cast<size_t>(prop.accessPolicyMaxWindowSize) : 0;
  if (h->cfg.n_interp_points <= 1) h->cfg.n_interp_points = 10000;
  if (h->cfg.n_energy_cells <= 0) h->cfg.n_energy_cells = 1000;
  if (h->cfg.n_cos_cells <= 0) h->cfg.n_cos_cells = 100;
  if (h->cfg.n_radial_cells <= 0) h->cfg.n_radial_cells = 200;
  if (h->cfg.n_axial_cells <= 0) h->cfg.n_axial_cells = 200;
  if (h->cfg.n_phases <= 0) h->cfg.n_phases = 100;
  h->sm_count = prop.multiProcessorCount;
  auto bail = [&](const char* what, cudaError_t e) { g_create_error = std::string(what) + ": " + cudaGetErrorString(e); lokib200_destroy(h); return LOKIB200_ERR_CUDA; };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  h->own_stream = true;
  const size_t n = static_cast<size_t>(cfg->n_electrons);
  if ((e = cudaMalloc(&h->d_state, 8 * n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc(state)", e);
  h->st = State{h->d_state, h->d_state + n, h->d_state + 2 * n, h->d_state + 3 * n, h->d_state + 4 * n, h->d_state + 5 * n, h->d_state + 6 * n, h->d_state + 7 * n};
  if ((e = cudaMalloc(&h->d_id, n * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc(id)", e);
  if ((e = cudaMalloc(&h->d_maxbits, sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&h->d_pc_result, 2 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemset(h->d_pc_result, 0, 2 * sizeof(double))) != cudaSuccess) return bail("cudaMemset", e);
  *out = h;
  return LOKIB200_OK;
}

void lokib200_destroy(lokib200_engine* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) { NcclApi* nc = nccl_api(); if (nc->ok) nc->CommDestroy(h->comm); h->comm = nullptr; }
  release_mailboxes(h);
  for (auto& g : h->ig) drop_graph(g);
  if (h->prof_calls > 1 && std::getenv("LOKIB200_PROFILE"))
    std::fprintf(stderr, "lokib200 engine %p: %lld blocking intervals; per interval: submit %.1f us, stream wait %.1f us, outside the call %.1f us\n", static_cast<void*>(h),
                 static_cast<long long>(h->prof_calls), h->prof_submit / h->prof_calls, h->prof_wait / h->prof_calls, h->prof_outside / h->prof_calls);
  void* ptrs[] = {h->d_type, h->d_angular, h->d_gas_first, h->d_gas_last, h->d_ap0, h->d_ap1, h->d_mass, h->d_redmass, h->d_eloss, h->d_thstd, h->d_wpar,
                  h->d_gas_fraction, h->d_tables, h->d_state, h->d_id, h->lists.birth, h->lists.dead, h->lists.freed, h->lists.claim, h->lists.dead_flag,
                  h->lists.growth_terms, h->lists.counters, h->pend.col, h->d_adv_part, h->d_birth_part, h->d_smp_part, h->d_result, h->d_pc_result, h->d_maxbits, h->d_eeh, h->d_eah,
                  h->d_evh, h->d_eeh_per, h->d_hist_red};
  const bool prof = std::getenv("LOKIB200_PROFILE") != nullptr;
  auto us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = us();
  for (void* p : ptrs) if (p) cudaFree(p);
  const double t1 = us();
  if (h->h_result) cudaFreeHost(h->h_result);
  const double t2 = us();
  for (auto& pr : h->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  const double t3 = us();
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (prof) std::fprintf(stderr, "lokib200 engine %p released: device memory %.0f us, pinned memory %.0f us, %zu events %.0f us, stream %.0f us\n", static_cast<void*>(h), t1 - t0, t2 - t1, 2 * h->ev_pool.size(), t3 - t2, us() - t3);
  delete h;
}

int lokib200_set_stream(lokib200_engine* h, void* cuda_stream) {
  if (!h) return LOKIB200_ERR_INVALID;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  CK(cudaSetDevice(h->cfg.device));
  if (h->stream) CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (cuda_stream) h->stream = static_cast<cudaStream_t>(cuda_stream);
  else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  apply_l2_window(h);
  return 0;
}

int lokib200_set_processes(lokib200_engine* h, const lokib200_process_soa* p) {
  if (!h || !p || p->n_processes <= 0 || p->n_gases <= 0) return fail(h, LOKIB200_ERR_INVALID, "bad process set");
  CK(cudaSetDevice(h->cfg.device));
  const int P = p->n_processes, nG = p->n_gases;
  h->P = P; h->nG = nG;
  auto cpI = [&](std::vector<int>& d, const int32_t* s, int n) { d.assign(s, s + n); };
  auto cpD = [&](std::vector<double>& d, const double* s, int64_t n) { d.assign(s, s + n); };
  cpI(h->type, p->type, P); cpI(h->superel, p->is_superelastic, P); cpI(h->angular, p->angular_model, P);
  cpD(h->ap0, p->angular_p0, P); cpD(h->ap1, p->angular_p1, P); cpD(h->swf, p->superelastic_weight_factor, P);
  cpD(h->emin, p->energy_min, P); cpD(h->emax, p->energy_max, P); cpD(h->reldens, p->rel_density, P); cpD(h->mass, p->target_mass, P);
  cpD(h->redmass, p->reduced_mass, P); cpD(h->eloss, p->energy_loss, P); cpD(h->thstd, p->thermal_std, P); cpD(h->wpar, p->w_parameter, P);
  cpI(h->gas_first, p->gas_first, nG); cpI(h->gas_last, p->gas_last, nG); cpD(h->gas_fraction, p->gas_fraction, nG);
  h->xs_off.assign(p->xs_offset, p->xs_offset + P + 1);
  cpD(h->xs_e, p->xs_energy, h->xs_off[P]); cpD(h->xs_v, p->xs_value, h->xs_off[P]);
  for (int k = 0; k < P; ++k) {
    if (h->type[k] < 0 || h->type[k] > 2) return fail(h, LOKIB200_ERR_INVALID, "process type out of range");
    if (h->superel[k] && (k == 0 || h->superel[k - 1])) return fail(h, LOKIB200_ERR_INVALID, "a superelastic process must follow its inelastic (BMC.C:209-266)");
    if (!h->superel[k] && h->xs_off[k + 1] - h->xs_off[k] < 2) return fail(h, LOKIB200_ERR_INVALID, "a cross section needs at least two points");
  }
  h->has_pc = false;
  for (int k = 0; k < P; ++k) if (h->type[k] != T_CONSERVATIVE) h->has_pc = true;
  int rc;
  if ((rc = upload(h, &h->d_type, h->type)) || (rc = upload(h, &h->d_angular, h->angular)) || (rc = upload(h, &h->d_gas_first, h->gas_first)) ||
      (rc = upload(h, &h->d_gas_last, h->gas_last)) || (rc = upload(h, &h->d_ap0, h->ap0)) || (rc = upload(h, &h->d_ap1, h->ap1)) ||
      (rc = upload(h, &h->d_mass, h->mass)) || (rc = upload(h, &h->d_redmass, h->redmass)) || (rc = upload(h, &h->d_eloss, h->eloss)) ||
      (rc = upload(h, &h->d_thstd, h->thstd)) || (rc = upload(h, &h->d_wpar, h->wpar)) || (rc = upload(h, &h->d_gas_fraction, h->gas_fraction)))
    return rc;

  // per-interval buffers that depend on P
  h->part_len = R_HEADER + 3 * P;
  int per_sm = 2;
  h->adv_blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + ADV_THREADS - 1) / ADV_THREADS, static_cast<int64_t>(h->sm_count) * per_sm));
  h->tile_blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + POOL - 1) / POOL, static_cast<int64_t>(h->sm_count) * 2));
  h->birth_blocks = h->sm_count;
  h->smp_blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + ADV_THREADS - 1) / ADV_THREADS, static_cast<int64_t>(h->sm_count) * 4));
  // kernel choice, measured (tools/form_crossover.py, profiles/r2_form_crossover.txt): the streaming pool wins from ~1e5 electrons on (K1 95 vs 107 us
  // at 1e5, 111 vs 205 us at 2.5e5, 221 vs 700 us at 1e6 on the N2 workload), one electron per thread below (60 vs 93 us at 5e4).
  // LOKIB200_KERNEL=thread|stream overrides
  h->use_tile = h->cfg.n_electrons >= static_cast<int64_t>(96) * POOL;
  if (const char* env = std::getenv("LOKIB200_KERNEL")) { if (!std::strcmp(env, "thread")) h->use_tile = false; else if (!std::strcmp(env, "stream") || !std::strcmp(env, "tile")) h->use_tile = true; }
  if (stream_smem_bytes(P) > STREAM_SMEM_BUDGET) h->use_tile = false;   // more than half an SM's shared memory: fall back to one electron per thread
  for (auto& g : h->ig) drop_graph(g);   // the captured launches hold the buffers released below
  for (double** q : {&h->d_adv_part, &h->d_birth_part, &h->d_smp_part, &h->d_result}) if (*q) { cudaFree(*q); *q = nullptr; }
  if (h->h_result) { cudaFreeHost(h->h_result); h->h_result = nullptr; }
  CK(cudaMalloc(&h->d_adv_part, static_cast<size_t>(std::max(h->adv_blocks, h->tile_blocks)) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_birth_part, static_cast<size_t>(h->birth_blocks) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_smp_part, static_cast<size_t>(h->smp_blocks) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_result, h->part_len * sizeof(double)));
  CK(cudaMallocHost(&h->h_result, (h->part_len + 2) * sizeof(double)));   // pinned: a pageable target would make the copy synchronous

  // birth/death lists (only when a non-conservative channel exists)
  Lists& L = h->lists;
  for (void* q : {static_cast<void*>(L.birth), static_cast<void*>(L.dead), static_cast<void*>(L.freed), static_cast<void*>(L.claim), static_cast<void*>(L.dead_flag),
                  static_cast<void*>(L.growth_terms), static_cast<void*>(L.counters)})
    if (q) cudaFree(q);
  L = Lists{};
  const size_t n = static_cast<size_t>(h->cfg.n_electrons);
  L.birth_cap = h->has_pc ? static_cast<unsigned int>(std::max<size_t>(4096, n / 4)) : 1u;
  L.dead_cap = h->has_pc ? static_cast<unsigned int>(std::max<size_t>(4096, n / 4)) : 1u;
  CK(cudaMalloc(&L.counters, C_COUNT * sizeof(unsigned int)));
  CK(cudaMemset(L.counters, 0, C_COUNT * sizeof(unsigned int)));
  CK(cudaMalloc(&L.birth, 8ull * L.birth_cap * sizeof(double)));
  CK(cudaMalloc(&L.dead, static_cast<size_t>(L.dead_cap) * sizeof(unsigned int)));
  CK(cudaMalloc(&L.freed, static_cast<size_t>(L.birth_cap) * sizeof(unsigned int)));
  CK(cudaMalloc(&L.growth_terms, (static_cast<size_t>(L.birth_cap) + L.dead_cap) * sizeof(double)));
  if (h->pend.col) { cudaFree(h->pend.col); h->pend.col = nullptr; }
  h->pend.cap = L.birth_cap;
  CK(cudaMalloc(&h->pend.col, 9ull * h->pend.cap * sizeof(double)));
  const size_t n_claim = h->has_pc ? n + L.birth_cap : 1, n_flag = h->has_pc ? n : 1;
  CK(cudaMalloc(&L.claim, n_claim * sizeof(unsigned int)));
  CK(cudaMemset(L.claim, 0, n_claim * sizeof(unsigned int)));
  CK(cudaMalloc(&L.dead_flag, n_flag));
  CK(cudaMemset(L.dead_flag, 0, n_flag));
  h->have_processes = true;
  h->have_tables = false;
  h->maxE = LOKIB200_NON_DEF;
  return 0;
}

int lokib200_get_config(const lokib200_engine* h, lokib200_config* cfg) { if (!h || !cfg) return LOKIB200_ERR_INVALID; *cfg = h->cfg; return 0; }
int lokib200_process_count(const lokib200_engine* h) { return (h && h->have_processes) ? h->P : 0; }
int lokib200_get_rel_densities(const lokib200_engine* h, double* rd) {
  if (!h || !rd || !h->have_processes) return LOKIB200_ERR_INVALID;
  std::memcpy(rd, h->reldens.data(), sizeof(double) * h->P);
  return 0;
}

// interpolateCrossSections (BMC.C:561-615), restated on the host; only the cumulative table goes to the device
int lokib200_build_tables(lokib200_engine* h, double max_energy) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  if (!(max_energy > 0)) return fail(h, LOKIB200_ERR_INVALID, "max_energy must be positive");
  CK(cudaSetDevice(h->cfg.device));
  if (h->have_tables && max_energy == h->maxE) return 0;                  // BMC.C:567
  const int nE = h->cfg.n_interp_points, P = h->P;
  const int stride = (P + 15) / 16 * 16;                                  // rows start on 128-byte lines
  h->nE = nE; h->stride = stride; h->maxE = max_energy;
  h->dE = max_energy / static_cast<double>(nE - 1);                       // BMC.C:573
  h->h_cum.assign(static_cast<size_t>(nE) * stride, 0.0);
  h->h_nu_tot.assign(nE, 0.0); h->h_nu_max.assign(nE, 0.0);
  const double Ngas = h->cfg.gas_density;
  double running_max = 0;
  // The rows are visited with rising energy, so the bracket of every process only moves forward: a cursor per process replaces the bisection of
  // lin_interp (the table is rebuilt ~5 times per job while the swarm heats up; 1e4 rows x P processes).  Same bracket - the largest j <= n - 2 with
  // x[j] <= xv - hence the same doubles (checked on all 22 golden process sets, 1.3e7 brackets); a cross section whose energies are not sorted keeps the bisection.
  std::vector<int64_t> cursor(static_cast<size_t>(P), 0);
  std::vector<char> sorted(static_cast<size_t>(P), 1);
  for (int k = 0; k < P; ++k) {
    const int64_t o = h->xs_off[k], n = h->xs_off[k + 1] - o;
    for (int64_t j = 1; j < n; ++j) if (h->xs_e[o + j] < h->xs_e[o + j - 1]) { sorted[k] = 0; break; }
    if (n < 2) sorted[k] = 0;
  }
  auto interpolate = [&](int table, int k, double xv) {   // cross section `table` at xv; the cursor belongs to process k (a superelastic reads its parent's table)
    const int64_t o = h->xs_off[table], n = h->xs_off[table + 1] - o;
    const double* x = h->xs_e.data() + o;
    const double* y = h->xs_v.data() + o;
    if (!sorted[table]) return lin_interp(x, y, n, xv);
    int64_t& c = cursor[k];
    while (c + 1 <= n - 2 && x[c + 1] <= xv) ++c;
    const double dx = x[c + 1] - x[c];
    return y[c] + (xv - x[c]) / dx * (y[c + 1] - y[c]);
  };
  for (int i = 0; i < nE; ++i) {
    const double energy = i * h->dE;
    double acc = 0;
    double* row = h->h_cum.data() + static_cast<size_t>(i) * stride;
    for (int k = 0; k < P; ++k) {
      double value = 0;
      if (h->superel[k]) {                                                // Klein-Rosseland, BMC.C:584-595
        if (energy > h->emin[k] && energy <= h->emax[k])
          value = h->swf[k] * (1.0 + h->emin[k - 1] / energy) * interpolate(k - 1, k, energy + h->emin[k - 1]) * h->reldens[k];
      } else if (energy >= h->emin[k] && energy <= h->emax[k]) {          // BMC.C:597-603
        value = interpolate(k, k, energy) * h->reldens[k];
      }
      acc += value;
      row[k] = acc;
    }
    for (int k = P; k < stride; ++k) row[k] = acc;
    acc *= Ngas * std::sqrt(energy * 2.0 * QE / ME);                      // BMC.C:610
    h->h_nu_tot[i] = acc;
    running_max = std::fmax(acc, running_max);
    h->h_nu_max[i] = running_max;
  }
  return push_tables(h);
}

int lokib200_upload_tables(lokib200_engine* h, const double* cum, const double* nu_tot, const double* nu_max, int32_t nE, double dE) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  if (!cum || !nu_tot || !nu_max || nE < 2 || !(dE > 0)) return fail(h, LOKIB200_ERR_INVALID, "bad tables");
  CK(cudaSetDevice(h->cfg.device));
  const int P = h->P, stride = (P + 15) / 16 * 16;
  h->nE = nE; h->stride = stride; h->dE = dE; h->maxE = dE * (nE - 1);
  h->h_cum.assign(static_cast<size_t>(nE) * stride, 0.0);
  for (int i = 0; i < nE; ++i) {
    double* row = h->h_cum.data() + static_cast<size_t>(i) * stride;
    std::memcpy(row, cum + static_cast<size_t>(i) * P, sizeof(double) * P);
    for (int k = P; k < stride; ++k) row[k] = row[P - 1];
  }
  h->h_nu_tot.assign(nu_tot, nu_tot + nE); h->h_nu_max.assign(nu_max, nu_max + nE);
  return push_tables(h);
}

int lokib200_get_tables(lokib200_engine* h, double* cum, double* nu_tot, double* nu_max) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  CK(cudaSetDevice(h->cfg.device));
  if (cum) {   // read back from the DEVICE copy so that tests see what the kernels see
    std::vector<double> tmp(static_cast<size_t>(h->nE) * h->stride);
    CK(cudaMemcpy(tmp.data(), h->d_cum, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < h->nE; ++i) std::memcpy(cum + static_cast<size_t>(i) * h->P, tmp.data() + static_cast<size_t>(i) * h->stride, sizeof(double) * h->P);
  }
  if (nu_tot) CK(cudaMemcpy(nu_tot, h->d_nu_tot, static_cast<size_t>(h->nE) * sizeof(double), cudaMemcpyDeviceToHost));
  if (nu_max) std::memcpy(nu_max, h->h_nu_max.data(), static_cast<size_t>(h->nE) * sizeof(double));
  return 0;
}

int lokib200_table_info(const lokib200_engine* h, int32_t* nE, double* dE, double* max_energy, double* nu_max_last) {
  if (!h || !h->have_tables) return LOKIB200_ERR_INVALID;
  if (nE) *nE = h->nE;
  if (dE) *dE = h->dE;
  if (max_energy) *max_energy = h->maxE;
  if (nu_max_last) *nu_max_last = h->h_nu_max[h->nE - 1];
  return 0;
}

double lokib200_nu_max_at(const lokib200_engine* h, int32_t index) {
  if (!h || !h->have_tables) return LOKIB200_NON_DEF;
  return h->h_nu_max[std::min(std::max(index, 0), h->nE - 1)];
}

int lokib200_init_ensemble(lokib200_engine* h, double temp_ratio, double* max_energy) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  CK(cudaSetDevice(h->cfg.device));
  const double sd = std::sqrt(KB * temp_ratio * h->cfg.gas_temperature / ME);   // BMC.C:496
  CK(cudaMemsetAsync(h->d_maxbits, 0, sizeof(unsigned long long), h->stream));
  const int blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + 255) / 256, static_cast<int64_t>(h->sm_count) * 8));
  k_init_ensemble<<<blocks, 256, 0, h->stream>>>(h->st, h->cfg.n_electrons, h->cfg.first_electron_id, h->cfg.seed, sd, h->d_maxbits);
  k_identity_ids<<<blocks, 256, 0, h->stream>>>(h->d_id, h->cfg.n_electrons, h->cfg.first_electron_id);
  h->permuted = false;
  h->launches += 2;
  CK(cudaGetLastError());
  unsigned long long bits = 0;
  CK(cudaMemcpyAsync(&bits, h->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (max_energy) std::memcpy(max_energy, &bits, sizeof(double));
  h->time = 0; h->interval = 0;
  return 0;
}

int lokib200_set_ensemble(lokib200_engine* h, const double* soa8, double time) {
  if (!h || !soa8) return LOKIB200_ERR_INVALID;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->d_state, soa8, 8 * static_cast<size_t>(h->cfg.n_electrons) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + 255) / 256, static_cast<int64_t>(h->sm_count) * 8));
  k_identity_ids<<<blocks, 256, 0, h->stream>>>(h->d_id, h->cfg.n_electrons, h->cfg.first_electron_id);
  ++h->launches;
  CK(cudaStreamSynchronize(h->stream));
  h->permuted = false;
  h->time = time;
  return 0;
}

int lokib200_get_ensemble(lokib200_engine* h, double* soa8) {
  if (!h || !soa8) return LOKIB200_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  const size_t n = static_cast<size_t>(h->cfg.n_electrons);
  if (!h->permuted) {
    CK(cudaMemcpyAsync(soa8, h->d_state, 8 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  // the stream kernel permutes electrons inside each CTA's range: return them in electron-id order (index i <-> id first_id + i)
  std::vector<double> tmp(8 * n);
  std::vector<unsigned long long> ids(n);
  CK(cudaMemcpyAsync(tmp.data(), h->d_state, 8 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ids.data(), h->d_id, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t p = 0; p < n; ++p) {
    const size_t i = static_cast<size_t>(ids[p] - h->cfg.first_electron_id);
    if (i >= n) return fail(h, LOKIB200_ERR_INVALID, "corrupt electron id column");
    for (int c = 0; c < 8; ++c) soa8[c * n + i] = tmp[c * n + p];
  }
  return 0;
}

double lokib200_time(const lokib200_engine* h) { return h ? h->time : LOKIB200_NON_DEF; }

// node added to a capturing stream by the launch just made (null outside a capture)
static cudaGraphNode_t last_captured_node(cudaStream_t stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  const cudaGraphNode_t* deps = nullptr;
  size_t nd = 0;
  if (cudaStreamGetCaptureInfo(stream, &st, nullptr, nullptr, &deps, &nd) != cudaSuccess || st != cudaStreamCaptureStatusActive || nd != 1) return nullptr;
  return deps[0];
}

static bool graph_eligible(lokib200_engine* h);

// launch geometry of the histogram pass for the current grid
static void histogram_launch_shape(const lokib200_engine* h, int& blocks, size_t& smem) {
  smem = (static_cast<size_t>(h->hist.nEn) + hist_tile_words(h->hist)) * 4 + 16;
  blocks = static_cast<int>(std::min<int64_t>((h->cfg.n_electrons + HIST_THREADS - 1) / HIST_THREADS, static_cast<int64_t>(h->sm_count) * 2));
}
static HistGrid histogram_args(const lokib200_engine* h, int phase_index) {
  HistGrid g = h->hist;
  g.eeh_phase = (phase_index >= 0) ? h->d_eeh_per + static_cast<size_t>(phase_index) * g.nEn : nullptr;
  return g;
}

// getTimeDependDistributions' counting pass (BMC.C:1551-1571) over the current ensemble
static int enqueue_histograms(lokib200_engine* h, int phase_index, lokib200_engine::IntervalGraph* tap) {
  CK(cudaSetDevice(h->cfg.device));
  const HistGrid g = histogram_args(h, phase_index);
  h->hist_reduced = false;
  int hblocks; size_t hsmem;
  histogram_launch_shape(h, hblocks, hsmem);
  static std::atomic<bool> attr_set[64];
  if (!attr_set[h->cfg.device & 63]) { CK(cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_set[h->cfg.device & 63] = true; }
  if (hsmem > 100 * 1024) return fail(h, LOKIB200_ERR_INVALID, "histogram grids too fine for the shared-memory tiles");
  k_histogram<<<hblocks, HIST_THREADS, hsmem, h->stream>>>(h->st, h->cfg.n_electrons, g);
  if (tap) tap->hist = last_captured_node(h->stream); else ++h->launches;
  CK(cudaGetLastError());
  return 0;
}
static int flush_pending_hist(lokib200_engine* h) {
  if (!h || !h->hist_pending) return 0;
  h->hist_pending = false;
  return enqueue_histograms(h, h->hist_pending_phase, nullptr);
}

// population control of a small ensemble runs as ONE single-CTA kernel: its lists hold at most a few thousand entries and five dependent launches
// cost more than the work (LOKIB200_PC_SMALL=0 keeps the five grids)
static bool pc_in_one_cta(const lokib200_engine* h) {
  static const bool on = [] { const char* e = std::getenv("LOKIB200_PC_SMALL"); return !(e && e[0] == '0'); }();
  return on && h->cfg.n_electrons <= 262144;
}

// spin limit of the exchange kernel, ~60 s of SM clocks: ranks may reach their first exchange seconds apart; a rank that never comes ends as an
// error code in the vector, not as a hung GPU
constexpr long long EXCHANGE_TIMEOUT_CYCLES = 120000000000ll;

// the peer-memory exchange of one engine's result vector (all ranks of the communicator must enqueue theirs)
static int enqueue_exchange(lokib200_engine* e, double* v, lokib200_engine::IntervalGraph* tap) {
  lokib200_engine* h = e;
  CK(cudaSetDevice(e->cfg.device));
  ++e->exchange_epoch;
  k_exchange<<<1, EXCHANGE_THREADS, 0, e->stream>>>(v, e->part_len, e->mail_stride, e->mail, e->comm_rank, e->comm_size, e->exchange_epoch, EXCHANGE_TIMEOUT_CYCLES);
  if (tap) tap->exchange = last_captured_node(e->stream); else ++e->launches;
  CK(cudaGetLastError());
  return 0;
}

// everything one synchronisation interval launches, in stream order; `tap` (capture only) receives the nodes with per-interval arguments
static int enqueue_interval(lokib200_engine* h, const Model& m, const AdvArgs& a, bool sample, double* d_result, lokib200_engine::IntervalGraph* tap) {
  int rc = 0;
  const bool fused = sample && !h->has_pc;
  HistGrid no_hist{};   // histograms are sampled by lokib200_sample_histograms
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = !tap && h->timing && h->ev_used < lokib200_engine::EV_CAP;
  if (timed) {
    if (h->ev_used == h->ev_pool.size()) {
      cudaEvent_t a0, a1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
      h->ev_pool.emplace_back(a0, a1);
    }
    e0 = h->ev_pool[h->ev_used].first; e1 = h->ev_pool[h->ev_used].second; ++h->ev_used;
    CK(cudaEventRecord(e0, h->stream));
  }
  int kernels = 0;
  if (h->use_tile) { if ((rc = launch_stream(h, m, a, no_hist))) return rc; h->last_adv_blocks = h->tile_blocks; h->permuted = true; }
  else { if ((rc = launch_advance(h, fused, m, a, no_hist))) return rc; h->last_adv_blocks = h->adv_blocks; }
  if (tap) tap->k1 = last_captured_node(h->stream);
  if (timed) CK(cudaEventRecord(e1, h->stream));
  ++kernels;
  CK(cudaGetLastError());
  const double* smp = nullptr;
  const double* births = nullptr;
  if (h->has_pc && h->use_tile) {
    launch_births(h, m, a); ++kernels; births = h->d_birth_part; CK(cudaGetLastError());
    if (tap) tap->births = last_captured_node(h->stream);
  }
  if (h->has_pc && pc_in_one_cta(h)) {   // small ensemble: the five phases in one CTA (lk_kernels.cuh k_pc_small)
    k_pc_small<<<1, PC_SMALL_THREADS, 0, h->stream>>>(h->st, h->lists, a.n, a.first_id, a.seed, a.interval, h->d_pc_result);
    if (tap) tap->pc_small = last_captured_node(h->stream);
    ++kernels;
  } else if (h->has_pc) {
    const int pcb = std::max(1, std::min(h->sm_count * 2, static_cast<int>((h->lists.birth_cap + 255) / 256)));
    k_pc_fill<<<pcb, 256, 0, h->stream>>>(h->st, h->lists);
    k_pc_copy<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n, a.first_id, a.seed, a.interval);
    if (tap) tap->pc_copy = last_captured_node(h->stream);
    k_pc_lottery<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n, a.first_id, a.seed, a.interval);
    if (tap) tap->pc_lottery = last_captured_node(h->stream);
    k_pc_place<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n);
    k_pc_reset<<<1, 256, 0, h->stream>>>(h->lists, a.n, h->d_pc_result);
    kernels += 5;
  }
  if (sample && (h->has_pc || h->use_tile)) {   // separate sampling pass (the thread kernel fuses it when nothing can be born or lost)
    k_sample<<<h->smp_blocks, ADV_THREADS, 16, h->stream>>>(h->st, a.n, no_hist, h->P, h->d_smp_part);
    ++kernels;
    smp = h->d_smp_part;
  }
  k_finalize<<<h->part_len, 32, 0, h->stream>>>(h->d_adv_part, h->last_adv_blocks, births, h->birth_blocks, smp, h->smp_blocks,
                                                h->has_pc ? h->d_pc_result : nullptr, h->P, d_result ? d_result : h->d_result);
  ++kernels;
  CK(cudaGetLastError());
  if (tap) tap->kernels = kernels; else h->launches += kernels;
  return 0;
}

static int begin_interval(lokib200_engine* h, double nu_trial, double t_sync, Model& m, AdvArgs& a) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  if (!(nu_trial > 0) || !(t_sync > h->time)) return fail(h, LOKIB200_ERR_INVALID, "need nu_trial > 0 and t_sync > current time");
  CK(cudaSetDevice(h->cfg.device));
  if (h->fast_mode && (h->band_for_nu != nu_trial || h->band_for_version != h->table_version)) build_bands(h, nu_trial);
  m = make_model(h);
  ++h->interval;
  a = AdvArgs{};
  a.n = h->cfg.n_electrons; a.first_id = h->cfg.first_electron_id; a.seed = h->cfg.seed; a.interval = h->interval;
  a.nu_trial = nu_trial; a.t0 = h->time; a.t_sync = t_sync;
  return 0;
}

int lokib200_advance_to_sync_device(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* d_result) {
  Model m; AdvArgs a;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  int rc = begin_interval(h, nu_trial, t_sync, m, a);
  if (rc) return rc;
  if ((rc = enqueue_interval(h, m, a, sample != 0, d_result, nullptr))) return rc;
  h->time = t_sync;
  return 0;
}

static int enqueue_result_copy(lokib200_engine* h) {
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->h_result, h->d_result, h->part_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
static int wait_result(lokib200_engine* h, double* result) {
  CK(cudaStreamSynchronize(h->stream));
  if (result) std::memcpy(result, h->h_result, h->part_len * sizeof(double));
  if (h->h_result[R_OVERFLOW] == 2.0) return fail(h, LOKIB200_ERR_CUDA, "the exchange of the result vectors timed out: a rank of the communicator did not arrive");
  if (h->h_result[R_OVERFLOW] != 0) return fail(h, LOKIB200_ERR_OVERFLOW, "birth/death list overflow inside one synchronisation interval (on this or another rank)");
  return 0;
}

int lokib200_read_result(lokib200_engine* h, double* result) {
  if (!h || !h->d_result) return LOKIB200_ERR_INVALID;
  const int rc = enqueue_result_copy(h);
  return rc ? rc : wait_result(h, result);
}

// The blocking interval (lokib200_advance_to_sync: K1, the births pass of the streaming form, five population-control kernels, k_sample, k_finalize,
// the copy of the result vector) is a chain of ~10 launches.  For a small ensemble they cost more host time than device time as separate API calls
// and serialise the host threads of jobs that run side by side (host/run.cpp); for a large one they delay the start of K1 after the host's turn.
// The chain is captured ONCE per engine (and per sample / deferred-histogram flag) as a CUDA graph; every interval then costs the argument update
// of the nodes that see the interval (K1 and the births pass: model + interval arguments; the two lottery kernels: the interval number), one graph
// launch and one stream wait.  Same kernels, same arguments, same order: results are those of the plain launches, which
// lokib200_advance_to_sync_device keeps using.  LOKIB200_GRAPH=0 selects the plain launches here too, 1 restricts the graph to the thread form.
static bool graph_eligible(lokib200_engine* h) {
  if (h->graph_off || (h->comm && (!h->p2p || h->comm_local_group))) return false;   // (an NCCL all-reduce is not captured; local groups advance asynchronously)
  static const int mode = [] { const char* e = std::getenv("LOKIB200_GRAPH"); return e ? std::atoi(e) : 2; }();   // 0 = never, 1 = one-electron-per-thread form only, 2 = both forms (default)
  return mode >= 2 || (mode == 1 && !h->use_tile);
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int advance_graph(lokib200_engine* h, double nu_trial, double t_sync, bool sample, double* result) {
  const double t_in = now_us();
  Model m; AdvArgs a;
  int rc = begin_interval(h, nu_trial, t_sync, m, a);
  if (rc) return rc;
  const bool with_hist = h->hist_pending;
  const int hist_phase = h->hist_pending_phase;
  h->hist_pending = false;
  lokib200_engine::IntervalGraph& g = h->ig[(sample ? 1 : 0) + (with_hist ? 2 : 0)];
  const unsigned long long epoch0 = h->exchange_epoch;
  if (!g.exec) {
    bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      ok = (!with_hist || enqueue_histograms(h, hist_phase, &g) == 0) && enqueue_interval(h, m, a, sample, nullptr, &g) == 0 &&
           (!h->comm || enqueue_exchange(h, h->d_result, &g) == 0) &&
           cudaMemcpyAsync(h->h_result, h->d_result, h->part_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
      ok = (cudaStreamEndCapture(h->stream, &g.graph) == cudaSuccess) && ok && g.graph && g.k1 && (!h->has_pc || g.pc_small || (g.pc_copy && g.pc_lottery)) && (!with_hist || g.hist) &&
           (!(h->has_pc && h->use_tile) || g.births) && (!h->comm || g.exchange);
    }
    if (ok) ok = cudaGraphInstantiate(&g.exec, g.graph, 0) == cudaSuccess;
    if (ok) ok = cudaGraphKernelNodeGetParams(g.k1, &g.k1_p) == cudaSuccess;
    if (ok && g.pc_small) ok = cudaGraphKernelNodeGetParams(g.pc_small, &g.small_p) == cudaSuccess;
    if (ok && g.pc_copy) ok = cudaGraphKernelNodeGetParams(g.pc_copy, &g.copy_p) == cudaSuccess && cudaGraphKernelNodeGetParams(g.pc_lottery, &g.lot_p) == cudaSuccess;
    if (ok && with_hist) ok = cudaGraphKernelNodeGetParams(g.hist, &g.hist_p) == cudaSuccess;
    if (ok && g.births) ok = cudaGraphKernelNodeGetParams(g.births, &g.births_p) == cudaSuccess;
    if (ok && g.exchange) { ok = cudaGraphKernelNodeGetParams(g.exchange, &g.exchange_p) == cudaSuccess; ++g.kernels; }
    if (ok && with_hist) ++g.kernels;
    if (!ok) {   // no graph on this engine: the plain launches of the same interval
      cudaGetLastError();
      drop_graph(g);
      h->graph_off = true;
      if (with_hist && (rc = enqueue_histograms(h, hist_phase, nullptr))) return rc;
      if ((rc = enqueue_interval(h, m, a, sample, nullptr, nullptr))) return rc;
      h->time = t_sync;
      if (h->comm) { h->exchange_epoch = epoch0; if ((rc = enqueue_exchange(h, h->d_result, nullptr))) return rc; }   // (the failed capture may have counted this exchange)
      if ((rc = enqueue_result_copy(h))) return rc;
      return wait_result(h, result);
    }
  } else {
    HistGrid no_hist{};
    double* part = h->d_adv_part;
    cudaKernelNodeParams p = g.k1_p;
    p.extra = nullptr;
    if (h->use_tile) {   // k_advance_stream(m, {state, ids}, lists, pending, args, grid, partials): launch_stream_t
      StateId sid{h->st, h->d_id};
      AdvArgs as = a;
      as.pad = stream_stages_nu(h->P) ? static_cast<unsigned int>(std::min(h->nE, NU_STAGE_ROWS)) : 0u;
      void* k1_args[7] = {&m, &sid, &h->lists, &h->pend, &as, &no_hist, &part};
      p.kernelParams = k1_args;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.k1, &p));
      if (g.births) {
        double* bpart = h->d_birth_part;
        void* b_args[5] = {&m, &h->lists, &h->pend, &a, &bpart};
        p = g.births_p; p.kernelParams = b_args; p.extra = nullptr;
        CK(cudaGraphExecKernelNodeSetParams(g.exec, g.births, &p));
      }
    } else {             // k_advance(m, state, lists, args, grid, partials): launch_advance_t
      void* k1_args[6] = {&m, &h->st, &h->lists, &a, &no_hist, &part};
      p.kernelParams = k1_args;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.k1, &p));
    }
    if (g.pc_small) {
      double* pcr = h->d_pc_result;
      void* pc_args[7] = {&h->st, &h->lists, &a.n, &a.first_id, &a.seed, &a.interval, &pcr};
      p = g.small_p; p.kernelParams = pc_args; p.extra = nullptr;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.pc_small, &p));
    } else if (h->has_pc) {
      void* pc_args[6] = {&h->st, &h->lists, &a.n, &a.first_id, &a.seed, &a.interval};
      p = g.copy_p; p.kernelParams = pc_args; p.extra = nullptr;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.pc_copy, &p));
      p = g.lot_p; p.kernelParams = pc_args; p.extra = nullptr;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.pc_lottery, &p));
    }
    if (g.exchange) {   // the epoch every rank stamps on its flags
      ++h->exchange_epoch;
      double* v = h->d_result;
      int len = h->part_len, stride = h->mail_stride, me = h->comm_rank, nr = h->comm_size;
      long long timeout = EXCHANGE_TIMEOUT_CYCLES;
      void* x_args[8] = {&v, &len, &stride, &h->mail, &me, &nr, &h->exchange_epoch, &timeout};
      p = g.exchange_p; p.kernelParams = x_args; p.extra = nullptr;
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.exchange, &p));
    }
    if (with_hist) {   // phase row (AC fields) and, after a regrid, the grid itself
      HistGrid hg = histogram_args(h, hist_phase);
      int hblocks; size_t hsmem;
      histogram_launch_shape(h, hblocks, hsmem);
      if (hsmem > 100 * 1024) return fail(h, LOKIB200_ERR_INVALID, "histogram grids too fine for the shared-memory tiles");
      long long n_el = h->cfg.n_electrons;
      void* h_args[3] = {&h->st, &n_el, &hg};
      p = g.hist_p; p.kernelParams = h_args; p.extra = nullptr;
      p.gridDim = dim3(static_cast<unsigned>(hblocks)); p.sharedMemBytes = static_cast<unsigned>(hsmem);
      CK(cudaGraphExecKernelNodeSetParams(g.exec, g.hist, &p));
      h->hist_reduced = false;
    }
  }
  CK(cudaGraphLaunch(g.exec, h->stream));
  h->launches += g.kernels;
  h->last_adv_blocks = h->use_tile ? h->tile_blocks : h->adv_blocks;
  if (h->use_tile) h->permuted = true;
  h->time = t_sync;
  const double t_submitted = now_us();
  rc = wait_result(h, result);
  const double t_out = now_us();
  if (h->prof_calls++ > 0) h->prof_outside += t_in - h->prof_last_return;
  h->prof_submit += t_submitted - t_in; h->prof_wait += t_out - t_submitted; h->prof_last_return = t_out;
  return rc;
}

int lokib200_advance_to_sync(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* result) {
  if (h && graph_eligible(h)) return advance_graph(h, nu_trial, t_sync, sample != 0, result);
  const double t_in = now_us();
  int rc = lokib200_advance_to_sync_device(h, nu_trial, t_sync, sample, nullptr);
  if (rc) return rc;
  if (h->comm) {   // shards of one job: every rank returns the combined vector
    if (h->comm_local_group) return fail(h, LOKIB200_ERR_INVALID, "engines of one process share a communicator: advance them with lokib200_advance_to_sync_device and combine with lokib200_comm_allreduce_results");
    lokib200_engine* one[1] = {h};
    if ((rc = lokib200_comm_allreduce_results(one, 1, nullptr))) return rc;
  }
  if ((rc = enqueue_result_copy(h))) return rc;
  const double t_submitted = now_us();
  rc = wait_result(h, result);
  const double t_out = now_us();
  if (h->prof_calls++ > 0) h->prof_outside += t_in - h->prof_last_return;
  h->prof_submit += t_submitted - t_in; h->prof_wait += t_out - t_submitted; h->prof_last_return = t_out;
  return rc;
}

// ---------------------------------------------------------------- multi-GPU exchange ----------------------------------------------------------------
int lokib200_comm_unique_id(void* id128) {
  NcclApi* nc = nccl_api();
  if (!id128 || !nc->ok) { g_create_error = nc->ok ? "null argument" : nc->why; return nc->ok ? LOKIB200_ERR_INVALID : LOKIB200_ERR_CUDA; }
  static_assert(sizeof(ncclUniqueId) == LOKIB200_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = nc->GetUniqueId(&id);
  if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + nc->GetErrorString(r); return LOKIB200_ERR_CUDA; }
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

// ---- mailboxes of the peer-memory exchange ----
static size_t mailbox_bytes(int n_ranks, int stride) { return (2ull * n_ranks * stride) * sizeof(double) + 2ull * n_ranks * sizeof(unsigned long long); }
static bool p2p_wanted() { const char* e = std::getenv("LOKIB200_P2P"); return !(e && e[0] == '0'); }   // (read when a communicator is created)

static void release_mailboxes(lokib200_engine* h) {
  for (int r = 0; r < EXCHANGE_MAX_RANKS; ++r) { if (h->mail_ipc[r] && h->mail.box[r]) cudaIpcCloseMemHandle(h->mail.box[r]); h->mail_ipc[r] = false; h->mail.box[r] = nullptr; }
  if (h->d_mail) { cudaFree(h->d_mail); h->d_mail = nullptr; }
  h->p2p = false; h->exchange_epoch = 0;
}
static bool alloc_mailbox(lokib200_engine* h) {
  h->mail_stride = (h->part_len + 15) / 16 * 16;
  const size_t bytes = mailbox_bytes(h->comm_size, h->mail_stride);
  if (cudaSetDevice(h->cfg.device) != cudaSuccess || cudaMalloc(&h->d_mail, bytes) != cudaSuccess) { cudaGetLastError(); h->d_mail = nullptr; return false; }
  return cudaMemset(h->d_mail, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
}
// engines of ONE process: peer access between their devices, mailboxes addressed directly
static void setup_p2p_local(lokib200_engine* const* engines, int n) {
  if (!p2p_wanted() || n < 2 || n > EXCHANGE_MAX_RANKS) return;
  bool ok = true;
  for (int i = 0; i < n && ok; ++i) {
    for (int j = 0; j < n && ok; ++j) {
      if (i == j) continue;
      int can = 0;
      ok = cudaDeviceCanAccessPeer(&can, engines[i]->cfg.device, engines[j]->cfg.device) == cudaSuccess && can;
      if (ok) {
        cudaSetDevice(engines[i]->cfg.device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(engines[j]->cfg.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else ok = e == cudaSuccess;
      }
    }
  }
  for (int i = 0; i < n && ok; ++i) ok = alloc_mailbox(engines[i]);
  if (!ok) { cudaGetLastError(); for (int i = 0; i < n; ++i) release_mailboxes(engines[i]); return; }
  for (int i = 0; i < n; ++i) { for (int r = 0; r < n; ++r) engines[i]->mail.box[r] = engines[r]->d_mail; engines[i]->p2p = true; }
}
// one engine per process: the 64-byte IPC handles of the mailboxes travel through the communicator once (ncclAllGather)
static void setup_p2p_ipc(lokib200_engine* h, NcclApi* nc) {
  if (!p2p_wanted() || h->comm_size < 2 || h->comm_size > EXCHANGE_MAX_RANKS || !nc->AllGather) return;
  const int n = h->comm_size;
  // every rank must take the same decision, so failures are exchanged too: byte 64 of a rank's record says "my mailbox exists"
  constexpr int REC = 128;
  unsigned char mine[REC] = {};
  if (alloc_mailbox(h)) {
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, h->d_mail) == cudaSuccess) { static_assert(sizeof(hd) == 64, "IPC handle size"); std::memcpy(mine, &hd, 64); mine[64] = 1; }
    else cudaGetLastError();
  }
  unsigned char* d_rec = nullptr;
  std::vector<unsigned char> all(static_cast<size_t>(REC) * n);
  bool ok = cudaMalloc(&d_rec, all.size() + REC) == cudaSuccess;
  if (ok) ok = cudaMemcpy(d_rec + all.size(), mine, REC, cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) ok = nc->AllGather(d_rec + all.size(), d_rec, REC, ncclChar, h->comm, h->stream) == ncclSuccess;
  if (ok) ok = cudaStreamSynchronize(h->stream) == cudaSuccess && cudaMemcpy(all.data(), d_rec, all.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (d_rec) cudaFree(d_rec);
  for (int r = 0; r < n && ok; ++r) ok = all[static_cast<size_t>(r) * REC + 64] == 1;
  int opened = 1;
  for (int r = 0; r < n && ok; ++r) {
    if (r == h->comm_rank) { h->mail.box[r] = h->d_mail; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, all.data() + static_cast<size_t>(r) * REC, 64);
    void* q = nullptr;
    if (cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) { h->mail.box[r] = static_cast<double*>(q); h->mail_ipc[r] = true; ++opened; }
    else { cudaGetLastError(); break; }
  }
  // second round: a rank that could not map a peer must keep everybody on the NCCL path
  unsigned char* d_flag = nullptr;
  std::vector<unsigned char> flags(n);
  const unsigned char my_ok = (ok && opened == n) ? 1 : 0;
  bool ok2 = cudaMalloc(&d_flag, n + 1) == cudaSuccess && cudaMemcpy(d_flag + n, &my_ok, 1, cudaMemcpyHostToDevice) == cudaSuccess &&
             nc->AllGather(d_flag + n, d_flag, 1, ncclChar, h->comm, h->stream) == ncclSuccess && cudaStreamSynchronize(h->stream) == cudaSuccess &&
             cudaMemcpy(flags.data(), d_flag, n, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (d_flag) cudaFree(d_flag);
  for (int r = 0; r < n && ok2; ++r) ok2 = flags[r] == 1;
  if (ok2) h->p2p = true; else { cudaGetLastError(); release_mailboxes(h); }
}

int lokib200_comm_init_rank(lokib200_engine* h, const void* id128, int32_t rank, int32_t n_ranks) {
  if (!h || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, LOKIB200_ERR_INVALID, "bad communicator arguments");
  NcclApi* nc = nccl_api();
  if (!nc->ok) return fail(h, LOKIB200_ERR_CUDA, nc->why);
  if (h->comm) return fail(h, LOKIB200_ERR_INVALID, "the engine already has a communicator");
  CK(cudaSetDevice(h->cfg.device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NK(nc->CommInitRank(&h->comm, n_ranks, id, rank));
  h->comm_size = n_ranks; h->comm_rank = rank; h->comm_local_group = false;
  for (auto& g : h->ig) drop_graph(g);   // a captured interval knows nothing of the exchange
  setup_p2p_ipc(h, nc);   // (collective; without peer access every rank stays on the NCCL all-reduce)
  return 0;
}

int lokib200_comm_init_all(lokib200_engine* const* engines, int32_t n) {
  if (!engines || n < 1) return LOKIB200_ERR_INVALID;
  lokib200_engine* h = engines[0];
  NcclApi* nc = nccl_api();
  if (!nc->ok) return fail(h, LOKIB200_ERR_CUDA, nc->why);
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!engines[i] || engines[i]->comm) return fail(h, LOKIB200_ERR_INVALID, "null engine or engine with a communicator");
    devs[i] = engines[i]->cfg.device;
    for (int j = 0; j < i; ++j) if (devs[j] == devs[i]) return fail(h, LOKIB200_ERR_INVALID, "lokib200_comm_init_all needs one engine per device");
  }
  std::vector<ncclComm_t> comms(n);
  NK(nc->CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) { engines[i]->comm = comms[i]; engines[i]->comm_size = n; engines[i]->comm_rank = i; engines[i]->comm_local_group = n > 1; }
  for (int i = 0; i < n; ++i) for (auto& g : engines[i]->ig) drop_graph(g);
  setup_p2p_local(engines, n);
  return 0;
}

int lokib200_comm_destroy(lokib200_engine* h) {
  if (!h) return LOKIB200_ERR_INVALID;
  if (!h->comm) return 0;
  NcclApi* nc = nccl_api();
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  NK(nc->CommDestroy(h->comm));
  for (auto& g : h->ig) drop_graph(g);   // (their exchange node addresses the mailboxes released below)
  release_mailboxes(h);
  h->comm = nullptr; h->comm_size = 1; h->comm_rank = 0; h->comm_local_group = false;
  return 0;
}

int32_t lokib200_comm_size(const lokib200_engine* h) { return h ? h->comm_size : 0; }
const char* lokib200_comm_transport(const lokib200_engine* h) { return (!h || !h->comm) ? "none" : (h->p2p ? "peer-memory" : "nccl"); }

int lokib200_comm_allreduce_results(lokib200_engine* const* engines, int32_t n, double* const* d_results) {
  if (!engines || n < 1 || !engines[0]) return LOKIB200_ERR_INVALID;
  lokib200_engine* h = engines[0];
  NcclApi* nc = nccl_api();
  if (!nc->ok) return fail(h, LOKIB200_ERR_CUDA, nc->why);
  for (int i = 0; i < n; ++i) if (!engines[i] || !engines[i]->comm || engines[i]->part_len != h->part_len) return fail(h, LOKIB200_ERR_INVALID, "engine without a communicator / different process sets");
  bool p2p = true;
  for (int i = 0; i < n; ++i) p2p = p2p && engines[i]->p2p;
  if (p2p) {   // one k_exchange launch per engine: vectors pushed into the peers' mailboxes over NVLink, summed in rank order on every rank
    for (int i = 0; i < n; ++i) {
      const int rc = enqueue_exchange(engines[i], (d_results && d_results[i]) ? d_results[i] : engines[i]->d_result, nullptr);
      if (rc) { if (engines[i] != h) h->err = engines[i]->err; return rc; }
    }
    return 0;
  }
  // one NCCL group: [0, SUM_COUNT) by sum, [SUM_COUNT, HEADER) by max, [HEADER, L) by sum -- in place, on each engine's stream
  NK(nc->GroupStart());
  for (int i = 0; i < n; ++i) {
    lokib200_engine* e = engines[i];
    double* v = (d_results && d_results[i]) ? d_results[i] : e->d_result;
    ncclResult_t r = nc->AllReduce(v, v, R_SUM_COUNT, ncclDouble, ncclSum, e->comm, e->stream);
    if (r == ncclSuccess) r = nc->AllReduce(v + R_SUM_COUNT, v + R_SUM_COUNT, R_HEADER - R_SUM_COUNT, ncclDouble, ncclMax, e->comm, e->stream);
    if (r == ncclSuccess) r = nc->AllReduce(v + R_HEADER, v + R_HEADER, e->part_len - R_HEADER, ncclDouble, ncclSum, e->comm, e->stream);
    if (r != ncclSuccess) { nc->GroupEnd(); return fail(h, LOKIB200_ERR_CUDA, std::string("ncclAllReduce: ") + nc->GetErrorString(r)); }
  }
  NK(nc->GroupEnd());
  return 0;
}

int lokib200_comm_allreduce_histograms(lokib200_engine* const* engines, int32_t n) {
  if (!engines || n < 1 || !engines[0]) return LOKIB200_ERR_INVALID;
  lokib200_engine* h = engines[0];
  NcclApi* nc = nccl_api();
  if (!nc->ok) return fail(h, LOKIB200_ERR_CUDA, nc->why);
  for (int i = 0; i < n; ++i) if (!engines[i] || !engines[i]->comm || !engines[i]->hist.enabled) return fail(h, LOKIB200_ERR_INVALID, "engine without a communicator / histogram grid");
  for (int i = 0; i < n; ++i) { const int frc = flush_pending_hist(engines[i]); if (frc) return frc; }
  for (int i = 0; i < n; ++i) {
    lokib200_engine* e = engines[i];
    if (!e->d_hist_red) {
      const HistGrid& g = e->hist;
      const size_t tot = static_cast<size_t>(g.nEn) * (1 + g.nC + e->cfg.n_phases) + static_cast<size_t>(g.nR) * g.nA;
      if (cudaSetDevice(e->cfg.device) != cudaSuccess || cudaMalloc(&e->d_hist_red, tot * 8) != cudaSuccess) return fail(h, LOKIB200_ERR_CUDA, "cudaMalloc(histogram reduction buffer)");
    }
  }
  NK(nc->GroupStart());
  for (int i = 0; i < n; ++i) {
    lokib200_engine* e = engines[i];
    const HistGrid& g = e->hist;
    const size_t ne = g.nEn, nea = ne * g.nC, nev = static_cast<size_t>(g.nR) * g.nA, nep = ne * e->cfg.n_phases;
    unsigned long long* red = e->d_hist_red;   // the accumulators keep the rank's own counts: reducing twice must not double them
    ncclResult_t r = nc->AllReduce(e->d_eeh, red, ne, ncclUint64, ncclSum, e->comm, e->stream);
    if (r == ncclSuccess) r = nc->AllReduce(e->d_eah, red + ne, nea, ncclUint64, ncclSum, e->comm, e->stream);
    if (r == ncclSuccess) r = nc->AllReduce(e->d_evh, red + ne + nea, nev, ncclUint64, ncclSum, e->comm, e->stream);
    if (r == ncclSuccess) r = nc->AllReduce(e->d_eeh_per, red + ne + nea + nev, nep, ncclUint64, ncclSum, e->comm, e->stream);
    if (r != ncclSuccess) { nc->GroupEnd(); return fail(h, LOKIB200_ERR_CUDA, std::string("ncclAllReduce: ") + nc->GetErrorString(r)); }
    e->hist_reduced = true;
  }
  NK(nc->GroupEnd());
  return 0;
}

int lokib200_sample_moments_device(lokib200_engine* h) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  CK(cudaSetDevice(h->cfg.device));
  HistGrid no_hist{};
  CK(cudaMemsetAsync(h->d_adv_part, 0, static_cast<size_t>(h->adv_blocks) * h->part_len * sizeof(double), h->stream));
  k_sample<<<h->smp_blocks, ADV_THREADS, 16, h->stream>>>(h->st, h->cfg.n_electrons, no_hist, h->P, h->d_smp_part);
  k_finalize<<<h->part_len, 32, 0, h->stream>>>(h->d_adv_part, h->adv_blocks, nullptr, 0, h->d_smp_part, h->smp_blocks, nullptr, h->P, h->d_result);
  h->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int lokib200_sample_moments(lokib200_engine* h, double* result) {
  int rc = lokib200_sample_moments_device(h);
  if (rc) return rc;
  if (h->comm && !h->comm_local_group) {
    lokib200_engine* one[1] = {h};
    if ((rc = lokib200_comm_allreduce_results(one, 1, nullptr))) return rc;
  }
  return lokib200_read_result(h, result);
}

int lokib200_regrid_energy_histograms(lokib200_engine* h, double new_max) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  if (!h->hist.enabled || !(new_max > 0)) return fail(h, LOKIB200_ERR_INVALID, "no histogram grid / bad energy");
  CK(cudaSetDevice(h->cfg.device));
  HistGrid& g = h->hist;
  h->hist_reduced = false;
  g.e_step = (0.0 + 1 * (new_max - 0.0) / static_cast<double>(g.nEn)) - 0.0;   // BMC.C:1507-1508
  g.inv_e = 1.0 / g.e_step;
  h->max_eedf_energy = new_max;
  const size_t ne = g.nEn;
  CK(cudaMemsetAsync(h->d_eeh, 0, ne * 8, h->stream)); CK(cudaMemsetAsync(h->d_eah, 0, ne * g.nC * 8, h->stream));
  CK(cudaMemsetAsync(h->d_eeh_per, 0, ne * h->cfg.n_phases * 8, h->stream));
  return 0;
}

int lokib200_set_histogram_grid(lokib200_engine* h, double max_eedf_energy) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  if (!(max_eedf_energy > 0)) return fail(h, LOKIB200_ERR_INVALID, "max_eedf_energy must be positive");
  CK(cudaSetDevice(h->cfg.device));
  const lokib200_config& c = h->cfg;
  HistGrid& g = h->hist;
  h->hist_reduced = false;
  if (h->d_hist_red) { cudaFree(h->d_hist_red); h->d_hist_red = nullptr; }
  g.enabled = 1; g.cylindrical = c.is_cylindrically_symmetric; g.nEn = c.n_energy_cells; g.nC = c.n_cos_cells; g.nR = c.n_radial_cells; g.nA = c.n_axial_cells;
  // Eigen::LinSpaced(size, low, high)[1] - [0] with size = cells + 1 (BMC.C:1863-1882)
  g.e_step = (0.0 + 1 * (max_eedf_energy - 0.0) / static_cast<double>(g.nEn)) - 0.0;
  g.c_first = -1.0; g.c_step = (-1.0 + 1 * (1.0 - (-1.0)) / static_cast<double>(g.nC)) - g.c_first;
  const double max_speed = std::sqrt(2.0 * max_eedf_energy * QE / ME);
  g.r_step = (0.0 + 1 * (max_speed - 0.0) / static_cast<double>(g.nR)) - 0.0;
  g.a_first = -max_speed; g.a_step = (-max_speed + 1 * (max_speed - (-max_speed)) / static_cast<double>(g.nA)) - g.a_first;
  h->max_eedf_energy = max_eedf_energy;
  // shared-memory tiles of the 2-D grids: the lowest 256 energy rows, and the 96 innermost radial rows x the 128 central axial cells (75 KB)
  g.ea_rows = c.is_cylindrically_symmetric ? std::min(g.nEn, 256) : 0;
  g.ev_rows = c.is_cylindrically_symmetric ? std::min(g.nR, 96) : 0;
  g.ev_aw = c.is_cylindrically_symmetric ? std::min(g.nA, 128) : 0;
  g.ev_a0 = (g.nA - g.ev_aw) / 2;
  g.inv_e = 1.0 / g.e_step; g.inv_c = 1.0 / g.c_step; g.inv_r = 1.0 / g.r_step; g.inv_a = 1.0 / g.a_step;
  const size_t ne = g.nEn, nea = ne * g.nC, nev = static_cast<size_t>(g.nR) * g.nA, nep = ne * c.n_phases;
  if (!h->d_eeh) {
    CK(cudaMalloc(&h->d_eeh, ne * 8)); CK(cudaMalloc(&h->d_eah, nea * 8)); CK(cudaMalloc(&h->d_evh, nev * 8)); CK(cudaMalloc(&h->d_eeh_per, nep * 8));
  }
  CK(cudaMemsetAsync(h->d_eeh, 0, ne * 8, h->stream)); CK(cudaMemsetAsync(h->d_eah, 0, nea * 8, h->stream));
  CK(cudaMemsetAsync(h->d_evh, 0, nev * 8, h->stream)); CK(cudaMemsetAsync(h->d_eeh_per, 0, nep * 8, h->stream));
  g.eeh = h->d_eeh; g.eah = h->d_eah; g.evh = h->d_evh; g.eeh_phase = nullptr;
  return 0;
}

int lokib200_sample_histograms(lokib200_engine* h, int32_t phase_index) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  if (!h->hist.enabled) return fail(h, LOKIB200_ERR_INVALID, "call lokib200_set_histogram_grid first");
  if (phase_index >= h->cfg.n_phases) return fail(h, LOKIB200_ERR_INVALID, "phase_index out of range");
  if ((rc = flush_pending_hist(h))) return rc;      // (two samples without an interval in between)
  static const bool defer = [] { const char* e = std::getenv("LOKIB200_DEFER_HIST"); return !(e && e[0] == '0'); }();
  if (defer && graph_eligible(h) && h->ig[1].exec) {   // this engine advances by graph launches: the pass rides in front of the next interval
    h->hist_pending = true; h->hist_pending_phase = phase_index;
    return 0;
  }
  return enqueue_histograms(h, phase_index, nullptr);
}

int lokib200_fetch_histograms(lokib200_engine* h, double* eeh, double* eah, double* evh, double* eeh_periodic) {
  if (!h || !h->hist.enabled) return fail(h, LOKIB200_ERR_INVALID, "no histogram grid");
  { const int frc = flush_pending_hist(h); if (frc) return frc; }
  CK(cudaSetDevice(h->cfg.device));
  const HistGrid& g = h->hist;
  const size_t ne = g.nEn, nea = ne * g.nC, nev = static_cast<size_t>(g.nR) * g.nA, nep = ne * h->cfg.n_phases;
  const bool red = h->hist_reduced && h->d_hist_red;   // after lokib200_comm_allreduce_histograms: the counts of all shards
  const unsigned long long* src[4] = {red ? h->d_hist_red : h->d_eeh, red ? h->d_hist_red + ne : h->d_eah, red ? h->d_hist_red + ne + nea : h->d_evh,
                                      red ? h->d_hist_red + ne + nea + nev : h->d_eeh_per};
  double* dst[4] = {eeh, eah, evh, eeh_periodic};
  const size_t len[4] = {ne, nea, nev, nep};
  std::vector<unsigned long long> tmp[4];
  for (int a = 0; a < 4; ++a) {
    if (!dst[a]) continue;
    tmp[a].resize(len[a]);
    CK(cudaMemcpyAsync(tmp[a].data(), src[a], len[a] * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));   // one synchronisation for all four arrays
  for (int a = 0; a < 4; ++a)
    if (dst[a]) for (size_t i = 0; i < len[a]; ++i) dst[a][i] = static_cast<double>(tmp[a][i]);
  return 0;
}

int lokib200_step_injected(lokib200_engine* h, int32_t n, const lokib200_electron* in, double nu_trial, const double* t_sync, const double* draws,
                           int32_t n_draws, lokib200_electron* out, lokib200_event_out* ev) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  if (n <= 0 || !in || !t_sync || !draws || n_draws <= 0 || !out || !ev) return fail(h, LOKIB200_ERR_INVALID, "bad arguments");
  CK(cudaSetDevice(h->cfg.device));
  struct Scratch {   // released on every return path
    ElectronIO *in = nullptr, *out = nullptr; EventIO* ev = nullptr; double *ts = nullptr, *dr = nullptr;
    ~Scratch() { for (void* q : {static_cast<void*>(in), static_cast<void*>(out), static_cast<void*>(ev), static_cast<void*>(ts), static_cast<void*>(dr)}) if (q) cudaFree(q); }
  } sc;
  ElectronIO *&d_in = sc.in, *&d_out = sc.out; EventIO*& d_ev = sc.ev; double *&d_ts = sc.ts, *&d_dr = sc.dr;
  CK(cudaMalloc(&d_in, sizeof(ElectronIO) * n)); CK(cudaMalloc(&d_out, sizeof(ElectronIO) * n)); CK(cudaMalloc(&d_ev, sizeof(EventIO) * n));
  CK(cudaMalloc(&d_ts, sizeof(double) * n)); CK(cudaMalloc(&d_dr, sizeof(double) * static_cast<size_t>(n) * n_draws));
  CK(cudaMemcpyAsync(d_in, in, sizeof(ElectronIO) * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_ts, t_sync, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_dr, draws, sizeof(double) * static_cast<size_t>(n) * n_draws, cudaMemcpyHostToDevice, h->stream));
  const Model m = make_model(h);
  const int gt = h->cfg.gas_temperature_effect;
  switch (field_case(h->cfg)) {
    case F_DC: launch_injected_g<F_DC>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
#ifndef LK_BENCH_ONLY
    case F_AC: launch_injected_g<F_AC>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
    case F_DCB: launch_injected_g<F_DCB>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
    case F_ECR: launch_injected_g<F_ECR>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
    case F_ACB: launch_injected_g<F_ACB>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
#endif
    default: launch_injected_g<F_DC>(h, gt, m, n, d_in, nu_trial, d_ts, d_dr, n_draws, d_out, d_ev); break;
  }
  ++h->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_out, sizeof(ElectronIO) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ev, d_ev, sizeof(EventIO) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// maximizationAccelerationEnergy (BMC.C:765-802)
double lokib200_max_accel_energy(const lokib200_engine* h, double e0, double dt) {
  if (!h) return LOKIB200_NON_DEF;
  const double e_me = QE / ME;
  const double Ex0 = std::fabs(h->cfg.electric_field[0]), Ez0 = std::fabs(h->cfg.electric_field[2]);
  const double Ex02 = Ex0 * Ex0, Ez02 = Ez0 * Ez0, E02 = Ex02 + Ez02, E0 = std::sqrt(E02);
  const double v0 = std::sqrt(e0 * QE * 2.0 / ME);
  const double w = h->cfg.excitation_omega, W = h->cfg.cyclotron_omega;
  double gain = (E0 * v0 + 0.5 * e_me * E02 * dt) * dt;
  if (W == 0) {
    if (w != 0) gain = std::fmin(gain, 2.0 / w * (e_me * E02 / w + v0 * (Ex0 + Ez0)));
  } else if (w == 0) {
    gain = std::fmin(gain, 0.5 * e_me * Ez02 * dt * dt + (2.0 * e_me * Ex02 / W + 3.0 * v0 * Ex0) / W + v0 * dt * Ez0);
  } else if (std::fabs(w - W) / w < 1E-6) {
    const double W2 = W * W;
    gain = std::fmin(gain, 2.0 * e_me * Ez02 / W2 + e_me * Ex02 / (8.0 * W2) * (4.0 + W * dt * (2.0 + W * dt)) + (v0 * dt + v0 / W) * Ex0 + 2.0 * v0 * Ez0 / W);
  } else {
    const double w2 = w * w, W2 = W * W, d = w2 - W2;
    gain = std::fmin(gain, 2.0 * e_me * Ez02 / w2 + 0.5 * e_me * Ex02 / (d * d) * (5.0 * w2 + 8.0 * w * W + 5.0 * W * W) + 3.0 * v0 * Ex0 / std::fabs(w - W) + 2.0 * v0 * Ez0 / w);
  }
  return e0 + gain;
}

// checkMaxCollisionFrequency (BMC.C:716-763) for a caller-chosen look-ahead `horizon` (the reference uses 10/nu_trial per micro-pass;
// a kernel that runs a whole interval per launch passes interval + 10/nu_trial)
int lokib200_check_nu_trial(lokib200_engine* h, double max_energy_now, double horizon_events, double energy_max_elastic, double* nu_trial) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  if (!nu_trial || !(*nu_trial > 0)) return fail(h, LOKIB200_ERR_INVALID, "nu_trial must be positive");
  const int gt = h->cfg.gas_temperature_effect;
  const double gas_energy = 1.5 * KB * h->cfg.gas_temperature / QE;
  const double thermal = (gt == GT_TRUE || gt == GT_SMART) ? 10.0 * gas_energy : 0.0;   // BMC.C:724-728
  bool updated = true;
  while (updated) {
    updated = false;
    const double maxE = lokib200_max_accel_energy(h, max_energy_now, horizon_events / *nu_trial) + thermal;   // BMC.C:737
    if (maxE > h->maxE || 2.5 * maxE < h->maxE) {                          // BMC.C:741-750
      if ((rc = lokib200_build_tables(h, (2.0 * maxE < energy_max_elastic) ? 2.0 * maxE : energy_max_elastic))) return rc;
    }
    const int idx = static_cast<int>(std::fmin(std::ceil(maxE / h->dE), static_cast<double>(h->nE - 1)));   // BMC.C:754
    if (*nu_trial < h->h_nu_max[idx]) { updated = true; *nu_trial *= 1.1; }   // BMC.C:758-761
  }
  return 0;
}

// Fast mode: per half-octave band [.., Eu) the smallest trial frequency that bounds nu_tot over every energy an electron starting below Eu
// can reach within the look-ahead time tau = 3 / nu (fixed point of nu = max nu_tot up to maximizationAccelerationEnergy(Eu, 3 / nu) + the
// thermal margin of checkMaxCollisionFrequency, BMC.C:724-737), never above the global trial frequency, which needs no look-ahead limit.
static void build_bands(lokib200_engine* h, double nu_trial) {
  const int gt = h->cfg.gas_temperature_effect;
  const double thermal = (gt == GT_TRUE || gt == GT_SMART) ? 10.0 * (1.5 * KB * h->cfg.gas_temperature / QE) : 0.0;
  auto maxnu = [&](double E) { return h->h_nu_max[static_cast<int>(std::fmin(std::ceil(E / h->dE), static_cast<double>(h->nE - 1)))]; };
  for (int b = 0; b < N_BANDS; ++b) {
    const double Eu = std::ldexp(1.0, -12) * std::pow(2.0, 0.5 * (b + 1));
    double nu = nu_trial, tau = 1e300;
    if (Eu + thermal < h->maxE) {
      double cand = maxnu(Eu + thermal);
      for (int it = 0; it < 64 && cand < nu_trial; ++it) {
        const double reach = lokib200_max_accel_energy(h, Eu, 3.0 / cand) + thermal;
        if (reach >= h->maxE) { cand = nu_trial; break; }
        const double next = maxnu(reach);
        if (next <= cand) break;
        cand = next;
      }
      if (cand < nu_trial) { nu = cand; tau = 3.0 / cand; }
    }
    h->band_nu[b] = nu; h->band_tau[b] = tau;
  }
  h->band_for_nu = nu_trial; h->band_for_version = h->table_version;
}

int lokib200_set_fast_mode(lokib200_engine* h, int32_t on) {
  if (!h) return LOKIB200_ERR_INVALID;
  h->fast_mode = on != 0;
  h->band_for_version = ~0ull;
  return 0;
}

double lokib200_device_hbm_gbs(const lokib200_engine* h) {
  if (!h) return 0.0;
  int clock_khz = 0, bus_bits = 0;
  if (cudaDeviceGetAttribute(&clock_khz, cudaDevAttrMemoryClockRate, h->cfg.device) != cudaSuccess || cudaDeviceGetAttribute(&bus_bits, cudaDevAttrGlobalMemoryBusWidth, h->cfg.device) != cudaSuccess) return 0.0;
  return 2.0 * static_cast<double>(clock_khz) * 1e3 * (bus_bits / 8.0) / 1e9;   // double data rate
}
int32_t lokib200_kernel_form(const lokib200_engine* h) { return (h && h->have_processes && h->use_tile) ? 1 : 0; }
int64_t lokib200_launch_count(const lokib200_engine* h) { return h ? h->launches : 0; }

int lokib200_kernel_time_ms(lokib200_engine* h, double* advance_ms, int64_t* launches) {
  if (!h) return LOKIB200_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  double total = 0;
  for (size_t i = 0; i < h->ev_used; ++i) { float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev_pool[i].first, h->ev_pool[i].second)); total += ms; }
  if (advance_ms) *advance_ms = h->ev_used ? total / static_cast<double>(h->ev_used) : 0.0;
  if (launches) *launches = static_cast<int64_t>(h->ev_used);
  h->ev_used = 0;
  return 0;
}

int lokib200_measure_fp64_peak(lokib200_engine* h, double* tflops) {
  if (!h || !tflops) return LOKIB200_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  const int blocks = h->sm_count * 8, threads = 256, iters = 4096;
  struct Scratch {   // released on every return path
    double* d_out = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scratch() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (d_out) cudaFree(d_out); }
  } sc;
  CK(cudaMalloc(&sc.d_out, sizeof(double) * blocks));
  CK(cudaEventCreate(&sc.e0)); CK(cudaEventCreate(&sc.e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(sc.e0, h->stream));
    k_dfma_peak<<<blocks, threads, 0, h->stream>>>(sc.d_out, iters, 1.0000001);
    CK(cudaEventRecord(sc.e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0; CK(cudaEventElapsedTime(&ms, sc.e0, sc.e1));
    if (rep > 0 && ms < best) best = ms;
  }
  *tflops = 2.0 * 16.0 * static_cast<double>(iters) * threads * blocks / (best * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
