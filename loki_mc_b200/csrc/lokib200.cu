// lokib200.cu -- engine and C ABI (include/lokib200.h) of the B200-native electron Monte Carlo hot path.
//
// Host side of the path: owns device memory, flattens cross sections into the device tables (restating
// BoltzmannMC::interpolateCrossSections, BMC.C:561-615), mirrors the scalar trial-frequency logic (BMC.C:716-802) and
// launches the kernels of lk_kernels.cuh.  There is NO CPU fallback: every entry point fails with LOKIB200_ERR_NO_DEVICE
// when no CUDA device is usable.
#include "../../include/lokib200.h"

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
              LOKIB200_R_N_TABLE_CLAMPED == R_N_TABLE_CLAMPED && LOKIB200_R_N_NU_EXCEEDED == R_N_NU_EXCEEDED, "result layout");
static_assert(sizeof(lokib200_electron) == sizeof(ElectronIO) && sizeof(lokib200_event_out) == sizeof(EventIO), "parity structs");

static std::string g_create_error;

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
  double2 *d_pair = nullptr, *d_coarse = nullptr;   // row-pair form of the cumulative table + its coarse level (see lk_physics.cuh)
  int G = 0, gstride = 0;
  std::vector<double2> h_pair, h_coarse;
  size_t d_cum_cap = 0;

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
  bool use_tile = false;
  int last_adv_blocks = 0;

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
  size_t stream_smem_set = 0;   // dynamic shared memory the streaming kernel's attributes were last set for (0 = never)
};

namespace {

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
  m.pair = h->d_pair; m.coarse = h->d_coarse; m.G = h->G; m.gstride = h->gstride;
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
  return static_cast<size_t>(h->P) * (8 + 8 + 4) + ((sample && h->hist.enabled) ? static_cast<size_t>(h->hist.nEn) * 4 : 0) + 16;
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
  if (h->stream_smem_set != smem) {   // (an engine uses one instantiation on one device)
    CK(cudaFuncSetAttribute(k_advance_stream<F, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CK(cudaFuncSetAttribute(k_advance_stream<F, G, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    h->stream_smem_set = smem;
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

int ensure_ready(lokib200_engine* h, bool need_tables) {
  if (!h) return LOKIB200_ERR_INVALID;
  if (!h->have_processes) return fail(h, LOKIB200_ERR_INVALID, "lokib200_set_processes has not been called");
  if (need_tables && !h->have_tables) return fail(h, LOKIB200_ERR_INVALID, "no tables: call lokib200_build_tables or lokib200_upload_tables first");
  return 0;
}

int push_tables(lokib200_engine* h) {   // h_cum (padded) / h_nu_tot -> device
  const size_t need = static_cast<size_t>(h->nE) * h->stride;
  h->G = (h->P + 7) / 8; h->gstride = (h->G + 3) / 4 * 4;
  const size_t need_c = static_cast<size_t>(h->nE) * h->gstride;
  if (need > h->d_cum_cap) {
    for (void* q : {static_cast<void*>(h->d_cum), static_cast<void*>(h->d_nu_tot), static_cast<void*>(h->d_pair), static_cast<void*>(h->d_coarse)}) if (q) cudaFree(q);
    CK(cudaMalloc(&h->d_cum, need * sizeof(double)));
    CK(cudaMalloc(&h->d_nu_tot, static_cast<size_t>(h->nE) * sizeof(double)));
#if defined(LK_SELECT_2LEVEL) || !defined(LK_SELECT_SPLIT_ROWS)
    CK(cudaMalloc(&h->d_pair, need * sizeof(double2)));
    CK(cudaMalloc(&h->d_coarse, need_c * sizeof(double2)));
#else
    h->d_pair = nullptr; h->d_coarse = nullptr;
#endif
    h->d_cum_cap = need;
  }
#if defined(LK_SELECT_2LEVEL) || !defined(LK_SELECT_SPLIT_ROWS)
  // row-pair + coarse forms, derived from the same doubles (no arithmetic: the kernels see identical table values)
  h->h_pair.resize(need); h->h_coarse.resize(need_c);
  for (int i = 0; i < h->nE; ++i) {
    const double* r1 = h->h_cum.data() + static_cast<size_t>(i) * h->stride;
    const double* r2 = h->h_cum.data() + static_cast<size_t>(std::min(i + 1, h->nE - 1)) * h->stride;
    double2* pr = h->h_pair.data() + static_cast<size_t>(i) * h->stride;
    for (int k = 0; k < h->stride; ++k) pr[k] = make_double2(r1[k], r2[k]);
    double2* cr = h->h_coarse.data() + static_cast<size_t>(i) * h->gstride;
    for (int g = 0; g < h->gstride; ++g) cr[g] = pr[std::min(8 * g + 7, h->stride - 1)];
  }
  CK(cudaMemcpyAsync(h->d_pair, h->h_pair.data(), need * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_coarse, h->h_coarse.data(), need_c * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
#endif
  CK(cudaMemcpyAsync(h->d_cum, h->h_cum.data(), need * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_nu_tot, h->h_nu_tot.data(), static_cast<size_t>(h->nE) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_tables = true;
  return 0;
}

}  // namespace

extern "C" {

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
  void* ptrs[] = {h->d_type, h->d_angular, h->d_gas_first, h->d_gas_last, h->d_ap0, h->d_ap1, h->d_mass, h->d_redmass, h->d_eloss, h->d_thstd, h->d_wpar,
                  h->d_gas_fraction, h->d_cum, h->d_nu_tot, h->d_pair, h->d_coarse, h->d_state, h->d_id, h->lists.birth, h->lists.dead, h->lists.freed, h->lists.claim, h->lists.dead_flag,
                  h->lists.growth_terms, h->lists.counters, h->pend.col, h->d_adv_part, h->d_birth_part, h->d_smp_part, h->d_result, h->d_pc_result, h->d_maxbits, h->d_eeh, h->d_eah,
                  h->d_evh, h->d_eeh_per};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->h_result) cudaFreeHost(h->h_result);
  for (auto& pr : h->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int lokib200_set_stream(lokib200_engine* h, void* cuda_stream) {
  if (!h) return LOKIB200_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  if (h->stream) CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (cuda_stream) h->stream = static_cast<cudaStream_t>(cuda_stream);
  else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
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
  // kernel choice: the tile kernel needs enough tiles to fill the machine; LOKIB200_KERNEL=thread|tile overrides
  // kernel choice: the streaming-pool kernel needs several pools per CTA to amortise its fill/drain; LOKIB200_KERNEL=thread|stream overrides
  h->use_tile = h->cfg.n_electrons >= static_cast<int64_t>(4) * POOL * 2 * h->sm_count;
  if (const char* env = std::getenv("LOKIB200_KERNEL")) { if (!std::strcmp(env, "thread")) h->use_tile = false; else if (!std::strcmp(env, "stream") || !std::strcmp(env, "tile")) h->use_tile = true; }
  if (stream_smem_bytes(P) > STREAM_SMEM_BUDGET) h->use_tile = false;   // more than half an SM's shared memory: fall back to one electron per thread
  for (double** q : {&h->d_adv_part, &h->d_birth_part, &h->d_smp_part, &h->d_result}) if (*q) { cudaFree(*q); *q = nullptr; }
  if (h->h_result) { cudaFreeHost(h->h_result); h->h_result = nullptr; }
  CK(cudaMalloc(&h->d_adv_part, static_cast<size_t>(std::max(h->adv_blocks, h->tile_blocks)) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_birth_part, static_cast<size_t>(h->birth_blocks) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_smp_part, static_cast<size_t>(h->smp_blocks) * h->part_len * sizeof(double)));
  CK(cudaMalloc(&h->d_result, h->part_len * sizeof(double)));
  CK(cudaMallocHost(&h->h_result, (h->part_len + 2) * sizeof(double)));   // + the two population-control words (pinned: a pageable target would make the copy synchronous)

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
  for (int i = 0; i < nE; ++i) {
    const double energy = i * h->dE;
    double acc = 0;
    double* row = h->h_cum.data() + static_cast<size_t>(i) * stride;
    for (int k = 0; k < P; ++k) {
      double value = 0;
      if (h->superel[k]) {                                                // Klein-Rosseland, BMC.C:584-595
        if (energy > h->emin[k] && energy <= h->emax[k]) {
          const int64_t o = h->xs_off[k - 1], n = h->xs_off[k] - o;
          value = h->swf[k] * (1.0 + h->emin[k - 1] / energy) * lin_interp(h->xs_e.data() + o, h->xs_v.data() + o, n, energy + h->emin[k - 1]) * h->reldens[k];
        }
      } else if (energy >= h->emin[k] && energy <= h->emax[k]) {          // BMC.C:597-603
        const int64_t o = h->xs_off[k], n = h->xs_off[k + 1] - o;
        value = lin_interp(h->xs_e.data() + o, h->xs_v.data() + o, n, energy) * h->reldens[k];
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

int lokib200_advance_to_sync_device(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* d_result) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  if (!(nu_trial > 0) || !(t_sync > h->time)) return fail(h, LOKIB200_ERR_INVALID, "need nu_trial > 0 and t_sync > current time");
  CK(cudaSetDevice(h->cfg.device));
  const Model m = make_model(h);
  ++h->interval;
  AdvArgs a{};
  a.n = h->cfg.n_electrons; a.first_id = h->cfg.first_electron_id; a.seed = h->cfg.seed; a.interval = h->interval;
  a.nu_trial = nu_trial; a.t0 = h->time; a.t_sync = t_sync;
  const bool fused = sample && !h->has_pc;
  HistGrid no_hist{};   // histograms are sampled by lokib200_sample_histograms
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = h->timing && h->ev_used < lokib200_engine::EV_CAP;
  if (timed) {
    if (h->ev_used == h->ev_pool.size()) {
      cudaEvent_t a0, a1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
      h->ev_pool.emplace_back(a0, a1);
    }
    e0 = h->ev_pool[h->ev_used].first; e1 = h->ev_pool[h->ev_used].second; ++h->ev_used;
    CK(cudaEventRecord(e0, h->stream));
  }
  if (h->use_tile) { if ((rc = launch_stream(h, m, a, no_hist))) return rc; h->last_adv_blocks = h->tile_blocks; h->permuted = true; }
  else { if ((rc = launch_advance(h, fused, m, a, no_hist))) return rc; h->last_adv_blocks = h->adv_blocks; }
  if (timed) CK(cudaEventRecord(e1, h->stream));
  ++h->launches;
  CK(cudaGetLastError());
  const double* smp = nullptr;
  const double* births = nullptr;
  if (h->has_pc && h->use_tile) { launch_births(h, m, a); ++h->launches; births = h->d_birth_part; CK(cudaGetLastError()); }
  if (h->has_pc) {
    const int pcb = std::max(1, std::min(h->sm_count * 2, static_cast<int>((h->lists.birth_cap + 255) / 256)));
    k_pc_fill<<<pcb, 256, 0, h->stream>>>(h->st, h->lists);
    k_pc_copy<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n, a.first_id, a.seed, a.interval);
    k_pc_lottery<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n, a.first_id, a.seed, a.interval);
    k_pc_place<<<pcb, 256, 0, h->stream>>>(h->st, h->lists, a.n);
    k_pc_reset<<<1, 256, 0, h->stream>>>(h->lists, a.n, h->d_pc_result);
    h->launches += 5;
  }
  if (sample && (h->has_pc || h->use_tile)) {   // separate sampling pass (the thread kernel fuses it when nothing can be born or lost)
    k_sample<<<h->smp_blocks, ADV_THREADS, 16, h->stream>>>(h->st, a.n, no_hist, h->P, h->d_smp_part);
    ++h->launches;
    smp = h->d_smp_part;
  }
  k_finalize<<<h->part_len, 32, 0, h->stream>>>(h->d_adv_part, h->last_adv_blocks, births, h->birth_blocks, smp, h->smp_blocks,
                                                h->has_pc ? h->d_pc_result : nullptr, h->P, d_result ? d_result : h->d_result);
  ++h->launches;
  CK(cudaGetLastError());
  h->time = t_sync;
  return 0;
}

int lokib200_read_result(lokib200_engine* h, double* result) {
  if (!h || !h->d_result) return LOKIB200_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->h_result, h->d_result, h->part_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  double* pc = h->h_result + h->part_len;
  pc[0] = 0; pc[1] = 0;
  if (h->has_pc) CK(cudaMemcpyAsync(pc, h->d_pc_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (result) std::memcpy(result, h->h_result, h->part_len * sizeof(double));
  if (pc[1] != 0) return fail(h, LOKIB200_ERR_OVERFLOW, "birth/death list overflow inside one synchronisation interval");
  return 0;
}

int lokib200_advance_to_sync(lokib200_engine* h, double nu_trial, double t_sync, int32_t sample, double* result) {
  int rc = lokib200_advance_to_sync_device(h, nu_trial, t_sync, sample, nullptr);
  if (rc) return rc;
  return lokib200_read_result(h, result);
}

int lokib200_sample_moments(lokib200_engine* h, double* result) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  CK(cudaSetDevice(h->cfg.device));
  HistGrid no_hist{};
  CK(cudaMemsetAsync(h->d_adv_part, 0, static_cast<size_t>(h->adv_blocks) * h->part_len * sizeof(double), h->stream));
  k_sample<<<h->smp_blocks, ADV_THREADS, 16, h->stream>>>(h->st, h->cfg.n_electrons, no_hist, h->P, h->d_smp_part);
  k_finalize<<<h->part_len, 32, 0, h->stream>>>(h->d_adv_part, h->adv_blocks, nullptr, 0, h->d_smp_part, h->smp_blocks, nullptr, h->P, h->d_result);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_result, h->d_result, h->part_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (result) std::memcpy(result, h->h_result, h->part_len * sizeof(double));
  return 0;
}

int lokib200_regrid_energy_histograms(lokib200_engine* h, double new_max) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  if (!h->hist.enabled || !(new_max > 0)) return fail(h, LOKIB200_ERR_INVALID, "no histogram grid / bad energy");
  CK(cudaSetDevice(h->cfg.device));
  HistGrid& g = h->hist;
  g.e_step = (0.0 + 1 * (new_max - 0.0) / static_cast<double>(g.nEn)) - 0.0;   // BMC.C:1507-1508
  h->max_eedf_energy = new_max;
  const size_t ne = g.nEn;
  CK(cudaMemsetAsync(h->d_eeh, 0, ne * 8, h->stream)); CK(cudaMemsetAsync(h->d_eah, 0, ne * g.nC * 8, h->stream));
  CK(cudaMemsetAsync(h->d_eeh_per, 0, ne * h->cfg.n_phases * 8, h->stream));
  return 0;
}

int lokib200_set_histogram_grid(lokib200_engine* h, double max_eedf_energy) {
  int rc = ensure_ready(h, false);
  if (rc) return rc;
  if (!(max_eedf_energy > 0)) return fail(h, LOKIB200_ERR_INVALID, "max_eedf_energy must be positive");
  CK(cudaSetDevice(h->cfg.device));
  const lokib200_config& c = h->cfg;
  HistGrid& g = h->hist;
  g.enabled = 1; g.cylindrical = c.is_cylindrically_symmetric; g.nEn = c.n_energy_cells; g.nC = c.n_cos_cells; g.nR = c.n_radial_cells; g.nA = c.n_axial_cells;
  // Eigen::LinSpaced(size, low, high)[1] - [0] with size = cells + 1 (BMC.C:1863-1882)
  g.e_step = (0.0 + 1 * (max_eedf_energy - 0.0) / static_cast<double>(g.nEn)) - 0.0;
  g.c_first = -1.0; g.c_step = (-1.0 + 1 * (1.0 - (-1.0)) / static_cast<double>(g.nC)) - g.c_first;
  const double max_speed = std::sqrt(2.0 * max_eedf_energy * QE / ME);
  g.r_step = (0.0 + 1 * (max_speed - 0.0) / static_cast<double>(g.nR)) - 0.0;
  g.a_first = -max_speed; g.a_step = (-max_speed + 1 * (max_speed - (-max_speed)) / static_cast<double>(g.nA)) - g.a_first;
  h->max_eedf_energy = max_eedf_energy;
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
  CK(cudaSetDevice(h->cfg.device));
  HistGrid g = h->hist;
  g.eeh_phase = (phase_index >= 0) ? h->d_eeh_per + static_cast<size_t>(phase_index) * g.nEn : nullptr;
  k_sample<<<h->smp_blocks, ADV_THREADS, static_cast<size_t>(g.nEn) * 4 + 16, h->stream>>>(h->st, h->cfg.n_electrons, g, h->P, h->d_smp_part);
  ++h->launches;
  CK(cudaGetLastError());
  return 0;
}

int lokib200_fetch_histograms(lokib200_engine* h, double* eeh, double* eah, double* evh, double* eeh_periodic) {
  if (!h || !h->hist.enabled) return fail(h, LOKIB200_ERR_INVALID, "no histogram grid");
  CK(cudaSetDevice(h->cfg.device));
  const HistGrid& g = h->hist;
  auto fetch = [&](double* dst, const unsigned long long* src, size_t n) -> int {
    if (!dst) return 0;
    std::vector<unsigned long long> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), src, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) dst[i] = static_cast<double>(tmp[i]);
    return 0;
  };
  int rc;
  if ((rc = fetch(eeh, h->d_eeh, g.nEn)) || (rc = fetch(eah, h->d_eah, static_cast<size_t>(g.nEn) * g.nC)) ||
      (rc = fetch(evh, h->d_evh, static_cast<size_t>(g.nR) * g.nA)) || (rc = fetch(eeh_periodic, h->d_eeh_per, static_cast<size_t>(g.nEn) * h->cfg.n_phases)))
    return rc;
  return 0;
}

int lokib200_step_injected(lokib200_engine* h, int32_t n, const lokib200_electron* in, double nu_trial, const double* t_sync, const double* draws,
                           int32_t n_draws, lokib200_electron* out, lokib200_event_out* ev) {
  int rc = ensure_ready(h, true);
  if (rc) return rc;
  if (n <= 0 || !in || !t_sync || !draws || n_draws <= 0 || !out || !ev) return fail(h, LOKIB200_ERR_INVALID, "bad arguments");
  CK(cudaSetDevice(h->cfg.device));
  ElectronIO *d_in = nullptr, *d_out = nullptr; EventIO* d_ev = nullptr; double *d_ts = nullptr, *d_dr = nullptr;
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
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_ev); cudaFree(d_ts); cudaFree(d_dr);
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
