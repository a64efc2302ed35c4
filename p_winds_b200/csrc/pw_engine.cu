// pw_engine.cu -- sm_100a kernels and the C ABI (include/pwinds_b200.h) of the
// p-winds B200 forward-model engine.
//
// Kernel map (one retrieval-grid model = one (T0, Mdot, h_fraction) triple):
//   k_sed_setup     1 CTA       cross-section tables + SED integrals   (per context)
//   k_h_fields      1 CTA/model and pass: Parker + exact photoionisation integrals
//   k_h_solve       1 thread/model and pass: RK45 + mu-bar + relax-loop test
//   k_h_finish      1 CTA/model: outputs
//   k_helium        1 thread/model LSODA (Adams/BDF) + relax loop
//   k_he_density    elementwise  n(He 2^3S) / n(H I) and v in SI for the RT stage
//   k_rt_los        1 CTA/model LOS projection: column densities + broadening moments
//   k_rt_sum        (lambda tile, pixel split, model) Voigt + sum_pixels I0 exp(-tau)
//   k_rt_reduce     deterministic reduction of the pixel splits
//   k_pixel_table   per geometry: interp1d bracket/weights of every lit pixel
// No tensor cores: nothing on this path is a dense contraction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../../include/pwinds_b200.h"
#include "pw_common.cuh"
#include "pw_parker.cuh"
#include "pw_sed.cuh"
#include "pw_rk45.cuh"
#include "pw_lsoda.cuh"
#include "pw_hydrogen.cuh"
#include "pw_helium.cuh"
#include "pw_oxygen.cuh"
#include "pw_carbon.cuh"
#include "pw_heion.cuh"
#include "pw_transit.cuh"
#include "pw_draw.cuh"

using namespace pw;

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct pwb_ctx {
  int device = 0;
  int sm_count = 148;
  std::string err;
  long long launches = 0;
  // SED
  int sed_n = 0, sed_nh = 0;
  DevBuf sed_gw, sed_sh, sed_she, sed_scal;
  double sed_scal_host[PW_SC_COUNT] = {0};
  bool sed_ready = false;
  // LSODA coefficient tables
  DevBuf lsoda_tab;
  DevBuf exp_wide;   // 2^((i - 4096) / 256), i = 0 .. 4096 (pw_exp_neg_wide)
  // geometry / pixel table
  // One slot per transit phase (advanced_tutorial.ipynb cell 11 averages the spectrum over
  // `sample_phases`); slot 0 alone is the single-geometry case of transit.radiative_transfer_2d.
  static const int kMaxGeo = 16;
  struct Geo {
    DevBuf px_lo, px_whi, px_wlo, px_i0, px_meta;
    int npix = 0, n_active = 0;
    double s_out = 0.0;  // intensity of the lit pixels outside the radius grid (tau = 0)
    double depth = 0.0;  // geometric transit depth added to the spectrum of this phase
    bool ready = false;
  } geo[kMaxGeo];
  // chi^2 / log-likelihood of a model whose relax loop met an odeint warning (status 3): 0 =
  // failed model (+inf / -inf, the default: the reference's own result is undefined there),
  // 1 = score the spectrum of the last completed pass
  int accept_relax_warning = 0;
  int n_geo = 1;         // phases the pixel sum averages over
  int geo_nr = 0;
  DevBuf geo_r;
  // internal streams for the chunked forward pipeline
  // scratch
  DevBuf h_state, h_fields, he_order, he_key, he_work, o_work, rt_ncol, rt_sums, rt_partial, rt_tau, fw_f, fw_scal, fw_he, fw_n, fw_v, fw_T, misc;
};

static int fail(pwb_ctx* c, const char* fmt, const char* a = "") {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a);
  if (c) c->err = buf;
  return -1;
}
#define PWB_CUDA(c, call)                                                    \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) return fail((c), "CUDA error: %s", cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------
__global__ void k_sed_setup(const double* __restrict__ wl, const double* __restrict__ flux, int n,
                            double* gw, double* s_h, double* s_he, double* scalars) {
  __shared__ double red[40];
  sed_setup(wl, flux, n, gw, s_h, s_he, scalars, red);
}

__global__ void k_lsoda_tables(LsodaTables* t) {
  if (threadIdx.x == 0 && blockIdx.x == 0) lsoda_tables_init(t);
}

__global__ void k_parker(const double* __restrict__ r, int n, double* v, double* rho) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parker_point(r[i], &v[i], &rho[i]);
}

__global__ void k_parker_tidal(const double* __restrict__ r, int n, double cs, double rs, double mp,
                               double ms, double a, double* v, double* rho) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const TidalCoef c = tidal_coef(cs, rs, mp, ms, a);
  if (i < n) parker_point_tidal(c, r[i], &v[i], &rho[i]);
}

// ---- hydrogen stage: three kernels per batch instead of one CTA per model --------------
// k_h_fields  one CTA per model and pass: normalisation, Parker structure, Nr x N_lambda
//             photoionisation integrals; all 128 lanes busy.  Dynamic shared memory:
//             [3*n_h tables (optional)] [10*Nr work] [40 red]
// k_h_solve   one *thread* per model and pass: the RK45 solve (serial by nature), mu-bar,
//             convergence test; 32 independent models per warp
// k_h_finish  one CTA per model: outputs, Parker structure for the helium stage
// State between passes lives in global memory: HState per model + kHFieldsPerRadius*Nr doubles
// (rn, v, rho, phi, tau, f, fprev).  The host enqueues max_n_relax + 1 rounds; a finished
// model's CTA / thread returns at once.
constexpr int kHFieldsPerRadius = 10;  // rn, v, rho, phi, tau, f, fprev, and the three np.interp slope tables

__device__ __forceinline__ HParams h_params(const pwb_h_opts& o, double T, double mdot, double hfrac, double mu0) {
  HParams p;
  p.T = T; p.mdot = mdot; p.h_fraction = hfrac; p.mu0 = mu0;
  p.r_pl = o.planet_radius; p.m_pl = o.planet_mass; p.m_star = o.star_mass; p.a_au = o.semimajor_axis;
  p.initial_f_ion = o.initial_f_ion; p.convergence = o.convergence; p.max_n_relax = o.max_n_relax;
  p.relax = o.relax_solution; p.exact_phi = o.exact_phi; p.phi_abs = o.phi_abs; p.a0 = o.a_0;
  p.ivp.rtol = o.rtol; p.ivp.atol = o.atol; p.ivp.first_step = o.first_step; p.ivp.max_step = o.max_step;
  p.ivp.max_nfev = (o.max_nfev == 0) ? 2000000ll : (o.max_nfev < 0 ? 0ll : o.max_nfev);
  p.out_tidal = o.structure_tidal;
  return p;
}

__global__ void __launch_bounds__(128)
k_h_fields(int B, int pass, const double* __restrict__ T, const double* __restrict__ mdot,
           const double* __restrict__ hfrac, const double* __restrict__ mu0,
           const double* __restrict__ r, int nr, pwb_h_opts o, const double* __restrict__ gw,
           const double* __restrict__ s_h, const double* __restrict__ s_he, int n_h, int tables_in_smem,
           HState* __restrict__ state, double* __restrict__ fields) {
  extern __shared__ double smem[];
  __shared__ double s_e2[32];
  const int b = blockIdx.x;
  if (b >= B) return;
  if (threadIdx.x < 32) s_e2[threadIdx.x] = kExp2TabDev[threadIdx.x];
  const HParams p = h_params(o, T[b], mdot[b], hfrac[b], mu0[b]);
  if (pass == 0) {
    if (threadIdx.x == 0) h_state_init(p, &state[b]);
  } else if (state[b].done) {
    return;
  }
  double* base = smem;
  SedTables t;
  t.n_h = n_h; t.i1 = 0; t.i0 = 0;
  if (tables_in_smem) {
    double* tg = base; double* th = base + n_h; double* the = base + 2 * n_h;
    for (int j = threadIdx.x; j < n_h; j += blockDim.x) { tg[j] = gw[j]; th[j] = s_h[j]; the[j] = s_he[j]; }
    t.gw = tg; t.s_h = th; t.s_he = the;
    base += 3 * n_h;
  } else {
    t.gw = gw; t.s_h = s_h; t.s_he = s_he;
  }
  HWork w;
  w.rn = base; w.v = base + nr; w.rho = base + 2 * nr; w.phi = base + 3 * nr; w.tau = base + 4 * nr;
  w.f = base + 5 * nr; w.fprev = base + 6 * nr; w.colh = base + 7 * nr; w.colhe = base + 8 * nr;
  w.rcm = base + 9 * nr; w.red = base + 10 * nr;
  double* g = fields + (size_t)b * kHFieldsPerRadius * nr;
  if (pass > 0)
    for (int i = threadIdx.x; i < nr; i += blockDim.x) w.f[i] = g[5 * nr + i];
  __syncthreads();
  const double mu = (pass == 0) ? p.mu0 : state[b].mu;
  h_pass_fields(p, t, r, nr, w, pass, mu, (unsigned)__cvta_generic_to_shared(s_e2));
  for (int i = threadIdx.x; i < nr; i += blockDim.x) {
    g[i] = w.rn[i]; g[nr + i] = w.v[i]; g[2 * nr + i] = w.rho[i];
    g[3 * nr + i] = w.phi[i]; g[4 * nr + i] = w.tau[i];
    g[7 * nr + i] = w.colh[i]; g[8 * nr + i] = w.colhe[i]; g[9 * nr + i] = w.rcm[i];
  }
}

// hydrogen.ion_fraction(return_rates=True): one CTA per model, after the relax loop
__global__ void __launch_bounds__(128)
k_h_rates(int B, const double* __restrict__ T, const double* __restrict__ mdot, const double* __restrict__ hfrac,
          const double* __restrict__ mu0, const double* __restrict__ r, int nr, pwb_h_opts o,
          const double* __restrict__ gw, const double* __restrict__ s_h, const double* __restrict__ s_he, int n_h,
          const HState* __restrict__ state, const double* __restrict__ fields, double* __restrict__ rates) {
  extern __shared__ double smem[];
  __shared__ double s_e2[32];
  const int b = blockIdx.x;
  if (b >= B) return;
  if (threadIdx.x < 32) s_e2[threadIdx.x] = kExp2TabDev[threadIdx.x];
  const HParams p = h_params(o, T[b], mdot[b], hfrac[b], mu0[b]);
  double* out = rates + (size_t)b * 2 * nr;
  if (state[b].status != PW_OK) {
    for (int i = threadIdx.x; i < 2 * nr; i += blockDim.x) out[i] = nan("");
    return;
  }
  SedTables t;
  t.n_h = n_h; t.i1 = 0; t.i0 = 0; t.gw = gw; t.s_h = s_h; t.s_he = s_he;
  double* base = smem;
  HWork w;
  w.rn = base; w.v = base + nr; w.rho = base + 2 * nr; w.phi = base + 3 * nr; w.tau = base + 4 * nr;
  w.f = base + 5 * nr; w.fprev = base + 6 * nr; w.colh = base + 7 * nr; w.colhe = base + 8 * nr;
  w.rcm = base + 9 * nr; w.red = base + 10 * nr;
  const double* g = fields + (size_t)b * kHFieldsPerRadius * nr;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) { w.rho[i] = g[2 * nr + i]; w.f[i] = g[5 * nr + i]; }
  __syncthreads();
  h_rates(p, t, r, nr, w, state[b], out, (unsigned)__cvta_generic_to_shared(s_e2));
}

__global__ void __launch_bounds__(32)
k_h_solve(int B, int lanes, int pass, const double* __restrict__ T, const double* __restrict__ mdot,
          const double* __restrict__ hfrac, const double* __restrict__ mu0, const double* __restrict__ r,
          int nr, pwb_h_opts o, HState* __restrict__ state, double* __restrict__ fields) {
  if ((int)threadIdx.x >= lanes) return;
  const int b = blockIdx.x * lanes + threadIdx.x;
  if (b >= B) return;
  HState s = state[b];
  if (s.done) return;
  const HParams p = h_params(o, T[b], mdot[b], hfrac[b], mu0[b]);
  double* g = fields + (size_t)b * kHFieldsPerRadius * nr;
  HWork w;
  w.rn = g; w.v = g + nr; w.rho = g + 2 * nr; w.phi = g + 3 * nr; w.tau = g + 4 * nr;
  w.f = g + 5 * nr; w.fprev = g + 6 * nr; w.colh = g + 7 * nr; w.colhe = g + 8 * nr; w.rcm = g + 9 * nr;
  w.red = nullptr;
  h_pass_solve(p, r, nr, w, pass, &s);
  state[b] = s;
}

__global__ void __launch_bounds__(128)
k_h_finish(int B, const double* __restrict__ T, const double* __restrict__ mdot,
           const double* __restrict__ hfrac, const double* __restrict__ mu0, const double* __restrict__ r,
           int nr, pwb_h_opts o, const HState* __restrict__ state, const double* __restrict__ fields,
           double* f_r, double* v, double* rho, double* scal, int* status) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const HParams p = h_params(o, T[b], mdot[b], hfrac[b], mu0[b]);
  HOut out{f_r + (size_t)b * nr, v + (size_t)b * nr, rho + (size_t)b * nr, scal + (size_t)b * 8, status + b};
  h_finish(p, r, nr, fields + (size_t)b * kHFieldsPerRadius * nr + 5 * nr, state[b], out);
}

// Cost-ordered schedule of the helium stage.  A model's cost (LSODA steps x relax passes)
// grows ~5x with the column density of the wind, and the kernel is a batch of independent
// serial solves: what matters is when the *last* one ends.  k_he_key computes a cost proxy --
// the total column density of the wind, rho_s * sum_i dr_i rho_i, which sets the helium optical
// depths and with them the number of relax passes (rank correlation with the measured cost
// 0.95 on the 64 x 64 grid; rho_s * r_s alone: 0.92 and the wrong models on top) -- and
// k_he_order sorts by it (descending; bucket sort in one CTA, order inside a bucket is
// irrelevant).  The launch then walks the sorted list in up to four segments with a
// growing number of models per warp: the expensive head runs one model per warp (no
// divergence, scheduled first), the cheap tail is packed (its warps take as long as the
// head's even though the lanes of a warp serialise where their paths differ).  Placement
// never changes a model's result: models share nothing.
struct HeSched {
  int nseg;
  int slot0[5];  // first sorted slot of segment s (slot0[nseg] = B)
  int warp0[5];  // first warp (CTA) of segment s (warp0[nseg] = number of CTAs)
  int lanes[4];  // models per warp in segment s
};

__global__ void k_he_key(int B, const double* __restrict__ rhos, int sstride, const double* __restrict__ r, int nr,
                         const double* __restrict__ rho, const int* __restrict__ status, double* __restrict__ key) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double k = 0.0;
  if (status[b] == PW_OK) {
    const double* q = rho + (size_t)b * nr;
    for (int i = 0; i < nr; ++i) {
      const double dr = (i < nr - 1) ? (r[i + 1] - r[i]) : (r[nr - 1] - r[nr - 2]);
      k += dr * q[i];
    }
    k *= rhos[(size_t)b * sstride];
  }
  key[b] = k;
}

constexpr int kHeBuckets = 4096;
__global__ void __launch_bounds__(1024, 1)
k_he_order(int B, const double* __restrict__ key, const int* __restrict__ status, int* __restrict__ order, int coarse) {
  __shared__ int hist[kHeBuckets];
  __shared__ unsigned long long s_lo, s_hi;
  const int t = threadIdx.x;
  auto key_bits = [&](int b) -> unsigned long long {
    if (status[b] != PW_OK) return 0ull;  // nothing to integrate: last
    const double k = key[b];
    return (k > 0.0) ? (unsigned long long)__double_as_longlong(k) : 1ull;  // positive doubles order like their bits
  };
  if (t == 0) { s_lo = ~0ull; s_hi = 0ull; }
  for (int i = t; i < kHeBuckets; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  // bucket range over the models that will be integrated: a failed one (key 0) would stretch the range
  // to zero and squeeze every valid key into the top percent of the buckets
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int b = t; b < B; b += blockDim.x) {
    const unsigned long long k = key_bits(b);
    if (k == 0ull) continue;
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
  atomicMin(&s_lo, lo);
  atomicMax(&s_hi, hi);
  __syncthreads();
  lo = s_lo; hi = s_hi;
  const double scale = (hi > lo) ? (double)(kHeBuckets - 1) / (double)(hi - lo) : 0.0;
  auto bucket = [&](int b) {   // 0 = costliest; failed models last
    const unsigned long long k = key_bits(b);
    const int q = (k == 0ull || k < lo) ? kHeBuckets - 1 : (kHeBuckets - 1) - (int)((double)(k - lo) * scale);
    return (q / coarse) * coarse;   // `coarse` buckets merged into one class (order inside a class: as it falls)
  };
  for (int b = t; b < B; b += blockDim.x) atomicAdd(&hist[bucket(b)], 1);
  __syncthreads();
  // exclusive scan of the histogram (4 buckets per thread + block scan through shared memory)
  {
    const int per = kHeBuckets / 1024;
    int loc[per], sum = 0;
    for (int q = 0; q < per; ++q) { loc[q] = hist[t * per + q]; sum += loc[q]; }
    __shared__ int part[1024];
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const int v = (t >= off) ? part[t - off] : 0;
      __syncthreads();
      part[t] += v;
      __syncthreads();
    }
    int run = part[t] - sum;
    for (int q = 0; q < per; ++q) { hist[t * per + q] = run; run += loc[q]; }
  }
  __syncthreads();
  for (int b = t; b < B; b += blockDim.x) order[atomicAdd(&hist[bucket(b)], 1)] = b;
}

#ifndef PW_HE_MINBLOCKS
#define PW_HE_MINBLOCKS 1   // CTAs (= warps) per SM the register allocation has to leave room for
#endif
template <bool kRk45>
__global__ void __launch_bounds__(32, PW_HE_MINBLOCKS)
k_helium(int B, HeSched sched, const int* __restrict__ order, const double* __restrict__ T, const double* __restrict__ hfrac,
         const double* __restrict__ vs, const double* __restrict__ rs, const double* __restrict__ rhos,
         int sstride, const double* __restrict__ r, int nr, const double* __restrict__ v,
         const double* __restrict__ rho, const double* __restrict__ f_h, pwb_he_opts o,
         const LsodaTables* __restrict__ tab, double* work, size_t wstride, double* f1, double* f3, double* scal,
         double* rates, int* status) {
  __shared__ double s_e2[32];
  s_e2[threadIdx.x] = kExp2TabDev[threadIdx.x];
  __syncwarp();
  const int lane = threadIdx.x, warp = blockIdx.x;
  int seg = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (q < sched.nseg && warp >= sched.warp0[q]) seg = q;
  if (lane >= sched.lanes[seg]) return;
  const int slot = sched.slot0[seg] + (warp - sched.warp0[seg]) * sched.lanes[seg] + lane;
  if (slot >= sched.slot0[seg + 1]) return;
  const int b = order ? order[slot] : slot;
  if (status[b] != PW_OK) {  // upstream failure: propagate NaNs
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) { f1[(size_t)b * nr + i] = nan_; f3[(size_t)b * nr + i] = nan_; }
    for (int k = 0; k < 4; ++k) scal[(size_t)b * 4 + k] = 0.0;
    return;
  }
  HeParams p;
  p.e2tab = (unsigned)__cvta_generic_to_shared(s_e2);
  p.T = T[b]; p.h_fraction = hfrac[b];
  p.vs = vs[(size_t)b * sstride]; p.rs = rs[(size_t)b * sstride]; p.rhos = rhos[(size_t)b * sstride];
  p.r_pl = o.planet_radius; p.y0[0] = o.initial_state[0]; p.y0[1] = o.initial_state[1];
  p.relax = o.relax_solution; p.max_n_relax = o.max_n_relax; p.convergence = o.convergence;
  for (int k = 0; k < 6; ++k) p.rates[k] = o.rates[k];
  p.method = o.method; p.rtol = o.rtol; p.atol = o.atol; p.first_step = o.first_step; p.max_step = o.max_step;
  double* wb = work + (size_t)b * wstride;
  HeWork w{wb, wb + nr, wb + 2 * nr, wb + 3 * nr, wb + 4 * nr, wb + 5 * nr, wb + 10 * nr};
  HeOut out{f1 + (size_t)b * nr, f3 + (size_t)b * nr, scal + (size_t)b * 4,
            rates ? rates + (size_t)b * 11 * nr : nullptr, status + b};
#if defined(PW_HE_TIMING)
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
  he_population_model<kRk45>(p, *tab, r, v + (size_t)b * nr, rho + (size_t)b * nr, f_h + (size_t)b * nr, nr, w, out);
#if defined(PW_HE_TIMING)   // developer build: per-model start / end time (ns) in the two spare slots
  unsigned long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  out.scal[0] = (double)(t0 % 100000000000ull);
  out.scal[3] = (double)(t1 % 100000000000ull);
#endif
}

#if defined(PW_HE_PS_KERNEL)
// Experiment (developer build): the same stage with the lanes of a warp aligned block by block
// (he_population_ps): every lane of the warp stays in the kernel until the warp's last model is done,
// because the warp votes on the block to run next.
__global__ void __launch_bounds__(32, PW_HE_MINBLOCKS)
k_helium_ps(int B, HeSched sched, const int* __restrict__ order, const double* __restrict__ T,
            const double* __restrict__ hfrac, const double* __restrict__ vs, const double* __restrict__ rs,
            const double* __restrict__ rhos, int sstride, const double* __restrict__ r, int nr,
            const double* __restrict__ v, const double* __restrict__ rho, const double* __restrict__ f_h, pwb_he_opts o,
            const LsodaTables* __restrict__ tab, double* work, size_t wstride, double* f1, double* f3, double* scal,
            double* rates, int* status) {
  __shared__ double s_e2[32];
  s_e2[threadIdx.x] = kExp2TabDev[threadIdx.x];
  __syncwarp();
  const int lane = threadIdx.x, warp = blockIdx.x;
  int seg = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (q < sched.nseg && warp >= sched.warp0[q]) seg = q;
  const int slot = sched.slot0[seg] + (warp - sched.warp0[seg]) * sched.lanes[seg] + lane;
  bool has = lane < sched.lanes[seg] && slot < sched.slot0[seg + 1];
  int b = 0;
  if (has) {
    b = order ? order[slot] : slot;
    if (status[b] != PW_OK) {  // upstream failure: propagate NaNs
      const double nan_ = nan("");
      for (int i = 0; i < nr; ++i) { f1[(size_t)b * nr + i] = nan_; f3[(size_t)b * nr + i] = nan_; }
      for (int k = 0; k < 4; ++k) scal[(size_t)b * 4 + k] = 0.0;
      has = false;
    }
  }
  HeParams p;
  p.e2tab = (unsigned)__cvta_generic_to_shared(s_e2);
  p.T = T[b]; p.h_fraction = hfrac[b];
  p.vs = vs[(size_t)b * sstride]; p.rs = rs[(size_t)b * sstride]; p.rhos = rhos[(size_t)b * sstride];
  p.r_pl = o.planet_radius; p.y0[0] = o.initial_state[0]; p.y0[1] = o.initial_state[1];
  p.relax = o.relax_solution; p.max_n_relax = o.max_n_relax; p.convergence = o.convergence;
  for (int k = 0; k < 6; ++k) p.rates[k] = o.rates[k];
  p.method = o.method; p.rtol = o.rtol; p.atol = o.atol; p.first_step = o.first_step; p.max_step = o.max_step;
  double* wb = work + (size_t)b * wstride;
  HeWork w{wb, wb + nr, wb + 2 * nr, wb + 3 * nr, wb + 4 * nr, wb + 5 * nr, wb + 10 * nr};
  HeOut out{f1 + (size_t)b * nr, f3 + (size_t)b * nr, scal + (size_t)b * 4,
            rates ? rates + (size_t)b * 11 * nr : nullptr, status + b};
  he_population_ps(p, *tab, r, v + (size_t)b * nr, rho + (size_t)b * nr, f_h + (size_t)b * nr, nr, w, out, has);
}
#endif

// oxygen.ion_fraction: one thread per model (Radau / LSODA / RK45 replica), `lanes` models per warp
__global__ void __launch_bounds__(32)
k_oxygen(int B, int lanes, const double* __restrict__ T, const double* __restrict__ hfrac,
         const double* __restrict__ vs, const double* __restrict__ rs, const double* __restrict__ rhos,
         int sstride, const double* __restrict__ r, int nr, const double* __restrict__ v,
         const double* __restrict__ rho, const double* __restrict__ f_h, const double* __restrict__ f_he,
         pwb_o_opts o, const LsodaTables* __restrict__ tab, double* work, double* f_o, double* scal,
         double* rates, int* status) {
  if ((int)threadIdx.x >= lanes) return;
  const int b = blockIdx.x * lanes + threadIdx.x;
  if (b >= B) return;
  if (status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX) {  // upstream failure: propagate NaNs
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) f_o[(size_t)b * nr + i] = nan_;
    for (int k = 0; k < 4; ++k) scal[(size_t)b * 4 + k] = 0.0;
    return;
  }
  OParams p;
  p.T = T[b]; p.h_fraction = hfrac[b];
  p.vs = vs[(size_t)b * sstride]; p.rs = rs[(size_t)b * sstride]; p.rhos = rhos[(size_t)b * sstride];
  p.r_pl = o.planet_radius; p.o_fraction = o.o_fraction; p.y0 = o.initial_f_o_ion;
  p.relax = o.relax_solution; p.max_n_relax = o.max_n_relax; p.convergence = o.convergence;
  for (int k = 0; k < 4; ++k) p.rates[k] = o.rates[k];
  p.method = o.method; p.rtol = o.rtol; p.atol = o.atol; p.first_step = o.first_step; p.max_step = o.max_step;
  double* wb = work + (size_t)b * kOWorkPerRadius * nr;
  OWork w{wb, wb + nr, wb + 2 * nr, wb + 3 * nr, wb + 4 * nr, wb + 5 * nr, wb + 6 * nr};
  OOut out{f_o + (size_t)b * nr, scal + (size_t)b * 4, rates ? rates + (size_t)b * 5 * nr : nullptr, status + b};
  int st = PW_OK;
  OOut out2 = out;
  out2.status = &st;
  o_ion_fraction_model(p, *tab, r, v + (size_t)b * nr, rho + (size_t)b * nr, f_h + (size_t)b * nr,
                       f_he + (size_t)b * nr, nr, w, out2);
  if (st != PW_OK) status[b] = st;  // keeps an upstream PW_WARN_HE_RELAX otherwise
}

// helium.ion_fraction: one thread per model
__global__ void __launch_bounds__(32)
k_he_ion(int B, int lanes, const double* __restrict__ T, const double* __restrict__ hfrac,
         const double* __restrict__ vs, const double* __restrict__ rs, const double* __restrict__ rhos,
         int sstride, const double* __restrict__ r, int nr, const double* __restrict__ v,
         const double* __restrict__ rho, const double* __restrict__ f_h, pwb_o_opts o,
         const LsodaTables* __restrict__ tab, double* work, double* f_he, double* scal, int* status) {
  if ((int)threadIdx.x >= lanes) return;
  const int b = blockIdx.x * lanes + threadIdx.x;
  if (b >= B) return;
  if (status[b] != PW_OK) {
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) f_he[(size_t)b * nr + i] = nan_;
    for (int k = 0; k < 4; ++k) scal[(size_t)b * 4 + k] = 0.0;
    return;
  }
  HeIonParams p;
  p.T = T[b]; p.h_fraction = hfrac[b];
  p.vs = vs[(size_t)b * sstride]; p.rs = rs[(size_t)b * sstride]; p.rhos = rhos[(size_t)b * sstride];
  p.r_pl = o.planet_radius; p.y0 = o.initial_f_o_ion;
  p.relax = o.relax_solution; p.max_n_relax = o.max_n_relax; p.convergence = o.convergence;
  for (int k = 0; k < 3; ++k) p.rates[k] = o.rates[k];
  p.method = o.method; p.rtol = o.rtol; p.atol = o.atol; p.first_step = o.first_step; p.max_step = o.max_step;
  double* wb = work + (size_t)b * kOWorkPerRadius * nr;
  OWork w{wb, wb + nr, wb + 2 * nr, wb + 3 * nr, wb + 4 * nr, wb + 5 * nr, wb + 6 * nr};
  int st = PW_OK;
  he_ion_fraction_model(p, *tab, r, v + (size_t)b * nr, rho + (size_t)b * nr, f_h + (size_t)b * nr, nr, w,
                        f_he + (size_t)b * nr, scal + (size_t)b * 4, &st);
  if (st != PW_OK) status[b] = st;
}

// carbon.ion_fraction: one thread per model (LSODA / Radau / RK45 replica), `lanes` models per warp
__global__ void __launch_bounds__(32)
k_carbon(int B, int lanes, const double* __restrict__ T, const double* __restrict__ hfrac,
         const double* __restrict__ vs, const double* __restrict__ rs, const double* __restrict__ rhos,
         int sstride, const double* __restrict__ r, int nr, const double* __restrict__ v,
         const double* __restrict__ rho, const double* __restrict__ f_h, const double* __restrict__ f_he,
         pwb_c_opts o, const LsodaTables* __restrict__ tab, double* work, double* f_cii, double* f_ciii,
         double* scal, double* rates, int* status) {
  if ((int)threadIdx.x >= lanes) return;
  const int b = blockIdx.x * lanes + threadIdx.x;
  if (b >= B) return;
  if (status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX) {  // upstream failure: propagate NaNs
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) { f_cii[(size_t)b * nr + i] = nan_; f_ciii[(size_t)b * nr + i] = nan_; }
    for (int k = 0; k < 4; ++k) scal[(size_t)b * 4 + k] = 0.0;
    return;
  }
  CParams p;
  p.T = T[b]; p.h_fraction = hfrac[b];
  p.vs = vs[(size_t)b * sstride]; p.rs = rs[(size_t)b * sstride]; p.rhos = rhos[(size_t)b * sstride];
  p.r_pl = o.planet_radius; p.c_fraction = o.c_fraction; p.y0[0] = o.initial_f_c_ion[0]; p.y0[1] = o.initial_f_c_ion[1];
  p.relax = o.relax_solution; p.max_n_relax = o.max_n_relax; p.convergence = o.convergence;
  for (int k = 0; k < 7; ++k) p.rates[k] = o.rates[k];
  p.method = o.method; p.rtol = o.rtol; p.atol = o.atol; p.first_step = o.first_step; p.max_step = o.max_step;
  double* wb = work + (size_t)b * kCWorkPerRadius * nr;
  CWork w{wb, wb + nr, wb + 2 * nr, wb + 3 * nr, wb + 4 * nr, wb + 5 * nr, wb + 6 * nr, wb + 7 * nr, wb + 8 * nr,
          wb + 9 * nr};
  int st = PW_OK;
  COut out{f_cii + (size_t)b * nr, f_ciii + (size_t)b * nr, scal + (size_t)b * 4,
           rates ? rates + (size_t)b * 11 * nr : nullptr, &st};
  c_ion_fraction_model(p, *tab, r, v + (size_t)b * nr, rho + (size_t)b * nr, f_h + (size_t)b * nr,
                       f_he + (size_t)b * nr, nr, w, out);
  if (st != PW_OK) status[b] = st;
}

// n [m^-3] and v [m/s] for the RT stage (quickstart.ipynb cells 10-11; SURVEY 8(d) cfg 2)
__global__ void k_he_density(int B, int nr, int species, const double* __restrict__ hfrac,
                             const double* __restrict__ hscal, const double* __restrict__ f_r,
                             const double* __restrict__ f3, const double* __restrict__ rho,
                             const double* __restrict__ v, const double* __restrict__ T,
                             double* n_out, double* v_out, double* T_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * nr) return;
  const int b = (int)(idx / nr);
  const double h = hfrac[b], he = 1 - h;
  const double vs = hscal[(size_t)b * 8 + 1], rhos = hscal[(size_t)b * 8 + 3];
  double n;
  if (species == 1) {
    n = (1 - f_r[idx]) * rho[idx] * rhos * h / (h + 4 * he) / 1.67262192369e-24 * 1E6;
  } else {
    const double n_he = rho[idx] * rhos * he / (h + 4 * he) / 1.67262192369e-24;
    n = (f3[idx] * n_he) * 1E6;
  }
  n_out[idx] = n;
  v_out[idx] = v[idx] * vs * 1000;
  if (T_out && idx % nr == 0) T_out[b] = T[b];
}

// Per-geometry pixel table: for every pixel with I0 != 0 and r inside the radius grid,
// the interp1d bracket and weights; everything else contributes I0 * exp(0).
// meta[0] = number of active pixels, meta[1] = sum of I0 over the inactive ones.
__global__ void k_pixel_table(const double* __restrict__ i0, const double* __restrict__ rmap, int npix,
                              const double* __restrict__ r, int nr, int* lo, double* whi, double* wlo,
                              double* pi0, double* meta) {
  __shared__ int s_cnt[1024];
  __shared__ int s_base;
  __shared__ double red[40];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_base = 0;
  double s_out = 0.0;
  __syncthreads();
  for (int start = 0; start < npix; start += nt) {
    const int p = start + tid;
    int act = 0;
    double x = 0, in = 0;
    if (p < npix) {
      in = i0[p]; x = rmap[p];
      const bool inside = !(x < r[0] || x > r[nr - 1]) && (x == x);
      if (in != 0.0) { if (inside) act = 1; else s_out += in; }
    }
    s_cnt[tid] = act;
    __syncthreads();
    // inclusive scan (Hillis-Steele), deterministic order = pixel order
    for (int off = 1; off < nt; off <<= 1) {
      int vtmp = (tid >= off) ? s_cnt[tid - off] : 0;
      __syncthreads();
      s_cnt[tid] += vtmp;
      __syncthreads();
    }
    if (act) {
      const int q = s_base + s_cnt[tid] - 1;
      const int hi = interp1d_hi(r, nr, x);
      const double dx = r[hi] - r[hi - 1];
      lo[q] = hi - 1;
      whi[q] = (x - r[hi - 1]) / dx;
      wlo[q] = (r[hi] - x) / dx;
      pi0[q] = in;
    }
    __syncthreads();
    if (tid == nt - 1) s_base += s_cnt[tid];
    __syncthreads();
  }
  const double tot = block_sum(s_out, red);
  if (tid == 0) { meta[0] = (double)s_base; meta[1] = tot; }
}

__global__ void __launch_bounds__(256)
k_rt_los(int B, const double* __restrict__ r, const double* __restrict__ n, const double* __restrict__ v,
         const double* __restrict__ Tprof, int t_is_profile, int nr, int nz, const int* __restrict__ status,
         double* ncol, double* sums) {
  extern __shared__ double sm[];  // r, n, v, (T) staged: 4*nr + 40; one row of nz cells per warp
  const int b = blockIdx.x;
  if (b >= B) return;
  if (status && status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX) return;
  double* sr = sm; double* sn = sm + nr; double* sv = sm + 2 * nr; double* st = sm + 3 * nr;
  double* red = sm + 4 * nr;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) {
    sr[i] = r[i]; sn[i] = n[(size_t)b * nr + i]; sv[i] = v[(size_t)b * nr + i];
    st[i] = t_is_profile ? Tprof[(size_t)b * nr + i] : 0.0;
  }
  __syncthreads();
  rt_los_average(sr, sn, sv, t_is_profile ? st : nullptr, nr, nz, ncol + (size_t)b * nr, sums + (size_t)b * 4, red,
                 sm + 4 * nr + 40);
}


// 'formal' wind broadening (transit.py:463-473, :512-519): per LOS cell Doppler width
// and wind shift, one Voigt evaluation per (z, b, lambda, line).  grid = (Nr, B); the CTA
// stages the Nz cells of its impact parameter in shared memory, threads own wavelengths.
__global__ void __launch_bounds__(256)
k_rt_formal_tau(int B, int nw, int nr, pwb_rt_opts o, const double* __restrict__ r, const double* __restrict__ n,
                const double* __restrict__ v, const double* __restrict__ T, const double* __restrict__ wl,
                const int* __restrict__ status, double* tau /*[B][nr][nw]*/) {
  extern __shared__ double sm[];  // r, n, v, T (4*nr) | n_c, v_los, T_c, z (4*nz)
  const int i = blockIdx.x, b = blockIdx.y;
  if (status && status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX) return;
  const int nz = o.z_grid_size;
  double* sr = sm; double* sn = sm + nr; double* sv = sm + 2 * nr; double* st = sm + 3 * nr;
  double* cn = sm + 4 * nr; double* cv = cn + nz; double* ct = cv + nz; double* cz = ct + nz;
  const bool tp = o.temperature_is_profile != 0;
  for (int q = threadIdx.x; q < nr; q += blockDim.x) {
    sr[q] = r[q]; sn[q] = n[(size_t)b * nr + q]; sv[q] = v[(size_t)b * nr + q];
    st[q] = tp ? T[(size_t)b * nr + q] : 0.0;
  }
  __syncthreads();
  const double r_top = sr[nr - 1], bb = sr[i];
  for (int j = threadIdx.x; j < nz; j += blockDim.x) {
    const double z = los_z(r_top, nz, j);
    double nc, vl, tc = 0.0;
    los_cell(sr, sn, sv, tp ? st : nullptr, nr, bb, z, &nc, &vl, tp ? &tc : nullptr);
    cn[j] = nc; cv[j] = vl; ct[j] = tc; cz[j] = z;
  }
  __syncthreads();
  const double c_speed = 2.99792458e+08, k_b = 1.380649e-23;
  const double t_scalar = tp ? 0.0 : T[b];
  for (int k = threadIdx.x; k < nw; k += blockDim.x) {
    const double nu_rest = c_speed / wl[k];
    double tot = 0.0;
    for (int line = 0; line < o.n_lines; ++line) {
      const double nu0 = c_speed / o.central_wavelength[line];
      const double gamma = o.einstein_coefficient[line] / 4 / kPi;
      const double dnu = nu_rest - nu0;
      const double alpha0 = nu0 / c_speed * sqrt(k_b * t_scalar / o.particle_mass);
      double acc = 0.0, yprev = 0.0;
      for (int j = 0; j < nz; ++j) {
        double y = 0.0;
        if (cn[j] != 0.0) {
          const double alpha = tp ? nu0 / c_speed * sqrt(k_b * ct[j] / o.particle_mass) : alpha0;
          const double add = (cv[j] + o.bulk_los_velocity + o.planet_radial_velocity) / c_speed * nu0;
          y = voigt_profile(dnu + add, alpha, gamma) * cn[j];
        }
        if (j > 0) acc += (cz[j] - cz[j - 1]) * (y + yprev) / 2.0;
        yprev = y;
      }
      tot += acc * (2.654008854574474e-06 * o.oscillator_strength[line]);
    }
    tau[((size_t)b * nr + i) * nw + k] = tot;
  }
}

// Pixel sum of the 'formal' broadening: tau(b, lambda) comes from the table k_rt_formal_tau made.
// grid = (B, pixel segments, lambda tiles), thread <-> wavelength bin; every thread accumulates
// sum_p I0_p exp(-tau_p(lambda)) over its segment of the active pixels, k_rt_reduce adds the segments.
// (The 'average' method, where tau = N_p Phi(lambda) separates, runs in k_rt_sum2 below.)
__global__ void __launch_bounds__(256)
k_rt_sum(int B, int nw, int nr, pwb_rt_opts o, const double* __restrict__ wl, const double* __restrict__ T,
         const double* __restrict__ ncol, const double* __restrict__ sums, const int* __restrict__ status,
         const int* __restrict__ px_lo, const double* __restrict__ px_whi, const double* __restrict__ px_wlo,
         const double* __restrict__ px_i0, const double* __restrict__ meta, int nsplit, double* out /*[B][nsplit][nw]*/,
         double* tau_out /*[B][nr][nw] or null*/, const double* __restrict__ tau_tab /*'formal': [B][nr][nw]*/) {
  __shared__ double s_e2[64];
  if (threadIdx.x < 64) s_e2[threadIdx.x] = kExp2Tab64Dev[threadIdx.x];
  const unsigned e2 = (unsigned)__cvta_generic_to_shared(s_e2);
  const int b = blockIdx.x, split = blockIdx.y;
  const int k = blockIdx.z * blockDim.x + threadIdx.x;  // wavelength index
  const bool bad = status && status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX;
  if (bad) {
    if (k < nw) out[((size_t)b * nsplit + split) * nw + k] = (split == 0) ? nan("") : 0.0;
    return;
  }
  if (tau_tab) {
    // 'formal' broadening: tau(b, lambda) was tabulated by k_rt_formal_tau; interpolate it
    // per pixel (transit.py:220-222) -- two table reads + exp per (pixel, lambda)
    const int n_act = (int)meta[0];
    const int per = (n_act + nsplit - 1) / nsplit;
    const int p0 = split * per, p1 = min(n_act, p0 + per);
    double acc = 0.0;
    __syncthreads();
    if (k < nw) {
      const double* tb = tau_tab + (size_t)b * nr * nw + k;
      for (int q = p0; q < p1; ++q) {
        const int lo = px_lo[q];
        const double t = px_whi[q] * tb[(size_t)(lo + 1) * nw] + px_wlo[q] * tb[(size_t)lo * nw];
        acc = pw_fma(px_i0[q], pw_exp_neg_sum_sh(-t, e2), acc);
      }
      if (split == 0) acc += meta[1];
      out[((size_t)b * nsplit + split) * nw + k] = acc;
      if (tau_out && split == 0)
        for (int i = 0; i < nr; ++i) tau_out[((size_t)b * nr + i) * nw + k] = tb[(size_t)i * nw];
    }
    return;
  }
}

// ---- pixel sum, 'average' broadening -----------------------------------------------------------
// spectrum(lambda) = mean over phases of [ sum_p I0_p exp(-N_p Phi(lambda)) + depth ]  (transit.py:220-228,
// advanced_tutorial.ipynb cell 11).  The active pixels of a phase are cut into the same S <= 16
// segments whatever the batch (rt_nsplit), every segment is summed sequentially and the segment sums
// are combined in segment order: a model's spectrum is bit-identical alone or inside any grid.
//
// Layout: a CTA owns one (model, wavelength tile) and G <= 4 pixel groups of TPG threads; a thread owns
// TWO adjacent wavelength bins (2u, 2u + 1) of the tile, so one 16-byte shared-memory read (N_p, I0_p) feeds two
// independent exp chains, no warp idles at Nw = 200 (2 x 100 threads), and a thread's two results leave in
// one 16-byte store: a warp writes 512 contiguous bytes of the spectrum.
//   ext == 1  the CTA sums all S segments of all phases: group g takes segments g, g + G, ...; after
//             each round the G segment sums meet in shared memory and group 0 adds them in segment
//             order; depth, phase mean and the spectrum store follow -- no partial spectra in HBM, no
//             reduce kernel.
//   ext == S  (few models: fewer CTAs than SMs otherwise) blockIdx.z = segment, one phase per launch;
//             the segment sums go to `part` [B][S][nw] and k_rt_reduce adds them in the same order.
// Pixels are staged per group in chunks with their interpolated column density N_p (interp1d weights of
// pwb_set_geometry).  Per chunk the group checks the largest tau it can meet (largest N_p times the
// tile's largest Phi): below kExpWideCap the terms go through pw_exp_neg_wide (10 FP64 + 2 other
// instructions), otherwise through pw_exp_neg_sum_sh (any tau, exact zero below e^-708).
struct RtGeoSlots {
  int n_geo;
  const int* lo[16];
  const double* whi[16];
  const double* wlo[16];
  const double* i0[16];
  const double* meta[16];
  double depth[16];
};

// two adjacent results of one thread: one 16-byte store where both exist and the address allows it
__device__ __forceinline__ void store_pair(double* p, double a, double b, bool va, bool vb) {
  if (va && vb && (reinterpret_cast<size_t>(p) & 15) == 0) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
  } else {
    if (va) p[0] = a;
    if (vb) p[1] = b;
  }
}

template <int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads, (kMaxThreads <= 256) ? 3 : (kMaxThreads <= 512 ? 2 : 1))
k_rt_sum2(int B, int nw, int nr, const __grid_constant__ pwb_rt_opts o, const double* __restrict__ wl,
          const double* __restrict__ T, const double* __restrict__ ncol, const double* __restrict__ sums,
          const int* __restrict__ status, const __grid_constant__ RtGeoSlots geo, int geo0, int S, int ext, int G,
          int TPG, int Wt, int chunk, const double* __restrict__ exp_wide, double* __restrict__ spec /*[B][nw]*/,
          double* __restrict__ part /*[B][S][nw], ext > 1*/, double* tau_out /*[B][nr][nw] or null*/) {
  extern __shared__ double sm2[];
  __shared__ double s_e2[64];
  __shared__ int s_flag[4];
  double* s_wide = sm2;                       // [kExpWideN + 1]  2^(i/256 - 16)
  double* s_ncol = s_wide + kExpWideN + 1;    // [nr]
  double* s_phi = s_ncol + nr;                // [Wt]
  double* s_x = s_phi + Wt;                   // [G][Wt] segment sums of one round
  double* s_px = s_x + (size_t)G * Wt;        // [G][chunk] pairs (N_p, I0_p); 16-byte aligned (even offsets + 1)
  if ((reinterpret_cast<size_t>(s_px) & 15) != 0) s_px += 1;
  const int t = threadIdx.x, b = blockIdx.x, tile = blockIdx.y;
  const int g = t / TPG, u = t - g * TPG;
  const int kbase = tile * Wt;
  const int wt = min(Wt, nw - kbase);         // bins of this tile
  const bool bad = status && status[b] != PW_OK && status[b] != PW_WARN_HE_RELAX;
  if (bad) {
    if (ext == 1) {
      for (int k = t; k < wt; k += blockDim.x) spec[(size_t)b * nw + kbase + k] = nan("");
    } else {
      for (int k = t; k < wt; k += blockDim.x)
        part[((size_t)b * S + blockIdx.z) * nw + kbase + k] = (blockIdx.z == 0) ? nan("") : 0.0;
    }
    return;
  }
  for (int i = t; i < 64; i += blockDim.x) s_e2[i] = kExp2Tab64Dev[i];   // (a tile of < 128 bins has < 64 threads)
  for (int i = t; i <= kExpWideN; i += blockDim.x) s_wide[i] = exp_wide[i];
  for (int i = t; i < nr; i += blockDim.x) s_ncol[i] = ncol[(size_t)b * nr + i];
  {  // Phi(lambda) of the tile, 'average' broadening (transit.py:479-513), one bin per thread
    const double s_n = sums[(size_t)b * 4], s_v2n = sums[(size_t)b * 4 + 1], s_tn = sums[(size_t)b * 4 + 2];
    const double vw2 = s_v2n / s_n;
    const double temp = o.temperature_is_profile ? s_tn / s_n : T[b];
    for (int k = t; k < wt; k += blockDim.x)
      s_phi[k] = rt_phi_average_raw(o.central_wavelength, o.oscillator_strength, o.einstein_coefficient, o.n_lines,
                                    o.particle_mass, o.bulk_los_velocity, o.planet_radial_velocity,
                                    o.turbulence_broadening, wl[kbase + k], temp, vw2);
  }
  __syncthreads();
  if (tau_out && (ext == 1 || blockIdx.z == 0))
    for (int k = t; k < wt; k += blockDim.x) {
      const double ph = s_phi[k];
      for (int i = 0; i < nr; ++i) tau_out[((size_t)b * nr + i) * nw + kbase + k] = s_ncol[i] * ph;
    }
  const int k0 = 2 * u, k1 = 2 * u + 1;       // this thread's bins within the tile
  const bool v0 = k0 < wt, v1 = k1 < wt;
  const double phi0 = v0 ? s_phi[k0] : 0.0, phi1 = v1 ? s_phi[k1] : 0.0;
  double phi_max = 0.0;   // of the tile (every thread scans the shared copy once: Wt <= 512 reads)
  for (int k = 0; k < wt; ++k) { const double x = fabs(s_phi[k]); phi_max = (x > phi_max || x != x) ? x : phi_max; }
  const unsigned e2 = (unsigned)__cvta_generic_to_shared(s_e2);
  const unsigned wide_top = (unsigned)__cvta_generic_to_shared(s_wide + kExpWideN);
  double ec[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) ec[q] = kExpWideC[q];
  double* my_px = s_px + (size_t)g * 2 * chunk;
  const unsigned px_addr = (unsigned)__cvta_generic_to_shared(my_px);
  double tot0 = 0.0, tot1 = 0.0;   // group 0: running sum over the phases
  const int gi_end = (ext == 1) ? geo.n_geo : geo0 + 1;
  for (int gi = geo0; gi < gi_end; ++gi) {
    const int* __restrict__ px_lo = geo.lo[gi];
    const double* __restrict__ px_whi = geo.whi[gi];
    const double* __restrict__ px_wlo = geo.wlo[gi];
    const double* __restrict__ px_i0 = geo.i0[gi];
    const int n_act = (int)geo.meta[gi][0];
    const int per = (n_act + S - 1) / S;
    const int nchunk = (per + chunk - 1) / chunk;
    double ph0 = 0.0, ph1 = 0.0;   // group 0: sum of this phase
    const int n_round = (ext == 1) ? (S + G - 1) / G : 1;
    for (int round = 0; round < n_round; ++round) {
      const int seg = (ext == 1) ? round * G + g : (int)blockIdx.z;
      const int p0 = min(n_act, seg * per), p1 = (seg < S) ? min(n_act, p0 + per) : p0;
      double acc0 = 0.0, acc1 = 0.0;
      for (int c = 0; c < nchunk; ++c) {
        const int start = p0 + c * chunk;
        const int cnt = max(0, min(chunk, p1 - start));
        __syncthreads();
        if (u == 0) s_flag[g] = 0;
        __syncthreads();
        bool big = false;
        for (int j = u; j < cnt; j += TPG) {
          const int q = start + j;
          const int lo = px_lo[q];
          const double np = px_whi[q] * s_ncol[lo + 1] + px_wlo[q] * s_ncol[lo];
          my_px[2 * j] = np;
          my_px[2 * j + 1] = px_i0[q];
          big = big || !(np * phi_max < kExpWideCap);   // (NaN counts as big)
        }
        if (big) s_flag[g] = 1;
        __syncthreads();
        // two pixels x two bins per trip: four independent exp chains
        unsigned a = px_addr;
        const unsigned a_end = px_addr + ((unsigned)(cnt & ~1) << 4);
        if (s_flag[g] == 0) {
          for (; a != a_end; a += 32) {
            double np_a, i0_a, np_b, i0_b;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(np_a), "=d"(i0_a) : "r"(a));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(np_b), "=d"(i0_b) : "r"(a));
            const double e0a = pw_exp_neg_wide(-(np_a * phi0), wide_top, ec);
            const double e1a = pw_exp_neg_wide(-(np_a * phi1), wide_top, ec);
            const double e0b = pw_exp_neg_wide(-(np_b * phi0), wide_top, ec);
            const double e1b = pw_exp_neg_wide(-(np_b * phi1), wide_top, ec);
            acc0 = pw_fma(i0_a, e0a, acc0); acc1 = pw_fma(i0_a, e1a, acc1);
            acc0 = pw_fma(i0_b, e0b, acc0); acc1 = pw_fma(i0_b, e1b, acc1);
          }
          if (cnt & 1) {
            const double np = my_px[2 * (cnt - 1)], i0 = my_px[2 * (cnt - 1) + 1];
            acc0 = pw_fma(i0, pw_exp_neg_wide(-(np * phi0), wide_top, ec), acc0);
            acc1 = pw_fma(i0, pw_exp_neg_wide(-(np * phi1), wide_top, ec), acc1);
          }
        } else {
#pragma unroll 2
          for (int j = 0; j < cnt; ++j) {
            const double np = my_px[2 * j], i0 = my_px[2 * j + 1];
            acc0 = pw_fma(i0, pw_exp_neg_sum_sh(-(np * phi0), e2), acc0);
            acc1 = pw_fma(i0, pw_exp_neg_sum_sh(-(np * phi1), e2), acc1);
          }
        }
      }
      if (ext > 1) {   // one segment per CTA: hand its sum to k_rt_reduce
        if (seg == 0) { acc0 += geo.meta[gi][1]; acc1 += geo.meta[gi][1]; }   // lit pixels with tau = 0
        store_pair(part + ((size_t)b * S + seg) * nw + kbase + k0, acc0, acc1, v0, v1);
        return;
      }
      // the G segment sums of this round, added in segment order by group 0
      if (g > 0 && seg < S) {
        if (v0) s_x[(size_t)g * Wt + k0] = acc0;
        if (v1) s_x[(size_t)g * Wt + k1] = acc1;
      }
      __syncthreads();
      if (g == 0) {
        if (round == 0) { acc0 += geo.meta[gi][1]; acc1 += geo.meta[gi][1]; }   // lit pixels with tau = 0
        ph0 = (round == 0) ? acc0 : ph0 + acc0;
        ph1 = (round == 0) ? acc1 : ph1 + acc1;
        for (int q = 1; q < G && round * G + q < S; ++q) {
          if (v0) ph0 += s_x[(size_t)q * Wt + k0];
          if (v1) ph1 += s_x[(size_t)q * Wt + k1];
        }
      }
    }
    if (g == 0) {
      const double depth = geo.depth[gi];
      if (depth != 0.0) { ph0 += depth; ph1 += depth; }
      tot0 = (gi == 0) ? ph0 : tot0 + ph0;
      tot1 = (gi == 0) ? ph1 : tot1 + ph1;
    }
  }
  if (g == 0) {
    if (geo.n_geo > 1) { tot0 /= (double)geo.n_geo; tot1 /= (double)geo.n_geo; }
    store_pair(spec + (size_t)b * nw + kbase + k0, tot0, tot1, v0, v1);
  }
}

// Fixed-order sum of the pixel segments of one phase; `depth` (geometric transit depth of the
// phase) is added, the result accumulated over phases and divided by their number at the end.
__global__ void k_rt_reduce(int B, int nw, int nsplit, const double* __restrict__ part, double* spec, double depth,
                            int accumulate, int divide_by) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * nw) return;
  const int b = (int)(idx / nw), k = (int)(idx % nw);
  double s = 0.0;
  for (int q = 0; q < nsplit; ++q) s += part[((size_t)b * nsplit + q) * nw + k];
  if (depth != 0.0) s += depth;
  if (accumulate) s = spec[idx] + s;
  if (divide_by > 0) s /= (double)divide_by;
  spec[idx] = s;
}

__global__ void k_chi2(int B, int nw, const double* __restrict__ spec, const double* __restrict__ data,
                       const double* __restrict__ sigma, double norm, const int* __restrict__ status,
                       int accept_warn, double* chi2) {
  __shared__ double red[40];
  const int b = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < nw; k += blockDim.x) {
    const double d = (data[k] - spec[(size_t)b * nw + k] / norm) / sigma[k];
    s += d * d;
  }
  const double t = block_sum(s, red);
  if (threadIdx.x == 0) {
    const bool bad = status && status[b] != PW_OK && !(accept_warn && status[b] == PW_WARN_HE_RELAX);
    chi2[b] = bad ? (1.0 / 0.0) : t;
  }
}


// Instrumental broadening of the model spectrum: astropy.convolution.convolve(spec, kernel,
// boundary='extend') as the retrieval of docs/source/advanced_tutorial.ipynb (cell 13) calls it.
// astropy pads the array with its edge values by half the kernel width and evaluates the true
// convolution out[i] = sum_ii pad[i + ii] * g[nk - 1 - ii], ii ascending; `kern` arrives already
// divided by its sum (normalize_kernel=True, done with numpy on the host like astropy does).
// One CTA per model, the spectrum staged in shared memory.
__global__ void k_convolve_extend(int B, int nw, const double* __restrict__ spec, const double* __restrict__ kern,
                                  int nk, double* __restrict__ out) {
  extern __shared__ double sm[];  // nw spectrum + nk kernel
  const int b = blockIdx.x;
  double* sx = sm;
  double* sk = sm + nw;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sx[i] = spec[(size_t)b * nw + i];
  for (int i = threadIdx.x; i < nk; i += blockDim.x) sk[i] = kern[i];
  __syncthreads();
  const int w = nk / 2;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    double top = 0.0;
    for (int ii = 0; ii < nk; ++ii) {
      int src = i + ii - w;
      src = src < 0 ? 0 : (src > nw - 1 ? nw - 1 : src);
      top += sx[src] * sk[nk - 1 - ii];
    }
    out[(size_t)b * nw + i] = top;
  }
}

// log-likelihood of advanced_tutorial.ipynb cell 15: -0.5 * sum((y - m)^2 / yerr^2 + log(yerr^2)),
// -inf where the model failed (the notebook's `except RuntimeError`)
__global__ void k_loglike(int B, int nw, const double* __restrict__ model, const double* __restrict__ y,
                          const double* __restrict__ yerr, const int* __restrict__ status, int accept_warn,
                          double* out) {
  __shared__ double red[40];
  const int b = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < nw; k += blockDim.x) {
    const double d = y[k] - model[(size_t)b * nw + k];
    const double s2 = yerr[k] * yerr[k];
    s += d * d / s2 + log(s2);
  }
  const double t = block_sum(s, red);
  if (threadIdx.x == 0) {
    const bool bad = status && status[b] != PW_OK && !(accept_warn && status[b] == PW_WARN_HE_RELAX);
    out[b] = bad ? -(1.0 / 0.0) : -0.5 * t;
  }
}

// ------------------------------------------------------------------------------------
// transit.draw_transit (transit.py:20-116) with flatstar's rasteriser restated:
// star and planet disks are the pixel spans of Pillow's ImageDraw.ellipse (integer walk
// done on the host, O(grid) work), limb darkening per super-sampled pixel, flux
// conserving box re-binning (the reference's default resample method).
// ------------------------------------------------------------------------------------
struct DrawParams {
  int n;            // super-sampled grid size
  int g;            // output grid size
  int ss;           // n / g
  double cen, r_star;
  int law;
  double c[4];
};

__device__ __forceinline__ double limb_darkening(int law, const double* c, double mu) {
  switch (law) {
    case 1: return 1 - c[0] * (1 - mu);
    case 2: return 1 - c[0] * (1 - mu) - c[1] * ((1 - mu) * (1 - mu));
    case 3: return 1 - c[0] * (1 - mu) - c[1] * (1 - sqrt(mu));
    case 4: return 1 - c[0] * (1 - mu) - c[1] * ((mu > 0) ? mu * log(mu) : 0.0);
    case 5: return 1 - c[0] * (1 - mu) - c[1] / (1 - exp(mu));
    case 6: return 1 - c[0] * (1 - mu) - c[1] * (1 - mu * sqrt(mu)) - c[2] * (1 - mu * mu);
    case 7: return 1 - c[0] * (1 - sqrt(mu)) - c[1] * (1 - mu) - c[2] * (1 - mu * sqrt(mu)) - c[3] * (1 - mu * mu);
    default: return 1.0;
  }
}

__device__ __forceinline__ void draw_pixel(const DrawParams& p, const int* __restrict__ spans, int row, int col,
                                           double* star, double* lit) {
  // spans: [4][n] = star_lo, star_hi, planet_lo, planet_hi per row (hi < lo: empty)
  const int n = p.n;
  double in = 0.0, tr = 0.0;
  if (col >= spans[row] && col <= spans[n + row]) {
    const double dx = col - p.cen, dy = row - p.cen;
    double m2 = 1 - (dx * dx + dy * dy) / (p.r_star * p.r_star);
    m2 = fmin(fmax(m2, 0.0), 1.0);
    in = limb_darkening(p.law, p.c, sqrt(m2));
    const bool behind = col >= spans[2 * n + row] && col <= spans[3 * n + row];
    tr = behind ? in * 0.0 : in;   // inten * (1 - planet): keeps a NaN of the 'exp' law like numpy
  }
  *star = in;
  *lit = tr;
}

// per-block partial sums of the star and of the transit image over the super-sampled grid
__global__ void k_draw_sums(DrawParams p, const int* __restrict__ spans, double* partial) {
  __shared__ double red[40];
  double s_in = 0.0, s_tr = 0.0;
  const long long tot = (long long)p.n * p.n;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
    double a, b;
    draw_pixel(p, spans, (int)(q / p.n), (int)(q % p.n), &a, &b);
    s_in += a; s_tr += b;
  }
  const double t_in = block_sum(s_in, red);
  const double t_tr = block_sum(s_tr, red);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = t_in; partial[2 * blockIdx.x + 1] = t_tr; }
}

__global__ void k_draw_total(int nblocks, const double* __restrict__ partial, double* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblocks; ++i) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    total[0] = a; total[1] = b;
  }
}

// one thread per output pixel: box sum of its ss x ss block, normalised; radius map
__global__ void k_draw_bin(DrawParams p, const int* __restrict__ spans, const double* __restrict__ total,
                           double x_p, double y_p, double px_size, double* i0, double* r_map) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.g * p.g) return;
  const int row = q / p.g, col = q % p.g;
  const double norm = total[0];
  double acc = 0.0;
  for (int a = 0; a < p.ss; ++a)
    for (int b = 0; b < p.ss; ++b) {
      double in, tr;
      draw_pixel(p, spans, row * p.ss + a, col * p.ss + b, &in, &tr);
      acc += tr / norm;
    }
  i0[q] = acc;
  const double dx = col - x_p, dy = row - y_p;
  r_map[q] = sqrt(dx * dx + dy * dy) * px_size;
}

// FP64 peak probe: 8 independent DFMA chains per thread
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// The two arithmetic primitives the device build of the LSODA step swaps in for a / b and
// pow(a, 1 / k), exposed so that the GPU tests can hold them to IEEE division / libm over the
// operand ranges of the step (tests/test_gpu_parity.py::test_step_arithmetic_on_gpu).
__global__ void k_selftest_arith(int n, const double* __restrict__ a, const double* __restrict__ b,
                                 const int* __restrict__ k, double* __restrict__ quot, double* __restrict__ root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  quot[i] = PW_DIV(a[i], b[i]);
  root[i] = pw_root(a[i], k[i]);
}

// mu_0 = (1 + 4 he/h) / (1 + he/h + mean_f_ion), mean_f_ion = 0 (quickstart.ipynb cell 2)
__global__ void k_mu0(int B, const double* __restrict__ hfrac, double* mu0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double he_h = (1 - hfrac[b]) / hfrac[b];
  mu0[b] = (1 + 4 * he_h) / (1 + he_h + 0.0);
}


// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" {

// PW_SRC_HASH: sha256 (first 16 hex digits) of csrc/* and include/pwinds_b200.h, passed by the build
// (__graft_entry__.build / build.sh) so that a shipped binary can be matched to its sources.
#ifndef PW_SRC_HASH
#define PW_SRC_HASH "unknown"
#endif
const char* pwb_version(void) { return "p_winds_b200 0.2 (sm_100a) src:" PW_SRC_HASH; }

int pwb_sizeof_opts(int which) {
  switch (which) {
    case 0: return (int)sizeof(pwb_h_opts);
    case 1: return (int)sizeof(pwb_he_opts);
    case 2: return (int)sizeof(pwb_o_opts);
    case 3: return (int)sizeof(pwb_c_opts);
    case 4: return (int)sizeof(pwb_rt_opts);
    default: return -1;
  }
}

int pwb_create(int device, pwb_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -2;  // no CUDA device: no fallback
  if (device < 0 || device >= ndev) return -3;
  if (cudaSetDevice(device) != cudaSuccess) return -4;
  pwb_ctx* c = new pwb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (c->lsoda_tab.ensure(sizeof(LsodaTables)) != 0) { delete c; return -5; }
  k_lsoda_tables<<<1, 32>>>(c->lsoda_tab.as<LsodaTables>());
  c->launches++;
  {  // table of pw_exp_neg_wide: correctly rounded 2^(j/256 - 16) from the host's long double exp2
    std::vector<double> tab(kExpWideN + 1);
    for (int i = 0; i <= kExpWideN; ++i) tab[i] = (double)exp2l((long double)(i - kExpWideN) / 256.0L);
    if (c->exp_wide.ensure(sizeof(double) * tab.size()) != 0 ||
        cudaMemcpy(c->exp_wide.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete c;
      return -5;
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { delete c; return -6; }
  cudaFuncSetAttribute(k_h_fields, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_rt_los, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_rt_formal_tau, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  *out = c;
  return 0;
}

void pwb_destroy(pwb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  DevBuf* bufs[] = {&c->sed_gw, &c->sed_sh, &c->sed_she, &c->sed_scal, &c->lsoda_tab, &c->exp_wide, &c->geo_r, &c->h_state, &c->h_fields, &c->he_order, &c->he_key, &c->he_work, &c->o_work, &c->rt_ncol, &c->rt_sums,
                    &c->rt_partial, &c->rt_tau, &c->fw_f, &c->fw_scal, &c->fw_he, &c->fw_n, &c->fw_v, &c->fw_T, &c->misc};
  for (DevBuf* b : bufs) b->release();
  for (auto& G : c->geo) {
    G.px_lo.release(); G.px_whi.release(); G.px_wlo.release(); G.px_i0.release(); G.px_meta.release();
  }
  delete c;
}

const char* pwb_last_error(pwb_ctx* c) { return c ? c->err.c_str() : "null context"; }
long long pwb_launch_count(pwb_ctx* c) { return c ? c->launches : 0; }
int pwb_n_active_pixels(pwb_ctx* c) { return (c && c->geo[0].ready) ? c->geo[0].n_active : 0; }

int pwb_set_sed(pwb_ctx* c, const double* d_wl, const double* d_flux, int n, void* stream) {
  if (!c) return -1;
  if (n < 3 || !d_wl || !d_flux) return fail(c, "pwb_set_sed: need >= 3 spectral points");
  cudaStream_t s = (cudaStream_t)stream;
  PWB_CUDA(c, cudaSetDevice(c->device));
  if (c->sed_gw.ensure(sizeof(double) * n) || c->sed_sh.ensure(sizeof(double) * n) ||
      c->sed_she.ensure(sizeof(double) * n) || c->sed_scal.ensure(sizeof(double) * PW_SC_COUNT))
    return fail(c, "pwb_set_sed: out of device memory");
  k_sed_setup<<<1, 256, 0, s>>>(d_wl, d_flux, n, c->sed_gw.as<double>(), c->sed_sh.as<double>(),
                               c->sed_she.as<double>(), c->sed_scal.as<double>());
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  PWB_CUDA(c, cudaMemcpyAsync(c->sed_scal_host, c->sed_scal.p, sizeof(double) * PW_SC_COUNT,
                              cudaMemcpyDeviceToHost, s));
  PWB_CUDA(c, cudaStreamSynchronize(s));
  c->sed_n = n;
  c->sed_nh = (int)c->sed_scal_host[PW_SC_N_H];
  c->sed_ready = true;
  return 0;
}

int pwb_get_sed_scalars(pwb_ctx* c, double* out32) {
  if (!c || !c->sed_ready) return fail(c, "pwb_get_sed_scalars: no SED set");
  memcpy(out32, c->sed_scal_host, sizeof(double) * PW_SC_COUNT);
  return 0;
}

int pwb_parker_structure(pwb_ctx* c, const double* d_r, int n, double* d_v, double* d_rho, void* stream) {
  if (!c) return -1;
  if (n <= 0) return 0;
  k_parker<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_r, n, d_v, d_rho);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_parker_structure_tidal(pwb_ctx* c, const double* d_r, int n, double cs, double rs, double mp,
                               double ms, double a, double* d_v, double* d_rho, void* stream) {
  if (!c) return -1;
  if (n <= 0) return 0;
  k_parker_tidal<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_r, n, cs, rs, mp, ms, a, d_v, d_rho);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

static int launch_hydrogen(pwb_ctx* c, cudaStream_t s, int B, const double* d_T, const double* d_mdot,
                           const double* d_hfrac, const double* d_mu0, const double* d_r, int Nr,
                           const pwb_h_opts* o, double* d_f_r, double* d_v, double* d_rho, double* d_scal,
                           int* d_status) {
  const int n_h = c->sed_ready ? c->sed_nh : 0;
  const size_t work = sizeof(double) * (10 * (size_t)Nr + 40), tab = sizeof(double) * 3 * (size_t)n_h;
  const int in_smem = (o->exact_phi && work + tab <= 160 * 1024) ? 1 : 0;
  const size_t smem = work + (in_smem ? tab : 0);
  if (smem > 200 * 1024) return fail(c, "hydrogen stage: radius profile too long for shared memory");
  if (c->h_state.ensure(sizeof(HState) * (size_t)B) ||
      c->h_fields.ensure(sizeof(double) * kHFieldsPerRadius * (size_t)Nr * B))
    return fail(c, "out of device memory");
  HState* st = c->h_state.as<HState>();
  double* fields = c->h_fields.as<double>();
  const int n_pass = o->relax_solution ? o->max_n_relax + 1 : 1;
  // models per warp of the solve kernel: spread a small batch over all warp schedulers
  // (the solve is latency bound), pack a large one
  int lanes = (int)std::min<long long>(32, std::max<long long>(1, ((long long)B + 4LL * c->sm_count - 1) / (4LL * c->sm_count)));
  if (const char* e = getenv("PWB_H_LANES")) { const int v = atoi(e); if (v >= 1 && v <= 32) lanes = v; }
  for (int pass = 0; pass < n_pass; ++pass) {
    k_h_fields<<<B, 128, smem, s>>>(B, pass, d_T, d_mdot, d_hfrac, d_mu0, d_r, Nr, *o, c->sed_gw.as<double>(),
                                    c->sed_sh.as<double>(), c->sed_she.as<double>(), n_h, in_smem, st, fields);
    k_h_solve<<<(B + lanes - 1) / lanes, 32, 0, s>>>(B, lanes, pass, d_T, d_mdot, d_hfrac, d_mu0, d_r, Nr, *o, st, fields);
    c->launches += 2;
  }
  k_h_finish<<<B, 128, 0, s>>>(B, d_T, d_mdot, d_hfrac, d_mu0, d_r, Nr, *o, st, fields, d_f_r, d_v, d_rho, d_scal,
                               d_status);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_h_rates_batch(pwb_ctx* c, int B, const double* d_T, const double* d_mdot, const double* d_hfrac,
                      const double* d_mu0, const double* d_r, int Nr, const pwb_h_opts* o, double* d_rates,
                      void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (!o->exact_phi || !c->sed_ready) return fail(c, "pwb_h_rates_batch: needs exact_phi and a SED (hydrogen.py:540-543)");
  if (c->h_state.cap < sizeof(HState) * (size_t)B || c->h_fields.cap < sizeof(double) * kHFieldsPerRadius * (size_t)Nr * B)
    return fail(c, "pwb_h_rates_batch: call pwb_h_ion_fraction_batch with the same batch first");
  const size_t smem = sizeof(double) * (10 * (size_t)Nr + 40);
  if (smem > 200 * 1024) return fail(c, "pwb_h_rates_batch: radius profile too long for shared memory");
  if (smem > 48 * 1024) PWB_CUDA(c, cudaFuncSetAttribute(k_h_rates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_h_rates<<<B, 128, smem, (cudaStream_t)stream>>>(B, d_T, d_mdot, d_hfrac, d_mu0, d_r, Nr, *o, c->sed_gw.as<double>(),
                                                     c->sed_sh.as<double>(), c->sed_she.as<double>(), c->sed_nh,
                                                     c->h_state.as<HState>(), c->h_fields.as<double>(), d_rates);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_h_ion_fraction_batch(pwb_ctx* c, int B, const double* d_T, const double* d_mdot,
                             const double* d_hfrac, const double* d_mu0, const double* d_r, int Nr,
                             const pwb_h_opts* o, double* d_f_r, double* d_v, double* d_rho,
                             double* d_scal, int* d_status, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nr < 2) return fail(c, "pwb_h_ion_fraction_batch: radius profile needs >= 2 points");
  if (o->exact_phi && !c->sed_ready) return fail(c, "pwb_h_ion_fraction_batch: exact_phi needs pwb_set_sed");
  return launch_hydrogen(c, (cudaStream_t)stream, B, d_T, d_mdot, d_hfrac, d_mu0, d_r, Nr, o, d_f_r, d_v, d_rho,
                         d_scal, d_status);
}

// Models per warp along the cost-sorted list (see HeSched).  `frac[s]` of the models, from the
// costliest down, run `lanes[s]` to a warp.  The default pattern -- 5 % alone, 25 % in pairs,
// 40 % in fours, the cheapest 30 % eight to a warp -- is the measured optimum for the 64 x 64
// retrieval grid on one B200 (scripts/sweep_sched.py: 62.6 ms, flat within 3 % around it; one
// lane count for all: 75-90 ms); all lane counts are scaled up together when the
// batch would otherwise need more warps than the GPU takes in one wave (7 per SM, see he_schedule).
// PWB_HE_SCHED="f:L,f:L,..." overrides the pattern, PWB_HE_LANES=L forces one lane count and the
// caller's order.
static HeSched he_schedule(pwb_ctx* c, int B, bool* sorted) {
  double frac[4] = {0.05, 0.25, 0.40, 0.30};
  int lanes[4] = {1, 2, 4, 8};
  int nseg = 4;
  *sorted = true;
  if (const char* e = getenv("PWB_HE_SCHED")) {
    int n = 0;
    const char* q = e;
    while (*q && n < 4) {
      double f = 0; int l = 0; int used = 0;
      if (sscanf(q, "%lf:%d%n", &f, &l, &used) != 2 || l < 1 || l > 32) break;
      frac[n] = f; lanes[n] = l; ++n;
      q += used;
      if (*q == ',') ++q;
    }
    if (n > 0) nseg = n;
  }
  if (const char* e = getenv("PWB_HE_LANES")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 32) { nseg = 1; frac[0] = 1.0; lanes[0] = v; *sorted = false; }
  } else if (!getenv("PWB_HE_SCHED")) {
    // Warps the schedule may ask for per SM.  The register file holds 8 of these 255-register warps, but a
    // grid of exactly 8 x SMs does not land evenly: the stragglers start when the first warps finish and
    // the kernel takes a quarter longer (24 576 models: 67.5 ms at 8, 52.6 ms at 7, 53.9 at 6;
    // scripts/he_big.py).  7 leaves the slack.
    double per_sm = 7.0;
    if (const char* e2 = getenv("PWB_HE_CAP")) { const double v = atof(e2); if (v >= 1. && v <= 32.) per_sm = v; }
    const double capacity = per_sm * c->sm_count;
    double warps = 0;
    for (int s = 0; s < nseg; ++s) warps += frac[s] * B / lanes[s];
    if (warps > capacity) {
      const double scale = warps / capacity;
      for (int s = 0; s < nseg; ++s) lanes[s] = std::min(32, (int)std::ceil(lanes[s] * scale));
    }
  }
  HeSched h;
  h.nseg = nseg;
  int slot = 0, warp = 0;
  for (int s = 0; s < nseg; ++s) {
    int n = (s == nseg - 1) ? B - slot : std::min(B - slot, (int)std::llround(frac[s] * B));
    if (n < 0) n = 0;
    h.slot0[s] = slot; h.warp0[s] = warp; h.lanes[s] = lanes[s];
    slot += n;
    warp += (n + lanes[s] - 1) / lanes[s];
  }
  for (int s = nseg; s <= 4; ++s) { h.slot0[s] = slot; h.warp0[s] = warp; }
  for (int s = nseg; s < 4; ++s) h.lanes[s] = 1;
  return h;
}

static int launch_helium(pwb_ctx* c, cudaStream_t s, int B, const double* d_T, const double* d_hfrac,
                         const double* d_vs, const double* d_rs, const double* d_rhos, int sstride,
                         const double* d_r, int Nr, const double* d_v, const double* d_rho, const double* d_f_h,
                         const pwb_he_opts* o, double* work, int* order, double* d_f1, double* d_f3, double* d_scal,
                         double* d_rates, int* d_status) {
  bool sorted = true;
  const HeSched sched = he_schedule(c, B, &sorted);
  if (sorted) {
    if (c->he_key.ensure(sizeof(double) * (size_t)B)) return fail(c, "out of device memory");
    k_he_key<<<(B + 127) / 128, 128, 0, s>>>(B, d_rhos, sstride, d_r, Nr, d_rho, d_status, c->he_key.as<double>());
    // A batch that takes several waves of warps (more models than 32 lanes x 7 warps x SMs) is sorted into
    // 128 cost classes only: with the fine sort every warp of a wave holds models of the same cost, the
    // warps of an SM move through the same phases in step and finish together; mixing within a class
    // staggers them (196 608 models: 263 -> 233 ms; scripts/he_coarse.sh).  One wave: fine sort (55 vs 57 ms).
    int coarse = ((long long)B > 32LL * 7 * c->sm_count) ? 32 : 1;
    if (const char* e = getenv("PWB_HE_COARSE")) { const int v = atoi(e); if (v >= 1 && v <= kHeBuckets) coarse = v; }
    k_he_order<<<1, 1024, 0, s>>>(B, c->he_key.as<double>(), d_status, order, coarse);
    c->launches += 2;
  }
  const int* ord = sorted ? order : nullptr;
  const int blocks = sched.warp0[sched.nseg];
  const size_t ws = (size_t)kHeWorkPerRadius * Nr;
#if defined(PW_HE_PS_KERNEL)
  if (o->method != PW_HE_RK45 && getenv("PWB_HE_PS")) {
    k_helium_ps<<<blocks, 32, 0, s>>>(B, sched, ord, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho,
                                      d_f_h, *o, c->lsoda_tab.as<LsodaTables>(), work, ws, d_f1, d_f3, d_scal,
                                      d_rates, d_status);
    c->launches++;
    PWB_CUDA(c, cudaGetLastError());
    return 0;
  }
#endif
  if (o->method == PW_HE_RK45)
    k_helium<true><<<blocks, 32, 0, s>>>(B, sched, ord, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho,
                                         d_f_h, *o, c->lsoda_tab.as<LsodaTables>(), work, ws, d_f1, d_f3, d_scal,
                                         d_rates, d_status);
  else
    k_helium<false><<<blocks, 32, 0, s>>>(B, sched, ord, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho,
                                          d_f_h, *o, c->lsoda_tab.as<LsodaTables>(), work, ws, d_f1, d_f3, d_scal,
                                          d_rates, d_status);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_he_population_batch(pwb_ctx* c, int B, const double* d_T, const double* d_hfrac,
                            const double* d_vs, const double* d_rs, const double* d_rhos,
                            int sstride, const double* d_r, int Nr, const double* d_v,
                            const double* d_rho, const double* d_f_h, const pwb_he_opts* o,
                            double* d_f1, double* d_f3, double* d_scal, double* d_rates,
                            int* d_status, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nr < 2) return fail(c, "pwb_he_population_batch: radius profile needs >= 2 points");
  if (o->method < 0 || o->method > 2) return fail(c, "pwb_he_population_batch: unsupported method");
  if (c->he_work.ensure(sizeof(double) * (size_t)kHeWorkPerRadius * Nr * B) ||
      c->he_order.ensure(sizeof(int) * (size_t)B))
    return fail(c, "out of device memory");
  return launch_helium(c, (cudaStream_t)stream, B, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho,
                       d_f_h, o, c->he_work.as<double>(), c->he_order.as<int>(), d_f1, d_f3, d_scal, d_rates,
                       d_status);
}

int pwb_o_ion_fraction_batch(pwb_ctx* c, int B, const double* d_T, const double* d_hfrac, const double* d_vs,
                             const double* d_rs, const double* d_rhos, int sstride, const double* d_r, int Nr,
                             const double* d_v, const double* d_rho, const double* d_f_h,
                             const double* d_f_he_ion, const pwb_o_opts* o, double* d_f_o, double* d_scal,
                             double* d_rates, int* d_status, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nr < 2) return fail(c, "pwb_o_ion_fraction_batch: radius profile needs >= 2 points");
  if (o->method < 0 || o->method > 2) return fail(c, "pwb_o_ion_fraction_batch: unsupported method");
  if (c->o_work.ensure(sizeof(double) * kOWorkPerRadius * (size_t)Nr * B)) return fail(c, "out of device memory");
  int lanes = (int)std::min<long long>(32, std::max<long long>(1, ((long long)B + 4LL * c->sm_count - 1) / (4LL * c->sm_count)));
  k_oxygen<<<(B + lanes - 1) / lanes, 32, 0, (cudaStream_t)stream>>>(
      B, lanes, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho, d_f_h, d_f_he_ion, *o,
      c->lsoda_tab.as<LsodaTables>(), c->o_work.as<double>(), d_f_o, d_scal, d_rates, d_status);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_he_ion_fraction_batch(pwb_ctx* c, int B, const double* d_T, const double* d_hfrac, const double* d_vs,
                              const double* d_rs, const double* d_rhos, int sstride, const double* d_r, int Nr,
                              const double* d_v, const double* d_rho, const double* d_f_h,
                              const pwb_heion_opts* o, double* d_f_he_ion, double* d_scal, int* d_status,
                              void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nr < 2) return fail(c, "pwb_he_ion_fraction_batch: radius profile needs >= 2 points");
  if (o->method < 0 || o->method > 2) return fail(c, "pwb_he_ion_fraction_batch: unsupported method");
  if (c->o_work.ensure(sizeof(double) * kOWorkPerRadius * (size_t)Nr * B)) return fail(c, "out of device memory");
  int lanes = (int)std::min<long long>(32, std::max<long long>(1, ((long long)B + 4LL * c->sm_count - 1) / (4LL * c->sm_count)));
  k_he_ion<<<(B + lanes - 1) / lanes, 32, 0, (cudaStream_t)stream>>>(
      B, lanes, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho, d_f_h, *o,
      c->lsoda_tab.as<LsodaTables>(), c->o_work.as<double>(), d_f_he_ion, d_scal, d_status);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_c_ion_fraction_batch(pwb_ctx* c, int B, const double* d_T, const double* d_hfrac, const double* d_vs,
                             const double* d_rs, const double* d_rhos, int sstride, const double* d_r, int Nr,
                             const double* d_v, const double* d_rho, const double* d_f_h,
                             const double* d_f_he_ion, const pwb_c_opts* o, double* d_f_cii, double* d_f_ciii,
                             double* d_scal, double* d_rates, int* d_status, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nr < 2) return fail(c, "pwb_c_ion_fraction_batch: radius profile needs >= 2 points");
  if (o->method < 0 || o->method > 2) return fail(c, "pwb_c_ion_fraction_batch: unsupported method");
  if (c->o_work.ensure(sizeof(double) * kCWorkPerRadius * (size_t)Nr * B)) return fail(c, "out of device memory");
  int lanes = (int)std::min<long long>(32, std::max<long long>(1, ((long long)B + 4LL * c->sm_count - 1) / (4LL * c->sm_count)));
  k_carbon<<<(B + lanes - 1) / lanes, 32, 0, (cudaStream_t)stream>>>(
      B, lanes, d_T, d_hfrac, d_vs, d_rs, d_rhos, sstride, d_r, Nr, d_v, d_rho, d_f_h, d_f_he_ion, *o,
      c->lsoda_tab.as<LsodaTables>(), c->o_work.as<double>(), d_f_cii, d_f_ciii, d_scal, d_rates, d_status);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_set_geometry_phase(pwb_ctx* c, int slot, const double* d_i0, const double* d_r_map, int npix,
                           const double* d_r_si, int Nr, double transit_depth, void* stream) {
  if (!c) return -1;
  if (slot < 0 || slot >= pwb_ctx::kMaxGeo) return fail(c, "pwb_set_geometry_phase: slot out of range (0..15)");
  if (slot > 0 && (!c->geo[0].ready || c->geo_nr != Nr))
    return fail(c, "pwb_set_geometry_phase: set slot 0 first, with the same radius grid");
  pwb_ctx::Geo& G = c->geo[slot];
  if (npix <= 0 || Nr < 2 || Nr > 1024) return fail(c, "pwb_set_geometry: bad sizes (need 2 <= Nr <= 1024)");
  cudaStream_t s = (cudaStream_t)stream;
  if (G.px_lo.ensure(sizeof(int) * npix) || G.px_whi.ensure(sizeof(double) * npix) ||
      G.px_wlo.ensure(sizeof(double) * npix) || G.px_i0.ensure(sizeof(double) * npix) ||
      G.px_meta.ensure(sizeof(double) * 2) || c->geo_r.ensure(sizeof(double) * Nr))
    return fail(c, "pwb_set_geometry: out of device memory");
  PWB_CUDA(c, cudaMemcpyAsync(c->geo_r.p, d_r_si, sizeof(double) * Nr, cudaMemcpyDeviceToDevice, s));
  k_pixel_table<<<1, 1024, 0, s>>>(d_i0, d_r_map, npix, c->geo_r.as<double>(), Nr, G.px_lo.as<int>(),
                                   G.px_whi.as<double>(), G.px_wlo.as<double>(), G.px_i0.as<double>(),
                                   G.px_meta.as<double>());
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  double meta[2];
  PWB_CUDA(c, cudaMemcpyAsync(meta, G.px_meta.p, sizeof meta, cudaMemcpyDeviceToHost, s));
  PWB_CUDA(c, cudaStreamSynchronize(s));
  int n_act = (int)meta[0];
  // Pixels at the same distance from the planet's centre see the same optical depth at every
  // wavelength: sum_p I0_p e^{-tau(r_p)} only needs one term per distinct radius, with the
  // intensities added up.  A transit map is mirror symmetric about the line through both disc
  // centres (and 8-fold symmetric for a central transit), so this halves the pixel loop of the
  // per-model hot path or better.  Per-geometry set-up, done once on the host: sort the table
  // by (bracket, weight) -- equal keys <=> equal radius -- and merge runs, adding I0 in pixel
  // order.  (Changes the summation order of the spectrum: ~1e-16 relative.)
  if (n_act > 1 && !getenv("PWB_NO_PIXEL_MERGE")) {
    std::vector<int> lo(n_act), idx(n_act);
    std::vector<double> whi(n_act), wlo(n_act), pi0(n_act);
    PWB_CUDA(c, cudaMemcpy(lo.data(), G.px_lo.p, sizeof(int) * n_act, cudaMemcpyDeviceToHost));
    PWB_CUDA(c, cudaMemcpy(whi.data(), G.px_whi.p, sizeof(double) * n_act, cudaMemcpyDeviceToHost));
    PWB_CUDA(c, cudaMemcpy(wlo.data(), G.px_wlo.p, sizeof(double) * n_act, cudaMemcpyDeviceToHost));
    PWB_CUDA(c, cudaMemcpy(pi0.data(), G.px_i0.p, sizeof(double) * n_act, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_act; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
      if (lo[a] != lo[b]) return lo[a] < lo[b];
      if (whi[a] != whi[b]) return whi[a] < whi[b];
      return wlo[a] < wlo[b];
    });
    std::vector<int> mlo;
    std::vector<double> mwhi, mwlo, mi0;
    for (int q = 0; q < n_act; ++q) {
      const int i = idx[q];
      if (!mlo.empty() && mlo.back() == lo[i] && mwhi.back() == whi[i] && mwlo.back() == wlo[i]) {
        mi0.back() += pi0[i];
      } else {
        mlo.push_back(lo[i]); mwhi.push_back(whi[i]); mwlo.push_back(wlo[i]); mi0.push_back(pi0[i]);
      }
    }
    n_act = (int)mlo.size();
    meta[0] = (double)n_act;
    PWB_CUDA(c, cudaMemcpy(G.px_lo.p, mlo.data(), sizeof(int) * n_act, cudaMemcpyHostToDevice));
    PWB_CUDA(c, cudaMemcpy(G.px_whi.p, mwhi.data(), sizeof(double) * n_act, cudaMemcpyHostToDevice));
    PWB_CUDA(c, cudaMemcpy(G.px_wlo.p, mwlo.data(), sizeof(double) * n_act, cudaMemcpyHostToDevice));
    PWB_CUDA(c, cudaMemcpy(G.px_i0.p, mi0.data(), sizeof(double) * n_act, cudaMemcpyHostToDevice));
    PWB_CUDA(c, cudaMemcpy(G.px_meta.p, meta, sizeof meta, cudaMemcpyHostToDevice));
  }
  G.npix = npix; G.n_active = n_act; G.s_out = meta[1]; G.depth = transit_depth; G.ready = true;
  c->geo_nr = Nr;
  return 0;
}

int pwb_set_geometry(pwb_ctx* c, const double* d_i0, const double* d_r_map, int npix,
                     const double* d_r_si, int Nr, void* stream) {
  if (!c) return -1;
  c->n_geo = 1;
  return pwb_set_geometry_phase(c, 0, d_i0, d_r_map, npix, d_r_si, Nr, 0.0, stream);
}

int pwb_set_phase_count(pwb_ctx* c, int n_phases) {
  if (!c) return -1;
  if (n_phases < 1 || n_phases > pwb_ctx::kMaxGeo) return fail(c, "pwb_set_phase_count: 1..16 phases");
  for (int g = 0; g < n_phases; ++g)
    if (!c->geo[g].ready) return fail(c, "pwb_set_phase_count: a phase slot has no geometry yet");
  c->n_geo = n_phases;
  return 0;
}


static int rt_nsplit(const pwb_ctx* c) {
  // The active pixels are always cut into the same segments, whatever the batch size,
  // so that a model's spectrum is bit-identical alone or inside a grid (the partial
  // sums are combined in a fixed order by k_rt_reduce).  <= 16 segments of >= 1024 pixels.
  int n = 0;
  for (int g = 0; g < c->n_geo; ++g) n = std::max(n, c->geo[g].n_active);
  return std::min(16, std::max(1, (n + 1023) / 1024));
}

static int rt_check(pwb_ctx* c, const pwb_rt_opts* o) {
  if (!c->geo[0].ready) return fail(c, "pwb_rt2d_batch: call pwb_set_geometry first");
  if (o->n_lines < 1 || o->n_lines > 8) return fail(c, "pwb_rt2d_batch: 1..8 spectral lines supported");
  if (o->wind_broadening_method != 0 && o->wind_broadening_method != 1)
    return fail(c, "The chosen ``wind_broadening_method`` is not implemented.");
  if (o->z_grid_size < 2) return fail(c, "pwb_rt2d_batch: z_grid_size must be >= 2");
  return 0;
}

static int launch_rt(pwb_ctx* c, cudaStream_t s, int B, const double* d_n, const double* d_v, const double* d_T,
                     const double* d_wl, int Nw, const pwb_rt_opts* o, const int* d_status, double* d_spectrum,
                     double* d_tau, double* ncol, double* sums, double* partial, double* tau_tab) {
  const int Nr = c->geo_nr;
  const int tiles = (Nw + 255) / 256;
  const int nsplit = rt_nsplit(c);
  // one phase, no depth to add, one pixel segment: the sum goes straight to the spectrum
  const bool direct = (nsplit == 1 && c->n_geo == 1 && c->geo[0].depth == 0.0);
  double* part = direct ? d_spectrum : partial;
  if (o->wind_broadening_method == 1) {
    const size_t smem = sizeof(double) * (4 * (size_t)Nr + 4 * (size_t)o->z_grid_size);
    if (smem > 100 * 1024) return fail(c, "pwb_rt2d_batch: z_grid_size too large for the 'formal' kernel");
    dim3 g1(Nr, B);
    // threads own wavelengths: no more warps than the bins need (Nw = 200: 7 warps, not 8)
    const int nt_formal = std::min(256, ((Nw + 31) / 32) * 32);
    k_rt_formal_tau<<<g1, nt_formal, smem, s>>>(B, Nw, Nr, *o, c->geo_r.as<double>(), d_n, d_v, d_T, d_wl, d_status, tau_tab);
    c->launches++;
    PWB_CUDA(c, cudaGetLastError());
  } else {
    const size_t smem_los = sizeof(double) * (4 * (size_t)Nr + 40 + 8 * (size_t)o->z_grid_size);
    if (smem_los > 100 * 1024) return fail(c, "pwb_rt2d_batch: z_grid_size too large for the line-of-sight kernel");
    k_rt_los<<<B, 256, smem_los, s>>>(
        B, c->geo_r.as<double>(), d_n, d_v, d_T, o->temperature_is_profile, Nr, o->z_grid_size, d_status, ncol, sums);
    c->launches++;
    PWB_CUDA(c, cudaGetLastError());
    tau_tab = nullptr;
  }
  // 'average' broadening: k_rt_sum2 (see there).  With enough (model, tile) pairs to fill the GPU one
  // CTA sums all pixel segments and phases and writes the spectrum; with fewer, one CTA per segment and
  // phase + the reduce kernel.  Same segments, same order, same arithmetic: identical bits either way.
  if (o->wind_broadening_method == 0) {
    const int Wt = (Nw <= 512) ? Nw : 512;
    const int tiles2 = (Nw + Wt - 1) / Wt;
    const int TPG = (Wt + 1) / 2;
    // (a CTA that sums all segments of a wide tile runs long: below ~16 CTAs per SM the last wave
    // would leave SMs idle)
    bool split = (long long)B * tiles2 < 16LL * c->sm_count && nsplit > 1;
    if (getenv("PWB_RT_SPLIT")) split = nsplit > 1;
    if (getenv("PWB_RT_NOSPLIT")) split = false;
    int G = split ? 1 : std::max(1, std::min(std::min(2, nsplit), 512 / TPG));
    int chunk = 1024;
    if (const char* e = getenv("PWB_RT_G")) { const int v = atoi(e); if (!split && v >= 1 && v <= 4 && v * TPG <= 1024) G = std::min(v, nsplit); }
    if (const char* e = getenv("PWB_RT_CHUNK")) { const int v = atoi(e); if (v >= 32 && v <= 4096) chunk = v; }
    RtGeoSlots slots;
    slots.n_geo = c->n_geo;
    for (int g = 0; g < 16; ++g) {
      pwb_ctx::Geo& Gg = c->geo[g < c->n_geo ? g : 0];
      slots.lo[g] = Gg.px_lo.as<int>(); slots.whi[g] = Gg.px_whi.as<double>(); slots.wlo[g] = Gg.px_wlo.as<double>();
      slots.i0[g] = Gg.px_i0.as<double>(); slots.meta[g] = Gg.px_meta.as<double>(); slots.depth[g] = Gg.depth;
    }
    const size_t smem = sizeof(double) * ((size_t)kExpWideN + 1 + Nr + Wt + (size_t)G * Wt + 2 * (size_t)G * chunk + 2);
    if (smem > 200 * 1024) return fail(c, "pwb_rt2d_batch: radius profile too long for the pixel-sum kernel");
    const int nt = G * TPG;
    const int ext = split ? nsplit : 1;
    dim3 grid2(B, tiles2, ext);
    for (int g0 = 0; g0 < (split ? c->n_geo : 1); ++g0) {
#define PW_RT_LAUNCH(MAXT)                                                                                         \
      do {                                                                                                        \
        PWB_CUDA(c, cudaFuncSetAttribute(k_rt_sum2<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_rt_sum2<MAXT><<<grid2, nt, smem, s>>>(B, Nw, Nr, *o, d_wl, d_T, ncol, sums, d_status, slots, g0, nsplit,  \
                                                ext, G, TPG, Wt, chunk, c->exp_wide.as<double>(), d_spectrum,    \
                                                partial, g0 == 0 ? d_tau : nullptr);                              \
      } while (0)
      if (nt <= 256) PW_RT_LAUNCH(256);
      else if (nt <= 512) PW_RT_LAUNCH(512);
      else PW_RT_LAUNCH(1024);
#undef PW_RT_LAUNCH
      c->launches++;
      PWB_CUDA(c, cudaGetLastError());
      if (split) {
        const size_t tot = (size_t)B * Nw;
        k_rt_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(B, Nw, nsplit, partial, d_spectrum, c->geo[g0].depth,
                                                                  g0 > 0, (g0 == c->n_geo - 1 && c->n_geo > 1) ? c->n_geo : 0);
        c->launches++;
        PWB_CUDA(c, cudaGetLastError());
      }
    }
    return 0;
  }
  dim3 grid(B, nsplit, tiles);
  // spectrum = mean over the phases of (pixel sum + transit depth of the phase), summed in
  // phase order and divided once, as numpy.mean(spectra, axis=0) does
  for (int g = 0; g < c->n_geo; ++g) {
    pwb_ctx::Geo& G = c->geo[g];
    k_rt_sum<<<grid, 256, 0, s>>>(B, Nw, Nr, *o, d_wl, d_T, ncol, sums, d_status, G.px_lo.as<int>(),
                                  G.px_whi.as<double>(), G.px_wlo.as<double>(), G.px_i0.as<double>(),
                                  G.px_meta.as<double>(), nsplit, part, g == 0 ? d_tau : nullptr, tau_tab);
    c->launches++;
    PWB_CUDA(c, cudaGetLastError());
    if (!direct) {
      const size_t tot = (size_t)B * Nw;
      k_rt_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(B, Nw, nsplit, part, d_spectrum, G.depth, g > 0,
                                                                (g == c->n_geo - 1 && c->n_geo > 1) ? c->n_geo : 0);
      c->launches++;
      PWB_CUDA(c, cudaGetLastError());
    }
  }
  return 0;
}

int pwb_rt2d_batch(pwb_ctx* c, int B, const double* d_n, const double* d_v, const double* d_T,
                   const double* d_wl, int Nw, const pwb_rt_opts* o, const int* d_status,
                   double* d_spectrum, double* d_tau, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (rt_check(c, o)) return -1;
  const int Nr = c->geo_nr;
  const int nsplit = rt_nsplit(c);
  if (c->rt_ncol.ensure(sizeof(double) * (size_t)B * Nr) || c->rt_sums.ensure(sizeof(double) * (size_t)B * 4) ||
      c->rt_partial.ensure(sizeof(double) * (size_t)B * nsplit * Nw) ||
      (o->wind_broadening_method == 1 && c->rt_tau.ensure(sizeof(double) * (size_t)B * Nr * Nw)))
    return fail(c, "out of device memory");
  return launch_rt(c, (cudaStream_t)stream, B, d_n, d_v, d_T, d_wl, Nw, o, d_status, d_spectrum, d_tau,
                   c->rt_ncol.as<double>(), c->rt_sums.as<double>(), c->rt_partial.as<double>(),
                   c->rt_tau.as<double>());
}

int pwb_forward_batch(pwb_ctx* c, int B, const double* d_T, const double* d_mdot,
                      const double* d_hfrac, const double* d_r, int Nr, const pwb_h_opts* hopts,
                      const pwb_he_opts* heopts, const pwb_rt_opts* rtopts, int species,
                      const double* d_wl, int Nw, double* d_spectrum, double* d_f_r, double* d_f1,
                      double* d_f3, double* d_v, double* d_rho, double* d_hscal, double* d_hescal,
                      int* d_status, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (!d_spectrum || !d_status) return fail(c, "pwb_forward_batch: d_spectrum and d_status are required");
  if (Nr < 2) return fail(c, "pwb_forward_batch: radius profile needs >= 2 points");
  if (hopts->exact_phi && !c->sed_ready) return fail(c, "pwb_forward_batch: exact_phi needs pwb_set_sed");
  if (species == 0 && (heopts->method < 0 || heopts->method > 2)) return fail(c, "unsupported helium method");
  if (rt_check(c, rtopts)) return -1;
  if (c->geo_nr != Nr) return fail(c, "pwb_forward_batch: pwb_set_geometry was called with a different radius grid");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bn = (size_t)B * Nr;
  const int nsplit = rt_nsplit(c);
  // scratch for the whole batch; every chunk works on its own slice
  if (c->fw_f.ensure(sizeof(double) * (bn * 5 + B)) || c->fw_scal.ensure(sizeof(double) * (size_t)B * 8) ||
      c->fw_he.ensure(sizeof(double) * (size_t)B * 4) || c->fw_n.ensure(sizeof(double) * bn) ||
      c->fw_v.ensure(sizeof(double) * bn) || c->fw_T.ensure(sizeof(double) * B) ||
      c->he_work.ensure(sizeof(double) * (size_t)kHeWorkPerRadius * Nr * B) ||
      c->he_order.ensure(sizeof(int) * (size_t)B) || c->rt_ncol.ensure(sizeof(double) * bn) ||
      c->rt_sums.ensure(sizeof(double) * (size_t)B * 4) ||
      c->rt_partial.ensure(sizeof(double) * (size_t)B * nsplit * Nw) ||
      (rtopts->wind_broadening_method == 1 && c->rt_tau.ensure(sizeof(double) * bn * Nw)))
    return fail(c, "out of device memory");
  double* base = c->fw_f.as<double>();
  double* f_r = d_f_r ? d_f_r : base;
  double* f1 = d_f1 ? d_f1 : base + bn;
  double* f3 = d_f3 ? d_f3 : base + 2 * bn;
  double* v = d_v ? d_v : base + 3 * bn;
  double* rho = d_rho ? d_rho : base + 4 * bn;
  double* mu0 = base + 5 * bn;  // [B]
  double* hscal = d_hscal ? d_hscal : c->fw_scal.as<double>();
  double* hescal = d_hescal ? d_hescal : c->fw_he.as<double>();
  // hydrogen -> helium -> line-of-sight + pixel sum, all on the caller's stream.  (Overlapping
  // the stages of different models was measured twice and dropped: equal chunks on separate
  // streams were slower, and a cost-ordered two-chunk pipeline -- costly 30 % first on a
  // high-priority stream, the cheap chunk's hydrogen stage and pixel sum under the head and
  // tail of the first helium kernel -- gave 54.4 against 56.5 ms on the 64 x 64 grid and
  // 134 against 123 ms on 24 576 models: the FP64-bound kernels of one chunk slow the
  // latency-critical helium warps of the other by about what the overlap saves.)
  k_mu0<<<(B + 255) / 256, 256, 0, s>>>(B, d_hfrac, mu0);   // quickstart.ipynb cell 2
  c->launches++;
  if (launch_hydrogen(c, s, B, d_T, d_mdot, d_hfrac, mu0, d_r, Nr, hopts, f_r, v, rho, hscal, d_status)) return -1;
  if (species == 0) {
    if (launch_helium(c, s, B, d_T, d_hfrac, hscal + 1, hscal + 2, hscal + 3, 8, d_r, Nr, v, rho, f_r, heopts,
                      c->he_work.as<double>(), c->he_order.as<int>(), f1, f3, hescal, nullptr, d_status))
      return -1;
  }
  k_he_density<<<(unsigned)((bn + 255) / 256), 256, 0, s>>>(B, Nr, species, d_hfrac, hscal, f_r, f3, rho, v, d_T,
                                                             c->fw_n.as<double>(), c->fw_v.as<double>(),
                                                             c->fw_T.as<double>());
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return launch_rt(c, s, B, c->fw_n.as<double>(), c->fw_v.as<double>(), c->fw_T.as<double>(), d_wl, Nw, rtopts,
                   d_status, d_spectrum, nullptr, c->rt_ncol.as<double>(), c->rt_sums.as<double>(),
                   c->rt_partial.as<double>(), c->rt_tau.as<double>());
}

int pwb_set_relax_warning_policy(pwb_ctx* c, int accept) {
  if (!c) return -1;
  c->accept_relax_warning = accept ? 1 : 0;
  return 0;
}

int pwb_chi2_batch(pwb_ctx* c, int B, const double* d_spectrum, int Nw, const double* d_data,
                   const double* d_sigma, double norm, const int* d_status, double* d_chi2, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  k_chi2<<<B, 128, 0, (cudaStream_t)stream>>>(B, Nw, d_spectrum, d_data, d_sigma, norm, d_status,
                                              c->accept_relax_warning, d_chi2);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_convolve_extend_batch(pwb_ctx* c, int B, const double* d_spectrum, int Nw, const double* d_kernel, int Nk,
                              double* d_out, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  if (Nk < 1 || Nk % 2 == 0) return fail(c, "pwb_convolve_extend_batch: the kernel needs an odd number of samples");
  const size_t smem = sizeof(double) * ((size_t)Nw + Nk);
  if (smem > 200 * 1024) return fail(c, "pwb_convolve_extend_batch: spectrum + kernel exceed shared memory");
  if (smem > 48 * 1024)
    PWB_CUDA(c, cudaFuncSetAttribute(k_convolve_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_convolve_extend<<<B, 256, smem, (cudaStream_t)stream>>>(B, Nw, d_spectrum, d_kernel, Nk, d_out);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_loglike_batch(pwb_ctx* c, int B, const double* d_model, int Nw, const double* d_y, const double* d_yerr,
                      const int* d_status, double* d_out, void* stream) {
  if (!c) return -1;
  if (B <= 0) return 0;
  k_loglike<<<B, 128, 0, (cudaStream_t)stream>>>(B, Nw, d_model, d_y, d_yerr, d_status,
                                                 c->accept_relax_warning, d_out);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_selftest_arith(pwb_ctx* c, int n, const double* d_a, const double* d_b, const int* d_k, double* d_quot,
                        double* d_root, void* stream) {
  if (!c) return -1;
  if (n <= 0) return 0;
  k_selftest_arith<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, d_a, d_b, d_k, d_quot, d_root);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

// the two e^x of the pixel sum on arbitrary arguments (test hook)
__global__ void k_selftest_exp(int n, const double* __restrict__ x, const double* __restrict__ exp_wide,
                               double* __restrict__ wide, double* __restrict__ gen) {
  extern __shared__ double sm_e[];
  __shared__ double s_e2[64];
  double* s_wide = sm_e;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_e2[i] = kExp2Tab64Dev[i];
  for (int i = threadIdx.x; i <= kExpWideN; i += blockDim.x) s_wide[i] = exp_wide[i];
  __syncthreads();
  double ec[8];
  for (int q = 0; q < 8; ++q) ec[q] = kExpWideC[q];
  const unsigned top = (unsigned)__cvta_generic_to_shared(s_wide + kExpWideN);
  const unsigned e2 = (unsigned)__cvta_generic_to_shared(s_e2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = x[i];
    wide[i] = (v <= 0.0 && v > -kExpWideCap) ? pw_exp_neg_wide(v, top, ec) : nan("");
    gen[i] = pw_exp_neg_sum_sh(v, e2);
  }
}

int pwb_selftest_exp(pwb_ctx* c, int n, const double* d_x, double* d_wide, double* d_general, void* stream) {
  if (!c) return -1;
  if (n <= 0) return 0;
  const size_t smem = sizeof(double) * (kExpWideN + 1);
  PWB_CUDA(c, cudaFuncSetAttribute(k_selftest_exp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_selftest_exp<<<c->sm_count, 256, smem, (cudaStream_t)stream>>>(n, d_x, c->exp_wide.as<double>(), d_wide, d_general);
  c->launches++;
  PWB_CUDA(c, cudaGetLastError());
  return 0;
}

int pwb_fp64_peak(pwb_ctx* c, double* tflops, void* stream) {
  if (!c) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = c->sm_count * 8, threads = 256, iters = 20000;
  if (c->misc.ensure(sizeof(double) * (size_t)blocks * threads)) return fail(c, "out of device memory");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fp64_peak<<<blocks, threads, 0, s>>>(c->misc.as<double>(), 1000);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, s);
    k_fp64_peak<<<blocks, threads, 0, s>>>(c->misc.as<double>(), iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  c->launches += 4;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  PWB_CUDA(c, cudaGetLastError());
  *tflops = best;
  return 0;
}

int pwb_draw_transit(pwb_ctx* c, double k, double r_phys, double b, double phase, int grid_size,
                     int supersampling, int ld_law, const double* ld, double* d_i0, double* d_r_map,
                     double* transit_depth, void* stream) {
  if (!c) return -1;
  if (grid_size < 2 || supersampling < 1 || ld_law < 0 || ld_law > 7)
    return fail(c, "pwb_draw_transit: bad grid_size / supersampling / limb-darkening law");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = grid_size * supersampling;
  DrawParams p;
  p.n = n; p.g = grid_size; p.ss = supersampling; p.cen = n / 2.0; p.r_star = n * 0.5; p.law = ld_law;
  for (int i = 0; i < 4; ++i) p.c[i] = ld ? ld[i] : 0.0;
  double r_p = p.r_star * k;
  double x_p = n / 2.0 + 2 * phase * p.r_star;
  double y_p = n / 2.0 + b * p.r_star;
  std::vector<int> spans(4 * (size_t)n);
  pil_ellipse_rows(p.cen - p.r_star, p.cen - p.r_star, p.cen + p.r_star, p.cen + p.r_star, n, &spans[0], &spans[n]);
  pil_ellipse_rows(x_p - r_p, y_p - r_p, x_p + r_p, y_p + r_p, n, &spans[2 * n], &spans[3 * n]);
  const int nblocks = c->sm_count * 4;
  if (c->misc.ensure(sizeof(int) * 4 * (size_t)n + sizeof(double) * (2 * (size_t)nblocks + 2) + 64))
    return fail(c, "out of device memory");
  double* partial = c->misc.as<double>();
  double* total = partial + 2 * nblocks;
  int* d_spans = (int*)(total + 2);
  PWB_CUDA(c, cudaMemcpyAsync(d_spans, spans.data(), sizeof(int) * 4 * (size_t)n, cudaMemcpyHostToDevice, s));
  k_draw_sums<<<nblocks, 256, 0, s>>>(p, d_spans, partial);
  k_draw_total<<<1, 32, 0, s>>>(nblocks, partial, total);
  if (supersampling > 1) { const double f = 1.0 / supersampling; x_p *= f; y_p *= f; r_p *= f; }
  k_draw_bin<<<(grid_size * grid_size + 255) / 256, 256, 0, s>>>(p, d_spans, total, x_p, y_p, r_phys / r_p, d_i0, d_r_map);
  c->launches += 3;
  PWB_CUDA(c, cudaGetLastError());
  double tot[2];
  PWB_CUDA(c, cudaMemcpyAsync(tot, total, sizeof tot, cudaMemcpyDeviceToHost, s));
  PWB_CUDA(c, cudaStreamSynchronize(s));   // also keeps `spans` alive until the copy is done
  if (transit_depth) *transit_depth = 1 - tot[1] / tot[0];
  return 0;
}

}  // extern "C"

