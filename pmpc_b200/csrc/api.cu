// C ABI of libpmpc_b200.so: handle management, layout packing kernels, kernel dispatch and the
// byte-compatible `c_lqp_solve` / `c_lcone_solve` entry points (see include/pmpc_b200.h).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pmpc_b200.h"
#include "kernels.cuh"

#include <atomic>
static std::atomic<long long> g_launches{0};  // kernels launched by this library (pmpc_launch_count)

using namespace pmpc;

#include "variants.h"
namespace pmpc {
#define PMPC_DECL(nx, nu, kind, var) \
  cudaError_t launch_##nx##_##nu##_##kind##_##var(const DevProb&, int blocks, int threads, cudaStream_t, int*);
PMPC_VARIANTS(PMPC_DECL)
#undef PMPC_DECL
}  // namespace pmpc

// warp-per-particle kernels for the NX >= 8 shapes (wpp_inst.cu): X(NX, NU, SLEW)
#include "wpp_plan.h"
#define PMPC_WPP_VARIANTS(X) X(8, 4, 0) X(8, 4, 1) X(12, 4, 0) X(12, 4, 1) X(16, 8, 0) X(16, 8, 1) X(32, 8, 0) X(32, 8, 1)
namespace pmpc {
#define PMPC_WDECL(nx, nu, sl)                                                     \
  cudaError_t wpp_plan_##nx##_##nu##_##sl(const DevProb&, WppPlan*);               \
  cudaError_t wpp_launch_##nx##_##nu##_##sl(const DevProb&, const WppPlan&, cudaStream_t);
PMPC_WPP_VARIANTS(PMPC_WDECL)
#undef PMPC_WDECL
}  // namespace pmpc
struct WppEntry {
  int nx, nu, slew;
  cudaError_t (*plan)(const DevProb&, WppPlan*);
  cudaError_t (*launch)(const DevProb&, const WppPlan&, cudaStream_t);
};
static const WppEntry kWpp[] = {
#define PMPC_WENTRY(nx, nu, sl) {nx, nu, sl, &pmpc::wpp_plan_##nx##_##nu##_##sl, &pmpc::wpp_launch_##nx##_##nu##_##sl},
    PMPC_WPP_VARIANTS(PMPC_WENTRY)
#undef PMPC_WENTRY
};
static const int kWppDmax = 64;  // == Wpp::DMAX

typedef cudaError_t (*launch_fn)(const DevProb&, int, int, cudaStream_t, int*);
struct VariantEntry { int nx, nu, kind, var; launch_fn fn; };
static const VariantEntry kVariants[] = {
#define PMPC_ENTRY(nx, nu, kind, var) {nx, nu, kind, var, &pmpc::launch_##nx##_##nu##_##kind##_##var},
    PMPC_VARIANTS(PMPC_ENTRY)
#undef PMPC_ENTRY
};

// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(PMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

struct Shape { int NX, NU; };
static const Shape kShapes[] = {{2, 1}, {4, 2}, {8, 4}, {12, 4}, {16, 8}, {32, 8}};

static bool pick_shape(int n, int m, Shape* out) {
  for (const Shape& s : kShapes)
    if (s.NX >= n && s.NU >= m) { *out = s; return true; }
  return false;
}

struct pmpc_userdyn_s;  // userdyn.inl

struct pmpc_handle_s {
  int device = 0, num_sms = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t tev0 = nullptr, tev1 = nullptr;  // pmpc_timer_start / pmpc_timer_stop
  DevProb pb{};
  Shape shape{};
  long long Nc_orig = 0;
  size_t Ptot = 0;            // B*M
  void* slab = nullptr;       // all device arrays
  size_t slab_bytes = 0;
  double* stage = nullptr;    // device staging for raw (reference-layout) inputs / outputs
  size_t stage_doubles = 0;
  double* raw = nullptr;      // device copy of all raw pmpc_set_problem inputs (grown on demand)
  size_t raw_doubles = 0;
  int* flags = nullptr;       // device-side NaN-sentinel analysis of the optional arguments (FL_*)
  int row_major = 0;          // matrices given as C-order (rows, cols) blocks instead of column-major
  unsigned bc_all = 0, bc_prob = 0;  // pmpc_set_broadcast: inputs given once for the whole batch / once per problem
  double* comm = nullptr;     // particle sharding: this rank's exchange buffer (IPC-exported to the peers)
  size_t comm_bytes = 0;
  void* peer_open[8] = {nullptr};  // peers' buffers opened with cudaIpcOpenMemHandle (to close on destroy)
  unsigned long long seq = 0;  // exchange sequence base, advanced identically on every rank per launch
  double *lin_f = nullptr, *lin_fx = nullptr, *lin_fu = nullptr;  // reference-layout linearisation
  int* uniform = nullptr;     // per problem: cost / reference / bound data identical for all particles
  int* xerr = nullptr;        // device flag raised when an in-kernel exchange times out (particle sharding)
  // warp-per-particle kernel: team exchange buffers, consensus partials, hand-over flags (DevProb::tm / wscr / fb)
  double* tm = nullptr;
  size_t tm_doubles = 0;
  int tm_Tc = -1, tm_cap = -1;
  double* wscr = nullptr;
  size_t wscr_doubles = 0;
  int* fb = nullptr;
  unsigned long long tseq = 0;
  double* hist = nullptr;     // (4, hist_len, B) iteration history, grown to max_it by pmpc_scp_solve_async
  bool have_problem = false, have_dyn = false, solved = false;
  pmpc_settings settings{};
  float last_ms = 0.f;
  pmpc_userdyn_s* user = nullptr;              // registered user device dynamics (pmpc_register_dynamics)
  cudaEvent_t uev0 = nullptr, uev1 = nullptr;  // span of the host-driven SCP loop
  bool user_timed = false;                     // last_ms already holds that span
};

// ---------------------------------------------------------------------------------------------
// packing kernels: reference layout (particle-major, column-major blocks) -> warp-tiled SoA
// (solver.cuh "DATA LAYOUT").  One thread per (flat particle gf = b*M + i, stage j), gf fastest so
// that the tile stores are coalesced.
struct Lay { int M, Mpad, N, tile; };  // tile: particles per layout tile (32, or 1 for the NX >= 8 shapes)
__device__ __forceinline__ size_t slot_of(int gf, Lay L) { return (size_t)(gf / L.M) * L.Mpad + (size_t)(gf % L.M); }
__device__ __forceinline__ size_t tix(size_t slot, Lay L, int j, int E, int e) {
  return (((slot / L.tile) * (size_t)L.N + (size_t)j) * (size_t)E + (size_t)e) * L.tile + (slot % L.tile);
}

// pmpc_set_broadcast: which block of an input array particle gf reads — its own (mode 0), its problem's (1), the only one (2)
enum { BC_Q = 0, BC_R, BC_XREF, BC_UREF, BC_LX, BC_UX, BC_LU, BC_UU, BC_XPREV, BC_UPREV, BC_COUNT };
__host__ __device__ __forceinline__ size_t src_row(int gf, int M, int mode) { return mode == 2 ? 0 : (mode == 1 ? (size_t)(gf / M) : (size_t)gf); }

enum { FL_XBOX_ABSENT = 0, FL_UBOX_ABSENT, FL_SLEW_ABSENT, FL_SLEW0_ABSENT, FL_SLEW_NONZERO, FL_COUNT = 8 };

// flag = 1 when any element is NaN (the reference's "option absent" sentinel, c_interface.jl:56-70)
__global__ void nan_flag_kernel(const double* __restrict__ a, size_t n, int* flag) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= (a[i] != a[i]);
  if (bad) *flag = 1;
}
__global__ void nonzero_flag_kernel(const double* __restrict__ a, size_t n, int* flag) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool nz = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) nz |= (a[i] != 0.0);
  if (nz) *flag = 1;
}

// absent: optional device flag; when set the whole array counts as absent -> `fill`
__global__ void pack_vec_kernel(const double* __restrict__ in, double* __restrict__ out, int Ptot, Lay L, int Ei, int Eo,
                                double fill, const int* __restrict__ absent, int mode = 0) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = L.N;
  if (idx >= (size_t)Ptot * N) return;
  if (absent != nullptr && *absent) in = nullptr;
  const int gf = (int)(idx % Ptot), j = (int)(idx / Ptot);
  const size_t sl = slot_of(gf, L);
  const size_t gs = src_row(gf, L.M, mode);
  for (int e = 0; e < Eo; ++e) {
    double v = fill;
    if (in != nullptr && e < Ei) v = in[(gs * N + j) * Ei + e];
    out[tix(sl, L, j, Eo, e)] = v;
  }
}

// lead = 1 writes the trajectory form (N+1 stages per particle, stage 0 = x0: what `solve` returns)
__global__ void unpack_vec_kernel(const double* __restrict__ in, double* __restrict__ out, int Ptot, Lay L, int Ei,
                                  int Eo, int lead, const double* __restrict__ x0, size_t P) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = L.N;
  if (idx >= (size_t)Ptot * N) return;
  const int gf = (int)(idx % Ptot), j = (int)(idx / Ptot);
  const size_t sl = slot_of(gf, L);
  for (int e = 0; e < Eo; ++e) out[((size_t)gf * (N + lead) + j + lead) * Eo + e] = in[tix(sl, L, j, Ei, e)];
  if (lead && j == 0)
    for (int e = 0; e < Eo; ++e) out[(size_t)gf * (N + 1) * Eo + e] = x0[(size_t)e * P + sl];
}

__global__ void pack_particle_kernel(const double* __restrict__ in, double* __restrict__ out, int Ptot, size_t P, Lay L,
                                     int Ei, int Eo, double fill) {
  const int gf = blockIdx.x * blockDim.x + threadIdx.x;
  if (gf >= Ptot) return;
  const size_t sl = slot_of(gf, L);
  for (int e = 0; e < Eo; ++e) out[(size_t)e * P + sl] = (in != nullptr && e < Ei) ? in[(size_t)gf * Ei + e] : fill;
}

// cost terms (lqp_utils.jl:26-40, 130-142, 164-214): Qs = triu(Q) mirrored + reg_x I, etc.
__global__ void pack_cost_kernel(DevProb pb, int NX, int NU, const double* __restrict__ Q, const double* __restrict__ R,
                                 const double* __restrict__ Xref, const double* __restrict__ Uref,
                                 const double* __restrict__ s, const double* __restrict__ s0,
                                 const double* __restrict__ um1, int Ptot, int quirk_keep_um1, int row_major,
                                 const int* __restrict__ flags, unsigned bc_all, unsigned bc_prob) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = pb.N, n = pb.n, m = pb.m;
  if (idx >= (size_t)Ptot * N) return;
  const int g = (int)(idx % Ptot), j = (int)(idx / Ptot);
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  const size_t sl = slot_of(g, L);
  const size_t P = pb.P;
  auto mode = [&](int k) { return ((bc_all >> k) & 1u) ? 2 : (((bc_prob >> k) & 1u) ? 1 : 0); };
  const double* Qb = Q + (src_row(g, pb.M, mode(BC_Q)) * N + j) * n * n;  // column-major: Q[r,t] = Qb[t*n + r]
  const double* Rb = R + (src_row(g, pb.M, mode(BC_R)) * N + j) * m * m;
  // element (r, t) of an (a x a) block: column-major (reference C ABI) or row-major (NumPy C order)
  const int qr = row_major ? n : 1, qt = row_major ? 1 : n, rr_ = row_major ? m : 1, rt = row_major ? 1 : m;
  if (flags[FL_SLEW_ABSENT]) s = nullptr;
  if (flags[FL_SLEW0_ABSENT]) { s0 = nullptr; um1 = nullptr; }
  const double* xr = Xref ? Xref + (src_row(g, pb.M, mode(BC_XREF)) * N + j) * n : nullptr;
  const double* ur = Uref ? Uref + (src_row(g, pb.M, mode(BC_UREF)) * N + j) * m : nullptr;
  const int NQ = NX * (NX + 1) / 2, NR = NU * (NU + 1) / 2;
  for (int r = 0; r < NX; ++r) {
    for (int t = r; t < NX; ++t) {
      double v = (r < n && t < n) ? Qb[(size_t)t * qt + (size_t)r * qr] : 0.0;
      if (r == t) v += (r < n) ? pb.reg_x : 1.0;
      pb.Qs[tix(sl, L, j, NQ, pk(r, t, NX))] = v;
    }
    double acc = 0.0;
    if (r < n && xr)
      for (int t = 0; t < n; ++t) acc += Qb[(size_t)t * qt + (size_t)r * qr] * xr[t];
    pb.qref[tix(sl, L, j, NX, r)] = acc;
    pb.xref[tix(sl, L, j, NX, r)] = (r < n && xr) ? xr[r] : 0.0;
  }
  const double sv = s ? s[g] : 0.0, s0v = s0 ? s0[g] : 0.0;
  const double extra = (j == 0 ? s0v : 0.0) + ((N == 1 && j == 0) ? sv : 0.0);  // lqp_utils.jl:31-37
  for (int a = 0; a < NU; ++a) {
    for (int b = a; b < NU; ++b) {
      double v = (a < m && b < m) ? Rb[(size_t)b * rt + (size_t)a * rr_] : 0.0;
      if (a == b) v += (a < m) ? (pb.reg_u + extra) : 1.0;
      pb.Rs[tix(sl, L, j, NR, pk(a, b, NU))] = v;
    }
    double acc = 0.0;
    if (a < m && ur)
      for (int t = 0; t < m; ++t) acc += Rb[(size_t)t * rt + (size_t)a * rr_] * ur[t];
    if (a < m && j == 0 && quirk_keep_um1 && um1) acc += s0v * um1[(size_t)g * m + a];  // lqp_utils.jl:165
    pb.rref[tix(sl, L, j, NU, a)] = acc;
    pb.uref[tix(sl, L, j, NU, a)] = (a < m && ur) ? ur[a] : 0.0;
  }
  if (j == 0) {
    pb.slew[sl] = sv;
    pb.rdiag0[sl] = extra;
  }
  (void)P;
}

// caller-supplied linearisation -> A, B, c   (lqp_utils.jl:231-300: c_j = f_j - fu_j ubar_j - fx_j xbar_{j-1})
__global__ void pack_lin_kernel(DevProb pb, int NX, int NU, const double* __restrict__ f, const double* __restrict__ fx,
                                const double* __restrict__ fu, int Ptot, int c_is_f, int row_major) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = pb.N, n = pb.n, m = pb.m;
  if (idx >= (size_t)Ptot * N) return;
  const int g = (int)(idx % Ptot), j = (int)(idx / Ptot);
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  const size_t sl = slot_of(g, L);
  const size_t P = pb.P;
  const double* fxb = fx + ((size_t)g * N + j) * n * n;  // fx[r,t] = fxb[t*n + r]
  const double* fub = fu + ((size_t)g * N + j) * n * m;  // fu[r,t] = fub[t*n + r]
  const double* fb = f ? f + ((size_t)g * N + j) * n : nullptr;
  const int xr_ = row_major ? n : 1, xt = row_major ? 1 : n, ur_ = row_major ? m : 1, ut = row_major ? 1 : n;
  for (int r = 0; r < NX; ++r) {
    double cv = (r < n && fb) ? fb[r] : 0.0;
    for (int t = 0; t < NX; ++t) {
      const double a0 = (r < n && t < n) ? fxb[(size_t)t * xt + (size_t)r * xr_] : 0.0;
      const double a = (j >= 1) ? a0 : 0.0;  // stage 1 is linearised at x0 itself: no fx term (lqp_utils.jl:288-296)
      pb.A[tix(sl, L, j, NX * NX, r * NX + t)] = a;
      if (!c_is_f && j >= 1) cv -= a * pb.Xp[tix(sl, L, j - 1, NX, t)];
      if (c_is_f && j == 0) cv += a0 * pb.x0[(size_t)t * P + sl];  // LTV: f_1 = A_1 x0 + B_1 u + c_1
    }
    for (int t = 0; t < NU; ++t) {
      const double b = (r < n && t < m) ? fub[(size_t)t * ut + (size_t)r * ur_] : 0.0;
      pb.Bm[tix(sl, L, j, NX * NU, r * NU + t)] = b;
      if (!c_is_f) cv -= b * pb.Up[tix(sl, L, j, NU, t)];
    }
    pb.c[tix(sl, L, j, NX, r)] = cv;
  }
}

// K1 stand-alone: Dubins f, fx, fu at (X_ = [x0, X_prev[:-1]], U_prev) in reference layout
__global__ void linearize_dubins_kernel(DevProb pb, double* __restrict__ f, double* __restrict__ fx,
                                        double* __restrict__ fu, int Ptot, int row_major) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = pb.N;
  if (idx >= (size_t)Ptot * N) return;
  const int g = (int)(idx % Ptot), j = (int)(idx / Ptot);
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  const size_t sl = slot_of(g, L);
  const size_t P = pb.P;
  double x[4], u[2], fo[4], fxo[16], fuo[8];
  for (int e = 0; e < 4; ++e) x[e] = (j == 0) ? pb.x0[(size_t)e * P + sl] : pb.Xp[tix(sl, L, j - 1, 4, e)];
  for (int e = 0; e < 2; ++e) u[e] = pb.Up[tix(sl, L, j, 2, e)];
  dubins_f_fx_fu(x, u, pb.dynp[sl], pb.dynp[P + sl], pb.dynp[2 * P + sl], pb.dyn_eps, fo, fxo, fuo);
  for (int r = 0; r < 4; ++r) {
    f[((size_t)g * N + j) * 4 + r] = fo[r];
    for (int t = 0; t < 4; ++t) fx[((size_t)g * N + j) * 16 + (row_major ? r * 4 + t : t * 4 + r)] = fxo[r * 4 + t];
    for (int t = 0; t < 2; ++t) fu[((size_t)g * N + j) * 8 + (row_major ? r * 2 + t : t * 4 + r)] = fuo[r * 2 + t];
  }
}

// flags[b] = 0 when any particle of problem b differs from particle 0 in its packed cost, reference
// or bound data (flags pre-set to 1).  One thread per (flat particle, stage).
__global__ void uniform_check_kernel(DevProb pb, int NX, int NU, int* flags, int Ptot) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = pb.N;
  if (idx >= (size_t)Ptot * N) return;
  const int gf = (int)(idx % Ptot), j = (int)(idx / Ptot);
  if (gf % pb.M == 0) return;
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  const size_t sl = slot_of(gf, L), s0 = slot_of(gf - gf % pb.M, L);
  const int NQ = NX * (NX + 1) / 2, NR = NU * (NU + 1) / 2;
  bool same = pb.rdiag0[sl] == pb.rdiag0[s0];
  for (int e = 0; e < NQ; ++e) same &= pb.Qs[tix(sl, L, j, NQ, e)] == pb.Qs[tix(s0, L, j, NQ, e)];
  for (int e = 0; e < NR; ++e) same &= pb.Rs[tix(sl, L, j, NR, e)] == pb.Rs[tix(s0, L, j, NR, e)];
  for (int e = 0; e < NX; ++e) {
    same &= pb.qref[tix(sl, L, j, NX, e)] == pb.qref[tix(s0, L, j, NX, e)];
    same &= pb.xref[tix(sl, L, j, NX, e)] == pb.xref[tix(s0, L, j, NX, e)];
  }
  for (int e = 0; e < NU; ++e) {
    same &= pb.rref[tix(sl, L, j, NU, e)] == pb.rref[tix(s0, L, j, NU, e)];
    same &= pb.uref[tix(sl, L, j, NU, e)] == pb.uref[tix(s0, L, j, NU, e)];
    same &= pb.lu[tix(sl, L, j, NU, e)] == pb.lu[tix(s0, L, j, NU, e)];
    same &= pb.uu[tix(sl, L, j, NU, e)] == pb.uu[tix(s0, L, j, NU, e)];
  }
  if (!same) flags[gf / pb.M] = 0;
}

__global__ void fill_int_kernel(int* p, size_t n, int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void fill_kernel(double* p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
static inline int nblk(size_t n, int t = 256) { return (int)((n + t - 1) / t); }

extern "C" {

void pmpc_default_settings(pmpc_settings* s) {
  s->rho0 = 1.0;
  s->alpha = 1.6;
  s->eps_abs = 1e-9;
  s->eps_rel = 1e-9;
  s->kkt_tol = 1e-9;
  s->adapt_tol = 3.0;
  s->max_admm_iter = 4000;
  s->check_every = 10;
  s->polish = 1;
  s->polish_iters = 16;
  s->adapt_rho = 1;
  s->warm_start = 1;
  s->fused = 1;
  s->ring = 1;
}

const char* pmpc_last_error(void) { return g_err.c_str(); }
const char* pmpc_version(void) { return "pmpc_b200 0.1.0 (sm_100a)"; }

int pmpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static void apply_settings(pmpc_handle_t h) {
  const pmpc_settings& s = h->settings;
  DevProb& pb = h->pb;
  pb.rho0 = s.rho0; pb.alpha = s.alpha; pb.eps_abs = s.eps_abs; pb.eps_rel = s.eps_rel;
  pb.kkt_tol = s.kkt_tol; pb.adapt_tol = s.adapt_tol; pb.max_admm_iter = s.max_admm_iter;
  pb.check_every = s.check_every < 1 ? 1 : s.check_every; pb.polish = s.polish;
  pb.polish_iters = s.polish_iters < 1 ? 1 : s.polish_iters; pb.adapt_rho = s.adapt_rho;
  pb.warm_start = s.warm_start;
  pb.fused = s.fused;
  pb.ring = s.ring;
}

static void userdyn_free(pmpc_userdyn_s* u);

int pmpc_create(pmpc_handle_t* out, int device, size_t B, size_t M, size_t N, size_t xdim, size_t udim,
                long long Nc) {
  if (!out || B < 1 || M < 1 || N < 1 || xdim < 1 || udim < 1) return fail(PMPC_ERR_BAD_ARG, "bad sizes");
  if (pmpc_device_count() < 1)
    return fail(PMPC_ERR_NO_DEVICE, "no CUDA device: pmpc_b200 has no CPU fallback");
  Shape sh;
  if (!pick_shape((int)xdim, (int)udim, &sh))
    return fail(PMPC_ERR_UNSUPPORTED, "xdim/udim outside the compiled kernel table (max 32/8)");
  if (M > 1024) return fail(PMPC_ERR_UNSUPPORTED, "M > 1024 particles per problem per GPU: shard the particles");
  if (Nc < 0 || Nc > (long long)N) Nc = (long long)N;  // main.jl:127-128
  pmpc_handle_s* h = new pmpc_handle_s();
  // every error return below releases what has been created so far (the slab alone can be tens of GB)
  struct Guard {
    pmpc_handle_s* h;
    ~Guard() { if (h) pmpc_destroy(h); }
  } guard{h};
  if (device >= 0) {
    CU(cudaSetDevice(device));
    h->device = device;
  } else {
    CU(cudaGetDevice(&h->device));
  }
  CU(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaEventCreate(&h->tev0));
  CU(cudaEventCreate(&h->tev1));
  CU(cudaEventCreate(&h->uev0));
  CU(cudaEventCreate(&h->uev1));
  h->shape = sh;
  h->Nc_orig = Nc;
  h->Ptot = B * M;
  DevProb& pb = h->pb;
  pb.B = (int)B; pb.M = (int)M; pb.N = (int)N; pb.n = (int)xdim; pb.m = (int)udim;
  pb.Nc = (M == 1) ? 0 : (int)Nc;  // one particle: consensus is vacuous (the Nc==0 quirk is applied at pack time)
  pb.D = pb.Nc * sh.NU;
  pb.tile = (sh.NX >= 8) ? 1 : 32;  // == Solver<NX, NU>::TILE
  // 2 <= M <= 16 particles: two problems share a 32-slot tile (the two-problems-per-warp kernel, kind 3; the
  // one-problem-per-warp kernels read the same layout: their mirror lanes stay inside the problem's own 16 slots)
  pb.Mpad = (M == 1 || pb.tile == 1) ? (int)M : ((M <= 16) ? 16 : (int)(((M + 31) / 32) * 32));
  pb.P = (int)((((size_t)B * pb.Mpad + 31) / 32) * 32);
  pb.hist_len = 128;
  pb.rank = 0; pb.world = 1; pb.M_total = (int)M; pb.comm_cap = 0; pb.seq0 = 0;
  pmpc_default_settings(&h->settings);
  apply_settings(h);
  // ---- carve one slab
  const size_t P = pb.P, NX = sh.NX, NU = sh.NU, NQ = NX * (NX + 1) / 2, NR = NU * (NU + 1) / 2, Nn = N;
  const int T = (int)(((M + 31) / 32) * 32);
  const size_t D = pb.D;
  size_t cols = D < 8 ? D : 8;
  pb.scratch_stride = std::max<size_t>(D * D + D + 32, (size_t)T * std::max<size_t>(D * cols, (size_t)ST_COUNT));
  std::vector<std::pair<double**, size_t>> req;
  auto want = [&](double** p, size_t cnt) { req.push_back({p, cnt}); };
  want(&pb.Qs, Nn * NQ * P); want(&pb.Rs, Nn * NR * P); want(&pb.qref, Nn * NX * P); want(&pb.rref, Nn * NU * P);
  want(&pb.xref, Nn * NX * P); want(&pb.uref, Nn * NU * P);
  want(&pb.lu, Nn * NU * P); want(&pb.uu, Nn * NU * P); want(&pb.lx, Nn * NX * P); want(&pb.ux, Nn * NX * P);
  want(&pb.x0, NX * P); want(&pb.slew, P); want(&pb.dynp, 4 * P); want(&pb.rdiag0, P);
  want(&pb.Xp, Nn * NX * P); want(&pb.Up, Nn * NU * P); want(&pb.A, Nn * NX * NX * P); want(&pb.Bm, Nn * NX * NU * P);
  want(&pb.c, Nn * NX * P); want(&pb.q, Nn * NX * P); want(&pb.r, Nn * NU * P);
  for (int k = 0; k < 2; ++k) {
    FactorSet& F = pb.fs[k];
    want(&F.K, Nn * NU * NX * P); want(&F.E, Nn * NR * P); want(&F.d, Nn * NX * P); want(&F.e, Nn * NU * P);
    want(&F.Pc, NX * NX * P); want(&F.Wc, NX * NU * P); want(&F.Zc, NU * NU * P);
    want(&F.Hc, B * D * D + 1); want(&F.Hci, B * D * D + 1); want(&F.hb, B * D + 1);
  }
  want(&pb.kff, Nn * NU * P); want(&pb.hand, Nn * (((NX == 4 && NU == 2) ? 10 : 0) + NU * NX + NU + (NR - NU)) * P); want(&pb.X, Nn * NX * P); want(&pb.U, Nn * NU * P);
  want(&pb.wu, Nn * NU * P); want(&pb.yu, Nn * NU * P); want(&pb.wx, Nn * NX * P); want(&pb.yx, Nn * NX * P);
  want(&pb.gc, B * D + 1); want(&pb.uc, B * D + 1); want(&pb.muc, B * D + 1);
  want(&pb.scratch, B * pb.scratch_stride);
  want(&pb.rho, B);
  size_t total = 0;
  for (auto& r : req) total += (r.second + 31) / 32 * 32;
  const size_t act_bytes = 2 * (Nn * NU * P) + 2 * (B * D + 32) + 64;
  const size_t info_bytes = 8 * B * sizeof(long long);
  h->slab_bytes = total * sizeof(double) + act_bytes + info_bytes + 256;
  cudaError_t e = cudaMalloc(&h->slab, h->slab_bytes);
  if (e != cudaSuccess) {
    h->slab = nullptr;
    return fail(PMPC_ERR_CUDA, std::string("cudaMalloc slab: ") + cudaGetErrorString(e));
  }
  CU(cudaMemsetAsync(h->slab, 0, h->slab_bytes, h->stream));
  double* cur = (double*)h->slab;
  for (auto& r : req) {
    *r.first = cur;
    cur += (r.second + 31) / 32 * 32;
  }
  pb.info = (long long*)cur;
  signed char* ac = (signed char*)(pb.info + 8 * B);
  pb.act = ac; ac += Nn * NU * P;
  pb.actn = ac; ac += Nn * NU * P;
  pb.actc = ac; ac += B * D + 32;
  pb.actcn = ac;
  // staging: the largest raw input is Q or fx: Ptot*N*n*n doubles (+ room for X/U outputs)
  h->stage_doubles = h->Ptot * (N + 1) * std::max<size_t>(xdim * xdim, xdim * udim + xdim + udim + 16);
  CU(cudaMalloc(&h->stage, h->stage_doubles * sizeof(double)));
  CU(cudaMalloc(&h->lin_f, h->Ptot * N * xdim * sizeof(double)));
  CU(cudaMalloc(&h->lin_fx, h->Ptot * N * xdim * xdim * sizeof(double)));
  CU(cudaMalloc(&h->lin_fu, h->Ptot * N * xdim * udim * sizeof(double)));
  CU(cudaMalloc(&h->flags, FL_COUNT * sizeof(int)));
  CU(cudaMalloc(&h->uniform, B * sizeof(int)));
  CU(cudaMemsetAsync(h->uniform, 0, B * sizeof(int), h->stream));
  pb.uniform = h->uniform;
  CU(cudaMalloc(&h->hist, (size_t)4 * pb.hist_len * B * sizeof(double)));
  CU(cudaMemsetAsync(h->hist, 0, (size_t)4 * pb.hist_len * B * sizeof(double), h->stream));
  pb.hist = h->hist;
  CU(cudaMalloc(&h->xerr, sizeof(int)));
  CU(cudaMemsetAsync(h->xerr, 0, sizeof(int), h->stream));
  pb.xerr = h->xerr;
  pb.xtimeout_ns = 30ull * 1000000000ull;
  // (+ the pair queue of the two-problems-per-warp kernel, + one byte per particle: any state multiplier non-zero)
  CU(cudaMalloc(&h->fb, (B + 1) * sizeof(int) + P));
  CU(cudaMemsetAsync(h->fb, 0, (B + 1) * sizeof(int) + P, h->stream));
  pb.fb = h->fb;
  pb.queue = h->fb + B;
  pb.xany = reinterpret_cast<unsigned char*>(h->fb + B + 1);
  pb.resume = 0;
  pb.tm = nullptr; pb.wscr = nullptr; pb.tm_cap = 0; pb.tseq0 = 0;
  fill_kernel<<<nblk(B), 256, 0, h->stream>>>(pb.rho, B, pb.rho0);
  ++g_launches;
  CU(cudaGetLastError());
  guard.h = nullptr;
  *out = h;
  return PMPC_OK;
}

int pmpc_destroy(pmpc_handle_t h) {
  if (!h) return PMPC_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  userdyn_free(h->user);
  for (cudaEvent_t ev : {h->uev0, h->uev1, h->ev0, h->ev1, h->tev0, h->tev1})
    if (ev) cudaEventDestroy(ev);
  cudaFree(h->slab); cudaFree(h->stage); cudaFree(h->lin_f); cudaFree(h->lin_fx); cudaFree(h->lin_fu);
  cudaFree(h->uniform); cudaFree(h->flags); cudaFree(h->raw); cudaFree(h->hist); cudaFree(h->xerr); cudaFree(h->tm); cudaFree(h->wscr); cudaFree(h->fb);
  for (int r = 0; r < 8; ++r)
    if (h->peer_open[r]) cudaIpcCloseMemHandle(h->peer_open[r]);
  cudaFree(h->comm);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  return PMPC_OK;
}

int pmpc_set_settings(pmpc_handle_t h, const pmpc_settings* s) {
  if (!h || !s) return fail(PMPC_ERR_BAD_ARG, "null");
  h->settings = *s;
  apply_settings(h);
  CU(cudaSetDevice(h->device));
  fill_kernel<<<nblk(h->pb.B), 256, 0, h->stream>>>(h->pb.rho, h->pb.B, h->pb.rho0);
  ++g_launches;
  CU(cudaGetLastError());
  return PMPC_OK;
}

// upload `cnt` doubles into the staging buffer at offset `off` (returns device pointer)
static int upload(pmpc_handle_t h, const double* src, size_t cnt, size_t off, double** dst) {
  if (off + cnt > h->stage_doubles) return fail(PMPC_ERR_BAD_ARG, "staging overflow");
  *dst = h->stage + off;
  CU(cudaMemcpyAsync(*dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return PMPC_OK;
}

// stream-ordered: no synchronisation here; `off` keeps concurrent uploads apart in the staging buffer
static int pack_vec(pmpc_handle_t h, const double* host, double* out, int Ei, int Eo, double fill, size_t off = 0, int mode = 0) {
  const DevProb& pb = h->pb;
  double* dev = nullptr;
  if (host) {
    const size_t blocks = mode == 2 ? 1 : (mode == 1 ? (size_t)pb.B : h->Ptot);
    int rc = upload(h, host, blocks * pb.N * Ei, off, &dev);
    if (rc) return rc;
  }
  pack_vec_kernel<<<nblk(h->Ptot * pb.N), 256, 0, h->stream>>>(dev, out, (int)h->Ptot, Lay{pb.M, pb.Mpad, pb.N, pb.tile}, Ei, Eo, fill,
                                                              nullptr, mode);
  ++g_launches;
  CU(cudaGetLastError());
  return PMPC_OK;
}

int pmpc_set_layout(pmpc_handle_t h, int row_major_matrices) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  h->row_major = row_major_matrices ? 1 : 0;
  return PMPC_OK;
}

int pmpc_set_broadcast(pmpc_handle_t h, unsigned all_mask, unsigned per_problem_mask) {
  if (!h || (all_mask >> BC_COUNT) || (per_problem_mask >> BC_COUNT)) return fail(PMPC_ERR_BAD_ARG, "null handle or unknown input bit");
  h->bc_all = all_mask;
  h->bc_prob = per_problem_mask & ~all_mask;
  return PMPC_OK;
}

int pmpc_host_alloc(void** p, size_t bytes) {
  if (!p) return fail(PMPC_ERR_BAD_ARG, "null");
  if (pmpc_device_count() < 1) return fail(PMPC_ERR_NO_DEVICE, "no CUDA device");
  CU(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return PMPC_OK;
}

int pmpc_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return PMPC_OK;
}

// All raw inputs go to one device buffer with back-to-back async copies; the NaN-sentinel analysis
// (c_interface.jl:56-70) and the re-layout run on the device; ONE synchronisation at the end reads the flags.
int pmpc_set_problem(pmpc_handle_t h, const double* x0, const double* Q, const double* R, const double* X_ref,
                     const double* U_ref, const double* lx, const double* ux, const double* lu, const double* uu,
                     double reg_x, double reg_u, const double* slew_reg, const double* slew_reg0,
                     const double* slew_um1) {
  if (!h || !x0 || !Q || !R) return fail(PMPC_ERR_BAD_ARG, "x0, Q, R are required");
  CU(cudaSetDevice(h->device));
  DevProb& pb = h->pb;
  const int n = pb.n, m = pb.m, N = pb.N, NX = h->shape.NX, NU = h->shape.NU;
  const size_t Pt = h->Ptot;
  pb.reg_x = reg_x; pb.reg_u = reg_u;
  auto bmode = [&](int k) { return ((h->bc_all >> k) & 1u) ? 2 : (((h->bc_prob >> k) & 1u) ? 1 : 0); };
  auto rows = [&](int k) -> size_t { return bmode(k) == 2 ? 1 : (bmode(k) == 1 ? (size_t)pb.B : Pt); };  // blocks of input k
  const size_t nQ = rows(BC_Q) * N * n * n, nR = rows(BC_R) * N * m * m, nXr = rows(BC_XREF) * N * n, nUr = rows(BC_UREF) * N * m;
  const size_t nLx = rows(BC_LX) * N * n, nUx = rows(BC_UX) * N * n, nLu = rows(BC_LU) * N * m, nUu = rows(BC_UU) * N * m;
  const size_t total = nQ + nR + nXr + nUr + nLx + nUx + nLu + nUu + Pt * (2 + m) + Pt * n + 64;
  if (total > h->raw_doubles) {
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->raw);
    h->raw = nullptr;
    h->raw_doubles = 0;
    CU(cudaMalloc(&h->raw, total * sizeof(double)));
    h->raw_doubles = total;
  }
  double* cur = h->raw;
  auto up = [&](const double* src, size_t cnt, double** dst) -> int {
    *dst = nullptr;
    if (!src) return 0;
    *dst = cur;
    cur += (cnt + 1) / 2 * 2;  // keep 16-byte alignment
    cudaError_t e = cudaMemcpyAsync(*dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    return e == cudaSuccess ? 0 : fail(PMPC_ERR_CUDA, cudaGetErrorString(e));
  };
  double *dQ, *dR, *dXr, *dUr, *dlx, *dux, *dlu, *duu, *ds, *ds0, *dum1, *dx0;
  int rc;
  // absent-by-NULL is decided here, absent-by-NaN on the device
  int init[FL_COUNT] = {0};
  init[FL_XBOX_ABSENT] = (!lx || !ux);
  init[FL_UBOX_ABSENT] = (!lu || !uu);
  init[FL_SLEW_ABSENT] = !slew_reg;
  init[FL_SLEW0_ABSENT] = (!slew_reg0 || !slew_um1);
  CU(cudaMemcpyAsync(h->flags, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  if ((rc = up(lx, nLx, &dlx)) || (rc = up(ux, nUx, &dux)) || (rc = up(lu, nLu, &dlu)) || (rc = up(uu, nUu, &duu)) ||
      (rc = up(slew_reg, Pt, &ds)) || (rc = up(slew_reg0, Pt, &ds0)) || (rc = up(slew_um1, Pt * m, &dum1)) ||
      (rc = up(x0, Pt * n, &dx0)) || (rc = up(X_ref, nXr, &dXr)) || (rc = up(U_ref, nUr, &dUr)) || (rc = up(R, nR, &dR)) ||
      (rc = up(Q, nQ, &dQ)))
    return rc;
  auto nanflag = [&](const double* d, size_t cnt, int k) {
    if (d) { nan_flag_kernel<<<std::min(nblk(cnt), 1184), 256, 0, h->stream>>>(d, cnt, h->flags + k); ++g_launches; }
  };
  nanflag(dlx, nLx, FL_XBOX_ABSENT); nanflag(dux, nUx, FL_XBOX_ABSENT);
  nanflag(dlu, nLu, FL_UBOX_ABSENT); nanflag(duu, nUu, FL_UBOX_ABSENT);
  nanflag(ds, Pt, FL_SLEW_ABSENT);
  nanflag(ds0, Pt, FL_SLEW0_ABSENT); nanflag(dum1, Pt * m, FL_SLEW0_ABSENT);
  if (ds) { nonzero_flag_kernel<<<std::min(nblk(Pt), 1184), 256, 0, h->stream>>>(ds, Pt, h->flags + FL_SLEW_NONZERO); ++g_launches; }
  pack_cost_kernel<<<nblk(Pt * N), 256, 0, h->stream>>>(pb, NX, NU, dQ, dR, dXr, dUr, ds, ds0, dum1, (int)Pt,
                                                         h->Nc_orig >= 1 ? 1 : 0, h->row_major, h->flags, h->bc_all, h->bc_prob);
  ++g_launches;
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  pack_particle_kernel<<<nblk(Pt), 256, 0, h->stream>>>(dx0, pb.x0, (int)Pt, pb.P, L, n, NX, 0.0);
  ++g_launches;
  const double inf = INFINITY;
  const int nb = nblk(Pt * N);
  pack_vec_kernel<<<nb, 256, 0, h->stream>>>(dlu, pb.lu, (int)Pt, L, m, NU, -inf, h->flags + FL_UBOX_ABSENT, bmode(BC_LU));
  pack_vec_kernel<<<nb, 256, 0, h->stream>>>(duu, pb.uu, (int)Pt, L, m, NU, inf, h->flags + FL_UBOX_ABSENT, bmode(BC_UU));
  pack_vec_kernel<<<nb, 256, 0, h->stream>>>(dlx, pb.lx, (int)Pt, L, n, NX, -inf, h->flags + FL_XBOX_ABSENT, bmode(BC_LX));
  pack_vec_kernel<<<nb, 256, 0, h->stream>>>(dux, pb.ux, (int)Pt, L, n, NX, inf, h->flags + FL_XBOX_ABSENT, bmode(BC_UX));
  g_launches += 4;
  fill_int_kernel<<<nblk(pb.B), 256, 0, h->stream>>>(h->uniform, pb.B, 1);
  ++g_launches;
  uniform_check_kernel<<<nblk(Pt * N), 256, 0, h->stream>>>(pb, NX, NU, h->uniform, (int)Pt);
  ++g_launches;
  CU(cudaGetLastError());
  int fl[FL_COUNT];
  CU(cudaMemcpyAsync(fl, h->flags, sizeof(fl), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  pb.has_xbox = !fl[FL_XBOX_ABSENT];
  pb.has_ubox = !fl[FL_UBOX_ABSENT];
  pb.has_slew = (!fl[FL_SLEW_ABSENT] && fl[FL_SLEW_NONZERO]) ? 1 : 0;
  h->have_problem = true;
  return pmpc_set_warm_start(h, nullptr, nullptr);
}

int pmpc_set_warm_start(pmpc_handle_t h, const double* X_prev, const double* U_prev) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  DevProb& pb = h->pb;
  int rc;
  auto bmode = [&](int k) { return ((h->bc_all >> k) & 1u) ? 2 : (((h->bc_prob >> k) & 1u) ? 1 : 0); };
  if ((rc = pack_vec(h, X_prev, pb.Xp, pb.n, h->shape.NX, 0.0, 0, bmode(BC_XPREV)))) return rc;
  if ((rc = pack_vec(h, U_prev, pb.Up, pb.m, h->shape.NU, 0.0, h->Ptot * pb.N * pb.n, bmode(BC_UPREV)))) return rc;
  // reset ADMM / active-set state
  const size_t P = pb.P, N = pb.N, NX = h->shape.NX, NU = h->shape.NU;
  CU(cudaMemsetAsync(pb.wu, 0, N * NU * P * sizeof(double), h->stream));
  CU(cudaMemsetAsync(pb.yu, 0, N * NU * P * sizeof(double), h->stream));
  CU(cudaMemsetAsync(pb.wx, 0, N * NX * P * sizeof(double), h->stream));
  CU(cudaMemsetAsync(pb.yx, 0, N * NX * P * sizeof(double), h->stream));
  CU(cudaMemsetAsync(pb.xany, 0, P, h->stream));
  CU(cudaMemsetAsync(pb.act, 0, N * NU * P, h->stream));
  CU(cudaMemsetAsync(pb.actn, 0, N * NU * P, h->stream));
  CU(cudaMemsetAsync(pb.actc, 0, (size_t)pb.B * pb.D + 32, h->stream));
  CU(cudaMemsetAsync(pb.actcn, 0, (size_t)pb.B * pb.D + 32, h->stream));
  fill_kernel<<<nblk(pb.B), 256, 0, h->stream>>>(pb.rho, pb.B, pb.rho0);
  ++g_launches;
  CU(cudaGetLastError());
  if (X_prev || U_prev) CU(cudaStreamSynchronize(h->stream));  // the caller's buffers are not retained
  return PMPC_OK;
}

int pmpc_set_dynamics(pmpc_handle_t h, int kind, const double* params, size_t nparam, double eps) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  DevProb& pb = h->pb;
  if (kind == PMPC_DYN_DUBINS) {
    if (pb.n != 4 || pb.m != 2 || nparam != 3 || !params)
      return fail(PMPC_ERR_BAD_ARG, "Dubins car needs xdim=4, udim=2, params (3, B*M)");
    double* d;
    int rc;
    if ((rc = upload(h, params, h->Ptot * 3, 0, &d))) return rc;
    pack_particle_kernel<<<nblk(h->Ptot), 256, 0, h->stream>>>(d, pb.dynp, (int)h->Ptot, pb.P, Lay{pb.M, pb.Mpad, pb.N, pb.tile}, 3, 4, 0.0);
    ++g_launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    pb.dyn_kind = DYN_DUBINS;
    pb.dyn_eps = eps;
    h->have_dyn = true;
    userdyn_free(h->user);
    h->user = nullptr;
    return PMPC_OK;
  }
  return fail(PMPC_ERR_BAD_ARG, "unknown dynamics kind (use pmpc_set_linearization / pmpc_set_ltv for AFFINE)");
}

static int set_affine(pmpc_handle_t h, const double* f, const double* fx, const double* fu, int c_is_f) {
  if (!h || !fx || !fu) return fail(PMPC_ERR_BAD_ARG, "fx, fu required");
  if (!h->have_problem) return fail(PMPC_ERR_STATE, "pmpc_set_problem first (the linearisation is packed against x0, X_prev, U_prev)");
  CU(cudaSetDevice(h->device));
  DevProb& pb = h->pb;
  const size_t Pt = h->Ptot, N = pb.N, n = pb.n, m = pb.m;
  double *dfx, *dfu, *df = nullptr;
  int rc;
  if ((rc = upload(h, fx, Pt * N * n * n, 0, &dfx))) return rc;
  // f and fu go through the persistent lin_* buffers (sized exactly)
  CU(cudaMemcpyAsync(h->lin_fu, fu, Pt * N * n * m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  dfu = h->lin_fu;
  if (f) {
    CU(cudaMemcpyAsync(h->lin_f, f, Pt * N * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    df = h->lin_f;
  }
  pack_lin_kernel<<<nblk(Pt * N), 256, 0, h->stream>>>(pb, h->shape.NX, h->shape.NU, df, dfx, dfu, (int)Pt, c_is_f,
                                                        h->row_major);
  ++g_launches;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  pb.dyn_kind = DYN_AFFINE;
  h->have_dyn = true;
  userdyn_free(h->user);
  h->user = nullptr;
  return PMPC_OK;
}

int pmpc_set_linearization(pmpc_handle_t h, const double* f, const double* fx, const double* fu) {
  if (!f) return fail(PMPC_ERR_BAD_ARG, "f required");
  return set_affine(h, f, fx, fu, 0);
}

int pmpc_set_ltv(pmpc_handle_t h, const double* A, const double* Bm, const double* c) {
  return set_affine(h, c, A, Bm, 1);
}

static int userdyn_launch_lin(pmpc_handle_t h, bool want_ref_layout, const int* done);
static int userdyn_scp(pmpc_handle_t h, int max_it, double res_tol);

int pmpc_linearize(pmpc_handle_t h) {
  if (!h || !h->have_problem || !h->have_dyn) return fail(PMPC_ERR_STATE, "set problem and dynamics first");
  if (h->user) {
    CU(cudaSetDevice(h->device));
    return userdyn_launch_lin(h, true, nullptr);
  }
  if (h->pb.dyn_kind != DYN_DUBINS) return fail(PMPC_ERR_STATE, "pmpc_linearize needs built-in or registered dynamics");
  CU(cudaSetDevice(h->device));
  linearize_dubins_kernel<<<nblk(h->Ptot * h->pb.N), 256, 0, h->stream>>>(h->pb, h->lin_f, h->lin_fx, h->lin_fu,
                                                                           (int)h->Ptot, h->row_major);
  ++g_launches;
  CU(cudaGetLastError());
  return PMPC_OK;
}

int pmpc_get_linearization(pmpc_handle_t h, double* f, double* fx, double* fu) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  const size_t Pt = h->Ptot, N = h->pb.N, n = h->pb.n, m = h->pb.m;
  if (f) CU(cudaMemcpyAsync(f, h->lin_f, Pt * N * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (fx) CU(cudaMemcpyAsync(fx, h->lin_fx, Pt * N * n * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (fu) CU(cudaMemcpyAsync(fu, h->lin_fu, Pt * N * n * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return PMPC_OK;
}

static int launch(pmpc_handle_t h, int* occ_out = nullptr) {
  const DevProb& pb = h->pb;
  // CTA kind: one thread per single-particle problem, or one CTA per problem (one thread per particle)
  const int threads = (pb.M == 1) ? 128 : ((pb.M + 31) / 32) * 32;
  const int kind = (pb.M == 1) ? 0 : (threads <= 128 ? 1 : 2);
  const int blocks = (pb.M == 1) ? (pb.B + threads - 1) / threads : pb.B;
  // fused-path variant: bit 0 = on-the-fly Dubins linearisation, bit 1 = lean (no slew penalty)
  const int NX = h->shape.NX, NU = h->shape.NU;
  int var = 0;
  if (NX == 4 && NU == 2 && pb.dyn_kind == DYN_DUBINS) var |= 1;
  if (NX <= 4 && !pb.has_slew) var |= 2;
  launch_fn fn = nullptr, fn_half = nullptr;
  for (const VariantEntry& v : kVariants) {
    if (v.nx == NX && v.nu == NU && v.kind == kind && v.var == var) fn = v.fn;
    if (v.nx == NX && v.nu == NU && v.kind == 3 && v.var == var) fn_half = v.fn;
  }
  if (!fn) return fail(PMPC_ERR_UNSUPPORTED, "no kernel for this shape");
  // two problems per warp: M <= 16 particles in the 16-slot layout, fused active-set passes only (no state box, one
  // GPU); what it cannot finish is continued by the one-problem-per-warp kernel launched behind it in resume mode
  // It pays once the one-problem-per-warp kernel needs more than one wave (8 warps per SM): below that it only doubles
  // the latency of a warp.  PMPC_HALF=0 / 1 overrides the batch-size rule (tests, A/B timing).
  const char* half_env = getenv("PMPC_HALF");
  const bool half_wanted = half_env ? atoi(half_env) != 0 : pb.B > 8 * h->num_sms;
  const bool half = fn_half && pb.M >= 2 && pb.M <= 16 && pb.Mpad == 16 && pb.fused && pb.polish && !pb.has_xbox &&
                    pb.D <= DSMALL && pb.world <= 1 && half_wanted;  // (D <= DSMALL: shared block solved in registers)
  if (occ_out) {
    cudaError_t e = half ? fn_half(pb, (pb.B + 1) / 2, 32, h->stream, occ_out) : fn(pb, blocks, threads, h->stream, occ_out);
    if (half) *occ_out *= 2;  // problems per SM
    return e == cudaSuccess ? PMPC_OK : fail(PMPC_ERR_CUDA, std::string("occupancy query: ") + cudaGetErrorString(e));
  }
  if (pb.world > 1 && kind == 0) return fail(PMPC_ERR_UNSUPPORTED, "particle sharding needs at least 2 particles per rank");
  auto next_window = [&]() {  // every launch gets its own window of exchange sequence numbers (flags only ever grow)
    if (h->pb.world > 1) {
      h->seq += (1ull << 32);
      h->pb.seq0 = h->seq;
    }
  };
  // NX >= 8: the warp-per-particle kernel runs the SCP loop (wpp.cuh); what it hands over (state bound active,
  // active-set iteration stalls) is continued by the thread-per-particle kernel launched behind it in resume mode
  const WppEntry* we = nullptr;
  if (NX >= 8 && pb.fused && pb.polish && pb.dyn_kind == DYN_AFFINE && ((pb.M == 1) ? 0 : pb.D) <= kWppDmax)
    for (const WppEntry& w : kWpp)
      if (w.nx == NX && w.nu == NU && w.slew == (pb.has_slew ? 1 : 0)) we = &w;
  CU(cudaEventRecord(h->ev0, h->stream));
  h->pb.resume = 0;
  {
    const char* mu = getenv("PMPC_AL_MU");  // multiplier method for active state bounds (wpp.cuh): starting penalty; 0 = ADMM hand-over
    h->pb.al_mu = mu ? atof(mu) : 30.0;
  }
  if (we) {
    WppPlan pl;
    cudaError_t e = we->plan(pb, &pl);
    if (e != cudaSuccess) return fail(PMPC_ERR_CUDA, std::string("warp-per-particle plan: ") + cudaGetErrorString(e));
    if (pl.tm_doubles > h->tm_doubles || pl.wscr_doubles > h->wscr_doubles) {
      CU(cudaStreamSynchronize(h->stream));
      if (pl.tm_doubles > h->tm_doubles) {
        cudaFree(h->tm);
        h->tm = nullptr;
        h->tm_doubles = 0;
        CU(cudaMalloc(&h->tm, pl.tm_doubles * sizeof(double)));
        h->tm_doubles = pl.tm_doubles;
        h->tm_Tc = -1;
      }
      if (pl.wscr_doubles > h->wscr_doubles) {
        cudaFree(h->wscr);
        h->wscr = nullptr;
        h->wscr_doubles = 0;
        CU(cudaMalloc(&h->wscr, pl.wscr_doubles * sizeof(double)));
        h->wscr_doubles = pl.wscr_doubles;
      }
    }
    if (h->tm_Tc != pl.Tc || h->tm_cap != pl.cap) {  // the flag words move with the team layout: start from zero
      CU(cudaMemsetAsync(h->tm, 0, h->tm_doubles * sizeof(double), h->stream));
      h->tm_Tc = pl.Tc;
      h->tm_cap = pl.cap;
      h->tseq = 0;
    }
    h->tseq += (1ull << 32);
    h->pb.tm = h->tm;
    h->pb.wscr = h->wscr;
    h->pb.tm_cap = pl.cap;
    h->pb.tseq0 = h->tseq;
    next_window();
    e = we->launch(h->pb, pl, h->stream);
    if (e != cudaSuccess) return fail(PMPC_ERR_CUDA, std::string("warp-per-particle kernel launch: ") + cudaGetErrorString(e));
    ++g_launches;
    h->pb.resume = 1;
  }
  if (half) {
    const char* hi = getenv("PMPC_HALF_IPM");
    h->pb.half_ipm = hi ? atoi(hi) : 1;
    cudaError_t e = fn_half(h->pb, (pb.B + 1) / 2, 32, h->stream, nullptr);
    if (e != cudaSuccess) return fail(PMPC_ERR_CUDA, std::string("kernel launch (two problems per warp): ") + cudaGetErrorString(e));
    ++g_launches;
    h->pb.resume = 1;
  }
  next_window();
  cudaError_t e = fn(h->pb, blocks, threads, h->stream, nullptr);
  h->pb.resume = 0;
  if (e != cudaSuccess) return fail(PMPC_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  ++g_launches;
  CU(cudaEventRecord(h->ev1, h->stream));
  h->solved = true;
  return PMPC_OK;
}

int pmpc_scp_solve_async(pmpc_handle_t h, int max_it, double res_tol) {
  if (!h || !h->have_problem || !h->have_dyn) return fail(PMPC_ERR_STATE, "set problem and dynamics first");
  if (max_it < 1) return fail(PMPC_ERR_BAD_ARG, "max_it < 1");
  CU(cudaSetDevice(h->device));
  h->user_timed = false;
  if (max_it > h->pb.hist_len) {  // the history holds every iteration of the launch (the reference keeps them all)
    CU(cudaStreamSynchronize(h->stream));
    double* nh = nullptr;
    const size_t cnt = (size_t)4 * max_it * h->pb.B;
    CU(cudaMalloc(&nh, cnt * sizeof(double)));
    cudaFree(h->hist);
    h->hist = nh;
    h->pb.hist = nh;
    h->pb.hist_len = max_it;
    CU(cudaMemsetAsync(nh, 0, cnt * sizeof(double), h->stream));
  }
  if (h->user) return userdyn_scp(h, max_it, res_tol);  // host-driven loop: synchronous
  h->pb.max_it = max_it;
  h->pb.res_tol = res_tol;
  h->pb.keep_prev = 0;
  return launch(h);
}

int pmpc_sync(pmpc_handle_t h) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  if (h->solved && !h->user_timed) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  if (h->pb.world > 1) {
    int xe = 0;
    CU(cudaMemcpy(&xe, h->xerr, sizeof(int), cudaMemcpyDeviceToHost));
    if (xe) {
      CU(cudaMemset(h->xerr, 0, sizeof(int)));
      return fail(PMPC_ERR_STATE, "particle-sharded exchange timed out: a peer rank never arrived (failed rank, or the "
                                  "kernels of the ranks were not co-resident)");
    }
  }
  return PMPC_OK;
}

int pmpc_scp_solve(pmpc_handle_t h, int max_it, double res_tol) {
  int rc = pmpc_scp_solve_async(h, max_it, res_tol);
  if (rc) return rc;
  return pmpc_sync(h);
}

int pmpc_qp_solve(pmpc_handle_t h, int keep_prev) {
  if (!h || !h->have_problem || !h->have_dyn) return fail(PMPC_ERR_STATE, "set problem and linearisation first");
  CU(cudaSetDevice(h->device));
  const int kind = h->pb.dyn_kind;
  // one sub-problem at the *given* linearisation: for built-in dynamics linearize_prep recomputes it
  // at (X_prev, U_prev), which is the same point.
  h->pb.max_it = 1;
  h->pb.res_tol = -1.0;
  h->pb.keep_prev = keep_prev;
  h->user_timed = false;
  if (h->user) {  // registered dynamics: the sub-problem is the one at the handle's (X_prev, U_prev)
    int rl = userdyn_launch_lin(h, false, nullptr);
    if (rl) return rl;
  }
  int rc = launch(h);
  h->pb.dyn_kind = kind;
  if (rc) return rc;
  return pmpc_sync(h);
}

static int get_solution_impl(pmpc_handle_t h, double* X, double* U, int lead) {
  if (!h || !h->solved) return fail(PMPC_ERR_STATE, "no solution yet");
  CU(cudaSetDevice(h->device));
  const DevProb& pb = h->pb;
  const size_t Pt = h->Ptot, N = pb.N;
  const size_t nXo = Pt * (N + lead) * pb.n, nUo = Pt * N * pb.m;
  if (nXo + nUo > h->stage_doubles) return fail(PMPC_ERR_STATE, "staging too small");
  double* dX = h->stage;
  double* dU = h->stage + nXo;
  const Lay L{pb.M, pb.Mpad, pb.N, pb.tile};
  if (X) {
    unpack_vec_kernel<<<nblk(Pt * N), 256, 0, h->stream>>>(pb.X, dX, (int)Pt, L, h->shape.NX, pb.n, lead, pb.x0, pb.P);
    ++g_launches;
    CU(cudaMemcpyAsync(X, dX, nXo * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (U) {
    unpack_vec_kernel<<<nblk(Pt * N), 256, 0, h->stream>>>(pb.U, dU, (int)Pt, L, h->shape.NU, pb.m, 0, nullptr, pb.P);
    ++g_launches;
    CU(cudaMemcpyAsync(U, dU, nUo * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  return PMPC_OK;
}

int pmpc_get_solution(pmpc_handle_t h, double* X, double* U) { return get_solution_impl(h, X, U, 0); }
int pmpc_get_trajectory(pmpc_handle_t h, double* X, double* U) { return get_solution_impl(h, X, U, 1); }

int pmpc_get_info(pmpc_handle_t h, long long* info) {
  if (!h || !info) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  const size_t B = h->pb.B;
  std::vector<long long> tmp(8 * B);
  CU(cudaMemcpyAsync(tmp.data(), h->pb.info, 8 * B * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (size_t b = 0; b < B; ++b)
    for (int k = 0; k < 8; ++k) info[b * 8 + k] = tmp[(size_t)k * B + b];
  return PMPC_OK;
}

int pmpc_get_hist(pmpc_handle_t h, double* hist, int max_it) {
  if (!h || !hist) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  const size_t B = h->pb.B, L = h->pb.hist_len;
  std::vector<double> tmp(4 * L * B);
  CU(cudaMemcpyAsync(tmp.data(), h->pb.hist, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 4; ++k)
    for (int it = 0; it < max_it; ++it)
      for (size_t b = 0; b < B; ++b)
        hist[((size_t)k * max_it + it) * B + b] = (it < (int)L) ? tmp[((size_t)k * L + it) * B + b] : NAN;
  return PMPC_OK;
}

int pmpc_last_kernel_ms(pmpc_handle_t h, float* ms) {
  if (!h || !ms) return fail(PMPC_ERR_BAD_ARG, "null");
  if (h->solved && cudaEventQuery(h->ev1) == cudaSuccess) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  *ms = h->last_ms;
  return PMPC_OK;
}

// ---- particle sharding over the GPUs of one node -------------------------------------------------
int pmpc_comm_export(pmpc_handle_t h, void* ipc_handle_64) {
  if (!h || !ipc_handle_64) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!h->comm) {
    const DevProb& pb = h->pb;
    const size_t D = (size_t)pb.D;
    const size_t cap = std::max<size_t>(D * D + D, (size_t)ST_COUNT) + 32;
    h->comm_bytes = (size_t)pb.B * 2 * (cap + 1) * sizeof(double);
    CU(cudaMalloc(&h->comm, h->comm_bytes));
    CU(cudaMemset(h->comm, 0, h->comm_bytes));
    h->pb.comm_cap = (int)cap;
  }
  cudaIpcMemHandle_t ih;
  CU(cudaIpcGetMemHandle(&ih, h->comm));
  memcpy(ipc_handle_64, &ih, 64);
  return PMPC_OK;
}

int pmpc_comm_connect(pmpc_handle_t h, int rank, int world, const void* ipc_handles, size_t M_total) {
  if (!h || !ipc_handles || world < 1 || world > 8 || rank < 0 || rank >= world)
    return fail(PMPC_ERR_BAD_ARG, "rank/world (world <= 8) or handles");
  if (!h->comm) return fail(PMPC_ERR_STATE, "pmpc_comm_export first");
  CU(cudaSetDevice(h->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->pb.peer[r] = h->comm;
      continue;
    }
    cudaIpcMemHandle_t ih;
    memcpy(&ih, (const char*)ipc_handles + (size_t)r * 64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    h->peer_open[r] = p;
    h->pb.peer[r] = (double*)p;
  }
  h->pb.rank = rank;
  h->pb.world = world;
  h->pb.M_total = (int)M_total;
  return PMPC_OK;
}

int pmpc_comm_connect_local(pmpc_handle_t h, int rank, int world, const pmpc_handle_t* handles, size_t M_total) {
  if (!h || !handles || world < 1 || world > 8 || rank < 0 || rank >= world || handles[rank] != h)
    return fail(PMPC_ERR_BAD_ARG, "rank/world (world <= 8) or handles");
  for (int r = 0; r < world; ++r) {
    if (!handles[r] || !handles[r]->comm) return fail(PMPC_ERR_STATE, "pmpc_comm_export on every handle first");
    if (handles[r]->comm_bytes != h->comm_bytes) return fail(PMPC_ERR_BAD_ARG, "handles differ in B / Nc");
    if (handles[r]->device != h->device) {
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, h->device, handles[r]->device));
      if (!can) return fail(PMPC_ERR_UNSUPPORTED, "no peer access between the devices of the handles");
      CU(cudaSetDevice(h->device));
      cudaError_t e = cudaDeviceEnablePeerAccess(handles[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(PMPC_ERR_CUDA, cudaGetErrorString(e));
      cudaGetLastError();
    }
    h->pb.peer[r] = handles[r]->comm;
  }
  h->pb.rank = rank;
  h->pb.world = world;
  h->pb.M_total = (int)M_total;
  return PMPC_OK;
}

int pmpc_comm_set_timeout(pmpc_handle_t h, double seconds) {
  if (!h || !(seconds > 0.0)) return fail(PMPC_ERR_BAD_ARG, "null handle or non-positive timeout");
  h->pb.xtimeout_ns = (unsigned long long)(seconds * 1e9);
  return PMPC_OK;
}

long long pmpc_launch_count(void) { return g_launches.load(); }

#include "userdyn.inl"

int pmpc_get_occupancy(pmpc_handle_t h, int* ctas_per_sm, int* num_sms) {
  if (!h || !ctas_per_sm || !num_sms) return fail(PMPC_ERR_BAD_ARG, "null");
  if (!h->have_problem || !h->have_dyn) return fail(PMPC_ERR_STATE, "set problem and dynamics first (they select the kernel)");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, h->device));
  return launch(h, ctas_per_sm);
}

int pmpc_timer_start(pmpc_handle_t h) {
  if (!h) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->tev0, h->stream));
  return PMPC_OK;
}

int pmpc_timer_stop(pmpc_handle_t h, float* ms) {
  if (!h || !ms) return fail(PMPC_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->tev1, h->stream));
  CU(cudaEventSynchronize(h->tev1));
  CU(cudaEventElapsedTime(ms, h->tev0, h->tev1));
  return PMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// compat entry points.  Handles are cached per shape so repeated calls (one per SCP iteration from
// the Python loop) do not pay for cudaMalloc.
static std::mutex g_cache_mu;
static std::map<std::tuple<int, size_t, size_t, size_t, size_t, long long>, pmpc_handle_t> g_cache;

int pmpc_lqp_solve(double* X_out, double* U_out, size_t xdim, size_t udim, size_t N, size_t M, long long Nc,
                   const double* x0, const double* f, const double* fx, const double* fu, const double* X_prev,
                   const double* U_prev, const double* Q, const double* R, const double* X_ref, const double* U_ref,
                   const double* lx, const double* ux, const double* lu, const double* uu, double reg_x,
                   double reg_u, const double* slew_reg, const double* slew_reg0, const double* slew_um1,
                   const pmpc_settings* settings, long long* info8) {
  if (!X_out || !U_out) return fail(PMPC_ERR_BAD_ARG, "null outputs");
  std::lock_guard<std::mutex> lock(g_cache_mu);
  if (pmpc_device_count() < 1) return fail(PMPC_ERR_NO_DEVICE, "no CUDA device: pmpc_b200 has no CPU fallback");
  const long long Nce = (Nc < 0 || Nc > (long long)N) ? (long long)N : Nc;
  // handles live on the device that was current when they were made: the caller's current device is part of the key
  // and is current again when the call returns (the handle API switches to its own device internally)
  int dev = 0;
  CU(cudaGetDevice(&dev));
  struct Restore {
    int d;
    ~Restore() { cudaSetDevice(d); }
  } restore{dev};
  auto key = std::make_tuple(dev, xdim, udim, N, M, Nce);
  pmpc_handle_t h = nullptr;
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    int rc = pmpc_create(&h, -1, 1, M, N, xdim, udim, Nce);
    if (rc) return rc;
    if (g_cache.size() >= 8) {
      for (auto& kv : g_cache) pmpc_destroy(kv.second);
      g_cache.clear();
    }
    g_cache[key] = h;
  } else {
    h = it->second;
  }
  int rc;
  pmpc_settings s;
  if (settings) s = *settings; else pmpc_default_settings(&s);
  if ((rc = pmpc_set_settings(h, &s))) return rc;
  if ((rc = pmpc_set_problem(h, x0, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u, slew_reg, slew_reg0, slew_um1)))
    return rc;
  if ((rc = pmpc_set_warm_start(h, X_prev, U_prev))) return rc;
  if ((rc = pmpc_set_linearization(h, f, fx, fu))) return rc;
  if ((rc = pmpc_qp_solve(h, 1))) return rc;
  if ((rc = pmpc_get_solution(h, X_out, U_out))) return rc;
  if (info8 && (rc = pmpc_get_info(h, info8))) return rc;
  return PMPC_OK;
}

// The reference's pybind shim starts the embedded Julia runtime at import (module.cpp:122-125, julia_init.h:1-2).
// Exported as no-ops so that the UNMODIFIED module.cpp links against this library: there is no runtime to start.
void init_julia(int argc, char* argv[]) { (void)argc; (void)argv; }
void shutdown_julia(int retcode) { (void)retcode; }

static void nan_fill(double* p, size_t n) {
  if (p) for (size_t i = 0; i < n; ++i) p[i] = NAN;
}

void c_lqp_solve(double* X_out, double* U_out, size_t xdim, size_t udim, size_t N, size_t M, long long Nc, double* x0,
                 double* f, double* fx, double* fu, double* X_prev, double* U_prev, double* Q, double* R,
                 double* X_ref, double* U_ref, double* lx, double* ux, double* lu, double* uu, double reg_x,
                 double reg_u, double* slew_reg, double* slew_reg0, double* slew_um1, long long verbose) {
  long long info[8] = {0};
  int rc = pmpc_lqp_solve(X_out, U_out, xdim, udim, N, M, Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx,
                          ux, lu, uu, reg_x, reg_u, slew_reg, slew_reg0, slew_um1, nullptr, info);
  if (rc != PMPC_OK) {
    if (verbose) fprintf(stderr, "pmpc_b200: c_lqp_solve failed: %s\n", pmpc_last_error());
    nan_fill(X_out, xdim * N * M);
    nan_fill(U_out, udim * N * M);
  } else if (verbose) {
    fprintf(stderr, "pmpc_b200: admm_iters=%lld factorizations=%lld polish_passes=%lld status=%lld\n", info[1],
            info[2], info[3], info[4]);
  }
}

void c_lcone_solve(double* X_out, double* U_out, size_t xdim, size_t udim, size_t N, size_t M, long long Nc,
                   double* x0, double* f, double* fx, double* fu, double* X_prev, double* U_prev, double* Q,
                   double* R, double* X_ref, double* U_ref, double* lx, double* ux, double* lu, double* uu,
                   double reg_x, double reg_u, double* slew_reg, double* slew_reg0, double* slew_um1,
                   long long verbose, double smooth_alpha, char* solver) {
  (void)solver;
  // The epigraph objective of the cone route (main.jl:204-227) has the QP's minimiser only for
  // M == 1 with hard constraints (smooth_alpha NaN); anything else is out of scope -> NaN.
  if (M == 1 && smooth_alpha != smooth_alpha) {
    c_lqp_solve(X_out, U_out, xdim, udim, N, M, Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu,
                reg_x, reg_u, slew_reg, slew_reg0, slew_um1, verbose);
    return;
  }
  if (verbose) fprintf(stderr, "pmpc_b200: c_lcone_solve: M>1 / smoothed cone route is not provided\n");
  nan_fill(X_out, xdim * N * M);
  nan_fill(U_out, udim * N * M);
}

}  // extern "C"
