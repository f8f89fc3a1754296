// User device dynamics (textually included by api.cu inside `extern "C"`).
//
// BASELINE.json north_star: "plus a C-ABI hook for registering user device dynamics".  In the reference the
// dynamics are an arbitrary Python callable `f_fx_fu_fn(X_, U_prev) -> f, fx, fu` evaluated on the host once
// per SCP iteration (pmpc/scp_mpc.py:338-342; e.g. tests/dubins_car.py:48-90 through torch autograd).  Here
// the user hands over CUDA C++ source of ONE __device__ function; it is compiled at run time with NVRTC for
// sm_100a into a linearisation kernel (K1) that reads the handle's (x0, X_prev, U_prev) and writes A, B, c
// straight into the warp-tiled HBM layout the solve kernel streams — f, fx, fu never cross PCIe and never
// exist in the reference layout unless pmpc_get_linearization asks for them.
//
// The SCP loop is then driven from the host, one iteration = [user linearisation kernel] + [solve kernel with
// max_it = 1] + [two bookkeeping kernels], with ONE 4-byte read-back per iteration (the number of problems
// still iterating).  Problems of a batch that have met the stop rule are frozen (their X, U, counters stop
// changing) so that batched == individual solves, as on the built-in path.


namespace {

// ---- NVRTC, bound lazily (the library must load on machines without it) -------------------------------
struct Nvrtc {
  void* so = nullptr;
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static Nvrtc* nvrtc() {
  static Nvrtc rt;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names)
      if ((rt.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!rt.so) return;
#define PMPC_SYM(field, sym) *(void**)(&rt.field) = dlsym(rt.so, sym)
    PMPC_SYM(CreateProgram, "nvrtcCreateProgram");
    PMPC_SYM(CompileProgram, "nvrtcCompileProgram");
    PMPC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    PMPC_SYM(GetProgramLog, "nvrtcGetProgramLog");
    PMPC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    PMPC_SYM(GetCUBIN, "nvrtcGetCUBIN");
    PMPC_SYM(DestroyProgram, "nvrtcDestroyProgram");
    PMPC_SYM(GetErrorString, "nvrtcGetErrorString");
#undef PMPC_SYM
    if (!rt.CreateProgram || !rt.CompileProgram || !rt.GetCUBIN || !rt.GetCUBINSize || !rt.DestroyProgram) {
      dlclose(rt.so);
      rt.so = nullptr;
    }
  });
  return rt.so ? &rt : nullptr;
}

// The wrapper kernel around the user's function.  One thread per (particle slot, stage), slot fastest, so
// that every tile access is a coalesced 256-byte row (solver.cuh "DATA LAYOUT").  Same arithmetic as
// pack_lin_kernel: A_j = fx_j (A_0 = 0), B_j = fu_j, c_j = f_j - fu_j ubar_j - fx_j xbar_{j-1}
// (lqp_utils.jl:231-300; no fx term at stage 0: linearised at x0 itself, lqp_utils.jl:288-296).
static const char* kUserWrapper = R"SRC(
// trampoline with prefixed parameter names: the user's function may be called `f`, `x`, `c`, ... without being shadowed
// by the wrapper kernel's own identifiers
__device__ __forceinline__ void pmpc_call_user_(const double* pmpc_a0_, const double* pmpc_a1_, const double* pmpc_a2_,
                                                double* pmpc_a3_, double* pmpc_a4_, double* pmpc_a5_) {
  PMPC_USER_FN(pmpc_a0_, pmpc_a1_, pmpc_a2_, pmpc_a3_, pmpc_a4_, pmpc_a5_);
}
extern "C" __global__ void pmpc_user_linearize(
    const double* __restrict__ x0, const double* __restrict__ Xp, const double* __restrict__ Up,
    const double* __restrict__ prm, double* __restrict__ A, double* __restrict__ Bm, double* __restrict__ c,
    double* __restrict__ f_out, double* __restrict__ fx_out, double* __restrict__ fu_out,
    const int* __restrict__ done, int B, int M, int Mpad, int N, unsigned long long P, int row_major) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long sl = idx % P;
  const int j = (int)(idx / P);
  if (j >= N || sl >= (unsigned long long)B * Mpad || (int)(sl % Mpad) >= M) return;
  const int b = (int)(sl / Mpad);
  if (done != nullptr && done[b]) return;
  const unsigned long long tile = (sl / PMPC_TILE) * (unsigned long long)N;
  const int lane = (int)(sl % PMPC_TILE);
#define PMPC_TIX(jj, E, e) ((((tile + (jj)) * (E)) + (e)) * PMPC_TILE + lane)
  double x[PMPC_N], u[PMPC_M], p[PMPC_NP > 0 ? PMPC_NP : 1], f[PMPC_N], fx[PMPC_N * PMPC_N], fu[PMPC_N * PMPC_M];
  for (int e = 0; e < PMPC_N; ++e) x[e] = (j == 0) ? x0[(unsigned long long)e * P + sl] : Xp[PMPC_TIX(j - 1, PMPC_NX, e)];
  for (int e = 0; e < PMPC_M; ++e) u[e] = Up[PMPC_TIX(j, PMPC_NU, e)];
  for (int e = 0; e < PMPC_NP; ++e) p[e] = prm[(unsigned long long)e * P + sl];
  for (int e = 0; e < PMPC_N; ++e) f[e] = 0.0;
  for (int e = 0; e < PMPC_N * PMPC_N; ++e) fx[e] = 0.0;
  for (int e = 0; e < PMPC_N * PMPC_M; ++e) fu[e] = 0.0;
  pmpc_call_user_(x, u, p, f, fx, fu);
  for (int r = 0; r < PMPC_NX; ++r) {
    double cv = (r < PMPC_N) ? f[r] : 0.0;
    for (int t = 0; t < PMPC_NX; ++t) {
      const double a0 = (r < PMPC_N && t < PMPC_N) ? fx[r * PMPC_N + t] : 0.0;
      const double a = (j >= 1) ? a0 : 0.0;
      A[PMPC_TIX(j, PMPC_NX * PMPC_NX, r * PMPC_NX + t)] = a;
      if (t < PMPC_N) cv -= a * x[t];
    }
    for (int t = 0; t < PMPC_NU; ++t) {
      const double bb = (r < PMPC_N && t < PMPC_M) ? fu[r * PMPC_M + t] : 0.0;
      Bm[PMPC_TIX(j, PMPC_NX * PMPC_NU, r * PMPC_NU + t)] = bb;
      if (t < PMPC_M) cv -= bb * u[t];
    }
    c[PMPC_TIX(j, PMPC_NX, r)] = cv;
  }
  if (f_out != nullptr) {  // reference layout (pmpc_get_linearization): flat particle gf = b*M + i
    const unsigned long long gf = (unsigned long long)b * M + (sl % Mpad);
    const unsigned long long o = gf * N + j;
    for (int r = 0; r < PMPC_N; ++r) {
      f_out[o * PMPC_N + r] = f[r];
      for (int t = 0; t < PMPC_N; ++t)
        fx_out[o * PMPC_N * PMPC_N + (row_major ? r * PMPC_N + t : t * PMPC_N + r)] = fx[r * PMPC_N + t];
      for (int t = 0; t < PMPC_M; ++t)
        fu_out[o * PMPC_N * PMPC_M + (row_major ? r * PMPC_M + t : t * PMPC_N + r)] = fu[r * PMPC_M + t];
    }
  }
}
)SRC";

// after every host-driven iteration: keep the newest iterate of the problems still running
__global__ void ud_snapshot_kernel(DevProb pb, int NX, int NU, const int* __restrict__ done, double* __restrict__ Xf,
                                   double* __restrict__ Uf) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t P = pb.P;
  const size_t sl = idx % P;
  const int j = (int)(idx / P);
  if (j >= pb.N || sl >= (size_t)pb.B * pb.Mpad) return;
  if (done[sl / pb.Mpad]) return;
  const size_t TL = (size_t)pb.tile, tile = (sl / TL) * (size_t)pb.N + j, lane = sl % TL;
  for (int e = 0; e < NX; ++e) Xf[(tile * NX + e) * TL + lane] = pb.X[(tile * NX + e) * TL + lane];
  for (int e = 0; e < NU; ++e) Uf[(tile * NU + e) * TL + lane] = pb.U[(tile * NU + e) * TL + lane];
}

// per problem: accumulate the launch's counters, apply the stop rule (scp_mpc.py:424), count the survivors
__global__ void ud_flag_kernel(DevProb pb, int hrow, double res_tol, int* done, long long* acc, int* active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pb.B || done[b]) return;
  for (int k = 0; k < 8; ++k) {
    const long long v = pb.info[(size_t)k * pb.B + b];
    if (k == 4) acc[(size_t)k * pb.B + b] |= v;
    else acc[(size_t)k * pb.B + b] += v;
  }
  const double res = pb.hist[((size_t)1 * pb.hist_len + hrow) * (size_t)pb.B + b];
  const bool failed = (pb.info[(size_t)4 * pb.B + b] & 1) != 0;
  if (failed || res < res_tol) done[b] = 1;
  else atomicAdd(active, 1);
}

// after the loop: the frozen iterates become the handle's solution and warm start; counters -> info
__global__ void ud_restore_kernel(DevProb pb, int NX, int NU, const double* __restrict__ Xf, const double* __restrict__ Uf) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t P = pb.P;
  const size_t sl = idx % P;
  const int j = (int)(idx / P);
  if (j >= pb.N || sl >= (size_t)pb.B * pb.Mpad) return;
  const size_t TL = (size_t)pb.tile, tile = (sl / TL) * (size_t)pb.N + j, lane = sl % TL;
  for (int e = 0; e < NX; ++e) {
    const double v = Xf[(tile * NX + e) * TL + lane];
    pb.X[(tile * NX + e) * TL + lane] = v;
    pb.Xp[(tile * NX + e) * TL + lane] = v;
  }
  for (int e = 0; e < NU; ++e) {
    const double v = Uf[(tile * NU + e) * TL + lane];
    pb.U[(tile * NU + e) * TL + lane] = v;
    pb.Up[(tile * NU + e) * TL + lane] = v;
  }
}

__global__ void ud_info_kernel(DevProb pb, const long long* __restrict__ acc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)8 * pb.B) pb.info[i] = acc[i];
}

}  // namespace

struct pmpc_userdyn_s {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  double* prm = nullptr;     // (nparam, P) per-particle parameters
  size_t nparam = 0;
  double *Xf = nullptr, *Uf = nullptr;  // frozen iterates (warp-tiled, as pb.X / pb.U)
  int* done = nullptr;       // (B) + [B] = active counter
  long long* acc = nullptr;  // (8, B)
  int* h_active = nullptr;   // page-locked
};

static void userdyn_free(pmpc_userdyn_s* u) {
  if (!u) return;
  if (u->lib) cudaLibraryUnload(u->lib);
  cudaFree(u->prm); cudaFree(u->Xf); cudaFree(u->Uf); cudaFree(u->done); cudaFree(u->acc);
  if (u->h_active) cudaFreeHost(u->h_active);
  delete u;
}

static int userdyn_launch_lin(pmpc_handle_t h, bool want_ref_layout, const int* done) {
  pmpc_userdyn_s* u = h->user;
  const DevProb& pb = h->pb;
  const double *x0 = pb.x0, *Xp = pb.Xp, *Up = pb.Up, *prm = u->prm;
  double *A = pb.A, *Bm = pb.Bm, *c = pb.c;
  double *fo = want_ref_layout ? h->lin_f : nullptr, *fxo = want_ref_layout ? h->lin_fx : nullptr,
         *fuo = want_ref_layout ? h->lin_fu : nullptr;
  int B = pb.B, M = pb.M, Mpad = pb.Mpad, N = pb.N, rm = h->row_major;
  unsigned long long P = (unsigned long long)pb.P;
  void* args[] = {&x0, &Xp, &Up, &prm, &A, &Bm, &c, &fo, &fxo, &fuo, &done, &B, &M, &Mpad, &N, &P, &rm};
  const size_t total = (size_t)pb.P * pb.N;
  CU(cudaLaunchKernel((const void*)u->kern, dim3((unsigned)((total + 127) / 128)), dim3(128), args, 0, h->stream));
  ++g_launches;
  return PMPC_OK;
}

int pmpc_register_dynamics(pmpc_handle_t h, const char* cuda_src, const char* fn_name, const double* params,
                           size_t nparam) {
  if (!h || !cuda_src || !fn_name || !*fn_name) return fail(PMPC_ERR_BAD_ARG, "handle, source and function name required");
  if (nparam > 0 && !params) return fail(PMPC_ERR_BAD_ARG, "params is NULL but nparam > 0");
  if (nparam > 64) return fail(PMPC_ERR_BAD_ARG, "at most 64 dynamics parameters per particle");
  if (h->pb.world > 1) return fail(PMPC_ERR_UNSUPPORTED, "user dynamics on a particle-sharded handle");
  Nvrtc* rt = nvrtc();
  if (!rt) return fail(PMPC_ERR_UNSUPPORTED, "libnvrtc.so.12 not found: user device dynamics need NVRTC");
  CU(cudaSetDevice(h->device));
  const DevProb& pb = h->pb;
  std::string src = "#define PMPC_N " + std::to_string(pb.n) + "\n#define PMPC_M " + std::to_string(pb.m) +
                    "\n#define PMPC_NX " + std::to_string(h->shape.NX) + "\n#define PMPC_NU " + std::to_string(h->shape.NU) +
                    "\n#define PMPC_NP " + std::to_string(nparam) + "\n#define PMPC_TILE " + std::to_string(pb.tile) + "\n#define PMPC_USER_FN " + fn_name + "\n#line 1 \"user_dynamics.cu\"\n" +
                    cuda_src + "\n#line 1 \"pmpc_user_wrapper.cu\"\n" + kUserWrapper;
  void* prog = nullptr;
  int rc = rt->CreateProgram(&prog, src.c_str(), "pmpc_user.cu", 0, nullptr, nullptr);
  if (rc != 0) return fail(PMPC_ERR_CUDA, std::string("nvrtcCreateProgram: ") + (rt->GetErrorString ? rt->GetErrorString(rc) : "?"));
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "--fmad=true"};
  rc = rt->CompileProgram(prog, 4, opts);
  if (rc != 0) {
    std::string log;
    size_t ls = 0;
    if (rt->GetProgramLogSize && rt->GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
      log.resize(ls);
      rt->GetProgramLog(prog, &log[0]);
    }
    rt->DestroyProgram(&prog);
    return fail(PMPC_ERR_BAD_ARG, "user dynamics failed to compile (expected `__device__ void " + std::string(fn_name) +
                                      "(const double* x, const double* u, const double* p, double* f, double* fx, double* fu)`):\n" + log);
  }
  size_t cs = 0;
  rt->GetCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  rt->GetCUBIN(prog, cubin.data());
  rt->DestroyProgram(&prog);
  pmpc_userdyn_s* u = new pmpc_userdyn_s();
  cudaError_t e = cudaLibraryLoadData(&u->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&u->kern, u->lib, "pmpc_user_linearize");
  if (e != cudaSuccess) {
    userdyn_free(u);
    return fail(PMPC_ERR_CUDA, std::string("loading the compiled user dynamics: ") + cudaGetErrorString(e));
  }
  const size_t P = pb.P, N = pb.N, NX = h->shape.NX, NU = h->shape.NU, B = pb.B;
  u->nparam = nparam;
  e = cudaMalloc(&u->prm, std::max<size_t>(nparam, 1) * P * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&u->Xf, N * NX * P * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&u->Uf, N * NU * P * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&u->done, (B + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&u->acc, 8 * B * sizeof(long long));
  if (e == cudaSuccess) e = cudaHostAlloc(&u->h_active, sizeof(int), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    userdyn_free(u);
    return fail(PMPC_ERR_CUDA, std::string("user dynamics buffers: ") + cudaGetErrorString(e));
  }
  if (nparam > 0) {
    double* d;
    int r2;
    if ((r2 = upload(h, params, h->Ptot * nparam, 0, &d))) {
      userdyn_free(u);
      return r2;
    }
    // params arrive (B*M, nparam) row-major == (nparam, B*M) column-major, like pmpc_set_dynamics
    pack_particle_kernel<<<nblk(h->Ptot), 256, 0, h->stream>>>(d, u->prm, (int)h->Ptot, pb.P, Lay{pb.M, pb.Mpad, pb.N},
                                                                (int)nparam, (int)nparam, 0.0);
    ++g_launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
  }
  userdyn_free(h->user);
  h->user = u;
  h->pb.dyn_kind = DYN_AFFINE;
  h->have_dyn = true;
  return PMPC_OK;
}

// host-driven SCP loop for user dynamics (see the header of this file)
static int userdyn_scp(pmpc_handle_t h, int max_it, double res_tol) {
  pmpc_userdyn_s* u = h->user;
  DevProb& pb = h->pb;
  const int B = pb.B, NX = h->shape.NX, NU = h->shape.NU;
  const size_t total = (size_t)pb.P * pb.N;
  CU(cudaMemsetAsync(u->done, 0, (B + 1) * sizeof(int), h->stream));
  CU(cudaMemsetAsync(u->acc, 0, 8 * (size_t)B * sizeof(long long), h->stream));
  cudaEvent_t e0 = h->uev0, e1 = h->uev1;
  CU(cudaEventRecord(e0, h->stream));
  int rc = PMPC_OK;
  for (int it = 0; it < max_it; ++it) {
    const int hrow = std::min(it, pb.hist_len - 1);
    if ((rc = userdyn_launch_lin(h, false, u->done))) return rc;
    pb.max_it = 1;
    pb.res_tol = -1.0;
    pb.keep_prev = 0;
    pb.hist_off = hrow;
    rc = launch(h);
    pb.hist_off = 0;
    if (rc) return rc;
    ud_snapshot_kernel<<<nblk(total, 128), 128, 0, h->stream>>>(pb, NX, NU, u->done, u->Xf, u->Uf);
    CU(cudaMemsetAsync(u->done + B, 0, sizeof(int), h->stream));
    ud_flag_kernel<<<nblk(B, 128), 128, 0, h->stream>>>(pb, hrow, res_tol, u->done, u->acc, u->done + B);
    g_launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(u->h_active, u->done + B, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (*u->h_active == 0) break;
  }
  ud_restore_kernel<<<nblk(total, 128), 128, 0, h->stream>>>(pb, NX, NU, u->Xf, u->Uf);
  ud_info_kernel<<<nblk((size_t)8 * B, 128), 128, 0, h->stream>>>(pb, u->acc);
  g_launches += 2;
  CU(cudaGetLastError());
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->last_ms, e0, e1);
  h->user_timed = true;
  return PMPC_OK;
}
