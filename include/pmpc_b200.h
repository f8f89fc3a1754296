/*
 * pmpc_b200 — C ABI of the B200-native particle-SCP MPC solver.
 *
 * This library is the drop-in for the ONE native boundary of StanfordASL/pmpc's solve path:
 * the `pmpcjl` extension (reference `PMPC.jl/pmpcjl/module.cpp:9-23`) which forwards raw
 * `double*` to the ahead-of-time compiled Julia entry points `c_lqp_solve` / `c_lcone_solve`
 * (reference `PMPC.jl/src/c_interface.jl:77-141`, `:146-214`).  Plain pointers and sizes, no
 * torch / numpy types.  All computation runs in hand-written sm_100a CUDA kernels; there is no
 * CPU fallback: every entry point returns an error (or NaN-fills the outputs for the two
 * void-returning compat symbols, the reference's own failure convention,
 * `PMPC.jl/src/osqp_solver.jl:69-71`) when no CUDA device is usable.
 *
 * ARRAY LAYOUT (identical to the reference C ABI, `c_interface.jl:28-46`): column-major
 *     x0(n,M) f(n,N,M) fx(n,n,N,M) fu(n,m,N,M) X_prev(n,N,M) U_prev(m,N,M) Q(n,n,N,M)
 *     R(m,m,N,M) X_ref(n,N,M) U_ref(m,N,M) lx,ux(n,N,M) lu,uu(m,N,M)
 *     slew_reg(M) slew_reg0(M) slew_um1(m,M);  outputs X(n,N,M) (WITHOUT x0), U(m,N,M).
 * i.e. NumPy C-order (M,N,n) vectors and (M,N,cols,rows) matrices.  For the batched handle API a
 * leading problem dimension B is simply more particles: index g = b*M + i.
 *
 * OPTIONAL ARGUMENTS use the reference's NaN sentinels (`c_interface.jl:56-70`): any NaN in
 * lx|ux disables the state box, in lu|uu the control box, in slew_reg the slew penalty, in
 * slew_reg0|slew_um1 the u_{-1} term.  The handle API additionally accepts NULL for "absent".
 * Nc < 0 means Nc = N (`main.jl:127-128`).
 */
#ifndef PMPC_B200_H
#define PMPC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. Byte-compatible replacements of the reference entry points
 *    (prototype copied in meaning from `PMPC.jl/pmpcjl/module.cpp:9-23`).
 *    Host pointers in, host pointers out; caller owns every buffer; no pointer is retained.
 *    On failure X_out/U_out are NaN-filled (reference convention, `scp_mpc.py:391-394`).
 * ------------------------------------------------------------------------------------------ */
void c_lqp_solve(double *X_out, double *U_out, size_t xdim, size_t udim, size_t N, size_t M,
                 long long Nc, double *x0, double *f, double *fx, double *fu, double *X_prev,
                 double *U_prev, double *Q, double *R, double *X_ref, double *U_ref, double *lx,
                 double *ux, double *lu, double *uu, double reg_x, double reg_u, double *slew_reg,
                 double *slew_reg0, double *slew_um1, long long verbose);

/* Cone route (`c_interface.jl:146-214`).  Served only where its minimiser coincides with the
 * QP's: M == 1 (or Nc-coupled particles with identical costs are NOT assumed), hard constraints
 * (smooth_alpha NaN).  Otherwise outputs are NaN-filled — see DESIGN.md "out of scope". */
void c_lcone_solve(double *X_out, double *U_out, size_t xdim, size_t udim, size_t N, size_t M,
                   long long Nc, double *x0, double *f, double *fx, double *fu, double *X_prev,
                   double *U_prev, double *Q, double *R, double *X_ref, double *U_ref, double *lx,
                   double *ux, double *lu, double *uu, double reg_x, double reg_u,
                   double *slew_reg, double *slew_reg0, double *slew_um1, long long verbose,
                   double smooth_alpha, char *solver);

/* Runtime hooks of the reference's ahead-of-time compiled Julia library (`PMPC.jl/pmpcjl/include/julia_init.h:1-2`),
 * called by its pybind shim at import (`module.cpp:122-125`).  No-ops here — exported so that the reference's
 * UNMODIFIED `module.cpp` links against libpmpc_b200.so in place of libPMPC.so + libjulia.so (INTEGRATION.md level 1,
 * built by oracle/ref_install.sh and exercised by tests/test_reference_dropin.py). */
void init_julia(int argc, char *argv[]);
void shutdown_julia(int retcode);

/* ------------------------------------------------------------------------------------------
 * 2. Handle API (the fast path): problem data stays resident in HBM, the whole SCP loop of
 *    `pmpc/scp_mpc.py:337-428` runs on the device for built-in dynamics.
 * ------------------------------------------------------------------------------------------ */
typedef struct pmpc_handle_s *pmpc_handle_t;

enum {
  PMPC_OK = 0,
  PMPC_ERR_NO_DEVICE = 1,   /* no usable CUDA device / driver */
  PMPC_ERR_BAD_ARG = 2,
  PMPC_ERR_UNSUPPORTED = 3, /* dimensions outside the compiled kernel table */
  PMPC_ERR_CUDA = 4,
  PMPC_ERR_STATE = 5        /* call order (e.g. solve before data was set) */
};

enum { PMPC_DYN_AFFINE = 0, /* f,fx,fu supplied by the caller (pmpc_set_linearization) or LTV */
       PMPC_DYN_DUBINS = 1  /* reference `tests/dubins_car.py:48-90`, params (v_scale,w_scale,T) */ };

typedef struct {
  double rho0;         /* initial ADMM penalty                         (default 1.0)   */
  double alpha;        /* ADMM over-relaxation                          (1.6)          */
  double eps_abs;      /* ADMM primal/dual residual tolerances          (1e-9)         */
  double eps_rel;      /*                                               (1e-9)         */
  double kkt_tol;      /* relative KKT tolerance accepted from a polish (1e-9)         */
  double adapt_tol;    /* refactor when the rho estimate moves by this factor (3.0)    */
  int max_admm_iter;   /* per QP                                        (4000)         */
  int check_every;     /* ADMM iterations between residual checks       (10)           */
  int polish;          /* active-set (masked Riccati) polish on/off     (1)            */
  int polish_iters;    /* active-set passes per QP before the interior-point fallback (16) */
  int adapt_rho;       /*                                               (1)            */
  int warm_start;      /* carry (w,y,active set) across SCP iterations  (1)            */
  int fused;           /* fused 2-sweep polish passes with on-the-fly linearisation (1) */
  int ring;            /* bulk-async (TMA) prefetch ring in shared memory for the fused sweeps (1) */
} pmpc_settings;

void pmpc_default_settings(pmpc_settings *s);
const char *pmpc_last_error(void);
const char *pmpc_version(void);
int pmpc_device_count(void);

/* B independent problems of M particles each (N stages, xdim states, udim controls, Nc shared
 * leading controls).  device < 0 selects the current device. */
int pmpc_create(pmpc_handle_t *h, int device, size_t B, size_t M, size_t N, size_t xdim,
                size_t udim, long long Nc);
int pmpc_destroy(pmpc_handle_t h);
int pmpc_set_settings(pmpc_handle_t h, const pmpc_settings *s);

/* Cost / constraints / start (host pointers, reference layout with leading B).  Any of
 * lx,ux,lu,uu,slew_* may be NULL or NaN (absent).  X_ref/U_ref NULL = zeros. */
int pmpc_set_problem(pmpc_handle_t h, const double *x0, const double *Q, const double *R,
                     const double *X_ref, const double *U_ref, const double *lx, const double *ux,
                     const double *lu, const double *uu, double reg_x, double reg_u,
                     const double *slew_reg, const double *slew_reg0, const double *slew_um1);
/* Matrix block order of the handle API's inputs (Q, R, fx, fu, A, B) and of pmpc_get_linearization:
 * 0 (default) = column-major blocks, the reference C ABI; 1 = row-major blocks, i.e. NumPy C-order
 * (M,N,rows,cols) arrays as `pmpc.solve` receives them, so that no host-side transpose is needed. */
int pmpc_set_layout(pmpc_handle_t h, int row_major_matrices);
/* Inputs that are identical for many particles need not be tiled on the host (the reference's API makes the caller
 * tile Q, R, references and bounds over the particle axis, `tests/simple.py:20-30`; for a batch that is most of the
 * upload).  Bit k of all_mask: input k of pmpc_set_problem / pmpc_set_warm_start is ONE (N, .) block shared by every
 * particle of every problem; bit k of per_problem_mask: one block per problem, (B, N, .), shared by its particles.
 * Bits: 0 Q, 1 R, 2 X_ref, 3 U_ref, 4 lx, 5 ux, 6 lu, 7 uu, 8 X_prev, 9 U_prev.  The expansion happens in the device
 * packing kernels; sticky until changed; default 0 (every input carries the full (B*M, N, .) layout). */
int pmpc_set_broadcast(pmpc_handle_t h, unsigned all_mask, unsigned per_problem_mask);
/* Page-locked host buffers (cudaHostAlloc) for inputs / outputs of the handle API: H2D / D2H copies from
 * them run at full PCIe / C2C speed and asynchronously. */
int pmpc_host_alloc(void **p, size_t bytes);
int pmpc_host_free(void *p);
/* Warm start X_prev (n,N,B*M), U_prev (m,N,B*M); NULL = zeros.  Also resets the ADMM state. */
int pmpc_set_warm_start(pmpc_handle_t h, const double *X_prev, const double *U_prev);

/* Built-in device dynamics.  params: (nparam, B*M) column-major; DUBINS nparam = 3
 * (v_scale, w_scale, T), `eps` is the +-eps control offset (reference 1e-6). */
int pmpc_set_dynamics(pmpc_handle_t h, int kind, const double *params, size_t nparam, double eps);
/* Caller-supplied linearisation about (X_prev, U_prev) — the host-callable path that keeps
 * arbitrary Python `f_fx_fu_fn` working (`scp_mpc.py:338-342`).  Sets kind = AFFINE. */
int pmpc_set_linearization(pmpc_handle_t h, const double *f, const double *fx, const double *fu);
/* LTV system x+ = A_j x + B_j u + c_j given once (fx = A, fu = B, c may be NULL). */
int pmpc_set_ltv(pmpc_handle_t h, const double *A, const double *Bm, const double *c);

/* User device dynamics (north_star: "a C-ABI hook for registering user device dynamics").  Replaces the
 * reference's host callable `f_fx_fu_fn(X_, U_prev) -> f, fx, fu` (pmpc/scp_mpc.py:338-342; e.g.
 * tests/dubins_car.py:48-90 evaluated through torch autograd on the CPU every SCP iteration) by CUDA C++
 * source that defines
 *     __device__ void <fn_name>(const double *x, const double *u, const double *p,
 *                               double *f, double *fx, double *fu);
 * x (xdim) is the state the control u (udim) is applied to, p the particle's `nparam` parameters; the
 * function writes f (xdim) = the next state, fx[r*xdim + t] = df_r/dx_t and fu[r*udim + t] = df_r/du_t
 * (row-major; all three are zero on entry).  The source is compiled at run time with NVRTC for sm_100a into
 * a linearisation kernel that writes straight into the handle's HBM layout; compile errors are returned
 * through pmpc_last_error().  params: (nparam, B*M) column-major as for pmpc_set_dynamics, NULL if
 * nparam == 0.  Afterwards pmpc_linearize / pmpc_qp_solve / pmpc_scp_solve use it; the SCP loop is then
 * driven from the host (one solve-kernel launch per iteration, 4 bytes read back per iteration) and
 * pmpc_scp_solve_async is synchronous. */
int pmpc_register_dynamics(pmpc_handle_t h, const char *cuda_src, const char *fn_name,
                           const double *params, size_t nparam);

/* Linearise only (built-in or registered dynamics) about the handle's (X_prev, U_prev); results readable
 * with pmpc_get_linearization in reference layout. */
int pmpc_linearize(pmpc_handle_t h);
int pmpc_get_linearization(pmpc_handle_t h, double *f, double *fx, double *fu);

/* One convex sub-problem at the current linearisation (= one `aff_solve`, scp_mpc.py:370),
 * followed by the SCP bookkeeping of scp_mpc.py:397-407 (X_prev <- X unless keep_prev). */
int pmpc_qp_solve(pmpc_handle_t h, int keep_prev);
/* Whole SCP loop on the device (built-in dynamics or LTV): up to max_it iterations, stop when
 * max_res < res_tol (scp_mpc.py:424).  Asynchronous on the handle's stream; pmpc_sync waits. */
int pmpc_scp_solve(pmpc_handle_t h, int max_it, double res_tol);
int pmpc_scp_solve_async(pmpc_handle_t h, int max_it, double res_tol);
int pmpc_sync(pmpc_handle_t h);

/* Results.  X (n,N,B*M) without x0, U (m,N,B*M). */
int pmpc_get_solution(pmpc_handle_t h, double *X, double *U);
/* Same with X in trajectory form (n,N+1,B*M): stage 0 is x0, what `solve` returns (static_backend.py:311). */
int pmpc_get_trajectory(pmpc_handle_t h, double *X, double *U);
/* Per problem b: info[b*8 + k], k = 0 scp_iters, 1 total ADMM iterations, 2 factorisations,
 * 3 polish passes, 4 status bits (0 ok, 1 NaN failure, 2 ADMM hit max_iter in some QP, 4 sharded exchange timed out),
 * 5 QPs finished by polish, 6 QPs finished by ADMM tolerance, 7 sweeps over the horizon. */
int pmpc_get_info(pmpc_handle_t h, long long *info);
/* hist (4, max_it, B): obj, resid, admm_iters(it), kkt_violation of the accepted QP solution. */
int pmpc_get_hist(pmpc_handle_t h, double *hist, int max_it);

/* Device-time of the last solve kernel in milliseconds (CUDA events on the handle stream). */
int pmpc_last_kernel_ms(pmpc_handle_t h, float *ms);

/* Measurement helpers (bench.py): CUDA events recorded on the handle's own stream (torch's
 * events do not see it), the number of kernels this library has launched in this process, and
 * two micro-benchmarks that measure the roofline denominators on the device at hand:
 * kind 0 = HBM copy bandwidth in GB/s (read + write bytes of a 1 GiB device-to-device copy
 * kernel), kind 1 = fp64 FMA throughput in TFLOP/s (register-resident DFMA chains), kind 2 = fp64 tensor-core
 * (DMMA m8n8k4) throughput, kinds 3 / 4 = DFMA with one operand read from shared memory by a broadcast LDS.64 per FMA /
 * LDS.128 per two FMAs (the operand pattern of the warp-per-particle Riccati kernel). */
int pmpc_timer_start(pmpc_handle_t h);
int pmpc_timer_stop(pmpc_handle_t h, float *ms);
long long pmpc_launch_count(void);
/* Resident CTAs per SM of the solve kernel this handle would launch, and the SM count: batch sizes
 * that are multiples of ctas_per_sm * num_sms fill the GPU in whole waves (one CTA = one problem). */
int pmpc_get_occupancy(pmpc_handle_t h, int *ctas_per_sm, int *num_sms);
int pmpc_microbench(int device, int kind, double *result);

/* ------------------------------------------------------------------------------------------
 * 2b. Particle sharding of ONE consensus problem over the GPUs of a node (one process per GPU).
 *     Rank r creates its handle with its own slice of the particles (M = local count, same B, N, Nc) and
 *     gives every rank the shared-block bounds of global particle 0 in its local particle 0
 *     (`lqp_utils.jl:329-330`).  Each rank exports its exchange buffer as a 64-byte CUDA IPC handle, the
 *     host plumbing (torch.distributed / MPI / files) all-gathers the handles, pmpc_comm_connect maps the
 *     peers' buffers.  From then on every solve on the handle exchanges the Nc*m consensus terms
 *     (sum_i H_i once per factorisation, sum_i g_i once per pass) and the pass statistics INSIDE the
 *     kernel by loads from peer memory over NVLink, combined in rank order (bit-identical on all ranks).
 *     All ranks must issue the same sequence of solves.  world <= 8, B <= resident CTAs of the kernel.
 * ------------------------------------------------------------------------------------------ */
int pmpc_comm_export(pmpc_handle_t h, void *ipc_handle_64);
int pmpc_comm_connect(pmpc_handle_t h, int rank, int world, const void *ipc_handles, size_t M_total);
/* The same exchange between handles of ONE process (several handles on one device — their kernels run
 * concurrently on the handles' own streams — or on peer-accessible devices): plain device pointers, no
 * IPC.  handles[r] is rank r's handle (handles[rank] == h); pmpc_comm_export must have been called on
 * every one of them (it allocates the exchange buffer).  Launch the solves of all ranks with
 * pmpc_scp_solve_async before waiting on any of them. */
int pmpc_comm_connect_local(pmpc_handle_t h, int rank, int world, const pmpc_handle_t *handles,
                            size_t M_total);
/* The in-kernel exchange spins on the peers' sequence flags for at most this long (default 30 s); on
 * expiry the launch is abandoned, status bit 2 (value 4) is set in info[4] and pmpc_sync returns
 * PMPC_ERR_STATE — a failed rank does not hang the node. */
int pmpc_comm_set_timeout(pmpc_handle_t h, double seconds);

/* ------------------------------------------------------------------------------------------
 * 3. int-returning variant of the compat entry with explicit settings and statistics.
 * ------------------------------------------------------------------------------------------ */
int pmpc_lqp_solve(double *X_out, double *U_out, size_t xdim, size_t udim, size_t N, size_t M,
                   long long Nc, const double *x0, const double *f, const double *fx,
                   const double *fu, const double *X_prev, const double *U_prev, const double *Q,
                   const double *R, const double *X_ref, const double *U_ref, const double *lx,
                   const double *ux, const double *lu, const double *uu, double reg_x,
                   double reg_u, const double *slew_reg, const double *slew_reg0,
                   const double *slew_um1, const pmpc_settings *settings, long long *info8);

#ifdef __cplusplus
}
#endif
#endif /* PMPC_B200_H */
