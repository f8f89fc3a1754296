// Device-side particle-SCP solver: one thread per particle, one CTA per problem.
//
// Replaces, per SCP iteration, what the reference does in
//   pmpc/scp_mpc.py:337-428 (loop + bookkeeping)  ->  scp_kernel
//   PMPC.jl/src/lqp_utils.jl:2-393 (QP assembly) + osqp_solver.jl:34-72 (OSQP)  ->  Riccati-ADMM + polish
// See DESIGN.md for the maths and the data layout.
//
// DATA LAYOUT (HBM): warp-tiled structure-of-arrays (TILE = 32 particles per tile; for NX >= 8 TILE = 1, i.e. every
// particle's stage block is contiguous and Mpad = M — see Solver::TILE).  Particle i of problem b owns slot
// g = b*Mpad + i (Mpad = M rounded up to a warp for M > 1, so a warp never straddles two problems);
// per-(particle, stage) arrays are stored as
//     arr[(((g / 32) * N + j) * E + e) * 32 + g % 32]     j = stage, e = element of E
// i.e. for every warp of 32 particles the E elements of a stage form ONE contiguous block of E*256
// bytes (one bulk async copy), consecutive stages are adjacent, and each lane access is a fully
// coalesced 256-byte row at a compile-time offset from the warp's stage pointer.  Per-particle arrays
// (x0, dynamics parameters, boundary value function) are arr[e * P + g].
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dynamics.cuh"
#include "tma.cuh"

namespace pmpc {

enum { MODE_POLISH = 0, MODE_ADMM = 1 };
enum { DYN_AFFINE = 0, DYN_DUBINS = 1 };  // == PMPC_DYN_* of include/pmpc_b200.h
enum { ST_KKT = 0, ST_UVIOL, ST_XVIOL, ST_CHANGED, ST_RP, ST_RD, ST_SP, ST_SD, ST_DX, ST_DU, ST_OBJ, ST_NAN,
       ST_GSCALE, ST_ZSCALE, ST_XACT, ST_COUNT };
// small CTA reductions (Solver::cta_reduce): values per call, and the largest shared block solved in registers
enum { RCAP = 24, DSMALL = 4 };

struct FactorSet {
  double *K, *E, *d, *e;      // per (particle, stage): feedback gain, masked H^-1, S c (+fix), W'c (+fix)
  double *Pc, *Wc, *Zc;       // per particle: value function at the consensus boundary
  double *Hc, *Hci, *hb;      // per problem: shared-block Hessian, its (masked) inverse, Hc*beta
};

struct DevProb {
  int B, M, N, n, m, Nc, D;   // logical sizes; D = Nc * NU (padded control dim)
  int Mpad;                   // particle slots per problem (M == 1 ? 1 : M rounded up to the tile)
  int tile;                   // particles per layout tile: 32, or 1 for NX >= 8 (== Solver::TILE)
  int P;                      // number of particle slots, B * Mpad rounded up to 32
  int dyn_kind;
  int has_ubox, has_xbox, has_slew;
  double reg_x, reg_u, dyn_eps;
  // constants of the problem
  double *Qs, *Rs, *qref, *rref, *xref, *uref, *lu, *uu, *lx, *ux;
  double *x0, *slew, *dynp, *rdiag0;  // per particle; rdiag0 = extra stage-0 diagonal of Rs (s0 [+ s if N==1])
  // current iterate and its linearisation
  double *Xp, *Up, *A, *Bm, *c, *q, *r;
  FactorSet fs[2];            // [MODE_POLISH], [MODE_ADMM]
  double *kff, *X, *U;
  double *hand;               // fused path: per-stage hand-off block [sparse stage | K | k | H01] from bsweep to fsweep (fused.inl)
  double *wu, *yu, *wx, *yx;
  signed char *act, *actn;    // per (particle, stage, control): -1 lower, +1 upper, 0 free
  // per problem
  signed char *actc, *actcn;  // shared-block active set (D)
  double *gc, *uc, *muc;      // (D)
  double *scratch;            // reduction scratch: scratch_stride doubles per problem
  size_t scratch_stride;
  double *hist;               // (4, hist_len, B)
  int hist_len;
  long long *info;            // (8, B) -> see pmpc_b200.h
  double *rho;                // (B) ADMM penalty carried across QPs
  // settings
  double rho0, alpha, eps_abs, eps_rel, kkt_tol, adapt_tol;
  int max_admm_iter, check_every, polish, polish_iters, adapt_rho, warm_start, fused;
  // fused-path resources decided at launch: shared-memory ring on/off, uniform-table size, per-problem flags
  int ring, tab_doubles;
  const int* uniform;
  // run parameters
  int max_it, keep_prev;
  int hist_off;               // first hist row of this launch (host-driven SCP loop: one iteration per launch)
  double res_tol;
  // particle sharding of ONE consensus problem over the GPUs of a node (SURVEY.md 8e case 2): this rank
  // holds M of the M_total particles; the Nc*m consensus terms and the pass statistics are exchanged
  // INSIDE the kernel through peer-mapped buffers (NVLink loads), see Solver::xreduce
  int rank, world, M_total, comm_cap;   // comm_cap: doubles per exchange slot
  unsigned long long seq0;              // exchange sequence number before this launch (same on every rank)
  double* peer[8];                      // exchange buffer of every rank (peer[rank] is this GPU's own)
  unsigned long long xtimeout_ns;       // bounded spin of the exchange (pmpc_comm_set_timeout; default 30 s)
  // warp-per-particle kernel (wpp.cuh): team exchange buffers [per problem: Tc part slots | bcast | flags], their slot
  // size, the team sequence number before this launch, per-warp / per-CTA consensus partials, and the hand-over flags
  // (fb[b] = 1: problem b is continued by the thread-per-particle kernel launched behind it with resume = 1)
  double* tm;
  int tm_cap;
  unsigned long long tseq0;
  double* wscr;
  int* fb;
  int* queue;                 // two-problems-per-warp kernel: next pair of problems to hand out (persistent CTAs)
  unsigned char* xany;        // warp-per-particle kernel: per particle, some state multiplier (yx) is non-zero
  int resume;
  double al_mu;               // warp-per-particle kernel: starting penalty of the multiplier method that carries active state bounds (0 = hand over)
  int half_ipm;               // two-problems-per-warp kernel: interior-point fallback inside the kernel (1) or hand the problem over (0)
  int* xerr;                            // device flag: an exchange timed out (read back by pmpc_sync)
};

__host__ __device__ constexpr int pk(int r, int t, int n) { return r * n - (r * (r - 1)) / 2 + (t - r); }

template <int NU>
__device__ __forceinline__ void masked_spd_inverse(const double* H, const bool* fixed, double* E) {
  // E = embed( inv(H[free,free]) ), zeros on fixed rows/cols.  H symmetric positive definite.
  if (NU == 1) {
    E[0] = fixed[0] ? 0.0 : 1.0 / H[0];
    return;
  }
  if (NU == 2) {
    const double h00 = fixed[0] ? 1.0 : H[0], h11 = fixed[1] ? 1.0 : H[3];
    const double h01 = (fixed[0] || fixed[1]) ? 0.0 : H[1];
    const double idet = 1.0 / (h00 * h11 - h01 * h01);
    E[0] = fixed[0] ? 0.0 : h11 * idet;
    E[3] = fixed[1] ? 0.0 : h00 * idet;
    E[1] = E[2] = -h01 * idet;
    return;
  }
  double L[NU * NU];
  for (int a = 0; a < NU; ++a)
    for (int b = 0; b < NU; ++b) {
      const bool fx = fixed[a] || fixed[b];
      L[a * NU + b] = fx ? (a == b ? 1.0 : 0.0) : H[a * NU + b];
    }
  // Cholesky (lower), in place
  for (int k = 0; k < NU; ++k) {
    double dk = L[k * NU + k];
    for (int l = 0; l < k; ++l) dk -= L[k * NU + l] * L[k * NU + l];
    dk = sqrt(dk);
    L[k * NU + k] = dk;
    const double idk = 1.0 / dk;
    for (int i = k + 1; i < NU; ++i) {
      double v = L[i * NU + k];
      for (int l = 0; l < k; ++l) v -= L[i * NU + l] * L[k * NU + l];
      L[i * NU + k] = v * idk;
    }
  }
  // inverse column by column: solve L L' x = e_t
  for (int t = 0; t < NU; ++t) {
    double y[NU];
    for (int i = 0; i < NU; ++i) {
      double v = (i == t) ? 1.0 : 0.0;
      for (int l = 0; l < i; ++l) v -= L[i * NU + l] * y[l];
      y[i] = v / L[i * NU + i];
    }
    for (int i = NU - 1; i >= 0; --i) {
      double v = y[i];
      for (int l = i + 1; l < NU; ++l) v -= L[l * NU + i] * y[l];
      y[i] = v / L[i * NU + i];
    }
    for (int i = 0; i < NU; ++i) E[i * NU + t] = (fixed[i] || fixed[t]) ? 0.0 : y[i];
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// PPW: problems per warp.  2 = two problems of at most 16 particles share a warp (and its 32-slot layout tile): half
// the lanes of an M = 16 problem would otherwise mirror a neighbour (BASELINE config 3).  Only the fused active-set
// passes run in that mode (run_half / fused_qp_half); everything else is handed to the PPW = 1 kernel of the same shape.
template <int NX, int NU, bool GROUPED, int PPW = 1>
struct Solver {
  static constexpr int NQ = NX * (NX + 1) / 2;
  static constexpr int NR = NU * (NU + 1) / 2;
  // inner-loop unroll factor: full unroll keeps small systems in registers; for the larger shapes (per-thread matrices
  // in local memory) it turns the dynamic indices into constant offsets and lets the loads pipeline — C4 (n = 12):
  // 1.53 s without unrolling, 0.45 s at 8, 0.21 s fully unrolled; bounded above by compile time and code size
  static constexpr int UNR = (NX <= 4) ? 64 : (NX <= 12 ? 16 : (NX <= 16 ? 8 : 4));
  // particles per layout tile (DATA LAYOUT above): 32 for the thread-per-particle shapes; 1 for NX >= 8, where the
  // stage block of ONE particle is contiguous — the layout of the warp-per-particle kernel (wpp.cuh), which these
  // thread-per-particle kernels only back up (state-box ADMM, interior-point fallback)
  static constexpr int TILE = (NX >= 8) ? 1 : 32;

  const DevProb& pb;
  const size_t P;
  const int N, Nc, D, T, tid;
  int prob, g;
  bool valid;
  int hl;       // PPW = 2: lane within the half-warp that owns the problem (PPW = 1: == tid)
  bool pvalid;  // PPW = 2: this half-warp has a problem (the last warp of an odd batch has one empty half)
  double s;  // slew weight of this particle
  double* red;  // shared: ST_COUNT doubles (+ D handled in global)
  double* red2; // shared: RCAP doubles per warp
  // current iterate (Xc, Uc) and linearisation point (Xpv, Upv): the fused path swaps the two buffer
  // pairs instead of copying X_prev <- X (scp_mpc.py:407)
  double *Xc, *Uc, *Xpv, *Upv;
  // fused path (fused.inl): lane / warp ids, this lane's column in its warp tile, the warp's
  // shared-memory ring + mbarriers + running slot counter, the CTA's uniform cost table (or null)
  int lane, wid, col;
  double* ring;
  uint64_t* bars;
  uint32_t rcount;
  const double* tab;
  bool has_tab;
  bool flip;  // fused path: (Xc, Uc) / (Xpv, Upv) currently are (pb.Xp, pb.Up) / (pb.X, pb.U)
  bool dead;  // whole warp beyond the last problem (thread-per-problem launch tail)
  unsigned long long xseq;  // cross-GPU exchange counter (uniform over the CTA, identical on every rank)
  bool xdead;               // an exchange timed out (peer missing): no further exchanges, the SCP loop stops

  __device__ Solver(const DevProb& pb_, double* red_, double* red2_)
      : pb(pb_), P(pb_.P), N(pb_.N), Nc(GROUPED ? pb_.Nc : 0), D(GROUPED ? pb_.D : 0), T(blockDim.x),
        tid(threadIdx.x), red(red_), red2(red2_), Xc(pb_.X), Uc(pb_.U), Xpv(pb_.Xp), Upv(pb_.Up), xseq(pb_.seq0), xdead(false) {
    // lanes without a particle mirror the last valid lane of their own warp (same warp tile)
    dead = false;
    hl = tid;
    pvalid = true;
    if (GROUPED && PPW == 2) {
      // two problems per warp: lanes 0-15 own problem 2 b, lanes 16-31 problem 2 b + 1 (Mpad == 16: both in one tile)
      assign_pair((int)blockIdx.x * (int)(blockDim.x >> 5) + (tid >> 5));
      return;
    } else if (GROUPED) {
      prob = blockIdx.x;
      valid = tid < pb.M;
      g = prob * pb.Mpad + min(tid, min(pb.M - 1, tid | 31));
    } else {
      const int idx = blockIdx.x * blockDim.x + threadIdx.x;
      valid = idx < pb.B * pb.M;
      dead = (idx & ~31) >= pb.B * pb.M;
      g = dead ? pb.B * pb.M - 1 : min(idx, min(pb.B * pb.M - 1, idx | 31));
      prob = g;  // M == 1
    }
    s = pb.has_slew ? pb.slew[g] : 0.0;
  }

  // PPW = 2: this warp takes the pair of problems (2 pair, 2 pair + 1); lanes 0-15 own the first, lanes 16-31 the second
  // (Mpad == 16: both in one tile).  Resets everything the constructor derives from the problem index.
  __device__ __forceinline__ void assign_pair(int pair) {
    hl = tid & 15;
    const int pr = pair * 2 + ((tid >> 4) & 1);
    pvalid = pr < pb.B;
    prob = pvalid ? pr : pb.B - 1;  // an empty half mirrors the last problem and never stores
    valid = pvalid && hl < pb.M;
    g = prob * pb.Mpad + min(hl, pb.M - 1);
    s = pb.has_slew ? pb.slew[g] : 0.0;
    Xc = pb.X; Uc = pb.U; Xpv = pb.Xp; Upv = pb.Up;
  }

  // Helper instance for the warp-per-particle kernel (wpp.cuh): only the CTA-cooperative shared-block routines and the
  // cross-GPU exchange are used; it owns no particle.
  __device__ Solver(const DevProb& pb_, double* red_, double* red2_, int prob_)
      : pb(pb_), P(pb_.P), N(pb_.N), Nc(GROUPED ? pb_.Nc : 0), D(GROUPED ? pb_.D : 0), T(blockDim.x),
        tid(threadIdx.x), red(red_), red2(red2_), Xc(pb_.X), Uc(pb_.U), Xpv(pb_.Xp), Upv(pb_.Up), xseq(pb_.seq0), xdead(false) {
    dead = false;
    prob = prob_;
    valid = false;
    g = prob_ * pb.Mpad;
    s = 0.0;
  }

  // warp-tile SoA index of element e (of E) of stage j for particle slot gg (see DATA LAYOUT above)
  __device__ __forceinline__ size_t ixg(int j, int E, int e, size_t gg) const {
    return (((gg / TILE) * (size_t)N + (size_t)j) * (size_t)E + (size_t)e) * TILE + (gg % TILE);
  }
  __device__ __forceinline__ size_t ix(int j, int E, int e) const { return ixg(j, E, e, (size_t)g); }
  __device__ __forceinline__ double& at(double* a, int j, int E, int e) const { return a[ix(j, E, e)]; }
  __device__ __forceinline__ double ld(const double* a, int j, int E, int e) const { return a[ix(j, E, e)]; }
  __device__ __forceinline__ double ldp(const double* a, int e) const { return a[(size_t)e * P + g]; }
  __device__ __forceinline__ double* scr() const { return pb.scratch + (size_t)prob * pb.scratch_stride; }

  // ---- reductions over the particles of the problem (deterministic order) -------------------
  // every thread wrote scr()[k*T + tid], k < cnt.  out[k] = sum (or max) over threads.
  // max_mask_from < 0: every row is summed; otherwise rows are max-reduced except row `max_mask_from` and the
  // rows whose bit is set in sum_mask (summed).
  __device__ __noinline__ void reduce(int cnt, double* out, int max_mask_from = -1, unsigned sum_mask = 0u) {
    if (!GROUPED) return;
    __syncthreads();
    const int lane = tid & 31, w = tid >> 5, nw = T >> 5;
    const double* sc = scr();
    for (int k = w; k < cnt; k += nw) {
      double acc = 0.0;
      const bool is_sum = (max_mask_from < 0) || (k == max_mask_from) || (k < 32 && ((sum_mask >> k) & 1u));
      if (is_sum) {
        for (int t = lane; t < T; t += 32) acc += sc[(size_t)k * T + t];
        acc = warp_sum(acc);
      } else {
        acc = -1e300;
        for (int t = lane; t < T; t += 32) acc = fmax(acc, sc[(size_t)k * T + t]);
        acc = warp_max(acc);
      }
      if (lane == 0) out[k] = acc;
    }
    __syncthreads();
  }


  // ---- small CTA reduction without global memory ------------------------------------------------
  // v[k], k < cnt <= RCAP, of every thread -> v[k] = sum over the CTA (bit k of sum_mask set) or max, replicated in
  // every thread.  Warp butterfly, then the per-warp partials through shared memory, combined in warp order
  // (deterministic).  Two barriers, no global round trip — the per-pass bookkeeping used to go through the
  // problem's global scratch with one barrier pair per reduction plus ~20 more in the shared-block solve, which
  // cost about a quarter of the kernel time at C2 size (ncu source view: 8 % of the instructions, 33 % of the
  // stall samples).
  __device__ __forceinline__ void cta_reduce(double* v, int cnt, unsigned sum_mask) {
    if (!GROUPED) return;
    if (PPW == 2) {
      // the problem is a half-warp: butterfly over 16 lanes.  Same association as the one-problem-per-warp kernel,
      // whose first step only adds the zeros of the mirror lanes: results are bit-identical between the two.
      for (int k = 0; k < cnt; ++k) {
        const bool is_sum = (sum_mask >> k) & 1u;
        double r = v[k];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          const double t = __shfl_xor_sync(0xffffffffu, r, o);
          r = is_sum ? r + t : fmax(r, t);
        }
        v[k] = r;
      }
      return;
    }
    const int ln = tid & 31, w = tid >> 5, nw = T >> 5;
    for (int k = 0; k < cnt; ++k) {
      const bool is_sum = (sum_mask >> k) & 1u;
      const double r = is_sum ? warp_sum(v[k]) : warp_max(v[k]);
      if (ln == 0) red2[w * RCAP + k] = r;
    }
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const bool is_sum = (sum_mask >> k) & 1u;
      double acc = red2[k];
      for (int ww = 1; ww < nw; ++ww) acc = is_sum ? acc + red2[ww * RCAP + k] : fmax(acc, red2[ww * RCAP + k]);
      v[k] = acc;
    }
    __syncthreads();
  }

  // ---- cross-GPU all-reduce of a short vector (particle sharding) -------------------------------
  // Every rank publishes its partial vector in its own exchange slot (double-buffered by sequence
  // parity), raises the slot's flag to the sequence number with system-scope release, then reads the
  // slots of ALL ranks over NVLink and combines them in rank order — the same order on every GPU, so
  // the result is bit-identical everywhere and no broadcast is needed.  Payloads are D, D*D or a
  // handful of scalars: latency-bound, hence inside the kernel instead of a host-launched collective.
  // sum_mask: bit k set -> element k is summed, otherwise max-reduced; all_sum overrides.
  __device__ __noinline__ void xreduce(double* vec, int len, bool all_sum, unsigned sum_mask = 0u) {
    if (!GROUPED || pb.world <= 1) return;
    __syncthreads();
    if (xdead) return;  // a previous exchange timed out: the launch is being abandoned (status bit 2 of info[4])
    ++xseq;
    const size_t slot = (size_t)pb.comm_cap + 1;  // data + flag
    const size_t off = ((size_t)prob * 2 + (size_t)(xseq & 1ull)) * slot;
    volatile double* mine = pb.peer[pb.rank] + off;
    for (int k = tid; k < len; k += T) mine[k] = vec[k];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      unsigned long long* flag = reinterpret_cast<unsigned long long*>(pb.peer[pb.rank] + off + pb.comm_cap);
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(xseq) : "memory");
      // bounded spin: a peer that never arrives (failed rank, kernel not co-resident) must not hang the node
      unsigned long long t_start = 0, spins = 0;
      bool timed_out = false;
      for (int r = 0; r < pb.world && !timed_out; ++r) {
        const unsigned long long* pf = reinterpret_cast<const unsigned long long*>(pb.peer[r] + off + pb.comm_cap);
        unsigned long long v;
        for (;;) {
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(pf) : "memory");
          if (v >= xseq) break;
          if ((++spins & 1023ull) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_start == 0) t_start = now;
            else if (now - t_start > pb.xtimeout_ns) { timed_out = true; break; }
          }
        }
      }
      red2[RCAP - 1] = timed_out ? 1.0 : 0.0;
      if (timed_out) *pb.xerr = 1;
    }
    __syncthreads();
    if (red2[RCAP - 1] != 0.0) {
      xdead = true;
      __syncthreads();
      return;
    }
    __syncthreads();
    for (int k = tid; k < len; k += T) {
      const bool is_sum = all_sum || (k < 32 && ((sum_mask >> k) & 1u));
      double acc = is_sum ? 0.0 : -INFINITY;
      for (int r = 0; r < pb.world; ++r) {
        const volatile double* p = pb.peer[r] + off;
        const double v = p[k];
        acc = is_sum ? acc + v : fmax(acc, v);
      }
      vec[k] = acc;
    }
    __syncthreads();
  }

  // End of a launch in sharded mode: leave an EVEN number of exchanges behind.  A slot is overwritten two exchanges
  // later, which is safe inside a launch (finishing exchange n-1 proves every peer has read n-2); the next launch
  // starts its own sequence window with an odd number, i.e. in the slot of this launch's second-to-last exchange,
  // which every peer is known to have finished reading — without the padding a fast rank could overwrite the slot a
  // slow peer is still reading from the last exchange of the previous launch.
  __device__ __forceinline__ void xclose() {
    if (!GROUPED || pb.world <= 1) return;
    if ((xseq - pb.seq0) & 1ull) {
      red2[0] = 0.0;
      __syncthreads();
      xreduce(red2, 1, true);
    }
  }

  // ---- K1 + cost vectors: linearise about (X_prev, U_prev), build c, q, r --------------------
  __device__ __noinline__ void linearize_prep() {
    double xb[NX], xprev[NX];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) xb[e] = ldp(pb.x0, e);
    for (int j = 0; j < N; ++j) {
      double ub[NU];
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) ub[e] = ld(Upv, j, NU, e);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) xprev[e] = ld(Xpv, j, NX, e);
      if (pb.dyn_kind == DYN_DUBINS) {
        if (NX == 4 && NU == 2) {
          double f[4], fx[16], fu[8];
          dubins_f_fx_fu(xb, ub, ldp(pb.dynp, 0), ldp(pb.dynp, 1), ldp(pb.dynp, 2), pb.dyn_eps, f, fx, fu);
          double cc[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double v = f[r];
#pragma unroll
            for (int t = 0; t < 2; ++t) v -= fu[r * 2 + t] * ub[t];
            if (j >= 1) {
#pragma unroll
              for (int t = 0; t < 4; ++t) v -= fx[r * 4 + t] * xb[t];
            }
            cc[r] = v;
          }
          if (valid) {
#pragma unroll
            for (int e = 0; e < 16; ++e) at(pb.A, j, NX * NX, e) = (j >= 1) ? fx[e] : 0.0;
#pragma unroll
            for (int e = 0; e < 8; ++e) at(pb.Bm, j, NX * NU, e) = fu[e];
#pragma unroll
            for (int e = 0; e < 4; ++e) at(pb.c, j, NX, e) = cc[e];
          }
        }
      }
      // proximal / reference linear terms (lqp_utils.jl:164-214)
      if (valid) {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) at(pb.q, j, NX, e) = -ld(pb.qref, j, NX, e) - pb.reg_x * xprev[e];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) at(pb.r, j, NU, e) = -ld(pb.rref, j, NU, e) - pb.reg_u * ub[e];
      }
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) xb[e] = xprev[e];
    }
  }

  // ---- K2: backward matrix recursion over the free stages ------------------------------------
  __device__ __noinline__ void factor(int mode, double rho_u, double rho_x) {
    const FactorSet& F = pb.fs[mode];
    double Pm[NX * NX], W[NX * NU], Z[NU * NU];
#pragma unroll UNR
    for (int e = 0; e < NX * NX; ++e) Pm[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NX * NU; ++e) W[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU * NU; ++e) Z[e] = 0.0;
    for (int j = N - 1; j >= Nc; --j) {
      double S[NX * NX], A[NX * NX], Bm[NX * NU], SBW[NX * NU], H[NU * NU], E[NU * NU], cv[NX], ufix[NU];
      bool fixed[NU];
      // S = Qs + rho_x I + P
#pragma unroll UNR
      for (int r = 0; r < NX; ++r)
#pragma unroll UNR
        for (int t = r; t < NX; ++t) {
          const double v = ld(pb.Qs, j, NQ, pk(r, t, NX)) + Pm[r * NX + t] + (r == t ? rho_x : 0.0);
          S[r * NX + t] = v;
          S[t * NX + r] = v;
        }
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) Bm[e] = ld(pb.Bm, j, NX * NU, e);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) cv[e] = ld(pb.c, j, NX, e);
      // SBW = S B + W
#pragma unroll UNR
      for (int r = 0; r < NX; ++r)
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) {
          double acc = W[r * NU + t];
#pragma unroll UNR
          for (int k = 0; k < NX; ++k) acc += S[r * NX + k] * Bm[k * NU + t];
          SBW[r * NU + t] = acc;
        }
      // H = Rs + (rho_u + [j>=1] s) I + B'(SB+W) + W'B + Z
#pragma unroll UNR
      for (int a = 0; a < NU; ++a)
#pragma unroll UNR
        for (int b = a; b < NU; ++b) {
          double acc = ld(pb.Rs, j, NR, pk(a, b, NU)) + Z[a * NU + b];
          if (a == b) acc += rho_u + (j >= 1 ? s : 0.0);
#pragma unroll UNR
          for (int k = 0; k < NX; ++k) acc += Bm[k * NU + a] * SBW[k * NU + b] + W[k * NU + a] * Bm[k * NU + b];
          H[a * NU + b] = acc;
          H[b * NU + a] = acc;
        }
      // active set of this stage (polish only)
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        fixed[t] = false;
        ufix[t] = 0.0;
      }
      if (mode == MODE_POLISH && pb.has_ubox) {
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) {
          const signed char a = pb.act[ix(j, NU, t)];
          if (a != 0) {
            fixed[t] = true;
            ufix[t] = (a > 0) ? ld(pb.uu, j, NU, t) : ld(pb.lu, j, NU, t);
          }
        }
      }
      masked_spd_inverse<NU>(H, fixed, E);
      // d = S c + SBW ufix ; e = W'c + H ufix
      double dv[NX], ev[NU];
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = 0.0;
#pragma unroll UNR
        for (int k = 0; k < NX; ++k) acc += S[r * NX + k] * cv[k];
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += SBW[r * NU + t] * ufix[t];
        dv[r] = acc;
      }
      // the vector pass forms h = rh + B'eta + e + z with eta = d + qh + p, and d already carries
      // SBW*ufix (needed by the p recursion); remove its image B'(SBW ufix) here so that
      // h = h_u + H ufix exactly.
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) {
        double acc = 0.0;
#pragma unroll UNR
        for (int k = 0; k < NX; ++k) {
          double tk = 0.0;
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) tk += SBW[k * NU + t] * ufix[t];
          acc += W[k * NU + a] * cv[k] - Bm[k * NU + a] * tk;
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += H[a * NU + t] * ufix[t];
        ev[a] = acc;
      }
      double K[NU * NX];
      if (j >= 1) {
#pragma unroll UNR
        for (int e = 0; e < NX * NX; ++e) A[e] = ld(pb.A, j, NX * NX, e);
        // G = SBW' A ;  K = -E G
        double G[NU * NX];
#pragma unroll UNR
        for (int a = 0; a < NU; ++a)
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) acc += SBW[k * NU + a] * A[k * NX + t];
            G[a * NX + t] = acc;
          }
#pragma unroll UNR
        for (int a = 0; a < NU; ++a)
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int b = 0; b < NU; ++b) acc -= E[a * NU + b] * G[b * NX + t];
            K[a * NX + t] = acc;
          }
        // P <- A' S A + G' K   (S A staged in S row by row would alias; use a column buffer)
        double SA[NX * NX];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r)
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) acc += S[r * NX + k] * A[k * NX + t];
            SA[r * NX + t] = acc;
          }
#pragma unroll UNR
        for (int r = 0; r < NX; ++r)
#pragma unroll UNR
          for (int t = r; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) acc += A[k * NX + r] * SA[k * NX + t];
#pragma unroll UNR
            for (int a = 0; a < NU; ++a) acc += G[a * NX + r] * K[a * NX + t];
            Pm[r * NX + t] = acc;
            Pm[t * NX + r] = acc;
          }
      } else {
#pragma unroll UNR
        for (int e = 0; e < NU * NX; ++e) K[e] = 0.0;
      }
      // value-function slew blocks for the previous stage: W = -s K', Z = s I - s^2 E
#pragma unroll UNR
      for (int r = 0; r < NX; ++r)
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) W[r * NU + t] = -s * K[t * NX + r];
#pragma unroll UNR
      for (int a = 0; a < NU; ++a)
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) Z[a * NU + b] = (a == b ? s : 0.0) - s * s * E[a * NU + b];
      if (valid) {
#pragma unroll UNR
        for (int e = 0; e < NU * NX; ++e) at(F.K, j, NU * NX, e) = K[e];
#pragma unroll UNR
        for (int a = 0; a < NU; ++a)
#pragma unroll UNR
          for (int b = a; b < NU; ++b) at(F.E, j, NR, pk(a, b, NU)) = E[a * NU + b];
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) at(F.d, j, NX, e) = dv[e];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) at(F.e, j, NU, e) = ev[e];
      }
    }
    if (Nc >= 1 && valid) {
#pragma unroll UNR
      for (int e = 0; e < NX * NX; ++e) F.Pc[(size_t)e * P + g] = Pm[e];
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) F.Wc[(size_t)e * P + g] = W[e];
#pragma unroll UNR
      for (int e = 0; e < NU * NU; ++e) F.Zc[(size_t)e * P + g] = Z[e];
    }
  }

  // linear terms of the z-update at stage j
  __device__ __forceinline__ void lin_terms(int mode, int j, double rho_u, double rho_x, double* qh, double* rh) const {
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) {
      double v = ld(pb.q, j, NX, e);
      if (mode == MODE_ADMM && pb.has_xbox) v += -rho_x * ld(pb.wx, j, NX, e) + ld(pb.yx, j, NX, e);
      qh[e] = v;
    }
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) {
      double v = ld(pb.r, j, NU, e);
      if (mode == MODE_ADMM && pb.has_ubox && j >= Nc) v += -rho_u * ld(pb.wu, j, NU, e) + ld(pb.yu, j, NU, e);
      rh[e] = v;
    }
  }

  // ---- K4a: backward vector pass over the free stages; leaves (p, z) at the boundary ----------
  __device__ __noinline__ void backward(int mode, double rho_u, double rho_x, double* p, double* z) {
    const FactorSet& F = pb.fs[mode];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) p[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) z[e] = 0.0;
    for (int j = N - 1; j >= Nc; --j) {
      double qh[NX], rh[NU], eta[NX], h[NU], k[NU], Bm[NX * NU], E[NU * NU];
      lin_terms(mode, j, rho_u, rho_x, qh, rh);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) eta[e] = ld(F.d, j, NX, e) + qh[e] + p[e];
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) Bm[e] = ld(pb.Bm, j, NX * NU, e);
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) {
        double acc = rh[a] + ld(F.e, j, NU, a) + z[a];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) acc += Bm[r * NU + a] * eta[r];
        h[a] = acc;
      }
#pragma unroll UNR
      for (int a = 0; a < NU; ++a)
#pragma unroll UNR
        for (int b = a; b < NU; ++b) {
          const double v = ld(F.E, j, NR, pk(a, b, NU));
          E[a * NU + b] = v;
          E[b * NU + a] = v;
        }
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) {
        double acc = 0.0;
        if (mode == MODE_POLISH && pb.has_ubox) {
          const signed char ac = pb.act[ix(j, NU, a)];
          if (ac != 0) acc = (ac > 0) ? ld(pb.uu, j, NU, a) : ld(pb.lu, j, NU, a);
        }
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) acc -= E[a * NU + b] * h[b];
        k[a] = acc;
      }
      if (valid) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) at(pb.kff, j, NU, a) = k[a];
      }
      if (j >= 1) {
        // p <- A' eta + K' h
        double pn[NX];
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) {
          double acc = 0.0;
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) acc += ld(pb.A, j, NX * NX, r * NX + t) * eta[r];
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) acc += ld(F.K, j, NU * NX, a * NX + t) * h[a];
          pn[t] = acc;
        }
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) p[t] = pn[t];
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) z[a] = -s * k[a];
      }
    }
  }

  // ---- K3: gradient of particle cost w.r.t. the shared controls ------------------------------
  // hom = false: gradient at `uc` (NULL = 0) including all linear terms; hom = true: Hessian
  // action on the unit vector ek.  Result written to scratch rows [row0, row0 + D).
  __device__ __noinline__ void shared_grad(int mode, bool hom, int ek, const double* uc, double rho_u, double rho_x,
                              const double* pc, const double* zc, int row0, double* out = nullptr) {
    const FactorSet& F = pb.fs[mode];
    double x[NX];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) x[e] = 0.0;
    // forward over the shared stages (X used as scratch)
    for (int j = 0; j < Nc; ++j) {
      double u[NU], xn[NX];
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) u[t] = hom ? ((j * NU + t == ek) ? 1.0 : 0.0) : (uc ? uc[j * NU + t] : 0.0);
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = hom ? 0.0 : ld(pb.c, j, NX, r);
        if (j >= 1) {
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) acc += ld(pb.A, j, NX * NX, r * NX + t) * x[t];
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += ld(pb.Bm, j, NX * NU, r * NU + t) * u[t];
        xn[r] = acc;
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        x[r] = xn[r];
        if (valid) at(Xc, j, NX, r) = xn[r];
      }
    }
    double a[NX];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) a[e] = 0.0;
    double* sc = scr();
    for (int j = Nc - 1; j >= 0; --j) {
      double u[NU], up[NU], un[NU], an[NX], gu[NU], qh[NX], rh[NU];
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) x[r] = valid ? ld(Xc, j, NX, r) : 0.0;
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        u[t] = hom ? ((j * NU + t == ek) ? 1.0 : 0.0) : (uc ? uc[j * NU + t] : 0.0);
        up[t] = (j >= 1) ? (hom ? (((j - 1) * NU + t == ek) ? 1.0 : 0.0) : (uc ? uc[(j - 1) * NU + t] : 0.0)) : 0.0;
        un[t] = (j < Nc - 1) ? (hom ? (((j + 1) * NU + t == ek) ? 1.0 : 0.0) : (uc ? uc[(j + 1) * NU + t] : 0.0)) : 0.0;
      }
      if (!hom) {
        lin_terms(mode, j, rho_u, rho_x, qh, rh);
      } else {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) qh[e] = 0.0;
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) rh[e] = 0.0;
      }
      // an = (Qs + rho_x I) x + qh ; gu = Rs u + rh
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = qh[r] + rho_x * x[r];
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += ld(pb.Qs, j, NQ, r <= t ? pk(r, t, NX) : pk(t, r, NX)) * x[t];
        an[r] = acc;
      }
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = rh[b];
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += ld(pb.Rs, j, NR, b <= t ? pk(b, t, NU) : pk(t, b, NU)) * u[t];
        gu[b] = acc;
      }
      if (j == Nc - 1) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) {
          double acc = hom ? 0.0 : pc[r];
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) acc += F.Pc[(size_t)(r * NX + t) * P + g] * x[t];
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += F.Wc[(size_t)(r * NU + t) * P + g] * u[t];
          an[r] += acc;
        }
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) {
          double acc = hom ? 0.0 : zc[b];
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) acc += F.Wc[(size_t)(r * NU + b) * P + g] * x[r];
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += F.Zc[(size_t)(b * NU + t) * P + g] * u[t];
          gu[b] += acc;
        }
      } else {
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) {
          double acc = 0.0;
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) acc += ld(pb.A, j + 1, NX * NX, r * NX + t) * a[r];
          an[t] += acc;
        }
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) gu[b] += s * (u[b] - un[b]);
      }
      if (j >= 1) {
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) gu[b] += s * (u[b] - up[b]);
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) a[r] = an[r];
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = gu[b];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) acc += ld(pb.Bm, j, NX * NU, r * NU + b) * a[r];
        if (out != nullptr) out[j * NU + b] = valid ? acc : 0.0;  // caller reduces with cta_reduce
        else sc[(size_t)(row0 + j * NU + b) * T + tid] = valid ? acc : 0.0;
      }
    }
  }

  // ---- K3: assemble and invert the shared-block system (cooperative, in global scratch) -------
  __device__ __forceinline__ void build_shared_system(int mode, double rho_u, double rho_x) {
    assemble_shared(mode, rho_u, rho_x);
    factor_shared(mode);
  }

  // Hc = sum_i H_i (+ rho on the diagonal in ADMM mode), symmetrised
  __device__ __noinline__ void assemble_shared(int mode, double rho_u, double rho_x) {
    if (!GROUPED || Nc < 1) return;
    const FactorSet& F = pb.fs[mode];
    double* Hc = F.Hc + (size_t)prob * D * D;
    double* Hi = F.Hci + (size_t)prob * D * D;
    // columns of sum_i H_i; scratch holds D rows per column -> process in column chunks that fit
    const int rows_cap = (int)(pb.scratch_stride / T);
    const int cols_per_pass = max(1, rows_cap / D);
    for (int k0 = 0; k0 < D; k0 += cols_per_pass) {
      const int kc = min(cols_per_pass, D - k0);
      for (int k = 0; k < kc; ++k) shared_grad(mode, true, k0 + k, nullptr, rho_u, rho_x, nullptr, nullptr, k * D);
      __syncthreads();
      // reduce kc*D rows; row (k*D + l) -> Hc[l][k0+k]
      const int lane = tid & 31, w = tid >> 5, nw = T >> 5;
      const double* sc = scr();
      for (int idx = w; idx < kc * D; idx += nw) {
        double acc = 0.0;
        for (int t = lane; t < T; t += 32) acc += sc[(size_t)idx * T + t];
        acc = warp_sum(acc);
        if (lane == 0) Hc[(size_t)(idx % D) * D + (k0 + idx / D)] = acc;
      }
      __syncthreads();
    }
    xreduce(Hc, D * D, true);  // particle sharding: sum over the ranks
    // symmetrise, add the ADMM penalty of the shared block
    for (int idx = tid; idx < D * D; idx += T) {
      const int a = idx / D, b = idx % D;
      if (a <= b) {
        double v = 0.5 * (Hc[(size_t)a * D + b] + Hc[(size_t)b * D + a]);
        if (a == b && mode == MODE_ADMM && pb.has_ubox) v += rho_u;
        Hi[(size_t)a * D + b] = v;  // stage symmetric copy in Hi
      }
    }
    __syncthreads();
    for (int idx = tid; idx < D * D; idx += T) {
      const int a = idx / D, b = idx % D;
      Hc[idx] = (a <= b) ? Hi[(size_t)a * D + b] : Hi[(size_t)b * D + a];
    }
    __syncthreads();
  }

  // masked inverse of Hc for the current shared active set (polish) and hb = Hc * beta
  // use_mask = false: plain inverse of Hc (interior-point fallback: no component is fixed)
  __device__ __noinline__ void factor_shared(int mode, bool use_mask = true) {
    if (!GROUPED || Nc < 1) return;
    const FactorSet& F = pb.fs[mode];
    double* Hc = F.Hc + (size_t)prob * D * D;
    double* Hi = F.Hci + (size_t)prob * D * D;
    double* hb = F.hb + (size_t)prob * D;
    const signed char* actc = pb.actc + (size_t)prob * D;
    // hb = Hc * beta (beta = bound value on fixed comps) ; then masked matrix into Hi
    const bool masked = (mode == MODE_POLISH && pb.has_ubox && use_mask);
    for (int a = tid; a < D; a += T) {
      double acc = 0.0;
      if (masked) {
        for (int b = 0; b < D; ++b) {
          const signed char ac = actc[b];
          if (ac != 0) {
            const int j = b / NU, t = b % NU;
            const size_t g0 = (size_t)prob * pb.Mpad;  // particle 0 of the problem owns the shared bounds
            const double beta = (ac > 0) ? pb.uu[ixg(j, NU, t, g0)] : pb.lu[ixg(j, NU, t, g0)];
            acc += Hc[(size_t)a * D + b] * beta;
          }
        }
      }
      hb[a] = acc;
    }
    for (int idx = tid; idx < D * D; idx += T) {
      const int a = idx / D, b = idx % D;
      const bool fx = masked && (actc[a] != 0 || actc[b] != 0);
      Hi[idx] = fx ? (a == b ? 1.0 : 0.0) : Hc[idx];
    }
    __syncthreads();
    // cooperative Cholesky of Hi (lower), then explicit inverse column per thread
    for (int k = 0; k < D; ++k) {
      if (tid == 0) Hi[(size_t)k * D + k] = sqrt(Hi[(size_t)k * D + k]);
      __syncthreads();
      const double dk = Hi[(size_t)k * D + k];
      for (int i = k + 1 + tid; i < D; i += T) Hi[(size_t)i * D + k] /= dk;
      __syncthreads();
      const int cnt = D - k - 1;
      for (int idx = tid; idx < cnt * cnt; idx += T) {
        const int i = k + 1 + idx / cnt, j = k + 1 + idx % cnt;
        if (j <= i) Hi[(size_t)i * D + j] -= Hi[(size_t)i * D + k] * Hi[(size_t)j * D + k];
      }
      __syncthreads();
    }
    // inverse: column t of the inverse goes to the strictly-upper part + a side buffer is awkward;
    // use scratch as a D x D staging area (scratch_stride >= D*D is guaranteed by the host).
    double* inv = scr();
    for (int t = tid; t < D; t += T) {
      for (int i = 0; i < D; ++i) {
        double v = (i == t) ? 1.0 : 0.0;
        for (int l = 0; l < i; ++l) v -= Hi[(size_t)i * D + l] * inv[(size_t)l * D + t];
        inv[(size_t)i * D + t] = v / Hi[(size_t)i * D + i];
      }
      for (int i = D - 1; i >= 0; --i) {
        double v = inv[(size_t)i * D + t];
        for (int l = i + 1; l < D; ++l) v -= Hi[(size_t)l * D + i] * inv[(size_t)l * D + t];
        inv[(size_t)i * D + t] = v / Hi[(size_t)i * D + i];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < D * D; idx += T) {
      const int a = idx / D, b = idx % D;
      const bool fx = masked && (actc[a] != 0 || actc[b] != 0);
      Hi[idx] = fx ? 0.0 : inv[idx];
    }
    __syncthreads();
  }

  // solve for the shared controls: uc = beta - Hi (gc + hb [+ admm terms])
  __device__ __forceinline__ void solve_shared(int mode, double rho_u, double rho_x, const double* pc, const double* zc) {
    grad_shared(mode, rho_u, rho_x, pc, zc);
    apply_shared(mode, rho_u);
  }

  // gc = sum_i gradient of particle i's cost w.r.t. the shared controls at uc = 0
  __device__ __noinline__ void grad_shared(int mode, double rho_u, double rho_x, const double* pc, const double* zc) {
    if (!GROUPED || Nc < 1) return;
    shared_grad(mode, false, -1, nullptr, rho_u, rho_x, pc, zc, 0);
    reduce(D, pb.gc + (size_t)prob * D);
    xreduce(pb.gc + (size_t)prob * D, D, true);
  }

  __device__ __noinline__ void apply_shared(int mode, double rho_u, bool use_mask = true) {
    if (!GROUPED || Nc < 1) return;
    const FactorSet& F = pb.fs[mode];
    const double* gc = pb.gc + (size_t)prob * D;
    double* uc = pb.uc + (size_t)prob * D;
    const double* Hi = F.Hci + (size_t)prob * D * D;
    const double* hb = F.hb + (size_t)prob * D;
    const signed char* actc = pb.actc + (size_t)prob * D;
    const size_t g0 = (size_t)prob * pb.Mpad;
    for (int a = tid; a < D; a += T) {
      double acc = 0.0;
      const int ja = a / NU, ta = a % NU;
      if (mode == MODE_POLISH && pb.has_ubox && use_mask && actc[a] != 0)
        acc = (actc[a] > 0) ? pb.uu[ixg(ja, NU, ta, g0)] : pb.lu[ixg(ja, NU, ta, g0)];
      for (int b = 0; b < D; ++b) {
        double gb = gc[b] + hb[b];
        if (mode == MODE_ADMM && pb.has_ubox) {
          const int j = b / NU, t = b % NU;
          gb += -rho_u * pb.wu[ixg(j, NU, t, g0)] + pb.yu[ixg(j, NU, t, g0)];
        }
        acc -= Hi[(size_t)a * D + b] * gb;
      }
      uc[a] = acc;
    }
    __syncthreads();
  }

  // ---- K4b/K5: forward rollout (+ fused projection / dual update in ADMM mode) ---------------
  __device__ __noinline__ void forward(int mode, double rho, double* st) {
    const FactorSet& F = pb.fs[mode];
    const double* uc = (GROUPED && Nc >= 1) ? pb.uc + (size_t)prob * D : nullptr;
    const double alpha = pb.alpha;
    double x[NX], up[NU];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) x[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) up[e] = 0.0;
    double rp = 0.0, rd = 0.0, sp = 0.0, sd = 0.0;
    for (int j = 0; j < N; ++j) {
      double u[NU], xn[NX];
      if (j < Nc) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) u[a] = uc[j * NU + a];
      } else {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = ld(pb.kff, j, NU, a);
          if (j >= 1) {
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) acc += ld(F.K, j, NU * NX, a * NX + t) * x[t];
            if (pb.has_slew) {
#pragma unroll UNR
              for (int b = 0; b < NU; ++b) acc += s * ld(F.E, j, NR, a <= b ? pk(a, b, NU) : pk(b, a, NU)) * up[b];
            }
          }
          u[a] = acc;
        }
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = ld(pb.c, j, NX, r);
        if (j >= 1) {
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) acc += ld(pb.A, j, NX * NX, r * NX + t) * x[t];
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += ld(pb.Bm, j, NX * NU, r * NU + t) * u[t];
        xn[r] = acc;
      }
      if (valid) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) at(Uc, j, NU, a) = u[a];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) at(Xc, j, NX, r) = xn[r];
      }
      if (mode == MODE_ADMM) {
        if (pb.has_ubox) {
          // shared stages: one multiplier per problem, owned by particle 0's slot / thread 0
          const size_t gs = (j < Nc) ? (size_t)prob * pb.Mpad : (size_t)g;
          if (j >= Nc || tid == 0) {
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) {
            const size_t o = ixg(j, NU, a, gs);
            const double w = pb.wu[o], y = pb.yu[o];
            const double zr = alpha * u[a] + (1.0 - alpha) * w;
            const double lo = pb.lu[o], hi = pb.uu[o];
            const double wn = fmin(fmax(zr + y / rho, lo), hi);
            const double yn = y + rho * (zr - wn);
            rp = fmax(rp, fabs(u[a] - wn));
            rd = fmax(rd, rho * fabs(wn - w));
            sp = fmax(sp, fmax(fabs(u[a]), fabs(wn)));
            sd = fmax(sd, fabs(yn));
            if (valid) {
              pb.wu[o] = wn;
              pb.yu[o] = yn;
            }
          }
          }
        }
        if (pb.has_xbox) {
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) {
            const size_t o = ix(j, NX, r);
            const double w = pb.wx[o], y = pb.yx[o];
            const double zr = alpha * xn[r] + (1.0 - alpha) * w;
            const double wn = fmin(fmax(zr + y / rho, pb.lx[o]), pb.ux[o]);
            const double yn = y + rho * (zr - wn);
            rp = fmax(rp, fabs(xn[r] - wn));
            rd = fmax(rd, rho * fabs(wn - w));
            sp = fmax(sp, fmax(fabs(xn[r]), fabs(wn)));
            sd = fmax(sd, fabs(yn));
            if (valid) {
              pb.wx[o] = wn;
              pb.yx[o] = yn;
            }
          }
        }
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) x[r] = xn[r];
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) up[a] = u[a];
    }
    st[ST_RP] = rp; st[ST_RD] = rd; st[ST_SP] = sp; st[ST_SD] = sd;
  }

  // ---- KKT check of (X, U) by an adjoint sweep with the true cost; proposes the next active set
  // (primal-dual active-set update) and the multipliers for the ADMM warm start ---------------
  __device__ __noinline__ void kkt_sweep(double* st) {
    double a[NX], un[NU];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) a[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) un[e] = 0.0;
    double kkt = 0.0, uviol = 0.0, xviol = 0.0, changed = 0.0, gscale = 1.0, zscale = 1.0, nanf = 0.0;
    double* sc = scr();
    // pass 1: scales (needed for the relative "at bound" test)
    for (int j = N - 1; j >= 0; --j) {
      double x[NX], u[NU], up[NU], an[NX];
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) x[r] = ld(Xc, j, NX, r);
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        u[t] = ld(Uc, j, NU, t);
        up[t] = (j >= 1) ? ld(Uc, j - 1, NU, t) : 0.0;
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        const double qv = ld(pb.q, j, NX, r);
        double acc = qv;
        gscale = fmax(gscale, fabs(qv));
        zscale = fmax(zscale, fabs(x[r]));
        if (!(x[r] == x[r])) nanf = 1.0;
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += ld(pb.Qs, j, NQ, r <= t ? pk(r, t, NX) : pk(t, r, NX)) * x[t];
        an[r] = acc;
      }
      if (j < N - 1) {
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) {
          double acc = 0.0;
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) acc += ld(pb.A, j + 1, NX * NX, r * NX + t) * a[r];
          an[t] += acc;
        }
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) a[r] = an[r];
      if (pb.has_xbox) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r)
          xviol = fmax(xviol, fmax(x[r] - ld(pb.ux, j, NX, r), ld(pb.lx, j, NX, r) - x[r]));
      }
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        const double rv = ld(pb.r, j, NU, b);
        double acc = rv;
        gscale = fmax(gscale, fabs(rv));
        zscale = fmax(zscale, fabs(u[b]));
        if (!(u[b] == u[b])) nanf = 1.0;
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += ld(pb.Rs, j, NR, b <= t ? pk(b, t, NU) : pk(t, b, NU)) * u[t];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) acc += ld(pb.Bm, j, NX * NU, r * NU + b) * a[r];
        if (j >= 1) acc += s * (u[b] - up[b]);
        if (j < N - 1) acc += s * (u[b] - un[b]);
        // acc = reduced gradient dJ/du_{j,b}
        if (j < Nc) {
          sc[(size_t)(j * NU + b) * T + tid] = valid ? acc : 0.0;  // summed over particles below
        } else {
          const double mu = -acc;
          double lo = -INFINITY, hi = INFINITY;
          if (pb.has_ubox) {
            lo = ld(pb.lu, j, NU, b);
            hi = ld(pb.uu, j, NU, b);
          }
          const double ub = u[b];
          uviol = fmax(uviol, fmax(ub - hi, lo - ub));
          const bool at_hi = ub >= hi - 1e-12 * fmax(1.0, fabs(hi));
          const bool at_lo = ub <= lo + 1e-12 * fmax(1.0, fabs(lo));
          double v;
          if (at_hi && at_lo) v = 0.0;
          else if (at_hi) v = fmax(acc, 0.0);
          else if (at_lo) v = fmax(-acc, 0.0);
          else v = fabs(acc);
          kkt = fmax(kkt, v);
          if (pb.has_ubox) {
            signed char na = 0;
            if (mu + (ub - hi) > 0.0 && hi < INFINITY) na = 1;
            else if (mu + (ub - lo) < 0.0 && lo > -INFINITY) na = -1;
            const size_t o = ix(j, NU, b);
            if (valid) {
              if (na != pb.act[o]) changed += 1.0;
              pb.actn[o] = na;
              pb.kff[o] = mu;  // kff is dead after the forward pass: stash the multiplier estimate
            }
          }
        }
      }
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) un[t] = u[t];
    }
    st[ST_KKT] = kkt; st[ST_UVIOL] = uviol; st[ST_XVIOL] = xviol; st[ST_CHANGED] = changed;
    st[ST_GSCALE] = gscale; st[ST_ZSCALE] = zscale; st[ST_NAN] = nanf;
    if (GROUPED && Nc >= 1) {
      double* muc = pb.muc + (size_t)prob * D;
      reduce(D, muc);  // muc = sum_i reduced gradient on the shared block
      xreduce(muc, D, true);
      const size_t g0 = (size_t)prob * pb.Mpad;
      const double* uc = pb.uc + (size_t)prob * D;
      const signed char* actc = pb.actc + (size_t)prob * D;
      signed char* actcn = pb.actcn + (size_t)prob * D;
      for (int k = tid; k < D; k += T) {
        const int j = k / NU, b = k % NU;
        const double acc = muc[k], mu = -acc, ub = uc[k];
        double lo = -INFINITY, hi = INFINITY;
        if (pb.has_ubox) {
          lo = pb.lu[ixg(j, NU, b, g0)];
          hi = pb.uu[ixg(j, NU, b, g0)];
        }
        st[ST_UVIOL] = fmax(st[ST_UVIOL], fmax(ub - hi, lo - ub));
        const bool at_hi = ub >= hi - 1e-12 * fmax(1.0, fabs(hi));
        const bool at_lo = ub <= lo + 1e-12 * fmax(1.0, fabs(lo));
        double v;
        if (at_hi && at_lo) v = 0.0;
        else if (at_hi) v = fmax(acc, 0.0);
        else if (at_lo) v = fmax(-acc, 0.0);
        else v = fabs(acc);
        st[ST_KKT] = fmax(st[ST_KKT], v);
        signed char na = 0;
        if (pb.has_ubox) {
          if (mu + (ub - hi) > 0.0 && hi < INFINITY) na = 1;
          else if (mu + (ub - lo) < 0.0 && lo > -INFINITY) na = -1;
        }
        if (na != actc[k]) st[ST_CHANGED] += 1.0;
        actcn[k] = na;
      }
    }
  }

  // after an accepted polish: make (w, y, act) consistent with the solution for later warm starts
  __device__ __noinline__ void commit_polish() {
    if (!pb.has_ubox) return;
    const double* muc = (GROUPED && Nc >= 1) ? pb.muc + (size_t)prob * D : nullptr;
    // recompute multipliers cheaply: y = 0 on free comps; on active comps the ADMM warm start only
    // needs the sign/magnitude, taken from the stored proposal (actn == act here).
    for (int j = 0; j < N; ++j) {
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        const size_t o = ix(j, NU, b);
        const double u = ld(Uc, j, NU, b);
        if (valid && (j >= Nc || tid == 0)) {
          pb.wu[o] = fmin(fmax(u, pb.lu[o]), pb.uu[o]);
          if (j < Nc) pb.yu[o] = (pb.actc[(size_t)prob * D + j * NU + b] != 0) ? -muc[j * NU + b] : 0.0;
          else pb.yu[o] = (pb.act[o] != 0) ? pb.kff[o] : 0.0;
        }
      }
    }
  }

  // ---- SCP bookkeeping (scp_mpc.py:397-407): residual, objective, X_prev <- X ------------------
  __device__ __noinline__ void scp_bookkeeping(double* st) {
    double dxm = 0.0, dum = 0.0, obj = 0.0;
    for (int j = 0; j < N; ++j) {
      double x[NX], u[NU], dx[NX], du[NU];
      double n2 = 0.0;
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        x[r] = ld(Xc, j, NX, r);
        const double d = x[r] - ld(Xpv, j, NX, r);
        n2 += d * d;
        dx[r] = x[r] - ld(pb.xref, j, NX, r);
      }
      dxm = fmax(dxm, sqrt(n2));
      n2 = 0.0;
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        u[t] = ld(Uc, j, NU, t);
        const double d = u[t] - ld(Upv, j, NU, t);
        n2 += d * d;
        du[t] = u[t] - ld(pb.uref, j, NU, t);
      }
      dum = fmax(dum, sqrt(n2));
      // obj += dx'Q dx + du'R du with the un-regularised Q, R (Qs - reg_x I, Rs - diag terms)
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += ld(pb.Qs, j, NQ, r <= t ? pk(r, t, NX) : pk(t, r, NX)) * dx[t];
        obj += dx[r] * (acc - pb.reg_x * dx[r]);
      }
      const double rdiag = pb.rdiag0[g];
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += ld(pb.Rs, j, NR, b <= t ? pk(b, t, NU) : pk(t, b, NU)) * du[t];
        obj += du[b] * (acc - (pb.reg_u + (j == 0 ? rdiag : 0.0)) * du[b]);
      }
      if (valid && !pb.keep_prev) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) at(Xpv, j, NX, r) = x[r];
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) at(Upv, j, NU, t) = u[t];
      }
    }
    st[ST_DX] = valid ? dxm : 0.0;
    st[ST_DU] = valid ? dum : 0.0;
    st[ST_OBJ] = valid ? obj : 0.0;
  }

  // reduce the per-thread statistics over the problem; result replicated in st[]
  __device__ __noinline__ void reduce_stats(double* st) {
    if (!GROUPED) return;
    if (pb.world <= 1) {  // one GPU: shuffles + shared memory (cta_reduce); ST_OBJ is summed, the rest max-reduced
      cta_reduce(st, ST_COUNT, 1u << ST_OBJ);
      return;
    }
    double* sc = scr();
    for (int k = 0; k < ST_COUNT; ++k) sc[(size_t)k * T + tid] = st[k];
    reduce(ST_COUNT, red, ST_OBJ);
    xreduce(red, ST_COUNT, false, 1u << ST_OBJ);
    for (int k = 0; k < ST_COUNT; ++k) st[k] = red[k];
    __syncthreads();
  }

  // active-set guess from the ADMM variables (OSQP polish rule: lower active if w - l < -y ...)
  __device__ __noinline__ double guess_active_from_admm() {
    double changed = 0.0;
    if (!pb.has_ubox) return 0.0;
    for (int j = Nc; j < N; ++j) {
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        const size_t o = ix(j, NU, b);
        const double w = pb.wu[o], y = pb.yu[o];
        signed char na = 0;
        if (pb.uu[o] - w < y) na = 1;
        else if (w - pb.lu[o] < -y) na = -1;
        if (valid) {
          if (pb.act[o] != na) changed += 1.0;
          pb.act[o] = na;
        }
      }
    }
    if (GROUPED && Nc >= 1) {
      const size_t g0 = (size_t)prob * pb.Mpad;
      signed char* actc = pb.actc + (size_t)prob * D;
      __syncthreads();
      for (int k = tid; k < D; k += T) {
        const size_t o = ixg(k / NU, NU, k % NU, g0);
        const double w = pb.wu[o], y = pb.yu[o];
        signed char na = 0;
        if (pb.uu[o] - w < y) na = 1;
        else if (w - pb.lu[o] < -y) na = -1;
        if (actc[k] != na) changed += 1.0;
        actc[k] = na;
      }
      __syncthreads();
    }
    return changed;
  }

  __device__ __noinline__ void adopt_proposed_active_set() {
    if (!pb.has_ubox) return;
    for (int j = Nc; j < N; ++j) {
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        const size_t o = ix(j, NU, b);
        if (valid) pb.act[o] = pb.actn[o];
      }
    }
    if (GROUPED && Nc >= 1) {
      __syncthreads();
      for (int k = tid; k < D; k += T) pb.actc[(size_t)prob * D + k] = pb.actcn[(size_t)prob * D + k];
      __syncthreads();
    }
  }

#include "fused.inl"

  // ---- one convex sub-problem: polish attempts globalised by ADMM blocks ----------------------
  // returns status bits; fills st with the statistics of the accepted solution
  __device__ __noinline__ int solve_qp(double* st, long long* cnt) {
    double rho = pb.rho[prob];
    double pc[NX], zc[NU];
    const bool constrained = pb.has_ubox || pb.has_xbox;
    int admm_it = 0, status = 0;
    int mode = (pb.polish || !constrained) ? MODE_POLISH : MODE_ADMM;
    int polish_try = 0, next_polish_block = 0, block = 0;
    bool admm_factored = false;
    double xact_hint = 0.0;
    for (;;) {
      if (mode == MODE_POLISH) {
        factor(MODE_POLISH, 0.0, 0.0);
        build_shared_system(MODE_POLISH, 0.0, 0.0);
        backward(MODE_POLISH, 0.0, 0.0, pc, zc);
        solve_shared(MODE_POLISH, 0.0, 0.0, pc, zc);
        forward(MODE_POLISH, 1.0, st);
        kkt_sweep(st);
        reduce_stats(st);
        cnt[2] += 1;
        cnt[3] += 1;
        cnt[7] += 4;
        const bool nan = st[ST_NAN] > 0.0;
        const double tol_g = pb.kkt_tol * st[ST_GSCALE], tol_z = pb.kkt_tol * st[ST_ZSCALE];
        const bool ok = !nan && st[ST_KKT] <= tol_g && st[ST_UVIOL] <= tol_z && st[ST_XVIOL] <= tol_z;
        if (ok || !constrained) {
          adopt_proposed_active_set();
          commit_polish();
          cnt[5] += 1;
          if (nan) status |= 1;
          break;
        }
        if (nan) { status |= 1; break; }
        ++polish_try;
        if (st[ST_XVIOL] > tol_z) {
          mode = MODE_ADMM;  // a state bound wants to be active: not representable in the polish
          xact_hint = 1.0;
        } else if (st[ST_CHANGED] > 0.0 && polish_try < pb.polish_iters) {
          adopt_proposed_active_set();
        } else {
          mode = MODE_ADMM;
        }
        if (mode == MODE_ADMM) {
          next_polish_block = block + (1 << min(polish_try / max(pb.polish_iters, 1), 6));
          polish_try = 0;
        }
        continue;
      }
      // ---------------- ADMM block ----------------
      const double rho_u = pb.has_ubox ? rho : 0.0, rho_x = pb.has_xbox ? rho : 0.0;
      if (!admm_factored) {
        factor(MODE_ADMM, rho_u, rho_x);
        build_shared_system(MODE_ADMM, rho_u, rho_x);
        admm_factored = true;
        cnt[2] += 1;
        cnt[7] += 1;
      }
      for (int k = 0; k < pb.check_every; ++k) {
        backward(MODE_ADMM, rho_u, rho_x, pc, zc);
        solve_shared(MODE_ADMM, rho_u, rho_x, pc, zc);
        forward(MODE_ADMM, rho, st);
        ++admm_it;
      }
      cnt[1] += pb.check_every;
      cnt[7] += 2 * pb.check_every;
      ++block;
      // x-activity indicator: any nonzero state-box multiplier
      st[ST_KKT] = st[ST_UVIOL] = st[ST_XVIOL] = st[ST_CHANGED] = st[ST_NAN] = 0.0;
      st[ST_GSCALE] = st[ST_ZSCALE] = 1.0;
      double xact = 0.0;
      if (pb.has_xbox) {
        for (int j = 0; j < N; ++j)
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) xact = fmax(xact, fabs(ld(pb.yx, j, NX, r)));
      }
      st[ST_XACT] = valid ? xact : 0.0;
      // gradient scale for the dual tolerance
      double gs = 0.0;
      for (int j = 0; j < N; ++j) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) gs = fmax(gs, fabs(ld(pb.q, j, NX, r)));
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) gs = fmax(gs, fabs(ld(pb.r, j, NU, b)));
      }
      st[ST_GSCALE] = gs;
      if (!(st[ST_RP] == st[ST_RP]) || !(st[ST_RD] == st[ST_RD])) st[ST_NAN] = 1.0;
      reduce_stats(st);
      if (st[ST_NAN] > 0.0) { status |= 1; break; }
      const double sp = fmax(st[ST_SP], 1e-12), sd = fmax(fmax(st[ST_SD], st[ST_GSCALE]), 1e-12);
      if (st[ST_RP] <= pb.eps_abs + pb.eps_rel * sp && st[ST_RD] <= pb.eps_abs + pb.eps_rel * sd) {
        guess_active_from_admm();
        cnt[6] += 1;
        break;
      }
      if (admm_it >= pb.max_admm_iter) {
        guess_active_from_admm();
        status |= 2;
        cnt[6] += 1;
        break;
      }
      if (pb.adapt_rho) {
        double est = rho * sqrt((st[ST_RP] / sp) / fmax(st[ST_RD] / sd, 1e-30));
        est = fmin(fmax(est, 1e-6), 1e6);
        if (est > rho * pb.adapt_tol || est * pb.adapt_tol < rho) {
          rho = est;
          admm_factored = false;
        }
      }
      if (pb.polish && st[ST_XACT] == 0.0 && block >= next_polish_block) {
        guess_active_from_admm();
        mode = MODE_POLISH;
        polish_try = 0;
      }
    }
    pb.rho[prob] = rho;
    (void)xact_hint;
    return status;
  }

  // ---- the SCP loop (scp_mpc.py:337-428) ----------------------------------------------------
  // kp: the kernel's own __grid_constant__ parameter.  The fused path reads it directly (constant bank
  // operands) instead of through the `pb` member, which the noinline helpers force into memory.
  template <int VAR>
  __device__ __forceinline__ void run(const DevProb& kp, unsigned char* smem) {
    fused_init(smem);
    flip = false;
    // Thread-per-problem launch: lanes beyond the last problem have nothing to mirror for — every decision
    // (active-set stall, interior-point / ADMM fallback, stop rule) is taken per thread there, and a lane whose
    // stores are masked never sees its own progress: it used to run the ADMM fallback to max_iter beside lane 0
    // (C1, B = 1: 230 ms per SCP iteration instead of 0.5 ms).  No warp-collective is executed on this path.
    if (dead || (!GROUPED && !valid)) return;
    double st[ST_COUNT];
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool writer = valid && (GROUPED ? tid == 0 : true);
    // the host launches the kernel variant that matches (dyn_kind, has_slew): see launch_scp_impl
    const bool fused_ok = pb.fused && pb.polish && ((VAR & 1) != 0) == (pb.dyn_kind == DYN_DUBINS) &&
                          !((VAR & 2) && pb.has_slew);
    double *solX = Xc, *solU = Uc;
    int it = 0;
    if (pb.resume) {
      // launched behind the warp-per-particle kernel: continue only the problems it handed over, at their iteration
      if (!valid || !pb.fb[prob]) {
        if (!GROUPED || !pb.fb[prob]) return;
      }
      for (int k = 0; k < 8; ++k) cnt[k] = pb.info[(size_t)k * pb.B + prob];
      it = (int)cnt[0];
    }
    for (; it < pb.max_it; ++it) {
      const long long admm0 = cnt[1];
      int status = -1;
      bool fused_done = false;
      if (fused_ok) {
        status = fused_qp<VAR>(kp, smem, st, cnt);
        fused_done = status >= 0;
      }
      if (!fused_done) {
        linearize_prep();
        status = solve_qp(st, cnt);
        cnt[7] += 2;
      }
      const double kkt_rel = st[ST_KKT] / fmax(st[ST_GSCALE], 1.0);
      solX = Xc;
      solU = Uc;
      if (status & 1) {
        cnt[4] |= 1;
        // reference convention: NaN output on failure (osqp_solver.jl:69-71)
        for (int j = 0; j < N; ++j) {
          if (valid) {
#pragma unroll UNR
            for (int r = 0; r < NX; ++r) at(Xc, j, NX, r) = NAN;
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) at(Uc, j, NU, t) = NAN;
          }
        }
        ++it;
        break;
      }
      if (status & 2) cnt[4] |= 2;
      if (xdead) {  // a peer never arrived at an exchange: abandon the launch (pmpc_sync reports PMPC_ERR_STATE)
        cnt[4] |= 4;
        ++it;
        break;
      }
      if (fused_done) {
        // residual / objective were accumulated by the forward sweep; X_prev <- X is a buffer swap
        if (!pb.keep_prev) {
          double* t0 = Xc; Xc = Xpv; Xpv = t0;
          double* t1 = Uc; Uc = Upv; Upv = t1;
          flip = !flip;
        }
      } else {
        scp_bookkeeping(st);
        reduce_stats(st);
      }
      const double max_res = fmax(st[ST_DX], st[ST_DU]);
      const double obj = st[ST_OBJ] / ((double)N * (double)(pb.world > 1 ? pb.M_total : pb.M));
      const int hrow = it + pb.hist_off;
      if (writer && hrow < pb.hist_len) {
        double* h = pb.hist;
        const size_t Bn = (size_t)pb.B;
        h[((size_t)0 * pb.hist_len + hrow) * Bn + prob] = obj;
        h[((size_t)1 * pb.hist_len + hrow) * Bn + prob] = max_res;
        h[((size_t)2 * pb.hist_len + hrow) * Bn + prob] = (double)(cnt[1] - admm0);
        h[((size_t)3 * pb.hist_len + hrow) * Bn + prob] = kkt_rel;
      }
      if (max_res < pb.res_tol) {
        ++it;
        break;
      }
    }
    // leave the last solution in pb.X / pb.U and, unless keep_prev, in pb.Xp / pb.Up as well
    if (valid) {
      for (int j = 0; j < N; ++j) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) {
          const double v = ld(solX, j, NX, r);
          if (solX != pb.X) at(pb.X, j, NX, r) = v;
          if (!pb.keep_prev && solX != pb.Xp) at(pb.Xp, j, NX, r) = v;
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) {
          const double v = ld(solU, j, NU, t);
          if (solU != pb.U) at(pb.U, j, NU, t) = v;
          if (!pb.keep_prev && solU != pb.Up) at(pb.Up, j, NU, t) = v;
        }
      }
    }
    xclose();
    cnt[0] = it;
    if (writer) {
      for (int k = 0; k < 8; ++k) pb.info[(size_t)k * pb.B + prob] = cnt[k];
    }
  }

  // ---- the SCP loop with two problems per warp (PPW == 2; see fused_qp_half) --------------------------------
  template <int VAR>
  __device__ __forceinline__ void run_half(const DevProb& kp, unsigned char* smem, bool first) {
    constexpr bool DUB = (VAR & 1) && NX == 4 && NU == 2;
    constexpr bool LEAN = (VAR & 2) != 0;
    fused_init(smem, first);
    flip = false;
    double st[ST_COUNT];
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool writer = pvalid && hl == 0;
    bool active = pvalid, handed = false;
    int its = 0;
    const double *solX = Xc, *solU = Uc, *linX = Xpv, *linU = Upv;
    for (int it = 0; it < pb.max_it; ++it) {
      if (!__any_sync(0xffffffffu, active)) break;  // (an empty warp of the last CTA)
      const int status = fused_qp_half<!LEAN, true, DUB>(kp, smem, st, cnt, active);
      if (active) {
        if (status & 8) {  // continued by the one-problem-per-warp kernel from this iteration's linearisation point
          handed = true;
          active = false;
          linX = Xpv;
          linU = Upv;
        } else {
          solX = Xc;
          solU = Uc;
          its = it + 1;
          if (status & 1) {
            cnt[4] |= 1;
            if (valid) {  // reference convention: NaN output on failure (osqp_solver.jl:69-71)
              for (int j = 0; j < N; ++j) {
#pragma unroll
                for (int r = 0; r < NX; ++r) at(Xc, j, NX, r) = NAN;
#pragma unroll
                for (int t = 0; t < NU; ++t) at(Uc, j, NU, t) = NAN;
              }
            }
            active = false;
          } else {
            const double max_res = fmax(st[ST_DX], st[ST_DU]);
            const int hrow = it + pb.hist_off;
            if (writer && hrow < pb.hist_len) {
              const size_t Bn = (size_t)pb.B;
              pb.hist[((size_t)0 * pb.hist_len + hrow) * Bn + prob] = st[ST_OBJ] / ((double)N * (double)pb.M);
              pb.hist[((size_t)1 * pb.hist_len + hrow) * Bn + prob] = max_res;
              pb.hist[((size_t)2 * pb.hist_len + hrow) * Bn + prob] = 0.0;
              pb.hist[((size_t)3 * pb.hist_len + hrow) * Bn + prob] = st[ST_KKT] / fmax(st[ST_GSCALE], 1.0);
            }
            if (max_res < pb.res_tol) active = false;
          }
        }
      }
      if (!pb.keep_prev) {  // X_prev <- X is a buffer swap, in lockstep for both problems (a frozen one stores nothing)
        double* t0 = Xc; Xc = Xpv; Xpv = t0;
        double* t1 = Uc; Uc = Upv; Upv = t1;
        flip = !flip;
      }
      if (!__any_sync(0xffffffffu, active)) break;
    }
    // leave the solution in pb.X / pb.U (and pb.Xp / pb.Up unless keep_prev); a handed-over problem only needs its
    // linearisation point in pb.Xp / pb.Up
    if (valid) {
      const double* sx = handed ? linX : solX;
      const double* su = handed ? linU : solU;
      for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          const double v = sx[ix(j, NX, r)];
          if (!handed && sx != pb.X) at(pb.X, j, NX, r) = v;
          if ((handed || !pb.keep_prev) && sx != pb.Xp) at(pb.Xp, j, NX, r) = v;
        }
#pragma unroll
        for (int t = 0; t < NU; ++t) {
          const double v = su[ix(j, NU, t)];
          if (!handed && su != pb.U) at(pb.U, j, NU, t) = v;
          if ((handed || !pb.keep_prev) && su != pb.Up) at(pb.Up, j, NU, t) = v;
        }
      }
    }
    cnt[0] = its;
    if (writer) {
      for (int k = 0; k < 8; ++k) pb.info[(size_t)k * pb.B + prob] = cnt[k];
      pb.fb[prob] = handed ? 1 : 0;
    }
  }

  // A warp without further work keeps answering the phase barriers of the CTA's other warps (half_phase) until all are done
  __device__ __forceinline__ void half_idle() {
    while (half_phase(0)) __syncthreads();
  }

  // Phase alignment of the warps of a CTA in the two-problems-per-warp kernel.  Every warp solves its own pair of problems
  // independently, but all of them start a backward sweep — and a forward sweep — together: the resident warps of an SM
  // then execute the same loop body, so that one instruction fetch serves them all.  (With eight independent one-warp
  // CTAs per SM at eight different places of a 330 KB kernel the instruction cache — 32 KB — misses all the time:
  // ncu `no_instruction` 2.5 stall cycles per issued instruction.)  All passes have the same length, so nobody waits
  // for long.  Two CTA barriers per pass: start of bsweep (returns whether any warp of the CTA still works) and start of
  // fsweep; everything else in this mode synchronises at warp level.  A warp whose pair is finished takes the next pair
  // from a queue (DevProb::queue, scp_kernel_half) and idles at the barriers only once the queue is empty.
  __device__ __forceinline__ int half_phase(int working) { return __syncthreads_or(working); }
};

}  // namespace pmpc
