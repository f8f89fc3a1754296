// Warp-per-particle SCP kernel for the larger shapes (NX >= 8: BASELINE configs 4 and 5).
//
// north_star: "the KKT system is factored per particle by a warp/CTA-resident fp64 Riccati recursion, with matrices
// staged in shared memory and reductions by warp shuffles; particles are coupled through a Schur complement on the
// shared u[0:Nc] block".  Replaces lqp_repr_Pq / lqp_repr_Ab / lqp_repr_Gla + OSQP (lqp_utils.jl:2-393,
// osqp_solver.jl:34-72) on 12x12 ... 32x32 stage blocks, where the thread-per-particle kernels of solver.cuh keep the
// matrices in per-thread local memory (7 447 LDL vs 4 660 DFMA in the n = 32 SASS) and one problem on ONE SM.
//
// * ONE WARP owns one particle.  Lane l owns column l of the n x n matrices (S = Q + P, A, S A, P') and element l of
//   the vectors; S, A, B, S B + W, G live in the warp's shared-memory slab, the second operand of every product is a
//   broadcast LDS, the accumulators (a whole column) stay in registers.  Small m x m systems (H, its masked inverse)
//   are solved redundantly per lane.
// * The particles of a problem are spread over a TEAM of CTAs (team size 1 for batches of small problems); the
//   consensus terms (sum_i H_i, sum_i g_i: the Schur complement on the shared block) and the pass statistics are
//   combined through global memory with a flag protocol (team_reduce / team_bcast), the team leader solves the
//   shared block and — for particles sharded over GPUs — runs the cross-GPU exchange of solver.cuh (Solver::xreduce).
// * Same algorithm, same hand-off format and same acceptance tests as the fused two-sweep active-set pass of
//   fused.inl (bsweep / fsweep): a QP whose active set is right costs one pass.  What this kernel does not do — a
//   state bound that wants to be active (ADMM), an active-set iteration that cycles (interior point) — is handed to
//   the thread-per-particle kernel of the same shape, which resumes the SCP loop at the failed iteration (DevProb::fb).
//
// Data layout: TILE = 1 (solver.cuh): element e of stage j of particle slot g is arr[(g * N + j) * E + e].
#pragma once
#ifdef PMPC_WPP_DEBUG
#include <cstdio>
#endif
#include <cooperative_groups.h>

#include "solver.cuh"
#include "tma.cuh"

namespace pmpc {

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The kernel's dynamic shared memory, named at namespace scope: every slab pointer is derived from this symbol inside the
// sweeps, so that the accesses compile to LDS / STS (a pointer parked in the object is a generic address: LD.E / ST.E
// through the generic path, measured 2.5x slower in the thread-per-particle kernels).
extern __shared__ __align__(16) unsigned char wpp_smem[];

template <int NX, int NU, bool SLEW>
struct Wpp {
  static constexpr int NQ = NX * (NX + 1) / 2, NR = NU * (NU + 1) / 2, NHO = NR - NU;
  // hand-off block of one (particle, stage), same format as Solver::NH with HL = 0: [K | k | strict upper H]
  static constexpr int H_K = 0, H_KF = NU * NX, H_H = H_KF + NU, NH = H_H + NHO;
  __host__ __device__ static constexpr int pko(int a, int b) { return a * NU - (a * (a + 1)) / 2 + (b - a - 1); }  // a < b
  // Per-warp shared-memory slab (doubles).  The matrix region is used in two ways:
  //   backward sweep: S | A | B | S B + W | W | G          (S, W survive into the consensus step as Pc, Wc)
  //   forward sweep : two ring slots of [A | B | hand-off | E]
  // Every stage block is brought in by bulk asynchronous copies (TMA engine, cp.async.bulk + mbarrier) issued by lane 0
  // as soon as its buffer is free and waited for right before its first use (bsweep: A at stage start for the second
  // half of the stage, Qs and B during the previous stage; fsweep: a two-deep ring).  All sub-buffers start at even
  // offsets (16-byte aligned, as bulk copies require).
  static constexpr int O_S = 0, O_A = O_S + NX * NX, O_B = O_A + NX * NX, O_SBW = O_B + NX * NU, O_W = O_SBW + NX * NU,
                       O_G = O_W + NX * NU, BMAT = O_G + NU * NX;
  static constexpr int F_A = 0, F_B = F_A + NX * NX, F_H = F_B + NX * NU, F_E = F_H + NH, FSLOT = F_E + NR;
  static constexpr int MAT = (BMAT > 2 * FSLOT) ? BMAT : 2 * FSLOT;
  static constexpr int O_Q = MAT;                    // packed Qs of one stage                     (NQ)
  static constexpr int O_V = O_Q + NQ;                // vectors: c, wv (pc), x, a / dx              (4 x NX)
  static constexpr int O_M = O_V + 4 * NX;            // H, E, Z (Zc)                                (3 x NU x NU)
  static constexpr int O_U = O_M + 3 * NU * NU;       // h0, h, kv, ufix, z (zc) | u, up, du         (5 x NU)
  static constexpr int O_BAR = O_U + 6 * NU;          // mbarriers: A, Q, B, A' (bsweep) / slot 0, slot 1, Q (fsweep)
  // small shapes: a second A buffer, so that A of stage j - 1 is requested a whole stage ahead (the stage of an n = 12
  // problem is shorter than an HBM round trip; at n = 32 the first half of the stage covers it and the slab has no room)
  static constexpr bool A2 = (NX <= 16);
  static constexpr int O_A2 = O_BAR + 8;
  static constexpr int SLAB = O_A2 + (A2 ? NX * NX : 0);
  static_assert(NX % 2 == 0 && NQ % 2 == 0 && NR % 2 == 0 && NH % 2 == 0 && SLAB % 2 == 0, "16-byte alignment of the sub-buffers");
  static constexpr int V_C = 0, V_WV = NX, V_X = 2 * NX, V_AD = 3 * NX, V_DX = 3 * NX;
  static constexpr int MM_H = 0, MM_E = NU * NU, MM_Z = 2 * NU * NU;
  static constexpr int U_H0 = 0, U_H = NU, U_KV = 2 * NU, U_UF = 3 * NU, U_Z = 4 * NU, U_U = 0, U_UP = NU, U_DU = 2 * NU;
  enum { BAR_A = 0, BAR_Q = 1, BAR_B = 2, BAR_S0 = 3, BAR_S1 = 4, BAR_FQ = 5, BAR_A2 = 6, NBAR = 7 };
  static constexpr int DMAX = 64;  // largest shared block (Nc * NU) this kernel takes; beyond it the old kernel runs

  // ------------------------------------------------------------------------------------------------------------
  // team exchange through global memory (all threads of the CTA call these)
  struct Team {
    double* part;                 // [Tc][cap]
    double* bc;                   // [cap]
    unsigned long long* pflag;    // [Tc]
    unsigned long long* bflag;    // [1]
    int Tc, idx, cap;
    unsigned long long seq;
    bool dead;
  };

  const DevProb& pb;
  unsigned char* smem;
  const int lane, wid, W, tid, T;
  const int lc;        // lane clamped to the matrix dimension (lanes >= NX mirror lane NX - 1 and never store)
  const bool lact;
  int prob, N, Nc, D, M;
  double *Xc, *Uc;             // iterate being written
  const double *Xpv, *Upv;     // linearisation point
  bool flip;
  double amu = 0.0;            // current penalty of the multiplier method for state bounds (xbox_terms; 0 = off)
  Team tm;
  double* cred;                // CTA exchange area in shared memory behind the slabs (cred_doubles)
  __host__ __device__ static constexpr int cred_doubles(int W_) { return 16 + (W_ + 1) * ST_COUNT; }

  __device__ Wpp(const DevProb& pb_, unsigned char* smem_)
      : pb(pb_), smem(smem_), lane(threadIdx.x & 31), wid(threadIdx.x >> 5), W(blockDim.x >> 5), tid(threadIdx.x),
        T(blockDim.x), lc((threadIdx.x & 31) < NX ? (threadIdx.x & 31) : NX - 1), lact((threadIdx.x & 31) < NX) {}

  __device__ __forceinline__ double* slab() const { return reinterpret_cast<double*>(wpp_smem) + (threadIdx.x >> 5) * SLAB; }
  __device__ __forceinline__ static size_t gix(size_t g, int N_, int j, int E) { return (g * (size_t)N_ + (size_t)j) * (size_t)E; }

  __device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long want) {
    unsigned long long t0 = 0, spins = 0;
    while (ld_acquire_gpu(flag) < want) {
      if ((++spins & 255ull) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > pb.xtimeout_ns) return false;
      }
    }
    return true;
  }

  // Every CTA of the team contributes vec[0..len); the leader (idx 0) receives the combination in out (CTA order:
  // deterministic).  sum_mask bit k (k < 32) set, or all_sum: element k is summed, otherwise max-reduced.
  __device__ __noinline__ void team_reduce(const double* vec, int len, double* out, bool all_sum, unsigned sum_mask) {
    if (tm.Tc == 1) {
      if (out != vec) {
        __syncthreads();
        for (int k = tid; k < len; k += T) out[k] = vec[k];
      }
      __syncthreads();
      return;
    }
    __syncthreads();
    ++tm.seq;
    if (tm.dead) return;
    double* mine = tm.part + (size_t)tm.idx * tm.cap;
    for (int k = tid; k < len; k += T) __stcg(mine + k, vec[k]);
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu(tm.pflag + tm.idx, tm.seq);
    if (tm.idx != 0) return;
    __shared__ int ok_s;
    if (tid == 0) ok_s = 1;
    __syncthreads();
    for (int c = tid; c < tm.Tc; c += T)
      if (!spin_until(tm.pflag + c, tm.seq)) ok_s = 0;
    __syncthreads();
    if (!ok_s) {
      if (tid == 0) *pb.xerr = 1;
      tm.dead = true;
      return;
    }
    for (int k = tid; k < len; k += T) {
      const bool is_sum = all_sum || (k < 32 && ((sum_mask >> k) & 1u));
      double acc = __ldcg(tm.part + k);
      for (int c = 1; c < tm.Tc; ++c) {
        const double v = __ldcg(tm.part + (size_t)c * tm.cap + k);
        acc = is_sum ? acc + v : fmax(acc, v);
      }
      out[k] = acc;
    }
    __syncthreads();
  }

  // leader: vec -> every CTA of the team (vec is a shared-memory or global buffer of each CTA)
  __device__ __noinline__ void team_bcast(double* vec, int len) {
    if (tm.Tc == 1) {
      __syncthreads();
      return;
    }
    __syncthreads();
    if (tm.idx == 0) {
      if (tm.dead) {  // release the others with a poisoned sequence so that they stop as well
        if (tid == 0) {
          __stcg(tm.bc + tm.cap - 1, 1.0);
          __threadfence();
          st_release_gpu(tm.bflag, tm.seq);
        }
        return;
      }
      for (int k = tid; k < len; k += T) __stcg(tm.bc + k, vec[k]);
      if (tid == 0) __stcg(tm.bc + tm.cap - 1, 0.0);
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_gpu(tm.bflag, tm.seq);
      return;
    }
    if (tm.dead) return;
    __shared__ int ok_b;
    if (tid == 0) ok_b = spin_until(tm.bflag, tm.seq) ? 1 : 0;
    __syncthreads();
    if (!ok_b || __ldcg(tm.bc + tm.cap - 1) != 0.0) {
      if (tid == 0 && !ok_b) *pb.xerr = 1;
      tm.dead = true;
      return;
    }
    for (int k = tid; k < len; k += T) vec[k] = __ldcg(tm.bc + k);
    __syncthreads();
  }

  // ------------------------------------------------------------------------------------------------------------
  // Masked inverse E = embed(inv(H[free, free])) of the m x m stage Hessian in shared memory (Hs): every lane
  // factors the masked matrix redundantly (registers, no exchange), lane t < NU solves for column t and stores it.
  // The solves multiply by the stored reciprocals of the diagonal instead of dividing.
  __device__ __forceinline__ static void masked_inverse(const double* Hs, unsigned fmask, int t, double* Es) {
    double L[NU][NU], idg[NU];  // lower triangle only; plain 2-D constant indices keep it in registers
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
      for (int b = 0; b < NU; ++b)
        if (b <= a) {
          const bool fx = ((fmask >> a) | (fmask >> b)) & 1u;
          L[a][b] = fx ? (a == b ? 1.0 : 0.0) : Hs[a * NU + b];
        }
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      double dk = L[k][k];
#pragma unroll
      for (int q = 0; q < NU; ++q)
        if (q < k) dk -= L[k][q] * L[k][q];
      const double idk = rsqrt(dk);
      idg[k] = idk;
#pragma unroll
      for (int i = 0; i < NU; ++i)
        if (i > k) {
          double v = L[i][k];
#pragma unroll
          for (int q = 0; q < NU; ++q)
            if (q < k) v -= L[i][q] * L[k][q];
          L[i][k] = v * idk;
        }
    }
    if (t < NU) {
      double y[NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double v = (i == t) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < NU; ++q)
          if (q < i) v -= L[i][q] * y[q];
        y[i] = v * idg[i];
      }
#pragma unroll
      for (int i = NU - 1; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int q = 0; q < NU; ++q)
          if (q > i) v -= L[q][i] * y[q];
        y[i] = v * idg[i];
      }
#pragma unroll
      for (int i = 0; i < NU; ++i) Es[i * NU + t] = (((fmask >> i) | (fmask >> t)) & 1u) ? 0.0 : y[i];
    }
  }

  __device__ __forceinline__ static double qs_at(const double* Qp, int r, int t) {
    return Qp[r <= t ? pk(r, t, NX) : pk(t, r, NX)];
  }
  __device__ __forceinline__ static double rs_at(const double* Rp, int a, int b) {
    return Rp[a <= b ? pk(a, b, NU) : pk(b, a, NU)];
  }

  // ------------------------------------------------------------------------------------------------------------
  // bulk-async staging (tma.cuh): lane 0 arms the warp's mbarrier with the byte count and issues the copies; every
  // lane waits on the barrier's phase.  ph[b] counts the completed uses of barrier b (uniform over the warp).
  unsigned phA, phQ, phB, phS0, phS1, phFQ, phA2;
  __device__ __forceinline__ uint64_t* bar(int b) const { return reinterpret_cast<uint64_t*>(slab() + O_BAR) + b; }
  __device__ __forceinline__ void bar_wait(int b, unsigned& ph) {
    mbar_wait(bar(b), ph & 1u);
    ++ph;
  }
  __device__ __forceinline__ void stage_init() {
    if (lane == 0) {
      for (int b = 0; b < NBAR; ++b) mbar_init(bar(b), 1);
      fence_mbar_init();
    }
    phA = phQ = phB = phS0 = phS1 = phFQ = phA2 = 0;
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------------------
  // Backward sweep of particle slot g over the free stages j = N-1 .. Nc (fused.inl bsweep, generic branch): masked
  // Riccati matrix + vector step; writes the hand-off block.  On return (Nc >= 1) the boundary value function sits in
  // the slab: Pc in O_S, Wc in O_W, Zc in MM_Z, pc in V_WV, zc in U_Z.
  // Active state bounds ride on the active-set passes by the method of multipliers: a state component whose
  // multiplier y (pb.yx, y > 0 upper, y < 0 lower bound) is non-zero carries the term (mu/2) (x - b + y/mu)^2, i.e. mu on
  // the diagonal of the stage cost and y - mu b in its linear term (qx).  The forward sweep then sets
  // y+ = mu (x + y/mu - clip(x + y/mu, lx, ux)): stationarity holds exactly with y+, the bound violation of an
  // held component is |y+ - y| / mu (its distance to the bound, on either side) and contracts by about (curvature / mu) per pass; a component whose sign of y
  // changes counts as an active-set change.  Converged: no change, violation below tolerance (qp()).
  __device__ __forceinline__ bool xbox_terms(size_t g, int j, double& qx, bool xa) const {
    qx = 0.0;
    if (xa) {
      const double y = pb.yx[gix(g, N, j, NX) + lc];
      if (y != 0.0) {
        const double b = (y > 0.0) ? pb.ux[gix(g, N, j, NX) + lc] : pb.lx[gix(g, N, j, NX) + lc];
        qx = y - amu * b;
        return true;
      }
    }
    return false;
  }

  __device__ __forceinline__ void bsweep(size_t g) {
    double* sl = slab();
    double *Sm = sl + O_S, *Bm = sl + O_B, *SBWm = sl + O_SBW, *Wm = sl + O_W, *Gm = sl + O_G, *Qb = sl + O_Q;
    double *cv = sl + O_V + V_C, *wvv = sl + O_V + V_WV;
    // pairs (a, b), a <= b, of the packed m x m matrices this lane assembles / stores (constant over the sweep)
    int pa[(NR + 31) / 32], pbb[(NR + 31) / 32];
#pragma unroll
    for (int kk = 0; kk < (NR + 31) / 32; ++kk) {
      int a = 0, rem = lane + 32 * kk;
      while (a < NU - 1 && rem >= NU - a) { rem -= NU - a; ++a; }
      pa[kk] = a;
      pbb[kk] = (a + rem < NU) ? a + rem : NU - 1;
    }
    double *Hm = sl + O_M + MM_H, *Em = sl + O_M + MM_E, *Zm = sl + O_M + MM_Z;
    double *h0m = sl + O_U + U_H0, *hm = sl + O_U + U_H, *kvm = sl + O_U + U_KV, *ufm = sl + O_U + U_UF, *zm = sl + O_U + U_Z;
    const int l = lane;
    const double s = SLEW ? pb.slew[g] : 0.0;
    const FactorSet& F = pb.fs[MODE_POLISH];
    double Pcol[NX], p_l = 0.0;
#pragma unroll
    for (int r = 0; r < NX; ++r) Pcol[r] = 0.0;
    for (int idx = l; idx < NX * NU; idx += 32) Wm[idx] = 0.0;
    for (int idx = l; idx < NU * NU; idx += 32) Zm[idx] = 0.0;
    if (l < NU) zm[l] = 0.0;
    __syncwarp();
    if (N - 1 < Nc) {
      if (lact) {
#pragma unroll
        for (int r = 0; r < NX; ++r) Sm[r * NX + l] = 0.0;
        wvv[l] = 0.0;
      }
      __syncwarp();
      return;
    }
    const bool xa = amu > 0.0 && pb.xany[g] != 0;  // some state multiplier of this particle is non-zero (xbox_terms)
    // per-lane stage vectors travel through registers, loaded one stage ahead
    struct Vec { double c, qref, xp, ub, rref, lo, hi, Rp[(NR + 31) / 32], qx; signed char act; bool ax; };
    auto load_vec = [&](int j) -> Vec {
      Vec v;
      v.c = pb.c[gix(g, N, j, NX) + lc];
      v.qref = pb.qref[gix(g, N, j, NX) + lc];
      v.xp = Xpv[gix(g, N, j, NX) + lc];
      v.ax = xbox_terms(g, j, v.qx, xa);
      const int la = l < NU ? l : NU - 1;
      v.ub = Upv[gix(g, N, j, NU) + la];
      v.rref = pb.rref[gix(g, N, j, NU) + la];
      v.lo = v.hi = 0.0;
      v.act = 0;
      if (pb.has_ubox) {
        v.lo = pb.lu[gix(g, N, j, NU) + la];
        v.hi = pb.uu[gix(g, N, j, NU) + la];
        v.act = pb.act[gix(g, N, j, NU) + la];
      }
#pragma unroll
      for (int k = 0; k < (NR + 31) / 32; ++k)  // packed Rs entries of the (a, b) pairs this lane assembles in H
        v.Rp[k] = pb.Rs[gix(g, N, j, NR) + ((l + 32 * k < NR) ? l + 32 * k : 0)];
      return v;
    };
    auto issue_q = [&](int j) {
      mbar_expect_tx(bar(BAR_Q), NQ * 8);
      bulk_g2s(Qb, pb.Qs + gix(g, N, j, NQ), NQ * 8, bar(BAR_Q));
    };
    auto issue_b = [&](int j) {
      mbar_expect_tx(bar(BAR_B), NX * NU * 8);
      bulk_g2s(Bm, pb.Bm + gix(g, N, j, NX * NU), NX * NU * 8, bar(BAR_B));
    };
    fence_proxy_async();  // plain stores of the previous phase into buffers the bulk copies are about to overwrite
    __syncwarp();
    // A of stage j goes to buffer (j & 1) when there are two (A2), tracked by its own barrier
    auto issue_a = [&](int j) {
      const bool odd = A2 && (j & 1);
      uint64_t* b = bar(odd ? BAR_A2 : BAR_A);
      mbar_expect_tx(b, NX * NX * 8);
      bulk_g2s(sl + (odd ? O_A2 : O_A), pb.A + gix(g, N, j, NX * NX), NX * NX * 8, b);
    };
    if (l == 0) {
      issue_q(N - 1);
      issue_b(N - 1);
      if (A2 && N - 1 >= 1) issue_a(N - 1);
    }
    Vec vn = load_vec(N - 1);
    for (int j = N - 1; j >= Nc; --j) {
      const Vec vc = vn;
      if (j - 1 >= Nc) vn = load_vec(j - 1);
      const double* Am = sl + ((A2 && (j & 1)) ? O_A2 : O_A);
      // ---- P0: A is needed in the second half of the stage only.  One buffer: start the copy of this stage's A now;
      // two buffers: this stage's A was requested a stage ago, request the next one's
      if (l == 0) {
        if (!A2 && j >= 1) issue_a(j);
        if (A2 && j - 1 >= Nc && j - 1 >= 1) issue_a(j - 1);
      }
      bar_wait(BAR_Q, phQ);
      double Rp[(NR + 31) / 32];
#pragma unroll
      for (int k = 0; k < (NR + 31) / 32; ++k) Rp[k] = vc.Rp[k];
#pragma unroll
      for (int r = 0; r < NX; ++r) {
        const double v = qs_at(Qb, r, lc) + Pcol[r];
        if (lact) Sm[r * NX + l] = v;
      }
      if (vc.ax && lact) Sm[lc * NX + l] += amu;  // (rare: a held state component of this stage, xbox_terms)
      const double q_l = -vc.qref - pb.reg_x * vc.xp + vc.qx;
      const double r_l = -vc.rref - pb.reg_u * vc.ub;
      if (lact) cv[l] = vc.c;
      const bool fixed_l = (l < NU) && vc.act != 0;
      if (l < NU) ufm[l] = (vc.act > 0) ? vc.hi : ((vc.act < 0) ? vc.lo : 0.0);
      const unsigned fmask = __ballot_sync(0xffffffffu, fixed_l);  // bit a: control a of this stage sits at a bound
      __syncwarp();
      if (l == 0 && j - 1 >= Nc) issue_q(j - 1);  // Qb has been consumed
      bar_wait(BAR_B, phB);
      // ---- P1: wv = S c + q + p (element l), row l of S B + W.  S is symmetric: row l is read as column l
      {
        double wv = q_l + p_l;
        double acc[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) acc[a] = SLEW ? Wm[lc * NU + a] : 0.0;
#pragma unroll 4
        for (int k = 0; k < NX; ++k) {
          const double sk = Sm[k * NX + lc];
          wv += sk * cv[k];
#pragma unroll
          for (int a = 0; a < NU; ++a) acc[a] += sk * Bm[k * NU + a];
        }
        if (lact) {
          wvv[l] = wv;
#pragma unroll
          for (int a = 0; a < NU; ++a) SBWm[l * NU + a] = acc[a];
        }
      }
      __syncwarp();
      // ---- P2: H = Rs + [j>=1] s I + B'(S B + W) + W'B + Z (pairs over lanes), h0 = r + B'wv + W'c + z
#pragma unroll
      for (int kk = 0; kk < (NR + 31) / 32; ++kk) {
        const int pi = l + 32 * kk;
        if (pi < NR) {
          const int a = pa[kk], b = pbb[kk];
          double acc = Rp[kk];
          if (SLEW) {
            acc += Zm[a * NU + b];
            if (a == b && j >= 1) acc += s;
          }
          for (int k = 0; k < NX; ++k) {
            acc += Bm[k * NU + a] * SBWm[k * NU + b];
            if (SLEW) acc += Wm[k * NU + a] * Bm[k * NU + b];
          }
          Rp[kk] = acc;
        }
      }
      double h0 = 0.0;
      if (l < NU) {
        h0 = r_l + (SLEW ? zm[l] : 0.0);
        for (int k = 0; k < NX; ++k) {
          h0 += Bm[k * NU + l] * wvv[k];
          if (SLEW) h0 += Wm[k * NU + l] * cv[k];
        }
      }
      __syncwarp();  // every lane has finished with B
      if (l == 0 && j - 1 >= Nc) issue_b(j - 1);
#pragma unroll
      for (int kk = 0; kk < (NR + 31) / 32; ++kk) {
        const int pi = l + 32 * kk;
        if (pi < NR) {
          Hm[pa[kk] * NU + pbb[kk]] = Rp[kk];
          Hm[pbb[kk] * NU + pa[kk]] = Rp[kk];
        }
      }
      if (l < NU) h0m[l] = h0;
      __syncwarp();
      // ---- P3: masked inverse, h = h0 + H ufix, k = ufix - E h
      masked_inverse(Hm, fmask, l, Em);
      if (l < NU) {
        double hv = h0m[l];
#pragma unroll
        for (int t = 0; t < NU; ++t) hv += Hm[l * NU + t] * ufm[t];
        hm[l] = hv;
      }
      __syncwarp();
      if (l < NU) {
        double kv = ufm[l];
#pragma unroll
        for (int b = 0; b < NU; ++b) kv -= Em[l * NU + b] * hm[b];
        kvm[l] = kv;
      }
      __syncwarp();
      double Kc[NU], Gc[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) { Kc[a] = 0.0; Gc[a] = 0.0; }
      if (j >= 1) {
        if (A2 && (j & 1)) bar_wait(BAR_A2, phA2);
        else bar_wait(BAR_A, phA);
        // ---- P4: column l of G = (S B + W)' A, K = -E G, p' = A' wv + G' k, S A
        double pn = 0.0;
        double SAc[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) SAc[r] = 0.0;
#pragma unroll 2
        for (int k = 0; k < NX; k += 2) {
          const double a0 = Am[k * NX + lc], a1 = Am[(k + 1) * NX + lc];
          pn = fma(a0, wvv[k], pn);
          pn = fma(a1, wvv[k + 1], pn);
#pragma unroll
          for (int a = 0; a < NU; ++a) Gc[a] = fma(SBWm[k * NU + a], a0, Gc[a]);
#pragma unroll
          for (int a = 0; a < NU; ++a) Gc[a] = fma(SBWm[(k + 1) * NU + a], a1, Gc[a]);
#pragma unroll
          for (int r = 0; r < NX; ++r) {  // two multiply-adds per broadcast LDS.128 (not a*b + c*d: that costs three)
            const double2 sv = *reinterpret_cast<const double2*>(&Sm[r * NX + k]);
            SAc[r] = fma(sv.x, a0, SAc[r]);
            SAc[r] = fma(sv.y, a1, SAc[r]);
          }
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          double acc = 0.0;
#pragma unroll
          for (int b = 0; b < NU; ++b) acc -= Em[a * NU + b] * Gc[b];
          Kc[a] = acc;
          pn += Gc[a] * kvm[a];
        }
        __syncwarp();  // every lane has finished reading S
        if (lact) {
#pragma unroll
          for (int r = 0; r < NX; ++r) Sm[r * NX + l] = SAc[r];
#pragma unroll
          for (int a = 0; a < NU; ++a) Gm[a * NX + l] = Gc[a];
        }
        __syncwarp();
        // ---- P' = A' (S A) + G' K, column l
#pragma unroll
        for (int r = 0; r < NX; ++r) Pcol[r] = 0.0;
#pragma unroll 2
        for (int k = 0; k < NX; ++k) {
          const double sa = Sm[k * NX + lc];
#pragma unroll
          for (int r = 0; r < NX; r += 2) {
            const double2 av = *reinterpret_cast<const double2*>(&Am[k * NX + r]);
            Pcol[r] += av.x * sa;
            Pcol[r + 1] += av.y * sa;
          }
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          const double ka = Kc[a];
#pragma unroll
          for (int r = 0; r < NX; r += 2) {
            const double2 gv = *reinterpret_cast<const double2*>(&Gm[a * NX + r]);
            Pcol[r] += gv.x * ka;
            Pcol[r + 1] += gv.y * ka;
          }
        }
        p_l = pn;
      }
      // ---- P5: hand-off (fixed controls carry G[a,:], h[a] for the multiplier), next stage's slew blocks
      {
        double* hg = pb.hand + gix(g, N, j, NH);
        if (lact) {
#pragma unroll
          for (int a = 0; a < NU; ++a) hg[H_K + a * NX + l] = ((fmask >> a) & 1u) ? Gc[a] : Kc[a];
        }
        if (l < NU) hg[H_KF + l] = ((fmask >> l) & 1u) ? hm[l] : kvm[l];
#pragma unroll
        for (int kk = 0; kk < (NR + 31) / 32; ++kk) {
          const int pi = l + 32 * kk;
          if (pi < NR) {
            const int a = pa[kk], b = pbb[kk];
            if (pb.has_ubox && a < b) hg[H_H + pko(a, b)] = Hm[a * NU + b];
            if (SLEW) F.E[gix(g, N, j, NR) + pi] = Em[a * NU + b];
          }
        }
      }
      __syncwarp();  // all reads of W, Z, z, E, H and A of this stage are done
      if (SLEW) {
        if (lact) {
#pragma unroll
          for (int a = 0; a < NU; ++a) Wm[l * NU + a] = -s * Kc[a];
        }
        for (int idx = l; idx < NU * NU; idx += 32) Zm[idx] = ((idx / NU == idx % NU) ? s : 0.0) - s * s * Em[idx];
        if (l < NU) zm[l] = (j >= 1) ? -s * kvm[l] : 0.0;
      }
      __syncwarp();
    }
    // boundary value function for the consensus step
    if (lact) {
#pragma unroll
      for (int r = 0; r < NX; ++r) Sm[r * NX + l] = Pcol[r];
      wvv[l] = p_l;
    }
    fence_proxy_async();  // the hand-off written above is bulk-copied by the forward sweep
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------------------
  // Gradients of the particle's cost w.r.t. the shared controls (Solver::shared_grad), VB directions at a time:
  // direction v < D is the Hessian action on the unit vector e_v (homogeneous pass), direction v == D the gradient at
  // uc = 0 with all linear terms.  Blocking the directions lets every matrix entry (A, Qs, Pc, B rows, read per lane
  // from global memory / the slab) serve VB products and gives VB independent dependency chains.  The boundary value
  // function is in the slab (see bsweep).  emit(v, j, value) is called in lane b < NU with d/du_{j,b}.
  // Scratch for the shared-stage states of the VB directions: the ADMM factor set's K array of this particle
  // (re-built by the thread-per-particle kernel whenever it starts an ADMM block).
  static constexpr int VB = 4;
  template <typename Emit>
  __device__ __forceinline__ void shared_grads(size_t g, int v0, Emit emit) {
    double* sl = slab();
    double *Pc = sl + O_S, *Wc = sl + O_W, *Zc = sl + O_M + MM_Z, *pcv = sl + O_V + V_WV, *zc = sl + O_U + U_Z;
    double *xv = sl + O_A, *av = sl + O_A + VB * NX;  // [VB][NX] each (the A buffer is idle between the sweeps)
    double* xs = pb.fs[MODE_ADMM].K + gix(g, N, 0, NU * NX);  // [Nc][VB][NX]
    const int l = lane;
    const double s = SLEW ? pb.slew[g] : 0.0;
    const bool xa = amu > 0.0 && pb.xany[g] != 0;
    // u_{j,t} of direction vb: unit vector for v < D, zero for the inhomogeneous direction (and beyond it)
    auto uval = [&](int vb, int j, int t) -> double { return (v0 + vb < D && j >= 0 && j < Nc && j * NU + t == v0 + vb) ? 1.0 : 0.0; };
    auto inhom = [&](int vb) -> bool { return v0 + vb == D; };
    if (lact) {
#pragma unroll
      for (int vb = 0; vb < VB; ++vb) xv[vb * NX + l] = 0.0;
    }
    __syncwarp();
    for (int j = 0; j < Nc; ++j) {
      const double* Ag = pb.A + gix(g, N, j, NX * NX);
      const double* Bg = pb.Bm + gix(g, N, j, NX * NU);
      const double cj = pb.c[gix(g, N, j, NX) + lc];
      double acc[VB];
#pragma unroll
      for (int vb = 0; vb < VB; ++vb) acc[vb] = inhom(vb) ? cj : 0.0;
      if (j >= 1) {
#pragma unroll 4
        for (int t = 0; t < NX; ++t) {
          const double a = Ag[lc * NX + t];
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) acc[vb] = fma(a, xv[vb * NX + t], acc[vb]);
        }
      }
      for (int t = 0; t < NU; ++t) {
        const double b = Bg[lc * NU + t];
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) acc[vb] = fma(b, uval(vb, j, t), acc[vb]);
      }
      __syncwarp();
      if (lact) {
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) {
          xv[vb * NX + l] = acc[vb];
          xs[(j * VB + vb) * NX + l] = acc[vb];
        }
      }
      __syncwarp();
    }
    if (lact) {
#pragma unroll
      for (int vb = 0; vb < VB; ++vb) av[vb * NX + l] = 0.0;
    }
    __syncwarp();
    for (int j = Nc - 1; j >= 0; --j) {
      const double* Bg = pb.Bm + gix(g, N, j, NX * NU);
      const double* Qg = pb.Qs + gix(g, N, j, NQ);
      const double* Rg = pb.Rs + gix(g, N, j, NR);
      if (lact) {
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) xv[vb * NX + l] = xs[(j * VB + vb) * NX + l];
      }
      __syncwarp();
      double an[VB], gu[VB];
      double qxj;
      const double mxj = xbox_terms(g, j, qxj, xa) ? amu : 0.0;
      {
        const double q0 = -pb.qref[gix(g, N, j, NX) + lc] - pb.reg_x * Xpv[gix(g, N, j, NX) + lc] + qxj;
        const int la = l < NU ? l : NU - 1;
        const double r0 = -pb.rref[gix(g, N, j, NU) + la] - pb.reg_u * Upv[gix(g, N, j, NU) + la];
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) {
          an[vb] = inhom(vb) ? q0 : 0.0;
          gu[vb] = inhom(vb) ? r0 : 0.0;
        }
      }
#pragma unroll 4
      for (int t = 0; t < NX; ++t) {
        const double qv = qs_at(Qg, lc, t);
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) an[vb] = fma(qv, xv[vb * NX + t], an[vb]);
      }
#pragma unroll
      for (int vb = 0; vb < VB; ++vb) an[vb] = fma(mxj, xv[vb * NX + lc], an[vb]);
      if (l < NU) {
        for (int t = 0; t < NU; ++t) {
          const double rv = rs_at(Rg, l, t);
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) gu[vb] = fma(rv, uval(vb, j, t), gu[vb]);
        }
      }
      if (j == Nc - 1) {
        const double pcl = pcv[lc];
#pragma unroll
        for (int vb = 0; vb < VB; ++vb)
          if (inhom(vb)) an[vb] += pcl;
#pragma unroll 4
        for (int t = 0; t < NX; ++t) {
          const double pv = Pc[t * NX + lc];  // Pc symmetric: row l read as column l
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) an[vb] = fma(pv, xv[vb * NX + t], an[vb]);
        }
        if (SLEW) {
          for (int t = 0; t < NU; ++t) {
            const double wv = Wc[lc * NU + t];
#pragma unroll
            for (int vb = 0; vb < VB; ++vb) an[vb] = fma(wv, uval(vb, j, t), an[vb]);
          }
          if (l < NU) {
            const double zl = zc[l];
#pragma unroll
            for (int vb = 0; vb < VB; ++vb)
              if (inhom(vb)) gu[vb] += zl;
            for (int r = 0; r < NX; ++r) {
              const double wv = Wc[r * NU + l];
#pragma unroll
              for (int vb = 0; vb < VB; ++vb) gu[vb] = fma(wv, xv[vb * NX + r], gu[vb]);
            }
            for (int t = 0; t < NU; ++t) {
              const double zv = Zc[l * NU + t];
#pragma unroll
              for (int vb = 0; vb < VB; ++vb) gu[vb] = fma(zv, uval(vb, j, t), gu[vb]);
            }
          }
        }
      } else {
        const double* An = pb.A + gix(g, N, j + 1, NX * NX);
#pragma unroll 4
        for (int r = 0; r < NX; ++r) {
          const double a = An[r * NX + lc];
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) an[vb] = fma(a, av[vb * NX + r], an[vb]);
        }
        if (SLEW && l < NU) {
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) gu[vb] += s * (uval(vb, j, l) - uval(vb, j + 1, l));
        }
      }
      if (SLEW && j >= 1 && l < NU) {
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) gu[vb] += s * (uval(vb, j, l) - uval(vb, j - 1, l));
      }
      __syncwarp();
      if (lact) {
#pragma unroll
        for (int vb = 0; vb < VB; ++vb) av[vb * NX + l] = an[vb];
      }
      __syncwarp();
      if (l < NU) {
        for (int r = 0; r < NX; ++r) {
          const double b = Bg[r * NU + l];
#pragma unroll
          for (int vb = 0; vb < VB; ++vb) gu[vb] = fma(b, av[vb * NX + r], gu[vb]);
        }
#pragma unroll
        for (int vb = 0; vb < VB; ++vb)
          if (v0 + vb <= D) emit(v0 + vb, j, gu[vb]);
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------------------
  // Forward sweep of particle slot g (fused.inl fsweep, generic branch): rollout of the policy, multipliers and
  // primal-dual update of the free-stage active set, SCP bookkeeping.  st: per-warp statistics (valid in every lane).
  // The stage matrices [A | B | hand-off | E] stream through a two-deep ring of bulk copies; Qs (objective only) is
  // requested at stage start and waited for at its end; the per-lane vectors travel through registers.
  // Rows of A (and of K) are read with a rotated column index, (t + row) mod NX, which keeps lanes that read
  // different rows of an unpadded matrix on different banks.
  __device__ __forceinline__ void fsweep(size_t g, double* st) {
    const double iamu = amu > 0.0 ? 1.0 / amu : 0.0;
    const bool xa = amu > 0.0 && pb.xany[g] != 0;
    bool xnew = false;
    double* sl = slab();
    double* Qb = sl + O_Q;
    double *xv = sl + O_V + V_X, *dxv = sl + O_V + V_DX;
    double *um = sl + O_U + U_U, *upm = sl + O_U + U_UP, *dum_ = sl + O_U + U_DU;
    const int l = lane;
    const double s = SLEW ? pb.slew[g] : 0.0;
    const FactorSet& F = pb.fs[MODE_POLISH];
    const double* uc = (Nc >= 1) ? pb.uc + (size_t)prob * D : nullptr;  // written by the team leader: read through L2
    const double rdiag = pb.rdiag0[g];
    double dxm = 0.0, dum = 0.0, obj = 0.0, nanf = 0.0, kkt_v = 0.0, uviol = 0.0, xviol = 0.0, changed = 0.0, gscale = 1.0,
           zscale = 1.0, nact = 0.0;
    struct Vec { double c, xp, xref, qref, lx, ux, yx, ub, uref, rref, lo, hi, kf; signed char act; };
    auto load_vec = [&](int j) -> Vec {
      Vec v;
      v.c = pb.c[gix(g, N, j, NX) + lc];
      v.xp = Xpv[gix(g, N, j, NX) + lc];
      v.xref = pb.xref[gix(g, N, j, NX) + lc];
      v.qref = pb.qref[gix(g, N, j, NX) + lc];
      v.lx = -INFINITY; v.ux = INFINITY; v.yx = 0.0;
      if (pb.has_xbox) {
        v.lx = pb.lx[gix(g, N, j, NX) + lc];
        v.ux = pb.ux[gix(g, N, j, NX) + lc];
        if (xa) v.yx = pb.yx[gix(g, N, j, NX) + lc];
      }
      const int la = l < NU ? l : NU - 1;
      v.ub = Upv[gix(g, N, j, NU) + la];
      v.uref = pb.uref[gix(g, N, j, NU) + la];
      v.rref = pb.rref[gix(g, N, j, NU) + la];
      v.lo = -INFINITY; v.hi = INFINITY; v.act = 0; v.kf = 0.0;
      if (pb.has_ubox) {
        v.lo = pb.lu[gix(g, N, j, NU) + la];
        v.hi = pb.uu[gix(g, N, j, NU) + la];
      }
      if (j >= Nc) {
        v.kf = pb.hand[gix(g, N, j, NH) + H_KF + la];
        if (pb.has_ubox) v.act = pb.act[gix(g, N, j, NU) + la];
      }
      return v;
    };
    auto issue_f = [&](int j) {  // lane 0: stage j into ring slot j & 1
      double* dst = sl + (j & 1) * FSLOT;
      uint64_t* b = bar(BAR_S0 + (j & 1));
      uint32_t bytes = NX * NU * 8;
      if (j >= 1) bytes += NX * NX * 8;
      if (j >= Nc) bytes += NH * 8 + (SLEW ? NR * 8 : 0);
      mbar_expect_tx(b, bytes);
      if (j >= 1) bulk_g2s(dst + F_A, pb.A + gix(g, N, j, NX * NX), NX * NX * 8, b);
      bulk_g2s(dst + F_B, pb.Bm + gix(g, N, j, NX * NU), NX * NU * 8, b);
      if (j >= Nc) {
        bulk_g2s(dst + F_H, pb.hand + gix(g, N, j, NH), NH * 8, b);
        if (SLEW) bulk_g2s(dst + F_E, F.E + gix(g, N, j, NR), NR * 8, b);
      }
    };
    if (lact) xv[l] = 0.0;
    if (l < NU) upm[l] = 0.0;
    fence_proxy_async();  // the ring slots overlay buffers the backward sweep / consensus step wrote with plain stores
    __syncwarp();
    if (l == 0) {
      issue_f(0);
      if (N > 1) issue_f(1);
    }
    Vec vn = load_vec(0);
    for (int j = 0; j < N; ++j) {
      const Vec vc = vn;
      if (j + 1 < N) vn = load_vec(j + 1);
      const double* slot = sl + (j & 1) * FSLOT;
      const double *Am = slot + F_A, *Bm = slot + F_B, *Km = slot + F_H + H_K, *Hoff = slot + F_H + H_H, *Es = slot + F_E;
      if (l == 0) {  // Qs of this stage is needed for the objective at the end of the stage only
        mbar_expect_tx(bar(BAR_FQ), (NQ + NR) * 8);
        bulk_g2s(Qb, pb.Qs + gix(g, N, j, NQ), NQ * 8, bar(BAR_FQ));
        bulk_g2s(sl + O_M, pb.Rs + gix(g, N, j, NR), NR * 8, bar(BAR_FQ));
      }
      const signed char act = (l < NU) ? vc.act : 0;
      const double lo = vc.lo, hi = vc.hi;
      const unsigned amask = __ballot_sync(0xffffffffu, act != 0);
      if (j & 1) bar_wait(BAR_S1, phS1);
      else bar_wait(BAR_S0, phS0);
      // ---- controls
      double u = 0.0, gx = 0.0;
      if (l < NU) {
        if (j < Nc) {
          u = __ldcg(uc + j * NU + l);
        } else {
          if (j >= 1) {
            for (int t = 0; t < NX; ++t) {
              const int tt = (t + l >= NX) ? t + l - NX : t + l;
              gx += Km[l * NX + tt] * xv[tt];
            }
          }
          u = vc.kf + gx;
          if (SLEW && j >= 1)
            for (int b = 0; b < NU; ++b) u += s * rs_at(Es, l, b) * upm[b];
          if (act != 0) u = (act > 0) ? hi : lo;
        }
        um[l] = u;
      }
      __syncwarp();
      if (l < NU && j >= Nc && pb.has_ubox) {
        signed char na = 0;
        if (act != 0) {
          // dQ/du_a = H[a,:] u + G[a,:] x + h0[a] - s u_{j-1,a};  kf[a] already holds h0[a] + H[a, fixed] u_fixed
          double acc = vc.kf + gx;
          for (int b = 0; b < NU; ++b)
            if (b != l && !((amask >> b) & 1u)) acc += Hoff[l < b ? pko(l, b) : pko(b, l)] * um[b];
          if (SLEW && j >= 1) acc -= s * upm[l];
          const double mu = -acc;
          if (act > 0) {
            kkt_v = fmax(kkt_v, fmax(acc, 0.0));
            if (mu > 0.0) na = 1;
          } else {
            kkt_v = fmax(kkt_v, fmax(-acc, 0.0));
            if (mu < 0.0) na = -1;
          }
        } else {
          uviol = fmax(uviol, fmax(u - hi, lo - u));
          if (u - hi > 0.0 && hi < INFINITY) na = 1;
          else if (u - lo < 0.0 && lo > -INFINITY) na = -1;
          if (na != 0) nact += 1.0;
        }
        if (na != act) {
          changed += 1.0;
          pb.act[gix(g, N, j, NU) + l] = na;
        }
      }
      // ---- x+ = A x + B u + c (row l; A_0 = 0: lqp_utils.jl:288-296)
      double xn = vc.c;
      if (j >= 1) {
#pragma unroll 4
        for (int t = 0; t < NX; ++t) {
          const int tt = (t + lc >= NX) ? t + lc - NX : t + lc;
          xn += Am[lc * NX + tt] * xv[tt];
        }
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) xn += Bm[lc * NU + a] * um[a];
      // ---- bookkeeping (scp_mpc.py:397-405) and the scales of the acceptance test
      double d2 = 0.0, dx = 0.0;
      if (lact) {
        Xc[gix(g, N, j, NX) + l] = xn;
        const double d = xn - vc.xp;
        d2 = d * d;
        dx = xn - vc.xref;
        if (!(xn == xn)) nanf = 1.0;
        zscale = fmax(zscale, fabs(xn));
        gscale = fmax(gscale, fabs(vc.qref + pb.reg_x * vc.xp));
        if (pb.has_xbox) {
          xviol = fmax(xviol, fmax(xn - vc.ux, vc.lx - xn));
          if (amu > 0.0 && (xa || xn > vc.ux || xn < vc.lx)) {  // multiplier update (xbox_terms); nothing to do while no
                                                                  // multiplier of the particle is set and x is inside
            const double yo = vc.yx;
            const double t = fma(yo, iamu, xn);
            double yn = amu * (t - fmin(fmax(t, vc.lx), vc.ux));
            if ((yn > 0.0) != (yo > 0.0) || (yn < 0.0) != (yo < 0.0)) {
              changed += 1.0;
              if (yo == 0.0) {
                nact += 1.0;
              } else {
                // a component that was held and comes off its bound: the pass was stationary with the multiplier
                // y + mu (x - b) of the wrong sign, not with y+ = 0
                kkt_v = fmax(kkt_v, fabs(amu * (t - (yo > 0.0 ? vc.ux : vc.lx)) - yn));
              }
            }
            if (yn != 0.0) xviol = fmax(xviol, fabs(yn - yo) * iamu);  // held component: distance to its bound, either side
            if (yn != yo) {
              pb.yx[gix(g, N, j, NX) + l] = yn;
              xnew = true;
            }
          }
        }
      }
      double e2 = 0.0, du = 0.0;
      if (l < NU) {
        Uc[gix(g, N, j, NU) + l] = u;
        const double d = u - vc.ub;
        e2 = d * d;
        du = u - vc.uref;
        if (!(u == u)) nanf = 1.0;
        zscale = fmax(zscale, fabs(u));
        gscale = fmax(gscale, fabs(vc.rref + pb.reg_u * vc.ub));
      }
      __syncwarp();  // every lane has read x, u_prev and the ring slot of this stage
      if (l == 0 && j + 2 < N) issue_f(j + 2);
      if (lact) {
        xv[l] = xn;
        dxv[l] = dx;
      }
      if (l < NU) {
        upm[l] = u;
        dum_[l] = du;
      }
      dxm = fmax(dxm, warp_sum(d2));
      dum = fmax(dum, warp_sum(e2));
      __syncwarp();
      bar_wait(BAR_FQ, phFQ);
      {
        const double* Rb = sl + O_M;
        if (lact) {
          double acc = 0.0;
          for (int t = 0; t < NX; ++t) {
            const int tt = (t + l >= NX) ? t + l - NX : t + l;
            acc += qs_at(Qb, l, tt) * dxv[tt];
          }
          obj += dx * (acc - pb.reg_x * dx);
        }
        if (l < NU) {
          double acc = 0.0;
          for (int t = 0; t < NU; ++t) acc += rs_at(Rb, l, t) * dum_[t];
          obj += du * (acc - (pb.reg_u + (j == 0 ? rdiag : 0.0)) * du);
        }
      }
      __syncwarp();  // Qb / Rs staging free for the next stage
    }
    if (!xa && __any_sync(0xffffffffu, xnew) && lane == 0) pb.xany[g] = 1;
    st[ST_DX] = sqrt(dxm);
    st[ST_DU] = sqrt(dum);
    st[ST_OBJ] = warp_sum(obj);
    st[ST_NAN] = warp_max(nanf);
    st[ST_KKT] = warp_max(kkt_v);
    st[ST_UVIOL] = warp_max(uviol);
    st[ST_XVIOL] = warp_max(xviol);
    st[ST_CHANGED] = warp_max(changed);
    st[ST_GSCALE] = warp_max(gscale);
    st[ST_ZSCALE] = warp_max(zscale);
    st[ST_XACT] = warp_max(nact);  // (here: components newly put on a bound by this sweep)
  }

  // ------------------------------------------------------------------------------------------------------------
  // One convex sub-problem: active-set passes (fused.inl fused_qp_impl).  Returns status bits, or -1 when the problem
  // has to be handed to the thread-per-particle kernel (state bound active / active-set iteration stalls).
  // S0: a Solver object of the same shape whose CTA-cooperative shared-block routines (factor_shared, apply_shared,
  // shared_active_update, xreduce) are reused by the team leader.
  template <typename S0T>
  __device__ __forceinline__ int qp(S0T& S0, int ppw, int ppc, double* st, long long* cnt) {
    const bool constrained = pb.has_ubox || pb.has_xbox;
    const FactorSet& F = pb.fs[MODE_POLISH];
    const int HG = D * D + D;  // consensus payload: columns of sum_i H_i, then sum_i g_i
    double* wacc = (Nc >= 1) ? pb.wscr + ((size_t)blockIdx.x * W + wid) * (size_t)tm.cap : nullptr;
    double* cacc = (Nc >= 1) ? pb.wscr + ((size_t)gridDim.x * W + blockIdx.x) * (size_t)tm.cap : nullptr;
    int unsettled = 0, settled = 0;
    amu = (pb.has_xbox && pb.al_mu > 0.0) ? pb.al_mu : 0.0;
    double base = amu;
    int need = 2;
    for (int tries = 0;; ++tries) {
      // ---- backward sweeps + consensus terms of this warp's particles
      for (int q = 0; q < ppw; ++q) {
        const int i = tm.idx * ppc + q * W + wid;
        if (i < M) {
          const size_t g = (size_t)prob * pb.Mpad + i;
          bsweep(g);
          if (Nc >= 1) {
            for (int v0 = 0; v0 <= D; v0 += VB) {
              shared_grads(g, v0, [&](int v, int j, double val) {
                double* dst = wacc + (size_t)v * D + j * NU + lane;
                *dst = (q == 0) ? val : *dst + val;
              });
            }
          }
        } else if (q == 0 && Nc >= 1) {
          for (int k = lane; k < HG; k += 32) wacc[k] = 0.0;
        }
      }
      cnt[2] += 1;
      cnt[7] += 1;
      double kk_c = 0.0, uv_c = 0.0, ch_c = 0.0;
      if (Nc >= 1) {
        // ---- CTA partial (warp order), team reduction, shared block solved by the team leader
        __syncthreads();
        for (int k = tid; k < HG; k += T) {
          double acc = 0.0;
          for (int w = 0; w < W; ++w) acc += pb.wscr[((size_t)blockIdx.x * W + w) * (size_t)tm.cap + k];
          cacc[k] = acc;
        }
        __syncthreads();
        double* Hc = F.Hc + (size_t)prob * D * D;
        double* Hi = F.Hci + (size_t)prob * D * D;
        double* lead = (tm.idx == 0) ? S0.scr() : nullptr;  // leader staging: scratch_stride >= D * D + D (host)
        team_reduce(cacc, HG, lead, true, 0u);
        double* res = cred;  // [kk, uv, ch] broadcast with uc
        if (tm.idx == 0 && !tm.dead) {
          S0.xreduce(lead, HG, true);  // particles sharded over GPUs: sum over the ranks (no-op for one GPU)
          // column v of sum H_i sits in lead[v*D ..]; symmetrise (Solver::assemble_shared)
          for (int idx = tid; idx < D * D; idx += T) {
            const int a = idx / D, b = idx % D;
            Hi[idx] = 0.5 * (lead[(size_t)b * D + a] + lead[(size_t)a * D + b]);
          }
          for (int a = tid; a < D; a += T) pb.gc[(size_t)prob * D + a] = lead[(size_t)D * D + a];
          __syncthreads();
          for (int idx = tid; idx < D * D; idx += T) Hc[idx] = Hi[idx];
          __syncthreads();
          // the shared block is a D-dimensional box QP once the particles' value functions are fixed: iterate its
          // active set here, without sweeping the horizon again (fused_qp_impl)
          for (int t = 0; t <= 2 * D; ++t) {
            S0.factor_shared(MODE_POLISH);
            S0.apply_shared(MODE_POLISH, 0.0);
            if (!pb.has_ubox) break;
            kk_c = uv_c = 0.0;
            ch_c = S0.shared_active_update(&kk_c, &uv_c);
            if (ch_c == 0.0) break;
          }
          if (tid == 0) { res[0] = kk_c; res[1] = uv_c; res[2] = ch_c; }
          __threadfence();
          __syncthreads();
        }
        double* bb = cred + 8;
        if (tm.idx == 0 && tid < 3) bb[tid] = res[tid];
        if (tm.idx == 0 && S0.xdead) tm.dead = true;  // a peer rank never arrived: release the team as well
        team_bcast(bb, 3);  // (orders the leader's pb.uc stores before the flag the others wait for)
        if (tm.dead) return 4;
        kk_c = bb[0]; uv_c = bb[1]; ch_c = bb[2];
        __syncthreads();
      }
      // ---- forward sweeps
      double acc_st[ST_COUNT];
      for (int k = 0; k < ST_COUNT; ++k) acc_st[k] = 0.0;
      acc_st[ST_GSCALE] = acc_st[ST_ZSCALE] = 1.0;
      for (int q = 0; q < ppw; ++q) {
        const int i = tm.idx * ppc + q * W + wid;
        if (i < M) {
          double s1[ST_COUNT];
          for (int k = 0; k < ST_COUNT; ++k) s1[k] = 0.0;
          fsweep((size_t)prob * pb.Mpad + i, s1);
          for (int k = 0; k < ST_COUNT; ++k) acc_st[k] = (k == ST_OBJ) ? acc_st[k] + s1[k] : fmax(acc_st[k], s1[k]);
        }
      }
      cnt[3] += 1;
      cnt[7] += 1;
      // ---- statistics: warps -> CTA -> team (-> ranks) -> everybody
      __syncthreads();
      double* wst = cred + 16;  // [W][ST_COUNT]
      if (lane == 0)
        for (int k = 0; k < ST_COUNT; ++k) wst[wid * ST_COUNT + k] = acc_st[k];
      __syncthreads();
      double* cst = wst + W * ST_COUNT;  // [ST_COUNT]
      if (tid < ST_COUNT) {
        double a = wst[tid];
        for (int w = 1; w < W; ++w) a = (tid == ST_OBJ) ? a + wst[w * ST_COUNT + tid] : fmax(a, wst[w * ST_COUNT + tid]);
        cst[tid] = a;
      }
      __syncthreads();
      team_reduce(cst, ST_COUNT, cst, false, 1u << ST_OBJ);
      if (tm.idx == 0 && !tm.dead) S0.xreduce(cst, ST_COUNT, false, 1u << ST_OBJ);
      if (tm.idx == 0 && S0.xdead) tm.dead = true;
      team_bcast(cst, ST_COUNT);
      if (tm.dead || S0.xdead) return 4;
      for (int k = 0; k < ST_COUNT; ++k) st[k] = cst[k];
      __syncthreads();
      st[ST_KKT] = fmax(st[ST_KKT], kk_c);
      st[ST_UVIOL] = fmax(st[ST_UVIOL], uv_c);
      st[ST_CHANGED] += ch_c;
#ifdef PMPC_WPP_DEBUG
      if (prob == 0 && tm.idx == 0 && tid == 0)
        printf("wpp qp try %d: kkt %.3e uviol %.3e xviol %.3e changed %.0f nact %.0f mu %.0e\n", tries, st[ST_KKT],
               st[ST_UVIOL], st[ST_XVIOL], st[ST_CHANGED], st[ST_XACT], amu);
#endif
      if (st[ST_NAN] > 0.0) return 1;
      if (!constrained) {
        cnt[5] += 1;
        return 0;
      }
      const double tol_g = pb.kkt_tol * st[ST_GSCALE], tol_z = pb.kkt_tol * st[ST_ZSCALE];
      if (st[ST_KKT] <= tol_g && st[ST_UVIOL] <= tol_z && st[ST_XVIOL] <= tol_z) {
        cnt[5] += 1;
        return 0;
      }
      if (st[ST_XVIOL] > tol_z) {
        // a state bound is active.  Multiplier method (xbox_terms): sweep on.  The penalty starts small, where the
        // undamped active-set iteration over controls and state bounds together is stable, and goes up a notch after
        // two passes without a change of the set (the violation contracts by about curvature / mu per pass); when the
        // set starts moving again it goes back down and waits longer before the next raise.  Without the method
        // (al_mu = 0), or when it does not get there: ADMM in the thread-per-particle kernel
        if (!(pb.al_mu > 0.0) || tries + 1 >= pb.polish_iters + 150) {
#ifdef PMPC_WPP_DEBUG
          if (prob == 0 && tm.idx == 0 && tid == 0) printf("wpp qp: hands over (state bounds) after try %d\n", tries);
#endif
          return -1;
        }
        if (st[ST_CHANGED] <= 0.0) {
          unsettled = 0;
          if (++settled >= need && amu < 3e5) {  // quiet passes at this penalty: one notch up
            amu *= 10.0;
            settled = 0;
          }
        } else {
          settled = 0;
          ++unsettled;
          if (unsettled >= 2 && amu > base) {  // the set moves again: back down, and more patience before the next raise
            amu = fmax(amu * 0.1, base);
            need = min(2 * need, 16);
            unsettled = 0;
          } else if (unsettled >= 24 && base > 1.0) {  // cycling at the starting penalty itself: a softer start
            base = fmax(base / 3.0, 1.0);
            amu = base;
            unsettled = 0;
          }
        }
        continue;
      }
      if (st[ST_CHANGED] <= 0.0 || tries + 1 >= pb.polish_iters + (amu > 0.0 ? 150 : 0)) return -1;  // stalls or cycles: interior point there
    }
  }
  // ------------------------------------------------------------------------------------------------------------
  // The SCP loop (scp_mpc.py:337-428; Solver::run)
  template <typename S0T>
  __device__ __forceinline__ void run(S0T& S0, int Tc, int ppw) {
    M = pb.M;
    N = pb.N;
    Nc = (M == 1) ? 0 : pb.Nc;
    D = Nc * NU;
    prob = blockIdx.x / Tc;
    S0.prob = prob;
    const int ppc = W * ppw;
    tm.Tc = Tc;
    tm.idx = blockIdx.x % Tc;
    tm.cap = pb.tm_cap;
    tm.seq = pb.tseq0;
    tm.dead = false;
    {
      // per problem: [Tc part slots | bcast slot | Tc + 1 flags]
      const size_t per = (size_t)(Tc + 1) * tm.cap + (size_t)Tc + 1;
      double* base = pb.tm + (size_t)prob * per;
      tm.part = base;
      tm.bc = base + (size_t)Tc * tm.cap;
      tm.pflag = reinterpret_cast<unsigned long long*>(tm.bc + tm.cap);
      tm.bflag = tm.pflag + Tc;
    }
    cred = reinterpret_cast<double*>(wpp_smem) + (size_t)W * SLAB;
    stage_init();
    Xc = pb.X; Uc = pb.U; Xpv = pb.Xp; Upv = pb.Up;
    flip = false;
    double st[ST_COUNT];
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool writer = (tm.idx == 0 && tid == 0);
    int it = 0;
    bool handed = false;
    const double *solX = Xc, *solU = Uc;
    for (; it < pb.max_it; ++it) {
      const int status = qp(S0, ppw, ppc, st, cnt);
      if (status == 4) {  // a team member or a peer rank never arrived
        cnt[4] |= 4;
        ++it;
        break;
      }
      if (status < 0) {  // hand the problem to the thread-per-particle kernel at this iteration
        handed = true;
        break;
      }
      solX = Xc;
      solU = Uc;
      if (status & 1) {
        cnt[4] |= 1;
        for (int q = 0; q < ppw; ++q) {  // reference convention: NaN output on failure (osqp_solver.jl:69-71)
          const int i = tm.idx * ppc + q * W + wid;
          if (i < M) {
            const size_t g = (size_t)prob * pb.Mpad + i;
            for (int idx = lane; idx < N * NX; idx += 32) Xc[gix(g, N, 0, NX) + idx] = NAN;
            for (int idx = lane; idx < N * NU; idx += 32) Uc[gix(g, N, 0, NU) + idx] = NAN;
          }
        }
        ++it;
        break;
      }
      const double kkt_rel = st[ST_KKT] / fmax(st[ST_GSCALE], 1.0);
      if (!pb.keep_prev) {  // X_prev <- X is a buffer swap
        double* t0 = Xc; Xc = const_cast<double*>(Xpv); Xpv = t0;
        double* t1 = Uc; Uc = const_cast<double*>(Upv); Upv = t1;
        flip = !flip;
      }
      const double max_res = fmax(st[ST_DX], st[ST_DU]);
      const double obj = st[ST_OBJ] / ((double)N * (double)(pb.world > 1 ? pb.M_total : pb.M));
      const int hrow = it + pb.hist_off;
      if (writer && hrow < pb.hist_len) {
        const size_t Bn = (size_t)pb.B;
        pb.hist[((size_t)0 * pb.hist_len + hrow) * Bn + prob] = obj;
        pb.hist[((size_t)1 * pb.hist_len + hrow) * Bn + prob] = max_res;
        pb.hist[((size_t)2 * pb.hist_len + hrow) * Bn + prob] = 0.0;
        pb.hist[((size_t)3 * pb.hist_len + hrow) * Bn + prob] = kkt_rel;
      }
      if (max_res < pb.res_tol) {
        ++it;
        break;
      }
    }
    // Leave the last solution in pb.X / pb.U and, unless keep_prev, in pb.Xp / pb.Up as well.  A handed-over problem
    // needs only its linearisation point in pb.Xp / pb.Up (the other kernel starts with unswapped buffers).
    if (handed) {
      solX = Xpv;
      solU = Upv;
    }
    for (int q = 0; q < ppw; ++q) {
      const int i = tm.idx * ppc + q * W + wid;
      if (i >= M) continue;
      const size_t g = (size_t)prob * pb.Mpad + i;
      for (int idx = lane; idx < N * NX; idx += 32) {
        const double v = solX[gix(g, N, 0, NX) + idx];
        if (!handed && solX != pb.X) pb.X[gix(g, N, 0, NX) + idx] = v;
        if ((handed || !pb.keep_prev) && solX != pb.Xp) pb.Xp[gix(g, N, 0, NX) + idx] = v;
      }
      for (int idx = lane; idx < N * NU; idx += 32) {
        const double v = solU[gix(g, N, 0, NU) + idx];
        if (!handed && solU != pb.U) pb.U[gix(g, N, 0, NU) + idx] = v;
        if ((handed || !pb.keep_prev) && solU != pb.Up) pb.Up[gix(g, N, 0, NU) + idx] = v;
      }
    }
    if (tm.idx == 0) S0.xclose();
    cnt[0] = it;
    if (writer) {
      for (int k = 0; k < 8; ++k) pb.info[(size_t)k * pb.B + prob] = cnt[k];
      pb.fb[prob] = handed ? 1 : 0;
    }
  }
};

}  // namespace pmpc
