  // =============================================================================================
  // FUSED PATH (textually included inside struct Solver).
  //
  // An active-set (polish) pass is TWO sweeps over the horizon:
  //   bsweep  backward: linearisation (recomputed on the fly for built-in dynamics instead of
  //           streaming f, fx, fu through HBM), masked Riccati matrix step and vector step;
  //   fsweep  forward: rollout of the feedback policy, multipliers of the fixed controls from the
  //           stage Q-function gradient  mu = -(H u + G x + h0 - s u_{j-1}), primal-dual active-set
  //           update, and the SCP residual / objective bookkeeping (scp_mpc.py:397-405).
  // A QP whose active set is already right therefore costs exactly one pass; every change of the
  // active set costs one more.  Between the sweeps only K, k, H (and E with a slew penalty), X, U,
  // X_prev, U_prev and the active set live in HBM.  For a fixed control a the slots K[a,:], k[a]
  // (which would be 0 and the bound) carry G[a,:], h0[a] for the multiplier.
  //
  // Streaming: every warp owns 32 particle slots whose stage blocks are contiguous (warp-tiled SoA),
  // so lane 0 prefetches them DEPTH stages ahead with bulk asynchronous copies (TMA engine,
  // cp.async.bulk + mbarrier complete_tx) into a per-warp ring in shared memory; the lanes read
  // their column conflict-free.  Costs, references and bounds that are identical for all particles
  // of a problem (the usual case: particles differ in their dynamics only) are served from a
  // per-CTA table in shared memory and never streamed.
  // ADMM (active-set iteration stalls, or a state bound active) falls back to the unfused path.
  // =============================================================================================
  static constexpr bool RING_SHAPE = (NX <= 4);
#ifndef PMPC_DEPTH
#define PMPC_DEPTH 2
#endif
  static constexpr int DEPTH = PMPC_DEPTH;
  // ring slot layout (doubles): backward sweep [Xp | Up | act], forward sweep [Xp | Up | K | k | H | (L) | act]
  // hand-off block of one stage, ONE array `pb.hand` with NH rows per (warp tile, stage) so that a single bulk copy brings
  // it in: [sparse stage (HL rows, Dubins shapes only) | K (NU*NX) | k (NU) | H01 (NR-NU)]
  static constexpr int HL = (NX == 4 && NU == 2) ? 10 : 0;
  static constexpr int H_LIN = 0, H_K = HL, H_KF = H_K + NU * NX, H_H = H_KF + NU, NH = H_H + (NR - NU);
  static constexpr int R_XP = 0, R_UP = R_XP + NX * 32, R_HAND = R_UP + NU * 32, R_LIN = R_HAND + H_LIN * 32,
                       R_K = R_HAND + H_K * 32, R_KF = R_HAND + H_KF * 32, R_H = R_HAND + H_H * 32, R_ACTF = R_HAND + NH * 32,
                       R_ACTB = R_UP + NU * 32;
  static constexpr int SLOT = R_ACTF + NU * 4;
  // Of the stage Hessian H only the strictly-upper triangle is handed to the forward sweep: a fixed control's multiplier
  // is h0 + G x + H u, and the part of H u that multiplies FIXED controls (its own diagonal entry included) is known in
  // the backward sweep and folded into the k-slot there (h = h0 + H ufix); the free controls need the off-diagonals.
  static constexpr int NHO = NR - NU;
  __host__ __device__ static constexpr int pko(int a, int b) { return a * NU - (a * (a + 1)) / 2 + (b - a - 1); }  // a < b
  // uniform cost table row: [Qs | Rs | qref | rref | xref | uref | lu | uu]
  static constexpr int T_QS = 0, T_RS = T_QS + NQ, T_QR = T_RS + NR, T_RR = T_QR + NX, T_XR = T_RR + NU,
                       T_UR = T_XR + NX, T_LU = T_UR + NU, T_UU = T_LU + NU, TW = T_UU + NU;

  __host__ __device__ static constexpr int table_doubles(int N_) { return N_ * TW; }
  __host__ __device__ static constexpr int ring_doubles(int nwarps) { return nwarps * DEPTH * SLOT; }

  // shared-memory view rebuilt from the kernel's own dynamic-smem base inside every sweep, so that the
  // accesses compile to LDS / mbarrier on shared addresses (pointers kept in the Solver object are generic)
  struct Sm {
    const double* tab;
    double* ring;
    uint64_t* bars;
  };
  __device__ __forceinline__ Sm smem_view(const DevProb& pb, unsigned char* smem) const {
    double* base = reinterpret_cast<double*>(smem);
    Sm v;
    v.tab = has_tab ? base + ((PPW == 2) ? (tid >> 4) * (pb.tab_doubles / (2 * (T >> 5))) : 0) : nullptr;
    v.ring = pb.ring ? base + pb.tab_doubles + (size_t)wid * DEPTH * SLOT : nullptr;
    v.bars = pb.ring ? reinterpret_cast<uint64_t*>(base + pb.tab_doubles + (size_t)(T >> 5) * DEPTH * SLOT) + wid * DEPTH : nullptr;
    return v;
  }

  // element 0 of the warp's stage-j block of an E-element array (no lane offset): bulk-copy source
  // wb = (g >> 5) * N: first stage block of this warp's particle tile
  template <typename Tp>
  __device__ __forceinline__ static const Tp* blk(const Tp* a, int j, int E, size_t wb) {
    return a + ((wb + (size_t)j) * (size_t)E) * TILE;
  }
  // warp-tile SoA index from sweep-local invariants (wb as above, col = g & 31).  The sweeps use these
  // instead of ix()/ld()/at(): the Solver object lives in local memory (it escapes into the noinline
  // helpers), and every asm volatile "memory" clobber of the ring protocol would force its members to be
  // re-read from there (LDL + long-scoreboard stall) at each use.
  __device__ __forceinline__ static size_t ixw(size_t wb, int col, int j, int E, int e) {
    return ((wb + (size_t)j) * (size_t)E + (size_t)e) * TILE + (size_t)col;
  }

  // dynamic shared memory: [table][ring of every warp][mbarriers of every warp]
  // first: false when a warp of the two-problems-per-warp kernel takes its next pair (the ring and its barriers carry on)
  __device__ __forceinline__ void fused_init(unsigned char* smem, bool first = true) {
    lane = tid & 31;
    wid = tid >> 5;
    col = g % TILE;
    tab = nullptr;
    has_tab = false;
    double* base = reinterpret_cast<double*>(smem);
    const int nw = T >> 5;
    double* tb = base;
    if (first) {
      rcount = 0;
      ring = nullptr;
      bars = nullptr;
    }
    if (first && pb.ring) {
      ring = base + pb.tab_doubles + (size_t)wid * DEPTH * SLOT;
      bars = reinterpret_cast<uint64_t*>(base + pb.tab_doubles + (size_t)nw * DEPTH * SLOT) + wid * DEPTH;
      if (lane == 0) {
        for (int k = 0; k < DEPTH; ++k) mbar_init(bars + k, 1);
        fence_mbar_init();
      }
    }
    if (PPW == 2) tb = base + (tid >> 4) * (pb.tab_doubles / (2 * (T >> 5)));  // one table per half-warp problem
    if (GROUPED && pb.tab_doubles > 0 && pb.uniform[prob]) {
      // particle 0 of the problem speaks for all (checked on the device by uniform_check_kernel)
      const size_t g0 = (size_t)prob * pb.Mpad;
      for (int idx = (PPW == 2 ? hl : tid); idx < N * TW; idx += (PPW == 2 ? 16 : T)) {
        const int j = idx / TW, k = idx % TW;
        double v;
        if (k < T_RS) v = pb.Qs[ixg(j, NQ, k - T_QS, g0)];
        else if (k < T_QR) v = pb.Rs[ixg(j, NR, k - T_RS, g0)];
        else if (k < T_RR) v = pb.qref[ixg(j, NX, k - T_QR, g0)];
        else if (k < T_XR) v = pb.rref[ixg(j, NU, k - T_RR, g0)];
        else if (k < T_UR) v = pb.xref[ixg(j, NX, k - T_XR, g0)];
        else if (k < T_LU) v = pb.uref[ixg(j, NU, k - T_UR, g0)];
        else if (k < T_UU) v = pb.lu[ixg(j, NU, k - T_LU, g0)];
        else v = pb.uu[ixg(j, NU, k - T_UU, g0)];
        tb[idx] = v;
      }
      tab = tb;
      has_tab = true;
    }
    if (PPW == 2) __syncwarp(); else __syncthreads();  // (PPW 2: tables, ring and barriers are per half-warp / per warp)
  }

  // one field (E doubles at table offset `off`) of the stage's cost / reference / bound data
  __device__ __forceinline__ static void cost_field(const double* tab, int j, int off, int E, const double* arr, double* out,
                                                    size_t wb, int col) {
    if (tab != nullptr) {
      const double* t = tab + j * TW + off;
#pragma unroll UNR
      for (int e = 0; e < E; ++e) out[e] = t[e];
    } else {
#pragma unroll UNR
      for (int e = 0; e < E; ++e) out[e] = arr[ixw(wb, col, j, E, e)];
    }
  }

  // x_j = A x_{j-1} + B u_j + c about (xm1 = X_[j], ub = U_prev[j]); A_0 = 0 (lqp_utils.jl:288-296)
  template <bool DUB>
  __device__ __forceinline__ static void lin_stage(const DevProb& pb, int j, const double* xm1, const double* ub, const double* dp,
                                                   double* A, double* Bm, double* c, size_t wb, int col) {
    if constexpr (DUB && NX == 4 && NU == 2) {
      double f[4], fx[16], fu[8];
      dubins_f_fx_fu(xm1, ub, dp[0], dp[1], dp[2], pb.dyn_eps, f, fx, fu);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double v = f[r];
#pragma unroll
        for (int t = 0; t < 2; ++t) v -= fu[r * 2 + t] * ub[t];
        if (j >= 1) {
#pragma unroll
          for (int t = 0; t < 4; ++t) v -= fx[r * 4 + t] * xm1[t];
        }
        c[r] = v;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) A[e] = (j >= 1) ? fx[e] : 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) Bm[e] = fu[e];
    } else {
#pragma unroll UNR
      for (int e = 0; e < NX * NX; ++e) A[e] = pb.A[ixw(wb, col, j, NX * NX, e)];
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) Bm[e] = pb.Bm[ixw(wb, col, j, NX * NU, e)];
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) c[e] = pb.c[ixw(wb, col, j, NX, e)];
    }
  }

  // ---- structure of the Dubins linearisation (lean Dubins variant) ----------------------------------
  // fx = I + N with N non-zero only in rows 0,1 x columns 2,3; fu non-zero at (0,0) (0,1) (1,0) (1,1) (2,0) (3,1)
  // (tests/dubins_car.py:48-90 differentiated).  The stage is carried as 14 doubles
  //     L = [n02 n03 n12 n13 | b00 b01 b10 b11 b20 b31 | c0 c1 c2 c3]
  // which (a) cuts the Riccati algebra from ~330 to ~135 FMAs per stage and (b) is small enough to hand
  // from the backward to the forward sweep through HBM + the TMA ring, so the forward sweep does not
  // evaluate the dynamics again.  Stage 0 has A = 0 (lqp_utils.jl:288-296): its n-entries are unused.
  static constexpr int NLIN = 14;
  // handed through HBM: 10 of the 14 — b20 = v_scale T and b31 = -w_scale T are constants of the particle, and c2, c3
  // (the affine terms of the speed and heading rows) need no trigonometry: the forward sweep rebuilds the four from
  // the particle's parameters and the linearisation point (hand index h -> L index h < 8 ? h : h + 2)
  static constexpr int NLINH = 10;
  static constexpr bool CAN_SPARSE = (NX == 4 && NU == 2);
  __device__ __forceinline__ static void lin_sparse(const DevProb& pb, int j, const double* xm1, const double* ub,
                                                    const double* dp, double* L) {
    if constexpr (CAN_SPARSE) {
      double f[4], fx[16], fu[8];
      dubins_f_fx_fu(xm1, ub, dp[0], dp[1], dp[2], pb.dyn_eps, f, fx, fu);
      L[0] = fx[2]; L[1] = fx[3]; L[2] = fx[6]; L[3] = fx[7];
      L[4] = fu[0]; L[5] = fu[1]; L[6] = fu[2]; L[7] = fu[3]; L[8] = fu[4]; L[9] = fu[7];
      double c0 = f[0] - L[4] * ub[0] - L[5] * ub[1], c1 = f[1] - L[6] * ub[0] - L[7] * ub[1];
      double c2 = f[2] - L[8] * ub[0], c3 = f[3] - L[9] * ub[1];
      if (j >= 1) {
        c0 -= xm1[0] + L[0] * xm1[2] + L[1] * xm1[3];
        c1 -= xm1[1] + L[2] * xm1[2] + L[3] * xm1[3];
        c2 -= xm1[2];
        c3 -= xm1[3];
      }
      L[10] = c0; L[11] = c1; L[12] = c2; L[13] = c3;
    }
  }
  // dense (A, B, c) of a sparse stage, for the shared stages the consensus step reads
  __device__ __forceinline__ static void lin_expand(int j, const double* L, double* A, double* Bm, double* c) {
    if constexpr (CAN_SPARSE) {
#pragma unroll
      for (int e = 0; e < 16; ++e) A[e] = 0.0;
      if (j >= 1) {
        A[0] = A[5] = A[10] = A[15] = 1.0;
        A[2] = L[0]; A[3] = L[1]; A[6] = L[2]; A[7] = L[3];
      }
      Bm[0] = L[4]; Bm[1] = L[5]; Bm[2] = L[6]; Bm[3] = L[7]; Bm[4] = L[8]; Bm[5] = 0.0; Bm[6] = 0.0; Bm[7] = L[9];
      c[0] = L[10]; c[1] = L[11]; c[2] = L[12]; c[3] = L[13];
    }
  }

  // ---- interior-point fallback (ALG = 1) -------------------------------------------------------
  // The primal-dual active-set iteration can cycle on a strictly convex box QP.  When it does, the same
  // two sweeps run a primal-dual interior-point method instead: with slacks sh = hi - u, sl = u - lo and
  // multipliers lh, ll the Newton step for target mu is the LQ problem with
  //     R~ = Rs + diag(lh/sh + ll/sl),   r~ = r - diag(.) u + mu/sh - mu/sl
  // (bsweep), its solution u+ comes out of the rollout (fsweep) together with the fraction-to-boundary
  // ratios and the sums that predict the complementarity after the step; the damped update
  //     u += ap (u+ - u),  lam += ad dlam
  // is applied lazily at the start of the next bsweep.  State: u -> pb.wu, lh -> pb.yu, ll -> pb.wx
  // (the ADMM arrays, indexed with E = NU), shared-block entries in particle 0's slots.
  struct Ipm {
    double mu, mu_prev;  // barrier parameter of the system being built / of the previous one
    double ap, ad;       // step lengths of the previous step (applied lazily)
    int init;            // first iteration: initialise (u, lh, ll) strictly inside the box
  };
  // updated (u, lh, ll) of one bounded scalar: lazy step of the previous iteration, or the initial point
  __device__ __forceinline__ void ipm_point(const Ipm& ip, double u0, double unew, double lo, double hi, double uprev,
                                            double& u, double& lh, double& ll) const {
    if (ip.init) {
      const double w = hi - lo;  // inf when one-sided or free
      const double dl = (w < INFINITY) ? 0.01 * w : 0.01 * fmax(1.0, fabs(lo));
      const double dh = (w < INFINITY) ? 0.01 * w : 0.01 * fmax(1.0, fabs(hi));
      u = fmin(fmax(uprev, lo + dl), hi - dh);
      lh = ip.mu_prev / (hi - u);
      ll = ip.mu_prev / (u - lo);
    } else {
      const double sh = hi - u0, sl = u0 - lo, du = unew - u0;
      const double dlh = ip.mu_prev / sh - lh + (lh / sh) * du;
      const double dll = ip.mu_prev / sl - ll - (ll / sl) * du;
      u = u0 + ip.ap * du;
      lh += ip.ad * dlh;
      ll += ip.ad * dll;
    }
  }
  // fraction-to-boundary ratios and complementarity-prediction sums of one bounded scalar
  __device__ __forceinline__ void ipm_ratios(const Ipm& ip, double u0, double unew, double lo, double hi, double lh,
                                             double ll, double* acc) const {
    const double sh = hi - u0, sl = u0 - lo, du = unew - u0;
    const double dlh = ip.mu / sh - lh + (lh / sh) * du;
    const double dll = ip.mu / sl - ll - (ll / sl) * du;
    if (du > 0.0 && hi < INFINITY) acc[0] = fmin(acc[0], sh / du);
    if (du < 0.0 && lo > -INFINITY) acc[0] = fmin(acc[0], sl / -du);
    if (dlh < 0.0) acc[1] = fmin(acc[1], lh / -dlh);
    if (dll < 0.0) acc[1] = fmin(acc[1], ll / -dll);
    if (hi < INFINITY) {
      acc[2] += lh * sh; acc[3] += dlh * sh; acc[4] += -lh * du; acc[5] += -dlh * du; acc[6] += 1.0;
    }
    if (lo > -INFINITY) {
      acc[2] += ll * sl; acc[3] += dll * sl; acc[4] += ll * du; acc[5] += dll * du; acc[6] += 1.0;
    }
  }

  // lane 0: stream the backward-sweep inputs of stage jj into ring slot `sl`
  template <bool UBOX>
  __device__ __forceinline__ static void issue_b(const DevProb& pb, const Sm& sm, const double* Xpv, const double* Upv, int jj, int sl,
                                                 size_t wb) {
    double* dst = sm.ring + sl * SLOT;
    uint64_t* bar = sm.bars + sl;
    uint32_t bytes = NU * 256;
    if (jj >= 1) bytes += NX * 256;
    if (UBOX) bytes += NU * 32;
    mbar_expect_tx(bar, bytes);
    if (jj >= 1) bulk_g2s(dst + R_XP, blk(Xpv, jj - 1, NX, wb), NX * 256, bar);
    bulk_g2s(dst + R_UP, blk(Upv, jj, NU, wb), NU * 256, bar);
    if (UBOX) bulk_g2s(dst + R_ACTB, blk(pb.act, jj, NU, wb), NU * 32, bar);
  }
  template <bool UBOX, bool SPD = false>
  __device__ __forceinline__ static void issue_f(const DevProb& pb, const Sm& sm, const double* Xpv, const double* Upv, int jj, int sl,
                                                 size_t wb, int Nc) {
    double* dst = sm.ring + sl * SLOT;
    uint64_t* bar = sm.bars + sl;
    // rows of the hand-off block this stage needs: the sparse stage on every stage (SPD), K | k | H01 on the free ones
    const int r0 = SPD ? 0 : H_K, r1 = (jj >= Nc) ? NH : (SPD ? H_K : 0);
    uint32_t bytes = (NX + NU) * 256 + (r1 > r0 ? (r1 - r0) * 256 : 0);
    if (jj >= Nc && UBOX) bytes += NU * 32;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst + R_XP, blk(Xpv, jj, NX, wb), NX * 256, bar);
    bulk_g2s(dst + R_UP, blk(Upv, jj, NU, wb), NU * 256, bar);
    if (r1 > r0) bulk_g2s(dst + R_HAND + r0 * 32, blk(pb.hand, jj, NH, wb) + r0 * 32, (r1 - r0) * 256, bar);
    if (jj >= Nc && UBOX) bulk_g2s(dst + R_ACTF, blk(pb.act, jj, NU, wb), NU * 32, bar);
  }

  // Backward sweep over the stages j = N-1 .. (first ? 0 : Nc).  Free stages (j >= Nc): masked
  // Riccati matrix + vector step; shared stages (j < Nc, only when the linearisation point changed
  // since the last sweep): store the stage data the consensus Schur step (K3) reads.
  template <bool SLEW, bool UBOX, bool DUB, int ALG = 0>
  __device__ __forceinline__ void bsweep(const DevProb& pb, unsigned char* smem, bool first, double* pc, double* zc,
                                         const Ipm ip = Ipm()) {
    // sweep-local copies of the per-thread invariants (see ixw): registers instead of local-memory reloads
    const int N = this->N, Nc = this->Nc, g = this->g, lane = this->lane, col = this->col;
    const size_t P = this->P;
    const bool valid = this->valid, flip = this->flip;
    const double s = this->s;
    const size_t wb = (size_t)(g / TILE) * (size_t)N;
    uint32_t rcount = this->rcount;
    auto ix = [=](int j, int E, int e) -> size_t { return ixw(wb, col, j, E, e); };
    auto ld = [=](const double* a, int j, int E, int e) -> double { return a[ixw(wb, col, j, E, e)]; };
    auto at = [=](double* a, int j, int E, int e) -> double& { return a[ixw(wb, col, j, E, e)]; };
    auto ldp = [=](const double* a, int e) -> double { return a[(size_t)e * P + g]; };
    (void)ix; (void)ld; (void)at; (void)ldp; (void)s; (void)lane;
    const FactorSet& F = pb.fs[MODE_POLISH];
    const Sm sm = smem_view(pb, smem);
    const bool use_ring = RING_SHAPE && sm.ring != nullptr;
    const int jmin = first ? 0 : Nc;
    // linearisation point, re-derived from the kernel parameter (global-memory pointers, not the generic members)
    const double* const Xpv = flip ? pb.X : pb.Xp;
    const double* const Upv = flip ? pb.U : pb.Up;
    double dp[3] = {0.0, 0.0, 0.0};
    if (DUB) {
      dp[0] = ldp(pb.dynp, 0); dp[1] = ldp(pb.dynp, 1); dp[2] = ldp(pb.dynp, 2);
    }
    double Pm[NQ], W[SLEW ? NX * NU : 1], Z[SLEW ? NU * NU : 1], p[NX], z[SLEW ? NU : 1], xq[NX];
#pragma unroll UNR
    for (int e = 0; e < NQ; ++e) Pm[e] = 0.0;
    if (SLEW) {
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) W[e] = 0.0;
#pragma unroll UNR
      for (int e = 0; e < NU * NU; ++e) Z[e] = 0.0;
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) z[e] = 0.0;
    }
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) p[e] = 0.0;
    if (use_ring) {
      // X_prev, U_prev and the active set may have been written with plain stores since the last sweep (the
      // unfused ADMM path of the previous sub-problem, ipm_finish): order them before the bulk reads below
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        for (int d = 0; d < DEPTH; ++d)
          if (N - 1 - d >= jmin) issue_b<UBOX>(pb, sm, Xpv, Upv, N - 1 - d, (rcount + d) % DEPTH, wb);
      }
    }
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) xq[e] = ld(Xpv, N - 1, NX, e);
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) pc[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) zc[e] = 0.0;
    if (Nc >= N && valid) {  // every stage shared: the value function at the boundary is zero
#pragma unroll UNR
      for (int e = 0; e < NX * NX; ++e) F.Pc[(size_t)e * P + g] = 0.0;
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) F.Wc[(size_t)e * P + g] = 0.0;
#pragma unroll UNR
      for (int e = 0; e < NU * NU; ++e) F.Zc[(size_t)e * P + g] = 0.0;
    }
    for (int j = N - 1; j >= jmin; --j) {
      // ---- stage inputs
      double xm1[NX], ub[NU];
      signed char act[NU];
      const int sl = rcount % DEPTH;  // ring slot of this stage; refilled at the END of the iteration (see below)
      if (use_ring) {
        mbar_wait(sm.bars + sl, (rcount / DEPTH) & 1);
        const double* b = sm.ring + sl * SLOT + col;
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xm1[e] = (j >= 1) ? b[R_XP + e * 32] : ldp(pb.x0, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = b[R_UP + e * 32];
        if (UBOX) {
          const signed char* ab = reinterpret_cast<const signed char*>(sm.ring + sl * SLOT + R_ACTB) + col;
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) act[t] = (j >= Nc && ALG == 0) ? ab[t * 32] : 0;
        }
        ++rcount;
      } else {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xm1[e] = (j >= 1) ? ld(Xpv, j - 1, NX, e) : ldp(pb.x0, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = ld(Upv, j, NU, e);
        if (UBOX) {
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) act[t] = (j >= Nc && ALG == 0) ? pb.act[ix(j, NU, t)] : 0;
        }
      }
      constexpr bool SPD = DUB && !SLEW && CAN_SPARSE;  // sparse Dubins stage (lean Dubins variant)
      double A[NX * NX], Bm[NX * NU], cv[NX], L[NLIN];
      if constexpr (SPD) {
        lin_sparse(pb, j, xm1, ub, dp, L);
        if (valid) {
#pragma unroll
          for (int e = 0; e < NLINH; ++e) at(pb.hand, j, NH, H_LIN + e) = L[e < 8 ? e : e + 2];
        }
#pragma unroll
        for (int e = 0; e < NX; ++e) cv[e] = L[10 + e];
        if (j < Nc) lin_expand(j, L, A, Bm, cv);
      } else {
        lin_stage<DUB>(pb, j, xm1, ub, dp, A, Bm, cv, wb, col);
      }
      // proximal / reference linear terms (lqp_utils.jl:164-214)
      double q[NX], r[NU];
      cost_field(sm.tab, j, T_QR, NX, pb.qref, q, wb, col);
      cost_field(sm.tab, j, T_RR, NU, pb.rref, r, wb, col);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) q[e] = -q[e] - pb.reg_x * xq[e];
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) r[e] = -r[e] - pb.reg_u * ub[e];
      double sig[NU];
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) sig[a] = 0.0;
      if (ALG == 1 && j >= Nc) {
        // interior point: (lazily) updated iterate, barrier diagonal and linear term of this stage
        double lo[NU], hi[NU];
        cost_field(sm.tab, j, T_LU, NU, pb.lu, lo, wb, col);
        cost_field(sm.tab, j, T_UU, NU, pb.uu, hi, wb, col);
        const double* Unew = flip ? pb.Up : pb.U;  // u+ of the previous interior-point iteration (fsweep's Uc)
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          const size_t o = ix(j, NU, a);
          double u = 0.0, lh = 0.0, ll = 0.0, u0 = 0.0, un = 0.0;
          if (!ip.init) {
            u0 = pb.wu[o]; lh = pb.yu[o]; ll = pb.wx[o]; un = Unew[o];
          }
          ipm_point(ip, u0, un, lo[a], hi[a], ub[a], u, lh, ll);
          if (valid) {
            pb.wu[o] = u; pb.yu[o] = lh; pb.wx[o] = ll;
          }
          const double sh = hi[a] - u, sl = u - lo[a];
          sig[a] = lh / sh + ll / sl;
          r[a] += -sig[a] * u + ip.mu / sh - ip.mu / sl;
        }
      }
      if (j >= Nc) {
        // ---- Riccati step with the stage's active set.  S = Qs + P (symmetric, packed)
        double S[NQ];
        cost_field(sm.tab, j, T_QS, NQ, pb.Qs, S, wb, col);
#pragma unroll UNR
        for (int e = 0; e < NQ; ++e) S[e] += Pm[e];
        // wv = S c + q + p
        double wv[NX];
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr) {
          double acc = q[rr] + p[rr];
#pragma unroll UNR
          for (int k = 0; k < NX; ++k) acc += S[rr <= k ? pk(rr, k, NX) : pk(k, rr, NX)] * cv[k];
          wv[rr] = acc;
        }
        // SBW = S B + W
        double SBW[NX * NU];
        if constexpr (SPD) {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const double s0 = S[rr <= 0 ? pk(rr, 0, 4) : pk(0, rr, 4)], s1 = S[rr <= 1 ? pk(rr, 1, 4) : pk(1, rr, 4)];
            const double s2 = S[rr <= 2 ? pk(rr, 2, 4) : pk(2, rr, 4)], s3 = S[rr <= 3 ? pk(rr, 3, 4) : pk(3, rr, 4)];
            SBW[rr * 2 + 0] = s0 * L[4] + s1 * L[6] + s2 * L[8];
            SBW[rr * 2 + 1] = s0 * L[5] + s1 * L[7] + s3 * L[9];
          }
        } else {
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) {
              double acc = SLEW ? W[rr * NU + t] : 0.0;
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) acc += S[rr <= k ? pk(rr, k, NX) : pk(k, rr, NX)] * Bm[k * NU + t];
              SBW[rr * NU + t] = acc;
            }
        }
        // H = Rs + [j>=1] s I + B'(SB + W) + W'B + Z ;  h0 = r + B'wv + W'c + z
        double H[NU * NU], h0[NU];
        if constexpr (SPD) {
          double Rp[NR];
          cost_field(sm.tab, j, T_RS, NR, pb.Rs, Rp, wb, col);
          H[0] = Rp[0] + (ALG == 1 ? sig[0] : 0.0) + L[4] * SBW[0] + L[6] * SBW[2] + L[8] * SBW[4];
          H[1] = H[2] = Rp[1] + L[4] * SBW[1] + L[6] * SBW[3] + L[8] * SBW[5];
          H[3] = Rp[2] + (ALG == 1 ? sig[1] : 0.0) + L[5] * SBW[1] + L[7] * SBW[3] + L[9] * SBW[7];
          h0[0] = r[0] + L[4] * wv[0] + L[6] * wv[1] + L[8] * wv[2];
          h0[1] = r[1] + L[5] * wv[0] + L[7] * wv[1] + L[9] * wv[3];
        } else {
          double Rp[NR];
          cost_field(sm.tab, j, T_RS, NR, pb.Rs, Rp, wb, col);
#pragma unroll UNR
          for (int a = 0; a < NU; ++a)
#pragma unroll UNR
            for (int b = a; b < NU; ++b) {
              double acc = Rp[pk(a, b, NU)];
              if (ALG == 1 && a == b) acc += sig[a];
              if (SLEW) {
                acc += Z[a * NU + b];
                if (a == b && j >= 1) acc += s;
              }
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) {
                acc += Bm[k * NU + a] * SBW[k * NU + b];
                if (SLEW) acc += W[k * NU + a] * Bm[k * NU + b];
              }
              H[a * NU + b] = acc;
              H[b * NU + a] = acc;
            }
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) {
            double acc = r[a] + (SLEW ? z[a] : 0.0);
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) {
              acc += Bm[k * NU + a] * wv[k];
              if (SLEW) acc += W[k * NU + a] * cv[k];
            }
            h0[a] = acc;
          }
        }
        bool fixed[NU];
        double ufix[NU], E[NU * NU];
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) {
          fixed[t] = false;
          ufix[t] = 0.0;
        }
        if (UBOX) {
          double lo[NU], hi[NU];
          cost_field(sm.tab, j, T_LU, NU, pb.lu, lo, wb, col);
          cost_field(sm.tab, j, T_UU, NU, pb.uu, hi, wb, col);
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) {
            fixed[t] = act[t] != 0;
            ufix[t] = (act[t] > 0) ? hi[t] : ((act[t] < 0) ? lo[t] : 0.0);
          }
        }
        masked_spd_inverse<NU>(H, fixed, E);
        // h = h0 + H ufix ; k = ufix - E h ; eta = wv + SBW k_full ... (the value-function recursion uses the
        // closed-loop form below, so only h and k are needed)
        double h[NU], kv[NU];
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = h0[a];
          if (UBOX) {
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) acc += H[a * NU + t] * ufix[t];
          }
          h[a] = acc;
        }
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = ufix[a];
#pragma unroll UNR
          for (int b = 0; b < NU; ++b) acc -= E[a * NU + b] * h[b];
          kv[a] = acc;
        }
        double K[NU * NX], G[NU * NX];
        if (j >= 1) {
          // G = (SB + W)' A ;  K = -E G
          if constexpr (SPD) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              G[a * 4 + 0] = SBW[0 * 2 + a];
              G[a * 4 + 1] = SBW[1 * 2 + a];
              G[a * 4 + 2] = SBW[2 * 2 + a] + SBW[0 * 2 + a] * L[0] + SBW[1 * 2 + a] * L[2];
              G[a * 4 + 3] = SBW[3 * 2 + a] + SBW[0 * 2 + a] * L[1] + SBW[1 * 2 + a] * L[3];
            }
          } else {
#pragma unroll UNR
            for (int a = 0; a < NU; ++a)
#pragma unroll UNR
              for (int t = 0; t < NX; ++t) {
                double acc = 0.0;
#pragma unroll UNR
                for (int k = 0; k < NX; ++k) acc += SBW[k * NU + a] * A[k * NX + t];
                G[a * NX + t] = acc;
              }
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
          // p <- A'(wv + SBW k) + G'... : with u = K x + k the cost-to-go gradient is
          //   p' = A' wv + G' k   (+ K'(h0 + H k + ...) = 0 on free comps, and fixed comps have K = 0)
          double pn[NX], Pn[NQ];
          if constexpr (SPD) {
            pn[0] = wv[0]; pn[1] = wv[1];
            pn[2] = wv[2] + L[0] * wv[0] + L[2] * wv[1];
            pn[3] = wv[3] + L[1] * wv[0] + L[3] * wv[1];
#pragma unroll
            for (int t = 0; t < 4; ++t) pn[t] += G[0 * 4 + t] * kv[0] + G[1 * 4 + t] * kv[1];
            // A'SA with A = I + N:  T = S[:,0:2] N2 ; U = S[0:2,2:4] + T[0:2,:] ; lower-right += T[2:4,:] + N2'U
            double Tm[4 * 2], Um[2 * 2];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const double s0 = S[rr <= 0 ? pk(rr, 0, 4) : pk(0, rr, 4)], s1 = S[rr <= 1 ? pk(rr, 1, 4) : pk(1, rr, 4)];
              Tm[rr * 2 + 0] = s0 * L[0] + s1 * L[2];
              Tm[rr * 2 + 1] = s0 * L[1] + s1 * L[3];
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
              for (int c = 0; c < 2; ++c) Um[rr * 2 + c] = S[pk(rr, 2 + c, 4)] + Tm[rr * 2 + c];
            Pn[pk(0, 0, 4)] = S[pk(0, 0, 4)]; Pn[pk(0, 1, 4)] = S[pk(0, 1, 4)]; Pn[pk(1, 1, 4)] = S[pk(1, 1, 4)];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
              for (int c = 0; c < 2; ++c) Pn[pk(rr, 2 + c, 4)] = Um[rr * 2 + c];
            // (2+d, 2+c), d <= c:  S + T[2+d][c] + N2[0][d] U[0][c] + N2[1][d] U[1][c]   (N2[0][.] = L0,L1 ; N2[1][.] = L2,L3)
            Pn[pk(2, 2, 4)] = S[pk(2, 2, 4)] + Tm[2 * 2 + 0] + L[0] * Um[0] + L[2] * Um[2];
            Pn[pk(2, 3, 4)] = S[pk(2, 3, 4)] + Tm[2 * 2 + 1] + L[0] * Um[1] + L[2] * Um[3];
            Pn[pk(3, 3, 4)] = S[pk(3, 3, 4)] + Tm[3 * 2 + 1] + L[1] * Um[1] + L[3] * Um[3];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
              for (int rr = 0; rr <= t; ++rr) Pn[pk(rr, t, 4)] += G[0 * 4 + rr] * K[0 * 4 + t] + G[1 * 4 + rr] * K[1 * 4 + t];
          } else {
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int rr = 0; rr < NX; ++rr) acc += A[rr * NX + t] * wv[rr];
#pragma unroll UNR
            for (int a = 0; a < NU; ++a) acc += G[a * NX + t] * kv[a];
            pn[t] = acc;
          }
          // P <- A' S A + G' K   (column t of S A at a time keeps the live set small)
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double sa[NX];
#pragma unroll UNR
            for (int rr = 0; rr < NX; ++rr) {
              double acc = 0.0;
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) acc += S[rr <= k ? pk(rr, k, NX) : pk(k, rr, NX)] * A[k * NX + t];
              sa[rr] = acc;
            }
#pragma unroll UNR
            for (int rr = 0; rr <= t; ++rr) {
              double acc = 0.0;
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) acc += A[k * NX + rr] * sa[k];
#pragma unroll UNR
              for (int a = 0; a < NU; ++a) acc += G[a * NX + rr] * K[a * NX + t];
              Pn[pk(rr, t, NX)] = acc;
            }
          }
          }
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) p[t] = pn[t];
#pragma unroll UNR
          for (int e = 0; e < NQ; ++e) Pm[e] = Pn[e];
          if (SLEW) {
#pragma unroll UNR
            for (int a = 0; a < NU; ++a) z[a] = -s * kv[a];
          }
        } else {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) { K[e] = 0.0; G[e] = 0.0; }
        }
        if (SLEW) {
          // value-function slew blocks for the previous stage: W = -s K', Z = s I - s^2 E
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) W[rr * NU + t] = -s * K[t * NX + rr];
#pragma unroll UNR
          for (int a = 0; a < NU; ++a)
#pragma unroll UNR
            for (int b = 0; b < NU; ++b) Z[a * NU + b] = (a == b ? s : 0.0) - s * s * E[a * NU + b];
        }
        if (valid) {
          // fixed controls: the (otherwise 0 / bound) slots carry G[a,:], h0[a] for the multiplier
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) {
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) at(pb.hand, j, NH, H_K + a * NX + t) = (UBOX && fixed[a]) ? G[a * NX + t] : K[a * NX + t];
            at(pb.hand, j, NH, H_KF + a) = (UBOX && fixed[a]) ? h[a] : kv[a];  // h = h0 + H ufix
          }
          if (UBOX) {
#pragma unroll UNR
            for (int a = 0; a < NU; ++a)
#pragma unroll UNR
              for (int b = a + 1; b < NU; ++b) at(pb.hand, j, NH, H_H + pko(a, b)) = H[a * NU + b];
          }
          if (SLEW) {
#pragma unroll UNR
            for (int a = 0; a < NU; ++a)
#pragma unroll UNR
              for (int b = a; b < NU; ++b) at(F.E, j, NR, pk(a, b, NU)) = E[a * NU + b];
          }
        }
        if (j == Nc && Nc >= 1) {
          if (valid) {
#pragma unroll UNR
            for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
              for (int t = 0; t < NX; ++t)
                F.Pc[(size_t)(rr * NX + t) * P + g] = Pm[rr <= t ? pk(rr, t, NX) : pk(t, rr, NX)];
#pragma unroll UNR
            for (int e = 0; e < NX * NU; ++e) F.Wc[(size_t)e * P + g] = SLEW ? W[e] : 0.0;
#pragma unroll UNR
            for (int e = 0; e < NU * NU; ++e) F.Zc[(size_t)e * P + g] = SLEW ? Z[e] : 0.0;
          }
#pragma unroll UNR
          for (int e = 0; e < NX; ++e) pc[e] = p[e];
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) zc[e] = SLEW ? z[e] : 0.0;
        }
      } else if (valid) {
        // shared stages: the consensus Schur step (build_shared_system / solve_shared) reads these
#pragma unroll UNR
        for (int e = 0; e < NX * NX; ++e) at(pb.A, j, NX * NX, e) = A[e];
#pragma unroll UNR
        for (int e = 0; e < NX * NU; ++e) at(pb.Bm, j, NX * NU, e) = Bm[e];
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) {
          at(pb.c, j, NX, e) = cv[e];
          at(pb.q, j, NX, e) = q[e];
        }
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) at(pb.r, j, NU, e) = r[e];
      }
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) xq[e] = xm1[e];
      // Refill the slot only now: the bulk copy (async proxy) overwrites shared memory the lanes read with
      // plain LDS above, and a converged __syncwarp does not wait for outstanding LDS to complete.  At
      // this point every loaded value has been consumed, so the write-after-read hazard is closed by
      // data dependence (issuing right after the loads raced once the LSU queue was backed up by the
      // stage's global stores).  The prefetch distance is still DEPTH - 1 full stages.
      if (use_ring) {
        __syncwarp();
        if (lane == 0 && j - DEPTH >= jmin) issue_b<UBOX>(pb, sm, Xpv, Upv, j - DEPTH, sl, wb);
      }
    }
    if (use_ring) fence_proxy_async();  // K, k, H written above are bulk-copied by the next sweep
    this->rcount = rcount;
  }

  // Forward sweep: rollout of the feedback policy, multipliers + primal-dual active-set update of the
  // free-stage controls, SCP bookkeeping against (Xpv, Upv).  Fills the per-thread statistics.
  template <bool SLEW, bool UBOX, bool DUB, int ALG = 0>
  __device__ __forceinline__ void fsweep(const DevProb& pb, unsigned char* smem, double* st, const Ipm ip = Ipm(),
                                         double* ipacc = nullptr) {
    // sweep-local copies of the per-thread invariants (see ixw): registers instead of local-memory reloads
    const int N = this->N, Nc = this->Nc, g = this->g, lane = this->lane, col = this->col;
    const size_t P = this->P;
    const bool valid = this->valid, flip = this->flip;
    const double s = this->s;
    const size_t wb = (size_t)(g / TILE) * (size_t)N;
    uint32_t rcount = this->rcount;
    auto ix = [=](int j, int E, int e) -> size_t { return ixw(wb, col, j, E, e); };
    auto ld = [=](const double* a, int j, int E, int e) -> double { return a[ixw(wb, col, j, E, e)]; };
    auto at = [=](double* a, int j, int E, int e) -> double& { return a[ixw(wb, col, j, E, e)]; };
    auto ldp = [=](const double* a, int e) -> double { return a[(size_t)e * P + g]; };
    (void)ix; (void)ld; (void)at; (void)ldp; (void)s; (void)lane;
    const FactorSet& F = pb.fs[MODE_POLISH];
    const Sm sm = smem_view(pb, smem);
    const bool use_ring = RING_SHAPE && sm.ring != nullptr;
    constexpr bool SPD = DUB && !SLEW && CAN_SPARSE;  // the backward sweep left the sparse stage in pb.lin
    const double* const Xpv = flip ? pb.X : pb.Xp;
    const double* const Upv = flip ? pb.U : pb.Up;
    double* const Xc = flip ? pb.Xp : pb.X;
    double* const Uc = flip ? pb.Up : pb.U;
    double dp[3] = {0.0, 0.0, 0.0};
    if (DUB) {
      dp[0] = ldp(pb.dynp, 0); dp[1] = ldp(pb.dynp, 1); dp[2] = ldp(pb.dynp, 2);
    }
    if (use_ring) {
      __syncwarp();
      if (lane == 0) {
        for (int d = 0; d < DEPTH; ++d)
          if (d < N) issue_f<UBOX, SPD>(pb, sm, Xpv, Upv, d, (rcount + d) % DEPTH, wb, Nc);
      }
    }
    const double* uc = (GROUPED && Nc >= 1) ? pb.uc + (size_t)prob * D : nullptr;
    double x[NX], up[NU], xm1[NX];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) { x[e] = 0.0; xm1[e] = ldp(pb.x0, e); }
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) up[e] = 0.0;
    double dxm = 0.0, dum = 0.0, obj = 0.0, nanf = 0.0, kkt_v = 0.0, uviol = 0.0, xviol = 0.0, changed = 0.0,
           gscale = 1.0, zscale = 1.0;
    const double rdiag = pb.rdiag0[g];
    for (int j = 0; j < N; ++j) {
      double ub[NU], xpj[NX], Kj[NU * NX], kf[NU], Hj[NHO > 0 ? NHO : 1], u[NU], xn[NX], L[NLIN];
      signed char act[NU];
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) act[t] = 0;
      const int sl = rcount % DEPTH;  // refilled at the end of the iteration (write-after-read, see bsweep)
      if (use_ring) {
        mbar_wait(sm.bars + sl, (rcount / DEPTH) & 1);
        const double* b = sm.ring + sl * SLOT + col;
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xpj[e] = b[R_XP + e * 32];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = b[R_UP + e * 32];
        if (SPD) {
#pragma unroll
          for (int e = 0; e < NLINH; ++e) L[e < 8 ? e : e + 2] = b[R_LIN + e * 32];
        }
        if (j >= Nc) {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) Kj[e] = b[R_K + e * 32];
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) kf[e] = b[R_KF + e * 32];
          if (UBOX) {
#pragma unroll UNR
            for (int e = 0; e < NHO; ++e) Hj[e] = b[R_H + e * 32];
            const signed char* ab = reinterpret_cast<const signed char*>(sm.ring + sl * SLOT + R_ACTF) + col;
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) act[t] = (ALG == 0) ? ab[t * 32] : 0;
          }
        }
        ++rcount;
      } else {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xpj[e] = ld(Xpv, j, NX, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = ld(Upv, j, NU, e);
        if (SPD) {
#pragma unroll
          for (int e = 0; e < NLINH; ++e) L[e < 8 ? e : e + 2] = ld(pb.hand, j, NH, H_LIN + e);
        }
        if (j >= Nc) {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) Kj[e] = ld(pb.hand, j, NH, H_K + e);
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) kf[e] = ld(pb.hand, j, NH, H_KF + e);
          if (UBOX) {
#pragma unroll UNR
            for (int e = 0; e < NHO; ++e) Hj[e] = ld(pb.hand, j, NH, H_H + e);
#pragma unroll UNR
            for (int t = 0; t < NU; ++t) act[t] = (ALG == 0) ? pb.act[ix(j, NU, t)] : 0;
          }
        }
      }
      if constexpr (SPD) {  // the four entries of the sparse stage that are not handed over (dynamics.cuh)
        dubins_cheap_entries(xm1, ub, dp[0], dp[1], dp[2], pb.dyn_eps, j >= 1, &L[8], &L[9], &L[12], &L[13]);
      }
      double lo[NU], hi[NU];
      if (UBOX) {
        cost_field(sm.tab, j, T_LU, NU, pb.lu, lo, wb, col);
        cost_field(sm.tab, j, T_UU, NU, pb.uu, hi, wb, col);
      }
      if (j < Nc) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) u[a] = uc[j * NU + a];
      } else {
        // policy on the free controls, bound on the fixed ones; then the multipliers of the fixed ones
        double gx[NU];  // K x (free) or G x (fixed)
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = 0.0;
          if (j >= 1) {
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) acc += Kj[a * NX + t] * x[t];
          }
          gx[a] = acc;
        }
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = kf[a] + gx[a];
          if (SLEW && j >= 1) {
#pragma unroll UNR
            for (int b = 0; b < NU; ++b) acc += s * ld(F.E, j, NR, a <= b ? pk(a, b, NU) : pk(b, a, NU)) * up[b];
          }
          u[a] = acc;
          if (UBOX && act[a] != 0) u[a] = (act[a] > 0) ? hi[a] : lo[a];
        }
        if (ALG == 1) {
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) {
            const size_t o = ix(j, NU, a);
            ipm_ratios(ip, pb.wu[o], u[a], lo[a], hi[a], pb.yu[o], pb.wx[o], ipacc);
          }
        }
        if (UBOX && ALG == 0) {
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) {
            const double ua = u[a];
            signed char na = 0;
            if (act[a] != 0) {
              // dQ/du_a = H[a,:] u + G[a,:] x + h0[a] - s u_{j-1,a};  kf[a] already holds h0[a] + H[a, fixed] u_fixed
              double acc = kf[a] + gx[a];
#pragma unroll UNR
              for (int b = 0; b < NU; ++b)
                if (b != a && act[b] == 0) acc += Hj[a < b ? pko(a, b) : pko(b, a)] * u[b];
              if (SLEW && j >= 1) acc -= s * up[a];
              const double mu = -acc;
              if (act[a] > 0) {
                kkt_v = fmax(kkt_v, fmax(acc, 0.0));
                if (mu > 0.0) na = 1;
              } else {
                kkt_v = fmax(kkt_v, fmax(-acc, 0.0));
                if (mu < 0.0) na = -1;
              }
            } else {
              uviol = fmax(uviol, fmax(ua - hi[a], lo[a] - ua));
              if (ua - hi[a] > 0.0 && hi[a] < INFINITY) na = 1;
              else if (ua - lo[a] < 0.0 && lo[a] > -INFINITY) na = -1;
            }
            if (na != act[a]) {
              changed += 1.0;
              if (valid) pb.act[ix(j, NU, a)] = na;
            }
          }
        }
      }
      if constexpr (SPD) {
        // x+ = (I + N) x + B u + c with the sparse stage (stage 0: A = 0)
        xn[0] = L[10] + L[4] * u[0] + L[5] * u[1];
        xn[1] = L[11] + L[6] * u[0] + L[7] * u[1];
        xn[2] = L[12] + L[8] * u[0];
        xn[3] = L[13] + L[9] * u[1];
        if (j >= 1) {
          xn[0] += x[0] + L[0] * x[2] + L[1] * x[3];
          xn[1] += x[1] + L[2] * x[2] + L[3] * x[3];
          xn[2] += x[2];
          xn[3] += x[3];
        }
      } else {
        double A[NX * NX], Bm[NX * NU], cv[NX];
        lin_stage<DUB>(pb, j, xm1, ub, dp, A, Bm, cv, wb, col);
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) {
          double acc = cv[r];
          if (j >= 1) {
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) acc += A[r * NX + t] * x[t];
          }
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += Bm[r * NU + t] * u[t];
          xn[r] = acc;
        }
      }
      if (valid) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) at(Uc, j, NU, a) = u[a];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) at(Xc, j, NX, r) = xn[r];
      }
      if (pb.has_xbox) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) xviol = fmax(xviol, fmax(xn[r] - ld(pb.ux, j, NX, r), ld(pb.lx, j, NX, r) - xn[r]));
      }
      // bookkeeping (scp_mpc.py:397-405) and the scales of the acceptance test
      double Qp[NQ], Rp[NR], xr[NX], ur[NU], qr[NX], rr_[NU];
      cost_field(sm.tab, j, T_QS, NQ, pb.Qs, Qp, wb, col);
      cost_field(sm.tab, j, T_RS, NR, pb.Rs, Rp, wb, col);
      cost_field(sm.tab, j, T_XR, NX, pb.xref, xr, wb, col);
      cost_field(sm.tab, j, T_UR, NU, pb.uref, ur, wb, col);
      cost_field(sm.tab, j, T_QR, NX, pb.qref, qr, wb, col);
      cost_field(sm.tab, j, T_RR, NU, pb.rref, rr_, wb, col);
      double dx[NX], du[NU], n2 = 0.0;
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        const double d = xn[r] - xpj[r];
        n2 += d * d;
        dx[r] = xn[r] - xr[r];
        if (!(xn[r] == xn[r])) nanf = 1.0;
        zscale = fmax(zscale, fabs(xn[r]));
        gscale = fmax(gscale, fabs(qr[r] + pb.reg_x * xpj[r]));
      }
      dxm = fmax(dxm, n2);  // squared per-stage norms; one sqrt after the sweep (sqrt is monotone and correctly rounded)
      n2 = 0.0;
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        const double d = u[t] - ub[t];
        n2 += d * d;
        du[t] = u[t] - ur[t];
        if (!(u[t] == u[t])) nanf = 1.0;
        zscale = fmax(zscale, fabs(u[t]));
        gscale = fmax(gscale, fabs(rr_[t] + pb.reg_u * ub[t]));
      }
      dum = fmax(dum, n2);
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += Qp[r <= t ? pk(r, t, NX) : pk(t, r, NX)] * dx[t];
        obj += dx[r] * (acc - pb.reg_x * dx[r]);
      }
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += Rp[b <= t ? pk(b, t, NU) : pk(t, b, NU)] * du[t];
        obj += du[b] * (acc - (pb.reg_u + (j == 0 ? rdiag : 0.0)) * du[b]);
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) { x[r] = xn[r]; xm1[r] = xpj[r]; }
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) up[a] = u[a];
      if (use_ring) {
        __syncwarp();
        if (lane == 0 && j + DEPTH < N) issue_f<UBOX, SPD>(pb, sm, Xpv, Upv, j + DEPTH, sl, wb, Nc);
      }
    }
    if (use_ring) fence_proxy_async();  // X, U, act written above are bulk-copied by the next sweep
    this->rcount = rcount;
    st[ST_DX] = valid ? sqrt(dxm) : 0.0;
    st[ST_DU] = valid ? sqrt(dum) : 0.0;
    st[ST_OBJ] = valid ? obj : 0.0;
    st[ST_NAN] = nanf;
    st[ST_KKT] = kkt_v; st[ST_UVIOL] = uviol; st[ST_XVIOL] = xviol; st[ST_CHANGED] = valid ? changed : 0.0;
    st[ST_GSCALE] = gscale; st[ST_ZSCALE] = zscale;
  }

  // Shared block: multipliers of the fixed shared controls from the assembled system,
  // mu = -(Hc uc + gc), and the primal-dual update of its active set.  Returns the number of changes
  // (identical in every thread); adds the block's KKT / bound violation into kk, uv.
  __device__ __forceinline__ double shared_active_update(double* kk, double* uv) {
    if (!GROUPED || Nc < 1) return 0.0;
    const FactorSet& F = pb.fs[MODE_POLISH];
    const double* Hc = F.Hc + (size_t)prob * D * D;
    const double* gc = pb.gc + (size_t)prob * D;
    const double* uc = pb.uc + (size_t)prob * D;
    signed char* actc = pb.actc + (size_t)prob * D;
    signed char* actcn = pb.actcn + (size_t)prob * D;
    const size_t g0 = (size_t)prob * pb.Mpad;
    double ch = 0.0;
    for (int k = 0; k < D; ++k) {  // every thread evaluates the (small) shared block identically
      const int jj = k / NU, b = k % NU;
      const double ubv = uc[k];
      double lo = -INFINITY, hi = INFINITY;
      if (pb.has_ubox) {
        lo = pb.lu[ixg(jj, NU, b, g0)];
        hi = pb.uu[ixg(jj, NU, b, g0)];
      }
      signed char na = 0;
      if (actc[k] != 0) {
        double acc = gc[k];
        for (int l = 0; l < D; ++l) acc += Hc[(size_t)k * D + l] * uc[l];
        if (actc[k] > 0) {
          *kk = fmax(*kk, fmax(acc, 0.0));
          if (-acc > 0.0) na = 1;
        } else {
          *kk = fmax(*kk, fmax(-acc, 0.0));
          if (-acc < 0.0) na = -1;
        }
      } else {
        *uv = fmax(*uv, fmax(ubv - hi, lo - ubv));
        if (ubv - hi > 0.0 && hi < INFINITY) na = 1;
        else if (ubv - lo < 0.0 && lo > -INFINITY) na = -1;
      }
      if (na != actc[k]) ch += 1.0;
      if (tid == 0) actcn[k] = na;
    }
    __syncthreads();  // every thread has read actc
    for (int k = tid; k < D; k += T) actc[k] = actcn[k];
    __syncthreads();
    return ch;
  }

  // Shared block of D <= DSMALL controls (C2: D = 2) solved in registers.  Same mathematics as
  // assemble_shared / grad_shared / factor_shared / apply_shared / shared_active_update — sum_i H_i and sum_i g_i by
  // adjoint passes over the shared stages, then the primal-dual active-set iteration of the D-dimensional box QP —
  // but with ONE small CTA reduction (cta_reduce) and the whole D x D iteration replicated in every thread, instead
  // of ~25 barriers and as many round trips through the problem's global scratch per pass.
  // Returns the number of active-set changes of the last iteration; kk / uv = KKT and bound violation of the block.
  template <bool UBOX>
  __device__ __noinline__ double shared_small(const double* pc, const double* zc, double* kk, double* uv) {
    const FactorSet& F = pb.fs[MODE_POLISH];
    const int Dn = D;
    double hv[DSMALL * DSMALL + DSMALL];
    if (Nc == 1 && NU <= DSMALL) {
      // One shared stage (the MPC-typical case): H_i and g_i in closed form from ONE read of the stage-0 data and of the
      // particle's boundary value function — the generic adjoint passes below evaluate the same expressions D + 1
      // times through scratch stages of X.  With S0 = Qs_0 + Pc (stage 0 has no A: lqp_utils.jl:288-296):
      //   H_i = Rs_0 + Zc + B0' (S0 B0 + Wc) + Wc' B0         g_i = r_0 + zc + Wc' c0 + B0' (q_0 + S0 c0 + pc)
      double B0[NX * NU], S0[NX * NX], Wc[NX * NU], c0[NX], an[NX];
#pragma unroll UNR
      for (int e = 0; e < NX * NU; ++e) { B0[e] = ld(pb.Bm, 0, NX * NU, e); Wc[e] = F.Wc[(size_t)e * P + g]; }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        c0[r] = ld(pb.c, 0, NX, r);
#pragma unroll UNR
        for (int t = 0; t < NX; ++t)
          S0[r * NX + t] = ld(pb.Qs, 0, NQ, r <= t ? pk(r, t, NX) : pk(t, r, NX)) + F.Pc[(size_t)(r * NX + t) * P + g];
      }
#pragma unroll UNR
      for (int k = 0; k < NU; ++k) {
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) {
          double acc = Wc[r * NU + k];
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) acc += S0[r * NX + t] * B0[t * NU + k];
          an[r] = acc;
        }
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) {
          double acc = ld(pb.Rs, 0, NR, b <= k ? pk(b, k, NU) : pk(k, b, NU)) + F.Zc[(size_t)(b * NU + k) * P + g];
#pragma unroll UNR
          for (int r = 0; r < NX; ++r) acc += Wc[r * NU + b] * B0[r * NU + k] + B0[r * NU + b] * an[r];
          hv[k * Dn + b] = valid ? acc : 0.0;
        }
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = ld(pb.q, 0, NX, r) + pc[r];
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += S0[r * NX + t] * c0[t];
        an[r] = acc;
      }
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = ld(pb.r, 0, NU, b) + zc[b];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) acc += Wc[r * NU + b] * c0[r] + B0[r * NU + b] * an[r];
        hv[Dn * Dn + b] = valid ? acc : 0.0;
      }
    } else {
      for (int k = 0; k < Dn; ++k) shared_grad(MODE_POLISH, true, k, nullptr, 0.0, 0.0, nullptr, nullptr, 0, &hv[k * Dn]);
      shared_grad(MODE_POLISH, false, -1, nullptr, 0.0, 0.0, pc, zc, 0, &hv[Dn * Dn]);
    }
    cta_reduce(hv, Dn * Dn + Dn, 0xffffffffu);
    double H[DSMALL * DSMALL], gcv[DSMALL], lo[DSMALL], hi[DSMALL], ucv[DSMALL];
    signed char act[DSMALL];
    const size_t g0 = (size_t)prob * pb.Mpad;  // particle 0 of the problem owns the shared bounds (lqp_utils.jl:329-330)
    for (int a = 0; a < Dn; ++a) {
      for (int b = 0; b < Dn; ++b) H[a * DSMALL + b] = 0.5 * (hv[b * Dn + a] + hv[a * Dn + b]);  // column k of sum H_i sits in hv[k*D ..]
      gcv[a] = hv[Dn * Dn + a];
      lo[a] = -INFINITY; hi[a] = INFINITY;
      if (pb.has_ubox) {
        lo[a] = pb.lu[ixg(a / NU, NU, a % NU, g0)];
        hi[a] = pb.uu[ixg(a / NU, NU, a % NU, g0)];
      }
      act[a] = (UBOX && pb.has_ubox) ? pb.actc[(size_t)prob * Dn + a] : 0;
      ucv[a] = 0.0;
    }
    double ch = 0.0;
    for (int t = 0; t <= 2 * Dn; ++t) {
      // masked system: fixed components at their bound beta, free ones solve  H_ff u_f = -(g_f + H_fb beta)
      double L[DSMALL * DSMALL], rhs[DSMALL];
      for (int a = 0; a < Dn; ++a) {
        double acc = gcv[a];
        for (int b = 0; b < Dn; ++b)
          if (act[b] != 0) acc += H[a * DSMALL + b] * ((act[b] > 0) ? hi[b] : lo[b]);
        rhs[a] = -acc;
        for (int b = 0; b < Dn; ++b) {
          const bool fx = act[a] != 0 || act[b] != 0;
          L[a * DSMALL + b] = fx ? (a == b ? 1.0 : 0.0) : H[a * DSMALL + b];
        }
      }
      for (int k = 0; k < Dn; ++k) {  // Cholesky (lower) of the masked matrix
        double dk = L[k * DSMALL + k];
        for (int l = 0; l < k; ++l) dk -= L[k * DSMALL + l] * L[k * DSMALL + l];
        dk = sqrt(dk);
        L[k * DSMALL + k] = dk;
        for (int i = k + 1; i < Dn; ++i) {
          double v = L[i * DSMALL + k];
          for (int l = 0; l < k; ++l) v -= L[i * DSMALL + l] * L[k * DSMALL + l];
          L[i * DSMALL + k] = v / dk;
        }
      }
      double y[DSMALL];
      for (int i = 0; i < Dn; ++i) {
        double v = (act[i] != 0) ? 0.0 : rhs[i];
        for (int l = 0; l < i; ++l) v -= L[i * DSMALL + l] * y[l];
        y[i] = v / L[i * DSMALL + i];
      }
      for (int i = Dn - 1; i >= 0; --i) {
        double v = y[i];
        for (int l = i + 1; l < Dn; ++l) v -= L[l * DSMALL + i] * y[l];
        y[i] = v / L[i * DSMALL + i];
      }
      for (int a = 0; a < Dn; ++a) ucv[a] = (act[a] != 0) ? ((act[a] > 0) ? hi[a] : lo[a]) : y[a];
      if (!(UBOX && pb.has_ubox)) break;
      // multipliers of the fixed components, violations of the free ones, primal-dual update (shared_active_update)
      *kk = 0.0; *uv = 0.0; ch = 0.0;
      signed char nact[DSMALL];
      for (int k = 0; k < Dn; ++k) {
        signed char na = 0;
        if (act[k] != 0) {
          double acc = gcv[k];
          for (int l = 0; l < Dn; ++l) acc += H[k * DSMALL + l] * ucv[l];
          if (act[k] > 0) {
            *kk = fmax(*kk, fmax(acc, 0.0));
            if (-acc > 0.0) na = 1;
          } else {
            *kk = fmax(*kk, fmax(-acc, 0.0));
            if (-acc < 0.0) na = -1;
          }
        } else {
          *uv = fmax(*uv, fmax(ucv[k] - hi[k], lo[k] - ucv[k]));
          if (ucv[k] - hi[k] > 0.0 && hi[k] < INFINITY) na = 1;
          else if (ucv[k] - lo[k] < 0.0 && lo[k] > -INFINITY) na = -1;
        }
        if (na != act[k]) ch += 1.0;
        nact[k] = na;
      }
      for (int k = 0; k < Dn; ++k) act[k] = nact[k];
      if (ch == 0.0) break;
    }
    if (hl == 0 && pvalid && (PPW == 1 || wr_en)) {  // every thread holds the same values; the forward sweep and later passes read them from HBM
      for (int a = 0; a < Dn; ++a) {
        pb.uc[(size_t)prob * Dn + a] = ucv[a];
        pb.gc[(size_t)prob * Dn + a] = gcv[a];
        pb.actc[(size_t)prob * Dn + a] = act[a];
        for (int b = 0; b < Dn; ++b) F.Hc[(size_t)prob * Dn * Dn + (size_t)a * Dn + b] = H[a * DSMALL + b];
      }
    }
    if (PPW == 2) { __threadfence_block(); __syncwarp(); } else __syncthreads();  // (PPW 2: a problem lives in a half-warp)
    return ch;
  }

  // ---- interior-point fallback: shared block and driver ----------------------------------------
  // before the shared solve: lazy step / initial point of the shared controls, barrier terms into Hc, gc
  __device__ __forceinline__ void ipm_shared_pre(const DevProb& pb, const Ipm& ip) {
    if (!GROUPED || Nc < 1) return;
    const FactorSet& F = pb.fs[MODE_POLISH];
    double* Hc = F.Hc + (size_t)prob * D * D;
    double* gc = pb.gc + (size_t)prob * D;
    const double* ucn = pb.uc + (size_t)prob * D;  // u+ of the previous iteration
    const double* Up0 = flip ? pb.U : pb.Up;
    const size_t g0 = (size_t)prob * pb.Mpad;
    for (int k = tid; k < D; k += T) {
      const size_t o = ixg(k / NU, NU, k % NU, g0);
      const double lo = pb.lu[o], hi = pb.uu[o];
      double u = 0.0, lh = 0.0, ll = 0.0, u0 = 0.0;
      if (!ip.init) {
        u0 = pb.wu[o]; lh = pb.yu[o]; ll = pb.wx[o];
      }
      ipm_point(ip, u0, ucn[k], lo, hi, Up0[o], u, lh, ll);
      pb.wu[o] = u; pb.yu[o] = lh; pb.wx[o] = ll;
      const double sh = hi - u, sl = u - lo, sg = lh / sh + ll / sl;
      Hc[(size_t)k * D + k] += sg;
      gc[k] += -sg * u + ip.mu / sh - ip.mu / sl;
    }
    __syncthreads();
  }
  __device__ __forceinline__ void ipm_shared_post(const DevProb& pb, const Ipm& ip, double* acc) {
    if (!GROUPED || Nc < 1) return;
    const double* ucn = pb.uc + (size_t)prob * D;
    const size_t g0 = (size_t)prob * pb.Mpad;
    for (int k = (PPW == 2 ? hl : tid); k < D; k += (PPW == 2 ? 16 : T)) {
      const size_t o = ixg(k / NU, NU, k % NU, g0);
      ipm_ratios(ip, pb.wu[o], ucn[k], pb.lu[o], pb.uu[o], pb.yu[o], pb.wx[o], acc);
    }
  }
  // last step applied, active set read off the complementarity pairs (bound active iff slack < multiplier)
  __device__ __noinline__ void ipm_finish(const DevProb& pb, const Ipm ip) {
    const double* Unew = flip ? pb.Up : pb.U;
    for (int j = Nc; j < N; ++j) {
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) {
        const size_t o = ix(j, NU, a);
        const double lo = pb.lu[o], hi = pb.uu[o];
        double u = 0.0, lh = pb.yu[o], ll = pb.wx[o];
        ipm_point(ip, pb.wu[o], Unew[o], lo, hi, 0.0, u, lh, ll);
        const signed char na = (hi - u < lh) ? 1 : ((u - lo < ll) ? -1 : 0);
        if (valid) pb.act[o] = na;
      }
    }
    if (GROUPED && Nc >= 1) {
      const double* ucn = pb.uc + (size_t)prob * D;
      const size_t g0 = (size_t)prob * pb.Mpad;
      if (PPW == 2) __syncwarp(); else __syncthreads();  // (PPW 2: the problem is a half-warp; CTA barriers are phase barriers there)
      for (int k = (PPW == 2 ? hl : tid); k < D; k += (PPW == 2 ? 16 : T)) {
        const size_t o = ixg(k / NU, NU, k % NU, g0);
        const double lo = pb.lu[o], hi = pb.uu[o];
        double u = 0.0, lh = pb.yu[o], ll = pb.wx[o];
        ipm_point(ip, pb.wu[o], ucn[k], lo, hi, 0.0, u, lh, ll);
        if (PPW == 1 || (wr_en && pvalid)) pb.actc[(size_t)prob * D + k] = (hi - u < lh) ? 1 : ((u - lo < ll) ? -1 : 0);
      }
      if (PPW == 2) { __threadfence_block(); __syncwarp(); } else __syncthreads();
    }
  }

  // Interior-point solve of the current QP down to complementarity tol_c; leaves its active-set estimate in
  // act / actc for the polish pass that follows.  Returns 0, 1 (NaN) or -1 (no convergence).
  template <bool SLEW, bool UBOX, bool DUB>
  __device__ __noinline__ int ipm_solve(const DevProb& pb, unsigned char* smem, double* st, long long* cnt, double scale) {
    double pc[NX], zc[NU];
    Ipm ip;
    ip.mu_prev = scale;
    ip.mu = 0.2 * scale;
    ip.ap = ip.ad = 0.0;
    ip.init = 1;
    const double tol_c = 1e-9 * scale;
    int rc = -1;
    for (int it = 0; it < 80; ++it) {
      bsweep<SLEW, UBOX, DUB, 1>(pb, smem, false, pc, zc, ip);
      if (GROUPED && Nc >= 1) {
        assemble_shared(MODE_POLISH, 0.0, 0.0);
        grad_shared(MODE_POLISH, 0.0, 0.0, pc, zc);
        ipm_shared_pre(pb, ip);
        factor_shared(MODE_POLISH, false);
        apply_shared(MODE_POLISH, 0.0, false);
      }
      double acc[7] = {INFINITY, INFINITY, 0.0, 0.0, 0.0, 0.0, 0.0};
      fsweep<SLEW, UBOX, DUB, 1>(pb, smem, st, ip, acc);
      const double nanf = st[ST_NAN];
      if (!valid) {
        acc[0] = acc[1] = INFINITY;
        acc[2] = acc[3] = acc[4] = acc[5] = acc[6] = 0.0;
      }
      ipm_shared_post(pb, ip, acc);
      cnt[1] += 1; cnt[2] += 1; cnt[3] += 1; cnt[7] += 2;
      if (GROUPED) {
        double* sc = scr();
        sc[(size_t)0 * T + tid] = -acc[0];
        sc[(size_t)1 * T + tid] = -acc[1];
        for (int k = 2; k < 7; ++k) sc[(size_t)k * T + tid] = acc[k];
        sc[(size_t)7 * T + tid] = nanf;
        reduce(8, red, 1000, 0x7cu);
        xreduce(red, 8, false, 0x7cu);
        acc[0] = -red[0]; acc[1] = -red[1];
        for (int k = 2; k < 7; ++k) acc[k] = red[k];
        const double nf = red[7];
        __syncthreads();
        if (nf > 0.0) return 1;
      } else if (nanf > 0.0) {
        return 1;
      }
      const double ap = fmin(1.0, 0.995 * acc[0]), ad = fmin(1.0, 0.995 * acc[1]);
      const double comp = (acc[2] + ad * acc[3] + ap * acc[4] + ap * ad * acc[5]) / fmax(acc[6], 1.0);
      if (!(comp == comp)) return 1;
      ip.mu_prev = ip.mu;
      ip.ap = ap;
      ip.ad = ad;
      ip.init = 0;
      if (comp <= tol_c && ap >= 0.9) {
        rc = 0;
        break;
      }
      ip.mu = fmax(0.2 * comp, 1e-3 * tol_c);
    }
    ipm_finish(pb, ip);
    return rc;
  }

  template <bool SLEW, bool UBOX, bool DUB>
  __device__ __forceinline__ int fused_qp_impl(const DevProb& pb, unsigned char* smem, double* st, long long* cnt) {
    double pc[NX], zc[NU];
    const bool constrained = pb.has_ubox || pb.has_xbox;
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    st[ST_GSCALE] = st[ST_ZSCALE] = 1.0;
    bool ipm_used = false, first = true;
    for (int tries = 0;; ++tries) {
      bsweep<SLEW, UBOX, DUB>(pb, smem, first, pc, zc);
      first = false;
      cnt[2] += 1;
      cnt[7] += 1;
      double kk_c = 0.0, uv_c = 0.0, ch_c = 0.0;
      if (GROUPED && Nc >= 1 && D <= DSMALL && pb.world <= 1) {
        ch_c = shared_small<UBOX>(pc, zc, &kk_c, &uv_c);
      } else if (GROUPED && Nc >= 1) {
        assemble_shared(MODE_POLISH, 0.0, 0.0);
        grad_shared(MODE_POLISH, 0.0, 0.0, pc, zc);
        // the shared block is a D-dimensional box QP once the particles' value functions are fixed:
        // iterate its active set here, without sweeping the horizon again
        for (int t = 0; t <= 2 * D; ++t) {
          factor_shared(MODE_POLISH);
          apply_shared(MODE_POLISH, 0.0);
          if (!UBOX) break;
          kk_c = uv_c = 0.0;
          ch_c = shared_active_update(&kk_c, &uv_c);
          if (ch_c == 0.0) break;
        }
      }
      fsweep<SLEW, UBOX, DUB>(pb, smem, st);
      cnt[3] += 1;
      cnt[7] += 1;
      reduce_stats(st);
      st[ST_KKT] = fmax(st[ST_KKT], kk_c);
      st[ST_UVIOL] = fmax(st[ST_UVIOL], uv_c);
      st[ST_CHANGED] += ch_c;
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
      if (st[ST_XVIOL] > tol_z) return -1;  // a state bound wants to be active: ADMM's job (unfused path)
      if (st[ST_CHANGED] <= 0.0 || tries + 1 >= pb.polish_iters) {
        // the active-set iteration stalls or cycles: one interior-point solve, then polish from its estimate
        if (ipm_used || !pb.has_ubox) return -1;
        ipm_used = true;
        const int rc = ipm_solve<SLEW, UBOX, DUB>(pb, smem, st, cnt, fmax(st[ST_GSCALE], 1.0));
        if (rc == 1) return 1;
        tries = 0;  // (the loop increment makes the next pass a "not first" one: same linearisation point)
      }
    }
  }

  // ---- two problems per warp (PPW == 2) -------------------------------------------------------------------
  // The two half-warps execute every sweep in lockstep (one ring, one instruction stream); what differs per problem is
  // DATA: the statistics are reduced per half-warp, and a problem that is finished — sub-problem accepted, SCP loop
  // converged, failed, or handed over — is frozen: its lanes keep computing but store nothing (`valid`, `wr_en`).
  // A warp's time is the longer of its two problems either way, so lockstep costs nothing over divergence.
  // Returns this half's status: 0 accepted, 1 NaN, 8 hand the problem to the one-problem-per-warp kernel (a state
  // bound wants to be active, or the active-set iteration stalls: interior point / ADMM live there).
  bool wr_en = true;  // this half's shared-block results may be written (shared_small, interior-point state)

  // Shared block of an interior-point iteration in registers (the half-warp counterpart of assemble_shared + grad_shared
  // + ipm_shared_pre + factor_shared + apply_shared for D <= DSMALL): sum_i H_i, sum_i g_i by one half-warp reduction, the
  // lazily updated iterate and the barrier terms of the shared controls, an unmasked Cholesky solve, replicated per lane.
  template <bool UBOX>
  __device__ __noinline__ void shared_small_ipm(const Ipm ip, const double* pc, const double* zc) {
    const FactorSet& F = pb.fs[MODE_POLISH];
    const int Dn = D;
    double hv[DSMALL * DSMALL + DSMALL];
    for (int k = 0; k < Dn; ++k) shared_grad(MODE_POLISH, true, k, nullptr, 0.0, 0.0, nullptr, nullptr, 0, &hv[k * Dn]);
    shared_grad(MODE_POLISH, false, -1, nullptr, 0.0, 0.0, pc, zc, 0, &hv[Dn * Dn]);
    cta_reduce(hv, Dn * Dn + Dn, 0xffffffffu);
    double H[DSMALL * DSMALL], gcv[DSMALL], un[DSMALL], lhn[DSMALL], lln[DSMALL];
    const size_t g0 = (size_t)prob * pb.Mpad;
    const double* Up0 = flip ? pb.U : pb.Up;
    for (int a = 0; a < Dn; ++a) {
      for (int b = 0; b < Dn; ++b) H[a * DSMALL + b] = 0.5 * (hv[b * Dn + a] + hv[a * Dn + b]);
      gcv[a] = hv[Dn * Dn + a];
      const size_t o = ixg(a / NU, NU, a % NU, g0);
      const double lo = pb.lu[o], hi = pb.uu[o];
      double u = 0.0, lh = 0.0, ll = 0.0, u0 = 0.0;
      if (!ip.init) {
        u0 = pb.wu[o]; lh = pb.yu[o]; ll = pb.wx[o];
      }
      ipm_point(ip, u0, pb.uc[(size_t)prob * Dn + a], lo, hi, Up0[o], u, lh, ll);
      un[a] = u; lhn[a] = lh; lln[a] = ll;
      const double sh = hi - u, sl = u - lo, sg = lh / sh + ll / sl;
      H[a * DSMALL + a] += sg;
      gcv[a] += -sg * u + ip.mu / sh - ip.mu / sl;
    }
    double L[DSMALL * DSMALL], y[DSMALL];
    for (int k = 0; k < Dn; ++k) {  // Cholesky (lower)
      double dk = H[k * DSMALL + k];
      for (int l = 0; l < k; ++l) dk -= L[k * DSMALL + l] * L[k * DSMALL + l];
      dk = sqrt(dk);
      L[k * DSMALL + k] = dk;
      for (int i = k + 1; i < Dn; ++i) {
        double v = H[i * DSMALL + k];
        for (int l = 0; l < k; ++l) v -= L[i * DSMALL + l] * L[k * DSMALL + l];
        L[i * DSMALL + k] = v / dk;
      }
    }
    for (int i = 0; i < Dn; ++i) {
      double v = -gcv[i];
      for (int l = 0; l < i; ++l) v -= L[i * DSMALL + l] * y[l];
      y[i] = v / L[i * DSMALL + i];
    }
    for (int i = Dn - 1; i >= 0; --i) {
      double v = y[i];
      for (int l = i + 1; l < Dn; ++l) v -= L[l * DSMALL + i] * y[l];
      y[i] = v / L[i * DSMALL + i];
    }
    __syncwarp();  // every lane has read the interior-point state of the shared controls
    if (hl == 0 && pvalid && wr_en) {
      for (int a = 0; a < Dn; ++a) {
        const size_t o = ixg(a / NU, NU, a % NU, g0);
        pb.wu[o] = un[a]; pb.yu[o] = lhn[a]; pb.wx[o] = lln[a];
        pb.uc[(size_t)prob * Dn + a] = y[a];
        pb.gc[(size_t)prob * Dn + a] = gcv[a];
        for (int b = 0; b < Dn; ++b) F.Hc[(size_t)prob * Dn * Dn + (size_t)a * Dn + b] = H[a * DSMALL + b];
      }
    }
    if (PPW == 2) { __threadfence_block(); __syncwarp(); } else __syncthreads();
  }

  // Interior-point solve (ipm_solve) for the halves with want == true, the other half frozen.  rc per half: 0 converged,
  // 1 NaN, -1 no convergence (the polish passes that follow decide).
  template <bool SLEW, bool UBOX, bool DUB>
  __device__ __noinline__ int ipm_solve_half(const DevProb& pb, unsigned char* smem, double* st, long long* cnt, double scale,
                                              bool want, bool base_valid) {
    double pc[NX], zc[NU];
    Ipm ip;
    ip.mu_prev = scale;
    ip.mu = 0.2 * scale;
    ip.ap = ip.ad = 0.0;
    ip.init = 1;
    const double tol_c = 1e-9 * scale;
    int rc = -1;
    bool fin = !want;
    for (int it = 0; it < 80; ++it) {
      this->valid = base_valid && want && !fin;
      wr_en = want && !fin;
      half_phase(1);
      bsweep<SLEW, UBOX, DUB, 1>(pb, smem, false, pc, zc, ip);
      if (Nc >= 1) shared_small_ipm<UBOX>(ip, pc, zc);
      double acc[7] = {INFINITY, INFINITY, 0.0, 0.0, 0.0, 0.0, 0.0};
      __syncthreads();  // (aligning the forward sweeps as well: 122 vs 129 ms on C3)
      fsweep<SLEW, UBOX, DUB, 1>(pb, smem, st, ip, acc);
      double v8[8];
      if (!this->valid) {
        acc[0] = acc[1] = INFINITY;
        acc[2] = acc[3] = acc[4] = acc[5] = acc[6] = 0.0;
      }
      if (want && !fin) ipm_shared_post(pb, ip, acc);
      v8[0] = -acc[0]; v8[1] = -acc[1];
      for (int k = 2; k < 7; ++k) v8[k] = acc[k];
      v8[7] = this->valid ? st[ST_NAN] : 0.0;
      cta_reduce(v8, 8, 0x7cu);
      if (!fin) {
        cnt[1] += 1; cnt[2] += 1; cnt[3] += 1; cnt[7] += 2;
        const double a0 = -v8[0], a1 = -v8[1];
        const double ap = fmin(1.0, 0.995 * a0), ad = fmin(1.0, 0.995 * a1);
        const double comp = (v8[2] + ad * v8[3] + ap * v8[4] + ap * ad * v8[5]) / fmax(v8[6], 1.0);
        if (v8[7] > 0.0 || !(comp == comp)) {
          rc = 1;
          fin = true;
        } else {
          ip.mu_prev = ip.mu;
          ip.ap = ap;
          ip.ad = ad;
          ip.init = 0;
          if (comp <= tol_c && ap >= 0.9) {
            rc = 0;
            fin = true;
          } else {
            ip.mu = fmax(0.2 * comp, 1e-3 * tol_c);
          }
        }
      }
      if (__all_sync(0xffffffffu, fin)) break;
    }
    this->valid = base_valid && want && rc != 1;
    wr_en = want && rc != 1;
    ipm_finish(pb, ip);
    return rc;
  }

  template <bool SLEW, bool UBOX, bool DUB>
  __device__ __forceinline__ int fused_qp_half(const DevProb& pb, unsigned char* smem, double* st, long long* cnt, bool active) {
    double pc[NX], zc[NU];
    double k_dx = 0.0, k_du = 0.0, k_obj = 0.0, k_kkt = 0.0, k_gs = 1.0, k_nan = 0.0;  // what the SCP loop reads of the accepted pass
    const bool constrained = pb.has_ubox || pb.has_xbox;
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    const bool base_valid = this->valid;
    enum { RUN = 0, WANT_IPM = 1, DONE = 2 };
    int state = active ? RUN : DONE, status = 0, tries = 0;
    bool first = true, ipm_used = false;
    double scale = 1.0;
    for (;;) {
      this->valid = base_valid && state == RUN;
      wr_en = state == RUN;
      half_phase(1);
      bsweep<SLEW, UBOX, DUB>(pb, smem, first, pc, zc);
      first = false;
      double kk_c = 0.0, uv_c = 0.0, ch_c = 0.0;
      if (Nc >= 1) ch_c = shared_small<UBOX>(pc, zc, &kk_c, &uv_c);
      __syncthreads();
      fsweep<SLEW, UBOX, DUB>(pb, smem, st);
      reduce_stats(st);
      if (state == RUN) {
        cnt[2] += 1; cnt[3] += 1; cnt[7] += 2;
        st[ST_KKT] = fmax(st[ST_KKT], kk_c);
        st[ST_UVIOL] = fmax(st[ST_UVIOL], uv_c);
        st[ST_CHANGED] += ch_c;
        const double tol_g = pb.kkt_tol * st[ST_GSCALE], tol_z = pb.kkt_tol * st[ST_ZSCALE];
        if (st[ST_NAN] > 0.0) {
          status = 1;
          state = DONE;
        } else if (!constrained || (st[ST_KKT] <= tol_g && st[ST_UVIOL] <= tol_z && st[ST_XVIOL] <= tol_z)) {
          cnt[5] += 1;
          state = DONE;
        } else if (st[ST_XVIOL] > tol_z) {
          status = 8;  // a state bound wants to be active: ADMM lives in the one-problem-per-warp kernel
          state = DONE;
        } else if (st[ST_CHANGED] <= 0.0 || tries + 1 >= pb.polish_iters) {
          // the active-set iteration stalls or cycles: one interior-point solve, then polish from its estimate
          if (ipm_used || !pb.has_ubox || !pb.half_ipm) {
            status = 8;
            state = DONE;
          } else {
            state = WANT_IPM;
            scale = fmax(st[ST_GSCALE], 1.0);
          }
        }
        ++tries;
        if (state == DONE)
          k_dx = st[ST_DX], k_du = st[ST_DU], k_obj = st[ST_OBJ], k_kkt = st[ST_KKT], k_gs = st[ST_GSCALE], k_nan = st[ST_NAN];
      }
      if (__any_sync(0xffffffffu, state == WANT_IPM)) {
        const int rc = ipm_solve_half<SLEW, UBOX, DUB>(pb, smem, st, cnt, scale, state == WANT_IPM, base_valid);
        if (state == WANT_IPM) {
          if (rc == 1) {
            status = 1;
            state = DONE;
            k_dx = k_du = k_obj = k_kkt = 0.0;
            k_gs = 1.0;
            k_nan = 1.0;
          } else {
            state = RUN;
            ipm_used = true;
            tries = 0;
          }
        }
      }
      if (__all_sync(0xffffffffu, state == DONE)) break;
    }
    this->valid = base_valid;
    wr_en = true;
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    st[ST_DX] = k_dx, st[ST_DU] = k_du, st[ST_OBJ] = k_obj, st[ST_KKT] = k_kkt, st[ST_GSCALE] = k_gs, st[ST_NAN] = k_nan;
    st[ST_ZSCALE] = 1.0;
    return status;
  }

  // One convex sub-problem on the fused path.  Returns status bits (as solve_qp), or -1 when the
  // active-set iteration has to be globalised by ADMM (caller falls back to the unfused path).
  // VAR (kernel template parameter, chosen by the host from the problem's flags): bit 0 = built-in
  // Dubins linearisation on the fly, bit 1 = lean (no slew penalty: W, Z, z and E vanish at compile time).
  template <int VAR>
  __device__ __forceinline__ int fused_qp(const DevProb& pb, unsigned char* smem, double* st, long long* cnt) {
    constexpr bool DUB = (VAR & 1) && NX == 4 && NU == 2;
    constexpr bool LEAN = (VAR & 2) != 0;
    return fused_qp_impl<!LEAN, true, DUB>(pb, smem, st, cnt);
  }

