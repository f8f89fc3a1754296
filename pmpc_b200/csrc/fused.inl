  // =============================================================================================
  // FUSED PATH (textually included inside struct Solver).
  //
  // One backward sweep does, per stage: (optionally) the adjoint/KKT check of the current iterate
  // with the primal-dual active-set update, the linearisation (recomputed on the fly for built-in
  // dynamics instead of streaming f, fx, fu through HBM), the masked Riccati matrix step and the
  // vector step.  One forward sweep does the rollout and the SCP residual/objective bookkeeping
  // (scp_mpc.py:397-405).  A polish pass is 2 sweeps over the horizon (the unfused path needs 4,
  // plus 2 per SCP iteration for linearise and X_prev <- X); the only per-(particle, stage) state
  // that lives in HBM between sweeps is K, k, X, U, X_prev, U_prev and the active set.
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
  static constexpr int DEPTH = 3;
  // ring slot layout (doubles): backward sweep [Xp | Up | X | U | act], forward sweep [Xp | Up | K | k]
  static constexpr int R_XP = 0, R_UP = R_XP + NX * 32, R_X = R_UP + NU * 32, R_U = R_X + NX * 32,
                       R_ACT = R_U + NU * 32, R_K = R_UP + NU * 32, R_KF = R_K + NU * NX * 32;
  static constexpr int SLOT_B = R_ACT + NU * 4, SLOT_F = R_KF + NU * 32;
  static constexpr int SLOT = SLOT_B > SLOT_F ? SLOT_B : SLOT_F;
  // uniform cost table row: [Qs | Rs | qref | rref | xref | uref | lu | uu]
  static constexpr int T_QS = 0, T_RS = T_QS + NQ, T_QR = T_RS + NR, T_RR = T_QR + NX, T_XR = T_RR + NU,
                       T_UR = T_XR + NX, T_LU = T_UR + NU, T_UU = T_LU + NU, TW = T_UU + NU;

  __host__ __device__ static constexpr int table_doubles(int N_) { return N_ * TW; }
  __host__ __device__ static constexpr int ring_doubles(int nwarps) { return nwarps * DEPTH * SLOT; }

  __device__ __forceinline__ double symQ(const double* Qp, int r, int t) const {
    return Qp[r <= t ? pk(r, t, NX) : pk(t, r, NX)];
  }
  __device__ __forceinline__ double symR(const double* Rp, int a, int b) const {
    return Rp[a <= b ? pk(a, b, NU) : pk(b, a, NU)];
  }
  // element 0 of the warp's stage-j block of an E-element array (no lane offset): bulk-copy source
  __device__ __forceinline__ const double* blk(const double* a, int j, int E) const {
    return a + (((size_t)(g >> 5) * (size_t)N + (size_t)j) * (size_t)E) * 32;
  }

  // dynamic shared memory: [table][ring of every warp][mbarriers of every warp]
  __device__ __forceinline__ void fused_init(unsigned char* smem) {
    lane = tid & 31;
    wid = tid >> 5;
    col = g & 31;
    rcount = 0;
    ring = nullptr;
    bars = nullptr;
    tab = nullptr;
    double* base = reinterpret_cast<double*>(smem);
    const int nw = T >> 5;
    double* tb = base;
    if (pb.ring) {
      ring = base + pb.tab_doubles + (size_t)wid * DEPTH * SLOT;
      bars = reinterpret_cast<uint64_t*>(base + pb.tab_doubles + (size_t)nw * DEPTH * SLOT) + wid * DEPTH;
      if (lane == 0) {
        for (int k = 0; k < DEPTH; ++k) mbar_init(bars + k, 1);
        fence_mbar_init();
      }
    }
    if (GROUPED && pb.tab_doubles > 0 && pb.uniform[prob]) {
      // particle 0 of the problem speaks for all (checked on the device by uniform_check_kernel)
      const size_t g0 = (size_t)prob * pb.Mpad;
      for (int idx = tid; idx < N * TW; idx += T) {
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
    }
    __syncthreads();
  }

  // per-stage cost / reference / bound data into registers, from the table or from HBM
  __device__ __forceinline__ void stage_cost(int j, double* Qp, double* Rp, double* qr, double* rr, double* lo,
                                             double* hi) const {
    if (tab != nullptr) {
      const double* t = tab + j * TW;
#pragma unroll UNR
      for (int e = 0; e < NQ; ++e) Qp[e] = t[T_QS + e];
#pragma unroll UNR
      for (int e = 0; e < NR; ++e) Rp[e] = t[T_RS + e];
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) qr[e] = t[T_QR + e];
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) {
        rr[e] = t[T_RR + e];
        lo[e] = t[T_LU + e];
        hi[e] = t[T_UU + e];
      }
    } else {
#pragma unroll UNR
      for (int e = 0; e < NQ; ++e) Qp[e] = ld(pb.Qs, j, NQ, e);
#pragma unroll UNR
      for (int e = 0; e < NR; ++e) Rp[e] = ld(pb.Rs, j, NR, e);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) qr[e] = ld(pb.qref, j, NX, e);
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) {
        rr[e] = ld(pb.rref, j, NU, e);
        lo[e] = ld(pb.lu, j, NU, e);
        hi[e] = ld(pb.uu, j, NU, e);
      }
    }
    if (!pb.has_ubox) {
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) {
        lo[e] = -INFINITY;
        hi[e] = INFINITY;
      }
    }
  }

  // x_j = A x_{j-1} + B u_j + c about (xm1 = X_[j], ub = U_prev[j]); A_0 = 0 (lqp_utils.jl:288-296)
  __device__ __forceinline__ void lin_stage(int j, const double* xm1, const double* ub, const double* dp, double* A,
                                            double* Bm, double* c) const {
    if constexpr (NX == 4 && NU == 2) {
      if (pb.dyn_kind == DYN_DUBINS) {
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
        return;
      }
    }
#pragma unroll UNR
    for (int e = 0; e < NX * NX; ++e) A[e] = ld(pb.A, j, NX * NX, e);
#pragma unroll UNR
    for (int e = 0; e < NX * NU; ++e) Bm[e] = ld(pb.Bm, j, NX * NU, e);
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) c[e] = ld(pb.c, j, NX, e);
  }

  // lane 0: stream the backward-sweep inputs of stage jj into ring slot `sl`
  __device__ __forceinline__ void issue_b(int jj, int sl, bool kkt) {
    double* dst = ring + sl * SLOT;
    uint64_t* bar = bars + sl;
    uint32_t bytes = NU * 256;
    if (jj >= 1) bytes += NX * 256;
    if (kkt) bytes += (NX + NU) * 256;
    if (pb.has_ubox) bytes += NU * 32;
    mbar_expect_tx(bar, bytes);
    if (jj >= 1) bulk_g2s(dst + R_XP, blk(Xpv, jj - 1, NX), NX * 256, bar);
    bulk_g2s(dst + R_UP, blk(Upv, jj, NU), NU * 256, bar);
    if (kkt) {
      bulk_g2s(dst + R_X, blk(Xc, jj, NX), NX * 256, bar);
      bulk_g2s(dst + R_U, blk(Uc, jj, NU), NU * 256, bar);
    }
    if (pb.has_ubox)
      bulk_g2s(dst + R_ACT, pb.act + (((size_t)(g >> 5) * (size_t)N + (size_t)jj) * NU) * 32, NU * 32, bar);
  }
  __device__ __forceinline__ void issue_f(int jj, int sl) {
    double* dst = ring + sl * SLOT;
    uint64_t* bar = bars + sl;
    const FactorSet& F = pb.fs[MODE_POLISH];
    uint32_t bytes = (NX + NU) * 256;
    if (jj >= Nc) bytes += (NU * NX + NU) * 256;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst + R_XP, blk(Xpv, jj, NX), NX * 256, bar);
    bulk_g2s(dst + R_UP, blk(Upv, jj, NU), NU * 256, bar);
    if (jj >= Nc) {
      bulk_g2s(dst + R_K, blk(F.K, jj, NU * NX), NU * NX * 256, bar);
      bulk_g2s(dst + R_KF, blk(pb.kff, jj, NU), NU * 256, bar);
    }
  }

  // Backward sweep.  kkt: check the iterate (Xc, Uc) and move to the proposed active set on the fly.
  // first: the linearisation point changed since the last sweep (store the shared-stage data).
  __device__ __forceinline__ void bsweep(bool kkt, bool first, double* st, double* pc, double* zc) {
    const FactorSet& F = pb.fs[MODE_POLISH];
    const bool on_the_fly = (NX == 4 && NU == 2 && pb.dyn_kind == DYN_DUBINS);
    const bool use_ring = RING_SHAPE && ring != nullptr;
    double dp[3] = {0.0, 0.0, 0.0};
    if (on_the_fly) {
      dp[0] = ldp(pb.dynp, 0); dp[1] = ldp(pb.dynp, 1); dp[2] = ldp(pb.dynp, 2);
    }
    double Pm[NX * NX], W[NX * NU], Z[NU * NU], p[NX], z[NU], Ata[NX], un[NU], xq[NX];
#pragma unroll UNR
    for (int e = 0; e < NX * NX; ++e) Pm[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NX * NU; ++e) W[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NU * NU; ++e) Z[e] = 0.0;
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) { p[e] = 0.0; Ata[e] = 0.0; }
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) { z[e] = 0.0; un[e] = 0.0; }
    if (use_ring) {
      __syncwarp();
      if (lane == 0) {
        for (int d = 0; d < DEPTH; ++d)
          if (N - 1 - d >= 0) issue_b(N - 1 - d, (rcount + d) % DEPTH, kkt);
      }
    }
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) xq[e] = ld(Xpv, N - 1, NX, e);
    double kkt_v = 0.0, uviol = 0.0, xviol = 0.0, changed = 0.0, gscale = 1.0, zscale = 1.0, nanf = 0.0;
    double* sc = scr();
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
    for (int j = N - 1; j >= 0; --j) {
      // ---- stage inputs
      double xm1[NX], ub[NU], xs[NX], us[NU], ups[NU];
      signed char act[NU];
      if (use_ring) {
        const int sl = rcount % DEPTH;
        mbar_wait(bars + sl, (rcount / DEPTH) & 1);
        const double* b = ring + sl * SLOT + col;
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xm1[e] = (j >= 1) ? b[R_XP + e * 32] : ldp(pb.x0, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = b[R_UP + e * 32];
        if (kkt) {
#pragma unroll UNR
          for (int e = 0; e < NX; ++e) xs[e] = b[R_X + e * 32];
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) us[e] = b[R_U + e * 32];
        }
        const signed char* ab = reinterpret_cast<const signed char*>(ring + sl * SLOT + R_ACT) + col;
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) act[t] = (pb.has_ubox && j >= Nc) ? ab[t * 32] : 0;
        __syncwarp();
        if (lane == 0 && j - DEPTH >= 0) issue_b(j - DEPTH, sl, kkt);
        ++rcount;
      } else {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xm1[e] = (j >= 1) ? ld(Xpv, j - 1, NX, e) : ldp(pb.x0, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = ld(Upv, j, NU, e);
        if (kkt) {
#pragma unroll UNR
          for (int e = 0; e < NX; ++e) xs[e] = ld(Xc, j, NX, e);
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) us[e] = ld(Uc, j, NU, e);
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) act[t] = (pb.has_ubox && j >= Nc) ? pb.act[ix(j, NU, t)] : 0;
      }
      if (kkt) {
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ups[e] = (pb.has_slew && j >= 1) ? ld(Uc, j - 1, NU, e) : 0.0;
      }
      double Qp[NQ], Rp[NR], q[NX], r[NU], lo[NU], hi[NU];
      stage_cost(j, Qp, Rp, q, r, lo, hi);
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) q[e] = -q[e] - pb.reg_x * xq[e];
#pragma unroll UNR
      for (int e = 0; e < NU; ++e) r[e] = -r[e] - pb.reg_u * ub[e];
      double A[NX * NX], Bm[NX * NU], cv[NX];
      lin_stage(j, xm1, ub, dp, A, Bm, cv);
      // ---- adjoint / KKT of the current iterate (kkt_sweep, fused)
      if (kkt) {
        double a[NX];
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr) {
          double acc = q[rr] + Ata[rr];
          gscale = fmax(gscale, fabs(q[rr]));
          zscale = fmax(zscale, fabs(xs[rr]));
          if (!(xs[rr] == xs[rr])) nanf = 1.0;
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) acc += symQ(Qp, rr, t) * xs[t];
          a[rr] = acc;
        }
        if (pb.has_xbox) {
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr)
            xviol = fmax(xviol, fmax(xs[rr] - ld(pb.ux, j, NX, rr), ld(pb.lx, j, NX, rr) - xs[rr]));
        }
#pragma unroll UNR
        for (int b = 0; b < NU; ++b) {
          double acc = r[b];
          gscale = fmax(gscale, fabs(r[b]));
          zscale = fmax(zscale, fabs(us[b]));
          if (!(us[b] == us[b])) nanf = 1.0;
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += symR(Rp, b, t) * us[t];
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr) acc += Bm[rr * NU + b] * a[rr];
          if (j >= 1) acc += s * (us[b] - ups[b]);
          if (j < N - 1) acc += s * (us[b] - un[b]);
          if (j < Nc) {
            sc[(size_t)(j * NU + b) * T + tid] = valid ? acc : 0.0;
          } else {
            const double mu = -acc, ubv = us[b];
            uviol = fmax(uviol, fmax(ubv - hi[b], lo[b] - ubv));
            const bool at_hi = ubv >= hi[b] - 1e-12 * fmax(1.0, fabs(hi[b]));
            const bool at_lo = ubv <= lo[b] + 1e-12 * fmax(1.0, fabs(lo[b]));
            double v;
            if (at_hi && at_lo) v = 0.0;
            else if (at_hi) v = fmax(acc, 0.0);
            else if (at_lo) v = fmax(-acc, 0.0);
            else v = fabs(acc);
            kkt_v = fmax(kkt_v, v);
            if (pb.has_ubox) {
              signed char na = 0;
              if (mu + (ubv - hi[b]) > 0.0 && hi[b] < INFINITY) na = 1;
              else if (mu + (ubv - lo[b]) < 0.0 && lo[b] > -INFINITY) na = -1;
              if (na != act[b]) {
                changed += 1.0;
                act[b] = na;
                if (valid) pb.act[ix(j, NU, b)] = na;
              }
            }
          }
        }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) un[t] = us[t];
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) {
          double acc = 0.0;
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr) acc += A[rr * NX + t] * a[rr];
          Ata[t] = acc;
        }
      }
      if (j >= Nc) {
        // ---- Riccati matrix step with the stage's active set, then the vector step
        double S[NX * NX], SBW[NX * NU], H[NU * NU], E[NU * NU], ufix[NU];
        bool fixed[NU];
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
          for (int t = rr; t < NX; ++t) {
            const double v = Qp[pk(rr, t, NX)] + Pm[rr * NX + t];
            S[rr * NX + t] = v;
            S[t * NX + rr] = v;
          }
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) {
            double acc = W[rr * NU + t];
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) acc += S[rr * NX + k] * Bm[k * NU + t];
            SBW[rr * NU + t] = acc;
          }
#pragma unroll UNR
        for (int a = 0; a < NU; ++a)
#pragma unroll UNR
          for (int b = a; b < NU; ++b) {
            double acc = Rp[pk(a, b, NU)] + Z[a * NU + b];
            if (a == b) acc += (j >= 1 ? s : 0.0);
#pragma unroll UNR
            for (int k = 0; k < NX; ++k) acc += Bm[k * NU + a] * SBW[k * NU + b] + W[k * NU + a] * Bm[k * NU + b];
            H[a * NU + b] = acc;
            H[b * NU + a] = acc;
          }
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) {
          fixed[t] = act[t] != 0;
          ufix[t] = (act[t] > 0) ? hi[t] : ((act[t] < 0) ? lo[t] : 0.0);
        }
        masked_spd_inverse<NU>(H, fixed, E);
        double tk[NX], eta[NX], h[NU], kv[NU];
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr) {
          double acc = 0.0;
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += SBW[rr * NU + t] * ufix[t];
          tk[rr] = acc;
        }
        // eta = S c + SBW ufix + q + p
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr) {
          double acc = tk[rr] + q[rr] + p[rr];
#pragma unroll UNR
          for (int k = 0; k < NX; ++k) acc += S[rr * NX + k] * cv[k];
          eta[rr] = acc;
        }
        // h = r + [W'c - B' tk + H ufix] + z + B' eta
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = r[a] + z[a];
#pragma unroll UNR
          for (int k = 0; k < NX; ++k) acc += W[k * NU + a] * cv[k] + Bm[k * NU + a] * (eta[k] - tk[k]);
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) acc += H[a * NU + t] * ufix[t];
          h[a] = acc;
        }
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = ufix[a];
#pragma unroll UNR
          for (int b = 0; b < NU; ++b) acc -= E[a * NU + b] * h[b];
          kv[a] = acc;
        }
        double K[NU * NX];
        if (j >= 1) {
          double G[NU * NX], SA[NX * NX];
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
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) {
              double acc = 0.0;
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) acc += S[rr * NX + k] * A[k * NX + t];
              SA[rr * NX + t] = acc;
            }
#pragma unroll UNR
          for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
            for (int t = rr; t < NX; ++t) {
              double acc = 0.0;
#pragma unroll UNR
              for (int k = 0; k < NX; ++k) acc += A[k * NX + rr] * SA[k * NX + t];
#pragma unroll UNR
              for (int a = 0; a < NU; ++a) acc += G[a * NX + rr] * K[a * NX + t];
              Pm[rr * NX + t] = acc;
              Pm[t * NX + rr] = acc;
            }
          // p <- A' eta + K' h ; z <- -s k
#pragma unroll UNR
          for (int t = 0; t < NX; ++t) {
            double acc = 0.0;
#pragma unroll UNR
            for (int rr = 0; rr < NX; ++rr) acc += A[rr * NX + t] * eta[rr];
#pragma unroll UNR
            for (int a = 0; a < NU; ++a) acc += K[a * NX + t] * h[a];
            p[t] = acc;
          }
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) z[a] = -s * kv[a];
        } else {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) K[e] = 0.0;
        }
#pragma unroll UNR
        for (int rr = 0; rr < NX; ++rr)
#pragma unroll UNR
          for (int t = 0; t < NU; ++t) W[rr * NU + t] = -s * K[t * NX + rr];
#pragma unroll UNR
        for (int a = 0; a < NU; ++a)
#pragma unroll UNR
          for (int b = 0; b < NU; ++b) Z[a * NU + b] = (a == b ? s : 0.0) - s * s * E[a * NU + b];
        if (valid) {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) at(F.K, j, NU * NX, e) = K[e];
#pragma unroll UNR
          for (int a = 0; a < NU; ++a) at(pb.kff, j, NU, a) = kv[a];
          if (pb.has_slew) {
#pragma unroll UNR
            for (int a = 0; a < NU; ++a)
#pragma unroll UNR
              for (int b = a; b < NU; ++b) at(F.E, j, NR, pk(a, b, NU)) = E[a * NU + b];
          }
        }
      } else if (first && valid) {
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
      if (j == Nc && Nc >= 1) {
        if (valid) {
#pragma unroll UNR
          for (int e = 0; e < NX * NX; ++e) F.Pc[(size_t)e * P + g] = Pm[e];
#pragma unroll UNR
          for (int e = 0; e < NX * NU; ++e) F.Wc[(size_t)e * P + g] = W[e];
#pragma unroll UNR
          for (int e = 0; e < NU * NU; ++e) F.Zc[(size_t)e * P + g] = Z[e];
        }
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) pc[e] = p[e];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) zc[e] = z[e];
      }
#pragma unroll UNR
      for (int e = 0; e < NX; ++e) xq[e] = xm1[e];
    }
    if (use_ring) fence_proxy_async();  // K, k, act written above are bulk-copied by the next sweep
    if (kkt) {
      st[ST_KKT] = kkt_v; st[ST_UVIOL] = uviol; st[ST_XVIOL] = xviol; st[ST_CHANGED] = changed;
      st[ST_GSCALE] = gscale; st[ST_ZSCALE] = zscale; st[ST_NAN] = fmax(st[ST_NAN], nanf);
      if (GROUPED && Nc >= 1) {
        // shared block: summed reduced gradient -> multipliers / proposed active set (kkt_sweep tail)
        double* muc = pb.muc + (size_t)prob * D;
        reduce(D, muc);
        const size_t g0 = (size_t)prob * pb.Mpad;
        const double* uc = pb.uc + (size_t)prob * D;
        signed char* actc = pb.actc + (size_t)prob * D;
        double ch = 0.0, kk = 0.0, uv = 0.0;
        for (int k = 0; k < D; ++k) {  // every thread evaluates the (tiny) shared block identically
          const int jj = k / NU, b = k % NU;
          const double acc = muc[k], mu = -acc, ubv = uc[k];
          double lo = -INFINITY, hi = INFINITY;
          if (pb.has_ubox) {
            lo = pb.lu[ixg(jj, NU, b, g0)];
            hi = pb.uu[ixg(jj, NU, b, g0)];
          }
          uv = fmax(uv, fmax(ubv - hi, lo - ubv));
          const bool at_hi = ubv >= hi - 1e-12 * fmax(1.0, fabs(hi));
          const bool at_lo = ubv <= lo + 1e-12 * fmax(1.0, fabs(lo));
          double v;
          if (at_hi && at_lo) v = 0.0;
          else if (at_hi) v = fmax(acc, 0.0);
          else if (at_lo) v = fmax(-acc, 0.0);
          else v = fabs(acc);
          kk = fmax(kk, v);
          signed char na = 0;
          if (pb.has_ubox) {
            if (mu + (ubv - hi) > 0.0 && hi < INFINITY) na = 1;
            else if (mu + (ubv - lo) < 0.0 && lo > -INFINITY) na = -1;
          }
          if (na != actc[k]) ch += 1.0;
          pb.actcn[(size_t)prob * D + k] = na;  // adopted below, after every thread has read actc
        }
        st[ST_KKT] = fmax(st[ST_KKT], kk);
        st[ST_UVIOL] = fmax(st[ST_UVIOL], uv);
        st[ST_CHANGED] += ch;
        __syncthreads();
        for (int k = tid; k < D; k += T) actc[k] = pb.actcn[(size_t)prob * D + k];
        __syncthreads();
      }
    }
  }

  // Forward sweep: rollout of the feedback policy + SCP bookkeeping against (Xpv, Upv).
  __device__ __forceinline__ void fsweep(double* st) {
    const FactorSet& F = pb.fs[MODE_POLISH];
    const bool on_the_fly = (NX == 4 && NU == 2 && pb.dyn_kind == DYN_DUBINS);
    const bool use_ring = RING_SHAPE && ring != nullptr;
    double dp[3] = {0.0, 0.0, 0.0};
    if (on_the_fly) {
      dp[0] = ldp(pb.dynp, 0); dp[1] = ldp(pb.dynp, 1); dp[2] = ldp(pb.dynp, 2);
    }
    if (use_ring) {
      __syncwarp();
      if (lane == 0) {
        for (int d = 0; d < DEPTH; ++d)
          if (d < N) issue_f(d, (rcount + d) % DEPTH);
      }
    }
    const double* uc = (GROUPED && Nc >= 1) ? pb.uc + (size_t)prob * D : nullptr;
    double x[NX], up[NU], xm1[NX];
#pragma unroll UNR
    for (int e = 0; e < NX; ++e) { x[e] = 0.0; xm1[e] = ldp(pb.x0, e); }
#pragma unroll UNR
    for (int e = 0; e < NU; ++e) up[e] = 0.0;
    double dxm = 0.0, dum = 0.0, obj = 0.0, nanf = 0.0;
    const double rdiag = pb.rdiag0[g];
    for (int j = 0; j < N; ++j) {
      double ub[NU], xpj[NX], Kj[NU * NX], kf[NU], u[NU], xn[NX];
      if (use_ring) {
        const int sl = rcount % DEPTH;
        mbar_wait(bars + sl, (rcount / DEPTH) & 1);
        const double* b = ring + sl * SLOT + col;
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xpj[e] = b[R_XP + e * 32];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = b[R_UP + e * 32];
        if (j >= Nc) {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) Kj[e] = b[R_K + e * 32];
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) kf[e] = b[R_KF + e * 32];
        }
        __syncwarp();
        if (lane == 0 && j + DEPTH < N) issue_f(j + DEPTH, sl);
        ++rcount;
      } else {
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xpj[e] = ld(Xpv, j, NX, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ub[e] = ld(Upv, j, NU, e);
        if (j >= Nc) {
#pragma unroll UNR
          for (int e = 0; e < NU * NX; ++e) Kj[e] = ld(F.K, j, NU * NX, e);
#pragma unroll UNR
          for (int e = 0; e < NU; ++e) kf[e] = ld(pb.kff, j, NU, e);
        }
      }
      if (j < Nc) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) u[a] = uc[j * NU + a];
      } else {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) {
          double acc = kf[a];
          if (j >= 1) {
#pragma unroll UNR
            for (int t = 0; t < NX; ++t) acc += Kj[a * NX + t] * x[t];
            if (pb.has_slew) {
#pragma unroll UNR
              for (int b = 0; b < NU; ++b) acc += s * ld(F.E, j, NR, a <= b ? pk(a, b, NU) : pk(b, a, NU)) * up[b];
            }
          }
          u[a] = acc;
        }
      }
      double A[NX * NX], Bm[NX * NU], cv[NX];
      lin_stage(j, xm1, ub, dp, A, Bm, cv);
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
      if (valid) {
#pragma unroll UNR
        for (int a = 0; a < NU; ++a) at(Uc, j, NU, a) = u[a];
#pragma unroll UNR
        for (int r = 0; r < NX; ++r) at(Xc, j, NX, r) = xn[r];
      }
      // bookkeeping (scp_bookkeeping, fused)
      double Qp[NQ], Rp[NR], xr[NX], ur[NU];
      if (tab != nullptr) {
        const double* t = tab + j * TW;
#pragma unroll UNR
        for (int e = 0; e < NQ; ++e) Qp[e] = t[T_QS + e];
#pragma unroll UNR
        for (int e = 0; e < NR; ++e) Rp[e] = t[T_RS + e];
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xr[e] = t[T_XR + e];
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ur[e] = t[T_UR + e];
      } else {
#pragma unroll UNR
        for (int e = 0; e < NQ; ++e) Qp[e] = ld(pb.Qs, j, NQ, e);
#pragma unroll UNR
        for (int e = 0; e < NR; ++e) Rp[e] = ld(pb.Rs, j, NR, e);
#pragma unroll UNR
        for (int e = 0; e < NX; ++e) xr[e] = ld(pb.xref, j, NX, e);
#pragma unroll UNR
        for (int e = 0; e < NU; ++e) ur[e] = ld(pb.uref, j, NU, e);
      }
      double dx[NX], du[NU], n2 = 0.0;
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        const double d = xn[r] - xpj[r];
        n2 += d * d;
        dx[r] = xn[r] - xr[r];
        if (!(xn[r] == xn[r])) nanf = 1.0;
      }
      dxm = fmax(dxm, sqrt(n2));
      n2 = 0.0;
#pragma unroll UNR
      for (int t = 0; t < NU; ++t) {
        const double d = u[t] - ub[t];
        n2 += d * d;
        du[t] = u[t] - ur[t];
        if (!(u[t] == u[t])) nanf = 1.0;
      }
      dum = fmax(dum, sqrt(n2));
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NX; ++t) acc += symQ(Qp, r, t) * dx[t];
        obj += dx[r] * (acc - pb.reg_x * dx[r]);
      }
#pragma unroll UNR
      for (int b = 0; b < NU; ++b) {
        double acc = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NU; ++t) acc += symR(Rp, b, t) * du[t];
        obj += du[b] * (acc - (pb.reg_u + (j == 0 ? rdiag : 0.0)) * du[b]);
      }
#pragma unroll UNR
      for (int r = 0; r < NX; ++r) { x[r] = xn[r]; xm1[r] = xpj[r]; }
#pragma unroll UNR
      for (int a = 0; a < NU; ++a) up[a] = u[a];
    }
    if (use_ring) fence_proxy_async();  // X, U written above are bulk-copied by the next sweep
    st[ST_DX] = valid ? dxm : 0.0;
    st[ST_DU] = valid ? dum : 0.0;
    st[ST_OBJ] = valid ? obj : 0.0;
    st[ST_NAN] = nanf;
  }

  // One convex sub-problem on the fused path.  Returns status bits (as solve_qp), or -1 when the
  // active-set iteration has to be globalised by ADMM (caller falls back to the unfused path).
  __device__ __forceinline__ int fused_qp(double* st, long long* cnt) {
    double pc[NX], zc[NU];
    const bool constrained = pb.has_ubox || pb.has_xbox;
    for (int k = 0; k < ST_COUNT; ++k) st[k] = 0.0;
    st[ST_GSCALE] = st[ST_ZSCALE] = 1.0;
    bool kkt = false;
    for (int tries = 0;; ++tries) {
      // backward: (verify the iterate, move the active set,) factor and solve
      bsweep(kkt, !kkt, st, pc, zc);
      cnt[2] += 1;
      cnt[7] += 1;
      if (kkt) {
        reduce_stats(st);
        if (st[ST_NAN] > 0.0) return 1;
        const double tol_g = pb.kkt_tol * st[ST_GSCALE], tol_z = pb.kkt_tol * st[ST_ZSCALE];
        if (st[ST_KKT] <= tol_g && st[ST_UVIOL] <= tol_z && st[ST_XVIOL] <= tol_z) {
          cnt[5] += 1;
          return 0;
        }
        if (st[ST_XVIOL] > tol_z || st[ST_CHANGED] <= 0.0 || tries >= pb.polish_iters) return -1;
      }
      build_shared_system(MODE_POLISH, 0.0, 0.0);
      solve_shared(MODE_POLISH, 0.0, 0.0, pc, zc);
      fsweep(st);
      cnt[3] += 1;
      cnt[7] += 1;
      if (!constrained) {
        reduce_stats(st);
        cnt[5] += 1;
        return (st[ST_NAN] > 0.0) ? 1 : 0;
      }
      kkt = true;
    }
  }

