/* CPU oracle of the particle-SCP solve path in plain C.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load this; the product library (libpmpc_b200.so) never does.
 *
 * Restates, for ONE hot path (pmpc.solve -> aff_solve -> c_lqp_solve -> lqp_solve -> OSQP), what the
 * reference (StanfordASL/pmpc) does on the CPU:
 *
 *   dubins_f_fx_fu      tests/dubins_car.py:48-90 (car) + its autograd Jacobians (:11-30), in the
 *                       cancellation-free closed form of oracle/dubins.py (same function)
 *   lqp_build           PMPC.jl/src/lqp_utils.jl:2-216 (P, q), :219-303 (A, b), :306-393 (G, l, u);
 *                       sentinels c_interface.jl:56-70; Nc < 0 -> N main.jl:127-128
 *   osqp_solve          OSQP 0.6.2 (OSQP_jll 0.600.200, PMPC.jl/Manifest.toml:294-304; call sites
 *                       osqp_solver.jl:40,48,50,67).  THIRD-PARTY, source not under /root/reference:
 *                       the *published* algorithm is restated (Stellato et al., Math. Prog. Comp. 2020,
 *                       Alg. 1 + Sec. 5: Ruiz equilibration, rho_eq = 1e3 rho, adaptive rho, residual
 *                       test every 25 iterations) with the defaults the reference runs with.  The KKT
 *                       system is factored by an up-looking sparse LDL' (the elimination-tree algorithm of
 *                       T. Davis' LDL package, which QDLDL also implements); the fill-reducing order is a
 *                       stage-interleaved order with the consensus block last instead of AMD.
 *   split / bookkeeping lqp_utils.jl:395-423; pmpc/scp_mpc.py:337-428 (loop, max_res, obj, stop rule)
 *
 * Parity status: pinned through the NumPy oracle (oracle/lqp.py, osqp_port.py, dubins.py, scp.py; tests/test_oracle_cpu.py),
 * which is pinned against the reference's own Dubins model and SCP loop executed here (tests/golden/) and against the
 * iteration tables the reference's real solver stack printed into its notebooks (4 digits;
 * tests/test_reference_notebook_traces.py).  No full-precision output of the reference's solver exists.
 *
 * Layouts are NumPy C-order as `pmpc.solve` sees them: vectors (M,N,d), matrices (M,N,rows,cols).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define OSQP_RHO_MIN 1e-6
#define OSQP_RHO_MAX 1e6
#define OSQP_RHO_EQ_FACTOR 1e3
#define OSQP_MIN_SCALING 1e-4
#define OSQP_MAX_SCALING 1e4
#define OSQP_RHO_TOL 1e-4
#define OSQP_INFTY_SCALED 1e26 /* OSQP_INFTY * MIN_SCALING */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------
 * sparse helpers: compressed sparse column, built from triplets (duplicates summed)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int nrow, ncol, nnz;
  int *p, *i;
  double* x;
} csc;

typedef struct {
  int n, cap;
  int *r, *c;
  double* v;
} trip;

static void trip_init(trip* t, int cap) {
  t->n = 0;
  t->cap = cap > 16 ? cap : 16;
  t->r = (int*)malloc(sizeof(int) * t->cap);
  t->c = (int*)malloc(sizeof(int) * t->cap);
  t->v = (double*)malloc(sizeof(double) * t->cap);
}
static void trip_add(trip* t, int r, int c, double v) {
  if (t->n == t->cap) {
    t->cap *= 2;
    t->r = (int*)realloc(t->r, sizeof(int) * t->cap);
    t->c = (int*)realloc(t->c, sizeof(int) * t->cap);
    t->v = (double*)realloc(t->v, sizeof(double) * t->cap);
  }
  t->r[t->n] = r; t->c[t->n] = c; t->v[t->n] = v; t->n++;
}
static void trip_free(trip* t) { free(t->r); free(t->c); free(t->v); }

static void csc_free(csc* A) {
  if (!A) return;
  free(A->p); free(A->i); free(A->x); free(A);
}

/* triplets -> CSC with sorted row indices and summed duplicates (two counting-sort passes) */
static csc* csc_from_trip(int nrow, int ncol, const trip* t) {
  const int nt = t->n;
  int* cnt = (int*)calloc((size_t)(nrow > ncol ? nrow : ncol) + 1, sizeof(int));
  int* ord = (int*)malloc(sizeof(int) * (nt > 0 ? nt : 1));
  int* ord2 = (int*)malloc(sizeof(int) * (nt > 0 ? nt : 1));
  /* sort by row */
  for (int k = 0; k < nt; ++k) cnt[t->r[k] + 1]++;
  for (int r = 0; r < nrow; ++r) cnt[r + 1] += cnt[r];
  for (int k = 0; k < nt; ++k) ord[cnt[t->r[k]]++] = k;
  /* stable sort by column */
  memset(cnt, 0, sizeof(int) * ((size_t)(nrow > ncol ? nrow : ncol) + 1));
  for (int k = 0; k < nt; ++k) cnt[t->c[k] + 1]++;
  for (int c = 0; c < ncol; ++c) cnt[c + 1] += cnt[c];
  for (int k = 0; k < nt; ++k) ord2[cnt[t->c[ord[k]]]++] = ord[k];
  csc* A = (csc*)malloc(sizeof(csc));
  A->nrow = nrow; A->ncol = ncol;
  A->p = (int*)calloc((size_t)ncol + 1, sizeof(int));
  A->i = (int*)malloc(sizeof(int) * (nt > 0 ? nt : 1));
  A->x = (double*)malloc(sizeof(double) * (nt > 0 ? nt : 1));
  int nnz = 0, k = 0;
  for (int c = 0; c < ncol; ++c) {
    A->p[c] = nnz;
    while (k < nt && t->c[ord2[k]] == c) {
      const int r = t->r[ord2[k]];
      double v = 0.0;
      while (k < nt && t->c[ord2[k]] == c && t->r[ord2[k]] == r) v += t->v[ord2[k++]];
      A->i[nnz] = r; A->x[nnz] = v; nnz++;
    }
  }
  A->p[ncol] = nnz;
  A->nnz = nnz;
  free(cnt); free(ord); free(ord2);
  return A;
}

static void csc_mv(const csc* A, const double* x, double* y) { /* y = A x */
  memset(y, 0, sizeof(double) * A->nrow);
  for (int c = 0; c < A->ncol; ++c) {
    const double xc = x[c];
    for (int p = A->p[c]; p < A->p[c + 1]; ++p) y[A->i[p]] += A->x[p] * xc;
  }
}
static void csc_mtv(const csc* A, const double* x, double* y) { /* y = A' x */
  for (int c = 0; c < A->ncol; ++c) {
    double acc = 0.0;
    for (int p = A->p[c]; p < A->p[c + 1]; ++p) acc += A->x[p] * x[A->i[p]];
    y[c] = acc;
  }
}
static void csc_symu_mv(const csc* Pu, const double* x, double* y) { /* y = P x, P symmetric, upper stored */
  memset(y, 0, sizeof(double) * Pu->ncol);
  for (int c = 0; c < Pu->ncol; ++c)
    for (int p = Pu->p[c]; p < Pu->p[c + 1]; ++p) {
      const int r = Pu->i[p];
      y[r] += Pu->x[p] * x[c];
      if (r != c) y[c] += Pu->x[p] * x[r];
    }
}
static double norm_inf(const double* v, int n) {
  double m = 0.0;
  for (int k = 0; k < n; ++k) { const double a = fabs(v[k]); if (a > m) m = a; }
  return m;
}

/* ------------------------------------------------------------------------------------------------
 * up-looking sparse LDL' of a symmetric quasi-definite matrix given by its upper triangle (CSC).
 * Elimination tree + column counts, then the numeric phase, as in T. Davis' LDL / QDLDL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n;
  int *Lp, *Li, *parent, *Lnz, *flag, *pattern;
  double *Lx, *D, *Dinv, *y;
} ldl_t;

static void ldl_free(ldl_t* F) {
  if (!F) return;
  free(F->Lp); free(F->Li); free(F->parent); free(F->Lnz); free(F->flag); free(F->pattern);
  free(F->Lx); free(F->D); free(F->Dinv); free(F->y); free(F);
}

static ldl_t* ldl_symbolic(const csc* K) {
  const int n = K->ncol;
  ldl_t* F = (ldl_t*)calloc(1, sizeof(ldl_t));
  F->n = n;
  F->Lp = (int*)malloc(sizeof(int) * ((size_t)n + 1));
  F->parent = (int*)malloc(sizeof(int) * n);
  F->Lnz = (int*)malloc(sizeof(int) * n);
  F->flag = (int*)malloc(sizeof(int) * n);
  F->pattern = (int*)malloc(sizeof(int) * n);
  F->D = (double*)malloc(sizeof(double) * n);
  F->Dinv = (double*)malloc(sizeof(double) * n);
  F->y = (double*)malloc(sizeof(double) * n);
  for (int k = 0; k < n; ++k) {
    F->parent[k] = -1; F->flag[k] = k; F->Lnz[k] = 0;
    for (int p = K->p[k]; p < K->p[k + 1]; ++p) {
      int i = K->i[p];
      if (i < k)
        for (; F->flag[i] != k; i = F->parent[i]) {
          if (F->parent[i] == -1) F->parent[i] = k;
          F->Lnz[i]++;
          F->flag[i] = k;
        }
    }
  }
  F->Lp[0] = 0;
  for (int k = 0; k < n; ++k) F->Lp[k + 1] = F->Lp[k] + F->Lnz[k];
  F->Li = (int*)malloc(sizeof(int) * (F->Lp[n] > 0 ? F->Lp[n] : 1));
  F->Lx = (double*)malloc(sizeof(double) * (F->Lp[n] > 0 ? F->Lp[n] : 1));
  return F;
}

static int ldl_numeric(ldl_t* F, const csc* K) {
  const int n = F->n;
  double* Y = F->y;
  for (int k = 0; k < n; ++k) { Y[k] = 0.0; F->Lnz[k] = 0; }
  for (int k = 0; k < n; ++k) {
    int top = n;
    F->flag[k] = k;
    for (int p = K->p[k]; p < K->p[k + 1]; ++p) {
      int i = K->i[p];
      if (i <= k) {
        Y[i] += K->x[p];
        int len = 0;
        for (; F->flag[i] != k; i = F->parent[i]) {
          F->pattern[len++] = i;
          F->flag[i] = k;
        }
        while (len > 0) F->pattern[--top] = F->pattern[--len];
      }
    }
    F->D[k] = Y[k];
    Y[k] = 0.0;
    for (; top < n; ++top) {
      const int i = F->pattern[top];
      const double yi = Y[i];
      Y[i] = 0.0;
      const int p2 = F->Lp[i] + F->Lnz[i];
      for (int p = F->Lp[i]; p < p2; ++p) Y[F->Li[p]] -= F->Lx[p] * yi;
      const double lki = yi * F->Dinv[i];
      F->D[k] -= lki * yi;
      F->Li[p2] = k;
      F->Lx[p2] = lki;
      F->Lnz[i]++;
    }
    if (F->D[k] == 0.0) return -1;
    F->Dinv[k] = 1.0 / F->D[k];
  }
  return 0;
}

static void ldl_solve(const ldl_t* F, double* x) { /* in place: x <- (L D L')^-1 x */
  const int n = F->n;
  for (int j = 0; j < n; ++j) {
    const double xj = x[j];
    for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) x[F->Li[p]] -= F->Lx[p] * xj;
  }
  for (int j = 0; j < n; ++j) x[j] *= F->Dinv[j];
  for (int j = n - 1; j >= 0; --j) {
    double xj = x[j];
    for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) xj -= F->Lx[p] * x[F->Li[p]];
    x[j] = xj;
  }
}

/* ------------------------------------------------------------------------------------------------
 * the QP of one SCP iteration (lqp_utils.jl) :  min 1/2 z'Pz + q'z   s.t.  la <= Aa z <= ua,
 * Aa = [A; G] (osqp_solver.jl:20-32), z = [u_c (Nc*m); u_free particle-major; x particle-major]
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int M, N, n, m, Nc, nz, ne, ni;
  csc *Pu, *Aa;
  double *q, *la, *ua;
  int* order; /* elimination order key of every KKT index (variables then constraint rows) */
} qp_t;

static void qp_free(qp_t* Q) {
  if (!Q) return;
  csc_free(Q->Pu); csc_free(Q->Aa);
  free(Q->q); free(Q->la); free(Q->ua); free(Q->order); free(Q);
}

static int any_nan(const double* a, size_t n) {
  if (!a) return 1;
  for (size_t k = 0; k < n; ++k) if (a[k] != a[k]) return 1;
  return 0;
}

static inline int iu_of(int M, int N, int m, int Nc, int i, int j, int a) {
  (void)M;
  return (j < Nc) ? j * m + a : Nc * m + (N - Nc) * m * i + (j - Nc) * m + a;
}
static inline int ix_of(int M, int N, int n, int m, int Nc, int i, int j, int r) {
  return Nc * m + M * (N - Nc) * m + N * n * i + n * j + r;
}

/* elimination-order key: particle-major, stage-interleaved, consensus block last */
static inline int okey(int M, int N, int i, int j, int kind) { /* kind 0..4 within a stage */
  (void)M;
  return ((i * N + j) * 5 + kind);
}

static qp_t* lqp_build(int M, int N, int n, int m, long long Nc_in, const double* x0, const double* f, const double* fx,
                       const double* fu, const double* X_prev, const double* U_prev, const double* Q, const double* R,
                       const double* X_ref, const double* U_ref, const double* lx, const double* ux, const double* lu,
                       const double* uu, double reg_x, double reg_u, const double* slew_reg, const double* slew_reg0,
                       const double* slew_um1) {
  (void)x0;
  const int Nc = (Nc_in < 0 || Nc_in > N) ? N : (int)Nc_in; /* main.jl:127-128 */
  const int Nf = N - Nc;
  const int nz = Nc * m + M * (Nf * m + N * n);
  const int ne = M * N * n;
  /* sentinels: c_interface.jl:56-70 */
  const int has_s = !any_nan(slew_reg, (size_t)M);
  const int has_s0 = !(any_nan(slew_reg0, (size_t)M) || any_nan(slew_um1, (size_t)M * m));
  const int has_ubox = !(any_nan(lu, (size_t)M * N * m) || any_nan(uu, (size_t)M * N * m));
  const int has_xbox = !(any_nan(lx, (size_t)M * N * n) || any_nan(ux, (size_t)M * N * n));
  const int ni = (has_ubox ? Nc * m + M * Nf * m : 0) + (has_xbox ? M * N * n : 0);
  qp_t* qp = (qp_t*)calloc(1, sizeof(qp_t));
  qp->M = M; qp->N = N; qp->n = n; qp->m = m; qp->Nc = Nc; qp->nz = nz; qp->ne = ne; qp->ni = ni;
  qp->q = (double*)calloc((size_t)nz, sizeof(double));
  qp->la = (double*)malloc(sizeof(double) * ((size_t)ne + ni));
  qp->ua = (double*)malloc(sizeof(double) * ((size_t)ne + ni));
  qp->order = (int*)malloc(sizeof(int) * ((size_t)nz + ne + ni));
  const int key_last = M * N * 5; /* consensus variables and their rows come after every particle */
  trip tp, ta;
  trip_init(&tp, M * N * (n * n + m * m + 2 * m));
  trip_init(&ta, M * N * (n * n + n * m + n) + ni);
  /* ---- P, upper triangle (OSQP reads the upper triangle only); lqp_utils.jl:17-160 */
  for (int i = 0; i < M; ++i) {
    const double s = has_s ? slew_reg[i] : 0.0, s0 = has_s0 ? slew_reg0[i] : 0.0;
    for (int j = 0; j < N; ++j) {
      const double* Rj = R + ((size_t)i * N + j) * m * m;
      const double* Qj = Q + ((size_t)i * N + j) * n * n;
      double dg = reg_u + 2.0 * s;               /* lqp_utils.jl:30-38 */
      if (j == N - 1) dg = reg_u + s;
      if (j == 0) dg = reg_u + s0 + s;           /* j == 1 wins over j == N when N == 1 */
      for (int a = 0; a < m; ++a) {
        const int ia = iu_of(M, N, m, Nc, i, j, a);
        for (int b = 0; b < m; ++b) {
          const int ib = iu_of(M, N, m, Nc, i, j, b);
          if (ia <= ib) trip_add(&tp, ia, ib, Rj[a * m + b] + (a == b ? dg : 0.0));
        }
        if (j + 1 < N) { /* -s between consecutive controls (:21-25,47-57,67-75,94-98) */
          const int in = iu_of(M, N, m, Nc, i, j + 1, a);
          trip_add(&tp, ia < in ? ia : in, ia < in ? in : ia, -s);
        }
      }
      for (int r = 0; r < n; ++r) {
        const int ir = ix_of(M, N, n, m, Nc, i, j, r);
        for (int t = r; t < n; ++t) trip_add(&tp, ir, ir + (t - r), Qj[r * n + t] + (r == t ? reg_x : 0.0));
      }
    }
  }
  qp->Pu = csc_from_trip(nz, nz, &tp);
  trip_free(&tp);
  /* ---- q (lqp_utils.jl:164-214) */
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      const double* Rj = R + ((size_t)i * N + j) * m * m;
      const double* Qj = Q + ((size_t)i * N + j) * n * n;
      const double* ur = U_ref + ((size_t)i * N + j) * m;
      const double* xr = X_ref + ((size_t)i * N + j) * n;
      for (int a = 0; a < m; ++a) {
        double acc = -reg_u * U_prev[((size_t)i * N + j) * m + a];
        for (int b = 0; b < m; ++b) acc -= Rj[a * m + b] * ur[b];
        qp->q[iu_of(M, N, m, Nc, i, j, a)] += acc;
      }
      for (int r = 0; r < n; ++r) {
        double acc = -reg_x * X_prev[((size_t)i * N + j) * n + r];
        for (int t = 0; t < n; ++t) acc -= Qj[r * n + t] * xr[t];
        qp->q[ix_of(M, N, n, m, Nc, i, j, r)] = acc;
      }
    }
  if (Nc >= 1 && has_s0) /* the -s0 u_{-1} term; with Nc == 0 the reference overwrites it (:165 vs :187, App. B.6) */
    for (int i = 0; i < M; ++i)
      for (int a = 0; a < m; ++a) qp->q[a] += -slew_reg0[i] * slew_um1[(size_t)i * m + a];
  /* ---- A, b (lqp_utils.jl:219-303) and G, l, u (:306-393) stacked */
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      const double* fuj = fu + ((size_t)i * N + j) * n * m;
      const double* fxj = fx + ((size_t)i * N + j) * n * n;
      for (int r = 0; r < n; ++r) {
        const int row = (i * N + j) * n + r;
        double b = -f[((size_t)i * N + j) * n + r];
        for (int a = 0; a < m; ++a) {
          trip_add(&ta, row, iu_of(M, N, m, Nc, i, j, a), fuj[r * m + a]);
          b += fuj[r * m + a] * U_prev[((size_t)i * N + j) * m + a];
        }
        trip_add(&ta, row, ix_of(M, N, n, m, Nc, i, j, r), -1.0);
        if (j >= 1) /* no fx term at the first stage: linearised at x0 itself (:288-296) */
          for (int t = 0; t < n; ++t) {
            trip_add(&ta, row, ix_of(M, N, n, m, Nc, i, j - 1, t), fxj[r * n + t]);
            b += fxj[r * n + t] * X_prev[((size_t)i * N + j - 1) * n + t];
          }
        qp->la[row] = qp->ua[row] = b;
        qp->order[nz + row] = okey(M, N, i, j, 2);
      }
    }
  int row = ne;
  if (has_ubox) {
    for (int j = 0; j < Nc; ++j) /* shared block: particle 1's bounds (:329-330) */
      for (int a = 0; a < m; ++a) {
        trip_add(&ta, row, iu_of(M, N, m, Nc, 0, j, a), 1.0);
        qp->la[row] = lu[((size_t)0 * N + j) * m + a];
        qp->ua[row] = uu[((size_t)0 * N + j) * m + a];
        qp->order[nz + row] = key_last + 1;
        row++;
      }
    for (int i = 0; i < M; ++i)
      for (int j = Nc; j < N; ++j)
        for (int a = 0; a < m; ++a) {
          trip_add(&ta, row, iu_of(M, N, m, Nc, i, j, a), 1.0);
          qp->la[row] = lu[((size_t)i * N + j) * m + a];
          qp->ua[row] = uu[((size_t)i * N + j) * m + a];
          qp->order[nz + row] = okey(M, N, i, j, 1);
          row++;
        }
  }
  if (has_xbox)
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j)
        for (int r = 0; r < n; ++r) {
          trip_add(&ta, row, ix_of(M, N, n, m, Nc, i, j, r), 1.0);
          qp->la[row] = lx[((size_t)i * N + j) * n + r];
          qp->ua[row] = ux[((size_t)i * N + j) * n + r];
          qp->order[nz + row] = okey(M, N, i, j, 4);
          row++;
        }
  qp->Aa = csc_from_trip(ne + ni, nz, &ta);
  trip_free(&ta);
  for (int k = 0; k < Nc * m; ++k) qp->order[k] = key_last;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      if (j >= Nc)
        for (int a = 0; a < m; ++a) qp->order[iu_of(M, N, m, Nc, i, j, a)] = okey(M, N, i, j, 0);
      for (int r = 0; r < n; ++r) qp->order[ix_of(M, N, n, m, Nc, i, j, r)] = okey(M, N, i, j, 3);
    }
  return qp;
}

/* ------------------------------------------------------------------------------------------------
 * OSQP 0.6.2 restated (see the header).  Settings struct mirrors the OSQP names.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, adaptive_rho_tolerance, adaptive_rho_fraction;
  int max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination;
} osqp_settings_t;

void pmpc_oracle_default_settings(osqp_settings_t* s) {
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->adaptive_rho_tolerance = 5.0; s->adaptive_rho_fraction = 0.4;
  s->max_iter = 4000; s->scaling = 10; s->adaptive_rho = 1; s->adaptive_rho_interval = 0; s->check_termination = 25;
}

typedef struct {
  int iters, status, rho_updates; /* status 1 solved, 0 max_iter reached, -1 factorisation failed */
  double r_prim, r_dual, setup_time, solve_time;
} osqp_info_t;

static double limit_scaling(double v) {
  if (v < OSQP_MIN_SCALING) v = 1.0;
  return v > OSQP_MAX_SCALING ? OSQP_MAX_SCALING : v;
}

/* permuted upper-triangular KKT = [[P + sigma I, Aa'], [Aa, -diag(1/rho)]] */
static csc* kkt_build(const csc* Pu, const csc* Aa, double sigma, const double* rho_vec, const int* pinv) {
  const int n = Pu->ncol, mm = Aa->nrow;
  trip t;
  trip_init(&t, Pu->nnz + Aa->nnz + n + mm);
  for (int c = 0; c < n; ++c) {
    int diag = 0;
    for (int p = Pu->p[c]; p < Pu->p[c + 1]; ++p) {
      const int r = Pu->i[p];
      double v = Pu->x[p];
      if (r == c) { v += sigma; diag = 1; }
      const int a = pinv[r], b = pinv[c];
      trip_add(&t, a < b ? a : b, a < b ? b : a, v);
    }
    if (!diag) trip_add(&t, pinv[c], pinv[c], sigma);
  }
  for (int c = 0; c < n; ++c)
    for (int p = Aa->p[c]; p < Aa->p[c + 1]; ++p) {
      const int a = pinv[c], b = pinv[n + Aa->i[p]];
      trip_add(&t, a < b ? a : b, a < b ? b : a, Aa->x[p]);
    }
  for (int r = 0; r < mm; ++r) trip_add(&t, pinv[n + r], pinv[n + r], -1.0 / rho_vec[r]);
  csc* K = csc_from_trip(n + mm, n + mm, &t);
  trip_free(&t);
  return K;
}

typedef struct { int key, idx; } keyidx;
static int cmp_keyidx(const void* a, const void* b) {
  const keyidx *x = (const keyidx*)a, *y = (const keyidx*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static void osqp_solve(const qp_t* qp, const osqp_settings_t* st, double* x_out, osqp_info_t* info) {
  const double t0 = now_s();
  const int n = qp->nz, mm = qp->ne + qp->ni;
  /* working (scaled) copies */
  csc Ps = *qp->Pu, As = *qp->Aa;
  Ps.x = (double*)malloc(sizeof(double) * (Ps.nnz > 0 ? Ps.nnz : 1));
  As.x = (double*)malloc(sizeof(double) * (As.nnz > 0 ? As.nnz : 1));
  memcpy(Ps.x, qp->Pu->x, sizeof(double) * Ps.nnz);
  memcpy(As.x, qp->Aa->x, sizeof(double) * As.nnz);
  double* q = (double*)malloc(sizeof(double) * n);
  memcpy(q, qp->q, sizeof(double) * n);
  double* D = (double*)malloc(sizeof(double) * n);
  double* E = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  double* dtmp = (double*)malloc(sizeof(double) * n);
  double* etmp = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  for (int k = 0; k < n; ++k) D[k] = 1.0;
  for (int k = 0; k < mm; ++k) E[k] = 1.0;
  double c = 1.0;
  /* Ruiz equilibration (OSQP scaling.c: scale_data) */
  for (int it = 0; it < st->scaling; ++it) {
    for (int k = 0; k < n; ++k) dtmp[k] = 0.0;
    for (int k = 0; k < mm; ++k) etmp[k] = 0.0;
    for (int cc = 0; cc < n; ++cc) {
      for (int p = Ps.p[cc]; p < Ps.p[cc + 1]; ++p) {
        const double a = fabs(Ps.x[p]);
        if (a > dtmp[cc]) dtmp[cc] = a;
        if (a > dtmp[Ps.i[p]]) dtmp[Ps.i[p]] = a;
      }
      for (int p = As.p[cc]; p < As.p[cc + 1]; ++p) {
        const double a = fabs(As.x[p]);
        if (a > dtmp[cc]) dtmp[cc] = a;
        if (a > etmp[As.i[p]]) etmp[As.i[p]] = a;
      }
    }
    for (int k = 0; k < n; ++k) dtmp[k] = 1.0 / sqrt(limit_scaling(dtmp[k]));
    for (int k = 0; k < mm; ++k) etmp[k] = 1.0 / sqrt(limit_scaling(etmp[k]));
    for (int cc = 0; cc < n; ++cc) {
      for (int p = Ps.p[cc]; p < Ps.p[cc + 1]; ++p) Ps.x[p] *= dtmp[cc] * dtmp[Ps.i[p]];
      for (int p = As.p[cc]; p < As.p[cc + 1]; ++p) As.x[p] *= dtmp[cc] * etmp[As.i[p]];
      q[cc] *= dtmp[cc];
      D[cc] *= dtmp[cc];
    }
    for (int k = 0; k < mm; ++k) E[k] *= etmp[k];
    /* cost scaling */
    for (int k = 0; k < n; ++k) dtmp[k] = 0.0;
    for (int cc = 0; cc < n; ++cc)
      for (int p = Ps.p[cc]; p < Ps.p[cc + 1]; ++p) {
        const double a = fabs(Ps.x[p]);
        if (a > dtmp[cc]) dtmp[cc] = a;
        if (a > dtmp[Ps.i[p]]) dtmp[Ps.i[p]] = a;
      }
    double mean = 0.0;
    for (int k = 0; k < n; ++k) mean += dtmp[k];
    mean = n > 0 ? mean / n : 0.0;
    double cost = fmax(mean, norm_inf(q, n));
    cost = cost < OSQP_MIN_SCALING ? 1.0 : fmin(cost, OSQP_MAX_SCALING);
    const double ct = 1.0 / cost;
    for (int p = 0; p < Ps.nnz; ++p) Ps.x[p] *= ct;
    for (int k = 0; k < n; ++k) q[k] *= ct;
    c *= ct;
  }
  double* l = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  double* u = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  double* rv = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  char* ctype = (char*)malloc((size_t)(mm > 0 ? mm : 1)); /* 0 inequality, 1 equality, 2 free */
  for (int k = 0; k < mm; ++k) {
    l[k] = E[k] * qp->la[k];
    u[k] = E[k] * qp->ua[k];
    ctype[k] = (qp->ua[k] - qp->la[k] < OSQP_RHO_TOL) ? 1 : ((qp->la[k] < -OSQP_INFTY_SCALED && qp->ua[k] > OSQP_INFTY_SCALED) ? 2 : 0);
  }
  double rho = st->rho;
#define SET_RHO_VEC() \
  for (int k = 0; k < mm; ++k) rv[k] = ctype[k] == 1 ? OSQP_RHO_EQ_FACTOR * rho : (ctype[k] == 2 ? OSQP_RHO_MIN : rho)
  SET_RHO_VEC();
  /* elimination order */
  const int nk = n + mm;
  keyidx* ki = (keyidx*)malloc(sizeof(keyidx) * nk);
  for (int k = 0; k < nk; ++k) { ki[k].key = qp->order[k]; ki[k].idx = k; }
  qsort(ki, nk, sizeof(keyidx), cmp_keyidx);
  int* perm = (int*)malloc(sizeof(int) * nk);
  int* pinv = (int*)malloc(sizeof(int) * nk);
  for (int k = 0; k < nk; ++k) { perm[k] = ki[k].idx; pinv[ki[k].idx] = k; }
  free(ki);
  csc* K = kkt_build(&Ps, &As, st->sigma, rv, pinv);
  ldl_t* F = ldl_symbolic(K);
  int fail = ldl_numeric(F, K);
  const double setup_time = now_s() - t0;
  double* x = (double*)calloc(n, sizeof(double));
  double* z = (double*)calloc(mm > 0 ? mm : 1, sizeof(double));
  double* y = (double*)calloc(mm > 0 ? mm : 1, sizeof(double));
  double* rhs = (double*)malloc(sizeof(double) * nk);
  double* sol = (double*)malloc(sizeof(double) * nk);
  double* Ax = (double*)malloc(sizeof(double) * (mm > 0 ? mm : 1));
  double* Px = (double*)malloc(sizeof(double) * n);
  double* Aty = (double*)malloc(sizeof(double) * n);
  info->status = fail ? -1 : 0;
  info->rho_updates = 0;
  info->r_prim = info->r_dual = INFINITY;
  const double t1 = now_s();
  int adapt_interval = st->adaptive_rho_interval, it = 0;
  for (it = 1; !fail && it <= st->max_iter; ++it) {
    for (int k = 0; k < n; ++k) rhs[pinv[k]] = st->sigma * x[k] - q[k];
    for (int k = 0; k < mm; ++k) rhs[pinv[n + k]] = z[k] - y[k] / rv[k];
    ldl_solve(F, rhs);
    for (int k = 0; k < nk; ++k) sol[perm[k]] = rhs[k];
    for (int k = 0; k < n; ++k) x[k] = st->alpha * sol[k] + (1.0 - st->alpha) * x[k];
    for (int k = 0; k < mm; ++k) {
      const double zt = z[k] + (sol[n + k] - y[k]) / rv[k];
      const double zr = st->alpha * zt + (1.0 - st->alpha) * z[k];
      double zn = zr + y[k] / rv[k];
      zn = zn < l[k] ? l[k] : (zn > u[k] ? u[k] : zn);
      y[k] += rv[k] * (zr - zn);
      z[k] = zn;
    }
    if (st->adaptive_rho && adapt_interval == 0 && now_s() - t1 > st->adaptive_rho_fraction * setup_time) {
      const int kk = st->check_termination ? st->check_termination : 25;
      int r = (int)floor((double)it / kk + 0.5) * kk;
      adapt_interval = r > kk ? r : kk;
    }
    const int check = st->check_termination && it % st->check_termination == 0;
    const int adapt = st->adaptive_rho && adapt_interval && it % adapt_interval == 0;
    if (check || adapt || it == st->max_iter) {
      csc_mv(&As, x, Ax);
      csc_symu_mv(&Ps, x, Px);
      csc_mtv(&As, y, Aty);
      double rp = 0, rd = 0, npa = 0, npz = 0, ndp = 0, nda = 0, ndq = 0;
      double rps = 0, rds = 0, nps = 0, nds = 0;
      for (int k = 0; k < mm; ++k) {
        const double ei = 1.0 / E[k];
        rp = fmax(rp, fabs(ei * (Ax[k] - z[k])));
        npa = fmax(npa, fabs(ei * Ax[k]));
        npz = fmax(npz, fabs(ei * z[k]));
        rps = fmax(rps, fabs(Ax[k] - z[k]));
        nps = fmax(nps, fmax(fabs(Ax[k]), fabs(z[k])));
      }
      for (int k = 0; k < n; ++k) {
        const double di = 1.0 / D[k];
        rd = fmax(rd, fabs(di * (Px[k] + q[k] + Aty[k])));
        ndp = fmax(ndp, fabs(di * Px[k]));
        nda = fmax(nda, fabs(di * Aty[k]));
        ndq = fmax(ndq, fabs(di * q[k]));
        rds = fmax(rds, fabs(Px[k] + q[k] + Aty[k]));
        nds = fmax(nds, fmax(fabs(Px[k]), fmax(fabs(Aty[k]), fabs(q[k]))));
      }
      rd /= c;
      const double n_prim = fmax(npa, npz), n_dual = fmax(ndp, fmax(nda, ndq)) / c;
      info->r_prim = rp;
      info->r_dual = rd;
      if (check && rp <= st->eps_abs + st->eps_rel * n_prim && rd <= st->eps_abs + st->eps_rel * n_dual) {
        info->status = 1;
        break;
      }
      if (adapt) { /* compute_rho_estimate on the scaled residuals */
        double est = rho * sqrt((rps / (nps + 1e-10)) / (rds / (nds + 1e-10) + 1e-30));
        est = est < OSQP_RHO_MIN ? OSQP_RHO_MIN : (est > OSQP_RHO_MAX ? OSQP_RHO_MAX : est);
        if (est > rho * st->adaptive_rho_tolerance || est < rho / st->adaptive_rho_tolerance) {
          rho = est;
          SET_RHO_VEC();
          csc_free(K);
          K = kkt_build(&Ps, &As, st->sigma, rv, pinv);
          fail = ldl_numeric(F, K); /* same pattern: numeric refactorisation only */
          if (fail) info->status = -1;
          info->rho_updates++;
        }
      }
    }
  }
#undef SET_RHO_VEC
  info->iters = it > st->max_iter ? st->max_iter : it;
  info->setup_time = setup_time;
  info->solve_time = now_s() - t1;
  for (int k = 0; k < n; ++k) x_out[k] = fail ? NAN : D[k] * x[k]; /* NaN on failure: osqp_solver.jl:69-71 */
  csc_free(K); ldl_free(F);
  free(Ps.x); free(As.x); free(q); free(D); free(E); free(dtmp); free(etmp); free(l); free(u); free(rv); free(ctype);
  free(perm); free(pinv); free(x); free(z); free(y); free(rhs); free(sol); free(Ax); free(Px); free(Aty);
}

/* one sub-problem: c_lqp_solve (c_interface.jl:77-141) in Python layout.  X_out (M,N,n) without x0, U_out (M,N,m). */
int pmpc_oracle_lqp_solve(double* X_out, double* U_out, int n, int m, int N, int M, long long Nc, const double* x0,
                          const double* f, const double* fx, const double* fu, const double* X_prev, const double* U_prev,
                          const double* Q, const double* R, const double* X_ref, const double* U_ref, const double* lx,
                          const double* ux, const double* lu, const double* uu, double reg_x, double reg_u,
                          const double* slew_reg, const double* slew_reg0, const double* slew_um1,
                          const osqp_settings_t* settings, osqp_info_t* info) {
  osqp_settings_t st;
  if (settings) st = *settings; else pmpc_oracle_default_settings(&st);
  qp_t* qp = lqp_build(M, N, n, m, Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u,
                       slew_reg, slew_reg0, slew_um1);
  double* zsol = (double*)malloc(sizeof(double) * qp->nz);
  osqp_info_t inf;
  osqp_solve(qp, &st, zsol, &inf);
  if (info) *info = inf;
  /* split_lqp_vars (lqp_utils.jl:395-423): the consensus block is broadcast to every particle */
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      for (int a = 0; a < m; ++a) U_out[((size_t)i * N + j) * m + a] = zsol[iu_of(M, N, m, qp->Nc, i, j, a)];
      for (int r = 0; r < n; ++r) X_out[((size_t)i * N + j) * n + r] = zsol[ix_of(M, N, n, m, qp->Nc, i, j, r)];
    }
  free(zsol);
  qp_free(qp);
  return inf.status;
}

/* ------------------------------------------------------------------------------------------------
 * Dubins car (tests/dubins_car.py:48-90), closed-form Jacobians in the cancellation-free form of
 * oracle/dubins.py::car_f_fx_fu.  fx row-major (4,4), fu (4,2).
 * ---------------------------------------------------------------------------------------------- */
static double entire_fn(double D, int j) { /* sum_k (-1)^k D^(2k) / (2k+j)!,  j = 1..5 */
  if (fabs(D) > 2.0) {
    switch (j) {
      case 1: return sin(D) / D;
      case 2: return (1.0 - cos(D)) / (D * D);
      case 3: return (D - sin(D)) / (D * D * D);
      case 4: return (D * D / 2.0 - 1.0 + cos(D)) / (D * D * D * D);
      default: return (D * D * D / 6.0 - D + sin(D)) / (D * D * D * D * D);
    }
  }
  const double D2 = D * D;
  double acc = 0.0;
  for (int k = 11; k >= 0; --k) {
    double fact = 1.0;
    for (int t = 2; t <= 2 * k + j; ++t) fact *= (double)t;
    acc = acc * D2 + ((k & 1) ? -1.0 : 1.0) / fact;
  }
  return acc;
}

void pmpc_oracle_dubins(const double* x, const double* uin, const double* p, double eps, double* f, double* fx, double* fu) {
  const double vs = p[0], ws = p[1], T = p[2];
  double b = vs * uin[0], a = ws * -uin[1];
  b += (b >= 0.0) ? eps : -eps;
  a += (a >= 0.0) ? eps : -eps;
  const double x0 = x[0], y0 = x[1], v0 = x[2], th0 = x[3];
  const double D = T * a;
  const double A = entire_fn(D, 1), B = entire_fn(D, 2), C = entire_fn(D, 3), Dd = entire_fn(D, 4), E = entire_fn(D, 5);
  const double s = sin(th0), c = cos(th0);
  const double P1 = A, P2 = D * B, P3 = D * (C - B), P4 = A - B;
  const double dP1 = P3, dP2 = B + D * D * (2 * Dd - C), dP3 = (C - B) + D * D * (3 * E - 3 * Dd + C), dP4 = D * (2 * C - B - 2 * Dd);
  const double vT = v0 * T, bT2 = b * T * T;
  f[0] = x0 + vT * (c * P1 - s * P2) + bT2 * (s * P3 + c * P4);
  f[1] = y0 + vT * (s * P1 + c * P2) + bT2 * (s * P4 - c * P3);
  f[2] = v0 + T * b;
  f[3] = T * a + th0;
  for (int k = 0; k < 16; ++k) fx[k] = 0.0;
  for (int k = 0; k < 8; ++k) fu[k] = 0.0;
  fx[0] = fx[5] = fx[10] = fx[15] = 1.0;
  fx[2] = T * (c * P1 - s * P2);
  fx[6] = T * (s * P1 + c * P2);
  fx[3] = vT * (-s * P1 - c * P2) + bT2 * (c * P3 - s * P4);
  fx[7] = vT * (c * P1 - s * P2) + bT2 * (c * P4 + s * P3);
  const double d1_b = T * T * (s * P3 + c * P4), d2_b = T * T * (s * P4 - c * P3);
  const double d1_a = T * (vT * (c * dP1 - s * dP2) + bT2 * (s * dP3 + c * dP4));
  const double d2_a = T * (vT * (s * dP1 + c * dP2) + bT2 * (s * dP4 - c * dP3));
  fu[0] = vs * d1_b; fu[1] = -ws * d1_a;
  fu[2] = vs * d2_b; fu[3] = -ws * d2_a;
  fu[4] = vs * T;
  fu[7] = -ws * T;
}

/* ------------------------------------------------------------------------------------------------
 * SCP loop (pmpc/scp_mpc.py:337-428) for Dubins particles, p (M,3).  X_out (M,N+1,n) with x0, U_out (M,N,m).
 * hist (max_it, 3): obj, resid, osqp iterations.  Returns the number of SCP iterations (<0: NaN failure).
 * ---------------------------------------------------------------------------------------------- */
int pmpc_oracle_scp_dubins(double* X_out, double* U_out, double* hist, int N, int M, long long Nc, const double* p, double eps,
                           const double* x0, const double* Q, const double* R, const double* X_ref, const double* U_ref,
                           const double* X_prev0, const double* U_prev0, const double* lu, const double* uu, double reg_x,
                           double reg_u, const double* slew_reg, int max_it, double res_tol,
                           const osqp_settings_t* settings) {
  const int n = 4, m = 2;
  const size_t nX = (size_t)M * N * n, nU = (size_t)M * N * m;
  double* Xp = (double*)malloc(sizeof(double) * nX);
  double* Up = (double*)malloc(sizeof(double) * nU);
  double* X = (double*)malloc(sizeof(double) * nX);
  double* U = (double*)malloc(sizeof(double) * nU);
  double* f = (double*)malloc(sizeof(double) * nX);
  double* fx = (double*)malloc(sizeof(double) * nX * n);
  double* fu = (double*)malloc(sizeof(double) * nX * m);
  memcpy(Xp, X_prev0 ? X_prev0 : X_ref, sizeof(double) * nX); /* default warm start: scp_mpc.py:313-314 */
  memcpy(Up, U_prev0 ? U_prev0 : U_ref, sizeof(double) * nU);
  int it = 0, rc = 0;
  while (it < max_it) {
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) { /* X_ = [x0, X_prev[:-1]]  (scp_mpc.py:338) */
        const double* xs = (j == 0) ? x0 + (size_t)i * n : Xp + ((size_t)i * N + j - 1) * n;
        const size_t o = (size_t)i * N + j;
        pmpc_oracle_dubins(xs, Up + o * m, p + (size_t)i * 3, eps, f + o * n, fx + o * n * n, fu + o * n * m);
      }
    osqp_info_t inf;
    pmpc_oracle_lqp_solve(X, U, n, m, N, M, Nc, x0, f, fx, fu, Xp, Up, Q, R, X_ref, U_ref, NULL, NULL, lu, uu, reg_x, reg_u,
                          slew_reg, NULL, NULL, settings, &inf);
    double max_res = 0.0, obj = 0.0;
    int nanf = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        const size_t o = (size_t)i * N + j;
        double n2 = 0.0;
        for (int r = 0; r < n; ++r) {
          const double d = X[o * n + r] - Xp[o * n + r];
          n2 += d * d;
          if (X[o * n + r] != X[o * n + r]) nanf = 1;
        }
        max_res = fmax(max_res, sqrt(n2)); /* per-stage 2-norms, scp_mpc.py:402-403 */
        n2 = 0.0;
        for (int a = 0; a < m; ++a) {
          const double d = U[o * m + a] - Up[o * m + a];
          n2 += d * d;
          if (U[o * m + a] != U[o * m + a]) nanf = 1;
        }
        max_res = fmax(max_res, sqrt(n2));
        for (int r = 0; r < n; ++r) { /* obj: un-augmented references, no 1/2 (scp_mpc.py:404-405) */
          double acc = 0.0;
          for (int t = 0; t < n; ++t) acc += Q[o * n * n + r * n + t] * (X[o * n + t] - X_ref[o * n + t]);
          obj += (X[o * n + r] - X_ref[o * n + r]) * acc;
        }
        for (int a = 0; a < m; ++a) {
          double acc = 0.0;
          for (int b = 0; b < m; ++b) acc += R[o * m * m + a * m + b] * (U[o * m + b] - U_ref[o * m + b]);
          obj += (U[o * m + a] - U_ref[o * m + a]) * acc;
        }
      }
    if (nanf) { rc = -1; break; } /* scp_mpc.py:391-394 */
    obj /= (double)N * (double)M;
    memcpy(Xp, X, sizeof(double) * nX);
    memcpy(Up, U, sizeof(double) * nU);
    if (hist) { hist[it * 3 + 0] = obj; hist[it * 3 + 1] = max_res; hist[it * 3 + 2] = (double)inf.iters; }
    ++it;
    if (max_res < res_tol) break; /* scp_mpc.py:424 */
  }
  if (rc == 0 && it > 0) {
    for (int i = 0; i < M; ++i) {
      memcpy(X_out + (size_t)i * (N + 1) * n, x0 + (size_t)i * n, sizeof(double) * n);
      memcpy(X_out + (size_t)i * (N + 1) * n + n, X + (size_t)i * N * n, sizeof(double) * N * n);
    }
    memcpy(U_out, U, sizeof(double) * nU);
  }
  free(Xp); free(Up); free(X); free(U); free(f); free(fx); free(fu);
  return rc < 0 ? rc : it;
}

/* ---- batch of independent problems on `nthreads` host threads (pmpc.solve_problems / remote.py worker pool:
 * the reference's CPU parallelism is across problems, OSQP itself is single-threaded) --------------------- */
typedef struct {
  int B, N, M, max_it, next;
  long long Nc;
  double eps, reg_x, reg_u, res_tol;
  const double *p, *x0, *Q, *R, *X_ref, *U_ref, *X_prev, *U_prev, *lu, *uu;
  double *X_out, *U_out;
  int* iters;
  const osqp_settings_t* st;
  pthread_mutex_t mu;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  const size_t n = 4, m = 2, M = b->M, N = b->N;
  for (;;) {
    pthread_mutex_lock(&b->mu);
    const int k = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (k >= b->B) break;
    const size_t oX = (size_t)k * M * N * n, oU = (size_t)k * M * N * m;
    b->iters[k] = pmpc_oracle_scp_dubins(b->X_out + (size_t)k * M * (N + 1) * n, b->U_out + oU, NULL, (int)N, (int)M, b->Nc,
                                         b->p + (size_t)k * M * 3, b->eps, b->x0 + (size_t)k * M * n, b->Q + oX * n, b->R + oU * m,
                                         b->X_ref + oX, b->U_ref + oU, b->X_prev ? b->X_prev + oX : NULL,
                                         b->U_prev ? b->U_prev + oU : NULL, b->lu ? b->lu + oU : NULL, b->uu ? b->uu + oU : NULL,
                                         b->reg_x, b->reg_u, NULL, b->max_it, b->res_tol, b->st);
  }
  return NULL;
}

int pmpc_oracle_scp_dubins_batch(int nthreads, int B, double* X_out, double* U_out, int* iters, int N, int M, long long Nc,
                                 const double* p, double eps, const double* x0, const double* Q, const double* R,
                                 const double* X_ref, const double* U_ref, const double* X_prev, const double* U_prev,
                                 const double* lu, const double* uu, double reg_x, double reg_u, int max_it, double res_tol,
                                 const osqp_settings_t* settings) {
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.B = B; b.N = N; b.M = M; b.Nc = Nc; b.max_it = max_it; b.eps = eps; b.reg_x = reg_x; b.reg_u = reg_u; b.res_tol = res_tol;
  b.p = p; b.x0 = x0; b.Q = Q; b.R = R; b.X_ref = X_ref; b.U_ref = U_ref; b.X_prev = X_prev; b.U_prev = U_prev; b.lu = lu; b.uu = uu;
  b.X_out = X_out; b.U_out = U_out; b.iters = iters; b.st = settings;
  pthread_mutex_init(&b.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &b);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&b.mu);
  return 0;
}
