// __global__ kernels of the CHMC hot path.  See mmd_kernels.cuh for the layout and the mapping.
#pragma once
#include "mmd_sweeps.cuh"
#include "mmd_philox.cuh"

namespace mmd {

// shared-memory carve-up of the phase kernels: [scratch NSCR*NT][crow NRMAX*NT][lam NRMAX*NT]; the scratch
// region is the cross-block reduction buffer or, inside a sweep, the thread-private prefetch ring
template <class M, int NRMAX, int UMAX>
struct SmemPlan {
  static constexpr int NRED = UMAX * (UMAX + 1) / 2 + 1;
  static constexpr int NRING = (MMD_PREFETCH_STEPS + 1) * RingRec<M>::NWP;
  static constexpr int NSCR = NRED > NRING ? NRED : NRING;
  // rows behind the scratch: residuals + multipliers of the projection solves (2 * NRMAX), reused by the linearisation
  // sweeps for the per-interval constants M_k, Lambda_k and the parameter-gradient accumulator
  static constexpr int NCONST = M::X * M::X + M::Z * M::X + M::Z;
  static constexpr int NROWS = 2 * NRMAX > NCONST ? 2 * NRMAX : NCONST;
  static constexpr int PER_THREAD = NSCR + NROWS;  // doubles per thread
};

#define MMD_SMEM_SETUP                                                  \
  extern __shared__ double smem[];                                      \
  const int NT = t.nslot * t.cpb;                                       \
  double* const sm_red = smem;                                          \
  double* const sm_ring = smem;                                         \
  double* const sm_c = smem + SmemPlan<M, NRMAX, UMAX>::NSCR * NT + t.tid; \
  double* const sm_l = sm_c + NRMAX * NT;                               \
  unsigned long long* const sm_bars = reinterpret_cast<unsigned long long*>(                     \
      reinterpret_cast<char*>(smem + SmemPlan<M, NRMAX, UMAX>::PER_THREAD * NT) + t.cpb * sizeof(typename M::Coef)); \
  (void)sm_red; (void)sm_ring; (void)sm_c; (void)sm_l; (void)sm_bars;

// ------------------------------------------------------------------------------------------
// dev_point: everything cached at a position.  Phase 1 = jacob_constr_blocks + chol_gram_blocks +
// log_det_sqrt_gram (mici_extensions.py:521-687, 800-820) in compressed form; phase 2 =
// grad_log_det_sqrt_gram (:1143-1146, :1173-1184) by a hand-derived second-order adjoint.
// ------------------------------------------------------------------------------------------
template <class M, int NRMAX, int RMAX, int UMAX>
MMD_PHASE void dev_point(const Dims& d, const Slots& S, const Work& W, const double* __restrict__ y, int part,
                     int which, int with_grad) {
  const Tid t = thread_id(d);
  MMD_SMEM_SETUP
  constexpr int X = M::X, V = M::V, Z = M::Z, XV = M::X * M::V;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  constexpr int UTRI = UMAX * (UMAX + 1) / 2;
  const int U = d.U, nta = t.nta, cpb = t.cpb;
  const int cur = S.cur[t.cix];
  const int sl = which ? 1 - cur : cur;
  const bool skip = !t.act || (W.status[t.cix] != 0);
  const QPtr q = qptr<M>(S.q + sl * S.s_q, d, t);
  const QPtr gq = qptr<M>(S.gradld + sl * S.s_q, d, t);
  double* Kc = tpr<XV>(S.K + sl * S.s_K, d.rmax * d.S * XV, t);
  double* Psibc = tp(S.Psib + sl * S.s_Psib, d.rmax * X * X, t);
  double* xendc = tp(S.xend + sl * S.s_xend, d.rmax * X, t);
  double* kapc = tp(S.kap + sl * S.s_xend, d.rmax * X, t);
  double* Ac = tp(S.A + sl * S.s_A, NRMAX * U, t);
  double* DinvAc = tp(S.DinvA + sl * S.s_A, NRMAX * U, t);
  double* Lc = tp(S.L + sl * S.s_L, NTRI, t);
  double* Dic = tp(S.Dinv + sl * S.s_L, NTRI, t);
  double* LCc = pc(S.LC + sl * S.s_LC, UTRI, t);
  const double* xoc = pc(W.xobs, d.T * X, t);
  double* xsc = tpr<X>(W.xs, d.rmax * d.S * X, t);
  double* Qc = tp(W.Qk, d.rmax * X * X, t);
  double* Ztc = tp(W.Zt, d.rmax * X * Z, t);

  ChainPar<M, UMAX> P;
  {
    double u[UMAX];
#pragma unroll
    for (int j = 0; j < UMAX; ++j) u[j] = (j < U) ? q.head[j * cpb] : 0.0;
    make_par<M, UMAX>(d, u, P);
  }
  const double sigy = P.sigy;
  const bool has_blk = t.slot < d.nb[part];
  Blk B;
  if (has_blk) B = get_block<M>(d, part, t.slot);
  // The model coefficients (22 doubles for FHN) and the per-interval constants of the adjoint would have to live in
  // registers across the sweep loops -- they do not fit next to the sweep's own state in 96 registers and end up
  // in local memory.  They are kept in shared memory instead (one coefficient block per chain of the tile behind
  // the per-thread regions; the per-thread constants in the rows the projection solves use for residuals and
  // multipliers) and read where they are used; MMD_SMEM_RELOAD keeps the compiler from hoisting those loads out
  // of the loops (which would put them back into registers / local memory).
  static_assert(X * X + Z * X + Z <= SmemPlan<M, NRMAX, UMAX>::NROWS, "per-interval constants must fit the rows behind the scratch");
  typename M::Coef* const sm_coef = reinterpret_cast<typename M::Coef*>(smem + SmemPlan<M, NRMAX, UMAX>::PER_THREAD * NT);
  if (t.slot == 0) sm_coef[t.cl] = P.C;
  __syncthreads();
  const typename M::Coef& CC = sm_coef[t.cl];

  double red[UTRI + 1];
#pragma unroll
  for (int i = 0; i < UTRI + 1; ++i) red[i] = 0.0;
  double xlast[X];  // state at the end of the block
  PH_T0

  if (has_blk && !skip) {
    double dx0_dv0[X * M::V0], dx0_dz[X * Z];
    M::gen_x0_jac(d.gen, P.z, dx0_dv0, dx0_dz);
    // ---------------- interval sweeps: trajectory, compressed Jacobian, interval summaries
    double x[X];
    {
      double v0[M::V0];
      ldcol<M::V0>(q.head + U * cpb, cpb, v0);
      block_start<M>(d, B, P.z, v0, xoc, cpb, x);
    }
    for (int k = 0; k < B.n; ++k) {
      const double* vp = q.body + k * d.S * V * nta;
      double* xk = xsc + k * d.S * X * nta;
      for (int tt = 0; tt < d.S; ++tt) {
        MMD_SMEM_RELOAD;
#if MMD_POINT_L2_PREFETCH > 0
        if (k * d.S + tt + MMD_POINT_L2_PREFETCH < B.n * d.S) prefetch_l2(vp + (tt + MMD_POINT_L2_PREFETCH) * V * nta);
#endif
#if MMD_POINT_L1_PREFETCH > 0
        if (k * d.S + tt + MMD_POINT_L1_PREFETCH < B.n * d.S) prefetch_l1(vp + (tt + MMD_POINT_L1_PREFETCH) * V * nta);
#endif
        strec<X>(xk + tt * X * nta, x);
        double v[V], xn[X];
        ldrec<V>(vp + tt * V * nta, v);
        M::step(CC, x, v, xn);
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = xn[i];
      }
      if (!M::OBS_LINEAR) stcol<X>(xendc + k * X * nta, nta, x);
      double Psi[X * X], Qk[X * X], kap[X];
#pragma unroll
      for (int i = 0; i < X * X; ++i) { Psi[i] = (i % (X + 1) == 0) ? 1.0 : 0.0; Qk[i] = 0.0; }
      // Z_k = sum_t Psi_{t+1} G_t accumulates in this thread's shared-memory rows (8 doubles less to keep live)
#pragma unroll
      for (int i = 0; i < X * Z; ++i) sm_c[i * NT] = 0.0;
#pragma unroll
      for (int i = 0; i < X; ++i) kap[i] = 0.0;
      double* Kk = Kc + k * d.S * XV * nta;
      for (int tt = d.S - 1; tt >= 0; --tt) {
        MMD_SMEM_RELOAD;
        // one Jacobian at a time (little live at once): K_t, then Z_k, then Psi
        double xt[X], v[V];
#if MMD_POINT_L1_PREFETCH > 0
        if (tt >= MMD_POINT_L1_PREFETCH) {
          prefetch_l1(xk + (tt - MMD_POINT_L1_PREFETCH) * X * nta);
          prefetch_l1(vp + (tt - MMD_POINT_L1_PREFETCH) * V * nta);
        }
#endif
        ldrec<X>(xk + tt * X * nta, xt);
        ldrec<V>(vp + tt * V * nta, v);
        {
          double Bm[X * V], Kt[X * V];
          M::jac_v(CC, xt, v, Bm);
          mm<X, V, X>(Psi, Bm, Kt);
          strec<XV>(Kk + tt * XV * nta, Kt);
          // Qk += Kt Kt^T ; kap = max |Kt| per row
#pragma unroll
          for (int i = 0; i < X; ++i)
#pragma unroll
            for (int j = 0; j < V; ++j) kap[i] = fmax(kap[i], fabs(Kt[i * V + j]));
#pragma unroll
          for (int i = 0; i < X; ++i)
#pragma unroll
            for (int j = 0; j < X; ++j) {
              double s2 = Qk[i * X + j];
#pragma unroll
              for (int l = 0; l < V; ++l) s2 = fma(Kt[i * V + l], Kt[j * V + l], s2);
              Qk[i * X + j] = s2;
            }
        }
        {
          double G[X * Z], Zk[X * Z];
          M::jac_z(CC, xt, v, G);
#pragma unroll
          for (int i = 0; i < X * Z; ++i) Zk[i] = sm_c[i * NT];
          mm_acc<X, Z, X>(Psi, G, Zk);
#pragma unroll
          for (int i = 0; i < X * Z; ++i) sm_c[i * NT] = Zk[i];
        }
        {
          double F[X * X], tmp[X * X];
          M::jac_x(CC, xt, v, F);
          mm<X, X, X>(Psi, F, tmp);
#pragma unroll
          for (int i = 0; i < X * X; ++i) Psi[i] = tmp[i];
        }
      }
      stcol<X * X>(Psibc + k * X * X * nta, nta, Psi);
      stcol<X * X>(Qc + k * X * X * nta, nta, Qk);
      {
        double Zk[X * Z];
#pragma unroll
        for (int i = 0; i < X * Z; ++i) Zk[i] = sm_c[i * NT];
        stcol<X * Z>(Ztc + k * X * Z * nta, nta, Zk);
      }
      stcol<X>(kapc + k * X * nta, nta, kap);
    }
#pragma unroll
    for (int i = 0; i < X; ++i) xlast[i] = x[i];
    PH(16);
    // ---------------- per-observation algebra: A_b = dc/du rows, D_b = J_v J_v^T (+ noise terms)
    double Su[X * Z], Pm[X * X], w[NRMAX * X], Dm[NTRI], Am[NRMAX * UMAX];
#pragma unroll
    for (int i = 0; i < X * Z; ++i) Su[i] = B.ini ? dx0_dz[i] : 0.0;
    if (B.ini) {
      mmt<X, X, M::V0>(dx0_dv0, dx0_dv0, Pm);
    } else {
#pragma unroll
      for (int i = 0; i < X * X; ++i) Pm[i] = 0.0;
    }
    int nrow_done = 0;
    for (int k = 0; k < B.n; ++k) {
      double Ps[X * X], Qk[X * X], Zk[X * Z], t1[X * Z], t2[X * X], t3[X * X];
      ldcol<X * X>(Psibc + k * X * X * nta, nta, Ps);
      ldcol<X * X>(Qc + k * X * X * nta, nta, Qk);
      ldcol<X * Z>(Ztc + k * X * Z * nta, nta, Zk);
      mm<X, Z, X>(Ps, Su, t1);
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Su[i] = t1[i] + Zk[i];
      mm<X, X, X>(Ps, Pm, t2);
      mmt<X, X, X>(t2, Ps, t3);
#pragma unroll
      for (int i = 0; i < X * X; ++i) Pm[i] = t3[i] + Qk[i];
      for (int r = 0; r < nrow_done; ++r) {
        double tw[X];
        mv<X, X>(Ps, &w[r * X], tw);
#pragma unroll
        for (int i = 0; i < X; ++i) w[r * X + i] = tw[i];
      }
      // new rows at this observation
      const int n_new = (k < B.ny ? 1 : 0) + ((k == B.n - 1) ? B.nx : 0);
      for (int a = 0; a < n_new; ++a) {
        double h[X];
        const bool yrow = (k < B.ny) && a == 0;
        if (yrow) {
          double xe[X];
          if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
          M::obs_grad(xe, h);
        } else {
          const int comp = a - ((k < B.ny) ? 1 : 0);
#pragma unroll
          for (int i = 0; i < X; ++i) h[i] = (i == comp) ? 1.0 : 0.0;
        }
        const int r = nrow_done;
        // w_r = P h ; D[r][j] = h . w_j ; Az[r] = h^T Su
        mv<X, X>(Pm, h, &w[r * X]);
        for (int j = 0; j <= r; ++j) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < X; ++i) s = fma(h[i], w[j * X + i], s);
          Dm[tri(r, j)] = s;
        }
        double az[Z];
        mtv<X, Z>(Su, h, az);
        for (int j = 0; j < U; ++j) {
          double s = 0.0;
          if (j < Z) {
#pragma unroll
            for (int m = 0; m < Z; ++m) s = fma(az[m], P.dzdu[m * Z + j], s);
          }
          Am[r * UMAX + j] = s;
        }
        if (d.noisy && yrow) {
          Dm[tri(r, r)] += sigy * sigy;
          if (d.noisy == 2) Am[r * UMAX + Z] = sigy * q.noise[k * nta];
        }
        nrow_done++;
      }
    }
    double ldpart;
    if (NRMAX <= 8) {
      // small blocks: compile-time bounds, everything unrolled (constant indices into Dm / Am / red: no dependent
      // local-memory addressing, independent solves interleave); identical operation order to the loops below
      const int nr1 = B.nrows;
#pragma unroll
      for (int r = 0; r < NRMAX; ++r)
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (r < nr1 && j < U) Ac[(r * U + j) * nta] = Am[r * UMAX + j];
      ldpart = chol_packed_invdiag_fixed<NRMAX>(Dm, nr1);
#pragma unroll
      for (int i = 0; i < NTRI; ++i)
        if (i < nr1 * (nr1 + 1) / 2) Lc[i * nta] = Dm[i];
#pragma unroll
      for (int c2 = 0; c2 < NRMAX; ++c2) {
        if (c2 < nr1) {
          double e[NRMAX];
#pragma unroll
          for (int r = 0; r < NRMAX; ++r) e[r] = (r == c2) ? 1.0 : 0.0;
          chol_solve_invdiag_fixed<NRMAX>(Dm, nr1, e);
#pragma unroll
          for (int r = c2; r < NRMAX; ++r)
            if (r < nr1) Dic[tri(r, c2) * nta] = e[r];
        }
      }
#pragma unroll
      for (int j = 0; j < UMAX; ++j) {
        if (j < U) {
          double col[NRMAX];
#pragma unroll
          for (int r = 0; r < NRMAX; ++r) col[r] = (r < nr1) ? Am[r * UMAX + j] : 0.0;
          chol_solve_invdiag_fixed<NRMAX>(Dm, nr1, col);
#pragma unroll
          for (int r = 0; r < NRMAX; ++r)
            if (r < nr1) DinvAc[(r * U + j) * nta] = col[r];
#pragma unroll
          for (int i = j; i < UMAX; ++i) {
            if (i < U) {
              double s = 0.0;
#pragma unroll
              for (int r = 0; r < NRMAX; ++r)
                if (r < nr1) s = fma(Am[r * UMAX + i], col[r], s);
              red[tri(i, j)] = s;
            }
          }
        }
      }
    } else {
    for (int r = 0; r < B.nrows; ++r)
      for (int j = 0; j < U; ++j) Ac[(r * U + j) * nta] = Am[r * UMAX + j];
    ldpart = chol_packed_invdiag(Dm, B.nrows);
    for (int i = 0; i < B.nrows * (B.nrows + 1) / 2; ++i) Lc[i * nta] = Dm[i];
    // explicit inverse of the block (lower triangle), column by column from the factor: what the Woodbury solves of
    // the projections use (inv_gram_block)
    for (int c2 = 0; c2 < B.nrows; ++c2) {
      double e[NRMAX];
      for (int r = 0; r < B.nrows; ++r) e[r] = (r == c2) ? 1.0 : 0.0;
      chol_solve_invdiag(Dm, B.nrows, e);
      for (int r = c2; r < B.nrows; ++r) Dic[tri(r, c2) * nta] = e[r];
    }
    // DinvA and C_b = A_b^T D_b^{-1} A_b
    for (int j = 0; j < U; ++j) {
      double col[NRMAX];
      for (int r = 0; r < B.nrows; ++r) col[r] = Am[r * UMAX + j];
      chol_solve_invdiag(Dm, B.nrows, col);
      for (int r = 0; r < B.nrows; ++r) DinvAc[(r * U + j) * nta] = col[r];
      for (int i = j; i < U; ++i) {
        double s = 0.0;
        for (int r = 0; r < B.nrows; ++r) s = fma(Am[r * UMAX + i], col[r], s);
        red[tri(i, j)] = s;
      }
    }
    }
    red[UTRI] = ldpart;
    PH(17);
  }
  block_reduce<UTRI + 1, 0>(red, sm_red, t);
  PH(18);
  double LCm[UTRI];
  for (int i = 0; i < U; ++i)
    for (int j = 0; j <= i; ++j) LCm[tri(i, j)] = red[tri(i, j)] + (i == j ? 1.0 : 0.0);
  const double ldC = chol_packed_invdiag(LCm, U);
  const double ldtot = red[UTRI] + ldC;
  if (t.slot == 0 && !skip) {
    for (int i = 0; i < U * (U + 1) / 2; ++i) LCc[i * cpb] = LCm[i];
    S.ldv[sl * S.s_ld + t.cix] = ldtot;
  }
  if (!with_grad) return;

  // =========================== phase 2: grad log det ======================================
  double gu[UMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) gu[j] = 0.0;
  if (has_blk && !skip) {
    double dx0_dv0[X * M::V0], dx0_dz[X * Z];
    M::gen_x0_jac(d.gen, P.z, dx0_dv0, dx0_dz);
    const int nr = B.nrows;
    // rows: compile-time bound and full unrolling for the small blocks (FHN), run-time bound for the large (SIR)
    constexpr bool STATIC_ROWS = NRMAX <= 8;
    constexpr int UR = STATIC_ROWS ? NRMAX : 1;
    const int NRL = STATIC_ROWS ? NRMAX : nr;
    // Block algebra of the adjoint.  Every array below is indexed by compile-time constants (rows / columns padded
    // to NRMAX / UMAX with zeros, the triangular factors with an identity) except for the interval index of `a`, so
    // the compiler keeps them in registers or fixed local slots and issues the independent loads and FMAs together
    // instead of one dependent local-memory access per operation.
    double Cinv[UMAX * UMAX];
    {
      double LCp[UTRI];
#pragma unroll
      for (int i = 0; i < UMAX; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) LCp[tri(i, j)] = (i < U) ? LCm[tri(i, j)] : (i == j ? 1.0 : 0.0);
#pragma unroll
      for (int j = 0; j < UMAX; ++j) {
        double e[UMAX];
#pragma unroll
        for (int i = 0; i < UMAX; ++i) e[i] = (i == j) ? 1.0 : 0.0;
        chol_solve_invdiag_fixed<UMAX>(LCp, UMAX, e);
#pragma unroll
        for (int i = 0; i < UMAX; ++i) Cinv[i * UMAX + j] = e[i];
      }
    }
    double DiA[NRMAX * UMAX], Om[NRMAX * UMAX], E[NRMAX * NRMAX];
#pragma unroll UR
    for (int r = 0; r < NRL; ++r)
#pragma unroll
      for (int j = 0; j < UMAX; ++j) DiA[r * UMAX + j] = (r < nr && j < U) ? DinvAc[(r * U + j) * nta] : 0.0;
#pragma unroll UR
    for (int r = 0; r < NRL; ++r)
#pragma unroll
      for (int j = 0; j < UMAX; ++j) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < UMAX; ++l) s = fma(DiA[r * UMAX + l], Cinv[l * UMAX + j], s);
        Om[r * UMAX + j] = s;
      }
    // E = D^{-1} - Om DiA^T from the stored block inverse (rows / columns beyond nr: identity, they are never used)
#pragma unroll UR
    for (int r = 0; r < NRL; ++r)
#pragma unroll UR
      for (int c2 = 0; c2 < NRL; ++c2) {
        double sv = (r < nr && c2 < nr) ? Dic[(r >= c2 ? tri(r, c2) : tri(c2, r)) * nta] : (r == c2 ? 1.0 : 0.0);
#pragma unroll
        for (int j = 0; j < UMAX; ++j) sv = fma(-Om[r * UMAX + j], DiA[c2 * UMAX + j], sv);
        E[r * NRMAX + c2] = sv;
      }
    // noise-scale terms (sigma = exp(u_Z) inferred): d/du_Z and d/dn of 1/2 <E, sigma^2 P_y> + <Om[:, Z], sigma n>
    if (d.noisy) {
#pragma unroll
      for (int k = 0; k < (RMAX < NRMAX ? RMAX : NRMAX); ++k)
        if (k < B.n) {
          double gn = 0.0;
          if (d.noisy == 2 && k < B.ny) {
            const double nk = q.noise[k * nta];
            gu[Z < UMAX ? Z : 0] += E[k * NRMAX + k] * sigy * sigy + Om[k * UMAX + (Z < UMAX ? Z : 0)] * sigy * nk;
            gn = Om[k * UMAX + (Z < UMAX ? Z : 0)] * sigy;
          }
          gq.noise[k * nta] = gn;
        }
    }
    // a[k][r] = Phi(t_kr, t_k)^T H_r^T (zero for k > kr; kr = r for observation rows, the last interval for the
    // rows of the conditioned full state): one backward recursion shared by all rows
    double a[RMAX * NRMAX * X];
    {
      double vec[NRMAX * X];
#pragma unroll
      for (int i = 0; i < NRMAX * X; ++i) vec[i] = 0.0;
      for (int k = B.n - 1; k >= 0; --k) {
        if (k < B.n - 1) {
          double Ps[X * X];
          ldcol<X * X>(Psibc + (k + 1) * X * X * nta, nta, Ps);
#pragma unroll UR
          for (int r = 0; r < NRL; ++r) {
            double tv[X];
            mtv<X, X>(Ps, &vec[r * X], tv);
#pragma unroll
            for (int i = 0; i < X; ++i) vec[r * X + i] = tv[i];
          }
        }
        if (k < B.ny) {
          double xe[X], h[X];
          if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
          M::obs_grad(xe, h);
#pragma unroll UR
          for (int r = 0; r < NRL; ++r)
            if (r == k) {
#pragma unroll
              for (int i = 0; i < X; ++i) vec[r * X + i] = h[i];
            }
        }
        if (k == B.n - 1) {
#pragma unroll UR
          for (int r = 0; r < NRL; ++r)
            if (r >= B.ny && r < nr) {
#pragma unroll
              for (int i = 0; i < X; ++i) vec[r * X + i] = (i == r - B.ny) ? 1.0 : 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < NRMAX * X; ++i) a[k * NRMAX * X + i] = vec[i];
      }
    }
    // u-directions in z-space: omz[r] = dzdu Om[r]
    double omz[NRMAX * Z];
#pragma unroll UR
    for (int r = 0; r < NRL; ++r)
#pragma unroll
      for (int m = 0; m < Z; ++m) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < Z; ++j) s = fma(P.dzdu[m * Z + j], Om[r * UMAX + j], s);
        omz[r * Z + m] = s;
      }
    // per-interval constants M_k, LamZ_k, Yb_k ; tangent recursion ; Az rows (for d2z/du2 term)
    double dprev[NRMAX * X];
#pragma unroll
    for (int i = 0; i < NRMAX * X; ++i) dprev[i] = 0.0;
    if (B.ini) {
      double Ps[X * X];
      ldcol<X * X>(Psibc, nta, Ps);
#pragma unroll UR
      for (int r = 0; r < NRL; ++r) {
        double be0[X], b0[X], t0[M::V0], t1[X], t2[X];
#pragma unroll
        for (int i = 0; i < X; ++i) {
          double s = 0.0;
#pragma unroll UR
          for (int s2 = 0; s2 < NRL; ++s2) s = fma(E[r * NRMAX + s2], a[s2 * X + i], s);
          be0[i] = s;
        }
        mtv<X, X>(Ps, be0, b0);
        mtv<X, M::V0>(dx0_dv0, b0, t0);
        mv<X, M::V0>(dx0_dv0, t0, t1);
        mv<X, Z>(dx0_dz, &omz[r * Z], t2);
#pragma unroll
        for (int i = 0; i < X; ++i) dprev[r * X + i] = t1[i] + t2[i];
      }
    }
    double Suz[X * Z], Gam[Z * UMAX];
    double gobs[M::OBS_LINEAR ? 1 : RMAX * X];
    (void)gobs;
#pragma unroll
    for (int i = 0; i < X * Z; ++i) Suz[i] = B.ini ? dx0_dz[i] : 0.0;
#pragma unroll
    for (int i = 0; i < Z * UMAX; ++i) Gam[i] = 0.0;
    double* Mkc = tp(W.Mk, d.rmax * X * X, t);
    double* LamZc = tp(W.LamZ, d.rmax * Z * X, t);
    double* Ybc = tp(W.Yb, d.rmax * X * X, t);
    for (int k = 0; k < B.n; ++k) {
      double Ak[NRMAX * X], bek[NRMAX * X];
#pragma unroll
      for (int i = 0; i < NRMAX * X; ++i) Ak[i] = a[k * NRMAX * X + i];
#pragma unroll UR
      for (int r = 0; r < NRL; ++r)
#pragma unroll
        for (int i = 0; i < X; ++i) {
          double s = 0.0;
#pragma unroll UR
          for (int s2 = 0; s2 < NRL; ++s2) s = fma(E[r * NRMAX + s2], Ak[s2 * X + i], s);
          bek[r * X + i] = s;
        }
      double Mk[X * X], Lam[Z * X], Yb[X * X];
#pragma unroll
      for (int i = 0; i < X * X; ++i) { Mk[i] = 0.0; Yb[i] = 0.0; }
#pragma unroll
      for (int i = 0; i < Z * X; ++i) Lam[i] = 0.0;
#pragma unroll UR
      for (int r = 0; r < NRL; ++r) {
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < X; ++j) {
            Mk[i * X + j] = fma(bek[r * X + i], Ak[r * X + j], Mk[i * X + j]);
            Yb[i * X + j] = fma(dprev[r * X + i], Ak[r * X + j], Yb[i * X + j]);
          }
#pragma unroll
        for (int m = 0; m < Z; ++m)
#pragma unroll
          for (int j = 0; j < X; ++j) Lam[m * X + j] = fma(omz[r * Z + m], Ak[r * X + j], Lam[m * X + j]);
      }
      stcol<X * X>(Mkc + k * X * X * nta, nta, Mk);
      stcol<Z * X>(LamZc + k * Z * X * nta, nta, Lam);
      stcol<X * X>(Ybc + k * X * X * nta, nta, Yb);
      double Ps[X * X], Qk[X * X], Zk[X * Z];
      ldcol<X * X>(Psibc + k * X * X * nta, nta, Ps);
      ldcol<X * X>(Qc + k * X * X * nta, nta, Qk);
      ldcol<X * Z>(Ztc + k * X * Z * nta, nta, Zk);
#pragma unroll UR
      for (int r = 0; r < NRL; ++r) {
        double t1[X], t2[X], t3[X];
        mv<X, X>(Ps, &dprev[r * X], t1);
        mv<X, X>(Qk, &bek[r * X], t2);
        mv<X, Z>(Zk, &omz[r * Z], t3);
#pragma unroll
        for (int i = 0; i < X; ++i) dprev[r * X + i] = t1[i] + t2[i] + t3[i];
      }
      if (!M::OBS_LINEAR && k < B.ny) {
        // tangent of the state at observation time k in the direction that weights row k: the observation
        // curvature enters the adjoint there (d H_k = Hess h(x_k) d x_k)
#pragma unroll UR
        for (int r = 0; r < NRL; ++r)
          if (r == k) {
#pragma unroll
            for (int i = 0; i < X; ++i) gobs[k * X + i] = dprev[r * X + i];
          }
      }
      double ts[X * Z];
      mm<X, Z, X>(Ps, Suz, ts);
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Suz[i] = ts[i] + Zk[i];
#pragma unroll UR
      for (int r = 0; r < NRL; ++r) {
        const int kr = (r < B.ny) ? r : B.n - 1;
        if (kr != k || r >= nr) continue;
        double az[Z];
        mtv<X, Z>(Suz, &Ak[r * X], az);  // a[kr][r] = H_r^T
#pragma unroll
        for (int m = 0; m < Z; ++m)
#pragma unroll
          for (int j = 0; j < UMAX; ++j) Gam[m * UMAX + j] = fma(az[m], Om[r * UMAX + j], Gam[m * UMAX + j]);
      }
    }
    PH(19);
    // ---------------- second-order sweeps, last interval first
    double gam[X], gz[Z];
#pragma unroll
    for (int i = 0; i < X; ++i) gam[i] = 0.0;
    // the parameter gradient accumulates in shared memory during the sweeps (rows behind M_k and Lambda_k)
    constexpr int GZ0 = X * X + Z * X;
    // contraction weights of one step: (X+V+Z)*X doubles per thread in the scratch region, thread stride odd in
    // doubles (at most 2-way bank conflicts); only when they fit (FHN: 16 + 1 <= 18)
    constexpr int NTH = (X + V + Z) * X;
    constexpr bool TH_SMEM = NTH + 1 <= SmemPlan<M, NRMAX, UMAX>::NSCR;
    double* const sm_th = smem + t.tid * (NTH + 1);
#pragma unroll
    for (int i = 0; i < Z; ++i) sm_c[(GZ0 + i) * NT] = 0.0;
    double* Ywc = tpr<X * X>(W.Yw, d.rmax * d.S * X * X, t);
    for (int k = B.n - 1; k >= 0; --k) {
      const double* vp = q.body + k * d.S * V * nta;
      const double* xk = xsc + k * d.S * X * nta;
      const double* Kk = Kc + k * d.S * XV * nta;
      double* Yk = Ywc + k * d.S * X * X * nta;
      double* gk = gq.body + k * d.S * V * nta;
      double Y[X * X];
      {
        double Mk[X * X], Lam[Z * X];
        ldcol<X * X>(Mkc + k * X * X * nta, nta, Mk);
        ldcol<Z * X>(LamZc + k * Z * X * nta, nta, Lam);
#pragma unroll
        for (int i = 0; i < X * X; ++i) sm_c[i * NT] = Mk[i];
#pragma unroll
        for (int i = 0; i < Z * X; ++i) sm_c[(X * X + i) * NT] = Lam[i];
      }
      ldcol<X * X>(Ybc + k * X * X * nta, nta, Y);
      if (!M::OBS_LINEAR && k < B.ny) {
        // curvature of the observation function: adjoint source at the observation time t_k
        double xe[X], hv[X];
        ldcol<X>(xendc + k * X * nta, nta, xe);
        M::obs_hess_vec(xe, &gobs[k * X], hv);
#pragma unroll
        for (int i = 0; i < X; ++i) gam[i] += hv[i];
      }
      PH(22);
      for (int tt = 0; tt < d.S; ++tt) {
        MMD_SMEM_RELOAD;
        strec<X * X>(Yk + tt * X * X * nta, Y);
#if MMD_POINT_L2_PREFETCH > 0
        if (tt + MMD_POINT_L2_PREFETCH < d.S) {
          prefetch_l2(xk + (tt + MMD_POINT_L2_PREFETCH) * X * nta);
          prefetch_l2(vp + (tt + MMD_POINT_L2_PREFETCH) * V * nta);
          prefetch_l2(Kk + (tt + MMD_POINT_L2_PREFETCH) * XV * nta);
        }
#endif
#if MMD_POINT_L1_PREFETCH > 0
        if (tt + MMD_POINT_L1_PREFETCH < d.S) {
          prefetch_l1(xk + (tt + MMD_POINT_L1_PREFETCH) * X * nta);
          prefetch_l1(vp + (tt + MMD_POINT_L1_PREFETCH) * V * nta);
          prefetch_l1(Kk + (tt + MMD_POINT_L1_PREFETCH) * XV * nta);
        }
#endif
        // Y <- F Y + B (K^T M) + G Lam, one Jacobian at a time
        double xt[X], v[V], Yn[X * X];
        ldrec<X>(xk + tt * X * nta, xt);
        ldrec<V>(vp + tt * V * nta, v);
        {
          double F[X * X];
          M::jac_x(CC, xt, v, F);
          mm<X, X, X>(F, Y, Yn);
        }
        {
          double Bm[X * V], Kt[X * V], KM[V * X], Mk[X * X];
          ldrec<XV>(Kk + tt * XV * nta, Kt);
#pragma unroll
          for (int i = 0; i < X * X; ++i) Mk[i] = sm_c[i * NT];
          M::jac_v(CC, xt, v, Bm);
          mtm<V, X, X>(Kt, Mk, KM);
          mm_acc<X, X, V>(Bm, KM, Yn);
        }
        {
          double G[X * Z], Lam[Z * X];
#pragma unroll
          for (int i = 0; i < Z * X; ++i) Lam[i] = sm_c[(X * X + i) * NT];
          M::jac_z(CC, xt, v, G);
          mm_acc<X, X, Z>(G, Lam, Yn);
        }
#pragma unroll
        for (int i = 0; i < X * X; ++i) Y[i] = Yn[i];
      }
      PH(20);
      double Psi[X * X];
#pragma unroll
      for (int i = 0; i < X * X; ++i) Psi[i] = (i % (X + 1) == 0) ? 1.0 : 0.0;
      for (int tt = d.S - 1; tt >= 0; --tt) {
        MMD_SMEM_RELOAD;
        // staged so that little is live at once: (1) the contraction weights Th = [Y Psi ; K^T M Psi ; Lam Psi] go to
        // this thread's shared-memory scratch, (2) the Hessian contraction reads them from there, (3) the
        // first-order terms follow one Jacobian at a time
        double xt[X], v[V], Bm[X * V], g[X + V + Z];
#if MMD_POINT_L2_PREFETCH > 0
        if (k > 0 && d.S - 1 - tt < MMD_POINT_L2_PREFETCH) {
          const int sp = (d.S - 1 - tt) - d.S;   // step of interval k - 1, relative to this interval's base
          prefetch_l2(xk + sp * X * nta);
          prefetch_l2(vp + sp * V * nta);
          prefetch_l2(Kk + sp * XV * nta);
        }
#endif
#if MMD_POINT_L1_PREFETCH > 0
        if (tt >= MMD_POINT_L1_PREFETCH) {
          prefetch_l1(xk + (tt - MMD_POINT_L1_PREFETCH) * X * nta);
          prefetch_l1(vp + (tt - MMD_POINT_L1_PREFETCH) * V * nta);
          prefetch_l1(Yk + (tt - MMD_POINT_L1_PREFETCH) * X * X * nta);
        }
#endif
        ldrec<X>(xk + tt * X * nta, xt);
        ldrec<V>(vp + tt * V * nta, v);
        M::jac_v(CC, xt, v, Bm);
        {
          double Th[NTH], MP[X * X], Kt[X * V], Yt[X * X], Mk[X * X], Lam[Z * X];
          ldrec<X * X>(Yk + tt * X * X * nta, Yt);
#pragma unroll
          for (int i = 0; i < X * X; ++i) Mk[i] = sm_c[i * NT];
#pragma unroll
          for (int i = 0; i < Z * X; ++i) Lam[i] = sm_c[(X * X + i) * NT];
          mm<X, V, X>(Psi, Bm, Kt);
          mm<X, X, X>(Yt, Psi, &Th[0]);
          mm<X, X, X>(Mk, Psi, MP);
          mtm<V, X, X>(Kt, MP, &Th[X * X]);
          mm<Z, X, X>(Lam, Psi, &Th[(X + V) * X]);
          if constexpr (TH_SMEM) {
#pragma unroll
            for (int i = 0; i < NTH; ++i) sm_th[i] = Th[i];
            MMD_SMEM_RELOAD;
            M::hess_contract(CC, xt, v, sm_th, g);
          } else {
            M::hess_contract(CC, xt, v, Th, g);
          }
        }
        {
          double gv[V];
          mtv<X, V>(Bm, gam, gv);
#pragma unroll
          for (int j = 0; j < V; ++j) gv[j] += g[X + j];
          strec<V>(gk + tt * V * nta, gv);
        }
        {
          double G[X * Z], tz[Z];
          M::jac_z(CC, xt, v, G);
          mtv<X, Z>(G, gam, tz);
#pragma unroll
          for (int m = 0; m < Z; ++m) sm_c[(GZ0 + m) * NT] += tz[m] + g[X + V + m];
        }
        {
          double F[X * X], gn[X], tmp[X * X];
          M::jac_x(CC, xt, v, F);
          mtv<X, X>(F, gam, gn);
#pragma unroll
          for (int i = 0; i < X; ++i) gam[i] = gn[i] + g[i];
          mm<X, X, X>(Psi, F, tmp);
#pragma unroll
          for (int i = 0; i < X * X; ++i) Psi[i] = tmp[i];
        }
      }
      PH(21);
    }
#pragma unroll
    for (int i = 0; i < Z; ++i) gz[i] = sm_c[(GZ0 + i) * NT];
    double gv0[M::V0];
    if (B.ini) {
      double tz[Z];
      mtv<X, M::V0>(dx0_dv0, gam, gv0);
      mtv<X, Z>(dx0_dz, gam, tz);
#pragma unroll
      for (int m = 0; m < Z; ++m) gz[m] += tz[m];
      stcol<M::V0>(gq.head + U * cpb, cpb, gv0);
    }
    double extra[Z], GamZ[Z * Z];
#pragma unroll
    for (int m = 0; m < Z; ++m)
#pragma unroll
      for (int j = 0; j < Z; ++j) GamZ[m * Z + j] = Gam[m * UMAX + j];
    M::gen_z_second(d.gen, P.u, P.z, GamZ, extra);
#pragma unroll
    for (int j = 0; j < Z; ++j) {
      double s = extra[j];
#pragma unroll
      for (int m = 0; m < Z; ++m) s = fma(P.dzdu[m * Z + j], gz[m], s);
      gu[j] += s;
    }
  }
  (void)xlast;
  // the sweeps use the scratch region thread by thread (contraction weights): every thread must have left them
  // before the reduction writes it column by column
  __syncthreads();
  block_reduce<UMAX, 0>(gu, sm_red, t);
  PH(23);
  if (t.slot == 0 && !skip)
    for (int j = 0; j < U; ++j) gq.head[j * cpb] = gu[j];
}

// ------------------------------------------------------------------------------------------
// dev_constr: c(q) at the resident position for the standalone `constr` op (mici_extensions.py:473-519);
// cout is thread-private [NRMAX]
// ------------------------------------------------------------------------------------------
template <class M, int NRMAX, int UMAX>
MMD_D void dev_constr(const Dims& d, const Slots& S, const Work& W, const double* __restrict__ y, int part,
                      double* __restrict__ cout) {
  const Tid t = thread_id(d);
  MMD_SMEM_SETUP
  constexpr int X = M::X;
  if (t.slot >= d.nb[part]) return;
  const int cur = S.cur[t.cix];
  const QPtr q = qptr<M>(S.q + cur * S.s_q, d, t);
  ChainPar<M, UMAX> P;
  {
    double u[UMAX];
#pragma unroll
    for (int j = 0; j < UMAX; ++j) u[j] = (j < d.U) ? q.head[j * t.cpb] : 0.0;
    make_par<M, UMAX>(d, u, P);
  }
  const Blk B = get_block<M>(d, part, t.slot);
  const double* xoc = pc(W.xobs, d.T * X, t);
  double x[X], v0[M::V0];
  ldcol<M::V0>(q.head + d.U * t.cpb, t.cpb, v0);
  block_start<M>(d, B, P.z, v0, xoc, t.cpb, x);
  {
    SweepArgs<M> a;
    a.C = P.C; a.sigma_y = P.sigy; a.sigma_lin = 0.0;
#pragma unroll
    for (int i = 0; i < X; ++i) a.xstart[i] = x[i];
    a.vb = q.body; a.nzb = q.noise; a.xoc = xoc; a.y = y; a.Kb = nullptr; a.alph = nullptr; a.lamtot = nullptr;
    a.crow = sm_c; a.ring = sm_ring; a.xend_out = nullptr; a.xs_out = nullptr;
    a.nta = t.nta; a.cpb = t.cpb; a.NT = NT; a.tid = t.tid; a.phase = nullptr; a.mask = 0u; a.bars = nullptr;
    constr_sweep<M, false>(d, B, a);
  }
  double* co = tp(cout, NRMAX, t);
  for (int r = 0; r < B.nrows; ++r) co[r * t.nta] = sm_c[r * NT];
}

// ------------------------------------------------------------------------------------------
// dev_project: p'' = proj(p - h (qcoef q + gradld)) with proj = I - J^T (J J^T)^{-1} J evaluated from
// the cached compressed Jacobian (normal_space_component / project_onto_cotangent_space, :983-993,
// :1243-1254) fused with the preceding h1_flow (Mici System.h1_flow; dh1_dpos :1192-1196) and,
// optionally, with the following h2_flow (:1222-1231):
//   flow 0: dst = p''
//   flow 1: dst = p'' and qw = fq * q + fp * p''                         (trial position only)
//   flow 2: dst = fq * p'' - fpm * q and qw = fq * q + fp * p''          (position and momentum)
// where (fq, fp, fpm) = (1, dt, 0) for the standard splitting and (cos dt, sin dt, sin dt) for the
// Gaussian splitting.  lin/src/dst select slots relative to cur (PSEL_*).
// ------------------------------------------------------------------------------------------

template <class M, int NRMAX, int RMAXP, int UMAX>
MMD_PHASE void dev_project(const Dims& d, const Slots& S, const Work& W, int part, int lin_sel, int src_sel,
                       int dst_sel, double h, double qcoef, FlowCoef fl, int force = -1) {
  const Tid t = thread_id(d);
  MMD_SMEM_SETUP
  constexpr int X = M::X, V = M::V, Z = M::Z, XV = M::X * M::V;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  constexpr int UTRI = UMAX * (UMAX + 1) / 2;
  const int U = d.U, nta = t.nta, cpb = t.cpb;
  const int cur = S.cur[t.cix];
  // force >= 0: the caller names the chains to work on (the deferred half kick of chains that failed later in a
  // multi-step launch); otherwise every chain without an error status
  const bool skip = !t.act || (force >= 0 ? force == 0 : (W.status[t.cix] != 0));
  const int sl = lin_sel ? 1 - cur : cur;
  auto psel = [&](int sel) -> double* {
    return sel == PSEL_WORK ? W.pw : S.p + (sel == PSEL_OTHER ? 1 - cur : cur) * S.s_q;
  };
  const QPtr q = qptr<M>(S.q + sl * S.s_q, d, t);
  const QPtr g = qptr<M>(S.gradld + sl * S.s_q, d, t);
  const QPtr ps = qptr<M>(psel(src_sel), d, t);
  const QPtr pd = qptr<M>(psel(dst_sel), d, t);
  const QPtr qw = qptr<M>(W.qw, d, t);
  const double* Kc = tpr<XV>(S.K + sl * S.s_K, d.rmax * d.S * XV, t);
  const double* Psibc = tp(S.Psib + sl * S.s_Psib, d.rmax * X * X, t);
  const double* xendc = tp(S.xend + sl * S.s_xend, d.rmax * X, t);
  const double* Ac = tp(S.A + sl * S.s_A, NRMAX * U, t);
  const double* DinvAc = tp(S.DinvA + sl * S.s_A, NRMAX * U, t);
  const double* Dic = tp(S.Dinv + sl * S.s_L, NTRI, t);
  const double* LCc = pc(S.LC + sl * S.s_LC, UTRI, t);
  double* alph = tp(W.alpha, d.rmax * X, t);
  const bool has_blk = t.slot < d.nb[part] && !skip;
  const bool kick = (h != 0.0);
  const bool write1 = kick || (src_sel != dst_sel);
  Blk B;
  if (t.slot < d.nb[part]) B = get_block<M>(d, part, t.slot);
  const int nsteps_blk = (t.slot < d.nb[part]) ? B.n * d.S : 0;
  (void)nsteps_blk;
  double pu[UMAX], sres[UMAX], pv0[M::V0];
  double dx0_dv0[X * M::V0], dx0_dz[X * Z];
  double sig = 0.0;
#pragma unroll
  for (int j = 0; j < UMAX; ++j) pu[j] = 0.0;
  if (has_blk) {
    double u[UMAX], z[Z], dzdu[Z * Z];
#pragma unroll
    for (int j = 0; j < UMAX; ++j) u[j] = (j < U) ? q.head[j * cpb] : 0.0;
    M::gen_z(d.gen, u, z, dzdu);
    M::gen_x0_jac(d.gen, z, dx0_dv0, dx0_dz);
    sig = sigma_of<M>(d, u);
    // head of the kicked momentum (u and v_0 components), kept in registers
#pragma unroll
    for (int j = 0; j < UMAX; ++j)
      if (j < U) {
        double pv = ps.head[j * cpb];
        if (kick) pv -= h * (qcoef * u[j] + g.head[j * cpb]);
        pu[j] = pv;
      }
#pragma unroll
    for (int j = 0; j < M::V0; ++j) {
      double pv = ps.head[(U + j) * cpb];
      if (kick) pv -= h * (qcoef * q.head[(U + j) * cpb] + g.head[(U + j) * cpb]);
      pv0[j] = pv;
    }
    double m[X];
    if (B.ini) {
      mv<X, M::V0>(dx0_dv0, pv0, m);
    } else {
#pragma unroll
      for (int i = 0; i < X; ++i) m[i] = 0.0;
    }
    // A_b p_u, three rows of A per batch of loads (see inv_gram_block)
    if (NRMAX <= 8) {
#pragma unroll
      for (int r0 = 0; r0 < NRMAX; r0 += 3) {
        if (r0 < B.nrows) {
          double av[3][UMAX];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (r0 + c < NRMAX) {
#pragma unroll
              for (int j = 0; j < UMAX; ++j) av[c][j] = ldg_vol(Ac + ((r0 + c) * U + (j < U ? j : U - 1)) * nta);
            }
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (r0 + c < NRMAX && r0 + c < B.nrows) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < UMAX; ++j)
                if (j < U) s = fma(av[c][j], pu[j], s);
              sm_c[(r0 + c) * NT] = s;
            }
        }
      }
    } else {
      for (int rr = 0; rr < B.nrows; ++rr) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (j < U) s = fma(Ac[(rr * U + j) * nta], pu[j], s);
        sm_c[rr * NT] = s;
      }
    }
#if MMD_FACTOR_PREFETCH > 1
    prefetch_col_l2(Dic, NTRI, nta);
    prefetch_col_l2(DinvAc, NRMAX * U, nta);
#endif
    // pass 1: r = J p'  (lmult_by_jacob_constr :822-877); p' written to dst when it differs from src
    for (int k = 0; k < B.n; ++k) {
      double sk[X];
#pragma unroll
      for (int i = 0; i < X; ++i) sk[i] = 0.0;
#pragma unroll 4
      for (int tt = 0; tt < d.S; ++tt) {
        const int so = (k * d.S + tt) * nta;
        double Kt[XV], pv[V];
#if MMD_POINTWISE_L2_PREFETCH > 0
        if (k * d.S + tt + MMD_POINTWISE_L2_PREFETCH < nsteps_blk) {   // pull a later step's records into L2
          const int sp = so + MMD_POINTWISE_L2_PREFETCH * nta;
          prefetch_l2(Kc + sp * XV);
          prefetch_l2(ps.body + sp * V);
          if (kick) { prefetch_l2(q.body + sp * V); prefetch_l2(g.body + sp * V); }
        }
#endif
        ldrec<XV>(Kc + so * XV, Kt);
        ldrec<V>(ps.body + so * V, pv);
        if (kick) {
          double qv[V], gv[V];
          ldrec<V>(q.body + so * V, qv);
          ldrec<V>(g.body + so * V, gv);
#pragma unroll
          for (int j = 0; j < V; ++j) pv[j] -= h * (qcoef * qv[j] + gv[j]);
        }
        if (write1) strec<V>(pd.body + so * V, pv);
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < V; ++j) sk[i] = fma(Kt[i * V + j], pv[j], sk[i]);
      }
      double Ps[X * X], t1[X];
      ldcol<X * X>(Psibc + k * X * X * nta, nta, Ps);
      mv<X, X>(Ps, m, t1);
#pragma unroll
      for (int i = 0; i < X; ++i) m[i] = t1[i] + sk[i];
      if (k < B.ny) {
        double dh[X], xe[X];
        if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
        M::obs_grad(xe, dh);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < X; ++i) s = fma(dh[i], m[i], s);
        if (d.noisy) {
          double pn = ps.noise[k * nta];
          if (kick) pn -= h * (qcoef * q.noise[k * nta] + g.noise[k * nta]);
          if (write1) pd.noise[k * nta] = pn;
          s = fma(sig, pn, s);
        }
        sm_c[k * NT] += s;
      }
      if (k == B.n - 1 && B.nx > 0) {
#pragma unroll
        for (int i = 0; i < X; ++i) sm_c[(B.ny + i) * NT] += m[i];
      }
    }
  }
  double rr[NRMAX];
#pragma unroll
  for (int i = 0; i < NRMAX; ++i) rr[i] = has_blk ? sm_c[i * NT] : 0.0;
  inv_gram_block<M, NRMAX, UMAX>(d, B, has_blk, Dic, DinvAc, LCc, rr, sres, nullptr, sm_red, t);
  if (has_blk) {
    // pass 2: p'' = p' - J^T lambda  (rmult_by_jacob_constr :879-913), fused h2_flow
    double a0[X];
#pragma unroll
    for (int i = 0; i < NRMAX; ++i) sm_c[i * NT] = rr[i];
    alpha_block<M, NRMAX, RMAXP>(B, rr, Psibc, xendc, nta, alph, a0);
    for (int k = 0; k < B.n; ++k) {
      double al[X];
      ldcol<X>(alph + k * X * nta, nta, al);
#pragma unroll 4
      for (int tt = 0; tt < d.S; ++tt) {
        const int so = (k * d.S + tt) * nta;
        double Kt[XV], pv[V];
#if MMD_POINTWISE_L2_PREFETCH > 0
        if (k * d.S + tt + MMD_POINTWISE_L2_PREFETCH < nsteps_blk) {
          const int sp = so + MMD_POINTWISE_L2_PREFETCH * nta;
          prefetch_l2(Kc + sp * XV);
          prefetch_l2((write1 ? pd.body : ps.body) + sp * V);
          if (fl.mode) prefetch_l2(q.body + sp * V);
        }
#endif
        ldrec<XV>(Kc + so * XV, Kt);
        ldrec<V>((write1 ? pd.body : ps.body) + so * V, pv);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int i = 0; i < X; ++i) pv[j] = fma(-Kt[i * V + j], al[i], pv[j]);
        if (fl.mode) {
          double qv[V], qo[V];
          ldrec<V>(q.body + so * V, qv);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            qo[j] = fl.fq * qv[j] + fl.fp * pv[j];
            if (fl.mode == 2) pv[j] = fl.fq * pv[j] - fl.fpm * qv[j];
          }
          strec<V>(qw.body + so * V, qo);
        }
        strec<V>(pd.body + so * V, pv);
      }
      if (d.noisy) {
        double pv = write1 ? pd.noise[k * nta] : ps.noise[k * nta];
        if (k < B.ny) pv = fma(-sig, sm_c[k * NT], pv);
        if (fl.mode) {
          const double qv = q.noise[k * nta];
          qw.noise[k * nta] = fl.fq * qv + fl.fp * pv;
          if (fl.mode == 2) pv = fl.fq * pv - fl.fpm * qv;
        }
        pd.noise[k * nta] = pv;
      }
    }
    if (B.ini) {
      double t0[M::V0];
      mtv<X, M::V0>(dx0_dv0, a0, t0);
#pragma unroll
      for (int j = 0; j < M::V0; ++j) {
        double pv = pv0[j] - t0[j];
        if (fl.mode) {
          const double qv = q.head[(U + j) * cpb];
          qw.head[(U + j) * cpb] = fl.fq * qv + fl.fp * pv;
          if (fl.mode == 2) pv = fl.fq * pv - fl.fpm * qv;
        }
        pd.head[(U + j) * cpb] = pv;
      }
#pragma unroll
      for (int j = 0; j < UMAX; ++j)
        if (j < U) {
          double pv = pu[j] - sres[j];
          if (fl.mode) {
            const double qv = q.head[j * cpb];
            qw.head[j * cpb] = fl.fq * qv + fl.fp * pv;
            if (fl.mode == 2) pv = fl.fq * pv - fl.fpm * qv;
          }
          pd.head[j * cpb] = pv;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Newton iteration body (newton_projection :1065-1135, lu_jacob_product_blocks :689-763,
// lmult_by_inv_jacob_product :944-981): linearise the constraint at the CURRENT iterate, form the
// non-symmetric block products D_b = J_v(q) J_v(q_lin)^T by the cross-covariance recursion
//   P_k = Psib_k(q) P_{k-1} Psib_k(q_lin)^T + sum_t K_t(q) K_t(q_lin)^T ,
// LU-factorise them with partial pivoting, and solve
//   (J(q) J(q_lin)^T) lam = c   via   C = I + sum_b A_b(q_lin)^T D_b^{-1} A_b(q)   (Woodbury).
// On entry the forward sweep has stored the trajectory in W.xs and c in `rr`; on return rr = lam_b.
// ------------------------------------------------------------------------------------------
template <class M, int NRMAX, int RMAX, int UMAX>
MMD_D void newton_solve_block(const Dims& d, const Blk& B, bool work, const ChainPar<M, UMAX>& P, double sig_lin,
                              const double* dx0_dv0_lin, const QPtr& qw, const double* __restrict__ Kc,
                              const double* __restrict__ Psibc, const double* __restrict__ xendc,
                              const double* __restrict__ Ac, const double* __restrict__ alph, const double* lamtot,
                              const Work& W, double* rr, double* s_out, double* extra_max, double* sm_red,
                              const Tid& t, int NT) {
  constexpr int X = M::X, V = M::V, Z = M::Z, XV = M::X * M::V;
  const int U = d.U, nta = t.nta;
  double Cb[UMAX * UMAX], gb[UMAX + 1];
#pragma unroll
  for (int i = 0; i < UMAX * UMAX; ++i) Cb[i] = 0.0;
#pragma unroll
  for (int i = 0; i < UMAX; ++i) gb[i] = 0.0;
  gb[UMAX] = extra_max ? *extra_max : 0.0;
  double Dm[NRMAX * NRMAX], DiA[NRMAX * UMAX], Am[NRMAX * UMAX];
  int piv[NRMAX];
  const int nr = work ? B.nrows : 0;
  if (work) {
    const double* xsc = tpr<X>(W.xs, d.rmax * d.S * X, t);
    double* Pcc = tp(W.Mk, d.rmax * X * X, t);   // Psib at the current iterate
    double* Qc = tp(W.Qk, d.rmax * X * X, t);    // cross sums  sum_t K_t(q) K_t(q_lin)^T
    double* Ztc = tp(W.Zt, d.rmax * X * Z, t);
    double dx0_dv0[X * M::V0], dx0_dz[X * Z];
    M::gen_x0_jac(d.gen, P.z, dx0_dv0, dx0_dz);
    // ---- backward sweeps per interval at the current iterate
    for (int k = 0; k < B.n; ++k) {
      double al[X];
      ldcol<X>(alph + k * X * nta, nta, al);
      double Psi[X * X], Qk[X * X], Zk[X * Z];
#pragma unroll
      for (int i = 0; i < X * X; ++i) { Psi[i] = (i % (X + 1) == 0) ? 1.0 : 0.0; Qk[i] = 0.0; }
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Zk[i] = 0.0;
      for (int tt = d.S - 1; tt >= 0; --tt) {
        const int so = (k * d.S + tt) * nta;
        double xt[X], v[V], Kp[XV], F[X * X], Bm[X * V], G[X * Z], Kt[X * V], tmp[X * X];
        ldrec<X>(xsc + so * X, xt);
        ldrec<V>(qw.body + so * V, v);
        ldrec<XV>(Kc + so * XV, Kp);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int i = 0; i < X; ++i) v[j] = fma(-Kp[i * V + j], al[i], v[j]);
        M::jac_x(P.C, xt, v, F);
        M::jac_v(P.C, xt, v, Bm);
        M::jac_z(P.C, xt, v, G);
        mm<X, V, X>(Psi, Bm, Kt);
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < X; ++j) {
            double sv = Qk[i * X + j];
#pragma unroll
            for (int l = 0; l < V; ++l) sv = fma(Kt[i * V + l], Kp[j * V + l], sv);
            Qk[i * X + j] = sv;
          }
        mm_acc<X, Z, X>(Psi, G, Zk);
        mm<X, X, X>(Psi, F, tmp);
#pragma unroll
        for (int i = 0; i < X * X; ++i) Psi[i] = tmp[i];
      }
      stcol<X * X>(Pcc + k * X * X * nta, nta, Psi);
      stcol<X * X>(Qc + k * X * X * nta, nta, Qk);
      stcol<X * Z>(Ztc + k * X * Z * nta, nta, Zk);
    }
    // ---- per-observation algebra: A_b(q) rows and the full (non-symmetric) D_b
    double Su[X * Z], Pm[X * X], wc[NRMAX * X], wp[NRMAX * X];
#pragma unroll
    for (int i = 0; i < X * Z; ++i) Su[i] = B.ini ? dx0_dz[i] : 0.0;
    if (B.ini) {
      mmt<X, X, M::V0>(dx0_dv0, dx0_dv0_lin, Pm);
    } else {
#pragma unroll
      for (int i = 0; i < X * X; ++i) Pm[i] = 0.0;
    }
    int nrow_done = 0;
    for (int k = 0; k < B.n; ++k) {
      double Pc[X * X], Pp[X * X], Qk[X * X], Zk[X * Z], t1[X * Z], t2[X * X], t3[X * X];
      ldcol<X * X>(Pcc + k * X * X * nta, nta, Pc);
      ldcol<X * X>(Psibc + k * X * X * nta, nta, Pp);
      ldcol<X * X>(Qc + k * X * X * nta, nta, Qk);
      ldcol<X * Z>(Ztc + k * X * Z * nta, nta, Zk);
      mm<X, Z, X>(Pc, Su, t1);
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Su[i] = t1[i] + Zk[i];
      mm<X, X, X>(Pc, Pm, t2);
      mmt<X, X, X>(t2, Pp, t3);
#pragma unroll
      for (int i = 0; i < X * X; ++i) Pm[i] = t3[i] + Qk[i];
      for (int r = 0; r < nrow_done; ++r) {
        double tw[X];
        mv<X, X>(Pc, &wc[r * X], tw);       // Phi_c(k, k_r) P_{k_r} h_r(lin)
#pragma unroll
        for (int i = 0; i < X; ++i) wc[r * X + i] = tw[i];
        mv<X, X>(Pp, &wp[r * X], tw);       // Phi_lin(k, k_r) P_{k_r}^T h_r(cur)
#pragma unroll
        for (int i = 0; i < X; ++i) wp[r * X + i] = tw[i];
      }
      const int n_new = (k < B.ny ? 1 : 0) + ((k == B.n - 1) ? B.nx : 0);
      const int first_new = nrow_done;
      for (int a = 0; a < n_new; ++a) {
        double hc[X], hp[X];
        const bool yrow = (k < B.ny) && a == 0;
        if (yrow) {
          double xe[X];
          // current iterate: state at the end of the interval = start of the next one (or recomputed)
          if (!M::OBS_LINEAR) ldcol<X>(tp(W.Yb, d.rmax * X * X, t) + k * X * nta, nta, xe);
          M::obs_grad(xe, hc);
          if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
          M::obs_grad(xe, hp);
        } else {
          const int comp = a - ((k < B.ny) ? 1 : 0);
#pragma unroll
          for (int i = 0; i < X; ++i) hc[i] = hp[i] = (i == comp) ? 1.0 : 0.0;
        }
        const int r = nrow_done;
        mv<X, X>(Pm, hp, &wc[r * X]);       // P_k h_r(lin)
        mtv<X, X>(Pm, hc, &wp[r * X]);      // P_k^T h_r(cur)
        for (int j = 0; j < first_new; ++j) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int i = 0; i < X; ++i) {
            s1 = fma(hc[i], wc[j * X + i], s1);   // D[r][j], k_r > k_j
            s2 = fma(wp[j * X + i], hp[i], s2);   // D[j][r]
          }
          Dm[r * NRMAX + j] = s1;
          Dm[j * NRMAX + r] = s2;
        }
        double az[Z];
        mtv<X, Z>(Su, hc, az);
        for (int j = 0; j < U; ++j) {
          double sv = 0.0;
          if (j < Z) {
#pragma unroll
            for (int m = 0; m < Z; ++m) sv = fma(az[m], P.dzdu[m * Z + j], sv);
          }
          Am[r * UMAX + j] = sv;
        }
        if (d.noisy == 2 && yrow) {
          const double nk = qw.noise[k * nta] - sig_lin * lamtot[k * NT];
          Am[r * UMAX + Z] = P.sigy * nk;
        }
        nrow_done++;
      }
      // rows created at the same observation: D[r][j] = h_r(cur) . P_k h_j(lin)
      for (int r = first_new; r < nrow_done; ++r)
        for (int j = first_new; j < nrow_done; ++j) {
          double hc[X];
          const bool yr = (k < B.ny) && r == first_new;
          if (yr) {
            double xe[X];
            if (!M::OBS_LINEAR) ldcol<X>(tp(W.Yb, d.rmax * X * X, t) + k * X * nta, nta, xe);
            M::obs_grad(xe, hc);
          } else {
            const int comp = (r - first_new) - ((k < B.ny) ? 1 : 0);
#pragma unroll
            for (int i = 0; i < X; ++i) hc[i] = (i == comp) ? 1.0 : 0.0;
          }
          double sv = 0.0;
#pragma unroll
          for (int i = 0; i < X; ++i) sv = fma(hc[i], wc[j * X + i], sv);
          if (d.noisy && yr && j == r) sv += P.sigy * sig_lin;
          Dm[r * NRMAX + j] = sv;
        }
    }
    lu_factor<NRMAX>(Dm, piv, nr);
    lu_solve<NRMAX>(Dm, piv, nr, rr);       // t_b = D_b^{-1} c_b
    for (int j = 0; j < U; ++j) {
      double col[NRMAX];
      for (int r = 0; r < nr; ++r) col[r] = Am[r * UMAX + j];
      lu_solve<NRMAX>(Dm, piv, nr, col);
      for (int r = 0; r < nr; ++r) DiA[r * UMAX + j] = col[r];
    }
    for (int r = 0; r < nr; ++r) {
      for (int i = 0; i < U; ++i) {
        const double ap = Ac[(r * U + i) * nta];        // A_b(q_lin)
        gb[i] = fma(ap, rr[r], gb[i]);
        for (int j = 0; j < U; ++j) Cb[i * UMAX + j] = fma(ap, DiA[r * UMAX + j], Cb[i * UMAX + j]);
      }
    }
  }
  // cross-block sums: C (UMAX^2 values, in chunks that fit the reduction scratch) and [g | max |c|]
  constexpr int CH = 8;
#pragma unroll
  for (int c0 = 0; c0 < UMAX * UMAX; c0 += CH) {
    double chunk[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) chunk[i] = (c0 + i < UMAX * UMAX) ? Cb[c0 + i] : 0.0;
    block_reduce<CH, 0>(chunk, sm_red, t);
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (c0 + i < UMAX * UMAX) Cb[c0 + i] = chunk[i];
  }
  block_reduce<UMAX, 1>(gb, sm_red, t);
  if (extra_max) *extra_max = gb[UMAX];
  for (int i = 0; i < U; ++i) Cb[i * UMAX + i] += 1.0;   // M_0 = I
  int pivC[UMAX];
  lu_factor<UMAX>(Cb, pivC, U);
  lu_solve<UMAX>(Cb, pivC, U, gb);
#pragma unroll
  for (int j = 0; j < UMAX; ++j) s_out[j] = gb[j];
  if (work) {
    // lam_b = D_b^{-1} (c_b - A_b(q) s) = t_b - (D_b^{-1} A_b(q)) s
    for (int r = 0; r < nr; ++r) {
      double tv = rr[r];
      for (int j = 0; j < U; ++j) tv = fma(-DiA[r * UMAX + j], gb[j], tv);
      rr[r] = tv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// dev_qn: on-device symmetric quasi-Newton projection loop (quasi_newton_projection :1009-1063 and
// its host wrapper :1323-1402) for a tile of chains, masked per chain, no host round trips.
//   iterate:  c = constr(q) ; err = |c|_inf ; lam = G_lin^{-1} c ; q -= J_lin^T lam
// with q = qw - J_lin^T lam_tot never materialised inside the loop.
// mode 0 (forward):  linearisation = cur; on convergence q(other) = q_new, p(other) = pw - mom_coef * mu
// mode 1 (reverse):  linearisation = other; compare q_back with q(cur) -> revd, no writes (Mici reverse check)
// ------------------------------------------------------------------------------------------
// One pointwise pass of the projection solve for this thread's block: q = q_w - J_lin^T lam_tot (mode 0: written to
// the other slot together with the momentum update; mode 1: reverse check, only the distance to the reference
// position).  WITH_INC also returns |J_lin^T (last increment)|_inf, the position-change norm of the reference's
// convergence test (:1047-1055).
template <class M, int NRMAX, int UMAX, bool WITH_INC>
MMD_D double qn_pass(const Dims& d, const Slots& S, const Work& W, const Blk& B, const Tid& t, int mode, int cur, int sl,
                     double mom_coef, double sig_lin, const double* dx0_dv0, const double* __restrict__ alph,
                     const double* __restrict__ alphi, const double* a0tot, const double* a0inc, const double* u0,
                     const double* stot, const double* sres, const double* sm_l, const double* sm_c, int NT,
                     double* final_norm) {
  constexpr int X = M::X, V = M::V, XV = M::X * M::V;
  const int U = d.U, nta = t.nta, cpb = t.cpb;
  const double* Kc = tpr<XV>(S.K + sl * S.s_K, d.rmax * d.S * XV, t);
  const QPtr qw = qptr<M>(W.qw, d, t);
  const QPtr qout = qptr<M>(S.q + (1 - cur) * S.s_q, d, t);
  const QPtr pout = qptr<M>(S.p + (1 - cur) * S.s_q, d, t);
  const QPtr pin = qptr<M>(W.pw, d, t);
  const QPtr qref = qptr<M>(S.q + cur * S.s_q, d, t);
  double nm = 0.0, rv = 0.0;
  for (int k = 0; k < B.n; ++k) {
    double al[X], ai[X];
    ldcol<X>(alph + k * X * nta, nta, al);
    if (WITH_INC) ldcol<X>(alphi + k * X * nta, nta, ai);
#pragma unroll 4
    for (int tt = 0; tt < d.S; ++tt) {
      const int so = (k * d.S + tt) * nta;
      double Kt[XV], qv[V], ov[V], pv[V];
#if MMD_POINTWISE_L2_PREFETCH > 0
      if (k * d.S + tt + MMD_POINTWISE_L2_PREFETCH < B.n * d.S) {
        const int sp = so + MMD_POINTWISE_L2_PREFETCH * nta;
        prefetch_l2(Kc + sp * XV);
        prefetch_l2(qw.body + sp * V);
        prefetch_l2((mode == 0 ? pin.body : qref.body) + sp * V);
      }
#endif
      ldrec<XV>(Kc + so * XV, Kt);
      ldrec<V>(qw.body + so * V, qv);
      ldrec<V>((mode == 0 ? pin.body : qref.body) + so * V, ov);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        double mu = 0.0, inc = 0.0;
#pragma unroll
        for (int i = 0; i < X; ++i) {
          mu = fma(Kt[i * V + j], al[i], mu);
          if (WITH_INC) inc = fma(Kt[i * V + j], ai[i], inc);
        }
        if (WITH_INC) nm = fmax(nm, fabs(inc));
        qv[j] -= mu;
        if (mode == 0) pv[j] = fma(-mom_coef, mu, ov[j]);
        else rv = fmax(rv, fabs(qv[j] - ov[j]));
      }
      if (mode == 0) {
        strec<V>(qout.body + so * V, qv);
        strec<V>(pout.body + so * V, pv);
      }
    }
    if (d.noisy) {
      double mu = 0.0;
      if (k < B.ny) {
        mu = sig_lin * sm_l[k * NT];
        if (WITH_INC) nm = fmax(nm, fabs(sig_lin * sm_c[k * NT]));
      }
      const double qn = qw.noise[k * nta] - mu;
      if (mode == 0) {
        qout.noise[k * nta] = qn;
        pout.noise[k * nta] = fma(-mom_coef, mu, pin.noise[k * nta]);
      } else {
        rv = fmax(rv, fabs(qn - qref.noise[k * nta]));
      }
    }
  }
  if (B.ini) {
    double t0[M::V0], ti[M::V0];
    mtv<X, M::V0>(dx0_dv0, a0tot, t0);
    mtv<X, M::V0>(dx0_dv0, a0inc, ti);
#pragma unroll
    for (int j = 0; j < M::V0; ++j) {
      if (WITH_INC) nm = fmax(nm, fabs(ti[j]));
      const int row = (U + j) * cpb;
      const double qn = qw.head[row] - t0[j];
      if (mode == 0) {
        qout.head[row] = qn;
        pout.head[row] = fma(-mom_coef, t0[j], pin.head[row]);
      } else {
        rv = fmax(rv, fabs(qn - qref.head[row]));
      }
    }
#pragma unroll
    for (int j = 0; j < UMAX; ++j)
      if (j < U) {
        if (WITH_INC) nm = fmax(nm, fabs(sres[j]));
        const double qn = u0[j] - stot[j];
        if (mode == 0) {
          qout.head[j * cpb] = qn;
          pout.head[j * cpb] = fma(-mom_coef, stot[j], pin.head[j * cpb]);
        } else {
          rv = fmax(rv, fabs(qn - qref.head[j * cpb]));
        }
      }
  }
  *final_norm = rv;
  return nm;
}

template <class M, int NRMAX, int RMAXP, int UMAX, bool NEWTON>
MMD_PHASE void dev_qn(const Dims& d, const Slots& S, const Work& W, const double* __restrict__ y, int part, int mode,
                  double mom_coef, double ctol, double ptol, double dtol, int max_iters) {
  const Tid t = thread_id(d);
  MMD_SMEM_SETUP
  constexpr int X = M::X, Z = M::Z, XV = M::X * M::V;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  constexpr int UTRI = UMAX * (UMAX + 1) / 2;
  const int U = d.U, nta = t.nta, cpb = t.cpb;
  const int cur = S.cur[t.cix];
  const int sl = mode ? 1 - cur : cur;  // linearisation used: forward -> cur (prev point), reverse -> new point
  const double* Kc = tpr<XV>(S.K + sl * S.s_K, d.rmax * d.S * XV, t);
  const double* Psibc = tp(S.Psib + sl * S.s_Psib, d.rmax * X * X, t);
  const double* xendc = tp(S.xend + sl * S.s_xend, d.rmax * X, t);
  const double* kapc = tp(S.kap + sl * S.s_xend, d.rmax * X, t);
  const double* Ac = tp(S.A + sl * S.s_A, NRMAX * U, t);
  const double* DinvAc = tp(S.DinvA + sl * S.s_A, NRMAX * U, t);
  const double* Dic = tp(S.Dinv + sl * S.s_L, NTRI, t);
  const double* LCc = pc(S.LC + sl * S.s_LC, UTRI, t);
  const QPtr qw = qptr<M>(W.qw, d, t);
  const QPtr qlin = qptr<M>(S.q + sl * S.s_q, d, t);
  const double* xoc = pc(W.xobs, d.T * X, t);
  double* alph = tp(W.alpha, d.rmax * X, t);
  double* alphi = tp(W.alphi, d.rmax * X, t);
  const bool in_blk = t.slot < d.nb[part];
  Blk B;
  if (in_blk) B = get_block<M>(d, part, t.slot);
  bool done = !t.act || (W.status[t.cix] != 0);
  bool pend = false;  // converged through the bound: iterate not written yet
  int st = 0, it = 0;
  double u0[UMAX], stot[UMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) {
    u0[j] = (j < U) ? qw.head[j * cpb] : 0.0;
    stot[j] = 0.0;
  }
  // derivatives of generate_x_0 and the noise scale at the LINEARISATION point (J_lin's v_0 / n columns)
  double dx0_dv0[X * M::V0], sig_lin = 0.0;
  {
    double u[UMAX], z[Z], dzdu[Z * Z], dx0_dz[X * Z];
#pragma unroll
    for (int j = 0; j < UMAX; ++j) u[j] = (j < U) ? qlin.head[j * cpb] : 0.0;
    M::gen_z(d.gen, u, z, dzdu);
    M::gen_x0_jac(d.gen, z, dx0_dv0, dx0_dz);
    sig_lin = sigma_of<M>(d, u);
  }
  double a0tot[X];
#pragma unroll
  for (int i = 0; i < X; ++i) a0tot[i] = 0.0;
  if (in_blk) {
    for (int k = 0; k < B.n; ++k) {
#pragma unroll
      for (int i = 0; i < X; ++i) alph[(k * X + i) * nta] = 0.0;
    }
    for (int r = 0; r < NRMAX; ++r) sm_l[r * NT] = 0.0;
  }
  double final_norm = 0.0;
  ChainPar<M, UMAX> Pkeep;
  PH_T0
#if defined(MMD_NO_FIRST_PASS_SKIP)
  bool first_pass = false;
#else
  bool first_pass = true;
#endif
  PHX_T0
  while (true) {
    const bool work = in_blk && !done;
    double sres[UMAX], err = 0.0;
    PHX_RESET
    PH_ADD(30, work ? 1 : 0);
    const unsigned wmask = __ballot_sync(0xffffffffu, work);   // (all threads of the CTA are converged here)
    if (work) {
      ChainPar<M, UMAX> P;
      {
        double u[UMAX];
#pragma unroll
        for (int j = 0; j < UMAX; ++j) u[j] = u0[j] - stot[j];
        make_par<M, UMAX>(d, u, P);
      }
      double x[X], v0[M::V0];
      ldcol<M::V0>(qw.head + U * cpb, cpb, v0);
      if (B.ini) {
        double t0[M::V0];
        mtv<X, M::V0>(dx0_dv0, a0tot, t0);
#pragma unroll
        for (int j = 0; j < M::V0; ++j) v0[j] -= t0[j];
      }
      block_start<M>(d, B, P.z, v0, xoc, cpb, x);
      PHX(24);
#if MMD_FACTOR_PREFETCH > 0
      prefetch_col_l2(Dic, NTRI, nta);
      prefetch_col_l2(DinvAc, NRMAX * U, nta);
      prefetch_col_l2(Psibc, B.n * X * X, nta);
      prefetch_col_l2(kapc, B.n * X, nta);
#endif
      {
        SweepArgs<M> a;
        a.C = P.C; a.sigma_y = P.sigy; a.sigma_lin = sig_lin;
#pragma unroll
        for (int i = 0; i < X; ++i) a.xstart[i] = x[i];
        a.vb = qw.body; a.nzb = qw.noise; a.xoc = xoc; a.y = y; a.Kb = Kc; a.alph = alph; a.lamtot = sm_l;
        a.crow = sm_c; a.ring = sm_ring; a.xend_out = nullptr;
        a.xs_out = NEWTON ? tpr<X>(W.xs, d.rmax * d.S * X, t) : nullptr;
        if (NEWTON && !M::OBS_LINEAR) a.xend_out = tp(W.Yb, d.rmax * X * X, t);
        a.nta = nta; a.cpb = cpb; a.NT = NT; a.tid = t.tid; a.phase = W.phase; a.mask = wmask; a.bars = sm_bars;
        // first pass of the loop: lam_tot = 0, the iterate is q_w itself -- no need to stream K (the products with
        // alpha = 0 leave v bit-for-bit unchanged)
        if (first_pass) constr_sweep<M, false, NEWTON>(d, B, a);
        else constr_sweep<M, true, NEWTON>(d, B, a);
      }
      if (NEWTON) Pkeep = P;
      PHX(25);
    }
    double rr[NRMAX];
    {
      double e = 0.0;
#pragma unroll
      for (int r = 0; r < NRMAX; ++r) {
        rr[r] = (work && r < B.nrows) ? sm_c[r * NT] : 0.0;
        const double a = fabs(rr[r]);
        e = (a > e || a != a) ? a : e;
      }
      err = e;
    }
    // the prefetch ring of the sweep and the reduction scratch share shared memory: every thread must have
    // left its sweep before any thread starts the cross-block reduction.  The same barrier ends the loop once
    // every chain of the tile is done (their threads skipped the sweep).
    if (__syncthreads_and(done ? 1 : 0)) break;
    PHX(26);
    PH(9);
    PH_ADD(12, 1);
    if (NEWTON)
      newton_solve_block<M, NRMAX, RMAXP, UMAX>(d, B, work, Pkeep, sig_lin, dx0_dv0, qw, Kc, Psibc, xendc, Ac, alph,
                                                sm_l, W, rr, sres, &err, sm_red, t, NT);
    else
      // (no trailing barrier: the next write to the scratch comes after the __syncthreads_or below)
      inv_gram_block<M, NRMAX, UMAX, false>(d, B, work, Dic, DinvAc, LCc, rr, sres, &err, sm_red, t, W.phase);
    PHX(27);
    // Convergence test of the reference (:1047-1055): |c| < constraint_tol AND |delta_q|_inf < position_tol for THIS
    // iteration's update delta_q = J_lin^T (increment of the multipliers).  The exact norm needs a pass over K; a
    // cheap upper bound (per-interval row maxima kap of K from the linearisation, exact for the head / noise
    // entries) decides it whenever the bound is already below the tolerance -- the usual case, since |c| < 1e-9
    // makes the update tiny.  Only an undecided chain pays for the exact pass; the iterate of a chain accepted
    // through the bound is written once, after the loop, together with the other chains of the tile.
    const bool check = !done && (err < ctol);
    double nrm[1];
    nrm[0] = 0.0;
    if (work) {
      double lt[NRMAX];
#pragma unroll
      for (int r = 0; r < NRMAX; ++r) {
        lt[r] = sm_l[r * NT] + rr[r];
        sm_l[r * NT] = lt[r];
        sm_c[r * NT] = rr[r];
      }
#pragma unroll
      for (int j = 0; j < UMAX; ++j) stot[j] += sres[j];
      alpha_block<M, NRMAX, RMAXP>(B, lt, Psibc, xendc, nta, alph, a0tot);
    }
    PHX(28);
    const int any_check = __syncthreads_or(check ? 1 : 0);
    PHX(29);
    PH(10);
    bool converged = false;
    if (any_check) {
      PH_ADD(13, 1);
      double a0inc[X];
      if (work && check) {
        double ub = 0.0;
        alpha_block<M, NRMAX, RMAXP>(B, rr, Psibc, xendc, nta, alphi, a0inc, kapc, &ub);
        if (d.noisy)
          for (int k = 0; k < B.ny; ++k) ub = fmax(ub, fabs(sig_lin * rr[k < NRMAX ? k : 0]));
        if (B.ini) {
          double ti[M::V0];
          mtv<X, M::V0>(dx0_dv0, a0inc, ti);
#pragma unroll
          for (int j = 0; j < M::V0; ++j) ub = fmax(ub, fabs(ti[j]));
#pragma unroll
          for (int j = 0; j < UMAX; ++j)
            if (j < U) ub = fmax(ub, fabs(sres[j]));
        }
        nrm[0] = ub;
      }
      block_reduce<0, 1>(nrm, sm_red, t);
      const bool undecided = check && !(nrm[0] < ptol);
      if (check && !undecided) { converged = true; pend = true; }
      if (__syncthreads_or(undecided ? 1 : 0)) {
        PH_ADD(14, 1);
        double ex[1];
        ex[0] = 0.0;
        if (work && undecided)
          ex[0] = qn_pass<M, NRMAX, UMAX, true>(d, S, W, B, t, mode, cur, sl, mom_coef, sig_lin, dx0_dv0, alph, alphi,
                                                a0tot, a0inc, u0, stot, sres, sm_l, sm_c, NT, &final_norm);
        block_reduce<0, 1>(ex, sm_red, t);
        if (undecided) converged = ex[0] < ptol;
      }
      PH(11);
    }
    first_pass = false;
    if (!done) {
      it += 1;
      const bool diverged = (err > dtol) || (err != err);
      if (converged) {
        done = true;
      } else if (diverged) {
        done = true;
        st = ST_DIVERGED;
      } else if (it >= max_iters) {
        done = true;
        st = ST_NOTCONV;
      }
    }
  }
  // iterates (mode 0) / reverse-check distances (mode 1) of the chains accepted through the bound
  if (__syncthreads_or(pend ? 1 : 0)) {
    double a0inc[X], sres0[UMAX];
#pragma unroll
    for (int i = 0; i < X; ++i) a0inc[i] = 0.0;
#pragma unroll
    for (int j = 0; j < UMAX; ++j) sres0[j] = 0.0;
    if (in_blk && pend)
      qn_pass<M, NRMAX, UMAX, false>(d, S, W, B, t, mode, cur, sl, mom_coef, sig_lin, dx0_dv0, alph, alphi, a0tot, a0inc,
                                     u0, stot, sres0, sm_l, sm_c, NT, &final_norm);
    PH(15);
  }
  // epilogue
  const bool live = t.act && (W.status[t.cix] == 0);
  double rr[1];
  rr[0] = final_norm;
  block_reduce<0, 1>(rr, sm_red, t);
  if (t.slot == 0 && live) {
    W.iters[mode * d.n_tiles * cpb + t.cix] = it;
    W.itsum[t.cix] += it;
    if (st) W.status[t.cix] |= st;
    if (mode == 1 && st == 0) W.revd[t.cix] = rr[0];
  }
}

// commit / reject: successful chains flip to the new slot; reverse check (Mici
// ConstrainedLeapfrogIntegrator._step_b: reverse_check_norm(...) > reverse_check_tol)
MMD_D void dev_commit(const Dims& d, const Slots& S, const Work& W, double rev_tol, long long* __restrict__ n_ok,
                      int cix) {
  int st = W.status[cix];
  if (st == 0 && !(W.revd[cix] <= rev_tol)) {
    st |= ST_NONREV;
    W.status[cix] = st;
  }
  if (st == 0) {
    S.cur[cix] = 1 - S.cur[cix];
    n_ok[cix] += 1;
  }
}

// ------------------------------------------------------------------------------------------
// standalone launches of the phases (per-op API) and the fused persistent leapfrog kernel
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
template <class M, int NRMAX, int RMAX, int UMAX, int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_point(const __grid_constant__ Dims d, const __grid_constant__ Slots S, const __grid_constant__ Work W, const double* __restrict__ y, int part, int which, int with_grad) {
  dev_point<M, NRMAX, RMAX, UMAX>(d, S, W, y, part, which, with_grad);
}
template <class M, int NRMAX, int UMAX, int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_constr(const __grid_constant__ Dims d, const __grid_constant__ Slots S, const __grid_constant__ Work W, const double* __restrict__ y, int part, double* __restrict__ cout) {
  dev_constr<M, NRMAX, UMAX>(d, S, W, y, part, cout);
}
template <class M, int NRMAX, int RMAX, int UMAX, int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_project(const __grid_constant__ Dims d, const __grid_constant__ Slots S, const __grid_constant__ Work W, int part, int lin_sel, int src_sel, int dst_sel, double h, double qcoef,
          FlowCoef fl) {
  dev_project<M, NRMAX, RMAX, UMAX>(d, S, W, part, lin_sel, src_sel, dst_sel, h, qcoef, fl);
}
template <class M, int NRMAX, int RMAX, int UMAX, int NTMAX, int MINB, bool NEWTON>
__global__ void __launch_bounds__(NTMAX, MINB)
k_qn(const __grid_constant__ Dims d, const __grid_constant__ Slots S, const __grid_constant__ Work W, const double* __restrict__ y, int part, int mode, double mom_coef, double ctol,
     double ptol, double dtol, int max_iters) {
  dev_qn<M, NRMAX, RMAX, UMAX, NEWTON>(d, S, W, y, part, mode, mom_coef, ctol, ptol, dtol, max_iters);
}
static __global__ void k_commit(Dims d, Slots S, Work W, double rev_tol, long long* __restrict__ n_ok) {
  const int cix = blockIdx.x * blockDim.x + threadIdx.x;
  if (cix >= d.n_chains) return;
  dev_commit(d, S, W, rev_tol, n_ok, cix);
}

// One (or n_steps) full ConstrainedLeapfrogIntegrator.step per chain in ONE launch: a CTA carries
// its tile of chains through every phase with CTA-local barriers only, so chains that need many
// projection iterations delay just their own small CTA while the other resident CTAs keep the SM busy
// (the per-chain iteration count is long-tailed: mean ~8, 1 % > 30, max_iters = 50).
// Step order: Mici ConstrainedLeapfrogIntegrator._step = A(dt/2) B(dt) A(dt/2), SURVEY.md 3.3.
template <class M, int NRMAX, int RMAX, int UMAX, int NTMAX, int MINB, bool NEWTON>
#ifdef MMD_LEAPFROG_MAXNREG
__global__ void __maxnreg__(MMD_LEAPFROG_MAXNREG)
#else
__global__ void __launch_bounds__(NTMAX, MINB)
#endif
k_leapfrog(const __grid_constant__ Dims d, const __grid_constant__ Slots S, const __grid_constant__ Work W, const double* __restrict__ y, int part, double dt, double ctol, double ptol,
           double dtol, int max_iters, double rev_tol, long long* __restrict__ n_ok, int n_steps,
           int reset_status) {
  const Tid t = thread_id(d);
  const FlowCoef noflow = {0, 1.0, 0.0, 0.0};
  // per-chain step sizes may be signed (per-chain integration direction); a negative scalar flips them all
  const StepCoef sc = make_step_coef(d.gaussian, W.use_dt_chain ? (dt < 0.0 ? -W.dt_chain[t.cix] : W.dt_chain[t.cix]) : dt);
  // Consecutive steps of one launch share a half kick: A(dt/2) at the end of step s and A(dt/2) at the start of step
  // s + 1 act at the same point with the same linearisation, and the cotangent projection P is linear and idempotent,
  // so P(P(p - h g) - h g) = P(p - 2 h g): one kick + projection instead of two (one block solve and ~20 doubles per
  // SDE step less).  A chain that fails in step s + 1 must be left exactly as Mici leaves it after step s, i.e. with
  // the second half kick applied: `owed` remembers that and the kick is made up after the loop.
  const bool fuse = (n_steps > 1) && !reset_status;
  int owed = 0;
  for (int s = 0; s < n_steps; ++s) {
    if (reset_status && t.slot == 0) W.status[t.cix] = 0;
    __syncthreads();
    PH_T0
    const bool first = !(fuse && s > 0), last = !(fuse && s + 1 < n_steps);
    dev_project<M, NRMAX, RMAX, UMAX>(d, S, W, part, 0, PSEL_CUR, PSEL_WORK, first ? sc.half_dt : 2.0 * sc.half_dt,
                                      sc.qcoef, sc.fwd);
    __syncthreads();
    PH(0);
    dev_qn<M, NRMAX, RMAX, UMAX, NEWTON>(d, S, W, y, part, 0, sc.mom_coef, ctol, ptol, dtol, max_iters);
    __syncthreads();
    PH(1);
    dev_point<M, NRMAX, RMAX, UMAX>(d, S, W, y, part, 1, 1);
    __syncthreads();
    PH(2);
    dev_project<M, NRMAX, RMAX, UMAX>(d, S, W, part, 1, PSEL_OTHER, PSEL_OTHER, 0.0, 0.0, sc.back);
    __syncthreads();
    PH(3);
    dev_qn<M, NRMAX, RMAX, UMAX, NEWTON>(d, S, W, y, part, 1, 0.0, ctol, ptol, dtol, max_iters);
    __syncthreads();
    PH(4);
    if (last) {
      dev_project<M, NRMAX, RMAX, UMAX>(d, S, W, part, 1, PSEL_OTHER, PSEL_OTHER, sc.half_dt, sc.qcoef, noflow);
      __syncthreads();
    }
    PH(5);
    if (t.slot == 0 && t.act) dev_commit(d, S, W, rev_tol, n_ok, t.cix);
    __syncthreads();
    // the chain now sits at the new point; without the closing half kick it owes one (paid by the next step's
    // opening kick, or below if that step fails)
    if (t.act && W.status[t.cix] == 0) owed = last ? 0 : 1;
    PH(6);
    PH_ADD(7, 1);
  }
  if (fuse) {
    const int pay = (t.act && owed && W.status[t.cix] != 0) ? 1 : 0;
    if (__syncthreads_or(pay))
      dev_project<M, NRMAX, RMAX, UMAX>(d, S, W, part, 0, PSEL_CUR, PSEL_CUR, sc.half_dt, sc.qcoef, noflow, pay);
  }
}

// Hamiltonian h = h1 + h2 (mici_extensions.py:1186-1202) for the current slot; both splittings give
// 1/2 |q|^2 + log det^{1/2} + 1/2 |p|^2 with the identity metric
template <class M>
__global__ void k_hamiltonian(Dims d, Slots S, Work W, int part, int sel, double* __restrict__ hout) {
  const Tid t = thread_id(d);
  extern __shared__ double smem[];
  const int cur = S.cur[t.cix];
  const int sl = sel ? 1 - cur : cur;
  const QPtr q = qptr<M>(S.q + sl * S.s_q, d, t);
  const QPtr p = qptr<M>(S.p + sl * S.s_q, d, t);
  double acc[1];
  acc[0] = 0.0;
  if (t.slot < d.nb[part]) {
    const Blk B = get_block<M>(d, part, t.slot);
    const int nstep = B.n * d.S;
    for (int s = 0; s < nstep; ++s) {
      double a[M::V], b[M::V];
      ldrec<M::V>(q.body + s * M::V * t.nta, a);
      ldrec<M::V>(p.body + s * M::V * t.nta, b);
#pragma unroll
      for (int j = 0; j < M::V; ++j) acc[0] += 0.5 * a[j] * a[j] + 0.5 * b[j] * b[j];
    }
    if (d.noisy)
      for (int k = 0; k < B.n; ++k) {
        const double a = q.noise[k * t.nta], b = p.noise[k * t.nta];
        acc[0] += 0.5 * a * a + 0.5 * b * b;
      }
    if (B.ini)
      for (int r = 0; r < d.rows_head; ++r) {
        const double a = q.head[r * t.cpb], b = p.head[r * t.cpb];
        acc[0] += 0.5 * a * a + 0.5 * b * b;
      }
  }
  block_reduce<1, 0>(acc, smem, t);
  if (t.slot == 0 && t.act) hout[t.cix] = acc[0] + S.ldv[sl * S.s_ld + t.cix];
}

// ------------------------------------------------------------------------------------------
// vector primitives on q-like arrays for host-driven tree building (batched dynamic HMC / NUTS):
// masked copy / axpy and the U-turn inner products.  Array selectors: VEC_Q / VEC_P = the live
// position / momentum (slot resolved per chain), otherwise a pointer to an auxiliary q-like array.
// ------------------------------------------------------------------------------------------
MMD_D int chain_of_qlike_index(const Dims& d, long long e) {
  int tile, cl;
  if (e < d.off_body) {
    tile = (int)(e / ((long long)d.rows_head * d.cpb));
    cl = (int)(e % d.cpb);
  } else if (e < d.off_noise) {
    const long long r = e - d.off_body;
    tile = (int)(r / ((long long)d.rows_body * d.nta));
    cl = (int)(((r % (d.V * d.nta)) / d.V) % d.cpb);
  } else {
    const long long r = e - d.off_noise;
    tile = (int)(r / ((long long)d.rows_noise * d.nta));
    cl = (int)(r % d.cpb);
  }
  return tile * d.cpb + cl;
}
// dst = beta * dst + alpha * src for the chains with mask != 0 (mask == nullptr: all chains).
// live_dst / live_src: 0 = plain array, 1 = slot array S.q / S.p indexed by cur[chain]
static __global__ void k_vec_axpby(Dims d, double* __restrict__ dst, int live_dst, const double* __restrict__ src,
                                   int live_src, long long slot_stride, const int* __restrict__ cur, double alpha,
                                   double beta, const int* __restrict__ mask) {
  const long long n = d.qsize;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = chain_of_qlike_index(d, e);
    if (chain >= d.n_chains || (mask && !mask[chain])) continue;
    const long long od = live_dst ? cur[chain] * slot_stride + e : e;
    const long long os = live_src ? cur[chain] * slot_stride + e : e;
    const double sv = (alpha == 0.0) ? 0.0 : alpha * src[os];
    dst[od] = (beta == 0.0) ? sv : fma(beta, dst[od], sv);
  }
}
// U-turn inner products per chain: with s = c - dd + a,  out1 = a . s,  out2 = e . s   (only valid rows)
template <class M>
__global__ void k_vec_uturn(Dims d, int part, const double* __restrict__ a, const double* __restrict__ dd,
                            const double* __restrict__ c, const double* __restrict__ e, int live_e,
                            long long slot_stride, const int* __restrict__ cur, double* __restrict__ out1,
                            double* __restrict__ out2) {
  const Tid t = thread_id(d);
  extern __shared__ double smem[];
  const long long oe = live_e ? (long long)cur[t.cix] * slot_stride : 0;
  const QPtr pa = qptr<M>(const_cast<double*>(a), d, t), pd = qptr<M>(const_cast<double*>(dd), d, t);
  const QPtr pc = qptr<M>(const_cast<double*>(c), d, t), pe = qptr<M>(const_cast<double*>(e) + oe, d, t);
  double acc[2] = {0.0, 0.0};
  if (t.slot < d.nb[part]) {
    const Blk B = get_block<M>(d, part, t.slot);
    const int nstep = B.n * d.S;
    for (int s = 0; s < nstep; ++s) {
      double va[M::V], vd[M::V], vc[M::V], ve[M::V];
      ldrec<M::V>(pa.body + s * M::V * t.nta, va);
      ldrec<M::V>(pd.body + s * M::V * t.nta, vd);
      ldrec<M::V>(pc.body + s * M::V * t.nta, vc);
      ldrec<M::V>(pe.body + s * M::V * t.nta, ve);
#pragma unroll
      for (int j = 0; j < M::V; ++j) {
        const double sv = vc[j] - vd[j] + va[j];
        acc[0] = fma(va[j], sv, acc[0]);
        acc[1] = fma(ve[j], sv, acc[1]);
      }
    }
    if (d.noisy)
      for (int k = 0; k < B.n; ++k) {
        const double sv = pc.noise[k * t.nta] - pd.noise[k * t.nta] + pa.noise[k * t.nta];
        acc[0] = fma(pa.noise[k * t.nta], sv, acc[0]);
        acc[1] = fma(pe.noise[k * t.nta], sv, acc[1]);
      }
    if (B.ini)
      for (int r = 0; r < d.rows_head; ++r) {
        const double sv = pc.head[r * t.cpb] - pd.head[r * t.cpb] + pa.head[r * t.cpb];
        acc[0] = fma(pa.head[r * t.cpb], sv, acc[0]);
        acc[1] = fma(pe.head[r * t.cpb], sv, acc[1]);
      }
  }
  block_reduce<2, 0>(acc, smem, t);
  if (t.slot == 0 && t.act) {
    out1[t.cix] = acc[0];
    out2[t.cix] = acc[1];
  }
}
// park / release chains: sets or clears ST_INACTIVE according to mask; clear_errors also drops the error bits
static __global__ void k_set_inactive(Dims d, Work W, const int* __restrict__ mask, int clear_errors) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_chains) return;
  int st = W.status[c];
  if (clear_errors) st = 0;
  st = mask && mask[c] ? (st | ST_INACTIVE) : (st & ~ST_INACTIVE);
  W.status[c] = st;
}

// ------------------------------------------------------------------------------------------
// layout conversion: reference per-chain layout (canonical, [chain][dim_q] row-major, AoS) <-> tile
// layout of a q-like vector, for the given partition.  One thread per canonical element.
// ------------------------------------------------------------------------------------------
template <class M>
MMD_D long long tile_index(const Dims& d, int part, int chain, int i) {
  const int tile = chain >> d.lcpb, cl = chain & (d.cpb - 1);
  if (i < d.off_v) return ((long long)tile * d.rows_head + i) * d.cpb + cl;
  if (i < d.off_n) {
    const int g = (i - d.off_v) / M::V, j = (i - d.off_v) % M::V;
    const int o = g / d.S, tt = g % d.S;
    const int b = block_of_obs(d, part, o);
    const Blk B = get_block<M>(d, part, b);
    const int s = (o - B.o) * d.S + tt;
    return d.off_body + ((long long)tile * d.rows_body + s * M::V) * d.nta + ((b << d.lcpb) + cl) * M::V + j;
  }
  const int o = i - d.off_n;
  const int b = block_of_obs(d, part, o);
  const Blk B = get_block<M>(d, part, b);
  return d.off_noise + ((long long)tile * d.rows_noise + (o - B.o)) * d.nta + (b << d.lcpb) + cl;
}
// per-chain slot selection: sel 0 cur, 1 other, 2 none (base used as is)
template <class M>
__global__ void k_pack(Dims d, int part, const double* __restrict__ canon, double* __restrict__ base,
                       long long slot_stride, const int* __restrict__ cur, int sel) {
  const long long n = (long long)d.n_chains * d.dim_q;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / d.dim_q), i = (int)(e % d.dim_q);
    const int sl = sel == 2 ? 0 : (sel ? 1 - cur[chain] : cur[chain]);
    base[sl * slot_stride + tile_index<M>(d, part, chain, i)] = canon[e];
  }
}
template <class M>
__global__ void k_unpack(Dims d, int part, double* __restrict__ canon, const double* __restrict__ base,
                         long long slot_stride, const int* __restrict__ cur, int sel) {
  const long long n = (long long)d.n_chains * d.dim_q;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / d.dim_q), i = (int)(e % d.dim_q);
    const int sl = sel == 2 ? 0 : (sel ? 1 - cur[chain] : cur[chain]);
    canon[e] = base[sl * slot_stride + tile_index<M>(d, part, chain, i)];
  }
}
// re-tile the current position from partition `pa` to partition `pb` (SwitchPartitionTransition)
template <class M>
// newpos (optional): the chain in slot `chain` moves to slot newpos[chain] (chain regrouping, mmd_set_chain_regrouping)
__global__ void k_retile(Dims d, int pa, int pb, const double* __restrict__ src, double* __restrict__ dst,
                         long long slot_stride, const int* __restrict__ cur, const int* __restrict__ newpos) {
  const long long n = (long long)d.n_chains * d.dim_q;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / d.dim_q), i = (int)(e % d.dim_q);
    const int to = newpos ? newpos[chain] : chain;
    dst[tile_index<M>(d, pb, to, i)] = src[cur[chain] * slot_stride + tile_index<M>(d, pa, chain, i)];
  }
}
// per-slot arrays follow their chains: dst[newpos[i]] = src[i]
template <class T>
__global__ void k_permute(int n, const int* __restrict__ newpos, const T* __restrict__ src, T* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[newpos[i]] = src[i];
}
// per-chain arrays [chain][rows] (canonical) <-> [tile][rows][cpb]
static __global__ void k_pack_chain(Dims d, int rows, const double* __restrict__ canon, double* __restrict__ dst) {
  const long long n = (long long)d.n_chains * rows;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / rows), r = (int)(e % rows);
    dst[((long long)(chain >> d.lcpb) * rows + r) * d.cpb + (chain & (d.cpb - 1))] = canon[e];
  }
}
static __global__ void k_unpack_chain(Dims d, int rows, double* __restrict__ canon, const double* __restrict__ src) {
  const long long n = (long long)d.n_chains * rows;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / rows), r = (int)(e % rows);
    canon[e] = src[((long long)(chain >> d.lcpb) * rows + r) * d.cpb + (chain & (d.cpb - 1))];
  }
}
// thread-private array (rows per thread, block-local rows; records of width W, W = 1 for plain columns)
// -> [chain][nb][rows] for tests / factor export
static __global__ void k_unpack_tp(Dims d, int part, int rows, int W, double* __restrict__ out,
                            const double* __restrict__ base, long long slot_stride, const int* __restrict__ cur) {
  const int nb = d.nb[part];
  const long long n = (long long)d.n_chains * nb * rows;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e % rows);
    const int b = (int)((e / rows) % nb);
    const int chain = (int)(e / ((long long)rows * nb));
    const int tile = chain >> d.lcpb, cl = chain & (d.cpb - 1);
    const int tid = (b << d.lcpb) + cl;
    const int s = r / W, c = r % W;
    out[e] = base[cur[chain] * slot_stride + ((long long)tile * rows + s * W) * d.nta + tid * W + c];
  }
}

// full forward scan keeping the states at observation times (generate_x_obs_seq :384-397); one
// thread per chain walks through the blocks of the tile layout of the current position
template <class M, int UMAX>
__global__ void k_gen_xobs(Dims d, Slots S, Work W, int part) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= d.n_chains) return;
  constexpr int X = M::X, V = M::V;
  const int tile = chain >> d.lcpb, cl = chain & (d.cpb - 1);
  const double* base = S.q + S.cur[chain] * S.s_q;
  const double* head = base + ((long long)tile * d.rows_head) * d.cpb + cl;
  double u[UMAX], z[M::Z], dzdu[M::Z * M::Z], v0[M::V0], x[X];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) u[j] = (j < d.U) ? head[j * d.cpb] : 0.0;
  M::gen_z(d.gen, u, z, dzdu);
  typename M::Coef C;
  M::make_coef(z, d.sd, C);
  ldcol<M::V0>(head + d.U * d.cpb, d.cpb, v0);
  M::gen_x0(d.gen, z, v0, x);
  double* xo = W.xobs + ((long long)tile * d.T * X) * d.cpb + cl;
  for (int b = 0; b < d.nb[part]; ++b) {
    const Blk B = get_block<M>(d, part, b);
    const double* body = base + d.off_body + ((long long)tile * d.rows_body) * d.nta + ((b << d.lcpb) + cl) * V;
    for (int k = 0; k < B.n; ++k) {
      for (int tt = 0; tt < d.S; ++tt) {
        double v[V], xn[X];
        ldrec<V>(body + (k * d.S + tt) * V * d.nta, v);
        M::step(C, x, v, xn);
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = xn[i];
      }
      stcol<X>(xo + (B.o + k) * X * d.cpb, d.cpb, x);
    }
  }
}

// find_initial_state_by_linear_interpolation (mici_extensions.py:1479-1547), batched: one thread per
// (chain, observation block).  Given u, v_0 (already in the head of q) and the states at observation
// times, solves per step for the noise vector that makes the discretised path interpolate linearly
// between them (forward_func is affine in v; d f / d v square and invertible: X == V).
template <class M, int UMAX>
__global__ void k_init_interp(Dims d, Slots S, Work W, int part) {
  constexpr int X = M::X, V = M::V, Z = M::Z;
  static_assert(X == V, "linear-interpolation initialiser needs a square noise Jacobian");
  const Tid t = thread_id(d);
  if (!t.act || t.slot >= d.nb[part]) return;
  const QPtr q = qptr<M>(S.q + S.cur[t.cix] * S.s_q, d, t);
  const double* xoc = pc(W.xobs, d.T * X, t);
  double u[UMAX], z[Z], dzdu[Z * Z];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) u[j] = (j < d.U) ? q.head[j * t.cpb] : 0.0;
  M::gen_z(d.gen, u, z, dzdu);
  typename M::Coef C;
  M::make_coef(z, d.sd, C);
  const Blk B = get_block<M>(d, part, t.slot);
  for (int k = 0; k < B.n; ++k) {
    double xa[X], xb[X];
    const int o = B.o + k;
    if (o == 0) {
      double v0[M::V0];
      ldcol<M::V0>(q.head + d.U * t.cpb, t.cpb, v0);
      M::gen_x0(d.gen, z, v0, xa);
    } else {
      ldcol<X>(xoc + (o - 1) * X * t.cpb, t.cpb, xa);
    }
    ldcol<X>(xoc + o * X * t.cpb, t.cpb, xb);
    double dx[X];
#pragma unroll
    for (int i = 0; i < X; ++i) dx[i] = (xb[i] - xa[i]) / d.S;
    for (int s = 0; s < d.S; ++s) {
      double x[X], vz[V], m[X], Bm[X * V], rhs[X];
#pragma unroll
      for (int i = 0; i < X; ++i) x[i] = xa[i] + s * dx[i];
#pragma unroll
      for (int j = 0; j < V; ++j) vz[j] = 0.0;
      M::step(C, x, vz, m);
      M::jac_v(C, x, vz, Bm);
#pragma unroll
      for (int i = 0; i < X; ++i) rhs[i] = dx[i] - (m[i] - x[i]);
      // Gaussian elimination with partial pivoting on the X x X system Bm v = rhs
      for (int c = 0; c < X; ++c) {
        int piv = c;
        for (int r = c + 1; r < X; ++r)
          if (fabs(Bm[r * V + c]) > fabs(Bm[piv * V + c])) piv = r;
        if (piv != c) {
          for (int j = 0; j < V; ++j) { double tw = Bm[c * V + j]; Bm[c * V + j] = Bm[piv * V + j]; Bm[piv * V + j] = tw; }
          double tw = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = tw;
        }
        for (int r = c + 1; r < X; ++r) {
          const double f = Bm[r * V + c] / Bm[c * V + c];
          for (int j = c; j < V; ++j) Bm[r * V + j] -= f * Bm[c * V + j];
          rhs[r] -= f * rhs[c];
        }
      }
      double v[V];
      for (int r = X - 1; r >= 0; --r) {
        double tw = rhs[r];
        for (int j = r + 1; j < V; ++j) tw -= Bm[r * V + j] * v[j];
        v[r] = tw / Bm[r * V + r];
      }
      strec<V>(q.body + (k * d.S + s) * V * t.nta, v);
    }
    if (d.noisy) q.noise[k * t.nta] = 0.0;
  }
}

// sample_momentum (:1256-1259), draw part: p = standard normals, Philox-4x32-10 keyed by
// (seed, offset) with counter (chain, canonical element / 2): the stream does not depend on the
// tile shape, the partition or the number of GPUs.  One thread per (chain, canonical pair).
template <class M>
__global__ void k_philox_momentum(Dims d, int part, double* __restrict__ base, long long slot_stride,
                                  const int* __restrict__ cur, uint64_t seed, uint64_t offset, int chain0,
                                  const int* __restrict__ slot_chain) {
  const int npair = (d.dim_q + 1) / 2;
  const long long n = (long long)d.n_chains * npair;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(e / npair), pr = (int)(e % npair);
    double a, b;
    // the stream belongs to the chain, not to the slot it currently occupies
    const int cid = chain0 + (slot_chain ? slot_chain[chain] : chain);
    philox_normal_pair(seed, offset, ((uint64_t)cid << 32) | (uint64_t)pr, &a, &b);
    double* dst = base + cur[chain] * slot_stride;
    dst[tile_index<M>(d, part, chain, 2 * pr)] = a;
    if (2 * pr + 1 < d.dim_q) dst[tile_index<M>(d, part, chain, 2 * pr + 1)] = b;
  }
}

// Metropolis accept step of a static-trajectory constrained HMC transition, on device.
// accept iff the trajectory finished without an integrator error, the final Hamiltonian is finite
// and log(uniform) < h0 - h1 (Mici MetropolisStaticIntegrationTransition semantics; IntegratorError
// -> reject).  acc_prob is the `accept_stat` statistic min(1, exp(h0 - h1)).
// A rejected chain returns to the slot it started the transition in (cur0).
static __global__ void k_decide(Dims d, Slots S, Work W, const double* __restrict__ h0, const double* __restrict__ h1,
                         const int* __restrict__ cur0, uint64_t seed, uint64_t offset, int chain0,
                         int* __restrict__ accepted, double* __restrict__ acc_prob,
                         const int* __restrict__ slot_chain) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= d.n_chains) return;
  uint32_t c[4] = {(uint32_t)(chain0 + (slot_chain ? slot_chain[chain] : chain)), 0u, (uint32_t)offset,
                   (uint32_t)(offset >> 32)};
  philox4x32_10(c, (uint32_t)seed ^ 0x5bd1e995u, (uint32_t)(seed >> 32));
  const double uu = ((double)((((uint64_t)c[1] << 32) | c[0]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const int st = W.status[chain];
  const double dh = h0[chain] - h1[chain];
  const bool finite = (dh == dh) && (fabs(h1[chain]) < 1.0e300);
  double ap = 0.0;
  int acc = 0;
  if (st == 0 && finite) {
    ap = dh >= 0.0 ? 1.0 : exp(dh);
    acc = log(uu) < dh;
  } else if (st == 0) {
    W.status[chain] = ST_NONFINITE;
  }
  accepted[chain] = acc;
  acc_prob[chain] = ap;
  (void)S; (void)cur0;
}
// DualAveragingStepSizeAdapter.update (Mici 0.1.10) per chain, on device:
//   w = 1/(offset + n); err <- (1-w) err + w (target - accept_stat); log eps = reg_target - err sqrt(n) / reg_coef;
//   smoothed <- (1 - n^-decay) smoothed + n^-decay log eps; step size <- exp(log eps)
// state: [4][nc] = iteration, smoothed log step size, adapt-stat error, regularisation target
static __global__ void k_dual_averaging(Dims d, double* __restrict__ state, long long nc,
                                        const double* __restrict__ accept_stat, double target, double reg_coef,
                                        double decay, double offset, double* __restrict__ dt_chain) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_chains) return;
  const double n = state[c] + 1.0;
  state[c] = n;
  const double w = 1.0 / (offset + n);
  const double err = (1.0 - w) * state[2 * nc + c] + w * (target - accept_stat[c]);
  state[2 * nc + c] = err;
  const double log_eps = state[3 * nc + c] - err * sqrt(n) / reg_coef;
  const double sw = pow(n, -decay);
  state[nc + c] = (1.0 - sw) * state[nc + c] + sw * log_eps;
  dt_chain[c] = exp(log_eps);
}
// DualAveragingStepSizeAdapter.finalize: step size = exp(smoothed log step size) per chain, or (pool) the mean
// over chains like Mici's finalize with several chains
static __global__ void k_dual_averaging_finalize(Dims d, const double* __restrict__ state, long long nc, int pool,
                                                 double* __restrict__ dt_chain) {
  __shared__ double red[256];
  double mean = 0.0;
  if (pool) {
    double acc = 0.0;
    for (int c = threadIdx.x; c < d.n_chains; c += blockDim.x) acc += exp(state[nc + c]);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
      if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
      __syncthreads();
    }
    mean = red[0] / d.n_chains;
  }
  for (int c = threadIdx.x; c < d.n_chains; c += blockDim.x) dt_chain[c] = pool ? mean : exp(state[nc + c]);
}

// restore the position of rejected chains from the copy taken at the start of the transition
static __global__ void k_restore(Dims d, Slots S, const int* __restrict__ accepted, const double* __restrict__ qsave) {
  // qsave holds the start-of-transition position in tile layout (same partition); element e of a
  // q-like vector belongs to chain (tile(e), cl(e)); decode from the section it falls into
  const long long n = d.qsize;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    int tile, cl;
    if (e < d.off_body) {
      tile = (int)(e / ((long long)d.rows_head * d.cpb));
      cl = (int)(e % d.cpb);
    } else if (e < d.off_noise) {
      const long long r = e - d.off_body;
      tile = (int)(r / ((long long)d.rows_body * d.nta));
      cl = (int)(((r % (d.V * d.nta)) / d.V) % d.cpb);
    } else {
      const long long r = e - d.off_noise;
      tile = (int)(r / ((long long)d.rows_noise * d.nta));
      cl = (int)(r % d.cpb);
    }
    const int chain = tile * d.cpb + cl;
    if (chain < d.n_chains && !accepted[chain]) S.q[S.cur[chain] * S.s_q + e] = qsave[e];
  }
}
// snapshot of the current position (tile layout) at the start of a transition
// Snapshot / roll-back of position AND momentum of the live slot (steps with n_inner_step > 1: a chain that fails in a
// later inner step must be left exactly where it was before the whole step, mici ConstrainedLeapfrogIntegrator.step
// works on a copy).  restore: chains with a non-zero status only.
static __global__ void k_snapshot_qp(Dims d, Slots S, double* __restrict__ qs, double* __restrict__ ps) {
  const long long n = d.qsize;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int chain = chain_of_qlike_index(d, e);
    const long long o = (long long)S.cur[chain < d.n_tiles * d.cpb ? chain : 0] * S.s_q + e;
    qs[e] = S.q[o];
    ps[e] = S.p[o];
  }
}
static __global__ void k_restore_qp(Dims d, Slots S, Work W, const double* __restrict__ qs, const double* __restrict__ ps) {
  const long long n = d.qsize;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int chain = chain_of_qlike_index(d, e);
    if (chain < d.n_chains && W.status[chain] != 0) {
      const long long o = (long long)S.cur[chain] * S.s_q + e;
      S.q[o] = qs[e];
      S.p[o] = ps[e];
    }
  }
}
// park the chains without an error and clear the error of the others (so that the next k_point re-linearises exactly
// the rolled-back chains), and undo it afterwards
static __global__ void k_status_swap(Dims d, Work W, int* __restrict__ saved, int begin) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_chains) return;
  if (begin) {
    const int st = W.status[c];
    saved[c] = st;
    W.status[c] = st != 0 ? 0 : ST_INACTIVE;
  } else {
    W.status[c] = saved[c];
  }
}
static __global__ void k_snapshot(Dims d, Slots S, double* __restrict__ qsave) {
  const long long n = d.qsize;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    int tile, cl;
    if (e < d.off_body) {
      tile = (int)(e / ((long long)d.rows_head * d.cpb));
      cl = (int)(e % d.cpb);
    } else if (e < d.off_noise) {
      const long long r = e - d.off_body;
      tile = (int)(r / ((long long)d.rows_body * d.nta));
      cl = (int)(((r % (d.V * d.nta)) / d.V) % d.cpb);
    } else {
      const long long r = e - d.off_noise;
      tile = (int)(r / ((long long)d.rows_noise * d.nta));
      cl = (int)(r % d.cpb);
    }
    const int chain = tile * d.cpb + cl;
    qsave[e] = S.q[S.cur[chain < d.n_tiles * d.cpb ? chain : 0] * S.s_q + e];
  }
}
#endif  // __CUDACC__

}  // namespace mmd
