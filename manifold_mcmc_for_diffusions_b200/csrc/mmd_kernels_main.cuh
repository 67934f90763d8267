// __global__ kernels of the CHMC hot path.  See mmd_kernels.cuh for the layout and the mapping.
#pragma once
#include "mmd_kernels.cuh"
#include "mmd_philox.cuh"

namespace mmd {

#define MMD_THREAD_SETUP                                        \
  const int cl = threadIdx.x % CPB;                             \
  const int slot = threadIdx.x / CPB;                           \
  const int nslot = blockDim.x / CPB;                           \
  const int chain = blockIdx.x * CPB + cl;                      \
  const bool act = chain < d.n_chains;                          \
  const long long ld = d.ld;                                    \
  extern __shared__ double smem[];                              \
  (void)nslot; (void)act;

// ------------------------------------------------------------------------------------------
// k_point: everything cached at a position.  Phase 1 = jacob_constr_blocks + chol_gram_blocks +
// log_det_sqrt_gram (mici_extensions.py:521-687, 800-820) in compressed form; phase 2 =
// grad_log_det_sqrt_gram (:1143-1146, :1173-1184) by a hand-derived second-order adjoint.
// ------------------------------------------------------------------------------------------
template <class M, int CPB, int NRMAX, int RMAX, int UMAX>
MMD_D void dev_point(const Dims& d, const Slots& S, const Work& W, const double* __restrict__ xobs,
                     const double* __restrict__ y, int part, int which, int with_grad) {
  MMD_THREAD_SETUP
  constexpr int X = M::X, V = M::V, Z = M::Z;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  constexpr int UTRI = UMAX * (UMAX + 1) / 2;
  const int U = d.U;
  const int cur = S.cur[chain];
  const int sl = which ? 1 - cur : cur;
  const bool skip = (W.status[chain] != 0);
  const double* qc = S.q + sl * S.s_q + chain;
  double* Kc = S.K + sl * S.s_K + chain;
  double* Psibc = S.Psib + sl * S.s_Psib + chain;
  double* Ac = S.A + sl * S.s_A + chain;
  double* DinvAc = S.DinvA + sl * S.s_A + chain;
  double* LCc = S.LC + sl * S.s_LC + chain;
  double* gc = S.gradld + sl * S.s_q + chain;
  const double* xoc = xobs + chain;
  double* xsc = W.xs + chain;
  double* Qc = W.Qk + chain;
  double* Ztc = W.Zt + chain;

  double u[UMAX], z[Z], dzdu[Z * Z];
  for (int j = 0; j < U; ++j) u[j] = qc[(long long)j * ld];
  M::gen_z(u, z, dzdu);
  const double sigy = sigma_of<M>(d, u);
  const bool has_blk = slot < d.nb[part];
  Blk B;
  if (has_blk) B = get_block<M>(d, part, slot);
  double* Lc = S.L + sl * S.s_L + chain + (long long)slot * NTRI * ld;

  double red[UTRI + 1];
#pragma unroll
  for (int i = 0; i < UTRI + 1; ++i) red[i] = 0.0;

  if (has_blk && !skip) {
    double dx0_dv0[X * M::V0], dx0_dz[X * Z];
    M::gen_x0_jac(z, dx0_dv0, dx0_dz);
    // ---------------- interval sweeps: trajectory, compressed Jacobian, interval summaries
    double x[X];
    if (B.ini) {
      double v0[M::V0];
      ldcol<M::V0>(qc + (long long)d.off_v0 * ld, ld, v0);
      M::gen_x0(z, v0, x);
    } else {
      ldcol<X>(xoc + (long long)(B.o - 1) * X * ld, ld, x);
    }
    for (int k = 0; k < B.n; ++k) {
      const long long g0 = (long long)(B.o + k) * d.S;
      const double* vp = qc + ((long long)d.off_v + g0 * V) * ld;
      for (int t = 0; t < d.S; ++t) {
        stcol<X>(xsc + (g0 + t) * X * ld, ld, x);
        double v[V], xn[X];
        ldcol<V>(vp + (long long)t * V * ld, ld, v);
        M::step(z, d.sd, x, v, xn);
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = xn[i];
      }
      double Psi[X * X], Qk[X * X], Zk[X * Z];
#pragma unroll
      for (int i = 0; i < X * X; ++i) { Psi[i] = (i % (X + 1) == 0) ? 1.0 : 0.0; Qk[i] = 0.0; }
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Zk[i] = 0.0;
      for (int t = d.S - 1; t >= 0; --t) {
        double xt[X], v[V], F[X * X], Bm[X * V], G[X * Z], Kt[X * V], tmp[X * X];
        ldcol<X>(xsc + (g0 + t) * X * ld, ld, xt);
        ldcol<V>(vp + (long long)t * V * ld, ld, v);
        M::jac_x(z, d.sd, xt, v, F);
        M::jac_v(z, d.sd, xt, v, Bm);
        M::jac_z(z, d.sd, xt, v, G);
        mm<X, V, X>(Psi, Bm, Kt);
        stcol<X * V>(Kc + (g0 + t) * X * V * ld, ld, Kt);
        // Qk += Kt Kt^T ; Zk += Psi G ; Psi = Psi F
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < X; ++j) {
            double s = Qk[i * X + j];
#pragma unroll
            for (int l = 0; l < V; ++l) s = fma(Kt[i * V + l], Kt[j * V + l], s);
            Qk[i * X + j] = s;
          }
        mm_acc<X, Z, X>(Psi, G, Zk);
        mm<X, X, X>(Psi, F, tmp);
#pragma unroll
        for (int i = 0; i < X * X; ++i) Psi[i] = tmp[i];
      }
      stcol<X * X>(Psibc + (long long)(B.o + k) * X * X * ld, ld, Psi);
      stcol<X * X>(Qc + (long long)(B.o + k) * X * X * ld, ld, Qk);
      stcol<X * Z>(Ztc + (long long)(B.o + k) * X * Z * ld, ld, Zk);
      if (!M::OBS_LINEAR) {
        // state at the observation time is needed for the observation gradient
      }
    }
    // ---------------- per-observation algebra: A_b = dc/du rows, D_b = J_v J_v^T (+ noise terms)
    double Su[X * Z], P[X * X], w[NRMAX * X], Dm[NTRI], Am[NRMAX * UMAX];
#pragma unroll
    for (int i = 0; i < X * Z; ++i) Su[i] = B.ini ? dx0_dz[i] : 0.0;
    if (B.ini) {
      mmt<X, X, M::V0>(dx0_dv0, dx0_dv0, P);
    } else {
#pragma unroll
      for (int i = 0; i < X * X; ++i) P[i] = 0.0;
    }
    int nrow_done = 0;
    for (int k = 0; k < B.n; ++k) {
      double Ps[X * X], Qk[X * X], Zk[X * Z], t1[X * Z], t2[X * X], t3[X * X];
      ldcol<X * X>(Psibc + (long long)(B.o + k) * X * X * ld, ld, Ps);
      ldcol<X * X>(Qc + (long long)(B.o + k) * X * X * ld, ld, Qk);
      ldcol<X * Z>(Ztc + (long long)(B.o + k) * X * Z * ld, ld, Zk);
      mm<X, Z, X>(Ps, Su, t1);
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Su[i] = t1[i] + Zk[i];
      mm<X, X, X>(Ps, P, t2);
      mmt<X, X, X>(t2, Ps, t3);
#pragma unroll
      for (int i = 0; i < X * X; ++i) P[i] = t3[i] + Qk[i];
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
          if (k + 1 < B.n) ldcol<X>(xsc + (long long)(B.o + k + 1) * d.S * X * ld, ld, xe);
          else {
#pragma unroll
            for (int i = 0; i < X; ++i) xe[i] = x[i];
          }
          M::obs_grad(xe, h);
        } else {
          const int comp = a - ((k < B.ny) ? 1 : 0);
#pragma unroll
          for (int i = 0; i < X; ++i) h[i] = (i == comp) ? 1.0 : 0.0;
        }
        const int r = nrow_done;
        // w_r = P h ; D[r][j] = h . w_j ; Az[r] = h^T Su
        mv<X, X>(P, h, &w[r * X]);
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
            for (int m = 0; m < Z; ++m) s = fma(az[m], dzdu[m * Z + j], s);
          }
          Am[r * UMAX + j] = s;
        }
        if (d.noisy && yrow) {
          Dm[tri(r, r)] += sigy * sigy;
          if (d.noisy == 2) Am[r * UMAX + Z] = sigy * qc[((long long)d.off_n + B.o + k) * ld];
        }
        nrow_done++;
      }
    }
    for (int r = 0; r < B.nrows; ++r)
      for (int j = 0; j < U; ++j) Ac[((long long)(B.row0 + r) * U + j) * ld] = Am[r * UMAX + j];
    chol_packed<NRMAX>(Dm, B.nrows);
    for (int i = 0; i < B.nrows * (B.nrows + 1) / 2; ++i) Lc[(long long)i * ld] = Dm[i];
    double ldpart = 0.0;
    for (int r = 0; r < B.nrows; ++r) ldpart += log(fabs(Dm[tri(r, r)]));
    // DinvA and C_b = A_b^T D_b^{-1} A_b
    for (int j = 0; j < U; ++j) {
      double col[NRMAX];
      for (int r = 0; r < B.nrows; ++r) col[r] = Am[r * UMAX + j];
      chol_solve_packed(Dm, B.nrows, col);
      for (int r = 0; r < B.nrows; ++r) DinvAc[((long long)(B.row0 + r) * U + j) * ld] = col[r];
      for (int i = j; i < U; ++i) {
        double s = 0.0;
        for (int r = 0; r < B.nrows; ++r) s = fma(Am[r * UMAX + i], col[r], s);
        red[tri(i, j)] = s;
      }
    }
    red[UTRI] = ldpart;
  }
  block_reduce<UTRI + 1, CPB, false>(red, smem, nslot, slot, cl);
  double LCm[UTRI];
  for (int i = 0; i < U; ++i)
    for (int j = 0; j <= i; ++j) LCm[tri(i, j)] = red[tri(i, j)] + (i == j ? 1.0 : 0.0);
  chol_packed<UMAX>(LCm, U);
  double ldtot = red[UTRI];
  for (int i = 0; i < U; ++i) ldtot += log(fabs(LCm[tri(i, i)]));
  if (slot == 0 && !skip) {
    for (int i = 0; i < U * (U + 1) / 2; ++i) LCc[(long long)i * ld] = LCm[i];
    S.ldv[sl * S.s_ld + chain] = ldtot;
  }
  if (!with_grad) return;

  // =========================== phase 2: grad log det ======================================
  double gu[UMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) gu[j] = 0.0;
  if (has_blk && !skip) {
    double dx0_dv0[X * M::V0], dx0_dz[X * Z];
    M::gen_x0_jac(z, dx0_dv0, dx0_dz);
    const int nr = B.nrows;
    // C^{-1}
    double Cinv[UMAX * UMAX];
    for (int j = 0; j < U; ++j) {
      double e[UMAX];
      for (int i = 0; i < U; ++i) e[i] = (i == j) ? 1.0 : 0.0;
      chol_solve_packed(LCm, U, e);
      for (int i = 0; i < U; ++i) Cinv[i * UMAX + j] = e[i];
    }
    double Lm[NTRI], DiA[NRMAX * UMAX], Om[NRMAX * UMAX], E[NRMAX * NRMAX];
    for (int i = 0; i < nr * (nr + 1) / 2; ++i) Lm[i] = Lc[(long long)i * ld];
    for (int r = 0; r < nr; ++r)
      for (int j = 0; j < U; ++j) DiA[r * UMAX + j] = DinvAc[((long long)(B.row0 + r) * U + j) * ld];
    for (int r = 0; r < nr; ++r)
      for (int j = 0; j < U; ++j) {
        double s = 0.0;
        for (int l = 0; l < U; ++l) s = fma(DiA[r * UMAX + l], Cinv[l * UMAX + j], s);
        Om[r * UMAX + j] = s;
      }
    for (int c2 = 0; c2 < nr; ++c2) {  // E[:, c2] = D^{-1} e_c2 - Om DiA[c2]^T
      double e[NRMAX];
      for (int i = 0; i < nr; ++i) e[i] = (i == c2) ? 1.0 : 0.0;
      chol_solve_packed(Lm, nr, e);
      for (int r = 0; r < nr; ++r) {
        double s = e[r];
        for (int j = 0; j < U; ++j) s = fma(-Om[r * UMAX + j], DiA[c2 * UMAX + j], s);
        E[r * NRMAX + c2] = s;
      }
    }
    // a[r][k] = Phi(t_kr, t_k)^T H_r^T  (zero for k > kr)
    double a[NRMAX * RMAX * X], be[NRMAX * RMAX * X];
    for (int i = 0; i < NRMAX * RMAX * X; ++i) a[i] = 0.0;  // indexed [(r * RMAX + k) * X + i]
    for (int r = 0; r < nr; ++r) {
      const int kr = (r < B.ny) ? r : B.n - 1;
      double vec[X];
      if (r < B.ny) {
        double xe[X];
        if (!M::OBS_LINEAR) {
          // x at obs time kr: start of next interval or (last) recomputed below
          if (kr + 1 < B.n) ldcol<X>(xsc + (long long)(B.o + kr + 1) * d.S * X * ld, ld, xe);
        }
        M::obs_grad(xe, vec);
      } else {
#pragma unroll
        for (int i = 0; i < X; ++i) vec[i] = (i == r - B.ny) ? 1.0 : 0.0;
      }
      for (int k = kr; k >= 0; --k) {
        if (k < kr) {
          double Ps[X * X], t[X];
          ldcol<X * X>(Psibc + (long long)(B.o + k + 1) * X * X * ld, ld, Ps);
          mtv<X, X>(Ps, vec, t);
#pragma unroll
          for (int i = 0; i < X; ++i) vec[i] = t[i];
        }
#pragma unroll
        for (int i = 0; i < X; ++i) a[(r * RMAX + k) * X + i] = vec[i];
      }
    }
    for (int r = 0; r < nr; ++r)
      for (int k = 0; k < B.n; ++k)
#pragma unroll
        for (int i = 0; i < X; ++i) {
          double s = 0.0;
          for (int s2 = 0; s2 < nr; ++s2) s = fma(E[r * NRMAX + s2], a[(s2 * RMAX + k) * X + i], s);
          be[(r * RMAX + k) * X + i] = s;
        }
    // u-directions in z-space: omz[r] = dzdu Om[r]
    double omz[NRMAX * Z];
    for (int r = 0; r < nr; ++r)
#pragma unroll
      for (int m = 0; m < Z; ++m) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < Z; ++j) s = fma(dzdu[m * Z + j], Om[r * UMAX + j], s);
        omz[r * Z + m] = s;
      }
    // per-interval constants M_k, LamZ_k, Yb_k ; tangent recursion ; Az rows (for d2z/du2 term)
    double dprev[NRMAX * X];
    for (int r = 0; r < nr; ++r) {
      if (B.ini) {
        double Ps[X * X], b0[X], t0[M::V0], t1[X], t2[X];
        ldcol<X * X>(Psibc + (long long)B.o * X * X * ld, ld, Ps);
        mtv<X, X>(Ps, &be[(r * RMAX + 0) * X], b0);
        mtv<X, M::V0>(dx0_dv0, b0, t0);
        mv<X, M::V0>(dx0_dv0, t0, t1);
        mv<X, Z>(dx0_dz, &omz[r * Z], t2);
#pragma unroll
        for (int i = 0; i < X; ++i) dprev[r * X + i] = t1[i] + t2[i];
      } else {
#pragma unroll
        for (int i = 0; i < X; ++i) dprev[r * X + i] = 0.0;
      }
    }
    double Suz[X * Z], Gam[Z * UMAX];
#pragma unroll
    for (int i = 0; i < X * Z; ++i) Suz[i] = B.ini ? dx0_dz[i] : 0.0;
    for (int i = 0; i < Z * UMAX; ++i) Gam[i] = 0.0;
    for (int k = 0; k < B.n; ++k) {
      double Mk[X * X], Lam[Z * X], Yb[X * X];
#pragma unroll
      for (int i = 0; i < X * X; ++i) { Mk[i] = 0.0; Yb[i] = 0.0; }
#pragma unroll
      for (int i = 0; i < Z * X; ++i) Lam[i] = 0.0;
      for (int r = 0; r < nr; ++r) {
        const double* ar = &a[(r * RMAX + k) * X];
        const double* br = &be[(r * RMAX + k) * X];
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < X; ++j) {
            Mk[i * X + j] = fma(br[i], ar[j], Mk[i * X + j]);
            Yb[i * X + j] = fma(dprev[r * X + i], ar[j], Yb[i * X + j]);
          }
#pragma unroll
        for (int m = 0; m < Z; ++m)
#pragma unroll
          for (int j = 0; j < X; ++j) Lam[m * X + j] = fma(omz[r * Z + m], ar[j], Lam[m * X + j]);
      }
      stcol<X * X>(W.Mk + chain + (long long)(B.o + k) * X * X * ld, ld, Mk);
      stcol<Z * X>(W.LamZ + chain + (long long)(B.o + k) * Z * X * ld, ld, Lam);
      stcol<X * X>(W.Yb + chain + (long long)(B.o + k) * X * X * ld, ld, Yb);
      double Ps[X * X], Qk[X * X], Zk[X * Z];
      ldcol<X * X>(Psibc + (long long)(B.o + k) * X * X * ld, ld, Ps);
      ldcol<X * X>(Qc + (long long)(B.o + k) * X * X * ld, ld, Qk);
      ldcol<X * Z>(Ztc + (long long)(B.o + k) * X * Z * ld, ld, Zk);
      for (int r = 0; r < nr; ++r) {
        double t1[X], t2[X], t3[X];
        mv<X, X>(Ps, &dprev[r * X], t1);
        mv<X, X>(Qk, &be[(r * RMAX + k) * X], t2);
        mv<X, Z>(Zk, &omz[r * Z], t3);
#pragma unroll
        for (int i = 0; i < X; ++i) dprev[r * X + i] = t1[i] + t2[i] + t3[i];
      }
      double ts[X * Z];
      mm<X, Z, X>(Ps, Suz, ts);
#pragma unroll
      for (int i = 0; i < X * Z; ++i) Suz[i] = ts[i] + Zk[i];
      for (int r = 0; r < nr; ++r) {
        const int kr = (r < B.ny) ? r : B.n - 1;
        if (kr != k) continue;
        double az[Z];
        mtv<X, Z>(Suz, &a[(r * RMAX + k) * X], az);  // a[r][kr] = H_r^T
#pragma unroll
        for (int m = 0; m < Z; ++m)
          for (int j = 0; j < U; ++j) Gam[m * UMAX + j] = fma(az[m], Om[r * UMAX + j], Gam[m * UMAX + j]);
      }
    }
    // ---------------- second-order sweeps, last interval first
    double gam[X], gz[Z];
#pragma unroll
    for (int i = 0; i < X; ++i) gam[i] = 0.0;
#pragma unroll
    for (int i = 0; i < Z; ++i) gz[i] = 0.0;
    double* Ywc = W.Yw + chain;
    for (int k = B.n - 1; k >= 0; --k) {
      const long long g0 = (long long)(B.o + k) * d.S;
      const double* vp = qc + ((long long)d.off_v + g0 * V) * ld;
      double Mk[X * X], Lam[Z * X], Y[X * X];
      ldcol<X * X>(W.Mk + chain + (long long)(B.o + k) * X * X * ld, ld, Mk);
      ldcol<Z * X>(W.LamZ + chain + (long long)(B.o + k) * Z * X * ld, ld, Lam);
      ldcol<X * X>(W.Yb + chain + (long long)(B.o + k) * X * X * ld, ld, Y);
      for (int t = 0; t < d.S; ++t) {
        stcol<X * X>(Ywc + (g0 + t) * X * X * ld, ld, Y);
        double xt[X], v[V], F[X * X], Bm[X * V], G[X * Z], Kt[X * V], KM[V * X], Yn[X * X];
        ldcol<X>(xsc + (g0 + t) * X * ld, ld, xt);
        ldcol<V>(vp + (long long)t * V * ld, ld, v);
        ldcol<X * V>(Kc + (g0 + t) * X * V * ld, ld, Kt);
        M::jac_x(z, d.sd, xt, v, F);
        M::jac_v(z, d.sd, xt, v, Bm);
        M::jac_z(z, d.sd, xt, v, G);
        mm<X, X, X>(F, Y, Yn);
        mtm<V, X, X>(Kt, Mk, KM);
        mm_acc<X, X, V>(Bm, KM, Yn);
        mm_acc<X, X, Z>(G, Lam, Yn);
#pragma unroll
        for (int i = 0; i < X * X; ++i) Y[i] = Yn[i];
      }
      double Psi[X * X];
#pragma unroll
      for (int i = 0; i < X * X; ++i) Psi[i] = (i % (X + 1) == 0) ? 1.0 : 0.0;
      for (int t = d.S - 1; t >= 0; --t) {
        double xt[X], v[V], F[X * X], Bm[X * V], G[X * Z], Kt[X * V], Yt[X * X];
        ldcol<X>(xsc + (g0 + t) * X * ld, ld, xt);
        ldcol<V>(vp + (long long)t * V * ld, ld, v);
        ldcol<X * X>(Ywc + (g0 + t) * X * X * ld, ld, Yt);
        M::jac_x(z, d.sd, xt, v, F);
        M::jac_v(z, d.sd, xt, v, Bm);
        M::jac_z(z, d.sd, xt, v, G);
        mm<X, V, X>(Psi, Bm, Kt);
        double Th[(X + V + Z) * X], MP[X * X], g[X + V + Z];
        mm<X, X, X>(Yt, Psi, &Th[0]);
        mm<X, X, X>(Mk, Psi, MP);
        mtm<V, X, X>(Kt, MP, &Th[X * X]);
        mm<Z, X, X>(Lam, Psi, &Th[(X + V) * X]);
        M::hess_contract(z, d.sd, xt, v, Th, g);
        double gv[V], gn[X], tz[Z];
        mtv<X, V>(Bm, gam, gv);
#pragma unroll
        for (int j = 0; j < V; ++j) gv[j] += g[X + j];
        stcol<V>(gc + ((long long)d.off_v + (g0 + t) * V) * ld, ld, gv);
        mtv<X, Z>(G, gam, tz);
#pragma unroll
        for (int m = 0; m < Z; ++m) gz[m] += tz[m] + g[X + V + m];
        mtv<X, X>(F, gam, gn);
#pragma unroll
        for (int i = 0; i < X; ++i) gam[i] = gn[i] + g[i];
        double tmp[X * X];
        mm<X, X, X>(Psi, F, tmp);
#pragma unroll
        for (int i = 0; i < X * X; ++i) Psi[i] = tmp[i];
      }
    }
    if (B.ini) {
      double gv0[M::V0], tz[Z];
      mtv<X, M::V0>(dx0_dv0, gam, gv0);
      stcol<M::V0>(gc + (long long)d.off_v0 * ld, ld, gv0);
      mtv<X, Z>(dx0_dz, gam, tz);
#pragma unroll
      for (int m = 0; m < Z; ++m) gz[m] += tz[m];
    }
    double extra[Z], GamZ[Z * Z];
#pragma unroll
    for (int m = 0; m < Z; ++m)
#pragma unroll
      for (int j = 0; j < Z; ++j) GamZ[m * Z + j] = Gam[m * UMAX + j];
    M::gen_z_second(u, z, GamZ, extra);
#pragma unroll
    for (int j = 0; j < Z; ++j) {
      double s = extra[j];
#pragma unroll
      for (int m = 0; m < Z; ++m) s = fma(dzdu[m * Z + j], gz[m], s);
      gu[j] = s;
    }
  }
  block_reduce<UMAX, CPB, false>(gu, smem, nslot, slot, cl);
  if (slot == 0 && !skip)
    for (int j = 0; j < U; ++j) gc[(long long)j * ld] = gu[j];
}

// ------------------------------------------------------------------------------------------
// k_constr: c(q) for the standalone `constr` op (mici_extensions.py:473-519)
// ------------------------------------------------------------------------------------------
template <class M, int CPB, int NRMAX, int UMAX>
__global__ void k_constr(Dims d, const double* __restrict__ q, const double* __restrict__ xobs,
                         const double* __restrict__ y, int part, double* __restrict__ cout) {
  MMD_THREAD_SETUP
  constexpr int X = M::X, Z = M::Z;
  if (slot >= d.nb[part]) return;
  const double* qc = q + chain;
  double u[UMAX], z[Z], dzdu[Z * Z];
  for (int j = 0; j < d.U; ++j) u[j] = qc[(long long)j * ld];
  M::gen_z(u, z, dzdu);
  const double sigy = sigma_of<M>(d, u);
  Blk B = get_block<M>(d, part, slot);
  double x[X], crow[NRMAX];
  if (B.ini) {
    double v0[M::V0];
    ldcol<M::V0>(qc + (long long)d.off_v0 * ld, ld, v0);
    M::gen_x0(z, v0, x);
  } else {
    ldcol<X>(xobs + chain + (long long)(B.o - 1) * X * ld, ld, x);
  }
  constr_block<M, false>(d, B, z, sigy, x, qc, xobs + chain, y, nullptr, nullptr, ld, crow, nullptr);
  for (int r = 0; r < B.nrows; ++r) cout[(long long)(B.row0 + r) * ld + chain] = crow[r];
}

// ------------------------------------------------------------------------------------------
// k_project: p <- proj(p - h (a q + gradld)) with proj = I - J^T (J J^T)^{-1} J evaluated from the cached
// compressed Jacobian (normal_space_component / project_onto_cotangent_space, :983-993,
// :1243-1254) fused with the preceding h1_flow (Mici System.h1_flow; dh1_dpos :1192-1196).
// Reads p from slot `src`, writes slot `dst` (relative to cur: 0 = cur, 1 = other).
// ------------------------------------------------------------------------------------------
template <class M, int CPB, int NRMAX, int UMAX>
MMD_D void dev_project(const Dims& d, const Slots& S, const Work& W, int part, int lin_sel, int src_sel,
                       int dst_sel, double h, double qcoef) {
  MMD_THREAD_SETUP
  constexpr int X = M::X, V = M::V, Z = M::Z;
  const int U = d.U;
  const int cur = S.cur[chain];
  const bool skip = (W.status[chain] != 0);
  const int sl = lin_sel ? 1 - cur : cur;
  const int ss = src_sel ? 1 - cur : cur;
  const int sd_ = dst_sel ? 1 - cur : cur;
  const double* qc = S.q + sl * S.s_q + chain;
  const double* gc = S.gradld + sl * S.s_q + chain;
  const double* psrc = S.p + ss * S.s_q + chain;
  double* pdst = S.p + sd_ * S.s_q + chain;
  const double* Kc = S.K + sl * S.s_K + chain;
  const double* Psibc = S.Psib + sl * S.s_Psib + chain;
  const double* Ac = S.A + sl * S.s_A + chain;
  const double* DinvAc = S.DinvA + sl * S.s_A + chain;
  const double* LCc = S.LC + sl * S.s_LC + chain;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  const double* Lc = S.L + sl * S.s_L + chain + (long long)slot * NTRI * ld;
  const bool has_blk = slot < d.nb[part] && !skip;
  Blk B;
  if (slot < d.nb[part]) B = get_block<M>(d, part, slot);
  auto pval = [&](long long row) -> double {
    double pv = psrc[row * ld];
    if (h != 0.0) pv -= h * (qcoef * qc[row * ld] + gc[row * ld]);
    return pv;
  };
  double pu[UMAX], r[NRMAX], sres[UMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) pu[j] = 0.0;
  for (int j = 0; j < U; ++j) pu[j] = pval(j);
  double dx0_dv0[X * M::V0], dx0_dz[X * Z];
  if (has_blk) {
    double u[UMAX], z[Z], dzdu[Z * Z];
    for (int j = 0; j < U; ++j) u[j] = qc[(long long)j * ld];
    M::gen_z(u, z, dzdu);
    M::gen_x0_jac(z, dx0_dv0, dx0_dz);
    double m[X];
    if (B.ini) {
      double pv0[M::V0];
#pragma unroll
      for (int j = 0; j < M::V0; ++j) pv0[j] = pval(d.off_v0 + j);
      mv<X, M::V0>(dx0_dv0, pv0, m);
    } else {
#pragma unroll
      for (int i = 0; i < X; ++i) m[i] = 0.0;
    }
    for (int rr = 0; rr < B.nrows; ++rr) {
      double s = 0.0;
      for (int j = 0; j < U; ++j) s = fma(Ac[((long long)(B.row0 + rr) * U + j) * ld], pu[j], s);
      r[rr] = s;
    }
    for (int k = 0; k < B.n; ++k) {
      const long long g0 = (long long)(B.o + k) * d.S;
      double sk[X];
#pragma unroll
      for (int i = 0; i < X; ++i) sk[i] = 0.0;
      for (int t = 0; t < d.S; ++t) {
        double Kt[X * V], pv[V];
        ldcol<X * V>(Kc + (g0 + t) * X * V * ld, ld, Kt);
#pragma unroll
        for (int j = 0; j < V; ++j) pv[j] = pval(d.off_v + (g0 + t) * V + j);
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
          for (int j = 0; j < V; ++j) sk[i] = fma(Kt[i * V + j], pv[j], sk[i]);
      }
      double Ps[X * X], t1[X];
      ldcol<X * X>(Psibc + (long long)(B.o + k) * X * X * ld, ld, Ps);
      mv<X, X>(Ps, m, t1);
#pragma unroll
      for (int i = 0; i < X; ++i) m[i] = t1[i] + sk[i];
      if (k < B.ny) {
        double dh[X], xe[X];
        M::obs_grad(xe, dh);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < X; ++i) s = fma(dh[i], m[i], s);
        if (d.noisy) s += sigma_of<M>(d, u) * pval(d.off_n + B.o + k);
        r[k] += s;
      }
      if (k == B.n - 1 && B.nx > 0) {
#pragma unroll
        for (int i = 0; i < X; ++i) r[B.ny + i] += m[i];
      }
    }
  }
  inv_gram_block<M, NRMAX, UMAX, CPB>(d, B, has_blk, Ac, Lc, DinvAc, LCc, ld, r, sres, smem, nslot, slot, cl);
  if (has_blk) {
    double* alph = W.alpha + chain;
    double a0[X];
    alpha_block<M>(d, B, r, Psibc, nullptr, ld, alph, a0);
    for (int k = 0; k < B.n; ++k) {
      const long long g0 = (long long)(B.o + k) * d.S;
      double al[X];
      ldcol<X>(alph + (long long)(B.o + k) * X * ld, ld, al);
      for (int t = 0; t < d.S; ++t) {
        double Kt[X * V];
        ldcol<X * V>(Kc + (g0 + t) * X * V * ld, ld, Kt);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const long long row = d.off_v + (g0 + t) * V + j;
          double pv = pval(row);
#pragma unroll
          for (int i = 0; i < X; ++i) pv = fma(-Kt[i * V + j], al[i], pv);
          pdst[row * ld] = pv;
        }
      }
      if (d.noisy && k < B.ny) {
        double u[UMAX];
        for (int j = 0; j < U; ++j) u[j] = qc[(long long)j * ld];
        const long long row = d.off_n + B.o + k;
        pdst[row * ld] = pval(row) - sigma_of<M>(d, u) * r[k];
      }
    }
    if (d.noisy && !B.fin) {
      // noise variables of a non-final block's last observation are covered by k < ny above
    }
    if (B.ini) {
      double t0[M::V0];
      mtv<X, M::V0>(dx0_dv0, a0, t0);
#pragma unroll
      for (int j = 0; j < M::V0; ++j) pdst[(long long)(d.off_v0 + j) * ld] = pval(d.off_v0 + j) - t0[j];
      for (int j = 0; j < U; ++j) pdst[(long long)j * ld] = pu[j] - sres[j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_qn: on-device symmetric quasi-Newton projection loop (quasi_newton_projection :1009-1063 and
// its host wrapper :1323-1402) for a whole CTA of chains, masked per chain, no host round trips.
//   iterate:  c = constr(q) ; err = |c|_inf ; lam = G_prev^{-1} c ; q -= J_prev^T lam
// with q = qw - J_prev^T lam_tot never materialised inside the loop.
// mode 0 (forward):  on convergence write q_new -> slot(other).q and p(other) -= mom_coef * mu
// mode 1 (reverse):  compare q_back with q(cur) -> revd, no writes       (Mici reverse check)
// ------------------------------------------------------------------------------------------
template <class M, int CPB, int NRMAX, int UMAX>
MMD_D void dev_qn(const Dims& d, const Slots& S, const Work& W, const double* __restrict__ xobs,
                  const double* __restrict__ y, int part, int mode, double mom_coef, double ctol, double ptol,
                  double dtol, int max_iters) {
  MMD_THREAD_SETUP
  constexpr int X = M::X, V = M::V, Z = M::Z;
  constexpr int NTRI = NRMAX * (NRMAX + 1) / 2;
  const int U = d.U;
  const int cur = S.cur[chain];
  const int sl = mode ? 1 - cur : cur;  // linearisation used: forward -> cur (prev point), reverse -> new point
  const double* Kc = S.K + sl * S.s_K + chain;
  const double* Psibc = S.Psib + sl * S.s_Psib + chain;
  const double* Ac = S.A + sl * S.s_A + chain;
  const double* DinvAc = S.DinvA + sl * S.s_A + chain;
  const double* LCc = S.LC + sl * S.s_LC + chain;
  const double* Lc = S.L + sl * S.s_L + chain + (long long)slot * NTRI * ld;
  const double* qwc = W.qw + chain;
  const double* xoc = xobs + chain;
  double* alph = W.alpha + chain;
  double* alphi = W.alphi + chain;
  const bool in_blk = slot < d.nb[part];
  Blk B;
  if (in_blk) B = get_block<M>(d, part, slot);
  bool done = !act || (W.status[chain] != 0);
  int st = 0, it = 0;
  double u0[UMAX], stot[UMAX], lamtot[NRMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) { u0[j] = 0.0; stot[j] = 0.0; }
  for (int j = 0; j < U; ++j) u0[j] = qwc[(long long)j * ld];
#pragma unroll
  for (int r = 0; r < NRMAX; ++r) lamtot[r] = 0.0;
  double dx0_dv0[X * M::V0], dx0_dz[X * Z];
  double a0tot[X];
#pragma unroll
  for (int i = 0; i < X; ++i) a0tot[i] = 0.0;
  if (in_blk)
    for (int k = 0; k < B.n; ++k) {
      double zero[X];
#pragma unroll
      for (int i = 0; i < X; ++i) zero[i] = 0.0;
      stcol<X>(alph + (long long)(B.o + k) * X * ld, ld, zero);
    }
  double final_norm = 0.0;
  while (true) {
    if (__syncthreads_and(done ? 1 : 0)) break;
    const bool work = in_blk && !done;
    double crow[NRMAX], sres[UMAX], red[1];
    double z[Z], dzdu[Z * Z], u[UMAX];
    red[0] = 0.0;
    if (work) {
#pragma unroll
      for (int j = 0; j < UMAX; ++j) u[j] = u0[j] - stot[j];
      M::gen_z(u, z, dzdu);
      M::gen_x0_jac(z, dx0_dv0, dx0_dz);
      double x[X];
      if (B.ini) {
        double v0[M::V0], t0[M::V0];
        ldcol<M::V0>(qwc + (long long)d.off_v0 * ld, ld, v0);
        mtv<X, M::V0>(dx0_dv0, a0tot, t0);
#pragma unroll
        for (int j = 0; j < M::V0; ++j) v0[j] -= t0[j];
        M::gen_x0(z, v0, x);
      } else {
        ldcol<X>(xoc + (long long)(B.o - 1) * X * ld, ld, x);
      }
      constr_block<M, true>(d, B, z, sigma_of<M>(d, u), x, qwc, xoc, y, Kc, alph, ld, crow, nullptr);
      if (d.noisy) {
        // noise part of q: n_k = qw_n[k] - sigma_prev * lamtot[k]; constr_block read qw_n, correct here
        // (handled in noisy build; see k_qn_noisy_fixup)
      }
      double e = 0.0;
      for (int r = 0; r < B.nrows; ++r) {
        const double a = fabs(crow[r]);
        e = (a > e || a != a) ? a : e;
      }
      red[0] = e;
    }
    block_reduce<1, CPB, true>(red, smem, nslot, slot, cl);
    const double err = red[0];
    inv_gram_block<M, NRMAX, UMAX, CPB>(d, B, work, Ac, Lc, DinvAc, LCc, ld, crow, sres, smem, nslot, slot, cl);
    // norm of this iteration's update and (speculative) finalisation when the constraint is met
    const bool check = !done && (err < ctol);
    double nrm[1];
    nrm[0] = 0.0;
    double a0inc[X];
    if (work) {
#pragma unroll
      for (int r = 0; r < NRMAX; ++r)
        if (r < B.nrows) lamtot[r] += crow[r];
#pragma unroll
      for (int j = 0; j < UMAX; ++j) stot[j] += sres[j];
      double a0[X];
      alpha_block<M>(d, B, lamtot, Psibc, nullptr, ld, alph, a0);
#pragma unroll
      for (int i = 0; i < X; ++i) a0tot[i] = a0[i];
      if (check) {
        alpha_block<M>(d, B, crow, Psibc, nullptr, ld, alphi, a0inc);
        double nm = 0.0;
        double* qout = S.q + (1 - cur) * S.s_q + chain;
        double* pout = S.p + (1 - cur) * S.s_q + chain;
        const double* qref = S.q + cur * S.s_q + chain;
        double rv = 0.0;
        for (int k = 0; k < B.n; ++k) {
          const long long g0 = (long long)(B.o + k) * d.S;
          double al[X], ai[X];
          ldcol<X>(alph + (long long)(B.o + k) * X * ld, ld, al);
          ldcol<X>(alphi + (long long)(B.o + k) * X * ld, ld, ai);
          for (int t = 0; t < d.S; ++t) {
            double Kt[X * V];
            ldcol<X * V>(Kc + (g0 + t) * X * V * ld, ld, Kt);
#pragma unroll
            for (int j = 0; j < V; ++j) {
              double mu = 0.0, inc = 0.0;
#pragma unroll
              for (int i = 0; i < X; ++i) {
                mu = fma(Kt[i * V + j], al[i], mu);
                inc = fma(Kt[i * V + j], ai[i], inc);
              }
              nm = fmax(nm, fabs(inc));
              const long long row = d.off_v + (g0 + t) * V + j;
              const double qn = qwc[row * ld] - mu;
              if (mode == 0) {
                qout[row * ld] = qn;
                W.Yw[row * ld + chain] = mu;  // stash mu_v (momentum update applied once converged)
              } else {
                rv = fmax(rv, fabs(qn - qref[row * ld]));
              }
            }
          }
        }
        if (B.ini) {
          double t0[M::V0], ti[M::V0];
          mtv<X, M::V0>(dx0_dv0, a0tot, t0);
          mtv<X, M::V0>(dx0_dv0, a0inc, ti);
#pragma unroll
          for (int j = 0; j < M::V0; ++j) {
            nm = fmax(nm, fabs(ti[j]));
            const long long row = d.off_v0 + j;
            const double qn = qwc[row * ld] - t0[j];
            if (mode == 0) {
              qout[row * ld] = qn;
              W.Yw[row * ld + chain] = t0[j];
            } else {
              rv = fmax(rv, fabs(qn - qref[row * ld]));
            }
          }
          for (int j = 0; j < U; ++j) {
            nm = fmax(nm, fabs(sres[j]));
            const double qn = u0[j] - stot[j];
            if (mode == 0) {
              qout[(long long)j * ld] = qn;
              W.Yw[(long long)j * ld + chain] = stot[j];
            } else {
              rv = fmax(rv, fabs(qn - qref[(long long)j * ld]));
            }
          }
        }
        nrm[0] = nm;
        final_norm = rv;
        (void)pout;
      }
    }
    block_reduce<1, CPB, true>(nrm, smem, nslot, slot, cl);
    if (!done) {
      it += 1;
      const bool diverged = (err > dtol) || (err != err);
      const bool converged = check && (nrm[0] < ptol);
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
  // epilogue
  const bool live = act && (W.status[chain] == 0);
  if (mode == 0) {
    if (live && st == 0 && in_blk) {
      // p(other) -= mom_coef * mu   (state.mom -= dh2_flow_mom_dmom @ mu, :1388-1392)
      double* pout = S.p + (1 - cur) * S.s_q + chain;
      for (int k = 0; k < B.n; ++k) {
        const long long g0 = (long long)(B.o + k) * d.S;
        for (int t = 0; t < d.S * V; ++t) {
          const long long row = d.off_v + g0 * V + t;
          pout[row * ld] -= mom_coef * W.Yw[row * ld + chain];
        }
      }
      if (B.ini) {
        for (int j = 0; j < M::V0; ++j) {
          const long long row = d.off_v0 + j;
          pout[row * ld] -= mom_coef * W.Yw[row * ld + chain];
        }
        for (int j = 0; j < U; ++j) pout[(long long)j * ld] -= mom_coef * W.Yw[(long long)j * ld + chain];
      }
    }
  }
  double rr[1];
  rr[0] = final_norm;
  block_reduce<1, CPB, true>(rr, smem, nslot, slot, cl);
  if (slot == 0 && live) {
    W.iters[mode * ld + chain] = it;
    W.itsum[chain] += it;
    if (st) W.status[chain] |= st;
    if (mode == 1 && st == 0) W.revd[chain] = rr[0];
  }
}

// ------------------------------------------------------------------------------------------
// standalone launches of the phases (per-op API) and the fused persistent leapfrog kernel
// ------------------------------------------------------------------------------------------
template <class M, int CPB, int NRMAX, int RMAX, int UMAX, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_point(Dims d, Slots S, Work W, const double* __restrict__ xobs, const double* __restrict__ y, int part,
        int which, int with_grad) {
  dev_point<M, CPB, NRMAX, RMAX, UMAX>(d, S, W, xobs, y, part, which, with_grad);
}
template <class M, int CPB, int NRMAX, int UMAX, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_project(Dims d, Slots S, Work W, int part, int lin_sel, int src_sel, int dst_sel, double h, double qcoef) {
  dev_project<M, CPB, NRMAX, UMAX>(d, S, W, part, lin_sel, src_sel, dst_sel, h, qcoef);
}
template <class M, int CPB, int NRMAX, int UMAX, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_qn(Dims d, Slots S, Work W, const double* __restrict__ xobs, const double* __restrict__ y, int part,
     int mode, double mom_coef, double ctol, double ptol, double dtol, int max_iters) {
  dev_qn<M, CPB, NRMAX, UMAX>(d, S, W, xobs, y, part, mode, mom_coef, ctol, ptol, dtol, max_iters);
}

// h2_flow for this thread's rows: qw = q(q_sel) + dt * p(p_sel)   (mici_extensions.py:1222-1231)
template <class M, int CPB>
MMD_D void dev_flow(const Dims& d, const Slots& S, const Work& W, int part, int q_sel, int p_sel, double dt) {
  MMD_THREAD_SETUP
  if (slot >= d.nb[part]) return;
  const int cur = S.cur[chain];
  const double* qc = S.q + (q_sel ? 1 - cur : cur) * S.s_q + chain;
  const double* pc = S.p + (p_sel ? 1 - cur : cur) * S.s_q + chain;
  double* qw = W.qw + chain;
  const Blk B = get_block<M>(d, part, slot);
  const long long r0 = d.off_v + (long long)B.o * d.S * M::V, r1 = r0 + (long long)B.n * d.S * M::V;
  for (long long r = r0; r < r1; ++r) qw[r * ld] = fma(dt, pc[r * ld], qc[r * ld]);
  if (d.noisy)
    for (long long r = d.off_n + B.o; r < d.off_n + B.o + B.n; ++r) qw[r * ld] = fma(dt, pc[r * ld], qc[r * ld]);
  if (B.ini)
    for (long long r = 0; r < d.off_v; ++r) qw[r * ld] = fma(dt, pc[r * ld], qc[r * ld]);
}

// One (or n_steps) full ConstrainedLeapfrogIntegrator.step per chain in ONE launch: a CTA carries
// its CPB chains through every phase with CTA-local barriers only, so chains that need many
// projection iterations delay just their own small CTA while the other resident CTAs keep the SM busy
// (the per-chain iteration count is long-tailed: mean ~8, 1 % > 30, max_iters = 50).
// Step order: Mici ConstrainedLeapfrogIntegrator._step = A(dt/2) B(dt) A(dt/2), SURVEY.md 3.3.
template <class M, int CPB, int NRMAX, int RMAX, int UMAX, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_leapfrog(Dims d, Slots S, Work W, const double* __restrict__ xobs, const double* __restrict__ y, int part,
           double dt, double ctol, double ptol, double dtol, int max_iters, double rev_tol,
           long long* __restrict__ n_ok, int n_steps, int reset_status) {
  const int cl = threadIdx.x % CPB, slot = threadIdx.x / CPB;
  const int chain = blockIdx.x * CPB + cl;
  for (int s = 0; s < n_steps; ++s) {
    if (reset_status && slot == 0) W.status[chain] = 0;
    __syncthreads();
    dev_project<M, CPB, NRMAX, UMAX>(d, S, W, part, 0, 0, 1, 0.5 * dt, 1.0);
    __syncthreads();
    dev_flow<M, CPB>(d, S, W, part, 0, 1, dt);
    __syncthreads();
    dev_qn<M, CPB, NRMAX, UMAX>(d, S, W, xobs, y, part, 0, 1.0 / dt, ctol, ptol, dtol, max_iters);
    __syncthreads();
    dev_point<M, CPB, NRMAX, RMAX, UMAX>(d, S, W, xobs, y, part, 1, 1);
    __syncthreads();
    dev_project<M, CPB, NRMAX, UMAX>(d, S, W, part, 1, 1, 1, 0.0, 0.0);
    __syncthreads();
    dev_flow<M, CPB>(d, S, W, part, 1, 1, -dt);
    __syncthreads();
    dev_qn<M, CPB, NRMAX, UMAX>(d, S, W, xobs, y, part, 1, 0.0, ctol, ptol, dtol, max_iters);
    __syncthreads();
    dev_project<M, CPB, NRMAX, UMAX>(d, S, W, part, 1, 1, 1, 0.5 * dt, 1.0);
    __syncthreads();
    if (slot == 0 && chain < d.n_chains) {
      int st = W.status[chain];
      if (st == 0 && !(W.revd[chain] <= rev_tol)) {
        st |= ST_NONREV;
        W.status[chain] = st;
      }
      if (st == 0) {
        S.cur[chain] = 1 - S.cur[chain];
        n_ok[chain] += 1;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------
// h2_flow (mici_extensions.py:1222-1231), standard splitting: qw = q(sel_q) + dt * p(sel_p)
__global__ void k_flow(Dims d, Slots S, Work W, int q_sel, int p_sel, double dt) {
  const long long n = (long long)d.dim_q * d.ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(i % d.ld);
    const int cur = S.cur[chain];
    const int sq = q_sel ? 1 - cur : cur, sp = p_sel ? 1 - cur : cur;
    W.qw[i] = S.q[sq * S.s_q + i] + dt * S.p[sp * S.s_q + i];
  }
}

// commit / reject: successful chains flip to the new slot; reverse check (Mici
// ConstrainedLeapfrogIntegrator._step_b: reverse_check_norm(...) > reverse_check_tol)
__global__ void k_commit(Dims d, Slots S, Work W, double rev_tol, long long* __restrict__ n_ok) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= d.n_chains) return;
  int st = W.status[chain];
  if (st == 0 && !(W.revd[chain] <= rev_tol)) {
    st |= ST_NONREV;
    W.status[chain] = st;
  }
  if (st == 0) {
    S.cur[chain] = 1 - S.cur[chain];
    n_ok[chain] += 1;
  }
}

// Hamiltonian h = h1 + h2 (mici_extensions.py:1186-1202) for the current slot
template <int CPB>
__global__ void k_hamiltonian(Dims d, Slots S, Work W, int sel, double* __restrict__ hout) {
  MMD_THREAD_SETUP
  const int cur = S.cur[chain];
  const int sl = sel ? 1 - cur : cur;
  const double* qc = S.q + sl * S.s_q + chain;
  const double* pc = S.p + sl * S.s_q + chain;
  double acc[1];
  acc[0] = 0.0;
  for (long long r = slot; r < d.dim_q; r += nslot) {
    const double a = qc[r * ld], b = pc[r * ld];
    acc[0] += 0.5 * a * a + 0.5 * b * b;
  }
  block_reduce<1, CPB, false>(acc, smem, nslot, slot, cl);
  if (slot == 0 && act) hout[chain] = acc[0] + S.ldv[sl * S.s_ld + chain];
}

// [n_chains, rows] row-major (reference per-chain layout) <-> [rows][ld] structure of arrays
__global__ void k_aos_to_soa(const double* __restrict__ src, double* __restrict__ dst, int n_chains, int rows,
                             long long ld) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < n_chains && r < rows) ? src[(long long)c * rows + r] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < n_chains) dst[(long long)r * ld + c] = tile[threadIdx.x][j];
  }
}
__global__ void k_soa_to_aos(const double* __restrict__ src, double* __restrict__ dst, int n_chains, int rows,
                             long long ld) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < n_chains && r < rows) ? src[(long long)r * ld + c] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < n_chains) dst[(long long)c * rows + r] = tile[threadIdx.x][j];
  }
}

// full forward scan keeping the states at observation times (generate_x_obs_seq :384-397)
template <class M, int UMAX>
__global__ void k_gen_xobs(Dims d, const double* __restrict__ q, double* __restrict__ xobs) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= d.n_chains) return;
  constexpr int X = M::X, V = M::V, Z = M::Z;
  const long long ld = d.ld;
  const double* qc = q + chain;
  double u[UMAX], z[Z], dzdu[Z * Z], v0[M::V0], x[X];
  for (int j = 0; j < d.U; ++j) u[j] = qc[(long long)j * ld];
  M::gen_z(u, z, dzdu);
  ldcol<M::V0>(qc + (long long)d.off_v0 * ld, ld, v0);
  M::gen_x0(z, v0, x);
  for (int k = 0; k < d.T; ++k) {
    for (int t = 0; t < d.S; ++t) {
      double v[V], xn[X];
      ldcol<V>(qc + ((long long)d.off_v + ((long long)k * d.S + t) * V) * ld, ld, v);
      M::step(z, d.sd, x, v, xn);
#pragma unroll
      for (int i = 0; i < X; ++i) x[i] = xn[i];
    }
    stcol<X>(xobs + chain + (long long)k * X * ld, ld, x);
  }
}

// find_initial_state_by_linear_interpolation (mici_extensions.py:1479-1547), batched: one thread per
// (chain, observation interval).  Given u, v_0 and the states at observation times, solves per
// step for the noise vector that makes the discretised path interpolate linearly between them
// (forward_func is affine in v; d f / d v square and invertible: X == V).
template <class M, int UMAX>
__global__ void k_init_interp(Dims d, double* __restrict__ q, const double* __restrict__ xobs) {
  constexpr int X = M::X, V = M::V, Z = M::Z;
  static_assert(X == V, "linear-interpolation initialiser needs a square noise Jacobian");
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int chain = (int)(idx % d.ld);
  const int k = (int)(idx / d.ld);
  if (k >= d.T || chain >= d.n_chains) return;
  const long long ld = d.ld;
  double* qc = q + chain;
  double u[UMAX], z[Z], dzdu[Z * Z], xa[X], xb[X];
  for (int j = 0; j < d.U; ++j) u[j] = qc[(long long)j * ld];
  M::gen_z(u, z, dzdu);
  if (k == 0) {
    double v0[M::V0];
    ldcol<M::V0>(qc + (long long)d.off_v0 * ld, ld, v0);
    M::gen_x0(z, v0, xa);
  } else {
    ldcol<X>(xobs + chain + (long long)(k - 1) * X * ld, ld, xa);
  }
  ldcol<X>(xobs + chain + (long long)k * X * ld, ld, xb);
  double dx[X];
#pragma unroll
  for (int i = 0; i < X; ++i) dx[i] = (xb[i] - xa[i]) / d.S;
  for (int s = 0; s < d.S; ++s) {
    double x[X], vz[V], m[X], Bm[X * V], rhs[X];
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = xa[i] + s * dx[i];
#pragma unroll
    for (int j = 0; j < V; ++j) vz[j] = 0.0;
    M::step(z, d.sd, x, vz, m);
    M::jac_v(z, d.sd, x, vz, Bm);
#pragma unroll
    for (int i = 0; i < X; ++i) rhs[i] = dx[i] - (m[i] - x[i]);
    // Gaussian elimination with partial pivoting on the X x X system Bm v = rhs
    for (int c = 0; c < X; ++c) {
      int piv = c;
      for (int r = c + 1; r < X; ++r)
        if (fabs(Bm[r * V + c]) > fabs(Bm[piv * V + c])) piv = r;
      if (piv != c) {
        for (int j = 0; j < V; ++j) { double t = Bm[c * V + j]; Bm[c * V + j] = Bm[piv * V + j]; Bm[piv * V + j] = t; }
        double t = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = t;
      }
      for (int r = c + 1; r < X; ++r) {
        const double f = Bm[r * V + c] / Bm[c * V + c];
        for (int j = c; j < V; ++j) Bm[r * V + j] -= f * Bm[c * V + j];
        rhs[r] -= f * rhs[c];
      }
    }
    double v[V];
    for (int r = X - 1; r >= 0; --r) {
      double t = rhs[r];
      for (int j = r + 1; j < V; ++j) t -= Bm[r * V + j] * v[j];
      v[r] = t / Bm[r * V + r];
    }
    stcol<V>(qc + ((long long)d.off_v + ((long long)k * d.S + s) * V) * ld, ld, v);
  }
  if (d.noisy) qc[((long long)d.off_n + k) * ld] = 0.0;
}

// Metropolis accept step of a static-trajectory constrained HMC transition, on device.
// accept iff the trajectory finished without an integrator error, the final Hamiltonian is finite
// and log(uniform) < h0 - h1 (Mici MetropolisStaticIntegrationTransition semantics; IntegratorError
// -> reject).  acc_prob is the `accept_stat` statistic min(1, exp(h0 - h1)).
__global__ void k_decide(Dims d, Work W, const double* __restrict__ h0, const double* __restrict__ h1,
                         uint64_t seed, uint64_t offset, int* __restrict__ accepted,
                         double* __restrict__ acc_prob) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= d.n_chains) return;
  double n0, n1;
  (void)n1;
  uint32_t c[4] = {(uint32_t)chain, 0u, (uint32_t)offset, (uint32_t)(offset >> 32)};
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
  n0 = ap;
  accepted[chain] = acc;
  acc_prob[chain] = n0;
}
__global__ void k_restore(Dims d, Slots S, const int* __restrict__ accepted, const double* __restrict__ qsave) {
  const long long n = (long long)d.dim_q * d.ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(i % d.ld);
    if (chain < d.n_chains && !accepted[chain]) S.q[S.cur[chain] * S.s_q + i] = qsave[i];
  }
}

}  // namespace mmd
