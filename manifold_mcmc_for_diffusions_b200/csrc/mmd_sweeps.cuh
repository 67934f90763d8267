// Time-stepping sweeps and per-block solves shared by the CHMC phases (mmd_kernels_main.cuh).
// Everything here is per thread = per (chain, observation block); see mmd_kernels.cuh for the layout.
#pragma once
#include "mmd_kernels.cuh"

namespace mmd {

// per-thread parameters of the chain at some value of u
template <class M, int UMAX>
struct ChainPar {
  double u[UMAX];
  double z[M::Z];
  double dzdu[M::Z * M::Z];
  typename M::Coef C;
  double sigy;
};
template <class M, int UMAX>
MMD_D void make_par(const Dims& d, const double* u, ChainPar<M, UMAX>& P) {
#pragma unroll
  for (int j = 0; j < UMAX; ++j) P.u[j] = u[j];
  M::gen_z(P.u, P.z, P.dzdu);
  M::make_coef(P.z, d.sd, P.C);
  P.sigy = sigma_of<M>(d, P.u);
}

// state at the start of the thread's block: generate_x_0(z, v_0) for the first block, otherwise the
// conditioned state at the end of the previous block (partition_into_subseqs, :413-471)
template <class M>
MMD_D void block_start(const Dims& d, const Blk& B, const double* z, const double* v0, const double* xoc, int cpb,
                       double* x) {
  if (B.ini) {
    M::gen_x0(z, v0, x);
  } else {
    ldcol<M::X>(xoc + (B.o - 1) * M::X * cpb, cpb, x);
  }
}

// ------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS): the sweeps prefetch the rows of future time steps into a
// thread-private shared-memory ring, so the recursion never waits on HBM and no registers are spent
// on data in flight
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
MMD_D void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
MMD_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
MMD_D void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#endif

// ------------------------------------------------------------------------------------------
// forward constraint sweep for one block:  c_b(q)   (generate_y_bar + constr, :399-411, :473-519).
// With WITH_K the position is the quasi-Newton parametrisation q = qw - J_lin^T lambda_tot, never
// materialised:   v_t = qw_v[t] - K_t^T alpha_k ,  n_k = qw_n[k] - sigma_lin * lambda_tot[k].
// The rows of step s + PF are fetched by cp.async into a shared-memory ring while step s is computed
// (the loads do not depend on the recursion; the sweep would otherwise be bound by global-load
// latency).  ring: thread-private, word w of ring slot i at ring[(i * NW + w) * NT].
// crow / lamtot are thread-private columns in shared memory: element r at [r * NT].
// ------------------------------------------------------------------------------------------
template <class M, bool WITH_K>
MMD_D void constr_sweep(const Dims& d, const Blk& B, const typename M::Coef& C, double sigma_y, double sigma_lin,
                        const double* xstart, const double* __restrict__ vb, const double* __restrict__ nzb,
                        const double* __restrict__ xoc, const double* __restrict__ y,
                        const double* __restrict__ Kb, const double* __restrict__ alph, const double* lamtot,
                        int nta, int cpb, int NT, double* crow, double* ring, double* xend_out) {
  constexpr int X = M::X, V = M::V, XV = M::X * M::V;
  constexpr int PF = MMD_PREFETCH_STEPS, NSL = PF + 1, NW = V + (WITH_K ? XV : 0);
  const int ns = B.n * d.S;
  double x[X];
#pragma unroll
  for (int i = 0; i < X; ++i) x[i] = xstart[i];
  auto issue = [&](int s, int slot) {
    double* dst = ring + slot * NW * NT;
#pragma unroll
    for (int j = 0; j < V; ++j) cp_async8(dst + j * NT, vb + (s * V + j) * nta);
    if (WITH_K) {
#pragma unroll
      for (int j = 0; j < XV; ++j) cp_async8(dst + (V + j) * NT, Kb + (s * XV + j) * nta);
    }
  };
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    if (i < ns) issue(i, i);
    cp_async_commit();
  }
  double al[X];
  if (WITH_K) ldcol<X>(alph, nta, al);
  int k = 0, t = 0, slot = 0, fill = PF;
  for (int s = 0; s < ns; ++s) {
    cp_async_wait<PF - 1>();
    const double* src = ring + slot * NW * NT;
    double v[V], xn[X];
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = src[j * NT];
    if (WITH_K) {
      double Kt[XV];
#pragma unroll
      for (int j = 0; j < XV; ++j) Kt[j] = src[(V + j) * NT];
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int a = 0; a < X; ++a) v[j] = fma(-Kt[a * V + j], al[a], v[j]);
    }
    if (s + PF < ns) issue(s + PF, fill);  // the slot consumed one step ago
    cp_async_commit();
    slot = (slot + 1 == NSL) ? 0 : slot + 1;
    fill = (fill + 1 == NSL) ? 0 : fill + 1;
    M::step(C, x, v, xn);
#pragma unroll
    for (int a = 0; a < X; ++a) x[a] = xn[a];
    if (++t == d.S) {  // end of observation interval k
      t = 0;
      if (xend_out) stcol<X>(xend_out + k * X * nta, nta, x);
      if (k < B.ny) {
        double cy = M::obs(x) - y[B.o + k];
        if (d.noisy) {
          double nk = nzb[k * nta];
          if (WITH_K) nk = fma(-sigma_lin, lamtot[k * NT], nk);
          cy = fma(sigma_y, nk, cy);
        }
        crow[k * NT] = cy;
      }
      if (k == B.n - 1 && B.nx > 0) {
        double xo[X];
        ldcol<X>(xoc + (B.o + k) * X * cpb, cpb, xo);
#pragma unroll
        for (int a = 0; a < X; ++a) crow[(B.ny + a) * NT] = x[a] - xo[a];
      }
      ++k;
      if (WITH_K && k < B.n) ldcol<X>(alph + k * X * nta, nta, al);
    }
  }
  cp_async_wait<0>();
}

// obs-level backward recursion: alpha_k = H_k^T lambda_k + Psib_{k+1}^T alpha_{k+1}  (J^T lambda in
// compressed form, rmult_by_jacob_constr :879-913).  Writes alpha for the block's intervals and
// returns alpha at the block start (needed for the v_0 columns of block 0).  lam: the block's rows in
// registers.  Fully unrolled over RMAX intervals so that the Psib loads (independent of the
// recursion) are issued together instead of one L2 round trip per interval.
template <class M, int NRMAX, int RMAX>
MMD_D void alpha_block(const Blk& B, const double* lam, const double* __restrict__ Psibc,
                       const double* __restrict__ xendc, int nta, double* alph_out, double* alpha_start) {
  constexpr int X = M::X;
  double al[X];
#pragma unroll
  for (int i = 0; i < X; ++i) al[i] = 0.0;
#pragma unroll
  for (int k = RMAX - 1; k >= 0; --k) {
    if (k < B.n) {
      if (k < B.n - 1) {
        double Ps[X * X], t[X];
        ldcol<X * X>(Psibc + (k + 1) * X * X * nta, nta, Ps);
        mtv<X, X>(Ps, al, t);
#pragma unroll
        for (int i = 0; i < X; ++i) al[i] = t[i];
      }
      if (k < B.ny) {
        double dh[X], xe[X];
        if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
        M::obs_grad(xe, dh);
        const double lk = lam[k < NRMAX ? k : 0];
#pragma unroll
        for (int i = 0; i < X; ++i) al[i] = fma(dh[i], lk, al[i]);
      }
      if (k == B.n - 1 && B.nx > 0) {
        // rows ny .. ny + X - 1 hold the multipliers of the conditioned full state
#pragma unroll
        for (int r = 0; r < NRMAX; ++r)
#pragma unroll
          for (int i = 0; i < X; ++i)
            if (r == B.ny + i) al[i] += lam[r];
      }
      stcol<X>(alph_out + k * X * nta, nta, al);
    }
  }
  {
    double Ps[X * X];
    ldcol<X * X>(Psibc, nta, Ps);
    mtv<X, X>(Ps, al, alpha_start);
  }
}

// Woodbury solve G^{-1} r for this thread's block (lmult_by_inv_gram :915-942):
//   t_b = D_b^{-1} r_b ; s = C^{-1} sum_b A_b^T t_b ; lam_b = t_b - (D_b^{-1} A_b) s
// `r` (registers, NRMAX entries, entries >= nrows ignored) is overwritten by lam_b; s (= u-part of
// J^T G^{-1} r) is returned in `s_out`.  `extra_max` rides along the same cross-block reduction as a
// maximum (the solver's |c|_inf).  Everything is unrolled over NRMAX x UMAX with guards so that the
// factor loads (L, A, D^{-1}A: thread-private, L2-resident) are issued up front, not one per
// dependent multiply-add.
template <class M, int NRMAX, int UMAX>
MMD_D void inv_gram_block(const Dims& d, const Blk& B, bool has_blk, const double* __restrict__ Ac,
                          const double* __restrict__ Lc, const double* __restrict__ DinvAc,
                          const double* __restrict__ LCc, double* r, double* s_out, double* extra_max,
                          double* smem_red, const Tid& t) {
  const int U = d.U, nta = t.nta;
  const int n = has_blk ? B.nrows : 0;
  double g[UMAX + 1];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) g[j] = 0.0;
  g[UMAX] = extra_max ? *extra_max : 0.0;
  if (has_blk) {
    // forward / backward substitution with the packed factor (diagonal stored inverted)
#pragma unroll
    for (int i = 0; i < NRMAX; ++i) {
      if (i < n) {
        double s = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-Lc[tri(i, k) * nta], r[k], s);
        r[i] = s * Lc[tri(i, i) * nta];
      }
    }
#pragma unroll
    for (int i = NRMAX - 1; i >= 0; --i) {
      if (i < n) {
        double s = r[i];
#pragma unroll
        for (int k = i + 1; k < NRMAX; ++k)
          if (k < n) s = fma(-Lc[tri(k, i) * nta], r[k], s);
        r[i] = s * Lc[tri(i, i) * nta];
      }
    }
#pragma unroll
    for (int i = 0; i < NRMAX; ++i) {
      if (i < n) {
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (j < U) g[j] = fma(Ac[(i * U + j) * nta], r[i], g[j]);
      }
    }
  }
  block_reduce<UMAX, 1>(g, smem_red, t);
  if (extra_max) *extra_max = g[UMAX];
  double LCm[UMAX * (UMAX + 1) / 2];
#pragma unroll
  for (int i = 0; i < UMAX * (UMAX + 1) / 2; ++i) LCm[i] = (i < U * (U + 1) / 2) ? LCc[i * t.cpb] : 0.0;
  chol_solve_invdiag_fixed<UMAX>(LCm, U, g);
#pragma unroll
  for (int j = 0; j < UMAX; ++j) s_out[j] = g[j];
  if (has_blk) {
#pragma unroll
    for (int i = 0; i < NRMAX; ++i) {
      if (i < n) {
        double ti = r[i];
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (j < U) ti = fma(-DinvAc[(i * U + j) * nta], g[j], ti);
        r[i] = ti;
      }
    }
  }
}

}  // namespace mmd
