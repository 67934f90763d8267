// Hand-written FP64 CUDA kernels (sm_100a) for the constrained-HMC hot path of
// sde/mici_extensions.py (reference), batched over chains.
//
// Mapping: one thread = one (chain, observation block); a warp = 32 (or CPB) consecutive chains
// of the same block, so every global access is a coalesced row of a structure-of-arrays
// [row][chain] matrix.  A CTA owns CPB chains x all of their blocks, so the cross-block
// reductions (Woodbury capacitance matrix C, u-part of J^T lambda, norms) are shared-memory
// reductions and no kernel ever needs a grid-wide sync or a host round trip.
//
// The block Jacobian dc/dv (reference: jacob_constr_blocks, mici_extensions.py:521-624, dense
// [rows, R*S*dim_v] per block) is never materialised.  For step t inside observation interval k
//     d c_r / d v_t = H_r Phi(t_r, t_k) K_t ,     K_t = Psi_{t+1} B_t ,  Psi_{t+1} = Phi(t_k, t+1)
// so the kernels keep the "compressed Jacobian" K_t (X x V per step) plus the per-interval
// transition matrices Psib_k = Phi(t_k, t_{k-1}); everything else is small per-observation algebra.
// See DESIGN.md for the derivation (incl. the second-order adjoint used for grad log det).
#pragma once
#include "mmd_common.cuh"
#include <stdint.h>

#ifndef MMD_PREFETCH_STEPS
#define MMD_PREFETCH_STEPS 5
#endif

namespace mmd {

// ------------------------------------------------------------------------------------------
// problem description (host fills, passed by value to kernels)
// ------------------------------------------------------------------------------------------
struct Dims {
  int T, S, R;        // num_obs, num_steps_per_obs, num_obs_per_subseq  (mici_extensions.py:317-351)
  int U;              // dim_u
  int noisy;          // 0 noiseless, 1 fixed sigma, 2 sigma = exp(u[Z])   (generate_sigma, :353-358)
  int gaussian;       // use_gaussian_splitting (:303)
  double sigma_fixed;
  int dim_q;
  int nb[2];          // number of blocks per partition
  int init_size[2];   // obs in first block
  int fin_size[2];    // obs in last block
  int n_c[2];         // constraint rows per partition
  int num_partition;
  int n_chains, ld;   // chains and leading dimension (>= n_chains, multiple of 32)
  double delta, sd;   // step delta = obs_interval / S and sqrt(delta)
  int off_v0, off_v, off_n;  // row offsets into q: [u | v_0 | v_seq | n]  (:476-484)
};

// Everything Mici caches at a position (jacob_constr_blocks, chol_gram_blocks, log_det_sqrt_gram,
// grad_log_det_sqrt_gram; mici_extensions.py:1151-1184) in compressed form, double-buffered so a
// failed step leaves the chain where it was.  Arrays are [2][rows][ld]; `cur[chain]` selects.
struct Slots {
  double* q;       // [dim_q]
  double* p;       // [dim_q]
  double* K;       // [T*S*X*V]
  double* Psib;    // [T*X*X]
  double* A;       // [NCMAX*U]      dc/du rows
  double* L;       // [NBMAX*NRTRI]  packed lower Cholesky factors of D_b
  double* DinvA;   // [NCMAX*U]
  double* LC;      // [U(U+1)/2]     packed lower Cholesky factor of C
  double* gradld;  // [dim_q]
  double* ldv;     // [1]
  long long s_q, s_K, s_Psib, s_A, s_L, s_LC, s_ld;  // slot strides in elements
  int* cur;        // [ld]
};

struct Work {
  double* xs;     // [T*S*X]   trajectory x_t at the point being linearised
  double* Yw;     // [T*S*X*X] forward tangent accumulator of the second-order sweep
  double* Qk;     // [T*X*X]
  double* Zt;     // [T*X*Z]
  double* Mk;     // [T*X*X]
  double* LamZ;   // [T*Z*X]
  double* Yb;     // [T*X*X]
  double* alpha;  // [T*X]
  double* alphi;  // [T*X]
  double* qw;     // [dim_q]   work position for the projection solves
  double* cvec;   // [NCMAX]
  int* status;    // [ld]  bit 1 not converged, 2 diverged, 4 non-reversible, 8 non-finite H
  int* iters;     // [2][ld] projection iterations (forward, reverse) of the last step
  double* revd;   // [ld] reverse-check distance of the last step
  double* hval;   // [ld]
  long long* itsum;  // [ld] total quasi-Newton iterations executed (both directions)
};

enum : int { ST_NOTCONV = 1, ST_DIVERGED = 2, ST_NONREV = 4, ST_NONFINITE = 8 };

// ------------------------------------------------------------------------------------------
// tiny dense helpers (row-major, fully unrolled)
// ------------------------------------------------------------------------------------------
template <int R, int C, int K>
MMD_D void mm(const double* A, const double* B, double* O) {  // O[RxC] = A[RxK] B[KxC]
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i * K + k], B[k * C + j], s);
      O[i * C + j] = s;
    }
}
template <int R, int C, int K>
MMD_D void mm_acc(const double* A, const double* B, double* O) {  // O += A B
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = O[i * C + j];
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i * K + k], B[k * C + j], s);
      O[i * C + j] = s;
    }
}
template <int R, int C, int K>
MMD_D void mtm(const double* A, const double* B, double* O) {  // O[RxC] = A[KxR]^T B[KxC]
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[k * R + i], B[k * C + j], s);
      O[i * C + j] = s;
    }
}
template <int R, int C, int K>
MMD_D void mmt(const double* A, const double* B, double* O) {  // O[RxC] = A[RxK] B[CxK]^T
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i * K + k], B[j * K + k], s);
      O[i * C + j] = s;
    }
}
template <int R, int C>
MMD_D void mtv(const double* A, const double* x, double* y) {  // y[C] = A[RxC]^T x[R]
#pragma unroll
  for (int j = 0; j < C; ++j) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < R; ++i) s = fma(A[i * C + j], x[i], s);
    y[j] = s;
  }
}
template <int R, int C>
MMD_D void mv(const double* A, const double* x, double* y) {  // y[R] = A[RxC] x[C]
#pragma unroll
  for (int i = 0; i < R; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < C; ++j) s = fma(A[i * C + j], x[j], s);
    y[i] = s;
  }
}
template <int N>
MMD_D void ldcol(const double* g, long long ld, double* r) {  // gather N consecutive rows of one chain
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = g[i * ld];
}
template <int N>
MMD_D void stcol(double* g, long long ld, const double* r) {
#pragma unroll
  for (int i = 0; i < N; ++i) g[i * ld] = r[i];
}
MMD_D int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower index, j <= i

// ------------------------------------------------------------------------------------------
// block geometry (restates the shape logic at mici_extensions.py:321-351 in closed form)
// ------------------------------------------------------------------------------------------
struct Blk {
  int o, n, ny, nx, nrows, row0;
  bool ini, fin;
};
template <class M>
MMD_D Blk get_block(const Dims& d, int part, int b) {
  Blk B;
  const int nb = d.nb[part];
  B.ini = (b == 0);
  B.fin = (b == nb - 1);
  const int nz = d.noisy ? 1 : 0;
  if (nb == 1) {
    B.o = 0;
    B.n = d.T;
    B.row0 = 0;
  } else {
    const int i0 = d.init_size[part];
    B.o = B.ini ? 0 : i0 + (b - 1) * d.R;
    B.n = B.ini ? i0 : (B.fin ? d.fin_size[part] : d.R);
    const int r0 = i0 - 1 + nz + M::X, rm = d.R - 1 + nz + M::X;
    B.row0 = B.ini ? 0 : r0 + (b - 1) * rm;
  }
  B.ny = (B.fin || d.noisy) ? B.n : B.n - 1;
  B.nx = B.fin ? 0 : M::X;
  B.nrows = B.ny + B.nx;
  return B;
}

// cross-block (same chain) reductions through shared memory.  All threads of the CTA must call.
// vals[NV] is replaced by the sum over block slots (deterministic ascending order).
template <int NV, int CPB, bool MAXRED>
MMD_D void block_reduce(double* vals, double* smem, int nslot, int slot, int cl) {
#pragma unroll
  for (int i = 0; i < NV; ++i) smem[(slot * NV + i) * CPB + cl] = vals[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = MAXRED ? 0.0 : 0.0;
    for (int sl = 0; sl < nslot; ++sl) {
      const double v = smem[(sl * NV + i) * CPB + cl];
      if (MAXRED) {
        s = (v > s || v != v) ? v : s;  // NaN-propagating max of non-negative values
      } else {
        s += v;
      }
    }
    vals[i] = s;
  }
  __syncthreads();
}

// in-place packed Cholesky (lower) of an n x n SPD matrix
template <int NRMAX>
MMD_D void chol_packed(double* Dm, int n) {
  for (int j = 0; j < n; ++j) {
    double s = Dm[tri(j, j)];
    for (int k = 0; k < j; ++k) s -= Dm[tri(j, k)] * Dm[tri(j, k)];
    const double ljj = sqrt(s);
    Dm[tri(j, j)] = ljj;
    const double inv = 1.0 / ljj;
    for (int i = j + 1; i < n; ++i) {
      double t = Dm[tri(i, j)];
      for (int k = 0; k < j; ++k) t -= Dm[tri(i, k)] * Dm[tri(j, k)];
      Dm[tri(i, j)] = t * inv;
    }
  }
}
// solve L L^T x = b in place (packed lower L)
MMD_D void chol_solve_packed(const double* Lm, int n, double* x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= Lm[tri(i, k)] * x[k];
    x[i] = s / Lm[tri(i, i)];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= Lm[tri(k, i)] * x[k];
    x[i] = s / Lm[tri(i, i)];
  }
}

// per-thread context
template <class M>
struct Ctx {
  int chain, cl, slot, nslot;
  bool active;
  long long ld;
};

// ------------------------------------------------------------------------------------------
// forward constraint sweep for one block:  c_b(q)   (generate_y_bar + constr, :399-411, :473-519)
// `vcol`/`alph` implement the quasi-Newton parametrisation q = qw - J_prev^T lambda_tot without
// materialising q:   v_t = qw_v[t] - K_t^T alpha_k.
// ------------------------------------------------------------------------------------------
template <class M, bool WITH_K>
MMD_D void constr_block(const Dims& d, const Blk& B, const double* z, double sigma_y, const double* xstart,
                        const double* qc, const double* xobs, const double* y, const double* Kc,
                        const double* alph, long long ld, double* crow, double* xend_out) {
  constexpr int X = M::X, V = M::V;
  double x[X];
#pragma unroll
  for (int i = 0; i < X; ++i) x[i] = xstart[i];
  for (int k = 0; k < B.n; ++k) {
    const long long g0 = (long long)(B.o + k) * d.S;
    const double* vp = qc + ((long long)d.off_v + g0 * V) * ld;
    double al[X];
    if (WITH_K) ldcol<X>(alph + (long long)(B.o + k) * X * ld, ld, al);
    const double* Kp = WITH_K ? Kc + g0 * X * V * ld : nullptr;
    // loads do not depend on the recursion: fetch PF steps' worth of rows first, then run the steps
    // (register-level software pipelining; the sweep is otherwise bound by global-load latency)
    constexpr int PF = MMD_PREFETCH_STEPS;
    int t = 0;
    for (; t + PF <= d.S; t += PF) {
      double vb[PF * V], Kb[WITH_K ? PF * X * V : 1];
#pragma unroll
      for (int i = 0; i < PF * V; ++i) vb[i] = vp[((long long)t * V + i) * ld];
      if (WITH_K) {
#pragma unroll
        for (int i = 0; i < PF * X * V; ++i) Kb[i] = Kp[((long long)t * X * V + i) * ld];
      }
#pragma unroll
      for (int g = 0; g < PF; ++g) {
        double v[V], xn[X];
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = vb[g * V + j];
        if (WITH_K) {
#pragma unroll
          for (int j = 0; j < V; ++j)
#pragma unroll
            for (int i = 0; i < X; ++i) v[j] = fma(-Kb[g * X * V + i * V + j], al[i], v[j]);
        }
        M::step(z, d.sd, x, v, xn);
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = xn[i];
      }
    }
    for (; t < d.S; ++t) {
      double v[V];
      ldcol<V>(vp + (long long)t * V * ld, ld, v);
      if (WITH_K) {
        double Kt[X * V];
        ldcol<X * V>(Kp + (long long)t * X * V * ld, ld, Kt);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int i = 0; i < X; ++i) v[j] = fma(-Kt[i * V + j], al[i], v[j]);
      }
      double xn[X];
      M::step(z, d.sd, x, v, xn);
#pragma unroll
      for (int i = 0; i < X; ++i) x[i] = xn[i];
    }
    if (xend_out) stcol<X>(xend_out + (long long)(B.o + k) * X * ld, ld, x);
    if (k < B.ny) {
      double cy = M::obs(x) - y[B.o + k];
      if (d.noisy) cy += sigma_y * qc[((long long)d.off_n + B.o + k) * ld];
      crow[k] = cy;
    }
    if (k == B.n - 1 && B.nx > 0) {
      double xo[X];
      ldcol<X>(xobs + (long long)(B.o + k) * X * ld, ld, xo);
#pragma unroll
      for (int i = 0; i < X; ++i) crow[B.ny + i] = x[i] - xo[i];
    }
  }
}

template <class M>
MMD_D double sigma_of(const Dims& d, const double* u) {
  if (d.noisy == 1) return d.sigma_fixed;
  if (d.noisy == 2) return exp(u[M::Z]);
  return 0.0;
}

// obs-level backward recursion: alpha_k = H_k^T lambda_k + Psib_{k+1}^T alpha_{k+1}  (J^T lambda in
// compressed form, rmult_by_jacob_constr :879-913).  Writes alpha for the block's intervals and
// returns alpha at the block start (needed for the v_0 columns of block 0).
template <class M>
MMD_D void alpha_block(const Dims& d, const Blk& B, const double* lam, const double* Psibc,
                       const double* xendc, long long ld, double* alph_out, double* alpha_start) {
  constexpr int X = M::X;
  double al[X];
#pragma unroll
  for (int i = 0; i < X; ++i) al[i] = 0.0;
  for (int k = B.n - 1; k >= 0; --k) {
    if (k < B.n - 1) {
      double Ps[X * X], t[X];
      ldcol<X * X>(Psibc + (long long)(B.o + k + 1) * X * X * ld, ld, Ps);
      mtv<X, X>(Ps, al, t);
#pragma unroll
      for (int i = 0; i < X; ++i) al[i] = t[i];
    }
    if (k < B.ny) {
      double dh[X], xe[X];
      if (!M::OBS_LINEAR) ldcol<X>(xendc + (long long)(B.o + k) * X * ld, ld, xe);
      M::obs_grad(xe, dh);
#pragma unroll
      for (int i = 0; i < X; ++i) al[i] = fma(dh[i], lam[k], al[i]);
    }
    if (k == B.n - 1 && B.nx > 0) {
#pragma unroll
      for (int i = 0; i < X; ++i) al[i] += lam[B.ny + i];
    }
    stcol<X>(alph_out + (long long)(B.o + k) * X * ld, ld, al);
  }
  {
    double Ps[X * X];
    ldcol<X * X>(Psibc + (long long)B.o * X * X * ld, ld, Ps);
    mtv<X, X>(Ps, al, alpha_start);
  }
}

// Woodbury solve G^{-1} r for this thread's block (lmult_by_inv_gram :915-942):
//   t_b = D_b^{-1} r_b ; s = C^{-1} sum_b A_b^T t_b ; lam_b = t_b - (D_b^{-1} A_b) s
// `r` is overwritten by lam_b; returns s (= u-part of J^T G^{-1} r) in `s_out`.
template <class M, int NRMAX, int UMAX, int CPB>
MMD_D void inv_gram_block(const Dims& d, const Blk& B, bool has_blk, const double* Ac, const double* Lc,
                          const double* DinvAc, const double* LCc, long long ld, double* r, double* s_out,
                          double* smem, int nslot, int slot, int cl) {
  const int U = d.U;
  double g[UMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) g[j] = 0.0;
  if (has_blk) {
    double Lm[NRMAX * (NRMAX + 1) / 2];
    for (int i = 0; i < B.nrows * (B.nrows + 1) / 2; ++i) Lm[i] = Lc[(long long)i * ld];
    chol_solve_packed(Lm, B.nrows, r);
    for (int i = 0; i < B.nrows; ++i)
      for (int j = 0; j < U; ++j) g[j] = fma(Ac[((long long)(B.row0 + i) * U + j) * ld], r[i], g[j]);
  }
  block_reduce<UMAX, CPB, false>(g, smem, nslot, slot, cl);
  double LCm[UMAX * (UMAX + 1) / 2];
  for (int i = 0; i < U * (U + 1) / 2; ++i) LCm[i] = LCc[(long long)i * ld];
  chol_solve_packed(LCm, U, g);
#pragma unroll
  for (int j = 0; j < UMAX; ++j) s_out[j] = g[j];
  if (has_blk) {
    for (int i = 0; i < B.nrows; ++i) {
      double t = r[i];
      for (int j = 0; j < U; ++j) t = fma(-DinvAc[((long long)(B.row0 + i) * U + j) * ld], g[j], t);
      r[i] = t;
    }
  }
}

}  // namespace mmd
