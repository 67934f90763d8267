// Time-stepping sweeps and per-block solves shared by the CHMC phases (mmd_kernels_main.cuh).
// Everything here is per thread = per (chain, observation block); see mmd_kernels.cuh for the layout.
#pragma once
#include "mmd_kernels.cuh"

// the per-iteration block solves: fully unrolled (factor loads issued up front: lowest latency for a lone
// CTA) or rolled (fewer registers: higher throughput with all SMs busy)
#if defined(MMD_ROLLED_SOLVES)
#define MMD_SOLVE_UNROLL _Pragma("unroll 1")
#else
#define MMD_SOLVE_UNROLL _Pragma("unroll")
#endif

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
// asynchronous global -> shared copies (LDGSTS): the sweeps prefetch the records of future time steps
// into a thread-private shared-memory ring, so the recursion never waits on HBM and no registers are
// spent on data in flight.  Shared addresses are 32-bit window offsets computed once per sweep.
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
MMD_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// L2 residency hints.  The per-iteration block factors (L, A, D^-1 A, Psi_k, kappa_k, alpha_k: ~100 doubles per
// thread) are re-read by every solver iteration while the streams (work position, compressed Jacobian: 6 KB per
// thread and sweep) flush the L2 in between; evict_last keeps the factors resident, evict_first marks the streams
// as the first candidates for replacement.
MMD_D unsigned long long l2_policy_keep() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
MMD_D unsigned long long l2_policy_stream() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
MMD_D double ldg_keep(const double* g) {
#if defined(MMD_HINT_KEEP)
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(g), "l"(l2_policy_keep()));
  return v;
#else
  return *g;
#endif
}
// The 16-byte copies allocate in L1 (.ca, not the L1-bypassing .cg): a thread's X*V-double record is fetched in
// 16-byte pieces at a lane stride of X*V*8 bytes, so every copy instruction of a warp touches half of each 32-byte
// sector and the next one the other half.  With .cg both requests go to L2 and the sweeps saturate the L2 request
// path at 4.5 TB/s of useful bytes; with .ca the second one hits L1: 6.9 TB/s, next to the 7.2 TB/s of CTA-wide
// cp.async.bulk copies (tools/stream_bench.cu, profiles/r2_stream_bench.json).
#if defined(MMD_CP_ASYNC_CG)
#define MMD_CP16 "cp.async.cg"
#else
#define MMD_CP16 "cp.async.ca"
#endif
template <int BYTES>
MMD_D void cp_async(unsigned sdst, const void* gsrc) {
#if defined(MMD_HINT_STREAM)
  if (BYTES == 16)
    asm volatile(MMD_CP16 ".shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(sdst), "l"(gsrc), "l"(l2_policy_stream()) : "memory");
  else
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(gsrc), "l"(l2_policy_stream()) : "memory");
#else
  if (BYTES == 16)
    asm volatile(MMD_CP16 ".shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst), "l"(gsrc) : "memory");
#endif
}
template <int N>
MMD_D void ldcol_keep(const double* g, int ld, double* r) {  // ldcol with the evict_last hint
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = ldg_keep(g + i * ld);
}
MMD_D void prefetch_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(g)); }
MMD_D void prefetch_l1(const void* g) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(g)); }
// prefetch `rows` rows of a thread-private column into L2 (the demand loads of the block solve that follows the
// sweep then hit L2 instead of queueing behind the sweeps' streams in DRAM)
MMD_D void prefetch_col_l2(const double* g, int rows, int ld) {
  for (int i = 0; i < rows; ++i) prefetch_l2(g + i * ld);
}
MMD_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
MMD_D void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// copy an N-double record (16-byte pieces when N is even)
template <int N>
MMD_D void cp_async_rec(unsigned sdst, const double* gsrc) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) cp_async<16>(sdst + 16 * i, gsrc + 2 * i);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) cp_async<8>(sdst + 8 * i, gsrc + i);
  }
}
template <int N>
MMD_D void lds_rec(unsigned saddr, double* r) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(r[2 * i]), "=d"(r[2 * i + 1]) : "r"(saddr + 16 * i) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(r[i]) : "r"(saddr + 8 * i) : "memory");
  }
}
#endif

// ring record per thread and slot: [v (V) | K (X*V)], padded to an even number of doubles
template <class M>
struct RingRec {
  static constexpr int NW = M::V + M::X * M::V;
  static constexpr int NWP = (NW + 1) / 2 * 2;
};

// ------------------------------------------------------------------------------------------
// forward constraint sweep for one block:  c_b(q)   (generate_y_bar + constr, :399-411, :473-519).
// With WITH_K the position is the quasi-Newton parametrisation q = qw - J_lin^T lambda_tot, never
// materialised:   v_t = qw_v[t] - K_t^T alpha_k ,  n_k = qw_n[k] - sigma_lin * lambda_tot[k].
// The records of step s + PF are fetched by cp.async into a shared-memory ring while step s is
// computed (the loads do not depend on the recursion; the sweep would otherwise be bound by
// global-load latency).  ring: CTA ring base; slot i of thread tid at ring[(i * NT + tid) * NWP].
// crow / lamtot are thread-private columns in shared memory: element r at [r * NT].
// Not inlined on purpose: the loop gets its own register allocation, independent of what the caller
// keeps alive across the sweep.
// ------------------------------------------------------------------------------------------
template <class M>
struct SweepArgs {
  typename M::Coef C;
  double sigma_y, sigma_lin;
  double xstart[M::X];
  const double* vb;      // q-like body records of the block (V per step)
  const double* nzb;     // noise column
  const double* xoc;     // per-chain x_obs_seq column
  const double* y;
  const double* Kb;      // compressed Jacobian records (X*V per step)
  const double* alph;    // thread-private alpha column [rmax*X]
  const double* lamtot;  // shared-memory column
  double* crow;          // shared-memory column (out)
  double* ring;          // CTA ring base (shared)
  double* xend_out;      // thread-private [rmax*X] or null
  double* xs_out;        // thread-private trajectory records [rmax*S][X] (state BEFORE each step) or null
  int nta, cpb, NT, tid;
};

template <class M, bool WITH_K, bool WITH_XS = false>
__device__ __noinline__ void constr_sweep(const Dims& d, const Blk& B, const SweepArgs<M>& a) {
  constexpr int X = M::X, V = M::V, XV = M::X * M::V;
  constexpr int PF = MMD_PREFETCH_STEPS, NSL = PF + 1, NWP = RingRec<M>::NWP;
  const int ns = B.n * d.S, nta = a.nta, NT = a.NT, S = d.S;
  const typename M::Coef C = a.C;
  double x[X];
#pragma unroll
  for (int i = 0; i < X; ++i) x[i] = a.xstart[i];
  const unsigned ring0 = smem_u32(a.ring) + (unsigned)a.tid * (NWP * 8);
  const unsigned sstride = (unsigned)NT * (NWP * 8);
  const unsigned ring_end = ring0 + NSL * sstride;
  const double* vp = a.vb;   // next record to fetch
  const double* Kp = a.Kb;
  const int vbump = V * nta, Kbump = XV * nta;
  unsigned wr = ring0;
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    if (i < ns) {
      cp_async_rec<V>(wr, vp);
      if (WITH_K) cp_async_rec<XV>(wr + V * 8, Kp);
      vp += vbump;
      Kp += Kbump;
    }
    cp_async_commit();
    wr += sstride;
  }
  unsigned rd = ring0;
  double al[X];
  if (WITH_K) ldcol<X>(a.alph, nta, al);
  int k = 0, t = 0;
  for (int s = 0; s < ns; ++s) {
    cp_async_wait<PF - 1>();
    double v[V], xn[X];
    if (WITH_XS && a.xs_out) strec<X>(a.xs_out + s * X * nta, x);   // (Newton only: keeps the test out of the loop)
    lds_rec<V>(rd, v);
    if (WITH_K) {
      double Kt[XV];
      lds_rec<XV>(rd + V * 8, Kt);
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int i = 0; i < X; ++i) v[j] = fma(-Kt[i * V + j], al[i], v[j]);
    }
    if (s + PF < ns) {  // refill the slot consumed one step ago
      cp_async_rec<V>(wr, vp);
      if (WITH_K) cp_async_rec<XV>(wr + V * 8, Kp);
#if MMD_L2_PREFETCH_STEPS > 0
      if (s + PF + MMD_L2_PREFETCH_STEPS < ns) {  // pull the records of a later step from HBM into L2
        prefetch_l2(vp + MMD_L2_PREFETCH_STEPS * vbump);
        if (WITH_K) prefetch_l2(Kp + MMD_L2_PREFETCH_STEPS * Kbump);
      }
#endif
      vp += vbump;
      Kp += Kbump;
    }
    cp_async_commit();
    rd += sstride;
    if (rd == ring_end) rd = ring0;
    wr += sstride;
    if (wr == ring_end) wr = ring0;
    M::step(C, x, v, xn);
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = xn[i];
    if (++t == S) {  // end of observation interval k
      t = 0;
      if (WITH_XS && a.xend_out) stcol<X>(a.xend_out + k * X * nta, nta, x);
      if (k < B.ny) {
        double cy = M::obs(x) - a.y[B.o + k];
        if (d.noisy) {
          double nk = a.nzb[k * nta];
          if (WITH_K) nk = fma(-a.sigma_lin, a.lamtot[k * NT], nk);
          cy = fma(a.sigma_y, nk, cy);
        }
        a.crow[k * NT] = cy;
      }
      if (k == B.n - 1 && B.nx > 0) {
        double xo[X];
        ldcol<X>(a.xoc + (B.o + k) * X * a.cpb, a.cpb, xo);
#pragma unroll
        for (int i = 0; i < X; ++i) a.crow[(B.ny + i) * NT] = x[i] - xo[i];
      }
      ++k;
      if (WITH_K && k < B.n) ldcol<X>(a.alph + k * X * nta, nta, al);
    }
  }
  cp_async_wait<0>();
}

// in-place LU factorisation with partial pivoting of a dense n x n matrix (row-major, leading dimension
// LD), as scipy.linalg.lu_factor does for the reference's Newton solver (mici_extensions.py:745-763)
template <int LD>
MMD_D void lu_factor(double* A, int* piv, int n) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = fabs(A[c * LD + c]);
    for (int r = c + 1; r < n; ++r) {
      const double a = fabs(A[r * LD + c]);
      if (a > best) { best = a; p = r; }
    }
    piv[c] = p;
    if (p != c)
      for (int j = 0; j < n; ++j) { const double tv = A[c * LD + j]; A[c * LD + j] = A[p * LD + j]; A[p * LD + j] = tv; }
    const double inv = 1.0 / A[c * LD + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * LD + c] * inv;
      A[r * LD + c] = f;
      for (int j = c + 1; j < n; ++j) A[r * LD + j] = fma(-f, A[c * LD + j], A[r * LD + j]);
    }
  }
}
template <int LD>
MMD_D void lu_solve(const double* A, const int* piv, int n, double* x) {
  for (int c = 0; c < n; ++c) {
    const int p = piv[c];
    if (p != c) { const double tv = x[c]; x[c] = x[p]; x[p] = tv; }
  }
  for (int i = 1; i < n; ++i) {
    double sv = x[i];
    for (int k = 0; k < i; ++k) sv = fma(-A[i * LD + k], x[k], sv);
    x[i] = sv;
  }
  for (int i = n - 1; i >= 0; --i) {
    double sv = x[i];
    for (int k = i + 1; k < n; ++k) sv = fma(-A[i * LD + k], x[k], sv);
    x[i] = sv / A[i * LD + i];
  }
}

// obs-level backward recursion: alpha_k = H_k^T lambda_k + Psib_{k+1}^T alpha_{k+1}  (J^T lambda in
// compressed form, rmult_by_jacob_constr :879-913).  Writes alpha for the block's intervals and
// returns alpha at the block start (needed for the v_0 columns of block 0).  lam: the block's rows in
// registers.  Fully unrolled over RMAX intervals so that the Psib loads (independent of the
// recursion) are issued together instead of one L2 round trip per interval.
template <class M, int NRMAX, int RMAX>
MMD_D void alpha_block(const Blk& B, const double* lam, const double* __restrict__ Psibc,
                       const double* __restrict__ xendc, int nta, double* alph_out, double* alpha_start,
                       const double* __restrict__ kapc = nullptr, double* body_bound = nullptr) {
  constexpr int X = M::X;
  double al[X], bnd = 0.0;
#pragma unroll
  for (int i = 0; i < X; ++i) al[i] = 0.0;
  MMD_SOLVE_UNROLL
  for (int k = RMAX - 1; k >= 0; --k) {
    if (k < B.n) {
      if (k < B.n - 1) {
        double Ps[X * X], t[X];
        ldcol_keep<X * X>(Psibc + (k + 1) * X * X * nta, nta, Ps);
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
      if (alph_out) stcol<X>(alph_out + k * X * nta, nta, al);
      if (kapc) {
        // |(J^T lam)_{v_t, j}| = |sum_i K_t[i][j] al[i]| <= sum_i kap_k[i] |al[i]| for every step t of interval k
        double kp[X], sb = 0.0;
        ldcol_keep<X>(kapc + k * X * nta, nta, kp);
#pragma unroll
        for (int i = 0; i < X; ++i) sb = fma(kp[i], fabs(al[i]), sb);
        bnd = (sb > bnd || sb != sb) ? sb : bnd;
      }
    }
  }
  {
    double Ps[X * X];
    ldcol_keep<X * X>(Psibc, nta, Ps);
    mtv<X, X>(Ps, al, alpha_start);
  }
  if (body_bound) *body_bound = bnd;
}

// Woodbury solve G^{-1} r for this thread's block (lmult_by_inv_gram :915-942):
//   t_b = D_b^{-1} r_b ; s = C^{-1} sum_b A_b^T t_b ; lam_b = t_b - (D_b^{-1} A_b) s
// evaluated with the explicit block inverse and sum_b A_b^T D_b^{-1} r_b = sum_b (D_b^{-1} A_b)^T r_b (D_b symmetric):
// both products take `r` directly, so every factor load (D^-1: NRMAX (NRMAX + 1) / 2, D^-1 A: NRMAX x U values per
// thread, thread-private columns) is independent of the arithmetic and can be in flight at once -- the triangular
// solves this replaces were a chain of dependent loads and multiply-adds, and the solver calls this once per
// iteration.  `r` (registers, NRMAX entries, entries >= nrows ignored) is overwritten by lam_b; s (= u-part of
// J^T G^{-1} r) is returned in `s_out`.  `extra_max` rides along the same cross-block reduction as a maximum (the
// solver's |c|_inf).
template <class M, int NRMAX, int UMAX, bool TAIL_SYNC = true>
MMD_D void inv_gram_block(const Dims& d, const Blk& B, bool has_blk, const double* __restrict__ Dic,
                          const double* __restrict__ DinvAc, const double* __restrict__ LCc, double* r, double* s_out,
                          double* extra_max, double* smem_red, const Tid& t) {
  const int U = d.U, nta = t.nta;
  const int n = has_blk ? B.nrows : 0;
  double g[UMAX + 1], tb[NRMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) g[j] = 0.0;
  g[UMAX] = extra_max ? *extra_max : 0.0;
#pragma unroll
  for (int i = 0; i < NRMAX; ++i) tb[i] = 0.0;
  if (has_blk) {
    MMD_SOLVE_UNROLL
    for (int i = 0; i < NRMAX; ++i) {
      if (i < n) {
        const double ri = r[i];
        // column i of the packed symmetric inverse: entries (k, i), k >= i, and by symmetry (i, k)
        MMD_SOLVE_UNROLL
        for (int k = i; k < NRMAX; ++k)
          if (k < n) {
            const double dv = Dic[tri(k, i) * nta];
            tb[k] = fma(dv, ri, tb[k]);
            if (k != i) tb[i] = fma(dv, r[k], tb[i]);
          }
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (j < U) g[j] = fma(DinvAc[(i * U + j) * nta], ri, g[j]);
      }
    }
  }
  block_reduce<UMAX, 1, TAIL_SYNC>(g, smem_red, t);
  if (extra_max) *extra_max = g[UMAX];
  double LCm[UMAX * (UMAX + 1) / 2];
#pragma unroll
  for (int i = 0; i < UMAX * (UMAX + 1) / 2; ++i) LCm[i] = (i < U * (U + 1) / 2) ? LCc[i * t.cpb] : 0.0;
  chol_solve_invdiag_fixed<UMAX>(LCm, U, g);
#pragma unroll
  for (int j = 0; j < UMAX; ++j) s_out[j] = g[j];
  if (has_blk) {
    MMD_SOLVE_UNROLL
    for (int i = 0; i < NRMAX; ++i) {
      if (i < n) {
        double ti = tb[i];
#pragma unroll
        for (int j = 0; j < UMAX; ++j)
          if (j < U) ti = fma(-DinvAc[(i * U + j) * nta], g[j], ti);
        r[i] = ti;
      }
    }
  }
}

}  // namespace mmd
