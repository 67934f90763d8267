// Hand-written FP64 CUDA kernels (sm_100a) for the constrained-HMC hot path of
// sde/mici_extensions.py (reference), batched over chains.
//
// Mapping: one thread = one (chain, observation block).  A CTA owns a *tile* of `cpb` chains x all of
// their blocks: thread tid = slot * cpb + cl  (slot = block index, cl = chain within the tile), so
// the cross-block reductions (Woodbury capacitance matrix C, u-part of J^T lambda, norms) stay inside
// the CTA and no kernel needs a grid-wide sync or a host round trip.
//
// Memory layout ("tile layout"): every per-(chain, block) quantity is a
// thread-private column of a [tile][row][nta] array (nta = threads allocated per tile), every per-chain
// quantity a column of a [tile][row][cpb] array.  A warp therefore always reads 32 consecutive
// doubles, a CTA streams one contiguous slab per array, and the inner loops address memory as
// base + row * nta with no index arithmetic beyond one multiply-add.
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
#define MMD_PREFETCH_STEPS 2      // cp.async ring of the recursion sweeps: MMD_PREFETCH_STEPS + 1 slots (= steps per group of the grouped / bulk sweeps)
#endif
#ifndef MMD_POINTWISE_L2_PREFETCH
#define MMD_POINTWISE_L2_PREFETCH 8   // look-ahead (steps) of the HBM -> L2 prefetch in the pointwise passes (0 = off)
#endif
#ifndef MMD_POINT_L2_PREFETCH
#define MMD_POINT_L2_PREFETCH 0   // look-ahead (steps) of the HBM -> L2 prefetch in the linearisation sweeps (0 = off)
#endif
#ifndef MMD_POINT_L1_PREFETCH
#define MMD_POINT_L1_PREFETCH 0   // look-ahead (steps) of the L1 prefetch in the linearisation sweeps (0 = off)
#endif
#ifndef MMD_POINTWISE_L1_PREFETCH
#define MMD_POINTWISE_L1_PREFETCH 0   // look-ahead (steps) of the L1 prefetch in the pointwise passes (0 = off)
#endif
// compiler-only barrier: values read from shared memory are re-read after it instead of being kept in registers
#ifndef MMD_SMEM_RELOAD
#define MMD_SMEM_RELOAD asm volatile("" ::: "memory")
#endif
// L2 residency hints (mmd_sweeps.cuh): evict_last for the per-iteration block factors, evict_first for the streams of
// the solver sweeps.  Measured on B200 (16,384 chains): 784 -> 790 k chain-steps/s, DRAM bytes per launch -2 %.
#if !defined(MMD_NO_L2_HINTS)
#define MMD_HINT_KEEP 1
#define MMD_HINT_STREAM 1
#endif
#ifndef MMD_FACTOR_PREFETCH
#define MMD_FACTOR_PREFETCH 0   // 1: pull the per-iteration block factors into L2 before every solver sweep; 2: also before the momentum projections
#endif
#ifndef MMD_L2_PREFETCH_STEPS
#define MMD_L2_PREFETCH_STEPS 8   // additional look-ahead of the HBM -> L2 prefetch (0 = off)
#endif

namespace mmd {

// ------------------------------------------------------------------------------------------
// problem description (host fills, passed by value to kernels)
// ------------------------------------------------------------------------------------------
// run-time parameters of the model's generators (generate_z, generate_x_0): see the model functor for their meaning
#define MMD_GEN_MAX 24
struct Dims {
  double gen[MMD_GEN_MAX];
  int T, S, R;        // num_obs, num_steps_per_obs, num_obs_per_subseq  (mici_extensions.py:317-351)
  int U;              // dim_u
  int X, V;           // dim_x, dim_v of the model
  int noisy;          // 0 noiseless, 1 fixed sigma, 2 sigma = exp(u[Z])   (generate_sigma, :353-358)
  int gaussian;       // use_gaussian_splitting (:303)
  double sigma_fixed;
  int dim_q;
  int nb[2];          // number of blocks per partition
  int init_size[2];   // obs in first block
  int fin_size[2];    // obs in last block
  int n_c[2];         // constraint rows per partition
  int num_partition;
  int n_chains;
  int cpb, lcpb;      // chains per tile (power of two) and its log2
  int n_tiles;        // ceil(n_chains / cpb)
  int nslot;          // block slots allocated per tile = max(nb[0], nb[1])
  int nta;            // nslot * cpb: threads allocated per tile = row stride of thread-private arrays
  int rmax;           // max observations per block (R, or T for a single block)
  double delta, sd;   // step delta = obs_interval / S and sqrt(delta)
  int off_v0, off_v, off_n;  // row offsets into the reference's q: [u | v_0 | v_seq | n]  (:476-484)
  // rows per thread / per chain of the array families
  int rows_body;      // rmax * S * V      q-like vectors, per-step part
  int rows_noise;     // rmax (noisy) or 0 q-like vectors, observation-noise part
  int rows_head;      // U + V0            q-like vectors, per-chain part [u | v_0]
  long long off_body, off_noise, qsize;  // section offsets / total size of one q-like vector
};

// q-like vector (q, p, grad log det, work position ...) in tile layout: three sections
//   head  [tile][rows_head][cpb]   (u, v_0: shared by all blocks of the chain)
//   body  [tile][step][nta][V]     (v_t of the thread's block, step = k*S + t: one 16-byte record per
//                                   thread and step, so a warp reads 32 consecutive records)
//   noise [tile][rows_noise][nta]  (n_k of the thread's block)
struct QPtr {
  double* head;
  double* body;
  double* noise;
};

// Everything Mici caches at a position (jacob_constr_blocks, chol_gram_blocks, log_det_sqrt_gram,
// grad_log_det_sqrt_gram; mici_extensions.py:1151-1184) in compressed form, double-buffered so a
// failed step leaves the chain where it was.  `cur[chain]` selects the live slot.
struct Slots {
  double* q;       // q-like
  double* p;       // q-like
  double* gradld;  // q-like
  double* K;       // thread-private [rmax*S*X*V]
  double* Psib;    // thread-private [rmax*X*X]
  double* xend;    // thread-private [rmax*X]      state at the observation times (nonlinear obs_func)
  double* kap;     // thread-private [rmax*X]      kap_k[i] = max_{t in interval k, j} |K_t[i][j]| (update-norm bound)
  double* A;       // thread-private [NRMAX*U]     dc/du rows
  double* L;       // thread-private [NRTRI]       packed lower Cholesky factor of D_b (diagonal stored inverted)
  double* Dinv;    // thread-private [NRTRI]       packed lower triangle of D_b^{-1} (the Woodbury solves multiply by it)
  double* DinvA;   // thread-private [NRMAX*U]
  double* LC;      // per-chain [U(U+1)/2]         packed lower Cholesky factor of C (diagonal stored inverted)
  double* ldv;     // per-chain [1]
  long long s_q, s_K, s_Psib, s_xend, s_A, s_L, s_LC, s_ld;  // slot strides in elements
  int* cur;        // [n_tiles * cpb]
};

struct Work {
  double* xs;     // thread-private [rmax*S*X]   trajectory x_t at the point being linearised
  double* Yw;     // thread-private [rmax*S*X*X] forward tangent accumulator of the second-order sweep
  double* Qk;     // thread-private [rmax*X*X]
  double* Zt;     // thread-private [rmax*X*Z]
  double* Mk;     // thread-private [rmax*X*X]
  double* LamZ;   // thread-private [rmax*Z*X]
  double* Yb;     // thread-private [rmax*X*X]
  double* alpha;  // thread-private [rmax*X]
  double* alphi;  // thread-private [rmax*X]
  double* qw;     // q-like: work position for the projection solves
  double* pw;     // q-like: work momentum (after the first half kick + projection)
  double* xobs;   // per-chain [T*X]  conditioned states at observation times (x_obs_seq)
  int* status;    // [chains]  bit 1 not converged, 2 diverged, 4 non-reversible, 8 non-finite H
  int* iters;     // [2][chains] projection iterations (forward, reverse) of the last step
  double* revd;   // [chains] reverse-check distance of the last step
  long long* itsum;  // [chains] total projection iterations executed (both directions)
  double* dt_chain;  // [chains] per-chain step sizes (used instead of the scalar step size when use_dt_chain)
  int use_dt_chain;
  unsigned long long* phase;  // [32] per-phase cycle counters of thread 0 of every CTA (MMD_PHASE_CLOCK builds only)
};

// Phase clocks (tools/phase_times.py): compiled in only with -DMMD_PHASE_CLOCK; thread 0 of each CTA adds the
// cycles between successive marks to W.phase[i]
#ifdef MMD_PHASE_CLOCK
#define PH_T0 long long _pc = clock64();
#define PH(i)                                                                     \
  do {                                                                            \
    if (threadIdx.x == 0 && W.phase) {                                            \
      const long long _n = clock64();                                             \
      atomicAdd(&W.phase[i], (unsigned long long)(_n - _pc));                     \
      _pc = _n;                                                                   \
    }                                                                             \
  } while (0)
#define PH_ADD(i, n)                                                              \
  do {                                                                            \
    if (threadIdx.x == 0 && W.phase) atomicAdd(&W.phase[i], (unsigned long long)(n)); \
  } while (0)
// second, independent clock for finer marks inside a phase (slots 24..31)
#define PHX_T0 long long _px = clock64();
#define PHX_RESET _px = clock64();
#define PHX(i)                                                                    \
  do {                                                                            \
    if (threadIdx.x == 0 && W.phase) {                                            \
      const long long _n = clock64();                                             \
      atomicAdd(&W.phase[i], (unsigned long long)(_n - _px));                     \
      _px = _n;                                                                   \
    }                                                                             \
  } while (0)
#else
#define PH_T0
#define PH(i)
#define PH_ADD(i, n)
#define PHX_T0
#define PHX_RESET
#define PHX(i)
#endif

struct FlowCoef {
  int mode;
  double fq, fp, fpm;
};

// step-size dependent constants of one leapfrog step (standard / Gaussian splitting, :1186-1238)
struct StepCoef {
  double half_dt;    // h1 kick
  double qcoef;      // 1: h1 contains 1/2 |q|^2 (standard splitting); 0: Gaussian splitting
  FlowCoef fwd;      // h2_flow(dt) fused into the first projection
  FlowCoef back;     // h2_flow(-dt) for the reverse check (trial position only)
  double mom_coef;   // dh2_flow_mom_dmom / (dt or sin dt): momentum update after the projection solve
};


// step-size dependent constants for the signed step dt (host and device: with per-chain step sizes every
// thread derives its own)
MMD_HD StepCoef make_step_coef(int gaussian, double dt) {
  StepCoef sc;
  sc.half_dt = 0.5 * dt;
  if (gaussian) {
    // h2_flow = exact rotation by dt; dh2_flow_dmom = (sin dt, cos dt) (mici_extensions.py:1222-1238)
    const double c = cos(dt), sn = sin(dt);
    sc.qcoef = 0.0;
    sc.fwd = FlowCoef{2, c, sn, sn};
    sc.back = FlowCoef{1, c, -sn, -sn};
    sc.mom_coef = c / sn;
  } else {
    sc.qcoef = 1.0;
    sc.fwd = FlowCoef{2, 1.0, dt, 0.0};
    sc.back = FlowCoef{1, 1.0, -dt, 0.0};
    sc.mom_coef = 1.0 / dt;
  }
  return sc;
}

enum : int { ST_NOTCONV = 1, ST_DIVERGED = 2, ST_NONREV = 4, ST_NONFINITE = 8,
             ST_INACTIVE = 16 /* parked by the host-side tree builder: not an error, the kernels just skip the chain */ };
enum : int { PSEL_CUR = 0, PSEL_OTHER = 1, PSEL_WORK = 2 };

// ------------------------------------------------------------------------------------------
// per-thread identity and tile-layout accessors
// ------------------------------------------------------------------------------------------
struct Tid {
  int tile, tid, cl, slot, nslot, chain, cix;  // cix = tile * cpb + cl = index into per-chain scalars
  bool act;
  int nta, cpb;
};
MMD_D Tid thread_id(const Dims& d) {
#if defined(__CUDACC__)
  Tid t;
  t.tile = blockIdx.x;
  t.tid = threadIdx.x;
  t.cpb = d.cpb;
  t.cl = t.tid & (d.cpb - 1);
  t.slot = t.tid >> d.lcpb;
  t.nslot = blockDim.x >> d.lcpb;
  t.chain = t.tile * d.cpb + t.cl;
  t.cix = t.chain;
  t.act = t.chain < d.n_chains;
  t.nta = d.nta;
  return t;
#else
  return Tid();
#endif
}
// thread-private array with `rows` rows per thread: element r at tp(...)[r * nta]
MMD_D double* tp(double* arr, int rows, const Tid& t) { return arr + ((long long)t.tile * rows) * t.nta + t.tid; }
MMD_D const double* tp(const double* arr, int rows, const Tid& t) {
  return arr + ((long long)t.tile * rows) * t.nta + t.tid;
}
// per-chain array with `rows` rows per chain: element r at pc(...)[r * cpb]
MMD_D double* pc(double* arr, int rows, const Tid& t) { return arr + ((long long)t.tile * rows) * t.cpb + t.cl; }
MMD_D const double* pc(const double* arr, int rows, const Tid& t) {
  return arr + ((long long)t.tile * rows) * t.cpb + t.cl;
}
// per-step record arrays [tile][step][nta][W]: record `s` of the thread at tpr<W>(...)[s * W * nta + c]
template <int W>
MMD_D double* tpr(double* arr, int rows, const Tid& t) {
  return arr + ((long long)t.tile * rows) * t.nta + t.tid * W;
}
template <int W>
MMD_D const double* tpr(const double* arr, int rows, const Tid& t) {
  return arr + ((long long)t.tile * rows) * t.nta + t.tid * W;
}
template <class M>
MMD_D QPtr qptr(double* base, const Dims& d, const Tid& t) {
  QPtr q;
  q.head = pc(base, d.rows_head, t);
  q.body = tpr<M::V>(base + d.off_body, d.rows_body, t);
  q.noise = tp(base + d.off_noise, d.rows_noise, t);
  return q;
}

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
MMD_D void ldcol(const double* g, int ld, double* r) {  // gather N consecutive rows of one column
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = g[i * ld];
}
template <int N>
MMD_D void stcol(double* g, int ld, const double* r) {
#pragma unroll
  for (int i = 0; i < N; ++i) g[i * ld] = r[i];
}
// contiguous N-double record (16-byte aligned when N is even): vector loads / stores
template <int N>
MMD_D void ldrec(const double* g, double* r) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const double2 v = reinterpret_cast<const double2*>(g)[i];
      r[2 * i] = v.x;
      r[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = g[i];
  }
}
template <int N>
MMD_D void strec(double* g, const double* r) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) reinterpret_cast<double2*>(g)[i] = make_double2(r[2 * i], r[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = r[i];
  }
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
MMD_HD Blk get_block(const Dims& d, int part, int b) {
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
// block that holds observation index o
MMD_HD int block_of_obs(const Dims& d, int part, int o) {
  if (d.nb[part] == 1) return 0;
  const int i0 = d.init_size[part];
  if (o < i0) return 0;
  const int b = 1 + (o - i0) / d.R;
  return b < d.nb[part] ? b : d.nb[part] - 1;
}

#if defined(__CUDACC__)
// cross-block (same chain) reductions through shared memory.  All threads of the CTA must call.
// vals[0..NSUM) are replaced by the sum over block slots (deterministic ascending order),
// vals[NSUM..NSUM+NMAX) by the NaN-propagating maximum of non-negative values.
// TAIL_SYNC = false leaves out the trailing barrier: only where the caller's next CTA-wide barrier comes before
// anything writes the scratch again.
template <int NSUM, int NMAX, bool TAIL_SYNC = true>
MMD_D void block_reduce(double* vals, double* smem, const Tid& t) {
  constexpr int NV = NSUM + NMAX;
  const int NT = t.nslot * t.cpb;
#pragma unroll
  for (int i = 0; i < NV; ++i) smem[i * NT + t.tid] = vals[i];
  __syncthreads();
  // slot loop outside, value loop unrolled inside: the NV shared-memory reads of one slot are independent, so a
  // slot costs one LDS latency instead of NV (the per-value order of the additions is unchanged: ascending slots)
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  const double* col = smem + t.cl;
#pragma unroll 3
  for (int sl = 0; sl < t.nslot; ++sl) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double v = col[i * NT + sl * t.cpb];
      if (i < NSUM) {
        acc[i] += v;
      } else {
        acc[i] = (v > acc[i] || v != v) ? v : acc[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = acc[i];
  if (TAIL_SYNC) __syncthreads();
}
#endif

// in-place packed Cholesky (lower) of an n x n SPD matrix; on return the DIAGONAL holds 1 / L_ii
// (the solves multiply instead of divide); returns sum_i log L_ii
MMD_D double chol_packed_invdiag(double* Dm, int n) {
  double ld = 0.0;
  for (int j = 0; j < n; ++j) {
    double s = Dm[tri(j, j)];
    for (int k = 0; k < j; ++k) s -= Dm[tri(j, k)] * Dm[tri(j, k)];
    const double ljj = sqrt(s);
    ld += log(fabs(ljj));
    const double inv = 1.0 / ljj;
    Dm[tri(j, j)] = inv;
    for (int i = j + 1; i < n; ++i) {
      double t = Dm[tri(i, j)];
      for (int k = 0; k < j; ++k) t -= Dm[tri(i, k)] * Dm[tri(j, k)];
      Dm[tri(i, j)] = t * inv;
    }
  }
  return ld;
}
// the same factorisation with compile-time loop bounds (n <= N, rows beyond n untouched): every index is a constant
// after unrolling, so the matrix stays in registers / fixed local slots; same operations in the same order
template <int N>
MMD_D double chol_packed_invdiag_fixed(double* Dm, int n) {
  double ld = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    if (j < n) {
      double s = Dm[tri(j, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= Dm[tri(j, k)] * Dm[tri(j, k)];
      const double ljj = sqrt(s);
      ld += log(fabs(ljj));
      const double inv = 1.0 / ljj;
      Dm[tri(j, j)] = inv;
#pragma unroll
      for (int i = j + 1; i < N; ++i) {
        if (i < n) {
          double t = Dm[tri(i, j)];
#pragma unroll
          for (int k = 0; k < j; ++k) t -= Dm[tri(i, k)] * Dm[tri(j, k)];
          Dm[tri(i, j)] = t * inv;
        }
      }
    }
  }
  return ld;
}
// solve L L^T x = b in place (packed lower L with inverted diagonal)
MMD_D void chol_solve_invdiag(const double* Lm, int n, double* x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= Lm[tri(i, k)] * x[k];
    x[i] = s * Lm[tri(i, i)];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= Lm[tri(k, i)] * x[k];
    x[i] = s * Lm[tri(i, i)];
  }
}
template <int N>
MMD_D void chol_solve_invdiag_fixed(const double* Lm, int n, double* x) {  // n <= N, fully unrolled
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < n) {
      double s = x[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= Lm[tri(i, k)] * x[k];
      x[i] = s * Lm[tri(i, i)];
    }
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    if (i < n) {
      double s = x[i];
#pragma unroll
      for (int k = i + 1; k < N; ++k)
        if (k < n) s -= Lm[tri(k, i)] * x[k];
      x[i] = s * Lm[tri(i, i)];
    }
  }
}

template <class M>
MMD_D double sigma_of(const Dims& d, const double* u) {
  if (d.noisy == 1) return d.sigma_fixed;
  if (d.noisy == 2) return exp(u[M::Z]);
  return 0.0;
}

}  // namespace mmd
