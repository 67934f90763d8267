// Standard-HMC target of the noisy-observation model and the Adam initialiser built on it.
//
//   conditioned_diffusion_neg_log_dens_and_grad  (sde/mici_extensions.py:82-205): the negative log posterior density
//       phi(q) = 1/2 sum_k ((y_k - h(x_k)) / sigma)^2 + T dim_y log sigma  [+ 1/2 |q|^2 unless Gaussian splitting]
//   over q = [u | v_0 | v_seq] (no observation-noise variables) and its gradient, which the reference obtains by
//   reverse-mode differentiation through lax.scan;
//   find_initial_state_by_gradient_descent_noisy_system (:1679-1801): Adam on the same objective (always with the
//   prior term) until the mean squared residual drops below a threshold; the residuals become the noise variables.
//
// The simulation is sequential over all T*S steps of a chain (no conditioning, hence no blocks), so the mapping is one
// thread = one chain; every array is transposed ([row][chain]) so a warp reads consecutive doubles.  Forward sweep:
// trajectory to `xs`; reverse sweep: lam_t = F_t^T lam_{t+1} (+ observation source), grad v_t = B_t^T lam_{t+1},
// grad z += G_t^T lam_{t+1} with the models' closed-form Jacobians (the same functors the CHMC kernels use).
#pragma once
#include "mmd_sweeps.cuh"

namespace mmd {

#if defined(__CUDACC__)
// transposes between the reference layout [chain][dim] and [dim][chain]
static __global__ void k_transpose_in(int n, int dim, int ld_src, const double* __restrict__ src, double* __restrict__ dst,
                                      const int* __restrict__ mask) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < n && r < dim) tile[i][threadIdx.x] = src[(long long)c * ld_src + r];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (c < n && r < dim && (!mask || mask[c])) dst[(long long)r * n + c] = tile[threadIdx.x][i];
  }
}
static __global__ void k_transpose_out(int n, int dim, int ld_dst, const double* __restrict__ src, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (c < n && r < dim) tile[i][threadIdx.x] = src[(long long)r * n + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < n && r < dim) dst[(long long)c * ld_dst + r] = tile[threadIdx.x][i];
  }
}

// value (always), residuals (optional) and gradient (optional) of phi for every chain with active[c] != 0
template <class M, int UMAX>
__global__ void __launch_bounds__(128)
k_hmc_target(Dims d, const double* __restrict__ y, const double* __restrict__ qT, double* __restrict__ xs,
             double* __restrict__ val, double* __restrict__ gT, double* __restrict__ resid, int n, int add_prior,
             const int* __restrict__ active) {
  constexpr int X = M::X, V = M::V, Z = M::Z, V0 = M::V0;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n || (active && !active[c])) return;
  const int U = d.U, T = d.T, S = d.S;
  const long long N = (long long)T * S;
  const double* qc = qT + c;
  ChainPar<M, UMAX> P;
  double sq = 0.0;
  {
    double u[UMAX];
#pragma unroll
    for (int j = 0; j < UMAX; ++j) {
      u[j] = (j < U) ? qc[(long long)j * n] : 0.0;
      sq = fma(u[j], u[j], sq);
    }
    make_par<M, UMAX>(d, u, P);
  }
  const double sig = P.sigy, isig = 1.0 / sig;
  double x[X], v0[V0];
#pragma unroll
  for (int j = 0; j < V0; ++j) {
    v0[j] = qc[(long long)(U + j) * n];
    sq = fma(v0[j], v0[j], sq);
  }
  M::gen_x0(d.gen, P.z, v0, x);
  const double* vq = qc + (long long)d.off_v * n;
  double* xc = xs + c;
  double acc = 0.0;
  for (int k = 0; k < T; ++k) {
    for (int tt = 0; tt < S; ++tt) {
      const long long t = (long long)k * S + tt;
      double v[V], xn[X];
#pragma unroll
      for (int i = 0; i < X; ++i) xc[(t * X + i) * n] = x[i];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        v[j] = vq[(t * V + j) * n];
        sq = fma(v[j], v[j], sq);
      }
      M::step(P.C, x, v, xn);
#pragma unroll
      for (int i = 0; i < X; ++i) x[i] = xn[i];
    }
    const double r = (y[k] - M::obs(x)) * isig;
    acc = fma(0.5 * r, r, acc);
    if (resid) resid[(long long)k * n + c] = r;
  }
#pragma unroll
  for (int i = 0; i < X; ++i) xc[(N * X + i) * n] = x[i];
  val[c] = acc + (double)T * M::Y * log(sig) + (add_prior ? 0.5 * sq : 0.0);
  if (!gT) return;

  // reverse sweep
  double* gc = gT + c;
  double* gv = gc + (long long)d.off_v * n;
  double lam[X], gz[Z], sum_r2 = 0.0;
#pragma unroll
  for (int i = 0; i < X; ++i) lam[i] = 0.0;
#pragma unroll
  for (int m = 0; m < Z; ++m) gz[m] = 0.0;
  for (int k = T - 1; k >= 0; --k) {
    {
      double xo[X], dh[X];
      const long long to = (long long)(k + 1) * S;
#pragma unroll
      for (int i = 0; i < X; ++i) xo[i] = xc[(to * X + i) * n];
      const double r = (y[k] - M::obs(xo)) * isig;
      M::obs_grad(xo, dh);
      sum_r2 = fma(r, r, sum_r2);
#pragma unroll
      for (int i = 0; i < X; ++i) lam[i] = fma(-r * isig, dh[i], lam[i]);
    }
    for (int tt = S - 1; tt >= 0; --tt) {
      const long long t = (long long)k * S + tt;
      double xt[X], v[V], F[X * X], Bm[X * V], G[X * Z], g2[V], t1[X], tz[Z];
#pragma unroll
      for (int i = 0; i < X; ++i) xt[i] = xc[(t * X + i) * n];
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = vq[(t * V + j) * n];
      M::jac_v(P.C, xt, v, Bm);
      mtv<X, V>(Bm, lam, g2);
#pragma unroll
      for (int j = 0; j < V; ++j) gv[(t * V + j) * n] = g2[j] + (add_prior ? v[j] : 0.0);
      M::jac_z(P.C, xt, v, G);
      mtv<X, Z>(G, lam, tz);
#pragma unroll
      for (int m = 0; m < Z; ++m) gz[m] += tz[m];
      M::jac_x(P.C, xt, v, F);
      mtv<X, X>(F, lam, t1);
#pragma unroll
      for (int i = 0; i < X; ++i) lam[i] = t1[i];
    }
  }
  {
    double dx0_dv0[X * V0], dx0_dz[X * Z], t0[V0], tz[Z];
    M::gen_x0_jac(d.gen, P.z, dx0_dv0, dx0_dz);
    mtv<X, V0>(dx0_dv0, lam, t0);
    mtv<X, Z>(dx0_dz, lam, tz);
#pragma unroll
    for (int j = 0; j < V0; ++j) gc[(long long)(U + j) * n] = t0[j] + (add_prior ? v0[j] : 0.0);
#pragma unroll
    for (int m = 0; m < Z; ++m) gz[m] += tz[m];
  }
#pragma unroll
  for (int j = 0; j < UMAX; ++j)
    if (j < U) {
      double s = 0.0;
      if (j < Z) {
#pragma unroll
        for (int m = 0; m < Z; ++m) s = fma(P.dzdu[m * Z + j], gz[m], s);
      } else if (d.noisy == 2) {
        s = (double)T * M::Y - sum_r2;   // d/du_Z of [1/2 sum r^2 + T log sigma] with sigma = exp(u_Z)
      }
      gc[(long long)j * n] = s + (add_prior ? P.u[j] : 0.0);
    }
}

// Adam update (jax.experimental.optimizers.adam as the reference uses it: b1 = 0.9, b2 = 0.999, eps = 1e-8) of the
// chains with upd[c] != 0; it[c] is that chain's iteration index
static __global__ void k_adam_update(int n, long long total, double* __restrict__ qT, const double* __restrict__ gT,
                                     double* __restrict__ mT, double* __restrict__ vT, const int* __restrict__ it,
                                     const int* __restrict__ upd, double step) {
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % n);
    if (!upd[c]) continue;
    const double g = gT[e];
    const double m = (1.0 - b1) * g + b1 * mT[e];
    const double v = (1.0 - b2) * (g * g) + b2 * vT[e];
    mT[e] = m;
    vT[e] = v;
    const double mhat = m / (1.0 - pow(b1, (double)(it[c] + 1)));
    const double vhat = v / (1.0 - pow(b2, (double)(it[c] + 1)));
    qT[e] = qT[e] - step * mhat / (sqrt(vhat) + eps);
  }
}
static __global__ void k_mean_sq_rows(int n, int rows, const double* __restrict__ a, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) {
    const double v = a[(long long)r * n + c];
    s += v * v;
  }
  out[c] = s / rows;
}
static __global__ void k_add_masked(int n, int* __restrict__ it, const int* __restrict__ mask) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && mask[c]) it[c] += 1;
}
static __global__ void k_zero_masked(int n, long long total, double* __restrict__ a, const int* __restrict__ mask) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    if (!mask || mask[(int)(e % n)]) a[e] = 0.0;
}
#endif

}  // namespace mmd
