// C ABI (include/mmd_b200.h) over the CHMC kernels.  Host-side orchestration only: allocation,
// layout conversion, kernel sequencing of one constrained leapfrog step.  No CPU compute path:
// every entry point fails with an error if CUDA is unavailable.
#include "../../include/mmd_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "mmd_kernels_main.cuh"
#include "mmd_model_fhn.cuh"
#include "mmd_philox.cuh"

using namespace mmd;

namespace {

thread_local std::string g_err;

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      g_err = buf_;                                                                           \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

#define FAIL(msg)    \
  do {               \
    g_err = (msg);   \
    return -1;       \
  } while (0)

constexpr int NRMAX = 8;  // max constraint rows per block  (R - 1 + noisy + dim_x)
constexpr int RMAX = 8;   // max observations per block
constexpr int UMAX = 5;   // max dim_u

}  // namespace

struct mmd_handle_s {
  Dims d;
  Slots S;
  Work W;
  int model;
  int X, V, Z, V0;
  double* xobs;   // [T*X][ld]
  double* y;      // [T]
  double* stage;  // [n_chains * max(dim_q, ...)] AoS staging
  double* stage2;
  double* hbuf;   // [ld]
  double* h0buf;  // [ld]
  double* qsave;  // [dim_q][ld]
  double* accp;   // [ld]
  int* accepted;  // [ld]
  int partition;
  int ncmax, nbmax;
  int cpb, nslot;
  bool fused;
  size_t smem_bytes;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  long long launches;
  bool lin_valid;
  std::vector<void*> allocs;
  // optional per-kernel event timing (bench.py's roofline leg)
  bool prof_on;
  std::vector<cudaEvent_t> prof_ev;   // pairs
  std::vector<int> prof_kid;
  size_t prof_used;
  long long* n_ok;   // [ld] successful leapfrog steps per chain (device counter)
};

namespace {

template <class T>
int dalloc(mmd_handle h, T** p, size_t n) {
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  CK(cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
  h->allocs.push_back((void*)*p);
  return 0;
}

void partition_shapes(int T, int R, int init, int* nb, int* fin) {
  // mici_extensions.py:327-351
  int num_full = (T - init) / R, num_rem = (T - init) % R;
  int num_middle = num_rem == 0 ? num_full - 1 : num_full;
  *fin = num_rem == 0 ? R : num_rem;
  *nb = 2 + (num_middle > 0 ? num_middle : 0);
}

enum { KID_POINT = 0, KID_PROJECT = 1, KID_QN = 2, KID_LEAPFROG = 3, KID_OTHER = 4, KID_COUNT = 5 };

struct ProfScope {
  mmd_handle h;
  size_t idx;
  bool on;
  ProfScope(mmd_handle h_, int kid) : h(h_), idx(0), on(false) {
    if (h->prof_on && h->prof_used + 2 <= h->prof_ev.size()) {
      on = true;
      idx = h->prof_used;
      h->prof_used += 2;
      h->prof_kid.push_back(kid);
      cudaEventRecord(h->prof_ev[idx], h->stream);
    }
  }
  ~ProfScope() {
    if (on) cudaEventRecord(h->prof_ev[idx + 1], h->stream);
  }
};

template <int CPB, int NSLOT, int MINB>
struct K {
  using Mdl = FhnModel;
  static int point(mmd_handle h, int which, int with_grad) {
    ProfScope ps(h, KID_POINT);
    auto kern = k_point<Mdl, CPB, NRMAX, RMAX, UMAX, CPB * NSLOT, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<h->d.ld / CPB, CPB * h->nslot, h->smem_bytes, h->stream>>>(h->d, h->S, h->W, h->xobs, h->y,
                                                                       h->partition, which, with_grad);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int constr(mmd_handle h, const double* q_soa, double* c_soa) {
    k_constr<Mdl, CPB, NRMAX, UMAX><<<h->d.ld / CPB, CPB * h->nslot, 0, h->stream>>>(
        h->d, q_soa, h->xobs, h->y, h->partition, c_soa);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int project(mmd_handle h, int lin, int src, int dst, double hh, double qcoef) {
    ProfScope ps(h, KID_PROJECT);
    auto kern = k_project<Mdl, CPB, NRMAX, UMAX, CPB * NSLOT, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<h->d.ld / CPB, CPB * h->nslot, h->smem_bytes, h->stream>>>(h->d, h->S, h->W, h->partition, lin, src,
                                                                       dst, hh, qcoef);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int qn(mmd_handle h, int mode, double mom_coef, const mmd_integrator_opts* o) {
    ProfScope ps(h, KID_QN);
    auto kern = k_qn<Mdl, CPB, NRMAX, UMAX, CPB * NSLOT, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<h->d.ld / CPB, CPB * h->nslot, h->smem_bytes, h->stream>>>(
        h->d, h->S, h->W, h->xobs, h->y, h->partition, mode, mom_coef, o->constraint_tol, o->position_tol,
        o->divergence_tol, o->max_iters);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int leapfrog(mmd_handle h, double dt, const mmd_integrator_opts* o, int n_steps, int reset_status) {
    ProfScope ps(h, KID_LEAPFROG);
    auto kern = k_leapfrog<Mdl, CPB, NRMAX, RMAX, UMAX, CPB * NSLOT, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<h->d.ld / CPB, CPB * h->nslot, h->smem_bytes, h->stream>>>(
        h->d, h->S, h->W, h->xobs, h->y, h->partition, dt, o->constraint_tol, o->position_tol, o->divergence_tol,
        o->max_iters, o->reverse_check_tol, h->n_ok, n_steps, reset_status);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int hamiltonian(mmd_handle h, int sel, double* out) {
    k_hamiltonian<CPB><<<h->d.ld / CPB, CPB * h->nslot, (size_t)h->nslot * CPB * sizeof(double), h->stream>>>(h->d, h->S, h->W, sel, out);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
};

// dispatch on the CTA shape chosen at create time
#define DISPATCH(h, CALL)                                                  \
  ((h)->nbmax <= 24 ? K<32, 24, 1>::CALL : K<8, 128, 1>::CALL)

int to_soa(mmd_handle h, const double* host, double* dst, int rows) {
  const size_t n = (size_t)h->d.n_chains * rows;
  CK(cudaMemcpyAsync(h->stage, host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  dim3 grid((h->d.n_chains + 31) / 32, (rows + 31) / 32), blk(32, 8);
  k_aos_to_soa<<<grid, blk, 0, h->stream>>>(h->stage, dst, h->d.n_chains, rows, h->d.ld);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int from_soa(mmd_handle h, const double* src, double* host, int rows) {
  const size_t n = (size_t)h->d.n_chains * rows;
  dim3 grid((h->d.n_chains + 31) / 32, (rows + 31) / 32), blk(32, 8);
  k_soa_to_aos<<<grid, blk, 0, h->stream>>>(src, h->stage, h->d.n_chains, rows, h->d.ld);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host, h->stage, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// slot-resolved copy: gathers S.<arr>[cur or other] rows into a contiguous [rows][ld] buffer
__global__ void k_gather_slot(const double* arr, long long stride, const int* cur, int sel, long long rows,
                              long long ld, double* out) {
  const long long n = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const int sl = sel ? 1 - cur[c] : cur[c];
    out[i] = arr[sl * stride + i];
  }
}
__global__ void k_scatter_slot(double* arr, long long stride, const int* cur, int sel, long long rows,
                               long long ld, const double* in) {
  const long long n = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const int sl = sel ? 1 - cur[c] : cur[c];
    arr[sl * stride + i] = in[i];
  }
}

int gather(mmd_handle h, const double* arr, long long stride, int sel, long long rows, double* out) {
  k_gather_slot<<<592, 256, 0, h->stream>>>(arr, stride, h->S.cur, sel, rows, h->d.ld, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int scatter(mmd_handle h, double* arr, long long stride, int sel, long long rows, const double* in) {
  k_scatter_slot<<<592, 256, 0, h->stream>>>(arr, stride, h->S.cur, sel, rows, h->d.ld, in);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

const char* mmd_last_error_string(void) { return g_err.c_str(); }

void mmd_default_integrator_opts(mmd_integrator_opts* o) {
  // scripts/utils.py:124-166 defaults
  o->solver = MMD_SOLVER_QUASI_NEWTON;
  o->constraint_tol = 1e-9;
  o->position_tol = 1e-8;
  o->divergence_tol = 1e10;
  o->max_iters = 50;
  o->reverse_check_tol = 2e-8;
}

int mmd_create(const mmd_config* cfg, mmd_handle* out) {
  if (!cfg || !out) FAIL("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    FAIL("no CUDA device: this library has no CPU path");
  if (cfg->model != MMD_MODEL_FHN) FAIL("model not supported in this build (FHN only)");
  if (cfg->device < 0 || cfg->device >= ndev) FAIL("bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  using Mdl = FhnModel;
  const int T = cfg->num_obs, S = cfg->num_steps_per_obs;
  int R = cfg->num_obs_per_subseq;
  if (T <= 0 || S <= 0 || cfg->n_chains <= 0) FAIL("bad sizes");
  if (R <= 0 || R >= T) R = T;
  if (cfg->noise != MMD_NOISE_NONE) FAIL("noisy observations not supported in this build");
  if (cfg->gaussian_splitting) FAIL("gaussian splitting not supported in this build");
  const int nz = cfg->noise != MMD_NOISE_NONE;
  if (cfg->dim_u != Mdl::Z + (cfg->noise == MMD_NOISE_PARAM ? 1 : 0)) FAIL("dim_u inconsistent with model/noise");
  if (cfg->dim_u > UMAX) FAIL("dim_u too large");
  if (R - 1 + nz + Mdl::X > NRMAX || R > RMAX) FAIL("num_obs_per_subseq too large for this build");

  mmd_handle h = new mmd_handle_s();
  memset(&h->d, 0, sizeof(Dims));
  h->model = cfg->model;
  h->X = Mdl::X; h->V = Mdl::V; h->Z = Mdl::Z; h->V0 = Mdl::V0;
  Dims& d = h->d;
  d.T = T; d.S = S; d.R = R; d.U = cfg->dim_u;
  d.noisy = cfg->noise; d.gaussian = cfg->gaussian_splitting; d.sigma_fixed = cfg->sigma_fixed;
  d.delta = cfg->obs_interval / S;
  d.sd = sqrt(d.delta);
  d.off_v0 = d.U; d.off_v = d.U + Mdl::V0; d.off_n = d.off_v + T * S * Mdl::V;
  d.dim_q = d.off_n + (nz ? T * Mdl::Y : 0);
  if (R == T) {
    d.num_partition = 1;
    d.nb[0] = d.nb[1] = 1;
    d.init_size[0] = d.init_size[1] = T;
    d.fin_size[0] = d.fin_size[1] = T;
  } else {
    d.num_partition = 2;
    const int inits[2] = {R, R / 2};
    for (int p = 0; p < 2; ++p) {
      if (inits[p] < 1) { delete h; FAIL("num_obs_per_subseq must be >= 2"); }
      d.init_size[p] = inits[p];
      partition_shapes(T, R, inits[p], &d.nb[p], &d.fin_size[p]);
    }
  }
  for (int p = 0; p < 2; ++p) {
    if (d.nb[p] == 1) d.n_c[p] = T * Mdl::Y;
    else
      d.n_c[p] = (d.init_size[p] - 1 + nz + Mdl::X) + (d.nb[p] - 2) * (R - 1 + nz + Mdl::X) + d.fin_size[p];
  }
  d.n_chains = cfg->n_chains;
  d.ld = (cfg->n_chains + 31) / 32 * 32;
  h->ncmax = d.n_c[0] > d.n_c[1] ? d.n_c[0] : d.n_c[1];
  h->nbmax = d.nb[0] > d.nb[1] ? d.nb[0] : d.nb[1];
  const char* cpb_env = getenv("MMD_CPB");  // tuning override
  const char* fused_env = getenv("MMD_FUSED");
  h->fused = !(fused_env && atoi(fused_env) == 0);
  if (h->nbmax <= 24) {
    h->cpb = 32;
    (void)cpb_env;
  } else if (h->nbmax <= 128) { h->cpb = 8; }
  else { delete h; FAIL("too many observation blocks for this build"); }
  h->nslot = h->nbmax;
  h->smem_bytes = (size_t)h->nslot * (UMAX * (UMAX + 1) / 2 + 1) * h->cpb * sizeof(double);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  h->launches = 0;
  h->prof_on = false;
  h->prof_used = 0;
  h->lin_valid = false;
  h->partition = 0;

  const size_t ld = d.ld, N = (size_t)T * S, X = Mdl::X, V = Mdl::V, Z = Mdl::Z;
  const size_t NTRI = NRMAX * (NRMAX + 1) / 2;
  Slots& Sx = h->S;
  Sx.s_q = (long long)d.dim_q * ld;
  Sx.s_K = (long long)(N * X * V) * ld;
  Sx.s_Psib = (long long)(T * X * X) * ld;
  Sx.s_A = (long long)(h->ncmax * d.U) * ld;
  Sx.s_L = (long long)(h->nbmax * NTRI) * ld;
  Sx.s_LC = (long long)(UMAX * (UMAX + 1) / 2) * ld;
  Sx.s_ld = (long long)ld;
  int rc = 0;
  rc |= dalloc(h, &Sx.q, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.p, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.K, 2 * Sx.s_K);
  rc |= dalloc(h, &Sx.Psib, 2 * Sx.s_Psib);
  rc |= dalloc(h, &Sx.A, 2 * Sx.s_A);
  rc |= dalloc(h, &Sx.L, 2 * Sx.s_L);
  rc |= dalloc(h, &Sx.DinvA, 2 * Sx.s_A);
  rc |= dalloc(h, &Sx.LC, 2 * Sx.s_LC);
  rc |= dalloc(h, &Sx.gradld, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.ldv, 2 * Sx.s_ld);
  rc |= dalloc(h, &Sx.cur, ld);
  Work& W = h->W;
  rc |= dalloc(h, &W.xs, N * X * ld);
  rc |= dalloc(h, &W.Yw, N * X * X * ld);
  rc |= dalloc(h, &W.Qk, (size_t)T * X * X * ld);
  rc |= dalloc(h, &W.Zt, (size_t)T * X * Z * ld);
  rc |= dalloc(h, &W.Mk, (size_t)T * X * X * ld);
  rc |= dalloc(h, &W.LamZ, (size_t)T * Z * X * ld);
  rc |= dalloc(h, &W.Yb, (size_t)T * X * X * ld);
  rc |= dalloc(h, &W.alpha, (size_t)T * X * ld);
  rc |= dalloc(h, &W.alphi, (size_t)T * X * ld);
  rc |= dalloc(h, &W.qw, (size_t)d.dim_q * ld);
  rc |= dalloc(h, &W.cvec, (size_t)h->ncmax * ld);
  rc |= dalloc(h, &W.status, ld);
  rc |= dalloc(h, &W.iters, 2 * ld);
  rc |= dalloc(h, &W.revd, ld);
  rc |= dalloc(h, &W.hval, ld);
  rc |= dalloc(h, &W.itsum, ld);
  rc |= dalloc(h, &h->xobs, (size_t)T * X * ld);
  rc |= dalloc(h, &h->y, (size_t)T * Mdl::Y);
  size_t stage_n = (size_t)d.n_chains * (d.dim_q > (int)(N * X * V) ? d.dim_q : N * X * V);
  rc |= dalloc(h, &h->stage, stage_n);
  rc |= dalloc(h, &h->stage2, (size_t)d.dim_q * ld);
  rc |= dalloc(h, &h->hbuf, ld);
  rc |= dalloc(h, &h->h0buf, ld);
  rc |= dalloc(h, &h->qsave, (size_t)d.dim_q * ld);
  rc |= dalloc(h, &h->accp, ld);
  rc |= dalloc(h, &h->accepted, ld);
  rc |= dalloc(h, &h->n_ok, ld);
  if (rc) { mmd_destroy(h); return -2; }
  CK(cudaMemcpyAsync(h->y, cfg->y_seq, (size_t)T * Mdl::Y * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

int mmd_destroy(mmd_handle h) {
  if (!h) return 0;
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int mmd_dim_q(mmd_handle h) { return h->d.dim_q; }
int mmd_num_partition(mmd_handle h) { return h->d.num_partition; }
int mmd_num_constraints(mmd_handle h, int p) { return h->d.n_c[p & 1]; }
int mmd_num_blocks(mmd_handle h, int p) { return h->d.nb[p & 1]; }
int mmd_n_chains(mmd_handle h) { return h->d.n_chains; }
int mmd_leading_dim(mmd_handle h) { return h->d.ld; }
int mmd_get_partition(mmd_handle h) { return h->partition; }
long long mmd_launch_count(mmd_handle h) { return h->launches; }

int mmd_set_state(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition) {
  if (partition < 0 || partition >= h->d.num_partition) FAIL("bad partition");
  CK(cudaMemsetAsync(h->S.cur, 0, h->d.ld * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->W.status, 0, h->d.ld * sizeof(int), h->stream));
  if (q) { if (to_soa(h, q, h->S.q, h->d.dim_q)) return -2; }
  if (p) { if (to_soa(h, p, h->S.p, h->d.dim_q)) return -2; }
  if (x_obs_seq) { if (to_soa(h, x_obs_seq, h->xobs, h->d.T * h->X)) return -2; }
  h->partition = partition;
  h->lin_valid = false;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_set_state_soa_dev(mmd_handle h, const double* q_dev, const double* p_dev, const double* x_dev,
                          int partition) {
  if (partition < 0 || partition >= h->d.num_partition) FAIL("bad partition");
  CK(cudaMemsetAsync(h->S.cur, 0, h->d.ld * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->W.status, 0, h->d.ld * sizeof(int), h->stream));
  const size_t nq = (size_t)h->d.dim_q * h->d.ld * sizeof(double);
  if (q_dev) CK(cudaMemcpyAsync(h->S.q, q_dev, nq, cudaMemcpyDeviceToDevice, h->stream));
  if (p_dev) CK(cudaMemcpyAsync(h->S.p, p_dev, nq, cudaMemcpyDeviceToDevice, h->stream));
  if (x_dev)
    CK(cudaMemcpyAsync(h->xobs, x_dev, (size_t)h->d.T * h->X * h->d.ld * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
  h->partition = partition;
  h->lin_valid = false;
  return 0;
}

int mmd_set_momentum(mmd_handle h, const double* p) {
  if (to_soa(h, p, h->stage2, h->d.dim_q)) return -2;
  return scatter(h, h->S.p, h->S.s_q, 0, h->d.dim_q, h->stage2);
}

int mmd_get_state(mmd_handle h, double* q, double* p, double* x_obs_seq) {
  if (q) {
    if (gather(h, h->S.q, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
    if (from_soa(h, h->stage2, q, h->d.dim_q)) return -2;
  }
  if (p) {
    if (gather(h, h->S.p, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
    if (from_soa(h, h->stage2, p, h->d.dim_q)) return -2;
  }
  if (x_obs_seq) { if (from_soa(h, h->xobs, x_obs_seq, h->d.T * h->X)) return -2; }
  return 0;
}

int mmd_linearize(mmd_handle h, int with_grad) {
  CK(cudaMemsetAsync(h->W.status, 0, h->d.ld * sizeof(int), h->stream));
  int rc = DISPATCH(h, point(h, 0, with_grad));
  if (rc) return rc;
  h->lin_valid = true;
  return 0;
}

int mmd_constr(mmd_handle h, double* c_out) {
  if (gather(h, h->S.q, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
  int rc = DISPATCH(h, constr(h, h->stage2, h->W.cvec));
  if (rc) return rc;
  return from_soa(h, h->W.cvec, c_out, h->d.n_c[h->partition]);
}

int mmd_log_det_sqrt_gram(mmd_handle h, double* out) {
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  if (gather(h, h->S.ldv, h->S.s_ld, 0, 1, h->hbuf)) return -2;
  CK(cudaMemcpyAsync(out, h->hbuf, h->d.n_chains * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_grad_log_det_sqrt_gram(mmd_handle h, double* out) {
  if (!h->lin_valid) FAIL("call mmd_linearize(with_grad=1) first");
  if (gather(h, h->S.gradld, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
  return from_soa(h, h->stage2, out, h->d.dim_q);
}

int mmd_hamiltonian(mmd_handle h, double* out) {
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  int rc = DISPATCH(h, hamiltonian(h, 0, h->hbuf));
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, h->hbuf, h->d.n_chains * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_project_momentum(mmd_handle h) {
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  return DISPATCH(h, project(h, 0, 0, 0, 0.0, 0.0));
}

int mmd_normal_space_component(mmd_handle h, const double* vct, double* out) {
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  // use the inactive slot's momentum array as scratch: proj(vct) -> p(other); nsc = vct - proj(vct)
  if (to_soa(h, vct, h->stage2, h->d.dim_q)) return -2;
  if (scatter(h, h->S.p, h->S.s_q, 1, h->d.dim_q, h->stage2)) return -2;
  int rc = DISPATCH(h, project(h, 0, 1, 1, 0.0, 0.0));
  if (rc) return rc;
  if (gather(h, h->S.p, h->S.s_q, 1, h->d.dim_q, h->stage2)) return -2;
  if (from_soa(h, h->stage2, out, h->d.dim_q)) return -2;
  const size_t n = (size_t)h->d.n_chains * h->d.dim_q;
  for (size_t i = 0; i < n; ++i) out[i] = vct[i] - out[i];
  return 0;
}

int mmd_update_x_obs_seq(mmd_handle h) {
  if (gather(h, h->S.q, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
  k_gen_xobs<FhnModel, UMAX><<<(h->d.n_chains + 63) / 64, 64, 0, h->stream>>>(h->d, h->stage2, h->xobs);
  h->launches++;
  CK(cudaGetLastError());
  h->lin_valid = false;
  return 0;
}

int mmd_switch_partition(mmd_handle h) {
  h->partition = (h->partition + 1) % h->d.num_partition;
  return mmd_update_x_obs_seq(h);
}

int mmd_sample_momentum(mmd_handle h, uint64_t seed, uint64_t offset) {
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  k_philox_normal<<<592, 256, 0, h->stream>>>(h->stage2, (long long)h->d.dim_q * h->d.ld, seed, offset);
  h->launches++;
  CK(cudaGetLastError());
  if (scatter(h, h->S.p, h->S.s_q, 0, h->d.dim_q, h->stage2)) return -2;
  return DISPATCH(h, project(h, 0, 0, 0, 0.0, 0.0));
}

int mmd_get_factor(mmd_handle h, const char* name, double* out, int* rows_out) {
  const Dims& d = h->d;
  const double* arr = nullptr;
  long long stride = 0, rows = 0;
  const long long NTRI = NRMAX * (NRMAX + 1) / 2;
  std::string nm(name);
  if (nm == "K") { arr = h->S.K; stride = h->S.s_K; rows = (long long)d.T * d.S * h->X * h->V; }
  else if (nm == "Psib") { arr = h->S.Psib; stride = h->S.s_Psib; rows = (long long)d.T * h->X * h->X; }
  else if (nm == "A") { arr = h->S.A; stride = h->S.s_A; rows = (long long)d.n_c[h->partition] * d.U; }
  else if (nm == "DinvA") { arr = h->S.DinvA; stride = h->S.s_A; rows = (long long)d.n_c[h->partition] * d.U; }
  else if (nm == "L") { arr = h->S.L; stride = h->S.s_L; rows = (long long)d.nb[h->partition] * NTRI; }
  else if (nm == "LC") { arr = h->S.LC; stride = h->S.s_LC; rows = d.U * (d.U + 1) / 2; }
  else FAIL("unknown factor name");
  if (rows_out) *rows_out = (int)rows;
  if (!out) return 0;
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  // gather into stage (large enough: stage holds n_chains * max(dim_q, N*X*V) and ld-padded rows fit
  // because every factor has rows <= N*X*V)
  double* tmp = nullptr;
  CK(cudaMalloc((void**)&tmp, (size_t)rows * d.ld * sizeof(double)));
  if (gather(h, arr, stride, 0, rows, tmp)) { cudaFree(tmp); return -2; }
  std::vector<double> hostbuf((size_t)rows * d.ld);
  CK(cudaMemcpyAsync(hostbuf.data(), tmp, hostbuf.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(tmp);
  for (long long r = 0; r < rows; ++r)
    memcpy(out + r * d.n_chains, hostbuf.data() + r * d.ld, d.n_chains * sizeof(double));
  return 0;
}

static int leapfrog_impl(mmd_handle h, double dt, const mmd_integrator_opts* opts, bool reset_status,
                         int n_steps = 1) {
  mmd_integrator_opts o;
  if (opts) o = *opts; else mmd_default_integrator_opts(&o);
  if (o.solver != MMD_SOLVER_QUASI_NEWTON) FAIL("only the quasi-Newton projection solver is built");
  if (!h->lin_valid) {
    int rc0 = mmd_linearize(h, 1);
    if (rc0) return rc0;
  }
  if (h->fused) return DISPATCH(h, leapfrog(h, dt, &o, n_steps, reset_status ? 1 : 0));
  if (reset_status) CK(cudaMemsetAsync(h->W.status, 0, h->d.ld * sizeof(int), h->stream));
  int rc = 0;
  for (int step_i = 0; step_i < n_steps; ++step_i) {
  // A(dt/2): h1_flow + cotangent projection at the current point; result -> p(other)
  rc = DISPATCH(h, project(h, 0, 0, 1, 0.5 * dt, 1.0)); if (rc) return rc;
  // B(dt): h2_flow, projection onto the manifold with the Jacobian at the previous point
  k_flow<<<592, 256, 0, h->stream>>>(h->d, h->S, h->W, 0, 1, dt); h->launches++;
  rc = DISPATCH(h, qn(h, 0, 1.0 / dt, &o)); if (rc) return rc;
  // pre-evaluate dh1_dpos at the new point (fills the new slot's cache), project momentum
  rc = DISPATCH(h, point(h, 1, 1)); if (rc) return rc;
  rc = DISPATCH(h, project(h, 1, 1, 1, 0.0, 0.0)); if (rc) return rc;
  // reversibility check: step back and project with the Jacobian at the new point
  k_flow<<<592, 256, 0, h->stream>>>(h->d, h->S, h->W, 1, 1, -dt); h->launches++;
  rc = DISPATCH(h, qn(h, 1, 0.0, &o)); if (rc) return rc;
  // A(dt/2)
  rc = DISPATCH(h, project(h, 1, 1, 1, 0.5 * dt, 1.0)); if (rc) return rc;
  k_commit<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->S, h->W, o.reverse_check_tol, h->n_ok);
  h->launches++;
  CK(cudaGetLastError());
  }
  return 0;
}

int mmd_leapfrog_step(mmd_handle h, double dt, const mmd_integrator_opts* opts) {
  return leapfrog_impl(h, dt, opts, true);
}

int mmd_transition_begin(mmd_handle h, uint64_t seed, uint64_t iter) {
  int rc = mmd_linearize(h, 1); if (rc) return rc;               // also clears status
  rc = mmd_sample_momentum(h, seed, 2 * iter); if (rc) return rc;
  rc = DISPATCH(h, hamiltonian(h, 0, h->h0buf)); if (rc) return rc;
  if (gather(h, h->S.q, h->S.s_q, 0, h->d.dim_q, h->qsave)) return -2;
  return 0;
}

int mmd_transition_steps(mmd_handle h, double dt, int n_steps, const mmd_integrator_opts* opts) {
  return leapfrog_impl(h, dt, opts, false, n_steps);
}

int mmd_transition_end(mmd_handle h, uint64_t seed, uint64_t iter, int switch_partition) {
  int rc = DISPATCH(h, hamiltonian(h, 0, h->hbuf)); if (rc) return rc;
  k_decide<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->W, h->h0buf, h->hbuf, seed, 2 * iter + 1,
                                                              h->accepted, h->accp);
  h->launches++;
  k_restore<<<592, 256, 0, h->stream>>>(h->d, h->S, h->accepted, h->qsave);
  h->launches++;
  CK(cudaGetLastError());
  h->lin_valid = false;
  if (switch_partition) return mmd_switch_partition(h);
  return 0;
}

int mmd_hmc_transition(mmd_handle h, double dt, int n_leapfrog, uint64_t seed, uint64_t iter,
                       const mmd_integrator_opts* opts, int switch_partition) {
  // IndependentMomentumTransition + static-trajectory integration with Metropolis accept +
  // SwitchPartitionTransition (scripts/utils.py:292-301 with a static instead of dynamic trajectory)
  int rc = mmd_transition_begin(h, seed, iter); if (rc) return rc;
  rc = leapfrog_impl(h, dt, opts, false, n_leapfrog);
  if (rc) return rc;
  return mmd_transition_end(h, seed, iter, switch_partition);
}

int mmd_profile_enable(mmd_handle h, int on, int max_launches) {
  if (on) {
    while ((int)h->prof_ev.size() < 2 * max_launches) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->prof_ev.push_back(e);
    }
    h->prof_used = 0;
    h->prof_kid.clear();
  }
  h->prof_on = on != 0;
  return 0;
}

int mmd_profile_summary(mmd_handle h, int kid, int* count, double* total_ms) {
  CK(cudaStreamSynchronize(h->stream));
  int n = 0;
  double tot = 0.0;
  for (size_t i = 0; i < h->prof_kid.size(); ++i) {
    if (h->prof_kid[i] != kid) continue;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    tot += ms;
    n++;
  }
  if (count) *count = n;
  if (total_ms) *total_ms = tot;
  return 0;
}

long long mmd_successful_steps(mmd_handle h, int reset) {
  std::vector<long long> host(h->d.ld);
  if (cudaMemcpyAsync(host.data(), h->n_ok, h->d.ld * sizeof(long long), cudaMemcpyDeviceToHost, h->stream) !=
      cudaSuccess)
    return -1;
  cudaStreamSynchronize(h->stream);
  long long tot = 0;
  for (int i = 0; i < h->d.n_chains; ++i) tot += host[i];
  if (reset) cudaMemsetAsync(h->n_ok, 0, h->d.ld * sizeof(long long), h->stream);
  return tot;
}

long long mmd_total_qn_iterations(mmd_handle h, int reset) {
  std::vector<long long> host(h->d.ld);
  if (cudaMemcpyAsync(host.data(), h->W.itsum, h->d.ld * sizeof(long long), cudaMemcpyDeviceToHost, h->stream) !=
      cudaSuccess)
    return -1;
  cudaStreamSynchronize(h->stream);
  long long tot = 0;
  for (int i = 0; i < h->d.n_chains; ++i) tot += host[i];
  if (reset) cudaMemsetAsync(h->W.itsum, 0, h->d.ld * sizeof(long long), h->stream);
  return tot;
}

int mmd_get_transition_stats(mmd_handle h, int* accepted, double* accept_prob, int* status) {
  const int n = h->d.n_chains;
  if (accepted) CK(cudaMemcpyAsync(accepted, h->accepted, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (accept_prob) CK(cudaMemcpyAsync(accept_prob, h->accp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (status) CK(cudaMemcpyAsync(status, h->W.status, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_get_step_info(mmd_handle h, int* status, int* iters_fwd, int* iters_rev, double* rev_dist) {
  const int n = h->d.n_chains;
  if (status) CK(cudaMemcpyAsync(status, h->W.status, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (iters_fwd) CK(cudaMemcpyAsync(iters_fwd, h->W.iters, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (iters_rev)
    CK(cudaMemcpyAsync(iters_rev, h->W.iters + h->d.ld, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (rev_dist) CK(cudaMemcpyAsync(rev_dist, h->W.revd, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_project_quasi_newton(mmd_handle h, const double* q_in, double dt, const mmd_integrator_opts* opts,
                             double* q_out, int* status, int* iters) {
  mmd_integrator_opts o;
  if (opts) o = *opts; else mmd_default_integrator_opts(&o);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  CK(cudaMemsetAsync(h->W.status, 0, h->d.ld * sizeof(int), h->stream));
  if (to_soa(h, q_in, h->W.qw, h->d.dim_q)) return -2;
  int rc = DISPATCH(h, qn(h, 0, 0.0, &o));
  if (rc) return rc;
  if (gather(h, h->S.q, h->S.s_q, 1, h->d.dim_q, h->stage2)) return -2;
  if (from_soa(h, h->stage2, q_out, h->d.dim_q)) return -2;
  (void)dt;
  return mmd_get_step_info(h, status, iters, nullptr, nullptr);
}

int mmd_init_linear_interpolation(mmd_handle h, const double* u, const double* v_0, const double* x_obs_seq,
                                  int partition) {
  // find_initial_state_by_linear_interpolation (:1479-1547) for all chains at once
  if (partition < 0 || partition >= h->d.num_partition) FAIL("bad partition");
  const Dims& d = h->d;
  CK(cudaMemsetAsync(h->S.cur, 0, d.ld * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->W.status, 0, d.ld * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->S.q, 0, (size_t)d.dim_q * d.ld * sizeof(double), h->stream));
  if (to_soa(h, u, h->S.q, d.U)) return -2;
  if (to_soa(h, v_0, h->S.q + (long long)d.off_v0 * d.ld, h->V0)) return -2;
  if (to_soa(h, x_obs_seq, h->xobs, d.T * h->X)) return -2;
  const long long n = (long long)d.T * d.ld;
  k_init_interp<FhnModel, UMAX><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(d, h->S.q, h->xobs);
  h->launches++;
  CK(cudaGetLastError());
  h->partition = partition;
  h->lin_valid = false;
  return 0;
}

int mmd_timer_start(mmd_handle h) { CK(cudaEventRecord(h->ev0, h->stream)); return 0; }
int mmd_timer_stop_ms(mmd_handle h, float* ms) {
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}
int mmd_synchronize(mmd_handle h) { CK(cudaStreamSynchronize(h->stream)); return 0; }

}  // extern "C"
