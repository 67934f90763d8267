// C ABI (include/mmd_b200.h) over the CHMC kernels.  Host-side orchestration only: allocation,
// layout conversion, kernel sequencing of one constrained leapfrog step.  No CPU compute path:
// every entry point fails with an error if CUDA is unavailable.
#include "mmd_host.h"
#include "mmd_kernels_main.cuh"
#include "mmd_hmc_target.cuh"

using namespace mmd;

std::string& mmd_err() {
  thread_local std::string e;
  return e;
}

namespace {

// Every entry point runs with the handle's device current and restores the caller's device on exit: handles of
// several GPUs can be driven from one host thread, and a caller (torch, ...) that switched devices is unaffected.
struct DevGuard {
  int prev, dev;
  explicit DevGuard(int d) : prev(-1), dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
#define MMD_GUARD(h) DevGuard guard_((h)->device)

template <class T>
int dalloc(mmd_handle h, T** p, size_t n) {
  if (n == 0) n = 1;
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


#define DISPATCH(h, CALL) ((h)->ops->CALL)

int h2d_stage(mmd_handle h, const double* host, double* stage, size_t n) {
  CK(cudaMemcpyAsync(stage, host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}
int d2h_sync(mmd_handle h, double* host, const double* dev, size_t n) {
  CK(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int pack_chain(mmd_handle h, int rows, const double* canon_dev, double* dst) {
  k_pack_chain<<<592, 256, 0, h->stream>>>(h->d, rows, canon_dev, dst);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int unpack_chain(mmd_handle h, int rows, double* canon_dev, const double* src) {
  k_unpack_chain<<<592, 256, 0, h->stream>>>(h->d, rows, canon_dev, src);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int reset_slot_order(mmd_handle h) {
  if (!h->regroup) return 0;
  for (int c = 0; c < h->d.n_chains; ++c) h->slot_chain_host[c] = c;
  CK(cudaMemcpyAsync(h->slot_chain, h->slot_chain_host.data(), h->d.n_chains * sizeof(int), cudaMemcpyHostToDevice,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int reset_flags(mmd_handle h) {   // new states from the caller (in the caller's chain order)
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  CK(cudaMemsetAsync(h->S.cur, 0, nc * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->W.status, 0, nc * sizeof(int), h->stream));
  return reset_slot_order(h);
}
template <class T>
int permute_slots(mmd_handle h, T* arr, T* tmp) {
  const int n = h->d.n_chains;
  k_permute<T><<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->newpos, arr, tmp);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(arr, tmp, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}
// Plan of the next slot assignment: chains sorted (stably) by the projection iterations of their last leapfrog step,
// failed chains last -- the per-tile iteration loops run as long as their slowest chain, and a chain's iteration
// count is persistent (it follows the stiffness of its parameters), tools/iter_corr.py.
int regroup_plan(mmd_handle h) {
  const int n = h->d.n_chains;
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  std::vector<int> it(2 * nc), st(nc), order(n), newpos(n), sc(n);
  CK(cudaMemcpyAsync(it.data(), h->W.iters, 2 * nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(st.data(), h->W.status, nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) order[i] = i;
  auto key = [&](int i) { return (st[i] & ~mmd::ST_INACTIVE) ? (1 << 20) : it[i] + it[nc + i]; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
  for (int j = 0; j < n; ++j) newpos[order[j]] = j;
  for (int i = 0; i < n; ++i) sc[newpos[i]] = h->slot_chain_host[i];
  h->slot_chain_host = sc;
  CK(cudaMemcpyAsync(h->newpos, newpos.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->slot_chain, sc.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // the host vectors go out of scope
  return 0;
}

// standard-HMC target / Adam initialiser scratch, allocated on first use
int ht_reserve(mmd_handle h) {
  if (h->ht_q) return 0;
  if (h->d.noisy == MMD_NOISE_NONE) FAIL("the standard-HMC target needs observation noise (noise = 1 or 2)");
  const size_t n = (size_t)h->d.n_chains;
  h->ht_dim = h->d.off_n;
  const size_t tot = (size_t)h->ht_dim * n;
  if (dalloc(h, &h->ht_q, tot) || dalloc(h, &h->ht_g, tot) || dalloc(h, &h->ht_m, tot) || dalloc(h, &h->ht_v, tot) ||
      dalloc(h, &h->ht_xs, ((size_t)h->d.T * h->d.S + 1) * h->X * n) || dalloc(h, &h->ht_val, n) ||
      dalloc(h, &h->ht_res, (size_t)h->d.T * n) || dalloc(h, &h->ht_msr, n) || dalloc(h, &h->ht_it, n) ||
      dalloc(h, &h->ht_mask, n) || dalloc(h, &h->ht_qin, tot) || dalloc(h, &h->ht_gout, tot) ||
      dalloc(h, &h->ht_res2, (size_t)h->d.T * n))
    return -2;
  return 0;
}
int ht_upload(mmd_handle h, const double* q_host, int ld, const int* mask_dev, double* dstT) {
  const int n = h->d.n_chains, dim = h->ht_dim;
  // canonical rows go through the staging buffer (ld doubles per chain, ld <= dim_q)
  if (h2d_stage(h, q_host, h->stage, (size_t)n * ld)) return -2;
  dim3 grid((n + 31) / 32, (dim + 31) / 32), blk(32, 8);
  k_transpose_in<<<grid, blk, 0, h->stream>>>(n, dim, ld, h->stage, dstT, mask_dev);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int ht_download(mmd_handle h, const double* srcT, double* host, int ld) {
  const int n = h->d.n_chains, dim = h->ht_dim;
  dim3 grid((n + 31) / 32, (dim + 31) / 32), blk(32, 8);
  k_transpose_out<<<grid, blk, 0, h->stream>>>(n, dim, ld, srcT, h->stage);
  h->launches++;
  CK(cudaGetLastError());
  return d2h_sync(h, host, h->stage, (size_t)n * ld);
}

}  // namespace

extern "C" {

const char* mmd_last_error_string(void) { return mmd_err().c_str(); }

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
  // MMD_MODEL_FHN_NOTEBOOK is the FHN model with the notebook's generator parameters (set below)
  const bool is_fhn = cfg->model == MMD_MODEL_FHN || cfg->model == MMD_MODEL_FHN_NOTEBOOK;
  const mmd_ops* ops = is_fhn ? mmd_ops_fhn() : cfg->model == MMD_MODEL_SIR ? mmd_ops_sir() : nullptr;
  if (!ops) FAIL("unknown model id");
  {
    // the smallest instantiation that holds the problem's blocks (fewer rows = less per-thread state)
    const int T_ = cfg->num_obs, R_ = cfg->num_obs_per_subseq;
    const int nz_ = cfg->noise != MMD_NOISE_NONE;
    if (is_fhn && R_ > 0 && R_ < T_ && !getenv("MMD_FHN_WIDE")) {
      const mmd_ops* small = mmd_ops_fhn_r5();
      const mmd_ops* wide = mmd_ops_fhn_r16();
      if (R_ - 1 + nz_ + small->X <= small->nrmax && R_ <= small->rmax) ops = small;
      else if ((R_ - 1 + nz_ + ops->X > ops->nrmax || R_ > ops->rmax) &&
               R_ - 1 + nz_ + wide->X <= wide->nrmax && R_ <= wide->rmax) ops = wide;
    }
  }
  if (cfg->device < 0 || cfg->device >= ndev) FAIL("bad device ordinal");
  DevGuard guard_(cfg->device);
  {
    // L2 set-aside for persisting lines: the evict_last hints on the per-iteration block factors (mmd_sweeps.cuh)
    // only retain anything when a persisting carve-out exists (default size 0).  MMD_L2_PERSIST_MB overrides.
    const char* e = getenv("MMD_L2_PERSIST_MB");
    const long long mb = e ? atoll(e) : MMD_L2_PERSIST_MB_DEFAULT;
    if (mb >= 0) {
      int maxp = 0;
      cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
      size_t want = (size_t)mb << 20;
      if (want > (size_t)maxp) want = (size_t)maxp;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaGetLastError();
    }
  }
  const int NRMAX = ops->nrmax, RMAX = ops->rmax;
  const int T = cfg->num_obs, S = cfg->num_steps_per_obs;
  int R = cfg->num_obs_per_subseq;
  if (T <= 0 || S <= 0 || cfg->n_chains <= 0) FAIL("bad sizes");
  if (R <= 0 || R >= T) R = T;
  const int nz = cfg->noise != MMD_NOISE_NONE;
  if (cfg->dim_u != ops->Z + (cfg->noise == MMD_NOISE_PARAM ? 1 : 0)) FAIL("dim_u inconsistent with model/noise");
  if (cfg->dim_u > UMAX) FAIL("dim_u too large");
  if (R < T && (R - 1 + nz + ops->X > NRMAX || R > RMAX)) FAIL("num_obs_per_subseq too large for this build");
  if (R == T && (T * ops->Y > NRMAX || T > RMAX)) FAIL("unblocked problem too large for this build");

  mmd_handle h = new mmd_handle_s();
  memset(&h->d, 0, sizeof(Dims));
  static_assert(MMD_GEN_MAX >= 22, "generator parameter block too small");
  ops->default_gen(h->d.gen);
  if (cfg->model == MMD_MODEL_FHN_NOTEBOOK) {
    // FitzHugh-Nagumo_example.ipynb cell 18: z = [exp(.5 u0 - 1), exp(.5 u1 - 2), .5 u2 + 1, .5 u3 + 1], x_0 = v_0 - .5
    const double nb[22] = {0.5, 0.5, 0.5, 0.5, -1.0, -2.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, -0.5, -0.5,
                           0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 22; ++i) h->d.gen[i] = nb[i];
  }
  h->model = cfg->model;
  h->device = cfg->device;
  h->ops = ops;
  h->X = ops->X; h->V = ops->V; h->Z = ops->Z; h->V0 = ops->V0;
  h->nrmax = NRMAX;
  Dims& d = h->d;
  d.T = T; d.S = S; d.R = R; d.U = cfg->dim_u; d.X = ops->X; d.V = ops->V;
  d.noisy = cfg->noise; d.gaussian = cfg->gaussian_splitting; d.sigma_fixed = cfg->sigma_fixed;
  d.delta = cfg->obs_interval / S;
  d.sd = sqrt(d.delta);
  d.off_v0 = d.U; d.off_v = d.U + ops->V0; d.off_n = d.off_v + T * S * ops->V;
  d.dim_q = d.off_n + (nz ? T * ops->Y : 0);
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
    if (d.nb[p] == 1) d.n_c[p] = T * ops->Y;
    else
      d.n_c[p] = (d.init_size[p] - 1 + nz + ops->X) + (d.nb[p] - 2) * (R - 1 + nz + ops->X) + d.fin_size[p];
  }
  d.n_chains = cfg->n_chains;
  h->ncmax = d.n_c[0] > d.n_c[1] ? d.n_c[0] : d.n_c[1];
  h->nbmax = d.nb[0] > d.nb[1] ? d.nb[0] : d.nb[1];
  if (h->nbmax > NTMAX) { delete h; FAIL("too many observation blocks for this build"); }
  // chains per tile: the largest power of two that keeps the CTA at or below ~192 threads, unless overridden
  int cpb = 32;
  while (cpb > 1 && cpb * h->nbmax > 192) cpb >>= 1;
  const char* cpb_env = getenv("MMD_CPB");  // tuning override
  if (cpb_env) {
    const int c = atoi(cpb_env);
    if (c >= 1 && c <= 32 && (c & (c - 1)) == 0 && c * h->nbmax <= NTMAX) cpb = c;
  }
  d.cpb = cpb;
  d.lcpb = 0;
  while ((1 << d.lcpb) < cpb) d.lcpb++;
  d.n_tiles = (d.n_chains + cpb - 1) / cpb;
  d.nslot = h->nbmax;
  d.nta = d.nslot * cpb;
  d.rmax = R;
  d.rows_body = d.rmax * S * ops->V;
  d.rows_noise = nz ? d.rmax : 0;
  d.rows_head = d.U + ops->V0;
  d.off_body = (long long)d.n_tiles * d.rows_head * cpb;
  d.off_noise = d.off_body + (long long)d.n_tiles * d.rows_body * d.nta;
  d.qsize = d.off_noise + (long long)d.n_tiles * d.rows_noise * d.nta;
  const char* fused_env = getenv("MMD_FUSED");
  h->fused = !(fused_env && atoi(fused_env) == 0);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
  h->launches = 0;
  h->prof_on = false;
  h->prof_used = 0;
  h->lin_valid = false;
  h->partition = 0;
  h->chain0 = 0;
  h->regroup = h->regroup_now = false;
  h->slot_chain = h->newpos = h->perm_i = nullptr;
  h->perm_d = nullptr;

  const size_t X = ops->X, V = ops->V, Z = ops->Z;
  const size_t NTRI = (size_t)NRMAX * (NRMAX + 1) / 2, UTRI = UMAX * (UMAX + 1) / 2;
  const size_t tpu = (size_t)d.n_tiles * d.nta;         // elements per thread-private row
  const size_t nc = (size_t)d.n_tiles * cpb;            // padded chain count
  const size_t RS = (size_t)d.rmax * S;
  Slots& Sx = h->S;
  Sx.s_q = d.qsize;
  Sx.s_K = (long long)(RS * X * V * tpu);
  Sx.s_Psib = (long long)(d.rmax * X * X * tpu);
  Sx.s_xend = (long long)(d.rmax * X * tpu);
  Sx.s_A = (long long)(NRMAX * d.U * tpu);
  Sx.s_L = (long long)(NTRI * tpu);
  Sx.s_LC = (long long)(UTRI * nc);
  Sx.s_ld = (long long)nc;
  int rc = 0;
  rc |= dalloc(h, &Sx.q, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.p, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.K, 2 * Sx.s_K);
  rc |= dalloc(h, &Sx.Psib, 2 * Sx.s_Psib);
  rc |= dalloc(h, &Sx.xend, 2 * Sx.s_xend);
  rc |= dalloc(h, &Sx.kap, 2 * Sx.s_xend);
  rc |= dalloc(h, &Sx.A, 2 * Sx.s_A);
  rc |= dalloc(h, &Sx.L, 2 * Sx.s_L);
  rc |= dalloc(h, &Sx.Dinv, 2 * Sx.s_L);
  rc |= dalloc(h, &Sx.DinvA, 2 * Sx.s_A);
  rc |= dalloc(h, &Sx.LC, 2 * Sx.s_LC);
  rc |= dalloc(h, &Sx.gradld, 2 * Sx.s_q);
  rc |= dalloc(h, &Sx.ldv, 2 * Sx.s_ld);
  rc |= dalloc(h, &Sx.cur, nc);
  Work& W = h->W;
  rc |= dalloc(h, &W.xs, RS * X * tpu);
  rc |= dalloc(h, &W.Yw, RS * X * X * tpu);
  rc |= dalloc(h, &W.Qk, d.rmax * X * X * tpu);
  rc |= dalloc(h, &W.Zt, d.rmax * X * Z * tpu);
  rc |= dalloc(h, &W.Mk, d.rmax * X * X * tpu);
  rc |= dalloc(h, &W.LamZ, d.rmax * Z * X * tpu);
  rc |= dalloc(h, &W.Yb, d.rmax * X * X * tpu);
  rc |= dalloc(h, &W.alpha, d.rmax * X * tpu);
  rc |= dalloc(h, &W.alphi, d.rmax * X * tpu);
  rc |= dalloc(h, &W.qw, (size_t)d.qsize);
  rc |= dalloc(h, &W.pw, (size_t)d.qsize);
  rc |= dalloc(h, &W.xobs, (size_t)T * X * nc);
  rc |= dalloc(h, &W.status, nc);
  rc |= dalloc(h, &W.iters, 2 * nc);
  rc |= dalloc(h, &W.revd, nc);
  rc |= dalloc(h, &W.itsum, nc);
  rc |= dalloc(h, &W.dt_chain, nc);
  W.use_dt_chain = 0;
  W.phase = nullptr;
#ifdef MMD_PHASE_CLOCK
  rc |= dalloc(h, &W.phase, 64);
#endif
  rc |= dalloc(h, &h->ad_state, 4 * nc);
  rc |= dalloc(h, &h->maskbuf, nc);
  h->adapting = false;
  rc |= dalloc(h, &h->y, (size_t)T * ops->Y);
  const size_t stage_n = (size_t)d.n_chains * (size_t)(d.dim_q > (int)(T * X) ? d.dim_q : T * X);
  rc |= dalloc(h, &h->stage, stage_n);
  rc |= dalloc(h, &h->stage2, stage_n);
  rc |= dalloc(h, &h->stage3, (size_t)d.n_chains * (T * X > (size_t)d.rows_head ? T * X : (size_t)d.rows_head));
  rc |= dalloc(h, &h->tpbuf, (size_t)NRMAX * tpu);
  rc |= dalloc(h, &h->hbuf, nc);
  rc |= dalloc(h, &h->h0buf, nc);
  rc |= dalloc(h, &h->qsave, (size_t)d.qsize);
  rc |= dalloc(h, &h->qtmp, (size_t)d.qsize);
  rc |= dalloc(h, &h->accp, nc);
  rc |= dalloc(h, &h->accepted, nc);
  rc |= dalloc(h, &h->cur0, nc);
  rc |= dalloc(h, &h->n_ok, nc);
  if (rc) { mmd_destroy(h); return -2; }
  CK(cudaMemcpyAsync(h->y, cfg->y_seq, (size_t)T * ops->Y * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

int mmd_destroy(mmd_handle h) {
  MMD_GUARD(h);
  if (!h) return 0;
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev_order);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int mmd_dim_q(mmd_handle h) { return h->d.dim_q; }
int mmd_num_partition(mmd_handle h) { return h->d.num_partition; }
int mmd_num_constraints(mmd_handle h, int p) { return h->d.n_c[p & 1]; }
int mmd_num_blocks(mmd_handle h, int p) { return h->d.nb[p & 1]; }
int mmd_n_chains(mmd_handle h) { return h->d.n_chains; }
int mmd_chains_per_tile(mmd_handle h) { return h->d.cpb; }
int mmd_get_partition(mmd_handle h) { return h->partition; }
long long mmd_launch_count(mmd_handle h) { return h->launches; }
int mmd_set_chain_offset(mmd_handle h, int chain0) { h->chain0 = chain0; return 0; }

static int set_state_dev_impl(mmd_handle h, const double* q_dev, const double* p_dev, const double* x_dev,
                              int partition) {
  if (partition < 0 || partition >= h->d.num_partition) FAIL("bad partition");
  if (!q_dev && partition != h->partition)
    FAIL("set_state: the partition can only change together with a new position (the resident q is tiled for the "
         "current partition; use mmd_switch_partition)");
  if (reset_flags(h)) return -2;
  h->partition = partition;
  h->lin_valid = false;
  if (q_dev) { if (DISPATCH(h, pack(h, q_dev, h->S.q, h->S.s_q, 0))) return -2; }
  if (p_dev) { if (DISPATCH(h, pack(h, p_dev, h->S.p, h->S.s_q, 0))) return -2; }
  if (x_dev) { if (pack_chain(h, h->d.T * h->X, x_dev, h->W.xobs)) return -2; }
  return 0;
}

static int set_state_host(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition,
                          bool blocking) {
  const size_t nq = (size_t)h->d.n_chains * h->d.dim_q;
  if (q) { if (h2d_stage(h, q, h->stage, nq)) return -2; }
  if (set_state_dev_impl(h, q ? h->stage : nullptr, nullptr, nullptr, partition)) return -2;
  if (p) {
    if (h2d_stage(h, p, h->stage2, nq)) return -2;
    if (DISPATCH(h, pack(h, h->stage2, h->S.p, h->S.s_q, 0))) return -2;
  }
  if (x_obs_seq) {
    // `stage` is reused: stream order guarantees the q pack has consumed it
    if (h2d_stage(h, x_obs_seq, h->stage, (size_t)h->d.n_chains * h->d.T * h->X)) return -2;
    if (pack_chain(h, h->d.T * h->X, h->stage, h->W.xobs)) return -2;
  }
  if (blocking) CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_set_state(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition) {
  MMD_GUARD(h);
  return set_state_host(h, q, p, x_obs_seq, partition, true);
}

int mmd_set_state_async(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition) {
  MMD_GUARD(h);
  return set_state_host(h, q, p, x_obs_seq, partition, false);
}

int mmd_set_state_dev(mmd_handle h, const double* q_dev, const double* p_dev, const double* x_dev, int partition) {
  MMD_GUARD(h);
  return set_state_dev_impl(h, q_dev, p_dev, x_dev, partition);
}

int mmd_set_momentum(mmd_handle h, const double* p) {
  MMD_GUARD(h);
  if (h2d_stage(h, p, h->stage2, (size_t)h->d.n_chains * h->d.dim_q)) return -2;
  return DISPATCH(h, pack(h, h->stage2, h->S.p, h->S.s_q, 0));
}

int mmd_get_state(mmd_handle h, double* q, double* p, double* x_obs_seq) {
  MMD_GUARD(h);
  const size_t nq = (size_t)h->d.n_chains * h->d.dim_q;
  if (q) {
    if (DISPATCH(h, unpack(h, h->stage, h->S.q, h->S.s_q, 0))) return -2;
    if (d2h_sync(h, q, h->stage, nq)) return -2;
  }
  if (p) {
    if (DISPATCH(h, unpack(h, h->stage, h->S.p, h->S.s_q, 0))) return -2;
    if (d2h_sync(h, p, h->stage, nq)) return -2;
  }
  if (x_obs_seq) {
    if (unpack_chain(h, h->d.T * h->X, h->stage, h->W.xobs)) return -2;
    if (d2h_sync(h, x_obs_seq, h->stage, (size_t)h->d.n_chains * h->d.T * h->X)) return -2;
  }
  return 0;
}

// Same as mmd_get_state without the final synchronisation: the three results go through separate staging buffers,
// the copies are queued on the handle's stream and the host arrays (page-locked for real overlap) are valid after
// mmd_synchronize or any blocking call.
int mmd_get_state_async(mmd_handle h, double* q, double* p, double* x_obs_seq) {
  MMD_GUARD(h);
  const size_t nq = (size_t)h->d.n_chains * h->d.dim_q;
  if (q) {
    if (DISPATCH(h, unpack(h, h->stage, h->S.q, h->S.s_q, 0))) return -2;
    CK(cudaMemcpyAsync(q, h->stage, nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (p) {
    if (DISPATCH(h, unpack(h, h->stage2, h->S.p, h->S.s_q, 0))) return -2;
    CK(cudaMemcpyAsync(p, h->stage2, nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (x_obs_seq) {
    if (unpack_chain(h, h->d.T * h->X, h->stage3, h->W.xobs)) return -2;
    CK(cudaMemcpyAsync(x_obs_seq, h->stage3, (size_t)h->d.n_chains * h->d.T * h->X * sizeof(double),
                       cudaMemcpyDeviceToHost, h->stream));
  }
  return 0;
}

// Ordering against other CUDA streams (DLPack producers / consumers).  mmd_get_stream returns the handle's
// cudaStream_t: a DLPack consumer passes it to the producer's __dlpack__(stream=...).  mmd_wait_stream makes the
// handle's stream wait for the work queued so far on `other`; mmd_stream_wait makes `other` wait for the handle.
void* mmd_get_stream(mmd_handle h) { return (void*)h->stream; }
int mmd_get_device(mmd_handle h) { return h->device; }
int mmd_wait_stream(mmd_handle h, void* other) {
  MMD_GUARD(h);
  CK(cudaEventRecord(h->ev_order, (cudaStream_t)other));
  CK(cudaStreamWaitEvent(h->stream, h->ev_order, 0));
  return 0;
}
int mmd_stream_wait(mmd_handle h, void* other) {
  MMD_GUARD(h);
  CK(cudaEventRecord(h->ev_order, h->stream));
  CK(cudaStreamWaitEvent((cudaStream_t)other, h->ev_order, 0));
  return 0;
}

// [u | v_0] of every chain's current position (the traced quantities of the reference's scripts are functions of
// these: generate_z(u), generate_x_0(z, v_0), scripts/fhn_model_noiseless_obs_chmc_experiment.py:102-117), without
// un-tiling the whole position: dim_u + dim_v_0 doubles per chain instead of dim_q.
static __global__ void k_get_head(mmd::Dims d, const double* __restrict__ qbase, long long s_q,
                                  const int* __restrict__ cur, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_chains) return;
  const int tile = c >> d.lcpb, cl = c & (d.cpb - 1);
  const double* src = qbase + (long long)cur[c] * s_q + ((long long)tile * d.rows_head) * d.cpb + cl;
  for (int r = 0; r < d.rows_head; ++r) out[(long long)c * d.rows_head + r] = src[r * d.cpb];
}
int mmd_get_head(mmd_handle h, double* out) {
  MMD_GUARD(h);
  if (!out) FAIL("null argument");
  if (h->regroup) FAIL("mmd_get_head: not available with chain regrouping (slot order)");
  const int n = h->d.n_chains;
  k_get_head<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d, h->S.q, h->S.s_q, h->S.cur, h->stage3);
  h->launches++;
  CK(cudaGetLastError());
  return d2h_sync(h, out, h->stage3, (size_t)n * h->d.rows_head);
}

int mmd_get_state_dev(mmd_handle h, double* q_dev, double* p_dev, double* x_dev) {
  MMD_GUARD(h);
  if (q_dev) { if (DISPATCH(h, unpack(h, q_dev, h->S.q, h->S.s_q, 0))) return -2; }
  if (p_dev) { if (DISPATCH(h, unpack(h, p_dev, h->S.p, h->S.s_q, 0))) return -2; }
  if (x_dev) { if (unpack_chain(h, h->d.T * h->X, x_dev, h->W.xobs)) return -2; }
  return 0;
}

int mmd_linearize(mmd_handle h, int with_grad) {
  MMD_GUARD(h);
  CK(cudaMemsetAsync(h->W.status, 0, (size_t)h->d.n_tiles * h->d.cpb * sizeof(int), h->stream));
  int rc = DISPATCH(h, point(h, 0, with_grad));
  if (rc) return rc;
  h->lin_valid = true;
  return 0;
}

int mmd_constr(mmd_handle h, double* c_out) {
  MMD_GUARD(h);
  int rc = DISPATCH(h, constr(h));
  if (rc) return rc;
  std::vector<double> buf((size_t)h->d.n_tiles * h->nrmax * h->d.nta);
  if (d2h_sync(h, buf.data(), h->tpbuf, buf.size())) return -2;
  DISPATCH(h, constr_rows(h, buf, c_out));
  return 0;
}

int mmd_log_det_sqrt_gram(mmd_handle h, double* out) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  std::vector<double> ld(2 * nc);
  std::vector<int> cur(nc);
  CK(cudaMemcpyAsync(ld.data(), h->S.ldv, 2 * nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cur.data(), h->S.cur, nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < h->d.n_chains; ++c) out[c] = ld[(size_t)cur[c] * nc + c];
  return 0;
}

int mmd_grad_log_det_sqrt_gram(mmd_handle h, double* out) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize(with_grad=1) first");
  if (DISPATCH(h, unpack(h, h->stage, h->S.gradld, h->S.s_q, 0))) return -2;
  return d2h_sync(h, out, h->stage, (size_t)h->d.n_chains * h->d.dim_q);
}

int mmd_hamiltonian(mmd_handle h, double* out) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  int rc = DISPATCH(h, hamiltonian(h, 0, h->hbuf));
  if (rc) return rc;
  return d2h_sync(h, out, h->hbuf, h->d.n_chains);
}

static const FlowCoef NOFLOW = {0, 1.0, 0.0, 0.0};

int mmd_project_momentum(mmd_handle h) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  return DISPATCH(h, project(h, 0, PSEL_CUR, PSEL_CUR, 0.0, 0.0, NOFLOW));
}

int mmd_normal_space_component(mmd_handle h, const double* vct, double* out) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  // use the work momentum as scratch: proj(vct) -> pw; nsc = vct - proj(vct)
  const size_t n = (size_t)h->d.n_chains * h->d.dim_q;
  if (h2d_stage(h, vct, h->stage2, n)) return -2;
  if (DISPATCH(h, pack(h, h->stage2, h->W.pw, 0, 2))) return -2;
  int rc = DISPATCH(h, project(h, 0, PSEL_WORK, PSEL_WORK, 0.0, 0.0, NOFLOW));
  if (rc) return rc;
  if (DISPATCH(h, unpack(h, h->stage, h->W.pw, 0, 2))) return -2;
  if (d2h_sync(h, out, h->stage, n)) return -2;
  for (size_t i = 0; i < n; ++i) out[i] = vct[i] - out[i];
  return 0;
}

int mmd_update_x_obs_seq(mmd_handle h) {
  MMD_GUARD(h);
  int rc = DISPATCH(h, gen_xobs(h));
  h->lin_valid = false;
  return rc;
}

int mmd_switch_partition(mmd_handle h) {
  MMD_GUARD(h);
  // SwitchPartitionTransition.sample (:1279-1282): next partition, regenerate x_obs_seq from the position
  const int pa = h->partition, pb = (h->partition + 1) % h->d.num_partition;
  if (h->regroup) {
    // the re-tiling pass moves every position anyway: let it also re-assign the slots
    const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
    if (regroup_plan(h)) return -2;
    h->regroup_now = true;
    const int rc = DISPATCH(h, retile(h, pa, pb));
    h->regroup_now = false;
    if (rc) return -2;
    h->partition = pb;
    // per-slot results of the transition that just ended follow their chains
    if (permute_slots(h, h->accepted, h->perm_i) || permute_slots(h, h->W.status, h->perm_i) ||
        permute_slots(h, h->W.iters, h->perm_i) || permute_slots(h, h->W.iters + nc, h->perm_i) ||
        permute_slots(h, h->accp, h->perm_d) || permute_slots(h, h->W.revd, h->perm_d))
      return -2;
  } else if (pa != pb) {
    if (DISPATCH(h, retile(h, pa, pb))) return -2;
    h->partition = pb;
  }
  return mmd_update_x_obs_seq(h);
}

int mmd_set_chain_regrouping(mmd_handle h, int on) {
  MMD_GUARD(h);
  if (on && (h->adapting || h->W.use_dt_chain || !h->aux.empty()))
    FAIL("chain regrouping cannot be combined with per-chain step sizes / adaptation / NUTS work vectors");
  if (on && !h->slot_chain) {
    const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
    if (dalloc(h, &h->slot_chain, nc) || dalloc(h, &h->newpos, nc) || dalloc(h, &h->perm_i, nc) ||
        dalloc(h, &h->perm_d, nc))
      return -2;
    h->slot_chain_host.assign(h->d.n_chains, 0);
  }
  if (on && !h->regroup) {
    h->regroup = true;
    return reset_slot_order(h);
  }
  if (!on && h->regroup) {
    for (int c = 0; c < h->d.n_chains; ++c)
      if (h->slot_chain_host[c] != c) FAIL("chains are regrouped: read the states back and set them again to return to chain order");
    h->regroup = false;
  }
  return 0;
}

int mmd_get_slot_chains(mmd_handle h, int* out) {
  MMD_GUARD(h);
  for (int c = 0; c < h->d.n_chains; ++c) out[c] = h->regroup ? h->slot_chain_host[c] : c;
  return 0;
}

int mmd_sample_momentum(mmd_handle h, uint64_t seed, uint64_t offset) {
  MMD_GUARD(h);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  if (DISPATCH(h, philox(h, seed, offset))) return -2;
  return DISPATCH(h, project(h, 0, PSEL_CUR, PSEL_CUR, 0.0, 0.0, NOFLOW));
}

int mmd_get_factor(mmd_handle h, const char* name, double* out, int* rows_out) {
  MMD_GUARD(h);
  const Dims& d = h->d;
  const double* arr = nullptr;
  long long stride = 0;
  int rows = 0;
  bool per_chain = false;
  int W = 1;
  const int NTRI = h->nrmax * (h->nrmax + 1) / 2;
  std::string nm(name);
  if (nm == "K") { arr = h->S.K; stride = h->S.s_K; rows = d.rmax * d.S * h->X * h->V; W = h->X * h->V; }
  else if (nm == "Psib") { arr = h->S.Psib; stride = h->S.s_Psib; rows = d.rmax * h->X * h->X; }
  else if (nm == "xend") { arr = h->S.xend; stride = h->S.s_xend; rows = d.rmax * h->X; }
  else if (nm == "A") { arr = h->S.A; stride = h->S.s_A; rows = h->nrmax * d.U; }
  else if (nm == "DinvA") { arr = h->S.DinvA; stride = h->S.s_A; rows = h->nrmax * d.U; }
  else if (nm == "L") { arr = h->S.L; stride = h->S.s_L; rows = NTRI; }
  else if (nm == "LC") { arr = h->S.LC; stride = h->S.s_LC; rows = UMAX * (UMAX + 1) / 2; per_chain = true; }
  else FAIL("unknown factor name");
  if (rows_out) *rows_out = rows;
  if (!out) return 0;
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  if (per_chain) {
    // [chain][rows]
    const size_t nc = (size_t)d.n_tiles * d.cpb;
    std::vector<double> buf(2 * (size_t)stride);
    std::vector<int> cur(nc);
    CK(cudaMemcpyAsync(buf.data(), arr, buf.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(cur.data(), h->S.cur, nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < d.n_chains; ++c)
      for (int r = 0; r < rows; ++r)
        out[(size_t)c * rows + r] =
            buf[(size_t)cur[c] * stride + ((size_t)(c / d.cpb) * rows + r) * d.cpb + c % d.cpb];
    return 0;
  }
  // [chain][block][rows]
  const size_t n = (size_t)d.n_chains * d.nb[h->partition] * rows;
  double* tmp = nullptr;
  CK(cudaMalloc((void**)&tmp, n * sizeof(double)));
  k_unpack_tp<<<592, 256, 0, h->stream>>>(d, h->partition, rows, W, tmp, arr, stride, h->S.cur);
  h->launches++;
  int rc = d2h_sync(h, out, tmp, n);
  cudaFree(tmp);
  return rc;
}

static int leapfrog_impl(mmd_handle h, double dt, const mmd_integrator_opts* opts, bool reset_status,
                         int n_steps = 1) {
  mmd_integrator_opts o;
  if (opts) o = *opts; else mmd_default_integrator_opts(&o);
  if (o.solver != MMD_SOLVER_QUASI_NEWTON && o.solver != MMD_SOLVER_NEWTON) FAIL("unknown projection solver");
  if (!h->lin_valid) {
    int rc0 = mmd_linearize(h, 1);
    if (rc0) return rc0;
  }
  if (h->fused || h->W.use_dt_chain) return DISPATCH(h, leapfrog(h, dt, &o, n_steps, reset_status ? 1 : 0));
  if (reset_status)
    CK(cudaMemsetAsync(h->W.status, 0, (size_t)h->d.n_tiles * h->d.cpb * sizeof(int), h->stream));
  const StepCoef sc = step_coef(h->d, dt);
  int rc = 0;
  for (int step_i = 0; step_i < n_steps; ++step_i) {
    // A(dt/2): h1_flow + cotangent projection at the current point, fused h2_flow -> pw, qw
    rc = DISPATCH(h, project(h, 0, PSEL_CUR, PSEL_WORK, sc.half_dt, sc.qcoef, sc.fwd)); if (rc) return rc;
    // B(dt): projection onto the manifold with the Jacobian at the previous point
    rc = DISPATCH(h, qn(h, 0, sc.mom_coef, &o)); if (rc) return rc;
    // pre-evaluate dh1_dpos at the new point (fills the new slot's cache), project momentum
    rc = DISPATCH(h, point(h, 1, 1)); if (rc) return rc;
    rc = DISPATCH(h, project(h, 1, PSEL_OTHER, PSEL_OTHER, 0.0, 0.0, sc.back)); if (rc) return rc;
    // reversibility check: step back and project with the Jacobian at the new point
    rc = DISPATCH(h, qn(h, 1, 0.0, &o)); if (rc) return rc;
    // A(dt/2)
    rc = DISPATCH(h, project(h, 1, PSEL_OTHER, PSEL_OTHER, sc.half_dt, sc.qcoef, NOFLOW)); if (rc) return rc;
    k_commit<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->S, h->W, o.reverse_check_tol, h->n_ok);
    h->launches++;
    CK(cudaGetLastError());
  }
  return 0;
}

int mmd_leapfrog_step(mmd_handle h, double dt, const mmd_integrator_opts* opts) {
  MMD_GUARD(h);
  return leapfrog_impl(h, dt, opts, true);
}

// ConstrainedLeapfrogIntegrator.step with n_inner_step > 1 (Mici 0.1.10 _step_b; exposed by the
// reference as --num-inner-h2-step, scripts/utils.py:132, 286):
//   A(dt/2);  n times [ h2_flow(dt/n) + projection onto the manifold from the previous inner point, cotangent
//   projection at the new point, reverse check over dt/n ];  A(dt/2)
// driven from the host over the per-phase kernels.  Every inner step commits into the other state slot, so the
// start-of-step q, p are snapshotted first; chains that fail in any inner step are rolled back to the snapshot and
// re-linearised there -- like Mici, whose step works on a copy of the state and raises.
int mmd_leapfrog_step_inner(mmd_handle h, double dt, int n_inner_step, const mmd_integrator_opts* opts) {
  MMD_GUARD(h);
  if (n_inner_step < 1) FAIL("n_inner_step must be >= 1");
  if (n_inner_step == 1) return leapfrog_impl(h, dt, opts, true);
  if (h->W.use_dt_chain) FAIL("n_inner_step > 1 is not available with per-chain step sizes");
  mmd_integrator_opts o;
  if (opts) o = *opts; else mmd_default_integrator_opts(&o);
  if (o.solver != MMD_SOLVER_QUASI_NEWTON && o.solver != MMD_SOLVER_NEWTON) FAIL("unknown projection solver");
  if (!h->lin_valid) {
    int rc0 = mmd_linearize(h, 1);
    if (rc0) return rc0;
  }
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  if (!h->inner_q) {
    int rc0 = dalloc(h, &h->inner_q, (size_t)h->d.qsize);
    rc0 |= dalloc(h, &h->inner_p, (size_t)h->d.qsize);
    rc0 |= dalloc(h, &h->inner_cnt, nc);
    rc0 |= dalloc(h, &h->inner_status, nc);
    if (rc0) return -2;
  }
  CK(cudaMemsetAsync(h->W.status, 0, nc * sizeof(int), h->stream));
  k_snapshot_qp<<<1184, 256, 0, h->stream>>>(h->d, h->S, h->inner_q, h->inner_p);
  h->launches++;
  const StepCoef sc = step_coef(h->d, dt);                       // half kicks over the whole step
  const StepCoef si = step_coef(h->d, dt / n_inner_step);        // flows / momentum coefficient of one inner step
  const int grid = (h->d.n_chains + 127) / 128;
  int rc = DISPATCH(h, project(h, 0, PSEL_CUR, PSEL_WORK, sc.half_dt, sc.qcoef, si.fwd)); if (rc) return rc;
  for (int i = 0; i < n_inner_step; ++i) {
    const bool last = i == n_inner_step - 1;
    rc = DISPATCH(h, qn(h, 0, si.mom_coef, &o)); if (rc) return rc;
    rc = DISPATCH(h, point(h, 1, last ? 1 : 0)); if (rc) return rc;           // dh1_dpos only where it is needed
    rc = DISPATCH(h, project(h, 1, PSEL_OTHER, PSEL_OTHER, 0.0, 0.0, si.back)); if (rc) return rc;
    rc = DISPATCH(h, qn(h, 1, 0.0, &o)); if (rc) return rc;
    if (last) {
      rc = DISPATCH(h, project(h, 1, PSEL_OTHER, PSEL_OTHER, sc.half_dt, sc.qcoef, NOFLOW)); if (rc) return rc;
    }
    k_commit<<<grid, 128, 0, h->stream>>>(h->d, h->S, h->W, o.reverse_check_tol, last ? h->n_ok : h->inner_cnt);
    h->launches++;
    CK(cudaGetLastError());
    if (!last) {
      // the committed point is the new "previous" point: flow on from it (its momentum is already tangent)
      rc = DISPATCH(h, project(h, 0, PSEL_CUR, PSEL_WORK, 0.0, 0.0, si.fwd)); if (rc) return rc;
    }
  }
  // roll back the chains that failed, and give them back the linearisation of the point they are left at
  k_restore_qp<<<1184, 256, 0, h->stream>>>(h->d, h->S, h->W, h->inner_q, h->inner_p);
  k_status_swap<<<grid, 128, 0, h->stream>>>(h->d, h->W, h->inner_status, 1);
  h->launches += 2;
  rc = DISPATCH(h, point(h, 0, 1)); if (rc) return rc;
  k_status_swap<<<grid, 128, 0, h->stream>>>(h->d, h->W, h->inner_status, 0);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int mmd_transition_begin(mmd_handle h, uint64_t seed, uint64_t iter) {
  MMD_GUARD(h);
  int rc = mmd_linearize(h, 1); if (rc) return rc;               // also clears status
  rc = mmd_sample_momentum(h, seed, 2 * iter); if (rc) return rc;
  rc = DISPATCH(h, hamiltonian(h, 0, h->h0buf)); if (rc) return rc;
  k_snapshot<<<1184, 256, 0, h->stream>>>(h->d, h->S, h->qsave);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int mmd_transition_steps(mmd_handle h, double dt, int n_steps, const mmd_integrator_opts* opts) {
  MMD_GUARD(h);
  return leapfrog_impl(h, dt, opts, false, n_steps);
}

int mmd_transition_end(mmd_handle h, uint64_t seed, uint64_t iter, int switch_partition) {
  MMD_GUARD(h);
  int rc = DISPATCH(h, hamiltonian(h, 0, h->hbuf)); if (rc) return rc;
  k_decide<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->S, h->W, h->h0buf, h->hbuf, h->cur0, seed,
                                                              2 * iter + 1, h->chain0, h->accepted, h->accp,
                                                              h->regroup ? h->slot_chain : nullptr);
  h->launches++;
  k_restore<<<1184, 256, 0, h->stream>>>(h->d, h->S, h->accepted, h->qsave);
  h->launches++;
  CK(cudaGetLastError());
  if (h->adapting) {
    const long long nc = (long long)h->d.n_tiles * h->d.cpb;
    k_dual_averaging<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->ad_state, nc, h->accp, h->ad_target,
                                                                        h->ad_reg_coef, h->ad_decay, h->ad_offset,
                                                                        h->W.dt_chain);
    h->launches++;
    CK(cudaGetLastError());
  }
  h->lin_valid = false;
  if (switch_partition) return mmd_switch_partition(h);
  return 0;
}

int mmd_set_step_sizes(mmd_handle h, const double* dt) {
  MMD_GUARD(h);
  if (!dt) { h->W.use_dt_chain = 0; return 0; }
  if (h->regroup) FAIL("per-chain step sizes cannot be combined with chain regrouping");
  CK(cudaMemcpyAsync(h->W.dt_chain, dt, h->d.n_chains * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->W.use_dt_chain = 1;
  return 0;
}

int mmd_get_step_sizes(mmd_handle h, double* dt) {
  MMD_GUARD(h);
  CK(cudaMemcpyAsync(dt, h->W.dt_chain, h->d.n_chains * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int adapt_start_impl(mmd_handle h, const double* init_per_chain, double init_scalar, double target,
                            double reg_coefficient, double iter_decay, double iter_offset) {
  if (h->regroup) FAIL("step-size adaptation cannot be combined with chain regrouping");
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  std::vector<double> st(4 * nc, 0.0), dt(nc, init_scalar);
  if (init_per_chain)
    for (int c = 0; c < h->d.n_chains; ++c) dt[c] = init_per_chain[c];
  for (size_t c = 0; c < nc; ++c) {
    if (!(dt[c] > 0.0)) FAIL("initial step sizes must be positive");
    st[3 * nc + c] = log(10.0 * dt[c]);   // log_step_size_reg_target
  }
  CK(cudaMemcpyAsync(h->ad_state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->W.dt_chain, dt.data(), nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->ad_target = target; h->ad_reg_coef = reg_coefficient; h->ad_decay = iter_decay; h->ad_offset = iter_offset;
  h->adapting = true;
  h->W.use_dt_chain = 1;
  return 0;
}

int mmd_adapt_start(mmd_handle h, double init_step_size, double target, double reg_coefficient, double iter_decay,
                    double iter_offset) {
  MMD_GUARD(h);
  return adapt_start_impl(h, nullptr, init_step_size, target, reg_coefficient, iter_decay, iter_offset);
}

int mmd_adapt_start_per_chain(mmd_handle h, const double* init_step_sizes, double target, double reg_coefficient,
                              double iter_decay, double iter_offset) {
  MMD_GUARD(h);
  if (!init_step_sizes) FAIL("init_step_sizes is NULL");
  return adapt_start_impl(h, init_step_sizes, 1.0, target, reg_coefficient, iter_decay, iter_offset);
}

int mmd_adapt_stop(mmd_handle h, int pool) {
  MMD_GUARD(h);
  if (!h->adapting) FAIL("mmd_adapt_start was not called");
  const long long nc = (long long)h->d.n_tiles * h->d.cpb;
  k_dual_averaging_finalize<<<1, 256, 0, h->stream>>>(h->d, h->ad_state, nc, pool, h->W.dt_chain);
  h->launches++;
  CK(cudaGetLastError());
  h->adapting = false;
  return 0;
}

// ---- vector primitives for host-driven tree building (batched dynamic HMC) ----------------------
static const int* upload_mask(mmd_handle h, const int* mask) {
  if (!mask) return nullptr;
  if (cudaMemcpyAsync(h->maskbuf, mask, h->d.n_chains * sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
    return nullptr;
  return h->maskbuf;
}
static int resolve_vec(mmd_handle h, int id, double** ptr, int* live) {
  if (id == MMD_VEC_Q) { *ptr = h->S.q; *live = 1; return 0; }
  if (id == MMD_VEC_P) { *ptr = h->S.p; *live = 1; return 0; }
  if (id < 0 || id >= (int)h->aux.size()) FAIL("bad vector id");
  *ptr = h->aux[id];
  *live = 0;
  return 0;
}

int mmd_aux_reserve(mmd_handle h, int n_arrays) {
  MMD_GUARD(h);
  if (h->regroup && n_arrays > 0) FAIL("work vectors cannot be combined with chain regrouping");
  while ((int)h->aux.size() < n_arrays) {
    double* p = nullptr;
    if (dalloc(h, &p, (size_t)h->d.qsize)) return -2;
    h->aux.push_back(p);
  }
  return 0;
}

int mmd_vec_axpby(mmd_handle h, int dst, int src, double alpha, double beta, const int* mask) {
  MMD_GUARD(h);
  double *pd, *ps;
  int ld, ls;
  if (resolve_vec(h, dst, &pd, &ld) || resolve_vec(h, src, &ps, &ls)) return -1;
  const int* m = upload_mask(h, mask);
  if (mask && !m) FAIL("mask upload failed");
  k_vec_axpby<<<1184, 256, 0, h->stream>>>(h->d, pd, ld, ps, ls, h->S.s_q, h->S.cur, alpha, beta, m);
  h->launches++;
  CK(cudaGetLastError());
  if (ld) h->lin_valid = h->lin_valid && dst != MMD_VEC_Q;   // overwriting the live position invalidates the cache
  return 0;
}

int mmd_vec_uturn(mmd_handle h, int a, int dd, int c, int e, double* out1, double* out2) {
  MMD_GUARD(h);
  double *pa, *pd, *pc, *pe;
  int la, ldd, lc, le;
  if (resolve_vec(h, a, &pa, &la) || resolve_vec(h, dd, &pd, &ldd) || resolve_vec(h, c, &pc, &lc) ||
      resolve_vec(h, e, &pe, &le))
    return -1;
  if (la || ldd || lc) FAIL("only the last operand of mmd_vec_uturn may be a live vector");
  int rc = DISPATCH(h, vec_uturn(h, pa, pd, pc, pe, le, h->hbuf, h->accp));
  if (rc) return rc;
  CK(cudaMemcpyAsync(out1, h->hbuf, h->d.n_chains * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(out2, h->accp, h->d.n_chains * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_set_inactive(mmd_handle h, const int* mask, int clear_errors) {
  MMD_GUARD(h);
  const int* m = upload_mask(h, mask);
  if (mask && !m) FAIL("mask upload failed");
  k_set_inactive<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->W, m, clear_errors);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int mmd_num_generator_params(mmd_handle h) { return h ? h->ops->ngen : -1; }

int mmd_get_generator_params(mmd_handle h, double* params) {
  MMD_GUARD(h);
  if (!params) FAIL("null argument");
  for (int i = 0; i < h->ops->ngen; ++i) params[i] = h->d.gen[i];
  return 0;
}

int mmd_set_generator_params(mmd_handle h, const double* params, int n) {
  MMD_GUARD(h);
  if (!params) FAIL("null argument");
  if (n != h->ops->ngen) FAIL("wrong number of generator parameters for this model");
  for (int i = 0; i < n; ++i)
    if (!(params[i] == params[i])) FAIL("generator parameter is NaN");
  for (int i = 0; i < n; ++i) h->d.gen[i] = params[i];
  h->lin_valid = false;   // everything cached at the position depends on the generators
  return 0;
}

int mmd_relinearize(mmd_handle h) {
  MMD_GUARD(h);
  int rc = DISPATCH(h, point(h, 0, 1));
  if (rc) return rc;
  h->lin_valid = true;
  return 0;
}

int mmd_adapt_update(mmd_handle h, const double* accept_stat) {
  MMD_GUARD(h);
  if (!h->adapting) FAIL("mmd_adapt_start was not called");
  CK(cudaMemcpyAsync(h->accp, accept_stat, h->d.n_chains * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const long long nc = (long long)h->d.n_tiles * h->d.cpb;
  k_dual_averaging<<<(h->d.n_chains + 127) / 128, 128, 0, h->stream>>>(h->d, h->ad_state, nc, h->accp, h->ad_target,
                                                                      h->ad_reg_coef, h->ad_decay, h->ad_offset,
                                                                      h->W.dt_chain);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int mmd_hmc_transition(mmd_handle h, double dt, int n_leapfrog, uint64_t seed, uint64_t iter,
                       const mmd_integrator_opts* opts, int switch_partition) {
  MMD_GUARD(h);
  // IndependentMomentumTransition + static-trajectory integration with Metropolis accept +
  // SwitchPartitionTransition (scripts/utils.py:292-301 with a static instead of dynamic trajectory)
  int rc = mmd_transition_begin(h, seed, iter); if (rc) return rc;
  rc = leapfrog_impl(h, dt, opts, false, n_leapfrog);
  if (rc) return rc;
  return mmd_transition_end(h, seed, iter, switch_partition);
}

int mmd_profile_enable(mmd_handle h, int on, int max_launches) {
  MMD_GUARD(h);
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
  MMD_GUARD(h);
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

static long long sum_counter(mmd_handle h, long long* dev, int reset) {
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  std::vector<long long> host(nc);
  if (cudaMemcpyAsync(host.data(), dev, nc * sizeof(long long), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
    return -1;
  cudaStreamSynchronize(h->stream);
  long long tot = 0;
  for (int i = 0; i < h->d.n_chains; ++i) tot += host[i];
  if (reset) cudaMemsetAsync(dev, 0, nc * sizeof(long long), h->stream);
  return tot;
}

long long mmd_successful_steps(mmd_handle h, int reset) { MMD_GUARD(h); return sum_counter(h, h->n_ok, reset); }
long long mmd_total_qn_iterations(mmd_handle h, int reset) { MMD_GUARD(h); return sum_counter(h, h->W.itsum, reset); }

int mmd_debug_phase_cycles(mmd_handle h, unsigned long long* out64, int reset) {
  MMD_GUARD(h);
  if (!h->W.phase) FAIL("phase clocks are compiled in only with -DMMD_PHASE_CLOCK (tools/phase_times.py)");
  CK(cudaMemcpyAsync(out64, h->W.phase, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (reset) CK(cudaMemsetAsync(h->W.phase, 0, 64 * sizeof(unsigned long long), h->stream));
  return 0;
}

int mmd_get_transition_stats(mmd_handle h, int* accepted, double* accept_prob, int* status) {
  MMD_GUARD(h);
  const int n = h->d.n_chains;
  if (accepted) CK(cudaMemcpyAsync(accepted, h->accepted, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (accept_prob) CK(cudaMemcpyAsync(accept_prob, h->accp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (status) CK(cudaMemcpyAsync(status, h->W.status, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_get_step_info(mmd_handle h, int* status, int* iters_fwd, int* iters_rev, double* rev_dist) {
  MMD_GUARD(h);
  const int n = h->d.n_chains;
  const size_t nc = (size_t)h->d.n_tiles * h->d.cpb;
  if (status) CK(cudaMemcpyAsync(status, h->W.status, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (iters_fwd) CK(cudaMemcpyAsync(iters_fwd, h->W.iters, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (iters_rev) CK(cudaMemcpyAsync(iters_rev, h->W.iters + nc, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (rev_dist) CK(cudaMemcpyAsync(rev_dist, h->W.revd, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int mmd_project_quasi_newton(mmd_handle h, const double* q_in, double dt, const mmd_integrator_opts* opts,
                             double* q_out, int* status, int* iters) {
  MMD_GUARD(h);
  mmd_integrator_opts o;
  if (opts) o = *opts; else mmd_default_integrator_opts(&o);
  if (!h->lin_valid) FAIL("call mmd_linearize first");
  const size_t n = (size_t)h->d.n_chains * h->d.dim_q;
  CK(cudaMemsetAsync(h->W.status, 0, (size_t)h->d.n_tiles * h->d.cpb * sizeof(int), h->stream));
  if (h2d_stage(h, q_in, h->stage2, n)) return -2;
  if (DISPATCH(h, pack(h, h->stage2, h->W.qw, 0, 2))) return -2;
  int rc = DISPATCH(h, qn(h, 0, 0.0, &o));
  if (rc) return rc;
  if (DISPATCH(h, unpack(h, h->stage, h->S.q, h->S.s_q, 1))) return -2;
  if (d2h_sync(h, q_out, h->stage, n)) return -2;
  (void)dt;
  return mmd_get_step_info(h, status, iters, nullptr, nullptr);
}

int mmd_init_linear_interpolation(mmd_handle h, const double* u, const double* v_0, const double* x_obs_seq,
                                  int partition) {
  MMD_GUARD(h);
  // find_initial_state_by_linear_interpolation (:1479-1547) for all chains at once
  if (partition < 0 || partition >= h->d.num_partition) FAIL("bad partition");
  const Dims& d = h->d;
  if (reset_flags(h)) return -2;
  h->partition = partition;
  h->lin_valid = false;
  CK(cudaMemsetAsync(h->S.q, 0, (size_t)d.qsize * sizeof(double), h->stream));
  // head rows [u | v_0] of slot 0: assemble [chain][U + V0] on the host side of the staging buffer
  std::vector<double> head((size_t)d.n_chains * d.rows_head);
  for (int c = 0; c < d.n_chains; ++c) {
    for (int j = 0; j < d.U; ++j) head[(size_t)c * d.rows_head + j] = u[(size_t)c * d.U + j];
    for (int j = 0; j < h->V0; ++j) head[(size_t)c * d.rows_head + d.U + j] = v_0[(size_t)c * h->V0 + j];
  }
  if (h2d_stage(h, head.data(), h->stage2, head.size())) return -2;
  if (pack_chain(h, d.rows_head, h->stage2, h->S.q)) return -2;
  if (h2d_stage(h, x_obs_seq, h->stage, (size_t)d.n_chains * d.T * h->X)) return -2;
  if (pack_chain(h, d.T * h->X, h->stage, h->W.xobs)) return -2;
  int rc = DISPATCH(h, init_interp(h));
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));  // `head` is pageable host memory owned by this call
  return 0;
}

int mmd_timer_start(mmd_handle h) { MMD_GUARD(h); CK(cudaEventRecord(h->ev0, h->stream)); return 0; }
int mmd_timer_stop_ms(mmd_handle h, float* ms) {
  MMD_GUARD(h);
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}
int mmd_synchronize(mmd_handle h) { MMD_GUARD(h); CK(cudaStreamSynchronize(h->stream)); return 0; }


// ---- standard-HMC target (conditioned_diffusion_neg_log_dens_and_grad, mici_extensions.py:82-205) --------------
int mmd_hmc_dim(mmd_handle h) { return h->d.off_n; }
int mmd_hmc_target(mmd_handle h, const double* q, int add_prior, double* val, double* grad, double* resid) {
  MMD_GUARD(h);
  if (!q || !val) FAIL("null argument");
  if (ht_reserve(h)) return -1;
  const int n = h->d.n_chains, dim = h->ht_dim;
  if (ht_upload(h, q, dim, nullptr, h->ht_qin)) return -2;
  if (DISPATCH(h, hmc_target(h, h->ht_qin, h->ht_xs, h->ht_val, grad ? h->ht_gout : nullptr,
                             resid ? h->ht_res2 : nullptr, n, add_prior, nullptr)))
    return -2;
  if (grad && ht_download(h, h->ht_gout, grad, dim)) return -2;
  if (resid) {
    // residuals come back as [chain][T]
    dim3 grid((n + 31) / 32, (h->d.T + 31) / 32), blk(32, 8);
    k_transpose_out<<<grid, blk, 0, h->stream>>>(n, h->d.T, h->d.T, h->ht_res2, h->stage);
    h->launches++;
    CK(cudaGetLastError());
    if (d2h_sync(h, resid, h->stage, (size_t)n * h->d.T)) return -2;
  }
  return d2h_sync(h, val, h->ht_val, (size_t)n);
}

// ---- Adam initialiser for noisy systems (find_initial_state_by_gradient_descent_noisy_system, :1679-1801) ------
// begin: (re)start the chains with mask != 0 (NULL: all) from u_v [n][dim]; their Adam moments and iteration
// counters are cleared.  eval: objective gradient and residuals at the current iterates, returns the mean squared
// residual per chain.  update: one Adam step for the chains with upd != 0 (their iteration counter advances).
int mmd_adam_begin(mmd_handle h, const double* u_v, const int* mask) {
  MMD_GUARD(h);
  if (!u_v) FAIL("null argument");
  if (ht_reserve(h)) return -1;
  const int n = h->d.n_chains;
  const long long tot = (long long)h->ht_dim * n;
  int* mdev = nullptr;
  if (mask) {
    CK(cudaMemcpyAsync(h->ht_mask, mask, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    mdev = h->ht_mask;
  }
  if (ht_upload(h, u_v, h->ht_dim, mdev, h->ht_q)) return -2;
  k_zero_masked<<<1184, 256, 0, h->stream>>>(n, tot, h->ht_m, mdev);
  k_zero_masked<<<1184, 256, 0, h->stream>>>(n, tot, h->ht_v, mdev);
  h->launches += 2;
  CK(cudaGetLastError());
  if (mask) {
    std::vector<int> it(n);
    CK(cudaMemcpyAsync(it.data(), h->ht_it, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < n; ++c)
      if (mask[c]) it[c] = 0;
    CK(cudaMemcpyAsync(h->ht_it, it.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  } else {
    CK(cudaMemsetAsync(h->ht_it, 0, n * sizeof(int), h->stream));
  }
  return 0;
}
int mmd_adam_eval(mmd_handle h, double* msr_out, double* val_out) {
  MMD_GUARD(h);
  if (!h->ht_q) FAIL("mmd_adam_begin first");
  const int n = h->d.n_chains;
  if (DISPATCH(h, hmc_target(h, h->ht_q, h->ht_xs, h->ht_val, h->ht_g, h->ht_res, n, 1, nullptr))) return -2;
  k_mean_sq_rows<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d.T, h->ht_res, h->ht_msr);
  h->launches++;
  CK(cudaGetLastError());
  if (val_out) CK(cudaMemcpyAsync(val_out, h->ht_val, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (msr_out) return d2h_sync(h, msr_out, h->ht_msr, (size_t)n);
  return 0;
}
int mmd_adam_update(mmd_handle h, double step_size, const int* upd) {
  MMD_GUARD(h);
  if (!h->ht_q) FAIL("mmd_adam_begin first");
  if (!upd) FAIL("null argument");
  const int n = h->d.n_chains;
  const long long tot = (long long)h->ht_dim * n;
  CK(cudaMemcpyAsync(h->ht_mask, upd, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  k_adam_update<<<1184, 256, 0, h->stream>>>(n, tot, h->ht_q, h->ht_g, h->ht_m, h->ht_v, h->ht_it, h->ht_mask, step_size);
  k_add_masked<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->ht_it, h->ht_mask);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));   // `upd` may be reused by the caller
  return 0;
}
int mmd_adam_get(mmd_handle h, double* u_v, double* resid) {
  MMD_GUARD(h);
  if (!h->ht_q) FAIL("mmd_adam_begin first");
  const int n = h->d.n_chains;
  if (u_v && ht_download(h, h->ht_q, u_v, h->ht_dim)) return -2;
  if (resid) {
    dim3 grid((n + 31) / 32, (h->d.T + 31) / 32), blk(32, 8);
    k_transpose_out<<<grid, blk, 0, h->stream>>>(n, h->d.T, h->d.T, h->ht_res, h->stage);
    h->launches++;
    CK(cudaGetLastError());
    if (d2h_sync(h, resid, h->stage, (size_t)n * h->d.T)) return -2;
  }
  return 0;
}
}  // extern "C"
