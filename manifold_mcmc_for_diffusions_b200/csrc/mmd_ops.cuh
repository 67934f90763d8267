// Kernel launchers for one model / block-size instantiation (included by mmd_ops_<model>.cu).
#pragma once
#include "mmd_host.h"
#include "mmd_kernels_main.cuh"
#include "mmd_hmc_target.cuh"

using namespace mmd;

namespace {

// kernel launchers for one model / block-size instantiation
template <class Mdl, int NRMAX, int RMAX>
struct Ops {
  // per-thread regions + one model-coefficient block per chain of the tile (used by the linearisation sweeps)
  static size_t smem(mmd_handle h, int nt) {
    return (size_t)SmemPlan<Mdl, NRMAX, UMAX>::PER_THREAD * nt * sizeof(double) +
           (size_t)h->d.cpb * sizeof(typename Mdl::Coef) + 64;   // + one mbarrier per warp (bulk-copy sweeps)
  }
  static int nt(mmd_handle h) { return h->d.nb[h->partition] * h->d.cpb; }
  template <class Kern>
  static int prep(Kern kern, size_t bytes) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // tuning aid: MMD_SMEM_CARVEOUT = preferred shared-memory carve-out in percent of the unified L1 / shared array
    // (the driver otherwise picks the smallest configuration that holds the resident CTAs)
    static const int carve = [] { const char* e = getenv("MMD_SMEM_CARVEOUT"); return e ? atoi(e) : -1; }();
    if (carve >= 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    return 0;
  }
  static int point(mmd_handle h, int which, int with_grad) {
    ProfScope ps(h, KID_POINT);
    auto kern = k_point<Mdl, NRMAX, RMAX, UMAX, NTMAX, MMD_MINB>;
    const int n = nt(h);
    if (prep(kern, smem(h, n))) return -2;
    kern<<<h->d.n_tiles, n, smem(h, n), h->stream>>>(h->d, h->S, h->W, h->y, h->partition, which, with_grad);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int constr(mmd_handle h) {
    auto kern = k_constr<Mdl, NRMAX, UMAX, NTMAX, MMD_MINB>;
    const int n = nt(h);
    if (prep(kern, smem(h, n))) return -2;
    kern<<<h->d.n_tiles, n, smem(h, n), h->stream>>>(h->d, h->S, h->W, h->y, h->partition, h->tpbuf);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int project(mmd_handle h, int lin, int src, int dst, double hh, double qcoef, FlowCoef fl) {
    ProfScope ps(h, KID_PROJECT);
    auto kern = k_project<Mdl, NRMAX, RMAX, UMAX, NTMAX, MMD_MINB>;
    const int n = nt(h);
    if (prep(kern, smem(h, n))) return -2;
    kern<<<h->d.n_tiles, n, smem(h, n), h->stream>>>(h->d, h->S, h->W, h->partition, lin, src, dst, hh, qcoef, fl);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  template <bool NEWTON>
  static int qn_t(mmd_handle h, int mode, double mom_coef, const mmd_integrator_opts* o) {
    ProfScope ps(h, KID_QN);
    auto kern = k_qn<Mdl, NRMAX, RMAX, UMAX, NTMAX, MMD_MINB, NEWTON>;
    const int n = nt(h);
    if (prep(kern, smem(h, n))) return -2;
    kern<<<h->d.n_tiles, n, smem(h, n), h->stream>>>(h->d, h->S, h->W, h->y, h->partition, mode, mom_coef,
                                                   o->constraint_tol, o->position_tol, o->divergence_tol,
                                                   o->max_iters);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int qn(mmd_handle h, int mode, double mom_coef, const mmd_integrator_opts* o) {
    return o->solver == MMD_SOLVER_NEWTON ? qn_t<true>(h, mode, mom_coef, o) : qn_t<false>(h, mode, mom_coef, o);
  }
  template <bool NEWTON>
  static int leapfrog_t(mmd_handle h, double dt, const mmd_integrator_opts* o, int n_steps, int reset_status) {
    ProfScope ps(h, KID_LEAPFROG);
    auto kern = k_leapfrog<Mdl, NRMAX, RMAX, UMAX, NTMAX, MMD_MINB, NEWTON>;
    const int n = nt(h);
    if (prep(kern, smem(h, n))) return -2;
    kern<<<h->d.n_tiles, n, smem(h, n), h->stream>>>(h->d, h->S, h->W, h->y, h->partition, dt,
                                                   o->constraint_tol, o->position_tol, o->divergence_tol,
                                                   o->max_iters, o->reverse_check_tol, h->n_ok, n_steps,
                                                   reset_status);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int leapfrog(mmd_handle h, double dt, const mmd_integrator_opts* o, int n_steps, int reset_status) {
    return o->solver == MMD_SOLVER_NEWTON ? leapfrog_t<true>(h, dt, o, n_steps, reset_status)
                                          : leapfrog_t<false>(h, dt, o, n_steps, reset_status);
  }
  static int hamiltonian(mmd_handle h, int sel, double* out) {
    const int n = nt(h);
    k_hamiltonian<Mdl><<<h->d.n_tiles, n, (size_t)n * sizeof(double), h->stream>>>(h->d, h->S, h->W, h->partition,
                                                                                   sel, out);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  // canonical [n_chains][dim_q] (device) -> tile layout of a q-like vector
  static int pack(mmd_handle h, const double* canon_dev, double* base, long long stride, int sel) {
    k_pack<Mdl><<<1184, 256, 0, h->stream>>>(h->d, h->partition, canon_dev, base, stride, h->S.cur, sel);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int unpack(mmd_handle h, double* canon_dev, const double* base, long long stride, int sel) {
    k_unpack<Mdl><<<1184, 256, 0, h->stream>>>(h->d, h->partition, canon_dev, base, stride, h->S.cur, sel);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int retile(mmd_handle h, int pa, int pb) {
    // q(cur) in the tiling of partition pa -> qtmp in the tiling of pb -> slot 0; cur := 0
    k_retile<Mdl><<<1184, 256, 0, h->stream>>>(h->d, pa, pb, h->S.q, h->qtmp, h->S.s_q, h->S.cur,
                                               h->regroup_now ? h->newpos : nullptr);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->S.q, h->qtmp, (size_t)h->d.qsize * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemsetAsync(h->S.cur, 0, (size_t)h->d.n_tiles * h->d.cpb * sizeof(int), h->stream));
    return 0;
  }
  static int gen_xobs(mmd_handle h) {
    k_gen_xobs<Mdl, UMAX><<<(h->d.n_chains + 63) / 64, 64, 0, h->stream>>>(h->d, h->S, h->W, h->partition);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int init_interp(mmd_handle h) {
    k_init_interp<Mdl, UMAX><<<h->d.n_tiles, nt(h), 0, h->stream>>>(h->d, h->S, h->W, h->partition);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int philox(mmd_handle h, uint64_t seed, uint64_t offset) {
    k_philox_momentum<Mdl><<<1184, 256, 0, h->stream>>>(h->d, h->partition, h->S.p, h->S.s_q, h->S.cur, seed, offset,
                                                        h->chain0,
                                                        h->regroup ? h->slot_chain : nullptr);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int vec_uturn(mmd_handle h, const double* a, const double* dd, const double* c, const double* e, int live_e,
                       double* out1, double* out2) {
    const int n = nt(h);
    k_vec_uturn<Mdl><<<h->d.n_tiles, n, (size_t)2 * n * sizeof(double), h->stream>>>(
        h->d, h->partition, a, dd, c, e, live_e, h->S.s_q, h->S.cur, out1, out2);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static int hmc_target(mmd_handle h, const double* qT, double* xs, double* val, double* gT, double* resid, int n,
                        int add_prior, const int* active) {
    k_hmc_target<Mdl, UMAX><<<(n + 127) / 128, 128, 0, h->stream>>>(h->d, h->y, qT, xs, val, gT, resid, n, add_prior,
                                                                    active);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  static void constr_rows(mmd_handle h, const std::vector<double>& buf, double* c_out) {
    // thread-private [tile][NRMAX][nta] -> [chain][n_c]: rows of block b start at its row0
    const Dims& d = h->d;
    const int part = h->partition, nc = d.n_c[part];
    for (int c = 0; c < d.n_chains; ++c) {
      const int tile = c / d.cpb, cl = c % d.cpb;
      for (int b = 0; b < d.nb[part]; ++b) {
        const Blk B = get_block<Mdl>(d, part, b);
        for (int r = 0; r < B.nrows; ++r)
          c_out[(size_t)c * nc + B.row0 + r] = buf[((size_t)tile * NRMAX + r) * d.nta + b * d.cpb + cl];
      }
    }
  }
};


template <class Mdl, int NRMAX, int RMAX>
mmd_ops make_ops() {
  using O = Ops<Mdl, NRMAX, RMAX>;
  mmd_ops t;
  t.X = Mdl::X; t.V = Mdl::V; t.Z = Mdl::Z; t.V0 = Mdl::V0; t.Y = Mdl::Y; t.nrmax = NRMAX; t.rmax = RMAX;
  t.ngen = Mdl::NGEN; t.default_gen = Mdl::default_gen;
  t.point = O::point; t.constr = O::constr; t.project = O::project; t.qn = O::qn; t.leapfrog = O::leapfrog;
  t.hamiltonian = O::hamiltonian; t.pack = O::pack; t.unpack = O::unpack; t.retile = O::retile;
  t.vec_uturn = O::vec_uturn; t.gen_xobs = O::gen_xobs; t.init_interp = O::init_interp; t.philox = O::philox; t.constr_rows = O::constr_rows; t.hmc_target = O::hmc_target;
  return t;
}

}  // namespace
