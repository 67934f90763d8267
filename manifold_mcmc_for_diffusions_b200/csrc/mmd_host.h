// Host-side declarations shared by the translation units of libmmd_b200.so: the handle, the error
// plumbing and the per-model table of kernel launchers (one translation unit per model, so the
// models compile in parallel and every kernel is instantiated exactly once).
#pragma once
#include "../../include/mmd_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "mmd_kernels.cuh"

std::string& mmd_err();   // thread-local last error text (defined in mmd_api.cu)

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      mmd_err() = buf_;                                                                       \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

#define FAIL(msg)      \
  do {                 \
    mmd_err() = (msg); \
    return -1;         \
  } while (0)

constexpr int UMAX = 5;      // max dim_u
#ifndef MMD_NTMAX
#define MMD_NTMAX 192         // max threads per CTA (= chains per tile x observation blocks)
#endif
constexpr int NTMAX = MMD_NTMAX;
#ifndef MMD_L2_PERSIST_MB_DEFAULT
#define MMD_L2_PERSIST_MB_DEFAULT (-1)   // persisting-L2 carve-out set by mmd_create in MB (-1: leave the device limit alone)
#endif
#ifndef MMD_MINB
#define MMD_MINB 3           // __launch_bounds__ min resident CTAs per SM for the phase kernels
#endif

struct mmd_ops;

struct mmd_handle_s {
  mmd::Dims d;
  mmd::Slots S;
  mmd::Work W;
  int model;
  int device;     // CUDA device ordinal the handle lives on (made current inside every entry point)
  int X, V, Z, V0;
  int nrmax;      // constraint rows per block the kernels are instantiated for
  double* y;      // [T]
  double* stage;  // [n_chains * dim_q] canonical staging (device)
  double* stage2;
  double* stage3; // [n_chains * T * X] third staging buffer (asynchronous read-back of x_obs_seq)
  double* tpbuf;  // thread-private scratch [n_tiles][nrmax][nta] (constraint values)
  double* hbuf;   // [chains]
  double* h0buf;  // [chains]
  double* qsave;  // q-like
  double* qtmp;   // q-like (re-tiling at a partition switch)
  double* accp;   // [chains]
  int* accepted;  // [chains]
  int* cur0;
  int partition;
  int ncmax, nbmax;
  int chain0;     // global index of this handle's first chain (Philox stream offset)
  // chain regrouping (mmd_set_chain_regrouping): slots are re-assigned at every partition switch so that chains
  // with similar projection iteration counts share a CTA tile
  bool regroup, regroup_now;
  int* slot_chain;   // device [chains]: local chain id living in each slot
  int* newpos;       // device [chains]: slot each slot's chain moves to at the pending re-tiling
  int* perm_i;       // device scratch [chains]
  double* perm_d;    // device scratch [chains]
  std::vector<int> slot_chain_host;
  bool fused;
  const mmd_ops* ops;  // kernel launchers of the handle's model
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  cudaEvent_t ev_order;   // cross-stream ordering (mmd_wait_stream / mmd_stream_wait)
  long long launches;
  bool lin_valid;
  std::vector<void*> allocs;
  // optional per-kernel event timing (bench.py's roofline leg)
  bool prof_on;
  std::vector<cudaEvent_t> prof_ev;   // pairs
  std::vector<int> prof_kid;
  size_t prof_used;
  long long* n_ok;   // [chains] successful leapfrog steps per chain (device counter)
  // on-device dual-averaging step-size adaptation (DualAveragingStepSizeAdapter, scripts/utils.py:303-306)
  bool adapting;
  double ad_target, ad_reg_coef, ad_decay, ad_offset;
  double* ad_state;  // [4][chains]: iteration count, smoothed log step size, adapt-stat error, reg target
  // standard-HMC target / Adam initialiser (allocated on first use): transposed [dim_uv][chains] arrays
  int ht_dim;                 // dim_u + dim_v_0 + T S dim_v
  double *ht_q, *ht_g, *ht_m, *ht_v, *ht_xs, *ht_val, *ht_res, *ht_msr;
  double *ht_qin, *ht_gout, *ht_res2;   // mmd_hmc_target's own input / outputs (the Adam iterates stay untouched)
  int *ht_it, *ht_mask;
  std::vector<double*> aux;   // auxiliary q-like arrays of the host-driven tree builder
  int* maskbuf;               // [chains] device copy of the caller's chain mask
  // steps with n_inner_step > 1 (allocated on first use): snapshot of q, p; scratch counter; saved status words
  double *inner_q, *inner_p;
  long long* inner_cnt;
  int* inner_status;
};


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

inline mmd::StepCoef step_coef(const mmd::Dims& d, double dt) { return mmd::make_step_coef(d.gaussian, dt); }

// model dimensions and the launcher table the C ABI dispatches through
struct mmd_ops {
  int X, V, Z, V0, Y, nrmax, rmax;
  int ngen;                        // number of run-time generator parameters of the model (0: none)
  void (*default_gen)(double*);    // the reference's generators (fills ngen values)
  int (*point)(mmd_handle, int, int);
  int (*constr)(mmd_handle);
  int (*project)(mmd_handle, int, int, int, double, double, mmd::FlowCoef);
  int (*qn)(mmd_handle, int, double, const mmd_integrator_opts*);
  int (*leapfrog)(mmd_handle, double, const mmd_integrator_opts*, int, int);
  int (*hamiltonian)(mmd_handle, int, double*);
  int (*pack)(mmd_handle, const double*, double*, long long, int);
  int (*unpack)(mmd_handle, double*, const double*, long long, int);
  int (*retile)(mmd_handle, int, int);
  int (*vec_uturn)(mmd_handle, const double*, const double*, const double*, const double*, int, double*, double*);
  int (*gen_xobs)(mmd_handle);
  int (*init_interp)(mmd_handle);
  int (*philox)(mmd_handle, uint64_t, uint64_t);
  void (*constr_rows)(mmd_handle, const std::vector<double>&, double*);
  // standard-HMC target: value / gradient / residuals for n chains in transposed layout (mmd_hmc_target.cuh)
  int (*hmc_target)(mmd_handle, const double*, double*, double*, double*, double*, int, int, const int*);
};
const mmd_ops* mmd_ops_fhn();
const mmd_ops* mmd_ops_fhn_r5();   // blocks of <= 5 observations / 6 constraint rows
const mmd_ops* mmd_ops_fhn_r16();  // blocks of <= 14 observations / 16 constraint rows
const mmd_ops* mmd_ops_sir();
