// FitzHugh-Nagumo instantiation for long blocks (9 <= num_obs_per_subseq <= 14: up to 16 constraint rows per block),
// the R = 10 point of the reference's operation-time sweep (scripts/run_fhn_model_noiseless_obs_experiments.sh:16-22).
// Same kernels as mmd_ops_fhn.cu; the per-block algebra uses run-time bounds as in the SIR instantiation.
#include "mmd_ops.cuh"
#include "mmd_model_fhn.cuh"

const mmd_ops* mmd_ops_fhn_r16() {
  static const mmd_ops t = make_ops<FhnModel, 16, 16>();
  return &t;
}
