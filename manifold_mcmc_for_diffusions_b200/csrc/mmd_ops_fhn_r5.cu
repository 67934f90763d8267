// FitzHugh-Nagumo instantiation for the reference's standard blocking (num_obs_per_subseq <= 5, noiseless
// observations: at most 6 constraint rows per block).  Same kernels as mmd_ops_fhn.cu with the per-block algebra
// unrolled over 6 rows / 5 intervals instead of 8 / 8: a third less local memory per thread in the linearisation.
#include "mmd_ops.cuh"
#include "mmd_model_fhn.cuh"

const mmd_ops* mmd_ops_fhn_r5() {
  static const mmd_ops t = make_ops<FhnModel, 6, 5>();
  return &t;
}
