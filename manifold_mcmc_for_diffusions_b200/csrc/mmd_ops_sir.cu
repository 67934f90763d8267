// SIR instantiation of the CHMC kernels (blocks of <= 16 observations / 16 constraint rows: the
// reference's SIR experiment uses one block of all 14 observations, scripts/sir_model_chmc_experiment.py).
#include "mmd_ops.cuh"
#include "mmd_model_sir.cuh"

const mmd_ops* mmd_ops_sir() {
  static const mmd_ops t = make_ops<SirModel, 16, 16>();
  return &t;
}
