// FitzHugh-Nagumo instantiation of the CHMC kernels (blocks of <= 8 observations / 8 constraint rows).
#include "mmd_ops.cuh"
#include "mmd_model_fhn.cuh"

const mmd_ops* mmd_ops_fhn() {
  static const mmd_ops t = make_ops<FhnModel, 8, 8>();
  return &t;
}
