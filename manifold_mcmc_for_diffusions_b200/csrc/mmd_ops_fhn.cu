// FitzHugh-Nagumo instantiation of the CHMC kernels (blocks of <= 8 observations / 8 constraint rows).
#include "mmd_ops.cuh"
#include "mmd_model_fhn.cuh"

#ifndef MMD_FHN_NRMAX
#define MMD_FHN_NRMAX 8
#define MMD_FHN_RMAX 8
#endif

const mmd_ops* mmd_ops_fhn() {
  static const mmd_ops t = make_ops<FhnModel, MMD_FHN_NRMAX, MMD_FHN_RMAX>();
  return &t;
}
