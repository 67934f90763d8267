// FitzHugh-Nagumo with the notebook's prior parametrisation (FitzHugh-Nagumo_example.ipynb): same kernels
// as mmd_ops_fhn.cu, different generate_z / generate_x_0.
#include "mmd_ops.cuh"
#include "mmd_model_fhn_notebook.cuh"

const mmd_ops* mmd_ops_fhn_notebook() {
  static const mmd_ops t = make_ops<FhnNotebookModel, 8, 8>();
  return &t;
}
