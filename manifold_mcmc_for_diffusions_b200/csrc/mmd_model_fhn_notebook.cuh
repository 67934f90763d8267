// FitzHugh-Nagumo model with the prior parametrisation of the reference's notebook
// (FitzHugh-Nagumo_example.ipynb cells 7-18): the strong-order-1.5 step, its derivatives and the observation
// function are those of FhnModel (mmd_model_fhn.cuh, inherited); only generate_z / generate_x_0 differ.  Used for the
// distributional known-answer test against the posterior table recorded in the notebook.
#pragma once
#include "mmd_model_fhn.cuh"

struct FhnNotebookModel : FhnModel {
  static constexpr int MODEL_ID = 2;

  // z = generate_z(u) (notebook cell 18): [exp(.5 u0 - 1), exp(.5 u1 - 2), .5 u2 + 1, .5 u3 + 1]
  MMD_HD static void gen_z(const double* u, double* z, double* dzdu /* Z x Z */) {
    z[0] = exp(0.5 * u[0] - 1.0); z[1] = exp(0.5 * u[1] - 2.0); z[2] = 0.5 * u[2] + 1.0; z[3] = 0.5 * u[3] + 1.0;
    for (int i = 0; i < 16; ++i) dzdu[i] = 0.0;
    dzdu[0] = 0.5 * z[0]; dzdu[5] = 0.5 * z[1]; dzdu[10] = 0.5; dzdu[15] = 0.5;
  }
  // extra[j'] = sum_{m,j} Gam[m*Z+j] * d2 z_m / du_j du_j'   (second derivative of generate_z)
  MMD_HD static void gen_z_second(const double* u, const double* z, const double* Gam, double* extra) {
    extra[0] = 0.25 * Gam[0] * z[0]; extra[1] = 0.25 * Gam[5] * z[1]; extra[2] = 0.0; extra[3] = 0.0;
  }
  // x_0 = generate_x_0(z, v_0) = [-.5, -.5] + v_0 (notebook cell 18); d/dv_0 = I, d/dz = 0
  MMD_HD static void gen_x0(const double* z, const double* v0, double* x0) {
    x0[0] = v0[0] - 0.5; x0[1] = v0[1] - 0.5;
  }
  MMD_HD static void gen_x0_jac(const double* z, double* dx0_dv0 /* X x V0 */, double* dx0_dz /* X x Z */) {
    dx0_dv0[0] = 1.0; dx0_dv0[1] = 0.0; dx0_dv0[2] = 0.0; dx0_dv0[3] = 1.0;
    for (int i = 0; i < 8; ++i) dx0_dz[i] = 0.0;
  }
};
