// Host-side (g++) build of the generated model functors so their arithmetic can be unit-tested
// on a machine without a GPU (tests/test_model_functor_host.py).  Not part of the product path.
#include "mmd_model_fhn.cuh"
#include "mmd_philox.cuh"
#define MMD_GEN_MAX_HOST 24

namespace {
FhnModel::Coef coef(const double* z, double sd) {
  FhnModel::Coef c;
  FhnModel::make_coef(z, sd, c);
  return c;
}
}  // namespace

extern "C" {
void fhn_step(const double* z, double sd, const double* x, const double* v, double* xn) { FhnModel::step(coef(z, sd), x, v, xn); }
void fhn_jac_x(const double* z, double sd, const double* x, const double* v, double* F) { FhnModel::jac_x(coef(z, sd), x, v, F); }
void fhn_jac_v(const double* z, double sd, const double* x, const double* v, double* B) { FhnModel::jac_v(coef(z, sd), x, v, B); }
void fhn_jac_z(const double* z, double sd, const double* x, const double* v, double* G) { FhnModel::jac_z(coef(z, sd), x, v, G); }
void fhn_hess_contract(const double* z, double sd, const double* x, const double* v, const double* Th, double* g) {
  FhnModel::hess_contract(coef(z, sd), x, v, Th, g);
}
void fhn_gen_z(const double* u, double* z, double* dzdu) {
  double gp[MMD_GEN_MAX_HOST];
  FhnModel::default_gen(gp);
  FhnModel::gen_z(gp, u, z, dzdu);
}
void fhn_gen_all(const double* gp, const double* u, const double* v0, const double* Gam, double* z, double* dzdu, double* extra, double* x0,
                 double* dx0_dv0, double* dx0_dz) {
  FhnModel::gen_z(gp, u, z, dzdu);
  FhnModel::gen_z_second(gp, u, z, Gam, extra);
  FhnModel::gen_x0(gp, z, v0, x0);
  FhnModel::gen_x0_jac(gp, z, dx0_dv0, dx0_dz);
}
void philox_normal_pair(unsigned long long seed, unsigned long long offset, unsigned long long idx, double* a, double* b) {
  mmd::philox_normal_pair(seed, offset, idx, a, b);
}
}

#include "mmd_model_sir.cuh"
namespace {
SirModel::Coef scoef(const double* z, double sd) {
  SirModel::Coef c;
  SirModel::make_coef(z, sd, c);
  return c;
}
}  // namespace
extern "C" {
void sir_step(const double* z, double sd, const double* x, const double* v, double* xn) { SirModel::step(scoef(z, sd), x, v, xn); }
void sir_jac_x(const double* z, double sd, const double* x, const double* v, double* F) { SirModel::jac_x(scoef(z, sd), x, v, F); }
void sir_jac_v(const double* z, double sd, const double* x, const double* v, double* B) { SirModel::jac_v(scoef(z, sd), x, v, B); }
void sir_jac_z(const double* z, double sd, const double* x, const double* v, double* G) { SirModel::jac_z(scoef(z, sd), x, v, G); }
void sir_hess_contract(const double* z, double sd, const double* x, const double* v, const double* Th, double* g) {
  SirModel::hess_contract(scoef(z, sd), x, v, Th, g);
}
void sir_gen_z(const double* u, double* z, double* dzdu) { SirModel::gen_z(nullptr, u, z, dzdu); }
void sir_gen_z_second(const double* u, const double* z, const double* Gam, double* extra) { SirModel::gen_z_second(nullptr, u, z, Gam, extra); }
}
