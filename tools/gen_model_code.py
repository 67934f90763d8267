#!/usr/bin/env python
"""Generate the CUDA device functors for the example diffusion models from SymPy expressions.

The reference builds its one-step maps symbolically (``sde/integrators.py`` through SymNum,
``sde/example_models/fhn.py:27-34``, ``sir.py:39-51``) and leaves every derivative to JAX autodiff.
Here the same symbolic step (re-derived in ``oracle/models.py:derive_*_step``) is differentiated
symbolically and printed, after common-subexpression elimination, as ``__host__ __device__``
functions: the step, its Jacobians with respect to state / noise / parameters, and the
second-order contraction ``sum_i Hess(f_i)[:, :] @ Theta[:, i]`` that the log-determinant gradient
sweep needs ("second-order adjoint").

Every sub-expression that depends only on the parameters ``z`` and the step size is hoisted into a
per-thread coefficient struct (``Coef``), computed once per sweep by ``make_coef``: the time-stepping
loops then contain no division, no power and no re-evaluation of parameter-only products.

Usage:  python tools/gen_model_code.py   (rewrites manifold_mcmc_for_diffusions_b200/csrc/mmd_model_*.cuh)
"""

import os
import sys

import sympy as sp
from sympy.printing.c import C99CodePrinter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT_DIR = os.path.join(ROOT, "manifold_mcmc_for_diffusions_b200", "csrc")


class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= abs(int(e)) <= 8:
            s = self._print(b)
            if not b.is_Atom:
                s = f"({s})"
            prod = "(" + "*".join([s] * abs(int(e))) + ")"
            return prod if int(e) > 0 else f"(1.0/{prod})"
        if e == -1:
            return f"(1.0/({self._print(b)}))"
        if b.is_Integer and e == sp.Rational(1, 2):
            return repr(float(sp.sqrt(b)))
        return super()._print_Pow(expr)


_pr = _Printer()


class Hoister:
    """Replaces maximal parameter-only sub-expressions by coefficient symbols k<i>."""

    def __init__(self, dyn_syms, prefix, recip_only=False):
        self.dyn = set(dyn_syms)
        self.table = {}
        self.prefix = prefix
        self.recip_only = recip_only  # hoist only sub-expressions that contain a division

    def _const(self, e):
        return not (e.free_symbols & self.dyn)

    def _sym(self, c):
        if not c.free_symbols:
            return c
        if self.recip_only:
            if not any(isinstance(a, sp.Pow) and a.exp.is_negative for a in sp.preorder_traversal(c)):
                return c
            # hoist just the reciprocal factors, keep the rest inline
            if isinstance(c, sp.Mul):
                rec = [a for a in c.args if isinstance(a, sp.Pow) and a.exp.is_negative]
                rest = [a for a in c.args if not (isinstance(a, sp.Pow) and a.exp.is_negative)]
                if rec and rest:
                    return sp.Mul(self._sym(sp.Mul(*rec)), *rest)
            if isinstance(c, sp.Add):
                return sp.Add(*[self._sym(a) for a in c.args])
        if c not in self.table:
            self.table[c] = sp.Symbol(f"{self.prefix}{len(self.table)}")
        return self.table[c]

    def __call__(self, e):
        if e.is_Atom:
            return e
        if self._const(e):
            return self._sym(e)
        if isinstance(e, (sp.Mul, sp.Add)):
            cargs = [a for a in e.args if self._const(a)]
            dargs = [self(a) for a in e.args if not self._const(a)]
            if cargs:
                return e.func(self._sym(e.func(*cargs)), *dargs)
            return e.func(*dargs)
        return e.func(*[self(a) for a in e.args])


def _emit(name, args_sig, outputs, unpack, hoister):
    """outputs: list of (lhs string, sympy expr).  Returns C source of one function."""
    exprs = [hoister(sp.sympify(e)) for _, e in outputs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    lines = [f"  MMD_HD static void {name}({args_sig}) {{"]
    lines.append(unpack)
    used = set()
    for _, e in repl:
        used |= e.free_symbols
    for e in red:
        used |= sp.sympify(e).free_symbols
    for k in sorted((s for s in used if s in set(hoister.table.values())), key=lambda s: int(s.name[len(hoister.prefix):])):
        lines.append(f"    const double {k} = c.{hoister.prefix}[{k.name[len(hoister.prefix):]}];")
    for s, e in repl:
        lines.append(f"    const double {s} = {_pr.doprint(e)};")
    for (lhs, _), e in zip(outputs, red):
        lines.append(f"    {lhs} = {_pr.doprint(e)};")
    lines.append("  }")
    return "\n".join(lines)


def gen_model(struct_name, fname, f, X, V, Z, sd, header, doc, extra_methods="", step_name="step", deriv_suffix=""):
    """Emit one model header.  f: sympy column of the step map in terms of X, V, Z symbols and sd."""
    nx, nv, nz = len(X), len(V), len(Z)
    dyn = list(X) + list(V)
    joint = list(X) + list(V) + list(Z)
    unpack = (
        "    " + " ".join(f"const double {s} = x[{i}];" for i, s in enumerate(X)) + "\n"
        "    " + " ".join(f"const double {s} = v[{i}];" for i, s in enumerate(V)) + "\n"
        "    " + " ".join(f"(void){s};" for s in dyn) + "\n"
        "    " + " ".join(f"const double {s} = c.z[{i}];" for i, s in enumerate(Z)) + f" const double {sd} = c.sd;\n"
        "    " + " ".join(f"(void){s};" for s in Z) + f" (void){sd};"
    )
    hs = Hoister(dyn, "ks")   # coefficients of the step itself (hot in every sweep)
    hd = Hoister(dyn + [sp.Symbol(f"Th{r}_{c}") for r in range(len(joint)) for c in range(nx)], "kd",
                 recip_only=True)
    sig = "const Coef& c, const double* x, const double* v, "
    parts = []
    parts.append(_emit(step_name, sig + "double* xn", [(f"xn[{i}]", f[i]) for i in range(nx)], unpack, hs))
    Fx = f.jacobian(X)
    parts.append(_emit("jac_x" + deriv_suffix, sig + "double* F",
                       [(f"F[{i*nx+j}]", Fx[i, j]) for i in range(nx) for j in range(nx)], unpack, hd))
    Fv = f.jacobian(V)
    parts.append(_emit("jac_v" + deriv_suffix, sig + "double* B",
                       [(f"B[{i*nv+j}]", Fv[i, j]) for i in range(nx) for j in range(nv)], unpack, hd))
    Fz = f.jacobian(Z)
    parts.append(_emit("jac_z" + deriv_suffix, sig + "double* G",
                       [(f"G[{i*nz+j}]", Fz[i, j]) for i in range(nx) for j in range(nz)], unpack, hd))
    # second-order contraction: g[a] = sum_i sum_b d2 f_i / d y_a d y_b * Th[b*X + i]
    nj = len(joint)
    Th = sp.Matrix(nj, nx, lambda r, c: sp.Symbol(f"Th{r}_{c}"))
    gout = []
    for a in range(nj):
        acc = 0
        for i in range(nx):
            for bb in range(nj):
                acc += sp.diff(f[i], joint[a], joint[bb]) * Th[bb, i]
        gout.append(acc)
    th_unpack = unpack + "\n" + "\n".join(
        "    " + " ".join(f"const double Th{r}_{c} = Th[{r*nx+c}];" for c in range(nx)) for r in range(nj)
    )
    parts.append(_emit("hess_contract" + deriv_suffix, sig + "const double* Th, double* g",
                       [(f"g[{a}]", gout[a]) for a in range(nj)], th_unpack, hd))
    jv_const = not (Fv.free_symbols & set(X) | Fv.free_symbols & set(V))

    def coef_lines(h):
        out = []
        for e, k in sorted(h.table.items(), key=lambda kv: int(kv[1].name[len(h.prefix):])):
            out.append(f"    c.{h.prefix}[{k.name[len(h.prefix):]}] = {_pr.doprint(e)};")
        return "\n".join(out)

    zunpack = " ".join(f"const double {s} = z[{i}];" for i, s in enumerate(Z)) + " " + \
        " ".join(f"(void){s};" for s in Z)
    body = "\n\n".join(parts)
    src = f"""// GENERATED by tools/gen_model_code.py -- do not edit by hand.
{doc}
#pragma once
#include "mmd_common.cuh"

struct {struct_name} {{
{header}
  static constexpr bool JV_CONST = {'true' if jv_const else 'false'};  // d f / d v independent of the state

  // per-thread constants of the step map: every parameter-only sub-expression of the generated code,
  // so no division / power / parameter product is left inside the time-stepping loops
  struct Coef {{
    double z[Z];
    double sd;
    double ks[{max(len(hs.table), 1)}];   // step
    double kd[{max(len(hd.table), 1)}];   // derivatives (jac_x, jac_v, jac_z, hess_contract)
  }};
  MMD_HD static void make_coef_step(const double* z, double {sd}, Coef& c) {{
    for (int i = 0; i < Z; ++i) c.z[i] = z[i];
    c.sd = {sd};
    {zunpack}
{coef_lines(hs)}
  }}
  MMD_HD static void make_coef(const double* z, double {sd}, Coef& c) {{
    make_coef_step(z, {sd}, c);
    {zunpack}
{coef_lines(hd)}
  }}

{body}
{extra_methods}
}};
"""
    if step_name != "step":
        src = src.replace("MMD_HD static void step_clipped(", "MMD_HD static void step(")
    path = os.path.join(OUT_DIR, fname)
    with open(path, "w") as fh:
        fh.write(src)
    print("wrote", path, "step coefficients:", len(hs.table), "derivative coefficients:", len(hd.table))


FHN_HEADER = """  static constexpr int X = 2;   // dim_x   (fhn.py:10)
  static constexpr int V = 2;   // dim_v   (fhn.py:14)
  static constexpr int Z = 4;   // dim_z   (fhn.py:12)  z = [sigma, epsilon, gamma, beta]
  static constexpr int V0 = 2;  // dim_v_0 (fhn.py:13)
  static constexpr int Y = 1;   // dim_y: obs_func(x) = x[0]  (fhn.py:37-38)
  static constexpr int MODEL_ID = 0;

  // Generators with run-time parameters gp[22] (Dims::gen; mmd_set_generator_params):
  //   z_i = a_i u_i + b_i, exponentiated where m_i != 0        gp = [a (4) | b (4) | m (4) | c (2) | E (2 x 4, row-major)]
  //   x_0 = v_0 + c + E z
  // The reference's fhn.py:41-51 (z = [exp u0, exp u1, exp u2, u3], x_0 = v_0 - [0, z3]) is a = 1, b = 0,
  // m = [1, 1, 1, 0], c = 0, E[1][3] = -1 (the default); the notebook's priors are another parameter set.
  static constexpr int NGEN = 22;
  MMD_HD static void default_gen(double* gp) {
    for (int i = 0; i < NGEN; ++i) gp[i] = 0.0;
    gp[0] = gp[1] = gp[2] = gp[3] = 1.0;
    gp[8] = gp[9] = gp[10] = 1.0;
    gp[14 + 1 * 4 + 3] = -1.0;
  }
  MMD_HD static void gen_z(const double* gp, const double* u, double* z, double* dzdu /* Z x Z */) {
    for (int i = 0; i < 16; ++i) dzdu[i] = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double lin = gp[i] * u[i] + gp[4 + i];
      const bool ex = gp[8 + i] != 0.0;
      z[i] = ex ? exp(lin) : lin;
      dzdu[5 * i] = ex ? gp[i] * z[i] : gp[i];
    }
  }
  // extra[j'] = sum_{m,j} Gam[m*Z+j] * d2 z_m / du_j du_j'   (second derivative of generate_z)
  MMD_HD static void gen_z_second(const double* gp, const double* u, const double* z, const double* Gam, double* extra) {
    for (int i = 0; i < 4; ++i) extra[i] = gp[8 + i] != 0.0 ? Gam[5 * i] * gp[i] * gp[i] * z[i] : 0.0;
  }
  MMD_HD static void gen_x0(const double* gp, const double* z, const double* v0, double* x0) {
    for (int i = 0; i < 2; ++i) {
      double s = v0[i] + gp[12 + i];
      for (int j = 0; j < 4; ++j)
        if (gp[14 + 4 * i + j] != 0.0) s += gp[14 + 4 * i + j] * z[j];
      x0[i] = s;
    }
  }
  MMD_HD static void gen_x0_jac(const double* gp, const double* z, double* dx0_dv0 /* X x V0 */, double* dx0_dz /* X x Z */) {
    dx0_dv0[0] = 1.0; dx0_dv0[1] = 0.0; dx0_dv0[2] = 0.0; dx0_dv0[3] = 1.0;
    for (int i = 0; i < 8; ++i) dx0_dz[i] = gp[14 + i];
  }
  // observation y = h(x) = x[0]; gradient e0; zero second derivative
  MMD_HD static double obs(const double* x) { return x[0]; }
  MMD_HD static void obs_grad(const double* x, double* dh) { dh[0] = 1.0; dh[1] = 0.0; }
  MMD_HD static void obs_hess_vec(const double* x, const double* d, double* out) { out[0] = 0.0; out[1] = 0.0; }
  static constexpr bool OBS_LINEAR = true;
"""


# (the notebook prior parametrisation is a hand-written struct derived from FhnModel: csrc/mmd_model_fhn_notebook.cuh)


def gen_fhn():
    from oracle.models import derive_fhn_step

    f, sy = derive_fhn_step(simplify=False)
    d = sy["delta"]
    sd = sp.symbols("sd", positive=True)  # sqrt(delta): keeps pow() out of the generated code
    f = f.subs(d, sd ** 2)
    gen_model(
        "FhnModel", "mmd_model_fhn.cuh", f, list(sy["x"]), list(sy["v"]), list(sy["z"]), "sd", FHN_HEADER,
        "// FitzHugh-Nagumo hypoelliptic diffusion, strong-order-1.5 Taylor step (additive noise).\n"
        "// Restates sde/example_models/fhn.py:10-51 and sde/integrators.py:46-63 of the reference; all\n"
        "// derivatives are symbolic derivatives of that one expression.",
    )


SIR_HEADER = """  static constexpr int X = 3;   // dim_x   (sir.py:9)   state [log S, log I, log contact rate]
  static constexpr int V = 3;   // dim_v = dim_w (sir.py:10-13)
  static constexpr int Z = 4;   // dim_z   z = [beta, gamma, zeta, epsilon]
  static constexpr int V0 = 1;  // dim_v_0
  static constexpr int Y = 1;   // dim_y: obs_func(x) = exp(x[1])  (sir.py:73-74)
  static constexpr int MODEL_ID = 1;

  // z = generate_z(u) (sir.py:77-85): [exp u0, exp u1, u2, exp(sqrt(.75) u3 + .5 u1 - 3)] and dz/du
  MMD_HD static void gen_z(const double* /*gp: no run-time generator parameters*/, const double* u, double* z, double* dzdu /* Z x Z */) {
    z[0] = exp(u[0]); z[1] = exp(u[1]); z[2] = u[2];
    z[3] = exp(0.8660254037844386 * u[3] + 0.5 * u[1] - 3.0);
    for (int i = 0; i < 16; ++i) dzdu[i] = 0.0;
    dzdu[0] = z[0]; dzdu[5] = z[1]; dzdu[10] = 1.0;
    dzdu[13] = 0.5 * z[3]; dzdu[15] = 0.8660254037844386 * z[3];
  }
  // extra[j'] = sum_{m,j} Gam[m*Z+j] * d2 z_m / du_j du_j'
  static constexpr int NGEN = 0;
  MMD_HD static void default_gen(double*) {}
  MMD_HD static void gen_z_second(const double*, const double* u, const double* z, const double* Gam, double* extra) {
    const double a = 0.5, b = 0.8660254037844386;
    extra[0] = Gam[0] * z[0];
    extra[1] = Gam[5] * z[1] + z[3] * (Gam[13] * a * a + Gam[15] * a * b);
    extra[2] = 0.0;
    extra[3] = z[3] * (Gam[13] * a * b + Gam[15] * b * b);
  }
  // x_0 = generate_x_0(z, v_0) = [log 762, log 1, v_0[0]] (sir.py:88-89)
  MMD_HD static void gen_x0(const double*, const double* z, const double* v0, double* x0) {
    x0[0] = 6.635946555686647; x0[1] = 0.0; x0[2] = v0[0];
  }
  MMD_HD static void gen_x0_jac(const double*, const double* z, double* dx0_dv0 /* X x V0 */, double* dx0_dz /* X x Z */) {
    dx0_dv0[0] = 0.0; dx0_dv0[1] = 0.0; dx0_dv0[2] = 1.0;
    for (int i = 0; i < 12; ++i) dx0_dz[i] = 0.0;
  }
  // observation y = h(x) = exp(x[1])
  MMD_HD static double obs(const double* x) { return exp(x[1]); }
  MMD_HD static void obs_grad(const double* x, double* dh) { dh[0] = 0.0; dh[1] = exp(x[1]); dh[2] = 0.0; }
  MMD_HD static void obs_hess_vec(const double* x, const double* d, double* out) {
    out[0] = 0.0; out[1] = exp(x[1]) * d[1]; out[2] = 0.0;
  }
  static constexpr bool OBS_LINEAR = false;
"""


def gen_sir():
    from oracle.models import derive_sir_step

    f, sy = derive_sir_step(simplify=False)
    d = sy["delta"]
    sd = sp.symbols("sd", positive=True)
    f = f.subs(d, sd ** 2)
    gen_model(
        "SirModel", "mmd_model_sir.cuh", f, list(sy["x"]), list(sy["v"]), list(sy["z"]), "sd", SIR_HEADER,
        "// SIR epidemic model with an Ornstein-Uhlenbeck log contact rate, Euler-Maruyama step of the\n"
        "// log-transformed SDE.  Restates sde/example_models/sir.py:9-93, sde/integrators.py:8-14 and\n"
        "// sde/transforms.py:9-63 of the reference (the clip of the first two state components at -500,\n"
        "// sir.py:54-70, is applied by `step_clipped`); all derivatives are symbolic.",
        extra_methods='''
  // forward_func with the reference's guards (sir.py:54-70): clip the first two components below at -500
  // before the step and keep them there afterwards
  MMD_HD static void step_clipped(const Coef& c, const double* x, const double* v, double* xn) {
    double xc[3] = {x[0] < -500.0 ? -500.0 : x[0], x[1] < -500.0 ? -500.0 : x[1], x[2]};
    double xr[3];
    step_raw(c, xc, v, xr);
    xn[0] = xc[0] > -500.0 ? xr[0] : xc[0];
    xn[1] = xc[1] > -500.0 ? xr[1] : xc[1];
    xn[2] = xr[2];
  }

  // Derivatives of the guarded step, as autodiff sees them (clip: zero slope below -500; select: the untouched
  // clipped value): a component held at -500 neither moves nor influences the others -- its row and its column of
  // d f / d x vanish, its rows of d f / d v and d f / d z vanish, and so does every second derivative that involves
  // it.  Everything else is the raw derivative at the clipped state.
  MMD_HD static void clip_state(const double* x, double* xc, bool* held) {
    held[0] = !(x[0] > -500.0); held[1] = !(x[1] > -500.0); held[2] = false;
    xc[0] = held[0] ? -500.0 : x[0]; xc[1] = held[1] ? -500.0 : x[1]; xc[2] = x[2];
  }
  MMD_HD static void jac_x(const Coef& c, const double* x, const double* v, double* F) {
    if (x[0] > -500.0 && x[1] > -500.0) { jac_x_raw(c, x, v, F); return; }
    double xc[3]; bool held[3];
    clip_state(x, xc, held);
    jac_x_raw(c, xc, v, F);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        if (held[i] || held[j]) F[i * 3 + j] = 0.0;
  }
  MMD_HD static void jac_v(const Coef& c, const double* x, const double* v, double* B) {
    if (x[0] > -500.0 && x[1] > -500.0) { jac_v_raw(c, x, v, B); return; }
    double xc[3]; bool held[3];
    clip_state(x, xc, held);
    jac_v_raw(c, xc, v, B);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < V; ++j)
        if (held[i]) B[i * V + j] = 0.0;
  }
  MMD_HD static void jac_z(const Coef& c, const double* x, const double* v, double* G) {
    if (x[0] > -500.0 && x[1] > -500.0) { jac_z_raw(c, x, v, G); return; }
    double xc[3]; bool held[3];
    clip_state(x, xc, held);
    jac_z_raw(c, xc, v, G);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < Z; ++j)
        if (held[i]) G[i * Z + j] = 0.0;
  }
  // g[a] = sum_i sum_b d2 f_i / d y_a d y_b Th[b * X + i],  y = (x, v, z)
  MMD_HD static void hess_contract(const Coef& c, const double* x, const double* v, const double* Th, double* g) {
    if (x[0] > -500.0 && x[1] > -500.0) { hess_contract_raw(c, x, v, Th, g); return; }
    double xc[3]; bool held[3];
    clip_state(x, xc, held);
    double Tm[(X + V + Z) * X];
    for (int b = 0; b < X + V + Z; ++b)
      for (int i = 0; i < X; ++i) Tm[b * X + i] = (held[i] || (b < X && held[b])) ? 0.0 : Th[b * X + i];
    hess_contract_raw(c, xc, v, Tm, g);
    for (int a = 0; a < X; ++a)
      if (held[a]) g[a] = 0.0;
  }
''',
        step_name="step_raw", deriv_suffix="_raw",
    )


if __name__ == "__main__":
    gen_fhn()
    gen_sir()
