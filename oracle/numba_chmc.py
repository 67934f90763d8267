"""Compiled CPU restatement of the constrained-HMC hot path for ONE chain (FHN, noiseless observations).

TEST INFRASTRUCTURE / CPU BASELINE ONLY: imported by tests/, bench.py's `cpu_baseline` / `--impl reference`
legs and tools/; never by the product package.

What it restates (reference `sde/mici_extensions.py`, line numbers of the reference):
  * `constr` :473-519, `jacob_constr_blocks` :521-624, `chol_gram_blocks` :626-687, `log_det_sqrt_gram` :800-820,
    `grad_log_det_sqrt_gram` :1143-1146, `lmult/rmult_by_jacob_constr` :822-913, `lmult_by_inv_gram` :915-942,
    `normal_space_component` :983-993, `quasi_newton_projection` :1009-1063 + wrapper :1323-1402,
    `h1 .. dh2_flow_dmom` :1186-1238, `generate_x_obs_seq` :384-397, partition logic :317-351;
  * Mici 0.1.10 `ConstrainedLeapfrogIntegrator.step` order, the same as oracle/torch_oracle.py.

The reference gets its Jacobians and the gradient of the log-determinant from `jax.jacrev` / `jax.value_and_grad` and
runs them through XLA:CPU; autodiff is not available to a numba loop, so the derivatives here use the block
structure directly (per-step transition matrices, adjoint recursions, a second-order adjoint for the gradient of the
log-determinant) -- the same mathematics as the CUDA kernels, written independently as scalar loops, and checked
against the autodiff oracle in tests/test_numba_oracle.py.  It is therefore a *faster* algorithm on the CPU than the
reference's dense reverse-mode sweeps: as a baseline it errs on the side of the CPU.
"""

import math
import os
import sys

import numpy as np
from numba import njit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle._gen_fhn_numba import fhn_fv, fhn_fx, fhn_fz, fhn_hess_contract, fhn_step  # noqa: E402

NRM = 6  # max constraint rows per block of the FHN model: (R - 1) observations + the conditioned 2-state, R <= 5 ...


def partition_layout(T, R):
    """Block offsets / sizes of the two partitions (mici_extensions.py:321-351)."""
    outs = []
    for init in (R, R // 2):
        if init <= 0 or T <= init:
            sizes = [T]
        else:
            nfull, nrem = divmod(T - init, R)
            nmid = nfull - 1 if nrem == 0 else nfull
            fin = R if nrem == 0 else nrem
            sizes = [init] + [R] * nmid + [fin]
        o = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
        outs.append((o, np.asarray(sizes, dtype=np.int64)))
    return outs


@njit(cache=True)
def _params(q):
    z = np.empty(4)
    z[0] = math.exp(q[0]); z[1] = math.exp(q[1]); z[2] = math.exp(q[2]); z[3] = q[3]
    dz = np.empty(4)
    dz[0] = z[0]; dz[1] = z[1]; dz[2] = z[2]; dz[3] = 1.0
    return z, dz


@njit(cache=True)
def generate_x_obs_seq(q, T, S, dl):
    z, dz = _params(q)
    x0 = q[4]; x1 = q[5] - z[3]
    out = np.empty((T, 2))
    for k in range(T):
        for t in range(S):
            i = 6 + 2 * (k * S + t)
            x0, x1 = fhn_step(z[0], z[1], z[2], z[3], x0, x1, q[i], q[i + 1], dl)
        out[k, 0] = x0; out[k, 1] = x1
    return out


@njit(cache=True)
def n_rows(bo, bn):
    return int(bn.sum()) + bo.shape[0] - 1


@njit(cache=True)
def constr(q, xobs, y, bo, bn, S, dl):
    z, dz = _params(q)
    nb = bo.shape[0]
    c = np.empty(n_rows(bo, bn))
    r = 0
    for b in range(nb):
        o = bo[b]; n = bn[b]; fin = b == nb - 1
        if b == 0:
            x0 = q[4]; x1 = q[5] - z[3]
        else:
            x0 = xobs[o - 1, 0]; x1 = xobs[o - 1, 1]
        for k in range(n):
            for t in range(S):
                i = 6 + 2 * ((o + k) * S + t)
                x0, x1 = fhn_step(z[0], z[1], z[2], z[3], x0, x1, q[i], q[i + 1], dl)
            if fin or k < n - 1:
                c[r] = x0 - y[o + k]; r += 1
            else:
                c[r] = x0 - xobs[o + k, 0]; c[r + 1] = x1 - xobs[o + k, 1]; r += 2
    return c


@njit(cache=True)
def linearize(q, xobs, y, bo, bn, S, dl):
    """Compressed Jacobian (per-step K_t = Psi_{t+1} B_t, per-interval transition matrices), block Gram matrices,
    their Cholesky factors, the capacitance matrix C = I + sum_b A_b^T D_b^-1 A_b and log det^{1/2}."""
    z, dz = _params(q)
    nb = bo.shape[0]
    T = xobs.shape[0]
    N = T * S
    xs = np.zeros((N, 2)); K = np.zeros((N, 2, 2)); Psib = np.zeros((T, 2, 2)); Q = np.zeros((T, 2, 2))
    Zt = np.zeros((T, 2, 4))
    A = np.zeros((nb, NRM, 4)); Dinv = np.zeros((nb, NRM, NRM)); DinvA = np.zeros((nb, NRM, 4))
    nr = np.zeros(nb, dtype=np.int64)
    rk = np.zeros((nb, NRM), dtype=np.int64); rh = np.zeros((nb, NRM), dtype=np.int64)
    C = np.eye(4)
    ld = 0.0
    for b in range(nb):
        o = bo[b]; n = bn[b]; fin = b == nb - 1; ini = b == 0
        if ini:
            x0 = q[4]; x1 = q[5] - z[3]
        else:
            x0 = xobs[o - 1, 0]; x1 = xobs[o - 1, 1]
        for k in range(n):
            g0 = (o + k) * S
            for t in range(S):
                xs[g0 + t, 0] = x0; xs[g0 + t, 1] = x1
                i = 6 + 2 * (g0 + t)
                x0, x1 = fhn_step(z[0], z[1], z[2], z[3], x0, x1, q[i], q[i + 1], dl)
            P00 = 1.0; P01 = 0.0; P10 = 0.0; P11 = 1.0
            for t in range(S - 1, -1, -1):
                i = 6 + 2 * (g0 + t)
                a = xs[g0 + t, 0]; bb = xs[g0 + t, 1]
                b00, b01, b10, b11 = fhn_fv(z[0], z[1], z[2], z[3], a, bb, q[i], q[i + 1], dl)
                G = fhn_fz(z[0], z[1], z[2], z[3], a, bb, q[i], q[i + 1], dl)
                f00, f01, f10, f11 = fhn_fx(z[0], z[1], z[2], z[3], a, bb, q[i], q[i + 1], dl)
                k00 = P00 * b00 + P01 * b10; k01 = P00 * b01 + P01 * b11
                k10 = P10 * b00 + P11 * b10; k11 = P10 * b01 + P11 * b11
                K[g0 + t, 0, 0] = k00; K[g0 + t, 0, 1] = k01; K[g0 + t, 1, 0] = k10; K[g0 + t, 1, 1] = k11
                Q[o + k, 0, 0] += k00 * k00 + k01 * k01; Q[o + k, 0, 1] += k00 * k10 + k01 * k11
                Q[o + k, 1, 0] += k10 * k00 + k11 * k01; Q[o + k, 1, 1] += k10 * k10 + k11 * k11
                for m in range(4):
                    Zt[o + k, 0, m] += P00 * G[m] + P01 * G[4 + m]
                    Zt[o + k, 1, m] += P10 * G[m] + P11 * G[4 + m]
                n00 = P00 * f00 + P01 * f10; n01 = P00 * f01 + P01 * f11
                n10 = P10 * f00 + P11 * f10; n11 = P10 * f01 + P11 * f11
                P00 = n00; P01 = n01; P10 = n10; P11 = n11
            Psib[o + k, 0, 0] = P00; Psib[o + k, 0, 1] = P01; Psib[o + k, 1, 0] = P10; Psib[o + k, 1, 1] = P11
        # rows of the block: (interval, component observed)
        m = 0
        for k in range(n):
            if fin or k < n - 1:
                rk[b, m] = k; rh[b, m] = 0; m += 1
            else:
                rk[b, m] = k; rh[b, m] = 0; rk[b, m + 1] = k; rh[b, m + 1] = 1; m += 2
        nr[b] = m
        Su = np.zeros((2, 4)); P = np.zeros((2, 2))
        if ini:
            Su[1, 3] = -dz[3]; P[0, 0] = 1.0; P[1, 1] = 1.0
        Sus = np.zeros((n, 2, 4)); Ps = np.zeros((n, 2, 2))
        for k in range(n):
            Su = Psib[o + k] @ Su + Zt[o + k] * dz
            P = Psib[o + k] @ P @ Psib[o + k].T + Q[o + k]
            Sus[k] = Su; Ps[k] = P
        D = np.zeros((m, m))
        for i in range(m):
            for j in range(4):
                A[b, i, j] = Sus[rk[b, i], rh[b, i], j]
        for i in range(m):
            for j in range(m):
                if rk[b, j] <= rk[b, i]:
                    Phi = np.eye(2)
                    for mm in range(rk[b, j] + 1, rk[b, i] + 1):
                        Phi = Psib[o + mm] @ Phi
                    val = (Phi @ Ps[rk[b, j]])[rh[b, i], rh[b, j]]
                    D[i, j] = val; D[j, i] = val
        L = np.linalg.cholesky(D)
        for i in range(m):
            ld += math.log(L[i, i])
        Di = np.linalg.inv(D)
        Dinv[b, :m, :m] = Di
        DA = Di @ A[b, :m, :]
        DinvA[b, :m, :] = DA
        C += A[b, :m, :].T @ DA
    LC = np.linalg.cholesky(C)
    for i in range(4):
        ld += math.log(LC[i, i])
    Cinv = np.linalg.inv(C)
    return xs, K, Psib, Q, Zt, A, Dinv, DinvA, nr, rk, rh, Cinv, ld


@njit(cache=True)
def inv_gram(A, Dinv, DinvA, nr, Cinv, c):
    """lmult_by_inv_gram (:915-942): Woodbury solve with the block factors."""
    nb = nr.shape[0]
    lam = np.empty_like(c)
    g = np.zeros(4)
    r = 0
    for b in range(nb):
        m = nr[b]
        t = Dinv[b, :m, :m] @ c[r:r + m]
        lam[r:r + m] = t
        g += A[b, :m, :].T @ t
        r += m
    sv = Cinv @ g
    r = 0
    for b in range(nb):
        m = nr[b]
        lam[r:r + m] -= DinvA[b, :m, :] @ sv
        r += m
    return lam


@njit(cache=True)
def jt(K, Psib, A, nr, rk, rh, bo, bn, S, lam, dim_q):
    """rmult_by_jacob_constr (:879-913) in compressed form."""
    out = np.zeros(dim_q)
    nb = nr.shape[0]
    r = 0
    for b in range(nb):
        o = bo[b]; n = bn[b]; m = nr[b]
        for j in range(4):
            s = 0.0
            for i in range(m):
                s += A[b, i, j] * lam[r + i]
            out[j] += s
        a0 = 0.0; a1 = 0.0
        for k in range(n - 1, -1, -1):
            if k < n - 1:
                P = Psib[o + k + 1]
                t0 = P[0, 0] * a0 + P[1, 0] * a1; t1 = P[0, 1] * a0 + P[1, 1] * a1
                a0 = t0; a1 = t1
            for i in range(m):
                if rk[b, i] == k:
                    if rh[b, i] == 0:
                        a0 += lam[r + i]
                    else:
                        a1 += lam[r + i]
            g0 = (o + k) * S
            for t in range(S):
                Kt = K[g0 + t]
                out[6 + 2 * (g0 + t)] = Kt[0, 0] * a0 + Kt[1, 0] * a1
                out[7 + 2 * (g0 + t)] = Kt[0, 1] * a0 + Kt[1, 1] * a1
        if b == 0:
            P = Psib[o]
            out[4] = P[0, 0] * a0 + P[1, 0] * a1
            out[5] = P[0, 1] * a0 + P[1, 1] * a1
        r += m
    return out


@njit(cache=True)
def jv(K, Psib, A, nr, rk, rh, bo, bn, S, p, n_c):
    """lmult_by_jacob_constr (:822-877) in compressed form."""
    out = np.zeros(n_c)
    nb = nr.shape[0]
    r = 0
    for b in range(nb):
        o = bo[b]; n = bn[b]; m = nr[b]
        for i in range(m):
            s = 0.0
            for j in range(4):
                s += A[b, i, j] * p[j]
            out[r + i] = s
        if b == 0:
            m0 = p[4]; m1 = p[5]
        else:
            m0 = 0.0; m1 = 0.0
        for k in range(n):
            g0 = (o + k) * S
            s0 = 0.0; s1 = 0.0
            for t in range(S):
                Kt = K[g0 + t]
                pa = p[6 + 2 * (g0 + t)]; pb = p[7 + 2 * (g0 + t)]
                s0 += Kt[0, 0] * pa + Kt[0, 1] * pb
                s1 += Kt[1, 0] * pa + Kt[1, 1] * pb
            P = Psib[o + k]
            t0 = P[0, 0] * m0 + P[0, 1] * m1 + s0; t1 = P[1, 0] * m0 + P[1, 1] * m1 + s1
            m0 = t0; m1 = t1
            for i in range(m):
                if rk[b, i] == k:
                    out[r + i] += m0 if rh[b, i] == 0 else m1
        r += m
    return out


@njit(cache=True)
def grad_log_det(q, lin, bo, bn, S, dl):
    """grad_log_det_sqrt_gram (:1143-1146) by the block-local second-order adjoint."""
    xs, K, Psib, Q, Zt, A, Dinv, DinvA, nrv, rk, rh, Cinv, ld = lin
    z, dz = _params(q)
    nb = nrv.shape[0]
    out = np.zeros(q.shape[0])
    for b in range(nb):
        o = bo[b]; n = bn[b]; nr = nrv[b]; ini = b == 0
        DA = DinvA[b, :nr, :]
        Om = DA @ Cinv
        E = Dinv[b, :nr, :nr] - Om @ DA.T
        a = np.zeros((nr, n, 2))
        for i in range(nr):
            kr = rk[b, i]
            vec = np.zeros(2); vec[rh[b, i]] = 1.0
            a[i, kr] = vec
            for k in range(kr - 1, -1, -1):
                vec = Psib[o + k + 1].T @ vec
                a[i, k] = vec
        beta = np.zeros((nr, n, 2))
        for r in range(nr):
            for s in range(nr):
                for k in range(n):
                    beta[r, k, 0] += E[r, s] * a[s, k, 0]; beta[r, k, 1] += E[r, s] * a[s, k, 1]
        Mk = np.zeros((n, 2, 2)); Lam = np.zeros((n, 4, 2))
        for r in range(nr):
            for k in range(n):
                for xa in range(2):
                    for xb in range(2):
                        Mk[k, xa, xb] += beta[r, k, xa] * a[r, k, xb]
                    for uu in range(4):
                        Lam[k, uu, xa] += Om[r, uu] * a[r, k, xa]
        d0 = np.zeros((nr, 2))
        if ini:
            for i in range(nr):
                b0 = Psib[o].T @ beta[i, 0]
                d0[i, 0] = b0[0]; d0[i, 1] = b0[1] - dz[3] * Om[i, 3]
        dobs = np.zeros((nr, n, 2))
        dprev = d0.copy()
        for j in range(n):
            Zd = Zt[o + j] * dz
            for i in range(nr):
                dprev[i] = Psib[o + j] @ dprev[i] + Q[o + j] @ beta[i, j] + Zd @ Om[i]
                dobs[i, j] = dprev[i]
        gz = np.zeros(4); gam0 = 0.0; gam1 = 0.0
        for k in range(n - 1, -1, -1):
            g0 = (o + k) * S
            Yb = np.zeros((2, 2))
            for i in range(nr):
                if rk[b, i] >= k:
                    for xa in range(2):
                        dst = d0[i, xa] if k == 0 else dobs[i, k - 1, xa]
                        for xb in range(2):
                            Yb[xa, xb] += dst * a[i, k, xb]
            LamZ = np.empty((4, 2))
            for uu in range(4):
                LamZ[uu, 0] = dz[uu] * Lam[k, uu, 0]; LamZ[uu, 1] = dz[uu] * Lam[k, uu, 1]
            Ys = np.zeros((S, 2, 2))
            Y = Yb
            for t in range(S):
                Ys[t] = Y
                i = 6 + 2 * (g0 + t)
                xa_ = xs[g0 + t, 0]; xb_ = xs[g0 + t, 1]
                F = np.array(fhn_fx(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 2)
                Bt = np.array(fhn_fv(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 2)
                Gt = np.array(fhn_fz(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 4)
                Y = F @ Y + Bt @ (K[g0 + t].T @ Mk[k]) + Gt @ LamZ
            Psi = np.eye(2)
            for t in range(S - 1, -1, -1):
                i = 6 + 2 * (g0 + t)
                xa_ = xs[g0 + t, 0]; xb_ = xs[g0 + t, 1]
                F = np.array(fhn_fx(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 2)
                Bt = np.array(fhn_fv(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 2)
                Gt = np.array(fhn_fz(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl)).reshape(2, 4)
                Kt = Psi @ Bt
                T1 = Ys[t] @ Psi; T2 = Kt.T @ Mk[k] @ Psi; T3 = LamZ @ Psi
                g = fhn_hess_contract(z[0], z[1], z[2], z[3], xa_, xb_, q[i], q[i + 1], dl,
                                      T1[0, 0], T1[0, 1], T1[1, 0], T1[1, 1], T2[0, 0], T2[0, 1], T2[1, 0], T2[1, 1],
                                      T3[0, 0], T3[0, 1], T3[1, 0], T3[1, 1], T3[2, 0], T3[2, 1], T3[3, 0], T3[3, 1])
                out[i] = Bt[0, 0] * gam0 + Bt[1, 0] * gam1 + g[2]
                out[i + 1] = Bt[0, 1] * gam0 + Bt[1, 1] * gam1 + g[3]
                for m in range(4):
                    gz[m] += Gt[0, m] * gam0 + Gt[1, m] * gam1 + g[4 + m]
                n0 = F[0, 0] * gam0 + F[1, 0] * gam1 + g[0]; n1 = F[0, 1] * gam0 + F[1, 1] * gam1 + g[1]
                gam0 = n0; gam1 = n1
                Psi = Psi @ F
        if ini:
            out[4] += gam0; out[5] += gam1; gz[3] += -gam1
        for m in range(4):
            out[m] += dz[m] * gz[m]
        for m in range(3):
            s = 0.0
            for r in range(nr):
                s += Om[r, m] * A[b, r, m]
            out[m] += s
    return out


@njit(cache=True)
def project_momentum(p, lin, bo, bn, S, n_c):
    """project_onto_cotangent_space (:1252-1254) with the identity metric."""
    xs, K, Psib, Q, Zt, A, Dinv, DinvA, nr, rk, rh, Cinv, ld = lin
    r = jv(K, Psib, A, nr, rk, rh, bo, bn, S, p, n_c)
    lam = inv_gram(A, Dinv, DinvA, nr, Cinv, r)
    return p - jt(K, Psib, A, nr, rk, rh, bo, bn, S, lam, p.shape[0])


@njit(cache=True)
def quasi_newton_projection(q, xobs, y, lin_prev, bo, bn, S, dl, ctol, ptol, dtol, max_iters):
    """quasi_newton_projection (:1009-1063): returns (q, mu, iterations, |dq|, |c|)."""
    xs, K, Psib, Q, Zt, A, Dinv, DinvA, nr, rk, rh, Cinv, ld = lin_prev
    mu = np.zeros_like(q)
    i = 0; ndq = np.inf; err = -1.0
    while True:
        diverged = err > dtol or err != err
        converged = err < ctol and ndq < ptol and err >= 0.0
        if i >= max_iters or diverged or converged:
            break
        c = constr(q, xobs, y, bo, bn, S, dl)
        err = np.max(np.abs(c))
        dq = jt(K, Psib, A, nr, rk, rh, bo, bn, S, inv_gram(A, Dinv, DinvA, nr, Cinv, c), q.shape[0])
        mu = mu + dq
        q = q - dq
        ndq = np.max(np.abs(dq))
        i += 1
    return q, mu, i, ndq, err


@njit(cache=True)
def leapfrog_step(q, p, xobs, y, lin, grad, bo, bn, S, dl, dt, gaussian, ctol, ptol, dtol, max_iters, rev_tol):
    """ConstrainedLeapfrogIntegrator.step, n_inner_step = 1 (Mici 0.1.10 order A(dt/2) B(dt) A(dt/2)).
    Returns (status, q, p, lin, grad, n_fwd, n_back): status 0 ok, 1 not converged / diverged, 4 non-reversible."""
    n_c = n_rows(bo, bn)
    qc = 0.0 if gaussian else 1.0
    p = p - 0.5 * dt * (qc * q + grad)
    p = project_momentum(p, lin, bo, bn, S, n_c)
    if gaussian:
        cs = math.cos(dt); sn = math.sin(dt)
        q_ = cs * q + sn * p; p_ = cs * p - sn * q
        mom_coef = cs / sn
    else:
        q_ = q + dt * p; p_ = p
        mom_coef = 1.0 / dt
    q_new, mu, n_fwd, ndq, err = quasi_newton_projection(q_, xobs, y, lin, bo, bn, S, dl, ctol, ptol, dtol, max_iters)
    if not (err < ctol and ndq < ptol):
        return 1, q, p, lin, grad, n_fwd, 0
    p_new = p_ - mom_coef * mu
    lin_new = linearize(q_new, xobs, y, bo, bn, S, dl)
    p_new = project_momentum(p_new, lin_new, bo, bn, S, n_c)
    if gaussian:
        q_b = cs * q_new - sn * p_new
    else:
        q_b = q_new - dt * p_new
    q_back, mu_b, n_back, ndq_b, err_b = quasi_newton_projection(q_b, xobs, y, lin_new, bo, bn, S, dl, ctol, ptol, dtol,
                                                                 max_iters)
    if not (err_b < ctol and ndq_b < ptol):
        return 1, q, p, lin, grad, n_fwd, n_back
    if np.max(np.abs(q_back - q)) > rev_tol:
        return 4, q, p, lin, grad, n_fwd, n_back
    grad_new = grad_log_det(q_new, lin_new, bo, bn, S, dl)
    p_new = p_new - 0.5 * dt * (qc * q_new + grad_new)
    p_new = project_momentum(p_new, lin_new, bo, bn, S, n_c)
    return 0, q_new, p_new, lin_new, grad_new, n_fwd, n_back


@njit(cache=True)
def hamiltonian(q, p, lin):
    return 0.5 * np.dot(q, q) + lin[12] + 0.5 * np.dot(p, p)


class NumbaChain:
    """One FHN noiseless chain driven like the batched device sampler: static-trajectory constrained HMC with
    momentum refresh, Metropolis accept and the partition switch (`SwitchPartitionTransition`, :1262-1282)."""

    def __init__(self, T, S, R, y, obs_interval, gaussian=False, ctol=1e-9, ptol=1e-8, dtol=1e10, max_iters=50,
                 rev_tol=2e-8):
        self.T, self.S, self.R = T, S, R
        self.y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(T))
        self.dl = obs_interval / S
        self.parts = partition_layout(T, R)
        self.gaussian = bool(gaussian)
        self.tol = (ctol, ptol, dtol, max_iters, rev_tol)
        self.partition = 0
        self.n_steps_ok = 0

    def set_state(self, q, xobs, partition=0, p=None):
        self.q = np.ascontiguousarray(q, dtype=np.float64)
        self.xobs = np.ascontiguousarray(xobs, dtype=np.float64)
        self.partition = partition
        self._relinearize()
        self.p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)

    def _relinearize(self):
        bo, bn = self.parts[self.partition]
        self.lin = linearize(self.q, self.xobs, self.y, bo, bn, self.S, self.dl)
        self.grad = grad_log_det(self.q, self.lin, bo, bn, self.S, self.dl)

    def step(self, dt):
        bo, bn = self.parts[self.partition]
        ctol, ptol, dtol, mi, rt = self.tol
        st, q, p, lin, grad, nf, nbk = leapfrog_step(self.q, self.p, self.xobs, self.y, self.lin, self.grad, bo, bn,
                                                     self.S, self.dl, dt, self.gaussian, ctol, ptol, dtol, mi, rt)
        if st == 0:
            self.q, self.p, self.lin, self.grad = q, p, lin, grad
            self.n_steps_ok += 1
        return st, nf, nbk

    def refresh_momentum(self, rng):
        bo, bn = self.parts[self.partition]
        self.p = project_momentum(rng.standard_normal(self.q.shape[0]), self.lin, bo, bn, self.S, n_rows(bo, bn))

    def hmc_transition(self, dt, n_steps, rng, switch_partition=True):
        self.refresh_momentum(rng)
        keep = (self.q, self.p, self.lin, self.grad)
        h0 = hamiltonian(self.q, self.p, self.lin)
        ok = True
        done = 0
        for _ in range(n_steps):
            st, _, _ = self.step(dt)
            if st != 0:
                ok = False
                break
            done += 1
        acc = False
        if ok:
            h1 = hamiltonian(self.q, self.p, self.lin)
            acc = np.log(rng.uniform()) < h0 - h1
        if not acc:
            self.q, self.p, self.lin, self.grad = keep
        if switch_partition:
            self.xobs = generate_x_obs_seq(self.q, self.T, self.S, self.dl)
            self.partition = 1 - self.partition
            self._relinearize()
        return acc, done


def linear_interpolation_init(T, S, y, obs_interval, rng, u=None, v_0=None):
    """find_initial_state_by_linear_interpolation (:1479-1547) for FHN: per step, solve the (linear in v) step map
    for the noise that moves x to the interpolated target."""
    dl = obs_interval / S
    u = rng.standard_normal(4) if u is None else np.asarray(u, dtype=np.float64)
    v_0 = rng.standard_normal(2) if v_0 is None else np.asarray(v_0, dtype=np.float64)
    z = np.array([math.exp(u[0]), math.exp(u[1]), math.exp(u[2]), u[3]])
    x0 = v_0 - np.array([0.0, z[3]])
    xobs = np.concatenate((np.asarray(y, dtype=np.float64).reshape(T, 1), 0.5 * rng.standard_normal((T, 1))), -1)
    q = np.empty(6 + 2 * T * S)
    q[:4] = u; q[4:6] = v_0
    _interp_fill(q, xobs, x0, z, S, dl)
    return q, xobs


@njit(cache=True)
def _interp_fill(q, xobs, x0, z, S, dl):
    T = xobs.shape[0]
    sa = x0[0]; sb = x0[1]
    for k in range(T):
        da = (xobs[k, 0] - sa) / S; db = (xobs[k, 1] - sb) / S
        for t in range(S):
            xa = sa + t * da; xb = sb + t * db
            m0, m1 = fhn_step(z[0], z[1], z[2], z[3], xa, xb, 0.0, 0.0, dl)
            b00, b01, b10, b11 = fhn_fv(z[0], z[1], z[2], z[3], xa, xb, 0.0, 0.0, dl)
            det = b00 * b11 - b01 * b10
            r0 = da - (m0 - xa); r1 = db - (m1 - xb)
            i = 6 + 2 * (k * S + t)
            q[i] = (b11 * r0 - b01 * r1) / det
            q[i + 1] = (-b10 * r0 + b00 * r1) / det
        sa = xobs[k, 0]; sb = xobs[k, 1]
