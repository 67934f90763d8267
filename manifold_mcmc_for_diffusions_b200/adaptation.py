"""Batched coarse search for the initial integrator step size of every chain: Mici 0.1.10
``DualAveragingStepSizeAdapter._find_and_set_init_step_size`` (the adapters are created at
``scripts/utils.py:303-306`` / notebook cell 43 without a step size, so Mici runs this search before warm-up).

Starting from ``init`` the step size of a chain is halved / doubled until one constrained leapfrog step from the
chain's state changes the Hamiltonian by more / less than log 2 on the other side of the first trial (a failed step
counts as "too big").  Every trial step of every chain runs on the device; the host keeps one step size and two flags
per chain."""

import numpy as np

_Q0, _P0 = 0, 1


def find_init_step_sizes(chains, seed, it, init=1.0, max_iters=100):
    """Returns (step_sizes [n_chains], found [n_chains] bool).  The chains' positions are left unchanged; momenta
    are refreshed with the Philox stream (seed, it).  Mici raises AdaptationError for a chain whose search does not
    end within ``max_iters`` trials: here such chains have ``found == False``."""
    bc, n = chains, chains.n_chains
    Q, P = bc.VEC_Q, bc.VEC_P
    thr = np.log(2.0)
    bc.aux_reserve(2)
    bc.transition_begin(seed, it)
    h0 = bc.hamiltonian()
    bc.vec_axpby(_Q0, Q)
    bc.vec_axpby(_P0, P)
    eps = np.full(n, float(init))
    too_big = np.zeros(n, dtype=bool)
    found = np.zeros(n, dtype=bool)
    searching = np.isfinite(h0)
    for s in range(max_iters):
        active = searching & ~found
        if not active.any():
            break
        bc.set_inactive(~active)
        bc.set_step_sizes(eps)
        bc.transition_steps(1.0, 1)
        status = bc.step_info()["status"]
        err = active & ((status & 7) != 0)
        ok = active & ~err
        with np.errstate(invalid="ignore"):
            dh = np.abs(h0 - bc.hamiltonian())
        nan = np.isnan(dh)
        if s == 0:
            too_big[ok] = (nan | (dh > thr))[ok]
        else:
            too_big[ok & nan] = True
        with np.errstate(invalid="ignore"):
            hit = ok & ((too_big & (dh <= thr)) | (~too_big & (dh > thr)))
        found |= hit
        too_big[err] = True
        halve = (ok & ~hit & too_big) | err
        double = ok & ~hit & ~too_big
        eps[halve] *= 0.5
        eps[double] *= 2.0
        # back to the initial state for the next trial
        bc.set_inactive(None, clear_errors=True)
        bc.vec_axpby(Q, _Q0)
        bc.vec_axpby(P, _P0)
        bc.relinearize()
    bc.set_inactive(None, clear_errors=True)
    bc.set_step_sizes(None)
    return eps, found
