"""Batched dynamic-HMC (NUTS) transition for resident chains: the reference's
``mici.transitions.MultinomialDynamicIntegrationTransition`` (call site ``scripts/utils.py:292-301``; Mici 0.1.10
semantics as in SURVEY.md appendix A) for thousands of chains at once.

The host keeps the per-chain tree bookkeeping (a few scalars per chain, NumPy); every state vector lives on the
device and every O(dim_q) operation -- constrained leapfrog steps, Hamiltonians, momentum sums, no-U-turn inner
products, proposal copies -- is a kernel of ``libmmd_b200.so`` (``mmd_vec_*`` entry points).  Trees are built by
iterative doubling with multinomial sampling inside the new sub-tree, biased progressive sampling between the old
tree and the new sub-tree, the no-U-turn criterion on momentum sums (sub-tree checks through momentum / momentum-sum
check-points at the power-of-two leaves, tree check on the two edges), termination on a Hamiltonian error above
``max_delta_h`` (divergence) or an integrator error (convergence / non-reversible step; the current proposal is
kept, like Mici).  All chains start their transition together and advance leaf by leaf; chains whose tree is
finished are parked (``mmd_set_inactive``) until the deepest tree of the batch is complete.
"""

import numpy as np

EQ0, EQ1, EP0, EP1, PROP, SUMP, SUBPROP, SUBSUM, TMP, CK0 = range(10)


def _popcount(x):
    return bin(x).count("1")


def _trailing_ones(x):
    n = 0
    while x & 1:
        n += 1
        x >>= 1
    return n


class BatchedNUTS:
    def __init__(self, chains, max_tree_depth=10, max_delta_h=1000.0, do_extra_subtree_checks=True,
                 error_accept_stat="partial"):
        """error_accept_stat: the `accept_stat` reported for a transition whose tree ended in an integrator error
        (projection not converged / non-reversible step).  "partial" (default) is Mici's rule as recorded in
        SURVEY.md appendix A: sum_acc_prob / n_step over the steps taken before the error, 0 when there were none.
        "zero" reports 0 for every such transition; it is an opt-in used by tools/notebook_nuts.py, which found that it
        reproduces the sampler statistics recorded in the reference's notebook better -- Mici's
        source is not available here, so the documented rule stays the default."""
        if error_accept_stat not in ("zero", "partial"):
            raise ValueError("error_accept_stat must be 'zero' or 'partial'")
        self.error_accept_stat = error_accept_stat
        self.bc = chains
        self.max_tree_depth = D = int(max_tree_depth)
        self.max_delta_h = float(max_delta_h)
        self.do_extra_subtree_checks = bool(do_extra_subtree_checks)
        # check-points per span level: momentum / running momentum sum at the first leaf of the span (ckp, cks), and
        # for the extra checks at the last leaf of its first half (midp, mids) and momentum at the first leaf of
        # its second half (inop)
        self.ckp = [CK0 + i for i in range(D)]
        self.cks = [CK0 + D + i for i in range(D)]
        self.midp = [CK0 + 2 * D + i for i in range(D)]
        self.mids = [CK0 + 3 * D + i for i in range(D)]
        self.inop = [CK0 + 4 * D + i for i in range(D)]
        # proposal stack of the sub-tree being built: Mici's _build_tree merges two sibling sub-trees by keeping the
        # outer one's proposal with probability w_outer / (w_inner + w_outer); with leaves arriving in order the merges
        # of the recursion are the carries of a binary counter, entry i of the stack = proposal of a finished sub-tree
        self.prp = [CK0 + 5 * D + i for i in range(D + 1)]
        chains.aux_reserve(CK0 + 5 * D + D + 1)

    def transition(self, step_size, rng, seed, it, switch_partition=True, trace=None):
        """One momentum refresh + dynamic integration transition (+ partition switch) for every chain.
        step_size: scalar or per-chain array; rng: NumPy Generator for the tree decisions (directions,
        multinomial / progressive sampling); (seed, it) key the on-device Philox momentum draw.
        Random numbers are drawn as one vector per decision in the order Mici's recursion draws them for a single
        chain (direction; one per sub-tree merge, innermost first; top-level acceptance), so chain c driven by column c
        of the draws reproduces the recursive transition (tests/test_gpu_nuts_parity.py).  `trace` (dict) receives the
        state after the momentum refresh.  Returns the per-chain statistics Mici reports."""
        bc, n = self.bc, self.bc.n_chains
        Q, P = bc.VEC_Q, bc.VEC_P
        eps = np.broadcast_to(np.asarray(step_size, dtype=np.float64), (n,)).copy()
        bc.transition_begin(seed, it)
        h0 = bc.hamiltonian()
        if trace is not None:
            trace["state0"] = bc.get_state()
            trace["partition"] = bc.partition
        for a in (EQ0, EQ1, PROP):
            bc.vec_axpby(a, Q)
        for a in (EP0, EP1, SUMP):
            bc.vec_axpby(a, P)
        logw = -h0
        active = np.isfinite(h0)
        last_dir = np.zeros(n, dtype=np.int64)
        n_step = np.zeros(n, dtype=np.int64)
        sum_acc = np.zeros(n)
        depth_reached = np.zeros(n, dtype=np.int64)
        diverging = np.zeros(n, dtype=bool)
        conv_err = np.zeros(n, dtype=bool)
        nonrev = np.zeros(n, dtype=bool)
        for depth in range(self.max_tree_depth):
            if not active.any():
                break
            dirs = np.where(rng.random(n) < 0.5, 1, -1)
            # put the live state of every chain at the edge its new sub-tree grows from
            sw = active & (last_dir != 0) & (dirs != last_dir)
            if sw.any():
                for sgn, eq, ep in ((1, EQ1, EP1), (-1, EQ0, EP0)):
                    m = sw & (dirs == sgn)
                    if m.any():
                        bc.vec_axpby(Q, eq, mask=m)
                        bc.vec_axpby(P, ep, mask=m)
                bc.set_inactive(~active)
                bc.relinearize()
            last_dir = np.where(active, dirs, last_dir)
            bc.set_step_sizes(dirs * eps)
            bc.vec_axpby(SUBSUM, SUBSUM, alpha=0.0, beta=0.0)
            sub_logw = np.full(n, -np.inf)
            stack_lw = np.full((self.max_tree_depth + 1, n), -np.inf)
            in_sub = active.copy()
            for leaf in range(2 ** depth):
                if not in_sub.any():
                    break
                bc.set_inactive(~in_sub)
                bc.transition_steps(1.0, 1)
                st = bc.step_info()["status"]
                err = in_sub & ((st & 7) != 0)
                conv_err |= err & ((st & 3) != 0)
                nonrev |= err & ((st & 4) != 0)
                ok = in_sub & ~err
                h = bc.hamiltonian()
                h = np.where(np.isnan(h), np.inf, h)
                n_step[ok] += 1
                with np.errstate(over="ignore", invalid="ignore"):
                    sum_acc[ok] += np.minimum(1.0, np.exp(h0[ok] - h[ok]))
                div = ok & (h - h0 > self.max_delta_h)
                diverging |= div
                ok &= ~div
                bc.vec_axpby(SUBSUM, P, 1.0, 1.0, mask=ok)
                lw = -h
                sub_logw = np.where(ok, np.logaddexp(sub_logw, lw), sub_logw)
                # push the leaf; the merges follow below, together with the no-U-turn checks of the same spans
                pos = _popcount(leaf)
                bc.vec_axpby(self.prp[pos], Q, mask=ok)
                stack_lw[pos] = np.where(ok, lw, -np.inf)
                turning = np.zeros(n, dtype=bool)
                idx_max, t_ones = _popcount(leaf >> 1), _trailing_ones(leaf)
                extra = self.do_extra_subtree_checks
                if leaf % 2 == 0:
                    bc.vec_axpby(self.ckp[idx_max], P, mask=ok)
                    bc.vec_axpby(self.cks[idx_max], SUBSUM, mask=ok)
                    if extra and leaf > 0:
                        # first leaf of the second half of the span whose first half ended at leaf - 1
                        bc.vec_axpby(self.inop[_trailing_ones(leaf - 1)], P, mask=ok)
                else:
                    # spans of 2, 4, ... leaves ending here (Mici _build_tree merges, innermost first)
                    for k, i in enumerate(range(idx_max, idx_max - t_ones, -1), start=1):
                        # merge the two sibling sub-trees on top of the stack (inner = older, outer = newer)
                        lw_i, lw_o = stack_lw[pos - 1], stack_lw[pos]
                        lw_m = np.logaddexp(lw_i, lw_o)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            take = ok & (np.log(rng.random(n)) < lw_o - lw_m)
                        if take.any():
                            bc.vec_axpby(self.prp[pos - 1], self.prp[pos], mask=take)
                        stack_lw[pos - 1] = lw_m
                        pos -= 1
                        d1, d2 = bc.vec_uturn(self.ckp[i], self.cks[i], SUBSUM, P)
                        turning |= ok & ((d1 < 0) | (d2 < 0))
                        if extra and k >= 2:
                            lv = k - 1
                            # (first leaf, first leaf of 2nd half) with sum(1st half) + that momentum
                            bc.vec_axpby(TMP, self.mids[lv])
                            bc.vec_axpby(TMP, self.inop[lv], 1.0, 1.0)
                            d1, d2 = bc.vec_uturn(self.ckp[i], self.cks[i], TMP, self.inop[lv])
                            turning |= ok & ((d1 < 0) | (d2 < 0))
                            # (last leaf of 1st half, last leaf) with sum(2nd half) + that momentum
                            d1, d2 = bc.vec_uturn(self.midp[lv], self.mids[lv], SUBSUM, P)
                            turning |= ok & ((d1 < 0) | (d2 < 0))
                    if extra and 2 ** (t_ones + 1) <= 2 ** depth:
                        # this leaf ends the first half of the span of 2^(t_ones+1) leaves
                        bc.vec_axpby(self.midp[t_ones], P, mask=ok)
                        bc.vec_axpby(self.mids[t_ones], SUBSUM, mask=ok)
                in_sub = ok & ~turning
                active &= ~(err | div | turning)
            done = in_sub & active
            with np.errstate(divide="ignore"):
                accept = done & (np.log(rng.random(n)) < sub_logw - logw)
            if accept.any():
                bc.vec_axpby(PROP, self.prp[0], mask=accept)
            logw = np.where(done, np.logaddexp(logw, sub_logw), logw)
            bc.vec_axpby(SUMP, SUBSUM, 1.0, 1.0, mask=done)
            for sgn, eq, ep in ((1, EQ1, EP1), (-1, EQ0, EP0)):
                m = done & (dirs == sgn)
                if m.any():
                    bc.vec_axpby(eq, Q, mask=m)
                    bc.vec_axpby(ep, P, mask=m)
            d1, d2 = bc.vec_uturn(EP0, EP0, SUMP, EP1)      # s = SUMP: (EP0 . SUMP, EP1 . SUMP)
            depth_reached[done] = depth + 1
            active &= done & ~((d1 < 0) | (d2 < 0))
        bc.set_inactive(None, clear_errors=True)
        bc.vec_axpby(Q, PROP)
        bc.set_step_sizes(eps)
        if switch_partition:
            bc.switch_partition()
        else:
            bc.relinearize()
        accept_stat = sum_acc / np.maximum(n_step, 1)
        if self.error_accept_stat == "zero":
            accept_stat = np.where(conv_err | nonrev, 0.0, accept_stat)
        return {"n_step": n_step, "accept_stat": accept_stat, "tree_depth": depth_reached,
                "diverging": diverging, "convergence_error": conv_err, "non_reversible_step": nonrev,
                "hamiltonian_init": h0}
