"""mici.transitions: momentum refresh and the multinomial dynamic integration (NUTS) transition
(Mici 0.1.10 as recalled in SURVEY.md appendix A; call sites scripts/utils.py:292-301)."""
import numpy as np

from .errors import ConvergenceError, IntegratorError, NonReversibleStepError


class Transition:
    state_variables = set()
    statistic_types = None

    def sample(self, state, rng):
        raise NotImplementedError


class IndependentMomentumTransition(Transition):
    state_variables = {"mom"}
    statistic_types = None

    def __init__(self, system):
        self.system = system

    def sample(self, state, rng):
        state.mom = self.system.sample_momentum(state, rng)
        return state, None


def riemannian_no_u_turn_criterion(system, state_1, state_2, sum_mom):
    return (np.sum(system.dh_dmom(state_1) * sum_mom) < 0) or (np.sum(system.dh_dmom(state_2) * sum_mom) < 0)


def _logaddexp(a, b):
    return np.logaddexp(a, b)


class MultinomialDynamicIntegrationTransition(Transition):
    """Dynamic (doubling-tree) integration transition with multinomial sampling from the trajectory:
    within-subtree proposals by multinomial sampling, top level by biased progressive sampling."""

    state_variables = {"pos", "mom"}
    statistic_types = {
        "hamiltonian": (np.float64, np.nan), "n_step": (np.int64, -1), "accept_stat": (np.float64, np.nan),
        "tree_depth": (np.int64, -1), "diverging": (bool, False), "non_reversible_step": (bool, False),
        "convergence_error": (bool, False),
    }

    def __init__(self, system, integrator, max_tree_depth=10, max_delta_h=1000,
                 termination_criterion=riemannian_no_u_turn_criterion, do_extra_subtree_checks=True,
                 error_accept_stat="partial"):
        # error_accept_stat: see nuts.BatchedNUTS (accept_stat of a transition that ended in an integrator error);
        # "partial" = sum_acc_prob / n_step, Mici's rule
        self.error_accept_stat = error_accept_stat
        self.system = system
        self.integrator = integrator
        self.max_tree_depth = max_tree_depth
        self.max_delta_h = max_delta_h
        self._termination_criterion = termination_criterion
        self.do_extra_subtree_checks = do_extra_subtree_checks

    def _leaf(self, state, h_init, stats):
        state = self.integrator.step(state)
        h = self.system.h(state)
        h = np.inf if np.isnan(h) else h
        stats["sum_acc_prob"] += min(1.0, np.exp(h_init - h))
        stats["n_step"] += 1
        if h - h_init > self.max_delta_h:
            stats["diverging"] = True
            return True, state, -h
        return False, state, -h

    def _build_tree(self, depth, state, h_init, stats, rng):
        """Returns (terminate, inner_edge, outer_edge, proposal, sum_mom, log_weight)."""
        if depth == 0:
            terminate, state, lw = self._leaf(state, h_init, stats)
            return terminate, state, state, state, np.array(state.mom, copy=True), lw
        t, in_i, out_i, prop_i, mom_i, lw_i = self._build_tree(depth - 1, state, h_init, stats, rng)
        if t:
            return True, None, None, None, None, None
        t, in_o, out_o, prop_o, mom_o, lw_o = self._build_tree(depth - 1, out_i, h_init, stats, rng)
        if t:
            return True, None, None, None, None, None
        lw = _logaddexp(lw_i, lw_o)
        proposal = prop_o if np.log(rng.uniform()) < lw_o - lw else prop_i
        sum_mom = mom_i + mom_o
        terminate = self._termination_criterion(self.system, in_i, out_o, sum_mom)
        if self.do_extra_subtree_checks and not terminate:
            terminate = (self._termination_criterion(self.system, in_i, in_o, mom_i + in_o.mom)
                         or self._termination_criterion(self.system, out_i, out_o, mom_o + out_i.mom))
        return terminate, in_i, out_o, proposal, sum_mom, lw

    def sample(self, state, rng):
        h_init = self.system.h(state)
        stats = {"n_step": 0, "sum_acc_prob": 0.0, "diverging": False, "non_reversible_step": False,
                 "convergence_error": False}
        sum_mom = np.array(state.mom, copy=True)
        log_weight = -h_init
        state_n, state_p, proposal = state, state, state
        depth = 0
        try:
            for depth in range(self.max_tree_depth):
                direction = 1 if rng.uniform() < 0.5 else -1
                edge = (state_p if direction == 1 else state_n).copy()
                edge.dir = direction
                t, _, out, new_prop, new_mom, new_lw = self._build_tree(depth, edge, h_init, stats, rng)
                if t:
                    break
                if direction == 1:
                    state_p = out
                else:
                    state_n = out
                if np.log(rng.uniform()) < new_lw - log_weight:
                    proposal = new_prop
                sum_mom = sum_mom + new_mom
                log_weight = _logaddexp(log_weight, new_lw)
                if self._termination_criterion(self.system, state_n, state_p, sum_mom):
                    depth += 1
                    break
            else:
                depth = self.max_tree_depth
        except ConvergenceError:
            stats["convergence_error"] = True
        except NonReversibleStepError:
            stats["non_reversible_step"] = True
        except IntegratorError:
            stats["convergence_error"] = True
        new_state = proposal.copy()
        new_state.dir = state.dir
        n_step = stats["n_step"]
        accept_stat = stats["sum_acc_prob"] / n_step if n_step > 0 else 0.0
        if self.error_accept_stat == "zero" and (stats["convergence_error"] or stats["non_reversible_step"]):
            accept_stat = 0.0
        out_stats = {
            "hamiltonian": self.system.h(new_state), "n_step": n_step,
            "accept_stat": accept_stat, "tree_depth": depth,
            "diverging": stats["diverging"], "non_reversible_step": stats["non_reversible_step"],
            "convergence_error": stats["convergence_error"],
        }
        return new_state, out_stats


class MetropolisStaticIntegrationTransition(Transition):
    """Static-trajectory HMC transition with a Metropolis accept step (integrator errors reject)."""

    state_variables = {"pos", "mom"}
    statistic_types = {"hamiltonian": (np.float64, np.nan), "n_step": (np.int64, -1),
                       "accept_stat": (np.float64, np.nan), "non_reversible_step": (bool, False),
                       "convergence_error": (bool, False)}

    def __init__(self, system, integrator, n_step):
        self.system, self.integrator, self.n_step = system, integrator, n_step

    def sample(self, state, rng):
        h_init = self.system.h(state)
        stats = {"n_step": self.n_step, "non_reversible_step": False, "convergence_error": False}
        new = state
        try:
            for _ in range(self.n_step):
                new = self.integrator.step(new)
            h = self.system.h(new)
            acc = min(1.0, np.exp(h_init - h)) if np.isfinite(h) else 0.0
        except NonReversibleStepError:
            stats["non_reversible_step"] = True
            acc = 0.0
        except IntegratorError:
            stats["convergence_error"] = True
            acc = 0.0
        if rng.uniform() < acc:
            state = new
        stats.update(hamiltonian=self.system.h(state), accept_stat=acc)
        return state, stats
