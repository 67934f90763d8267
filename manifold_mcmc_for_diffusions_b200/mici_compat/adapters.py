"""mici.adapters.DualAveragingStepSizeAdapter (Mici 0.1.10 as recalled in SURVEY.md appendix A;
call site scripts/utils.py:303-306 with target 0.8 and regularisation coefficient 0.1)."""
import numpy as np

from .errors import AdaptationError, IntegratorError


class Adapter:
    is_fast = True


class DualAveragingStepSizeAdapter(Adapter):
    is_fast = True

    def __init__(self, adapt_stat_target=0.8, adapt_stat_func=None, log_step_size_reg_target=None,
                 log_step_size_reg_coefficient=0.05, iter_decay_coeff=0.75, iter_offset=10,
                 max_init_step_size_iters=100):
        self.adapt_stat_target = adapt_stat_target
        self.adapt_stat_func = adapt_stat_func or (lambda stats: stats["accept_stat"])
        self.log_step_size_reg_target = log_step_size_reg_target
        self.log_step_size_reg_coefficient = log_step_size_reg_coefficient
        self.iter_decay_coeff = iter_decay_coeff
        self.iter_offset = iter_offset
        self.max_init_step_size_iters = max_init_step_size_iters

    def initialize(self, chain_state, transition):
        integrator = transition.integrator
        system = transition.system
        adapter_state = {"iter": 0, "smoothed_log_step_size": 0.0, "adapt_stat_error": 0.0}
        init_step_size = (self._find_and_set_init_step_size(chain_state, system, integrator)
                          if integrator.step_size is None else integrator.step_size)
        adapter_state["log_step_size_reg_target"] = (
            np.log(10 * init_step_size) if self.log_step_size_reg_target is None else self.log_step_size_reg_target)
        return adapter_state

    def _find_and_set_init_step_size(self, state, system, integrator):
        init_state = state.copy()
        h_init = system.h(init_state)
        if np.isnan(h_init):
            raise AdaptationError("Hamiltonian evaluating to NaN at initial state.")
        integrator.step_size = 1.0
        delta_h_threshold = np.log(2)
        for s in range(self.max_init_step_size_iters):
            try:
                state = integrator.step(init_state)
                delta_h = abs(h_init - system.h(state))
                if s == 0 or np.isnan(delta_h):
                    step_size_too_big = np.isnan(delta_h) or delta_h > delta_h_threshold
                if (step_size_too_big and delta_h <= delta_h_threshold) or (
                        not step_size_too_big and delta_h > delta_h_threshold):
                    return integrator.step_size
                elif step_size_too_big:
                    integrator.step_size /= 2.0
                else:
                    integrator.step_size *= 2.0
            except IntegratorError:
                step_size_too_big = True
                integrator.step_size /= 2.0
        raise AdaptationError("Could not find reasonable initial step size.")

    def update(self, adapter_state, chain_state, trans_stats, transition):
        adapter_state["iter"] += 1
        n = adapter_state["iter"]
        error_weight = 1.0 / (self.iter_offset + n)
        adapter_state["adapt_stat_error"] *= 1 - error_weight
        adapter_state["adapt_stat_error"] += error_weight * (self.adapt_stat_target - self.adapt_stat_func(trans_stats))
        smoothing_weight = (1.0 / n) ** self.iter_decay_coeff
        log_step_size = adapter_state["log_step_size_reg_target"] - (
            adapter_state["adapt_stat_error"] * n ** 0.5 / self.log_step_size_reg_coefficient)
        adapter_state["smoothed_log_step_size"] *= 1 - smoothing_weight
        adapter_state["smoothed_log_step_size"] += smoothing_weight * log_step_size
        transition.integrator.step_size = float(np.exp(log_step_size))

    def finalize(self, adapter_states, transition):
        if isinstance(adapter_states, dict):
            adapter_states = [adapter_states]
        transition.integrator.step_size = float(
            np.mean([np.exp(a["smoothed_log_step_size"]) for a in adapter_states]))
