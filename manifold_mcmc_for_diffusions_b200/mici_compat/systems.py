"""mici.systems.System base class (the reference subclass overrides everything else, mici_extensions.py:1186-1259)."""
from .states import cache_in_state, cache_in_state_with_aux  # noqa: F401  (re-exported like mici.systems)


class System:
    def __init__(self, neg_log_dens, grad_neg_log_dens=None):
        self._neg_log_dens = neg_log_dens
        self._grad_neg_log_dens = grad_neg_log_dens

    @cache_in_state("pos")
    def neg_log_dens(self, state):
        return self._neg_log_dens(state.pos)

    @cache_in_state_with_aux("pos", "neg_log_dens")
    def grad_neg_log_dens(self, state):
        out = self._grad_neg_log_dens(state.pos)
        return out if isinstance(out, tuple) else (out, self._neg_log_dens(state.pos))

    def h(self, state):
        return self.h1(state) + self.h2(state)

    def h1_flow(self, state, dt):
        state.mom -= dt * self.dh1_dpos(state)

    def dh_dmom(self, state):
        return self.dh2_dmom(state)
