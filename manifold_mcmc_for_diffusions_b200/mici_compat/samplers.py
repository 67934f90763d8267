"""mici.samplers.MarkovChainMonteCarloMethod, reduced to what scripts/utils.py:338-365 uses: sequential
chains, one adaptive warm-up stage with fast adapters followed by the main stage, traces per chain."""
import os

import numpy as np


class MarkovChainMonteCarloMethod:
    def __init__(self, rng, transitions):
        self.rng = rng
        self.transitions = transitions

    def _run_stage(self, state, rng, n_iter, trace_funcs, adapters, adapter_states):
        traces, stats = {}, {}
        for it in range(n_iter):
            for key, transition in self.transitions.items():
                state, st = transition.sample(state, rng)
                if st is not None:
                    for k, v in st.items():
                        stats.setdefault(key, {}).setdefault(k, []).append(v)
                    if adapters and key in adapters:
                        for ad, ast in zip(adapters[key], adapter_states[key]):
                            ad.update(ast, state, st, transition)
            if trace_funcs:
                for f in trace_funcs:
                    for k, v in f(state).items():
                        traces.setdefault(k, []).append(np.asarray(v))
        return state, traces, stats

    def sample_chains_with_adaptive_warm_up(self, n_warm_up_iter, n_main_iter, init_states, trace_funcs=None,
                                            adapters=None, trace_warm_up=False, memmap_enabled=False,
                                            memmap_path=None, **kwargs):
        seeds = self.rng.bit_generator.seed_seq.spawn(len(init_states)) if hasattr(
            self.rng.bit_generator, "seed_seq") else [None] * len(init_states)
        rngs = [np.random.default_rng(s) for s in seeds]
        adapters = adapters or {}
        final_states, all_traces, all_stats = [], {}, {}
        adapter_states_all = {key: [] for key in adapters}
        states = list(init_states)
        # warm-up stage (adaptive), chain by chain
        for c, state in enumerate(states):
            ast = {key: [ad.initialize(state, self.transitions[key]) for ad in ads] for key, ads in adapters.items()}
            state, _, _ = self._run_stage(state, rngs[c], n_warm_up_iter, None, adapters, ast)
            for key in adapters:
                adapter_states_all[key].append(ast[key])
            states[c] = state
        for key, ads in adapters.items():
            for i, ad in enumerate(ads):
                ad.finalize([chain_ast[i] for chain_ast in adapter_states_all[key]], self.transitions[key])
        # main stage
        for c, state in enumerate(states):
            state, traces, stats = self._run_stage(state, rngs[c], n_main_iter, trace_funcs, None, None)
            final_states.append(state)
            for k, v in traces.items():
                arr = np.stack(v) if v else np.empty((0,))
                if memmap_enabled and memmap_path is not None:
                    os.makedirs(memmap_path, exist_ok=True)
                    np.save(os.path.join(memmap_path, f"trace_{c}_{k}.npy"), arr)
                all_traces.setdefault(k, []).append(arr)
            for key, d in stats.items():
                for k, v in d.items():
                    all_stats.setdefault(key, {}).setdefault(k, []).append(np.asarray(v))
        return final_states, all_traces, all_stats
