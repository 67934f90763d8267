"""mici.states: chain state with a dependency-tracked cache."""
import copy
from functools import wraps


def _cache_key_func(system, method):
    name = method if isinstance(method, str) else method.__name__
    return (type(system).__name__, id(system), name)


class ChainState:
    def __init__(self, _call_counts=None, _read_only=False, _dependencies=None, _cache=None, **variables):
        object.__setattr__(self, "_variables", dict(variables))
        object.__setattr__(self, "_call_counts", _call_counts)
        object.__setattr__(self, "_read_only", _read_only)
        object.__setattr__(self, "_dependencies", {n: set() for n in variables} if _dependencies is None else _dependencies)
        object.__setattr__(self, "_cache", {} if _cache is None else _cache)

    def __getattr__(self, name):
        v = object.__getattribute__(self, "_variables")
        if name in v:
            return v[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._variables:
            if self._read_only:
                raise RuntimeError("read-only state")
            self._variables[name] = value
            for key in self._dependencies.get(name, ()):
                self._cache[key] = None
        else:
            object.__setattr__(self, name, value)

    def copy(self, read_only=False):
        return type(self)(
            **{n: copy.copy(v) for n, v in self._variables.items()},
            _call_counts=self._call_counts, _dependencies=self._dependencies, _cache=dict(self._cache),
            _read_only=read_only,
        )


def cache_in_state(*depends_on):
    def decorator(method):
        @wraps(method)
        def wrapper(self, state):
            key = _cache_key_func(self, method)
            if state._cache.get(key) is None:
                for d in depends_on:
                    state._dependencies.setdefault(d, set()).add(key)
                state._cache[key] = method(self, state)
                if state._call_counts is not None:
                    state._call_counts[key] = state._call_counts.get(key, 0) + 1
            return state._cache[key]

        return wrapper

    return decorator


def cache_in_state_with_aux(depends_on, auxiliary_outputs):
    if isinstance(depends_on, str):
        depends_on = (depends_on,)
    if isinstance(auxiliary_outputs, str):
        auxiliary_outputs = (auxiliary_outputs,)

    def decorator(method):
        @wraps(method)
        def wrapper(self, state):
            key = _cache_key_func(self, method)
            if state._cache.get(key) is None:
                aux_keys = [_cache_key_func(self, getattr(self, a)) for a in auxiliary_outputs]
                for d in depends_on:
                    state._dependencies.setdefault(d, set()).update([key] + aux_keys)
                outs = method(self, state)
                state._cache[key] = outs[0]
                for ak, val in zip(aux_keys, outs[1:]):
                    state._cache[ak] = val
                if state._call_counts is not None:
                    state._call_counts[key] = state._call_counts.get(key, 0) + 1
            return state._cache[key]

        return wrapper

    return decorator
