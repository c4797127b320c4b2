"""Chain state with per-variable cache invalidation (Mici `ChainState`, `cache_in_state`)."""
import copy


def _cache_key_func(system, method):
    return (type(system).__name__, id(system), method.__name__)


class ChainState:
    def __init__(self, _call_counts=None, _read_only=False, _dependencies=None,
                 _cache=None, **variables):  # fmt: skip
        d = self.__dict__
        d["_variables"] = variables
        d["_dependencies"] = (
            {name: set() for name in variables} if _dependencies is None else _dependencies
        )
        d["_cache"] = {} if _cache is None else _cache
        d["_call_counts"] = _call_counts
        d["_read_only"] = _read_only

    def __getattr__(self, name):
        if name in self.__dict__["_variables"]:
            return self.__dict__["_variables"][name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._variables:
            self._variables[name] = value
            for dep in self._dependencies[name]:
                self._cache[dep] = None
        else:
            self.__dict__[name] = value

    def copy(self, read_only=False):
        return type(self)(
            _call_counts=self._call_counts, _read_only=read_only,
            _dependencies=self._dependencies, _cache=dict(self._cache),
            **{k: copy.copy(v) for k, v in self._variables.items()})  # fmt: skip


def _bump(state, key):
    if state._call_counts is not None:
        state._call_counts[key] = state._call_counts.get(key, 0) + 1


def cache_in_state(*depends_on):
    def decorator(method):
        def wrapper(self, state):
            key = _cache_key_func(self, method)
            if key not in state._cache:
                for var in depends_on:
                    state._dependencies[var].add(key)
            if state._cache.get(key) is None:
                state._cache[key] = method(self, state)
                _bump(state, key)
            return state._cache[key]

        wrapper.__name__ = method.__name__
        return wrapper

    return decorator


def cache_in_state_with_aux(depends_on, auxiliary_outputs):
    if isinstance(depends_on, str):
        depends_on = (depends_on,)
    if isinstance(auxiliary_outputs, str):
        auxiliary_outputs = (auxiliary_outputs,)

    def decorator(method):
        def wrapper(self, state):
            prim = _cache_key_func(self, method)
            keys = [prim] + [_cache_key_func(self, getattr(self, a)) for a in auxiliary_outputs]
            for key in keys:
                if key not in state._cache:
                    for var in depends_on:
                        state._dependencies[var].add(key)
            if state._cache.get(prim) is None:
                vals = method(self, state)
                _bump(state, prim)
                if isinstance(vals, tuple):
                    for key, val in zip(keys, vals):
                        state._cache[key] = val
                else:
                    state._cache[prim] = vals
            return state._cache[prim]

        wrapper.__name__ = method.__name__
        return wrapper

    return decorator
