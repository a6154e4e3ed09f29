"""``gEconpy.pytensorf.compile`` (compile.py:11-136): graph compilation helpers, pytensor-guarded."""

from __future__ import annotations

import functools

from ..solvers._pt import HAVE_PYTENSOR, pytensor, require_pytensor


def rewrite_pregrad(graph):
    require_pytensor("rewrite_pregrad")
    from pytensor.graph.rewriting.utils import rewrite_graph

    return rewrite_graph(graph, include=("canonicalize", "stabilize"))


@functools.lru_cache(maxsize=128)
def _compile_cached(inputs, outputs, mode, kwargs_items):
    return pytensor.function(list(inputs), list(outputs), mode=mode, **dict(kwargs_items))


def compile_pytensor_function(inputs, outputs, mode=None, **kwargs):
    require_pytensor("compile_pytensor_function")
    return _compile_cached(tuple(inputs), tuple(outputs) if isinstance(outputs, (list, tuple)) else (outputs,), mode, tuple(sorted(kwargs.items())))


def clear_compile_cache() -> None:
    _compile_cached.cache_clear()


def compile_cache_info():
    return _compile_cached.cache_info()


__all__ = ["HAVE_PYTENSOR", "rewrite_pregrad", "compile_pytensor_function", "clear_compile_cache", "compile_cache_info"]
