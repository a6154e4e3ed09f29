"""Import shim that makes the *pure-sympy* part of the reference importable in this container.

TEST/ORACLE INFRASTRUCTURE ONLY.  Used by ``make_models.py`` / ``make_goldens.py`` to generate the
committed fixtures under ``tests/golden/``; nothing in the product package imports it, and nothing
that runs on the GPU box does (``/root/reference`` does not exist there).

The reference (``/root/reference/gEconpy``) imports pytensor, pymc, pymc_extras, sympytensor, preliz,
pyparsing, xarray, ... at module import time; none of those is installed here (SURVEY.md section 0,
fact 10 and Appendix D).  The GCN parser, FOC derivation, simplification and steady-state
propagation only need sympy + pyparsing, so this shim

* aliases ``pyparsing`` to pip's vendored copy, and
* answers every other missing third-party import with a permissive placeholder whose attributes
  are placeholders, which can be called, subscripted, used in ``X | None`` annotations, used as a
  base class, and which -- when used as a decorator -- returns the decorated function unchanged.

Anything that actually needs pytensor arithmetic will fail loudly when called; the pure numpy/sympy
functions (``_compile_gcn``, ``cycle_reduction_numpy``, ``_gensys_setup``, ...) run as written.
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"

_MISSING = (
    "pytensor",
    "pymc",
    "pymc_extras",
    "sympytensor",
    "preliz",
    "xarray",
    "matplotlib",
    "arviz",
    "better_optimize",
    "IPython",
    "statsmodels",
    "jax",
    "nutpie",
    "rich",
)


class _Placeholder:
    def __init__(self, name="placeholder"):
        object.__setattr__(self, "_name", name)

    def __getattr__(self, key):
        if key.startswith("__") and key.endswith("__"):
            raise AttributeError(key)
        return _Placeholder(f"{self._name}.{key}")

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and not kwargs and isinstance(args[0], types.FunctionType):
            return args[0]  # decorator use: leave the python function as it is
        return _Placeholder(f"{self._name}()")

    def __mro_entries__(self, bases):
        return (object,)

    def __getitem__(self, key):
        return _Placeholder(f"{self._name}[]")

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __iter__(self):
        return iter(())

    def __repr__(self):
        return f"<placeholder {self._name}>"


class _PlaceholderModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, key):
        if key.startswith("__") and key.endswith("__"):
            raise AttributeError(key)
        return _Placeholder(f"{self.__name__}.{key}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _PlaceholderModule(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make ``import gEconpy....`` resolve to the read-only reference tree (idempotent)."""
    global _installed
    if _installed:
        return
    missing = []
    for name in _MISSING:
        try:
            importlib.import_module(name)
        except Exception:
            missing.append(name)
    globals()["_MISSING"] = tuple(missing)
    sys.meta_path.append(_Finder())
    try:
        import pyparsing  # noqa: F401
    except ImportError:
        from pip._vendor import pyparsing as _pp

        sys.modules["pyparsing"] = _pp
        for sub in ("common", "exceptions", "helpers", "results", "core", "util", "actions", "unicode", "testing"):
            try:
                sys.modules[f"pyparsing.{sub}"] = importlib.import_module(f"pip._vendor.pyparsing.{sub}")
            except Exception:
                pass
    pkg = types.ModuleType("gEconpy")
    pkg.__path__ = [f"{REFERENCE_ROOT}/gEconpy"]
    sys.modules["gEconpy"] = pkg
    _installed = True
