"""Import guard for the optional pytensor layer (pytensor is not installed in the build image)."""

from __future__ import annotations

try:  # pragma: no cover - exercised only where pytensor exists
    import pytensor
    import pytensor.tensor as pt

    from pytensor.graph.basic import Apply
    from pytensor.graph.op import Op

    HAVE_PYTENSOR = True
except Exception:  # pragma: no cover
    pytensor = pt = None
    Apply = None
    Op = object
    HAVE_PYTENSOR = False


def require_pytensor(what: str):
    if not HAVE_PYTENSOR:
        raise ImportError(
            f"{what} builds a pytensor graph and needs pytensor, which is not installed. The numerical entry points "
            "(numpy arrays / torch CUDA tensors, with an optional leading draw axis) do not need it."
        )
