"""``gEconpy.pytensorf.block`` (block.py:53): assemble a matrix from nested lists of blocks, like ``numpy.block``.

The reference needs it to build the Sims pencil ``Gamma0 = [[B, C], [-I, 0]]``, ``Gamma1 = [[A, 0], [0, I]]`` symbolically
(perturbation.py:480-497).  Here the leaves may be numpy arrays, torch tensors (CUDA included, so the pencil of a whole
population of draws is assembled on the device) or -- when pytensor is installed -- symbolic tensors.  Leaves may carry
leading batch axes: the concatenation spans the LAST ``d`` axes (``d`` = nesting depth) and leading axes broadcast.
"""

from __future__ import annotations

import numpy as np

from ..solvers._pt import HAVE_PYTENSOR, pt

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _kind(x) -> str:
    if torch is not None and isinstance(x, torch.Tensor):
        return "torch"
    if HAVE_PYTENSOR and hasattr(x, "owner") and hasattr(x, "type"):
        return "pt"
    return "np"


def _depth(node, where=()) -> int:
    """Nesting depth of the leaves; every leaf must sit at the same depth, containers must be non-empty lists."""
    if isinstance(node, tuple):
        raise TypeError("Block: tuples are not allowed as nested containers; use lists")
    if not isinstance(node, list):
        return 0
    if not node:
        raise ValueError("Block: empty list is not allowed")
    depths = {_depth(child, (*where, i)) for i, child in enumerate(node)}
    if len(depths) != 1:
        raise ValueError(f"Block: all leaves must be at the same nesting depth (mixed depths {sorted(depths)} under index {where})")
    return 1 + depths.pop()


def _leaves(node, out):
    if isinstance(node, list):
        for child in node:
            _leaves(child, out)
    else:
        out.append(node)
    return out


def block(arrays):
    """Nested lists of blocks -> one tensor (see the module docstring).  A bare leaf comes back as ``atleast_1d``."""
    depth = _depth(arrays)
    leaves = _leaves(arrays, [])
    kinds = {_kind(x) for x in leaves}
    lib = "pt" if "pt" in kinds else ("torch" if "torch" in kinds else "np")
    if lib == "torch":
        ref = next(x for x in leaves if _kind(x) == "torch")
        conv = lambda x: x if _kind(x) == "torch" else torch.as_tensor(np.asarray(x), dtype=ref.dtype, device=ref.device)  # noqa: E731
        ndim_of, expand = (lambda x: x.dim()), (lambda x, k: x.reshape((1,) * (k - x.dim()) + tuple(x.shape)))
        cat = lambda xs, axis: torch.cat(_bcast_torch(xs, axis), dim=axis)  # noqa: E731
    elif lib == "pt":
        conv, ndim_of = pt.as_tensor_variable, (lambda x: x.type.ndim)
        expand = lambda x, k: pt.atleast_Nd(x, n=k)  # noqa: E731
        cat = lambda xs, axis: pt.concatenate(xs, axis=axis)  # noqa: E731
    else:
        conv, ndim_of = np.asarray, (lambda x: x.ndim)
        expand = lambda x, k: x.reshape((1,) * (k - x.ndim) + x.shape)  # noqa: E731
        cat = lambda xs, axis: np.concatenate(_bcast_np(xs, axis), axis=axis)  # noqa: E731
    if depth == 0:
        x = conv(arrays)
        return expand(x, max(1, ndim_of(x)))
    nd = max(depth, *(ndim_of(conv(x)) for x in leaves))

    def build(node, level):
        if level == depth:
            return expand(conv(node), nd)
        return cat([build(child, level + 1) for child in node], -(depth - level))

    return build(arrays, 0)


def _bcast_np(xs, axis):
    """Broadcast every axis except ``axis`` (leading batch axes of size 1 stretch to the batch size)."""
    nd = xs[0].ndim
    ax = axis % nd
    shape = [max(x.shape[d] for x in xs) if d != ax else None for d in range(nd)]
    return [np.broadcast_to(x, [x.shape[d] if d == ax else shape[d] for d in range(nd)]) for x in xs]


def _bcast_torch(xs, axis):
    nd = xs[0].dim()
    ax = axis % nd
    shape = [max(x.shape[d] for x in xs) if d != ax else None for d in range(nd)]
    return [x.expand([x.shape[d] if d == ax else shape[d] for d in range(nd)]) for x in xs]


__all__ = ["block"]
