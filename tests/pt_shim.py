"""A minimal stand-in for the handful of pytensor names the Op classes of ``geconpy_b200`` touch (``Op``, ``Apply``,
``pt.as_tensor``, ``pt.tensor``, ``pt.scalar``, ``pt.vector``): TEST INFRASTRUCTURE.  pytensor is not installable in the build
container, so without this the Op layer (``make_node`` / ``infer_shape`` / ``perform`` / ``pullback``) would never execute.
``install()`` patches the import guard ``geconpy_b200.solvers._pt`` and every module that imported names from it;
``run(op, *arrays)`` does what a compiled pytensor function does with a single Apply node: build it, allocate the output
storage cells, call ``perform`` and hand back the computed values."""

from __future__ import annotations

import importlib
import types

import numpy as np


class TensorType:
    def __init__(self, dtype, shape):
        self.dtype, self.shape, self.ndim = str(dtype), tuple(shape), len(tuple(shape))
        self.numpy_dtype = np.dtype(self.dtype)


class Variable:
    def __init__(self, type_, name=None, owner=None, value=None):
        self.type, self.name, self.owner, self.value = type_, name, owner, value
        self.ndim = type_.ndim


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for o in self.outputs:
            o.owner = self


class Op:
    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs


def _as_tensor(x, name=None):
    if isinstance(x, Variable):
        return x
    a = np.asarray(x)
    return Variable(TensorType(a.dtype, a.shape), name=name, value=a)


pt = types.SimpleNamespace(
    as_tensor=_as_tensor,
    as_tensor_variable=_as_tensor,
    tensor=lambda name=None, dtype="float64", shape=(): Variable(TensorType(dtype, shape), name=name),
    scalar=lambda name=None, dtype="float64": Variable(TensorType(dtype, ()), name=name),
    vector=lambda name=None, dtype="float64", shape=(None,): Variable(TensorType(dtype, shape), name=name),
    zeros_like=lambda v: Variable(TensorType(v.type.dtype, v.type.shape), value=np.zeros(v.type.shape)),
)

_MODULES = ["geconpy_b200.solvers.cycle_reduction", "geconpy_b200.solvers.gensys", "geconpy_b200.pytensorf.real_eig", "geconpy_b200.solvers.shared"]


def install(monkeypatch):
    """Patch the guard module and re-import its clients so that their Op classes derive from the shim's ``Op``."""
    guard = importlib.import_module("geconpy_b200.solvers._pt")
    for name, val in (("HAVE_PYTENSOR", True), ("pt", pt), ("Op", Op), ("Apply", Apply)):
        monkeypatch.setattr(guard, name, val)
    mods = {}
    for m in _MODULES:
        mod = importlib.import_module(m)
        mods[m.rsplit(".", 1)[1]] = importlib.reload(mod)
    return mods


def uninstall():
    """Undo ``install`` (after monkeypatch restored the guard): reload the clients against the real guard."""
    for m in _MODULES:
        importlib.reload(importlib.import_module(m))


def run(op, *arrays):
    node = op.make_node(*arrays)
    storage = [[None] for _ in node.outputs]
    op.perform(node, [np.asarray(a) for a in arrays], storage)
    return node, [cell[0] for cell in storage]
