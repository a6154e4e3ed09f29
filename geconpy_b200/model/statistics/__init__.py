"""Batch diagnostics over parameter draws (``gEconpy.model.statistics`` counterparts on the hot-path kernels)."""

from .perturbation_diagnostics import solvability_check  # noqa: F401
