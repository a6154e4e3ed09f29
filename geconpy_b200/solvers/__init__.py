"""Drop-in counterparts of ``gEconpy.solvers`` for the estimation hot path, backed by ``libgecon_b200.so``.

Same function names, argument meaning and failure conventions as the reference modules
(``cycle_reduction``, ``gensys``, ``backward_looking``, ``shared``); every numerical call goes to a CUDA kernel
through the C ABI -- there is no CPU implementation here.  All functions also accept a leading draw axis
(``A[N, n, n]``), in which case one kernel launch serves the whole batch.
"""

from . import backward_looking, cycle_reduction, gensys, shared  # noqa: F401
