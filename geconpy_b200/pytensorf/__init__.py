"""Import surface of ``gEconpy.pytensorf`` for the hot path.

``compile``  : ``compile_pytensor_function`` & cache helpers -- thin pass-throughs to pytensor when it is installed
               (graph compilation is host-side plumbing, not part of the CUDA path).
``real_eig`` : the reference's only hot-path use of ``RealEig`` is the Blanchard-Kahn count
               (perturbation.py:499-505), which the B200 path computes without eigenvalues (gecon_bk_count_*).
"""

from . import compile, real_eig  # noqa: F401
