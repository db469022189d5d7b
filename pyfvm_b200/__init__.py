"""pyfvm_b200: B200-native finite-volume assembly behind pyfvm's API.

Drop-in for pyfvm's hot path (README.md of the reference):
``discretize_linear(problem, mesh) -> (matrix, rhs)``, ``discretize(problem,
mesh) -> (f, jacobian)``, ``f.eval(u)``, ``jacobian.get_linear_operator(u0)``,
``newton`` and the ``form_language`` module.  The numeric path runs on the GPU
through ``libfvm_b200.so`` (C-ABI, ``include/fvm_b200.h``); there is no CPU
fallback.
"""
from . import form_language, fvm_matrix
from .fvm_matrix import EdgeMatrixKernel, get_fvm_matrix
from .newton import newton
from .problem import FvmProblem, Jacobian, LinearProblem, discretize, discretize_linear

__version__ = "0.1.0"
__all__ = [
    "form_language",
    "fvm_matrix",
    "EdgeMatrixKernel",
    "get_fvm_matrix",
    "discretize",
    "discretize_linear",
    "newton",
    "FvmProblem",
    "Jacobian",
    "LinearProblem",
]
