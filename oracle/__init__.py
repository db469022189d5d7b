"""CPU oracle: a restatement of pyfvm's assembly path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product
(``pyfvm_b200``) never does.

PARITY UNPINNED: ``/root/reference`` holds only ``README.md``.  The
reference's Python sources, tests and golden vectors are absent and pyfvm /
meshplex / meshzoo are not installable offline, so this oracle restates the
historically public pyfvm 0.3.x algorithm from the README contract
([README:25-54], [README:86-124]), the finite-volume mathematics and the very
third-party calls the reference used and that ARE installed here:
``sympy.lambdify``, ``numpy.add.at`` / ``subtract.at`` and
``scipy.sparse.coo_matrix(...).tocsr()`` (scipy 1.18.1, ``_coo.py:368-439``,
``_compressed.py:1072-1085``).  It is pinned only by the independent known
answers in SURVEY §4 (``tests/test_oracle.py``, ``tests/golden/``) and by
second-order convergence to manufactured solutions (``tests/test_convergence.py``,
the strategy of the reference's own test suite).

The public surface mirrors the reference: ``discretize_linear``,
``discretize``, ``newton`` and the ``form_language`` module.
"""
from . import form_language
from .discretize import discretize
from .discretize_linear import discretize_linear, split
from .discretize_linear import lower as lower_linear
from .fvm_matrix import EdgeMatrixKernel, get_fvm_matrix
from .linear_fvm_problem import get_linear_fvm_problem
from .newton import newton

__all__ = [
    "form_language",
    "discretize",
    "discretize_linear",
    "lower_linear",
    "get_linear_fvm_problem",
    "get_fvm_matrix",
    "EdgeMatrixKernel",
    "split",
    "newton",
]
