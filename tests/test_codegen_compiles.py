"""CPU tier: the generated functor preludes + the tile-kernel template compile for
sm_100a (nvcc cross-compiles without a GPU).  Catches codegen gotchas (SURVEY
App. A.6: M_PI, pow, Piecewise, cse temporaries) before a GPU is involved."""
import os
import shutil
import subprocess

import pytest
import sympy

from . import problems
from .test_convergence import manufactured

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def piecewise_problem(fl):
    class P:
        def apply(self, u):
            # (the reference's edge visitor knows no Piecewise: it stays in the dV term)
            k = lambda x: sympy.Piecewise((1.0, x[0] < 0.5), (10.0, True))
            return (
                fl.integrate(lambda x: -(1 + x[0] ** 2) * fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(
                    lambda x: k(x) * sympy.sqrt(1 + x[1] ** 2) * u(x) ** 3 - sympy.pi * x[0] ** 4, fl.dV
                )
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - sympy.exp(-x[0]), fl.Boundary())]

    return P()


CASES = [
    ("poisson", 2, "linear"), ("reaction_x", 3, "linear"), ("subdomain_problem", 2, "linear"),
    ("bratu", 3, "nonlinear"), ("nonlinear_conv", 2, "nonlinear"), ("manufactured", 2, "linear"),
    ("piecewise", 2, "nonlinear"),
]


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
@pytest.mark.parametrize("name,dim,kind", CASES)
def test_generated_kernel_compiles(name, dim, kind, tmp_path):
    import ctypes

    from pyfvm_b200 import codegen, engine, form_language, lowering

    if name == "manufactured":
        prob = manufactured(form_language, dim, "convection")
    elif name == "piecewise":
        prob = piecewise_problem(form_language)
    else:
        prob = getattr(problems, name)(form_language)
    low = (lowering.lower_linear if kind == "linear" else lowering.lower_nonlinear)(prob, dim)
    ebits = [None if t.subdomain is None else i for i, t in enumerate(low.edge)]
    vbits = [None if t.subdomain is None else i for i, t in enumerate(low.vertex)]
    modes = [codegen.MODE_LINEAR] if kind == "linear" else [codegen.MODE_RESIDUAL, codegen.MODE_JACOBIAN]
    lib = engine.load_library()
    for mode in modes:
        fun = codegen.generate(low, mode, ebits, vbits)
        need = ctypes.c_size_t()
        lib.fvm_kernel_source(fun.source.encode(), None, 0, ctypes.byref(need))
        buf = ctypes.create_string_buffer(need.value)
        lib.fvm_kernel_source(fun.source.encode(), buf, need.value, ctypes.byref(need))
        src = tmp_path / f"{name}_{mode}.cu"
        src.write_bytes(buf.value)
        r = subprocess.run(
            [NVCC, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false",
             "-cubin", "-o", str(tmp_path / "k.cubin"), str(src)],
            capture_output=True, text=True,
        )
        assert r.returncode == 0, r.stderr[-2000:]
