"""Generates the golden fixtures in this directory from the CPU oracle.

The reference checkout holds no golden vectors and pyfvm cannot be
imported here, so these are ORACLE outputs, not reference outputs; they pin the
oracle against accidental change and give the GPU tests a fixture that does not
need the oracle at run time.  Run:  python -m tests.golden.make_golden
"""
import os

import numpy
from scipy.sparse import linalg

import oracle
from pyfvm_b200 import meshes
from tests import problems

fl = oracle.form_language
HERE = os.path.dirname(os.path.abspath(__file__))


def _linear(problem, mesh):
    A, b = oracle.discretize_linear(problem, mesh)
    return {"indptr": A.indptr, "indices": A.indices, "data": A.data, "rhs": b}


def poisson_33x17():
    v, c = problems.rectangle(33, 17)
    return _linear(problems.poisson(fl), meshes.Mesh(v, c))


def convdiff_33x17():
    v, c = problems.rectangle(33, 17)
    return _linear(problems.convection_diffusion(fl), meshes.Mesh(v, c))


def poisson_cube5():
    p, c = problems.cube(5)
    return _linear(problems.poisson(fl), meshes.Mesh(p, c))


def bratu_21x11():
    v, c = problems.rectangle(21, 11)
    mesh = meshes.Mesh(v, c)
    f, jac = oracle.discretize(problems.bratu(fl), mesh)
    u = numpy.sin(numpy.arange(len(v)) * 0.37) * 0.5
    J = jac.get_linear_operator(u)
    return {"u": u, "F": f.eval(u), "indptr": J.indptr, "indices": J.indices, "data": J.data}


CASES = {
    "poisson_33x17": poisson_33x17,
    "convdiff_33x17": convdiff_33x17,
    "bratu_21x11": bratu_21x11,
    "poisson_cube5": poisson_cube5,
}

if __name__ == "__main__":
    for name, fn in CASES.items():
        numpy.savez_compressed(os.path.join(HERE, name + ".npz"), **fn())
        print("wrote", name)
