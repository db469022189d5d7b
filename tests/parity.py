"""Parity checks between the engine and the CPU oracle:
pattern arrays byte-identical; values within 1e-13 of the sum of the absolute
contributions of the slot / row."""
import numpy

RTOL = 1.0e-13  # north_star: "values within 1e-13 relative in fp64"


def assert_pattern_equal(A, B):
    assert A.shape == B.shape
    assert A.indptr.dtype == numpy.int32 and A.indices.dtype == numpy.int32
    assert B.indptr.dtype == numpy.int32 and B.indices.dtype == numpy.int32
    assert A.indptr.tobytes() == B.indptr.tobytes(), "indptr differs"
    assert A.indices.tobytes() == B.indices.tobytes(), "indices differ"


def assert_matrix_close(A, B, scale):
    """A: engine, B: oracle, scale: oracle's per-slot sum of |contribution|."""
    assert_pattern_equal(A, B)
    assert scale.indptr.tobytes() == B.indptr.tobytes()
    err = numpy.abs(A.data - B.data)
    # Dirichlet-replaced rows hold exact values (0 / coeff) on both sides; their
    # pre-replacement scale only makes the bound looser, never tighter than 0.
    bound = RTOL * scale.data
    bad = err > bound
    assert not bad.any(), (
        f"{bad.sum()} slots out of tolerance; worst err {err[bad].max():.3e} "
        f"vs bound {bound[bad].min():.3e}"
    )


def assert_vector_close(a, b, scale):
    err = numpy.abs(a - b)
    bound = RTOL * scale
    # rows overwritten by a Dirichlet condition: compare against |value| too
    bound = numpy.maximum(bound, RTOL * numpy.abs(b))
    bad = err > bound
    assert not bad.any(), (
        f"{bad.sum()} rows out of tolerance; worst err {err[bad].max():.3e} "
        f"vs bound {bound[bad].min():.3e}"
    )
