"""``pyfvm.newton(f, jacobian_solver, u0)`` [README:113-121], unchanged: an
undamped Newton iteration on host arrays (defaults per SURVEY §2: tol=1e-10,
max_iter=20, residual norms printed when verbose, convergence asserted)."""
import numpy


def newton(f, jacobian_solver, u0, tol=1.0e-10, max_iter=20, verbose=True):
    u = u0.copy()
    fu = f(u)
    nrm = numpy.sqrt(numpy.dot(fu, fu))
    if verbose:
        print(f"||F(u)|| = {nrm:e}")
    k = 0
    is_converged = False
    while k < max_iter:
        if nrm < tol:
            is_converged = True
            break
        du = jacobian_solver(u, -fu)
        u += du
        fu = f(u)
        nrm = numpy.sqrt(numpy.dot(fu, fu))
        k += 1
        if verbose:
            print(f"||F(u)|| = {nrm:e}")
    assert is_converged, "Newton's method did not converge"
    return u
