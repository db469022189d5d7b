"""Oracle restatement of pyfvm's edge-integral visitor.
Test infrastructure only; see ``oracle/__init__.py``.

Walks the sympy tree of a ``dS`` integrand on the edge x0 -> x1:
``n_dot_grad(f(x)) -> (f(x1) - f(x0))/edge_length``; bare ``f(x) -> (f0+f1)/2``;
``n -> (x1-x0)/edge_length``; ``x -> (x0+x1)/2``; the result is multiplied by
``edge_ce_ratio*edge_length`` (the dual-face measure)."""
import sympy

from . import form_language as fl


class DiscretizeEdgeIntegral:
    def __init__(self, x0, x1, edge_length, edge_ce_ratio):
        self.x0 = x0
        self.x1 = x1
        self.edge_length = edge_length
        self.edge_ce_ratio = edge_ce_ratio
        self.x = sympy.DeferredVector("x")

    def generate(self, integrand, index_functions):
        expr = integrand(self.x)
        out = self.edge_ce_ratio * self.edge_length * self.visit(expr)
        index_vars = []
        for f in index_functions:
            fk0 = sympy.Symbol(f"{f}k0")
            fk1 = sympy.Symbol(f"{f}k1")
            out = out.subs(f(self.x0), fk0)
            out = out.subs(f(self.x1), fk1)
            # f(x) not under a gradient: value at the edge midpoint
            out = out.subs(f(self.x), 0.5 * (fk0 + fk1))
            index_vars.append([fk0, fk1])
        # explicit x: the edge midpoint; n: the unit edge direction
        rep = {}
        for k in range(3):
            rep[self.x[k]] = 0.5 * (self.x0[k] + self.x1[k])
            rep[fl.n[k]] = (self.x1[k] - self.x0[k]) / self.edge_length
        out = out.subs(rep, simultaneous=True)
        return out, index_vars

    def visit(self, node):
        if isinstance(node, (int, float)):
            return node
        if isinstance(node, sympy.Basic):
            if node.is_Add:
                return self.visit_chain(node, sympy.Add)
            if node.is_Mul:
                return self.visit_chain(node, sympy.Mul)
            if node.is_Pow:
                return sympy.Pow(self.visit(node.args[0]), self.visit(node.args[1]))
            if node.is_Number or node.is_Symbol:
                return node
            if node.is_Function:
                return self.visit_call(node)
        raise RuntimeError(f"Unknown node type {type(node)}")

    def visit_chain(self, node, op):
        args = [self.visit(a) for a in node.args]
        ret = op(args[0], args[1])
        for a in args[2:]:
            ret = op(ret, a)
        return ret

    def visit_call(self, node):
        name = node.func.__name__
        if name == "n_dot_grad":
            return self.visit_n_dot_grad(node)
        if name == "n_dot":
            return sum(self.visit(a) * fl.n[k] for k, a in enumerate(node.args))
        # default: visit the arguments, keep the function
        return node.func(*[self.visit(a) for a in node.args])

    def visit_n_dot_grad(self, node):
        assert len(node.args) == 1
        fx = node.args[0]
        f = fx.func
        assert len(fx.args) == 1 and fx.args[0] == self.x, "n_dot_grad expects f(x)"
        return (f(self.x1) - f(self.x0)) / self.edge_length
