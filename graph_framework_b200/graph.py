"""Python view of include/graph_c_binding.h (the reference's C binding, same entry points).

Node objects wrap the opaque `graph_node` handles; arithmetic operators forward to graph_add /
graph_sub / graph_mul / graph_div.  Everything numeric happens inside libgfb200.so.
"""
import ctypes

import numpy as np

from ._lib import lib, c_double_p, GfbError

DOUBLE = 1


class Node:
    __slots__ = ("ctx", "h")

    def __init__(self, ctx, h):
        if not h:
            raise GfbError("null graph node")
        self.ctx, self.h = ctx, h

    def _other(self, o):
        return o if isinstance(o, Node) else self.ctx.constant(float(o))

    def __add__(self, o): return Node(self.ctx, lib.graph_add(self.ctx.c, self.h, self._other(o).h))
    def __radd__(self, o): return self._other(o) + self
    def __sub__(self, o): return Node(self.ctx, lib.graph_sub(self.ctx.c, self.h, self._other(o).h))
    def __rsub__(self, o): return self._other(o) - self
    def __mul__(self, o): return Node(self.ctx, lib.graph_mul(self.ctx.c, self.h, self._other(o).h))
    def __rmul__(self, o): return self._other(o)*self
    def __truediv__(self, o): return Node(self.ctx, lib.graph_div(self.ctx.c, self.h, self._other(o).h))
    def __rtruediv__(self, o): return self._other(o)/self
    def __neg__(self): return self.ctx.constant(-1.0)*self
    def __eq__(self, o): return isinstance(o, Node) and self.h == o.h
    def __hash__(self): return hash(self.h)

    def df(self, x):
        return Node(self.ctx, lib.graph_df(self.ctx.c, self.h, x.h))

    def evaluate(self, capacity=1 << 20):
        """leaf_node::evaluate on the host (node.hpp:378)."""
        buf = np.empty(capacity, dtype=np.float64)
        n = lib.graph_evaluate(self.ctx.c, self.h, buf.ctypes.data_as(c_double_p), capacity)
        return buf[:n].copy()


class Context:
    """graph_construct_context(DOUBLE, false) and the workflow calls that hang off it."""

    def __init__(self):
        self.c = lib.graph_construct_context(DOUBLE, False)
        if not self.c:
            raise GfbError("graph_construct_context failed")
        self._keep = []

    def close(self):
        if self.c:
            lib.graph_destroy_context(self.c)
            self.c = None

    def __del__(self):
        self.close()

    def variable(self, size, symbol, values=None):
        v = Node(self, lib.graph_variable(self.c, size, symbol.encode()))
        if values is not None:
            self.set_variable(v, values)
        return v

    def set_variable(self, var, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        lib.graph_set_variable(self.c, var.h, a.ctypes.data_as(ctypes.c_void_p))

    def constant(self, value):
        return Node(self, lib.graph_constant(self.c, float(value)))

    def pseudo_variable(self, x): return Node(self, lib.graph_pseudo_variable(self.c, x.h))
    def remove_pseudo(self, x): return Node(self, lib.graph_remove_pseudo(self.c, x.h))
    def sqrt(self, x): return Node(self, lib.graph_sqrt(self.c, x.h))
    def exp(self, x): return Node(self, lib.graph_exp(self.c, x.h))
    def log(self, x): return Node(self, lib.graph_log(self.c, x.h))
    def erfi(self, x): return Node(self, lib.graph_erfi(self.c, x.h))
    def sin(self, x): return Node(self, lib.graph_sin(self.c, x.h))
    def cos(self, x): return Node(self, lib.graph_cos(self.c, x.h))
    def pow(self, x, y): return Node(self, lib.graph_pow(self.c, x.h, x._other(y).h))
    def atan(self, x, y): return Node(self, lib.graph_atan(self.c, x.h, y.h))
    def fma(self, a, b, c): return Node(self, lib.graph_fma(self.c, a.h, b.h, c.h))

    def piecewise_1D(self, arg, scale, offset, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        return Node(self, lib.graph_piecewise_1D(self.c, arg.h, scale, offset, a.ctypes.data_as(ctypes.c_void_p), a.size))

    def piecewise_2D(self, num_cols, x, x_scale, x_offset, y, y_scale, y_offset, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        return Node(self, lib.graph_piecewise_2D(self.c, num_cols, x.h, x_scale, x_offset, y.h, y_scale, y_offset,
                                                 a.ctypes.data_as(ctypes.c_void_p), a.size))

    def index_1D(self, variable, arg, scale, offset):
        return Node(self, lib.graph_index_1D(self.c, variable.h, arg.h, scale, offset))

    def index_2D(self, variable, num_cols, x, x_scale, x_offset, y, y_scale, y_offset):
        return Node(self, lib.graph_index_2D(self.c, variable.h, num_cols, x.h, x_scale, x_offset, y.h, y_scale, y_offset))

    @staticmethod
    def _arr(nodes):
        t = ctypes.c_void_p*max(len(nodes), 1)
        return t(*[n.h for n in nodes])

    def _item(self, fn, inputs, outputs, setters, name, size, *extra):
        ins, outs = self._arr(inputs), self._arr(outputs)
        m_in = self._arr([v for _, v in setters])
        m_out = self._arr([e for e, _ in setters])
        fn(self.c, ins, len(inputs), outs, len(outputs), m_in, m_out, len(setters), None, name.encode(), size, *extra)

    def add_item(self, inputs, outputs, setters, name, size):
        """setters: list of (expression, variable) pairs, as graph::map_nodes."""
        self._item(lib.graph_add_item, inputs, outputs, setters, name, size)

    def add_pre_item(self, inputs, outputs, setters, name, size):
        self._item(lib.graph_add_pre_item, inputs, outputs, setters, name, size)

    def add_converge_item(self, inputs, outputs, setters, name, size, tol=1.0e-30, max_iter=1000):
        self._item(lib.graph_add_converge_item, inputs, outputs, setters, name, size, tol, max_iter)

    def set_fast_division(self, on):
        """graph_set_fast_division: False = IEEE division/sqrt and (x - offset)/scale table indices."""
        lib.graph_set_fast_division(self.c, bool(on))

    def compile(self): lib.graph_compile(self.c)
    def pre_run(self): lib.graph_pre_run(self.c)
    def run(self): lib.graph_run(self.c)
    def wait(self): lib.graph_wait(self.c)

    def copy_to_host(self, node, size):
        out = np.empty(size, dtype=np.float64)
        lib.graph_copy_to_host(self.c, node.h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def copy_to_device(self, node, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        lib.graph_copy_to_device(self.c, node.h, a.ctypes.data_as(ctypes.c_void_p))

    def source(self):
        return lib.graph_get_source(self.c).decode()

    def max_concurrency(self):
        return lib.graph_get_max_concurrency(self.c)
