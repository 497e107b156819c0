"""ctypes binding for oracle/liboracle_shell.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Comp(C.Structure):
    _fields_ = [("model", C.c_int), ("transform", C.c_int), ("axis", C.c_double * 3),
                ("Cs", C.c_double * 22), ("eth", C.c_double * 9), ("mom", C.c_double * 3),
                ("temperature", C.c_double)]


def make_comp(model, Cs, eth, mom=(0, 0, 0), temperature=0.0, transform=0, axis=(1.0, 0.0, 0.0)):
    c = Comp()
    c.model = int(model); c.transform = int(transform)
    ax = np.asarray(axis, dtype=np.float64)
    if transform == 1:
        ax = ax / np.sqrt(ax @ ax)
    for i in range(3):
        c.axis[i] = ax[i]; c.mom[i] = mom[i]
    for i in range(22):
        c.Cs[i] = Cs[i]
    for i in range(9):
        c.eth[i] = eth[i]
    c.temperature = float(temperature)
    return c


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_shell.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "oracle"])
        _LIB = C.CDLL(path)
    return _LIB


_LIB_FMA = None


def lib_fma():
    """the same C restatement compiled with FMA contraction (noise-floor measurements only)"""
    global _LIB_FMA
    if _LIB_FMA is None:
        path = os.path.join(_HERE, "liboracle_shell_fma.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "liboracle_shell_fma.so"])
        _LIB_FMA = C.CDLL(path)
    return _LIB_FMA


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fn(name, order):
    """entry point for the element order: 2 = 4-node (oracle_*), 3 = 9-node (oracle9_*)"""
    assert order in (2, 3)
    return getattr(lib(), ("oracle_" if order == 2 else "oracle9_") + name)


def strain(comp, X, q, order=2):
    nn = order * order
    e = np.zeros(9 * nn); det = np.zeros(nn)
    _fn("strain", order)(C.byref(comp), _p(np.ascontiguousarray(X, dtype=np.float64)),
                        _p(np.ascontiguousarray(q, dtype=np.float64)), _p(e), _p(det))
    return e.reshape(nn, 9), det


def residual(comp, X, q, order=2):
    r = np.zeros(6 * order * order)
    _fn("residual", order)(C.byref(comp), _p(np.ascontiguousarray(X, dtype=np.float64)),
                          _p(np.ascontiguousarray(q, dtype=np.float64)), _p(r))
    return r


def jacobian(comp, X, q, alpha=1.0, order=2):
    nv = 6 * order * order
    r = np.zeros(nv); m = np.zeros(nv * nv)
    _fn("jacobian", order)(C.byref(comp), C.c_double(alpha),
                           _p(np.ascontiguousarray(X, dtype=np.float64)),
                           _p(np.ascontiguousarray(q, dtype=np.float64)), _p(r), _p(m))
    return r, m.reshape(nv, nv)


def jacobian_dyn(comp, X, q, qdd, alpha=1.0, gamma=0.0, order=2):
    """res (static + M qdd) and alpha K + gamma M"""
    nv = 6 * order * order
    r = np.zeros(nv); m = np.zeros(nv * nv)
    _fn("jacobian_dyn", order)(C.byref(comp), C.c_double(alpha), C.c_double(gamma),
                              _p(np.ascontiguousarray(X, dtype=np.float64)),
                              _p(np.ascontiguousarray(q, dtype=np.float64)),
                              _p(np.ascontiguousarray(qdd, dtype=np.float64)), _p(r), _p(m))
    return r, m.reshape(nv, nv)


def mat_type(comp, type_, X, q, order=2):
    nv = 6 * order * order
    m = np.zeros(nv * nv)
    _fn("mat_type", order)(C.byref(comp), C.c_int(type_),
                           _p(np.ascontiguousarray(X, dtype=np.float64)),
                           _p(np.ascontiguousarray(q, dtype=np.float64)), _p(m))
    return m.reshape(nv, nv)


def pattern(n_nodes, conn, order=2):
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, order * order)
    rowp = np.zeros(n_nodes + 1, dtype=np.int32)
    f = _fn("pattern", order)
    nnz = f(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(rowp), None)
    cols = np.zeros(nnz, dtype=np.int32)
    f(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(rowp), _p(cols))
    return rowp, cols


def assemble(op, conn, elem_comp, comps, X, u, rowp, cols, bc_nodes=None, bc_vars=None,
             bc_vals=None, alpha=1.0, gamma=0.0, udd=None, fma=False, order=2):
    """op 0 res, 1 jacobian (alpha K + gamma M), 2 K, 3 G, 4 M -> (res[n,6] or None,
    A[nnz,6,6] or None); udd: second time derivatives (inertial term of the residual)."""
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, order * order)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 6)
    n = X.shape[0]
    f = getattr(lib_fma() if fma else lib(), "oracle_assemble_dyn" if order == 2 else "oracle9_assemble_dyn")
    arr = (Comp * len(comps))(*comps)
    ec = np.ascontiguousarray(elem_comp, dtype=np.int32)
    nb = 0 if bc_nodes is None else len(bc_nodes)
    bn = np.ascontiguousarray(bc_nodes if nb else np.zeros(0), dtype=np.int32)
    bv = np.ascontiguousarray(bc_vars if nb else np.zeros(0), dtype=np.int32)
    bx = np.ascontiguousarray(bc_vals if nb else np.zeros(0), dtype=np.float64)
    res = np.zeros((n, 6)) if op <= 1 else None
    A = np.zeros((len(cols), 6, 6)) if op >= 1 else None
    udd = None if udd is None else np.ascontiguousarray(udd, dtype=np.float64).reshape(-1, 6)
    miss = f(C.c_int(op), C.c_double(alpha), C.c_double(gamma), C.c_int(n),
                                     C.c_int(conn.shape[0]), _p(conn), _p(ec), arr, _p(X), _p(u),
                                     _p(udd), C.c_int(nb), _p(bn), _p(bv), _p(bx),
                                     _p(np.ascontiguousarray(rowp, dtype=np.int32)),
                                     _p(np.ascontiguousarray(cols, dtype=np.int32)), _p(res), _p(A))
    assert miss == 0, "element block missing from the pattern"
    return res, A


def pattern_dep(n_nodes, conn, dep_ptr, dep_conn, order=2):
    """pattern of a mesh with dependent nodes (connectivity entries -(d + 1))"""
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, order * order)
    dp = np.ascontiguousarray(dep_ptr, dtype=np.int32); dc = np.ascontiguousarray(dep_conn, dtype=np.int32)
    rowp = np.zeros(n_nodes + 1, dtype=np.int32)
    f = _fn("pattern_dep", order)
    nnz = f(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(dp), _p(dc), _p(rowp), None)
    cols = np.zeros(nnz, dtype=np.int32)
    f(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(dp), _p(dc), _p(rowp), _p(cols))
    return rowp, cols


def assemble_dep(op, conn, elem_comp, comps, X, u, dep, rowp, cols, bc_nodes=None, bc_vars=None,
                 bc_vals=None, alpha=1.0, gamma=0.0, udd=None, order=2):
    """assemble() on a mesh with dependent nodes, dep = (dep_ptr, dep_conn, dep_weights);
    X, u, udd and the residual have one row per independent node"""
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, order * order)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 6)
    n = X.shape[0]
    dp = np.ascontiguousarray(dep[0], dtype=np.int32); dc = np.ascontiguousarray(dep[1], dtype=np.int32)
    dw = np.ascontiguousarray(dep[2], dtype=np.float64)
    arr = (Comp * len(comps))(*comps)
    ec = np.ascontiguousarray(elem_comp, dtype=np.int32)
    nb = 0 if bc_nodes is None else len(bc_nodes)
    bn = np.ascontiguousarray(bc_nodes if nb else np.zeros(0), dtype=np.int32)
    bv = np.ascontiguousarray(bc_vars if nb else np.zeros(0), dtype=np.int32)
    bx = np.ascontiguousarray(bc_vals if nb else np.zeros(0), dtype=np.float64)
    res = np.zeros((n, 6)) if op <= 1 else None
    A = np.zeros((len(cols), 6, 6)) if op >= 1 else None
    udd = None if udd is None else np.ascontiguousarray(udd, dtype=np.float64).reshape(-1, 6)
    miss = _fn("assemble_dep", order)(
        C.c_int(op), C.c_double(alpha), C.c_double(gamma), C.c_int(n), C.c_int(conn.shape[0]),
        _p(conn), _p(ec), arr, _p(X), _p(u), _p(udd), C.c_int(len(dp) - 1), _p(dp), _p(dc), _p(dw),
        C.c_int(nb), _p(bn), _p(bv), _p(bx), _p(np.ascontiguousarray(rowp, dtype=np.int32)),
        _p(np.ascontiguousarray(cols, dtype=np.int32)), _p(res), _p(A))
    assert miss == 0, "element block missing from the pattern"
    return res, A
