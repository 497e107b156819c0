"""ctypes binding for oracle/_ref/libref_driver.so — TEST INFRASTRUCTURE ONLY.

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's
CPU arm may import this module.  It wraps the UNMODIFIED reference build
(oracle/_ref/liba2dshells_ref.so, see oracle/Makefile) behind numpy arrays.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NPROP = 40

_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_driver.so"))


def lib():
    global _LIB
    if _LIB is None:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        # the bundled OpenBLAS needs its sibling libgfortran: preload both by path
        import glob
        import sysconfig
        bl = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*", "libopenblasp-*.so"):
            for f in sorted(glob.glob(os.path.join(bl, pat))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
        C.CDLL(os.path.join(_HERE, "_ref", "liba2dshells_ref.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(os.path.join(_HERE, "_ref", "libref_driver.so"))
        L.refdrv_create.restype = C.c_void_p
        L.refdrv_time.restype = C.c_double
        L.refdrv_element_batch.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def iso_props(kind=0, E=72e9, nu=0.33, rho=2718.0, cte=10e-6, t=0.010, t_offset=0.0,
              temperature=0.0):
    """Component record for an isotropic shell (mechBuckling.cpp:41-58 defaults)."""
    p = np.zeros(NPROP)
    p[0] = kind
    p[1] = 0
    p[2] = temperature
    p[3:9] = [E, nu, rho, cte, t, t_offset]
    return p


def general_props(kind, Cs, eth3, moments, temperature=0.0):
    p = np.zeros(NPROP)
    p[0] = kind
    p[1] = 1
    p[2] = temperature
    p[9:31] = Cs
    p[31:34] = eth3
    p[34:37] = moments
    return p


def con_tables(props):
    Cs = np.zeros(22); eth = np.zeros(9); mom = np.zeros(3)
    lib().refdrv_con_tables(_p(np.ascontiguousarray(props)), _p(Cs), _p(eth), _p(mom))
    return Cs, eth, mom


def element(props, op, X, vars_, dvars=None, ddvars=None, alpha=1.0, beta=0.0, gamma=0.0,
            transform=0, axis=(1.0, 0.0, 0.0)):
    """op: 0 addResidual, 1 addJacobian, 2 K, 3 G, 4 M -> (res[nv], mat[nv,nv]); nv = 24 for
    the 4-node kinds (props[0] = 0, 1), 54 for the 9-node kinds (2: TACSQuad9Shell, 3: nonlinear)."""
    nv = 54 if int(props[0]) >= 2 else 24
    z = np.zeros(nv)
    dv = z if dvars is None else np.ascontiguousarray(dvars, dtype=np.float64)
    ddv = z if ddvars is None else np.ascontiguousarray(ddvars, dtype=np.float64)
    res = np.zeros(nv); mat = np.zeros(nv * nv)
    ax = np.asarray(axis, dtype=np.float64)
    lib().refdrv_element(_p(np.ascontiguousarray(props)), C.c_int(transform), _p(ax), C.c_int(op),
                         C.c_double(alpha), C.c_double(beta), C.c_double(gamma),
                         _p(np.ascontiguousarray(X, dtype=np.float64)),
                         _p(np.ascontiguousarray(vars_, dtype=np.float64)), _p(dv), _p(ddv),
                         _p(res), _p(mat))
    return res, mat.reshape(nv, nv)


def element_batch(props, op, X, vars_, alpha=1.0, beta=0.0, gamma=0.0, transform=0,
                  axis=(1.0, 0.0, 0.0), want_out=True):
    n = X.shape[0]
    X = np.ascontiguousarray(X.reshape(n, 12), dtype=np.float64)
    v = np.ascontiguousarray(vars_.reshape(n, 24), dtype=np.float64)
    res = np.zeros((n, 24)) if want_out else None
    mat = np.zeros((n, 24, 24)) if want_out else None
    ax = np.asarray(axis, dtype=np.float64)
    secs = lib().refdrv_element_batch(_p(np.ascontiguousarray(props)), C.c_int(transform), _p(ax),
                                      C.c_int(op), C.c_double(alpha), C.c_double(beta),
                                      C.c_double(gamma), C.c_int(n), _p(X), _p(v), _p(res), _p(mat))
    return res, mat, secs


class RefAssembler:
    """Reference TACSAssembler built through TACSCreator for a quad mesh."""

    def __init__(self, conn, X, elem_comp, comp_props, bc_nodes=None, bc_vars=None, bc_vals=None,
                 transform=0, axis=(1.0, 0.0, 0.0), nodes_per_elem=4, dep=None):
        """dep = (dep_ptr, dep_conn, dep_weights): dependent nodes (TACSCreator::setDependentNodes);
        connectivity entries -(d + 1) refer to dependent node d"""
        L = lib()
        if dep is not None:
            dp = np.ascontiguousarray(dep[0], dtype=np.int32)
            dc = np.ascontiguousarray(dep[1], dtype=np.int32)
            dw = np.ascontiguousarray(dep[2], dtype=np.float64)
            L.refdrv_next_dependent_nodes(C.c_int(len(dp) - 1), _p(dp), _p(dc), _p(dw))
        L.refdrv_create_n.restype = C.c_void_p
        self.nodes_per_elem = nodes_per_elem   # 9: TACSQuad9Shell components (props[0] = 2, 3)
        conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, nodes_per_elem)
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
        self.n_elems = conn.shape[0]
        self.n_nodes = X.shape[0]
        elem_comp = np.ascontiguousarray(elem_comp, dtype=np.int32)
        comp_props = np.ascontiguousarray(comp_props, dtype=np.float64).reshape(-1, NPROP)
        # BCs: bc_nodes[k]; bc_vars = list of lists of dof indices; bc_vals same shape
        if bc_nodes is None:
            bc_nodes = np.zeros(0, dtype=np.int32); bc_vars = []; bc_vals = []
        bc_nodes = np.ascontiguousarray(bc_nodes, dtype=np.int32)
        ptr = np.zeros(len(bc_nodes) + 1, dtype=np.int32)
        for k, v in enumerate(bc_vars):
            ptr[k + 1] = ptr[k] + len(v)
        flat_vars = np.ascontiguousarray(
            np.concatenate([np.asarray(v, dtype=np.int32) for v in bc_vars]) if len(bc_vars) else
            np.zeros(0, dtype=np.int32), dtype=np.int32)
        flat_vals = np.ascontiguousarray(
            np.concatenate([np.asarray(v, dtype=np.float64) for v in bc_vals]) if len(bc_vals) else
            np.zeros(0), dtype=np.float64)
        ax = np.asarray(axis, dtype=np.float64)
        self.h = C.c_void_p(L.refdrv_create_n(
            C.c_int(nodes_per_elem), C.c_int(self.n_nodes), C.c_int(self.n_elems), _p(conn), _p(elem_comp), _p(X),
            C.c_int(len(bc_nodes)), _p(bc_nodes), _p(ptr), _p(flat_vars), _p(flat_vals),
            C.c_int(comp_props.shape[0]), _p(comp_props), C.c_int(transform), _p(ax)))
        self.new_nodes = np.zeros(self.n_nodes, dtype=np.int32)
        L.refdrv_get_node_nums(self.h, _p(self.new_nodes))
        self._keep = (conn, X, elem_comp, comp_props)

    def conn(self):
        c = np.zeros((self.n_elems, self.nodes_per_elem), dtype=np.int32)
        lib().refdrv_get_conn(self.h, _p(c))
        return c

    def dep(self):
        """(dep_ptr, dep_conn, dep_weights) in the reference's numbering, or None"""
        nd = lib().refdrv_get_dep(self.h, None, None, None)
        if nd == 0:
            return None
        ptr = np.zeros(nd + 1, dtype=np.int32)
        lib().refdrv_get_dep(self.h, _p(ptr), None, None)
        cn = np.zeros(ptr[-1], dtype=np.int32); w = np.zeros(ptr[-1])
        lib().refdrv_get_dep(self.h, _p(ptr), _p(cn), _p(w))
        return ptr, cn, w

    def nodes(self):
        X = np.zeros((self.n_nodes, 3))
        lib().refdrv_get_nodes(self.h, _p(X))
        return X

    def bcs(self):
        nb = lib().refdrv_get_bcs(self.h, None, None, None)
        nodes = np.zeros(nb, dtype=np.int32); vars_ = np.zeros(nb, dtype=np.int32)
        vals = np.zeros((nb, 6))
        lib().refdrv_get_bcs(self.h, _p(nodes), _p(vars_), _p(vals))
        return nodes, vars_, vals

    def set_threads(self, nt):
        lib().refdrv_set_threads(self.h, C.c_int(nt))

    def set_state(self, u=None, ud=None, udd=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (u, ud, udd)]
        lib().refdrv_set_state(self.h, _p(a[0]), _p(a[1]), _p(a[2]))

    def set_temperature(self, T):
        lib().refdrv_set_temperature(self.h, C.c_double(T))

    def bc_state(self):
        u = np.zeros((self.n_nodes, 6))
        lib().refdrv_bc_state(self.h, _p(u))
        return u

    def mat_create(self, kind=0):
        return lib().refdrv_mat_create(self.h, C.c_int(kind))

    def mat_block(self, mat, which=0, values=True):
        nr = C.c_int(); nc = C.c_int(); nnz = C.c_int()
        if lib().refdrv_mat_info(self.h, C.c_int(mat), C.c_int(which), C.byref(nr), C.byref(nc),
                                 C.byref(nnz)):
            return None
        rowp = np.zeros(nr.value + 1, dtype=np.int32); cols = np.zeros(nnz.value, dtype=np.int32)
        A = np.zeros((nnz.value, 6, 6)) if values else None
        lib().refdrv_mat_get(self.h, C.c_int(mat), C.c_int(which), _p(rowp), _p(cols), _p(A))
        return dict(nrows=nr.value, ncols=nc.value, rowp=rowp, cols=cols, A=A)

    def mat_set(self, mat, which, A):
        lib().refdrv_mat_set(self.h, C.c_int(mat), C.c_int(which),
                             _p(np.ascontiguousarray(A, dtype=np.float64)))

    def schur_index(self, mat, which):
        n = lib().refdrv_schur_index(self.h, C.c_int(mat), C.c_int(which), None)
        out = np.zeros(max(n, 0), dtype=np.int32)
        if n > 0:
            lib().refdrv_schur_index(self.h, C.c_int(mat), C.c_int(which), _p(out))
        return out

    def assemble_res(self):
        r = np.zeros((self.n_nodes, 6))
        lib().refdrv_assemble_res(self.h, _p(r))
        return r

    def assemble_jacobian(self, mat, alpha=1.0, beta=0.0, gamma=0.0, transpose=False):
        r = np.zeros((self.n_nodes, 6))
        fn = lib().refdrv_assemble_jacobian_transpose if transpose else lib().refdrv_assemble_jacobian
        fn(self.h, C.c_double(alpha), C.c_double(beta), C.c_double(gamma), C.c_int(mat), _p(r))
        return r

    def assemble_mat_type(self, type_, mat):
        lib().refdrv_assemble_mat_type(self.h, C.c_int(type_), C.c_int(mat))

    def time(self, op, mat=0):
        return lib().refdrv_time(self.h, C.c_int(op), C.c_int(mat))

    def buckling(self, kmat, gmat, aux, mode, sigma=10.0, max_lanczos=100, num_eigs=10,
                 tol=1e-12, u0=None):
        eigs = np.zeros(num_eigs); errs = np.zeros(num_eigs)
        u0a = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
        self.path = np.zeros((self.n_nodes, 6))
        lib().refdrv_buckling(self.h, C.c_int(kmat), C.c_int(gmat), C.c_int(aux), C.c_int(mode),
                              C.c_double(sigma), C.c_int(max_lanczos), C.c_int(num_eigs),
                              C.c_double(tol), _p(u0a), _p(eigs), _p(errs), _p(self.path))
        return eigs, errs

    def frequency(self, kmat, mmat, sigma=0.0, max_lanczos=60, num_eigs=10, tol=1e-12):
        """TACSFrequencyAnalysis::solve (Lanczos); returns (omega^2[num_eigs], err)"""
        eig = np.zeros(num_eigs); err = np.zeros(num_eigs)
        lib().refdrv_frequency(self.h, C.c_int(kmat), C.c_int(mmat), C.c_double(sigma),
                               C.c_int(max_lanczos), C.c_int(num_eigs), C.c_double(tol),
                               _p(eig), _p(err))
        return eig, err

    def close(self):
        if self.h:
            lib().refdrv_destroy(self.h)
            self.h = None


def bdf_scan(path):
    """TACSMeshLoader::scanBDFFile + getConnectivity + getBCs of the unmodified reference.
    Returns (fail, dict of arrays, seconds spent in scanBDFFile)."""
    L = lib()
    h = C.c_void_p()
    sizes = np.zeros(6, dtype=np.int32)
    secs = C.c_double()
    fail = L.refdrv_bdf_scan(C.c_char_p(os.fsencode(path)), C.byref(h), _p(sizes), C.byref(secs))
    nn, ne, nc, nb, nv, ncomp = [int(x) for x in sizes]
    out = dict(elem_ptr=np.zeros(ne + 1, np.int32), elem_conn=np.zeros(nc, np.int32),
               elem_comp=np.zeros(ne, np.int32), X=np.zeros((nn, 3)),
               bc_nodes=np.zeros(nb, np.int32), bc_ptr=np.zeros(nb + 1, np.int32),
               bc_vars=np.zeros(nv, np.int32), bc_vals=np.zeros(nv))
    ed = C.create_string_buffer(9 * max(ncomp, 1))
    cd = C.create_string_buffer(33 * max(ncomp, 1))
    if fail == 0:
        L.refdrv_bdf_arrays(h, *[_p(out[k]) for k in ("elem_ptr", "elem_conn", "elem_comp", "X",
                                                       "bc_nodes", "bc_ptr", "bc_vars", "bc_vals")],
                            ed, cd)
    out["elem_descript"] = [ed.raw[9 * k:9 * k + 9].split(b"\0")[0].decode() for k in range(ncomp)]
    out["comp_descript"] = [cd.raw[33 * k:33 * k + 33].split(b"\0")[0].decode() for k in range(ncomp)]
    out["n_comp"] = ncomp
    L.refdrv_bdf_free(h)
    return fail, out, secs.value
