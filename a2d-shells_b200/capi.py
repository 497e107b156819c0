"""ctypes face of include/a2ds.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

STIFFNESS_MATRIX = 0
GEOMETRIC_STIFFNESS_MATRIX = 1
MASS_MATRIX = 2
QUAD4_SHELL = 0
QUAD4_NONLINEAR_SHELL = 1
TRANSFORM_NATURAL = 0
TRANSFORM_REF_AXIS = 1
SCATTER_ATOMIC = 0
SCATTER_COLORED = 1
SCATTER_ATOMIC_COLOR_ORDER = 2

# every symbol include/a2ds.h declares (tests check the library exports them all)
SYMBOLS = [
    "a2ds_last_error", "a2ds_version", "a2ds_create", "a2ds_destroy", "a2ds_synchronize",
    "a2ds_set_mesh", "a2ds_set_mesh_order", "a2ds_set_dependent_nodes", "a2ds_set_nodes", "a2ds_set_components", "a2ds_set_state",
    "a2ds_set_state_dev", "a2ds_set_bcs", "a2ds_set_scatter_mode", "a2ds_mat_create",
    "a2ds_mat_create_natural", "a2ds_mat_pattern", "a2ds_mat_nnz", "a2ds_mat_zero",
    "a2ds_mat_download", "a2ds_mat_values_dev", "a2ds_mat_download_rows", "a2ds_assemble_res", "a2ds_assemble_jacobian",
    "a2ds_assemble_mat_type", "a2ds_assemble_all", "a2ds_res_dev", "a2ds_state_dev",
    "a2ds_comm_unique_id", "a2ds_comm_init", "a2ds_set_halo", "a2ds_halo_forward",
    "a2ds_halo_from_distribute",
    "a2ds_last_timing", "a2ds_host_pattern", "a2ds_host_color_elements",
    "a2ds_host_color_elements_hashed", "a2ds_get_element_colors", "a2ds_set_double_buffer",
    "a2ds_last_kernel_ms", "a2ds_region_begin", "a2ds_region_end",
    "a2ds_mat_copy", "a2ds_mat_axpy", "a2ds_mat_apply_bcs", "a2ds_mat_mult_dev", "a2ds_mat_mult",
    "a2ds_add_jacobian_vec_product", "a2ds_add_jacobian_vec_product_dev",
    "a2ds_set_mass_moments", "a2ds_set_state_rates", "a2ds_assemble_mat_combo",
    "a2ds_mat_mult_dist_dev", "a2ds_mat_set_halo",
    "a2ds_mesh_read_bdf", "a2ds_mesh_read_bin", "a2ds_mesh_write_bin", "a2ds_mesh_from_arrays",
    "a2ds_mesh_free", "a2ds_mesh_sizes", "a2ds_mesh_connectivity", "a2ds_mesh_bcs",
    "a2ds_mesh_file_numbers", "a2ds_mesh_component", "a2ds_mesh_quad4",
    "a2ds_partition_build", "a2ds_partition_free", "a2ds_partition_sizes", "a2ds_partition_mesh",
    "a2ds_partition_halo", "a2ds_partition_apply", "a2ds_partition_rcb",
    "a2ds_partition_build_matrix", "a2ds_partition_matrix", "a2ds_partition_create_mat",
]

_LIB = None


class A2dsError(RuntimeError):
    pass


def lib_path():
    # A2DS_LIB: development switch to try an alternative build of the same library
    return os.environ.get("A2DS_LIB") or os.path.join(HERE, "lib", "liba2ds_b200.so")


def _prefer_bundled_nccl():
    """The library links libnccl.so.2.  If a newer NCCL ships with the Python environment
    (the one PyTorch is built against), make that the process-wide copy BEFORE ours resolves
    the soname: otherwise a later `import torch` finds the older system NCCL already loaded
    and fails on missing symbols.  Harmless when torch was imported first (already loaded)."""
    import sysconfig
    cand = os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "nccl", "lib", "libnccl.so.2")
    if os.path.exists(cand):
        try:
            C.CDLL(cand, mode=C.RTLD_GLOBAL)
        except OSError:
            pass


def load_library():
    """Load the CUDA library.  Fails loudly if it has not been built — there is no
    fallback implementation."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise A2dsError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        _prefer_bundled_nccl()
        L = C.CDLL(path)
        L.a2ds_last_error.restype = C.c_char_p
        L.a2ds_version.restype = C.c_char_p
        L.a2ds_mesh_free.restype = None
        L.a2ds_partition_free.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def host_pattern(n_nodes, conn):
    """natural-order BCSR pattern (rowp, cols) — host only, no device needed"""
    L = load_library()
    conn = _i32(conn).reshape(-1, 4)
    rowp = np.zeros(n_nodes + 1, dtype=np.int32)
    nnz = C.c_longlong()
    if L.a2ds_host_pattern(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(rowp), None,
                           C.byref(nnz)):
        raise A2dsError(L.a2ds_last_error().decode())
    cols = np.zeros(nnz.value, dtype=np.int32)
    L.a2ds_host_pattern(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(rowp), _p(cols),
                        C.byref(nnz))
    return rowp, cols


def host_color_elements(n_nodes, conn):
    L = load_library()
    conn = _i32(conn).reshape(-1, 4)
    color = np.zeros(conn.shape[0], dtype=np.int32)
    nc = C.c_int()
    if L.a2ds_host_color_elements(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(color),
                                  C.byref(nc)):
        raise A2dsError(L.a2ds_last_error().decode())
    return color, nc.value


def host_color_elements_hashed(n_nodes, conn):
    """the device's colouring rule (hashed-priority greedy) stepped on the host"""
    L = load_library()
    conn = _i32(conn).reshape(-1, 4)
    color = np.zeros(conn.shape[0], dtype=np.int32)
    nc = C.c_int()
    if L.a2ds_host_color_elements_hashed(C.c_int(n_nodes), C.c_int(conn.shape[0]), _p(conn), _p(color),
                                         C.byref(nc)):
        raise A2dsError(L.a2ds_last_error().decode())
    return color, nc.value


class Mesh:
    """A mesh container of the library (include/a2ds.h, a2ds_mesh_*): what
    TACSMeshLoader::scanBDFFile + getConnectivity + getBCs give (src/io/TACSMeshLoader.cpp).
    Host only — no device needed.  Arrays are copied out on construction:
      elem_ptr, elem_conn, elem_comp, X (n, 3), bc_nodes, bc_ptr, bc_vars, bc_vals,
      node_nums / elem_nums (the file's own numbers, 0-based), elem_descript, comp_descript."""

    def __init__(self, handle):
        L = self.L = load_library()
        self.h = handle
        n = [C.c_int() for _ in range(6)]
        self._chk(L.a2ds_mesh_sizes(self.h, *[C.byref(x) for x in n]))
        (self.n_nodes, self.n_elems, self.conn_size, self.n_bcs, self.bc_size,
         self.n_comp) = [x.value for x in n]
        IP, DP = C.POINTER(C.c_int), C.POINTER(C.c_double)

        def ints(ptr, k):
            return np.ctypeslib.as_array(ptr, shape=(k,)).copy() if k else np.zeros(0, np.int32)

        def reals(ptr, k):
            return np.ctypeslib.as_array(ptr, shape=(k,)).copy() if k else np.zeros(0)

        a, b, c, x = IP(), IP(), IP(), DP()
        self._chk(L.a2ds_mesh_connectivity(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(x)))
        self.elem_ptr = ints(a, self.n_elems + 1)
        self.elem_conn = ints(b, self.conn_size)
        self.elem_comp = ints(c, self.n_elems)
        self.X = reals(x, 3 * self.n_nodes).reshape(-1, 3)
        a, b, c, x = IP(), IP(), IP(), DP()
        self._chk(L.a2ds_mesh_bcs(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(x)))
        self.bc_nodes = ints(a, self.n_bcs)
        self.bc_ptr = ints(b, self.n_bcs + 1)
        self.bc_vars = ints(c, self.bc_size)
        self.bc_vals = reals(x, self.bc_size)
        a, b = IP(), IP()
        self._chk(L.a2ds_mesh_file_numbers(self.h, C.byref(a), C.byref(b)))
        self.node_nums = ints(a, self.n_nodes)
        self.elem_nums = ints(b, self.n_elems)
        self.elem_descript, self.comp_descript = [], []
        for k in range(self.n_comp):
            e, d = C.c_char_p(), C.c_char_p()
            self._chk(L.a2ds_mesh_component(self.h, C.c_int(k), C.byref(e), C.byref(d)))
            self.elem_descript.append(e.value.decode())
            self.comp_descript.append(d.value.decode())

    def _chk(self, rc):
        if rc != 0:
            raise A2dsError(self.L.a2ds_last_error().decode())

    @classmethod
    def _open(cls, fn, *args):
        L = load_library()
        h = C.c_void_p()
        if getattr(L, fn)(*args, C.byref(h)):
            raise A2dsError(L.a2ds_last_error().decode())
        return cls(h)

    @classmethod
    def read_bdf(cls, path, n_threads=0):
        """TACSMeshLoader::scanBDFFile"""
        return cls._open("a2ds_mesh_read_bdf", C.c_char_p(os.fsencode(path)), C.c_int(n_threads))

    @classmethod
    def read_bin(cls, path):
        return cls._open("a2ds_mesh_read_bin", C.c_char_p(os.fsencode(path)))

    @classmethod
    def from_arrays(cls, conn, X, elem_comp=None, bc_nodes=(), bc_vars=(), bc_vals=None):
        """conn: (n_elems, nodes per element); bc_vars: per BC entry a list of 0-based DOFs;
        bc_vals: per entry one value or a list matching bc_vars"""
        conn = _i32(conn)
        conn = conn.reshape(conn.shape[0], -1) if conn.size else conn.reshape(0, 4)
        X = _f64(X).reshape(-1, 3)
        ptr = _i32(np.arange(conn.shape[0] + 1) * conn.shape[1])
        ec = None if elem_comp is None else _i32(elem_comp)
        nb = len(bc_nodes)
        bptr = _i32(np.concatenate([[0], np.cumsum([len(v) for v in bc_vars])])) if nb else _i32([0])
        bvars = _i32(np.concatenate([np.asarray(v, dtype=np.int32) for v in bc_vars])) if nb else _i32([])
        if bc_vals is None:
            bvals = np.zeros(len(bvars))
        else:
            bvals = np.concatenate([np.broadcast_to(np.asarray(v, dtype=float), (len(w),))
                                    for v, w in zip(bc_vals, bc_vars)]) if nb else np.zeros(0)
        return cls._open("a2ds_mesh_from_arrays", C.c_int(X.shape[0]), C.c_int(conn.shape[0]),
                         _p(ptr), _p(conn), _p(ec), _p(X), C.c_int(nb), _p(_i32(bc_nodes)),
                         _p(bptr), _p(bvars), _p(_f64(bvals)))

    def write_bin(self, path):
        self._chk(self.L.a2ds_mesh_write_bin(self.h, C.c_char_p(os.fsencode(path))))

    def quad4(self):
        """(conn (n_elems, 4), bc_masks, bc_vals (n_bcs, 6)) for Assembler.set_mesh / set_bcs"""
        conn = np.zeros((self.n_elems, 4), dtype=np.int32)
        masks = np.zeros(self.n_bcs, dtype=np.int32)
        vals = np.zeros((self.n_bcs, 6))
        self._chk(self.L.a2ds_mesh_quad4(self.h, _p(conn), _p(masks), _p(vals)))
        return conn, masks, vals

    def close(self):
        if self.h:
            self.L.a2ds_mesh_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def halo_from_distribute(lo, n_owned, ext_proc, ext_ptr, ext_count, req_proc, req_ptr, req_count,
                         req_vars):
    """TACSBVecDistribute's slab lists (global node numbers) -> (peers, send_lists, recv_lists)
    in the device path's local numbering, as a2ds_set_halo takes them (host only)"""
    L = load_library()
    ext_proc, ext_ptr, ext_count = _i32(ext_proc), _i32(ext_ptr), _i32(ext_count)
    req_proc, req_ptr, req_count, req_vars = _i32(req_proc), _i32(req_ptr), _i32(req_count), _i32(req_vars)
    nmax = len(ext_proc) + len(req_proc)
    peers = np.zeros(max(nmax, 1), dtype=np.int32)
    sp = np.zeros(nmax + 1, dtype=np.int32); rp = np.zeros(nmax + 1, dtype=np.int32)
    sn = np.zeros(max(int(req_count.sum()), 1), dtype=np.int32)
    rn = np.zeros(max(int(ext_count.sum()), 1), dtype=np.int32)
    n = C.c_int()
    if L.a2ds_halo_from_distribute(C.c_int(lo), C.c_int(n_owned), C.c_int(len(ext_proc)), _p(ext_proc),
                                   _p(ext_ptr), _p(ext_count), C.c_int(len(req_proc)), _p(req_proc),
                                   _p(req_ptr), _p(req_count), _p(req_vars), C.byref(n), _p(peers),
                                   _p(sp), _p(sn), _p(rp), _p(rn)):
        raise A2dsError(L.a2ds_last_error().decode())
    k = n.value
    return (peers[:k].copy(), [sn[sp[i]:sp[i + 1]].copy() for i in range(k)],
            [rn[rp[i]:rp[i + 1]].copy() for i in range(k)])


def partition_rcb(conn, X, n_ranks):
    """element -> rank by recursive coordinate bisection (a2ds_partition_rcb); host only"""
    L = load_library()
    conn = _i32(conn).reshape(-1, 4)
    X = _f64(X).reshape(-1, 3)
    out = np.zeros(conn.shape[0], dtype=np.int32)
    if L.a2ds_partition_rcb(C.c_int(X.shape[0]), C.c_int(conn.shape[0]), _p(conn), _p(X),
                            C.c_int(n_ranks), _p(out)):
        raise A2dsError(L.a2ds_last_error().decode())
    return out


class Partition:
    """One rank's sub-mesh and halo plan from the global mesh (a2ds_partition_*; node ownership
    as TACSCreator, src/TACSCreator.cpp:1156-1205).  Host only.  Attributes: n_nodes, n_owned,
    elems, conn_local (n, 4), glob, ghost_owner, peers, send_lists, recv_lists."""

    def __init__(self, conn, n_nodes, elem_rank, n_ranks, rank, matrix_halo=False):
        """matrix_halo: TACSParallelMat flavour — extended ghost set, local pattern (rowp, cols)
        and per-peer block lists (mat_send_lists, mat_recv_lists); create_mat(asm) then gives a
        matrix whose owned rows are fully assembled by every assemble call"""
        L = self.L = load_library()
        conn = _i32(conn).reshape(-1, 4)
        er = _i32(elem_rank)
        assert len(er) == len(conn)
        self.h = C.c_void_p()
        build = L.a2ds_partition_build_matrix if matrix_halo else L.a2ds_partition_build
        if build(C.c_int(n_nodes), C.c_int(len(conn)), _p(conn), _p(er), C.c_int(n_ranks),
                 C.c_int(rank), C.byref(self.h)):
            raise A2dsError(L.a2ds_last_error().decode())
        n = [C.c_int() for _ in range(6)]
        L.a2ds_partition_sizes(self.h, *[C.byref(x) for x in n])
        self.n_nodes, self.n_owned, ne, npeer, ns, nr = [x.value for x in n]
        IP = C.POINTER(C.c_int)

        def ints(ptr, k):
            return np.ctypeslib.as_array(ptr, shape=(k,)).copy() if k else np.zeros(0, np.int32)

        a, b, c, d = IP(), IP(), IP(), IP()
        L.a2ds_partition_mesh(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.elems = ints(a, ne)
        self.conn_local = ints(b, 4 * ne).reshape(-1, 4)
        self.glob = ints(c, self.n_nodes)
        self.ghost_owner = ints(d, self.n_nodes - self.n_owned)
        a, b, c, d, e = IP(), IP(), IP(), IP(), IP()
        L.a2ds_partition_halo(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e))
        self.peers = ints(a, npeer)
        sp, sn, rp, rn = ints(b, npeer + 1), ints(c, ns), ints(d, npeer + 1), ints(e, nr)
        self.send_lists = [sn[sp[k]:sp[k + 1]] for k in range(npeer)]
        self.recv_lists = [rn[rp[k]:rp[k + 1]] for k in range(npeer)]
        if matrix_halo:
            a, b, c, d, e, f = IP(), IP(), IP(), IP(), IP(), IP()
            if L.a2ds_partition_matrix(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d),
                                       C.byref(e), C.byref(f)):
                raise A2dsError(L.a2ds_last_error().decode())
            self.rowp = ints(a, self.n_nodes + 1)
            self.cols = ints(b, int(self.rowp[-1]) if self.n_nodes else 0)
            sp, rp = ints(c, npeer + 1), ints(e, npeer + 1)
            sb, rb = ints(d, int(sp[-1])), ints(f, int(rp[-1]))
            self.mat_send_lists = [sb[sp[k]:sp[k + 1]] for k in range(npeer)]
            self.mat_recv_lists = [rb[rp[k]:rp[k + 1]] for k in range(npeer)]

    def create_mat(self, asm):
        """a matrix with the partition's pattern and matrix halo on an Assembler"""
        m = C.c_int()
        asm._chk(self.L.a2ds_partition_create_mat(asm.ctx, self.h, C.byref(m)))
        return m.value

    def apply(self, asm, elem_comp=None):
        """a2ds_set_mesh + a2ds_set_halo on an Assembler"""
        ec = None if elem_comp is None else _i32(elem_comp)
        asm._chk(self.L.a2ds_partition_apply(asm.ctx, self.h, _p(ec)))
        asm.n_nodes, asm.n_owned, asm.n_elems = self.n_nodes, self.n_owned, len(self.elems)

    def close(self):
        if self.h:
            self.L.a2ds_partition_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Assembler:
    """One device context = one rank's sub-mesh on one GPU.  Method names follow
    TACSAssembler (src/TACSAssembler.h:213-220)."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = C.c_void_p()
        self._chk(self.L.a2ds_create(C.c_int(device), C.byref(self.ctx)))
        self.n_nodes = self.n_owned = self.n_elems = 0
        self._keep = []

    def _chk(self, rc):
        if rc != 0:
            raise A2dsError(self.L.a2ds_last_error().decode())

    def close(self):
        if self.ctx:
            self.L.a2ds_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mesh -------------------------------------------------------------------
    def set_mesh(self, conn, n_nodes, n_owned=None, elem_comp=None, order=2):
        """order 2: conn[ne, 4] (MITC4); order 3: conn[ne, 9] (TACSQuad9Shell, residual + tangent)"""
        conn = _i32(conn).reshape(-1, order * order)
        self.n_elems = conn.shape[0]
        self.n_nodes = int(n_nodes)
        self.n_owned = self.n_nodes if n_owned is None else int(n_owned)
        ec = None if elem_comp is None else _i32(elem_comp)
        if order == 2:
            self._chk(self.L.a2ds_set_mesh(self.ctx, C.c_int(self.n_nodes), C.c_int(self.n_owned),
                                           C.c_int(self.n_elems), _p(conn), _p(ec)))
        else:
            self._chk(self.L.a2ds_set_mesh_order(self.ctx, C.c_int(order), C.c_int(self.n_nodes),
                                                 C.c_int(self.n_owned), C.c_int(self.n_elems),
                                                 _p(conn), _p(ec)))

    def set_dependent_nodes(self, dep_ptr, dep_conn, dep_weights):
        """TACSAssembler::setDependentNodes: dependent node d = sum_j w[j] * node dep_conn[j] over
        dep_ptr[d] .. dep_ptr[d + 1]; the connectivity of the FOLLOWING set_mesh refers to it as
        -(d + 1).  (None, None, None) withdraws the declaration."""
        if dep_ptr is None:
            self._chk(self.L.a2ds_set_dependent_nodes(self.ctx, C.c_int(0), None, None, None))
            return
        dp, dc, dw = _i32(dep_ptr), _i32(dep_conn), _f64(dep_weights)
        assert len(dc) == len(dw) == dp[-1]
        self._chk(self.L.a2ds_set_dependent_nodes(self.ctx, C.c_int(len(dp) - 1), _p(dp), _p(dc), _p(dw)))

    def set_nodes(self, X):
        X = _f64(X).reshape(-1, 3)
        assert X.shape[0] == self.n_nodes
        self._chk(self.L.a2ds_set_nodes(self.ctx, _p(X)))

    def set_components(self, Cs, eth, temperature=None, elem_class=None,
                       transform=TRANSFORM_NATURAL, ref_axis=None):
        Cs = _f64(Cs).reshape(-1, 22)
        eth = _f64(eth).reshape(-1, 9)
        n = Cs.shape[0]
        T = _f64(np.zeros(n) if temperature is None else np.broadcast_to(temperature, (n,)))
        cl = _i32(np.zeros(n) if elem_class is None else np.broadcast_to(elem_class, (n,)))
        ax = None if ref_axis is None else _f64(ref_axis)
        self._chk(self.L.a2ds_set_components(self.ctx, C.c_int(n), _p(Cs), _p(eth), _p(T), _p(cl),
                                             C.c_int(transform), _p(ax)))

    def set_state(self, u):
        u = _f64(u).reshape(-1, 6)
        self._chk(self.L.a2ds_set_state(self.ctx, C.c_int(u.shape[0]), _p(u)))
        self._keep = [u]  # the copy is asynchronous

    def set_mass_moments(self, moments):
        """mass moments per component (TACSShellConstitutive::evalMassMoments)"""
        m = _f64(moments).reshape(-1, 3)
        self._chk(self.L.a2ds_set_mass_moments(self.ctx, C.c_int(m.shape[0]), _p(m)))

    def set_state_rates(self, udot=None, uddot=None):
        """qdot, qddot of TACSAssembler::setVariables; uddot=None removes the inertial term"""
        n = self.n_nodes
        ud = None if udot is None else _f64(udot).reshape(-1, 6)
        udd = None if uddot is None else _f64(uddot).reshape(-1, 6)
        if udd is not None:
            n = udd.shape[0]
        self._chk(self.L.a2ds_set_state_rates(self.ctx, C.c_int(n), _p(ud), _p(udd)))
        self.synchronize()

    def set_state_ptr(self, n_given, host_ptr):
        """state from a raw (pinned) host pointer; asynchronous"""
        self._chk(self.L.a2ds_set_state(self.ctx, C.c_int(n_given), C.c_void_p(host_ptr)))

    def set_state_dev(self, n_given, dev_ptr):
        self._chk(self.L.a2ds_set_state_dev(self.ctx, C.c_int(n_given), C.c_void_p(dev_ptr)))

    def set_bcs(self, nodes, vars_mask, vals=None):
        nodes = _i32(nodes)
        vm = _i32(np.broadcast_to(vars_mask, nodes.shape))
        vals = _f64(np.zeros((len(nodes), 6)) if vals is None else vals).reshape(-1, 6)
        self._chk(self.L.a2ds_set_bcs(self.ctx, C.c_int(len(nodes)), _p(nodes), _p(vm), _p(vals)))

    def set_scatter_mode(self, mode):
        self._chk(self.L.a2ds_set_scatter_mode(self.ctx, C.c_int(mode)))

    def set_double_buffer(self, on):
        """matrices double buffered on the device (default on); off frees the spare value arrays"""
        self._chk(self.L.a2ds_set_double_buffer(self.ctx, C.c_int(1 if on else 0)))

    def element_colors(self):
        """element colours of the coloured scatter modes, computed on the device"""
        color = np.zeros(self.n_elems, dtype=np.int32)
        nc = C.c_int()
        self._chk(self.L.a2ds_get_element_colors(self.ctx, _p(color), C.byref(nc)))
        return color, nc.value

    # -- matrices -----------------------------------------------------------------
    def create_mat(self):
        """natural-order BCSR over the local nodes (TACSAssembler::createMat pattern)"""
        m = C.c_int()
        self._chk(self.L.a2ds_mat_create_natural(self.ctx, C.byref(m)))
        return m.value

    def create_mat_from_pattern(self, blocks):
        """blocks: list of dicts with nrows,rowp,cols and optional row_map,col_map,ident —
        the pattern of a host matrix (e.g. the B/E/F/C blocks of a TACSSchurMat)"""
        nb = len(blocks)
        keep = []
        IntP = C.POINTER(C.c_int)

        def arr(key, b):
            a = b.get(key)
            if a is None:
                return IntP()
            a = _i32(a)
            keep.append(a)
            return a.ctypes.data_as(IntP)

        nrows = _i32([b["nrows"] for b in blocks])
        rowp = (IntP * nb)(*[arr("rowp", b) for b in blocks])
        cols = (IntP * nb)(*[arr("cols", b) for b in blocks])
        rmap = (IntP * nb)(*[arr("row_map", b) for b in blocks])
        cmap = (IntP * nb)(*[arr("col_map", b) for b in blocks])
        ident = _i32([b.get("ident", 1 if i == 0 else 0) for i, b in enumerate(blocks)])
        m = C.c_int()
        self._chk(self.L.a2ds_mat_create(self.ctx, C.c_int(nb), _p(nrows), rowp, cols, rmap, cmap,
                                         _p(ident), C.byref(m)))
        return m.value

    def mat_pattern(self, mat, block=0):
        nr = C.c_int(); nnz = C.c_longlong()
        self._chk(self.L.a2ds_mat_pattern(self.ctx, C.c_int(mat), C.c_int(block), C.byref(nr),
                                          None, None))
        self._chk(self.L.a2ds_mat_nnz(self.ctx, C.c_int(mat), C.c_int(block), C.byref(nnz)))
        rowp = np.zeros(nr.value + 1, dtype=np.int32); cols = np.zeros(nnz.value, dtype=np.int32)
        self._chk(self.L.a2ds_mat_pattern(self.ctx, C.c_int(mat), C.c_int(block), C.byref(nr),
                                          _p(rowp), _p(cols)))
        return rowp, cols

    def mat_nnz(self, mat, block=0):
        nnz = C.c_longlong()
        self._chk(self.L.a2ds_mat_nnz(self.ctx, C.c_int(mat), C.c_int(block), C.byref(nnz)))
        return nnz.value

    def mat_zero(self, mat):
        self._chk(self.L.a2ds_mat_zero(self.ctx, C.c_int(mat)))

    def mat_values(self, mat, block=0, out=None):
        n = self.mat_nnz(mat, block)
        A = np.empty((n, 6, 6)) if out is None else out
        self._chk(self.L.a2ds_mat_download(self.ctx, C.c_int(mat), C.c_int(block), _p(A)))
        return A

    def mat_rows(self, mat, rows, rowp, block=0):
        """values of the listed block rows: list of arrays [nnz_row, 6, 6] (rowp: the block's
        row pointer, e.g. from mat_pattern)"""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cnt = (rowp[rows + 1] - rowp[rows]).astype(np.int64)
        A = np.empty((int(cnt.sum()), 6, 6))
        self._chk(self.L.a2ds_mat_download_rows(self.ctx, C.c_int(mat), C.c_int(block),
                                                C.c_int(len(rows)), _p(rows), _p(A)))
        return np.split(A, np.cumsum(cnt)[:-1])

    def mat_values_dev(self, mat, block=0):
        p = C.c_void_p()
        self._chk(self.L.a2ds_mat_values_dev(self.ctx, C.c_int(mat), C.c_int(block), C.byref(p)))
        return p.value

    # -- matrix algebra on the device values (TACSMat names) ------------------------------
    def mat_copy(self, dst, src):
        self._chk(self.L.a2ds_mat_copy(self.ctx, C.c_int(dst), C.c_int(src)))

    def mat_axpy(self, alpha, x, y):
        self._chk(self.L.a2ds_mat_axpy(self.ctx, C.c_double(alpha), C.c_int(x), C.c_int(y)))

    def mat_apply_bcs(self, mat):
        self._chk(self.L.a2ds_mat_apply_bcs(self.ctx, C.c_int(mat)))

    def mat_mult(self, mat, x, block=0):
        x = _f64(x).reshape(-1, 6)
        nr = C.c_int()
        self._chk(self.L.a2ds_mat_pattern(self.ctx, C.c_int(mat), C.c_int(block), C.byref(nr),
                                          None, None))
        y = np.empty((nr.value, 6))
        self._chk(self.L.a2ds_mat_mult(self.ctx, C.c_int(mat), C.c_int(block), C.c_int(x.shape[0]),
                                       _p(x), _p(y)))
        return y

    def mat_mult_dev(self, mat, x_dev, y_dev, block=0):
        self._chk(self.L.a2ds_mat_mult_dev(self.ctx, C.c_int(mat), C.c_int(block),
                                           C.c_void_p(x_dev), C.c_void_p(y_dev)))

    def mat_set_halo(self, mat, peers, send_lists, recv_lists):
        """matrix halo plan (ParallelMat flavour): per peer the indices of my ghost-row blocks
        to send and of the blocks arriving contributions are added to"""
        peers = _i32(peers)
        sp = _i32(np.concatenate([[0], np.cumsum([len(x) for x in send_lists])]))
        rp = _i32(np.concatenate([[0], np.cumsum([len(x) for x in recv_lists])]))
        sb = _i32(np.concatenate(list(send_lists))) if len(peers) and sp[-1] else _i32([])
        rb = _i32(np.concatenate(list(recv_lists))) if len(peers) and rp[-1] else _i32([])
        self._chk(self.L.a2ds_mat_set_halo(self.ctx, C.c_int(mat), C.c_int(len(peers)), _p(peers),
                                           _p(sp), _p(sb), _p(rp), _p(rb)))

    def mat_mult_dist_dev(self, mat, x_dev, y_dev):
        """distributed y = A x (owned rows), halo exchanges inside; device pointers"""
        self._chk(self.L.a2ds_mat_mult_dist_dev(self.ctx, C.c_int(mat), C.c_void_p(x_dev),
                                                C.c_void_p(y_dev)))

    # -- assembly (TACSAssembler names) ------------------------------------------------
    def _res_out(self, want):
        return np.empty((self.n_owned, 6)) if want else None

    def assembleRes(self, download=True):
        r = self._res_out(download)
        self._chk(self.L.a2ds_assemble_res(self.ctx, _p(r)))
        return r

    def assembleJacobian(self, alpha, beta, gamma, mat, download=True):
        r = self._res_out(download)
        self._chk(self.L.a2ds_assemble_jacobian(self.ctx, C.c_double(alpha), C.c_double(beta),
                                                C.c_double(gamma), _p(r), C.c_int(mat)))
        return r

    def assembleMatType(self, mat_type, mat):
        self._chk(self.L.a2ds_assemble_mat_type(self.ctx, C.c_int(mat_type), C.c_int(mat)))

    def assembleMatCombo(self, mat_types, scales, mat):
        """A = sum_i scales[i] * matType(mat_types[i])  (TACSAssembler::assembleMatCombo)"""
        t = _i32(mat_types)
        sc = _f64(scales)
        assert len(t) == len(sc)
        self._chk(self.L.a2ds_assemble_mat_combo(self.ctx, C.c_int(len(t)), _p(t), _p(sc),
                                                 C.c_int(mat)))

    def assembleAll(self, kmat, gmat, download=True, out_ptr=None):
        """residual + K + G in one pass; out_ptr: raw (pinned) host pointer for the residual"""
        if out_ptr is not None:
            self._chk(self.L.a2ds_assemble_all(self.ctx, C.c_void_p(out_ptr), C.c_int(kmat),
                                               C.c_int(gmat)))
            return None
        r = self._res_out(download)
        self._chk(self.L.a2ds_assemble_all(self.ctx, _p(r), C.c_int(kmat), C.c_int(gmat)))
        return r

    def addJacobianVecProduct(self, scale, alpha, x, y):
        """y <- y + scale * alpha * K x (matrix free), BC rows zeroed; returns the new y"""
        x = _f64(x).reshape(-1, 6)
        y = np.array(y, dtype=np.float64).reshape(-1, 6)
        assert x.shape[0] == self.n_nodes and y.shape[0] == self.n_nodes
        self._chk(self.L.a2ds_add_jacobian_vec_product(self.ctx, C.c_double(scale),
                                                       C.c_double(alpha), _p(x), _p(y)))
        return y

    def addJacobianVecProduct_dev(self, scale, alpha, x_dev, y_dev):
        self._chk(self.L.a2ds_add_jacobian_vec_product_dev(self.ctx, C.c_double(scale),
                                                           C.c_double(alpha), C.c_void_p(x_dev),
                                                           C.c_void_p(y_dev)))

    def res_dev(self):
        p = C.c_void_p()
        self._chk(self.L.a2ds_res_dev(self.ctx, C.byref(p)))
        return p.value

    def state_dev(self):
        p = C.c_void_p()
        self._chk(self.L.a2ds_state_dev(self.ctx, C.byref(p)))
        return p.value

    def synchronize(self):
        self._chk(self.L.a2ds_synchronize(self.ctx))

    def last_timing(self):
        ms = C.c_float(); n = C.c_int()
        self._chk(self.L.a2ds_last_timing(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self._chk(self.L.a2ds_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value

    def region_begin(self):
        self._chk(self.L.a2ds_region_begin(self.ctx))

    def region_end(self):
        ms = C.c_float()
        self._chk(self.L.a2ds_region_end(self.ctx, C.byref(ms)))
        return ms.value

    # -- multi-GPU ---------------------------------------------------------------------
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._chk(self.L.a2ds_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks, rank, uid):
        self._chk(self.L.a2ds_comm_init(self.ctx, C.c_int(n_ranks), C.c_int(rank),
                                        C.c_char_p(uid)))

    def set_halo(self, peers, send_lists, recv_lists):
        peers = _i32(peers)
        sp = _i32(np.concatenate([[0], np.cumsum([len(x) for x in send_lists])]))
        rp = _i32(np.concatenate([[0], np.cumsum([len(x) for x in recv_lists])]))
        sn = _i32(np.concatenate(send_lists)) if len(send_lists) else _i32([])
        rn = _i32(np.concatenate(recv_lists)) if len(recv_lists) else _i32([])
        self._chk(self.L.a2ds_set_halo(self.ctx, C.c_int(len(peers)), _p(peers), _p(sp), _p(sn),
                                       _p(rp), _p(rn)))

    def halo_forward(self):
        self._chk(self.L.a2ds_halo_forward(self.ctx))
