"""ctypes front-ends for the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (thirring2d_b200/) never does.

Two back-ends with the same method names:

* ``Oracle``  — oracle/liboracle.so, the C restatement in oracle/thirring_oracle.c (runtime lattice size).
* ``RefLib``  — oracle/_ref/libhmcref_<NT>x<NX>_<flavour>.so, the reference's own hmc.c compiled
  unmodified by oracle/build_ref.sh (compile-time lattice size, file-scope globals,
  /root/reference/hmc.c:38-50).  RefLib fills the globals exactly as main() does (hmc.c:888-925).

Array conventions: vectors are complex128 numpy arrays of shape (NT, NX); gauge fields are float64
arrays of shape (NT, NX, 2) with [...,0] the t-link angle and [...,1] the x-link angle.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

MODE_REF_COMPAT, MODE_ADJOINT = 0, 1
BC_ANTISYMMETRIC, BC_SYMMETRIC, BC_OPENX = 0, 1, 2   # Thirring.h:27-29
CG_CONVERGED, CG_MAXITER, CG_DIVERGED, CG_ZERO_SOURCE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def build(verbose=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


class Oracle:
    """The C restatement (oracle/thirring_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        ii, dd = C.c_int, C.c_double
        L.orc_fm_mul.argtypes = [ii, ii, dd, dd, _dp, _dp, _dp]
        L.orc_fm_dagger_mul.argtypes = [ii, ii, dd, dd, _dp, _dp, _dp]
        L.orc_fm_conjugate_mul.argtypes = [ii, ii, dd, dd, ii, _dp, _dp, _dp]
        L.orc_fmdm_invert_cg.argtypes = [ii, ii, dd, dd, ii, _dp, _dp, _dp, ii, C.POINTER(ii), _dp]
        L.orc_fmdm_invert_cg.restype = ii
        L.orc_fmdm_invert_cg_treesum.argtypes = L.orc_fmdm_invert_cg.argtypes
        L.orc_fmdm_invert_cg_treesum.restype = ii
        L.orc_fm_invert_cg.argtypes = L.orc_fmdm_invert_cg.argtypes
        L.orc_fm_invert_cg.restype = ii
        L.orc_fermion_matrix.argtypes = [ii, ii, dd, dd, _dp, _dp]
        L.orc_re_dot.argtypes = [ii, _dp, _dp]
        L.orc_re_dot.restype = dd
        ipp = C.POINTER(C.c_int)
        L.orc_fM.argtypes = [ii, ii, dd, dd, ipp, _dp, _dp]
        L.orc_fM_transpose.argtypes = [ii, ii, dd, dd, ipp, _dp, _dp]
        L.orc_cg_MdM.argtypes = [ii, ii, dd, dd, ipp, _dp, _dp, ipp, _dp]
        L.orc_cg_MdM.restype = ii
        L.orc_cg_propagator.argtypes = [ii, ii, dd, dd, ipp, _dp, _dp, ipp, _dp]
        L.orc_cg_propagator.restype = ii
        L.orc_fM_occupied.argtypes = [ii, ii, dd, ipp, _dp, _dp]
        L.orc_fM_occupied_sq.argtypes = [ii, ii, dd, ipp, _dp, _dp]
        L.orc_action.argtypes = [ii, _dp]
        L.orc_action.restype = dd
        L.orc_cg_MdM_occupied.argtypes = [ii, ii, dd, ipp, _dp, _dp, ipp]
        L.orc_cg_MdM_occupied.restype = ii

    def set_boundary(self, bc):
        """Family B boundary variant (a compile-time choice in the reference, Thirring.h:27-29); applies to every
        family-B call of this process until changed."""
        self.lib.orc_set_boundary(int(bc))

    @staticmethod
    def _prep(v, A):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        A = np.ascontiguousarray(A, dtype=np.float64)
        nt, nx = v.shape
        assert A.shape == (nt, nx, 2)
        return v, A, nt, nx

    def fm_mul(self, v, A, m, mu):
        v, A, nt, nx = self._prep(v, A)
        out = np.empty_like(v)
        self.lib.orc_fm_mul(nt, nx, m, mu, _p(v), _p(out), _p(A))
        return out

    def fm_dagger_mul(self, v, A, m, mu):
        v, A, nt, nx = self._prep(v, A)
        out = np.empty_like(v)
        self.lib.orc_fm_dagger_mul(nt, nx, m, mu, _p(v), _p(out), _p(A))
        return out

    def fm_conjugate_mul(self, v, A, m, mu, mode):
        v, A, nt, nx = self._prep(v, A)
        out = np.empty_like(v)
        self.lib.orc_fm_conjugate_mul(nt, nx, m, mu, mode, _p(v), _p(out), _p(A))
        return out

    def fmdm_invert_cg(self, b, A, m, mu, mode, max_iter=0, treesum=False):
        """Returns (x, status, iterations, final rr).  treesum: NOT the reference's arithmetic; the loop's two dot
        products are summed as a pairwise tree, to measure what the summation order alone does to the iteration count."""
        b, A, nt, nx = self._prep(b, A)
        x = np.empty_like(b)
        it, rr = C.c_int(0), C.c_double(0)
        fn = self.lib.orc_fmdm_invert_cg_treesum if treesum else self.lib.orc_fmdm_invert_cg
        st = fn(nt, nx, m, mu, mode, _p(b), _p(x), _p(A), max_iter, C.byref(it), C.byref(rr))
        return x, st, it.value, rr.value

    def fm_invert_cg(self, v, A, m, mu, mode, max_iter=0):
        v, A, nt, nx = self._prep(v, A)
        x = np.empty_like(v)
        it, rr = C.c_int(0), C.c_double(0)
        st = self.lib.orc_fm_invert_cg(nt, nx, m, mu, mode, _p(v), _p(x), _p(A), max_iter,
                                       C.byref(it), C.byref(rr))
        return x, st, it.value, rr.value

    def fermion_matrix(self, A, m, mu):
        """Dense V x V matrix with Mdense[row, col]; row/col = NX*t + x (hmc.c:269-310)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        nt, nx, _ = A.shape
        V = nt * nx
        M = np.empty((V, V), dtype=np.complex128)  # column-major in the reference => transpose below
        self.lib.orc_fermion_matrix(nt, nx, m, mu, _p(A), _p(M))
        return M.T.copy()


    # ---- family B (vec_ops.c): real vectors (NT, NX) float64, field (NT, NX) int32 --------------------------
    @staticmethod
    def _prep_b(v, field):
        v = np.ascontiguousarray(v, dtype=np.float64)
        field = np.ascontiguousarray(field, dtype=np.int32)
        assert v.shape == field.shape
        return v, field, v.shape[0], v.shape[1]

    def fM(self, psi, field, m, mu, transpose=False):
        psi, field, nt, nx = self._prep_b(psi, field)
        chi = np.empty_like(psi)
        fn = self.lib.orc_fM_transpose if transpose else self.lib.orc_fM
        fn(nt, nx, m, mu, field.ctypes.data_as(C.POINTER(C.c_int)), _p(psi), _p(chi))
        return chi

    def cg_MdM(self, source, field, m, mu, propagator=False):
        source, field, nt, nx = self._prep_b(source, field)
        inv = np.empty_like(source)
        it, rr = C.c_int(0), C.c_double(0)
        fn = self.lib.orc_cg_propagator if propagator else self.lib.orc_cg_MdM
        st = fn(nt, nx, m, mu, field.ctypes.data_as(C.POINTER(C.c_int)), _p(source), _p(inv), C.byref(it), C.byref(rr))
        return inv, st, it.value, rr.value


    # flat-array family (vec_ops.c:345-461); vectors still passed as (NT, NX) arrays, flattened t*NX+x
    def fM_occupied(self, psi, field, mu, sq=False):
        psi, field, nt, nx = self._prep_b(psi, field)
        chi = np.empty_like(psi)
        fn = self.lib.orc_fM_occupied_sq if sq else self.lib.orc_fM_occupied
        fn(nt, nx, mu, field.ctypes.data_as(C.POINTER(C.c_int)), _p(psi), _p(chi))
        return chi

    def action(self, psi):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        return self.lib.orc_action(psi.size, _p(psi))

    def cg_MdM_occupied(self, source, field, mu):
        """Returns (psi, ret, iterations); ret 0 = converged, 1 = not (psi is then None)."""
        source, field, nt, nx = self._prep_b(source, field)
        psi = np.full_like(source, np.nan)
        it = C.c_int(0)
        ret = self.lib.orc_cg_MdM_occupied(nt, nx, mu, field.ctypes.data_as(C.POINTER(C.c_int)),
                                           _p(source), _p(psi), C.byref(it))
        return (psi if ret == 0 else None), ret, it.value


def ref_available(nt, nx, flavour="compat", nsteps=10):
    return os.path.exists(_ref_path(nt, nx, flavour, nsteps))


def _ref_path(nt, nx, flavour, nsteps=10):
    name = f"libhmcref_{nt}x{nx}_{flavour}"
    if nsteps != 10:
        name += f"_ns{nsteps}"
    return os.path.join(REF_DIR, name + ".so")


class RefLib:
    """The reference's own hmc.c, compiled unmodified (see oracle/build_ref.sh).

    hmc.c keeps exp(+-mu) in function-local statics that freeze at the first call (hmc.c:124-130), so
    every RefLib instance dlopens a private temporary copy of the shared object: one instance == one
    fresh "process image" of the reference with its own (m, g, mu) and RNG stream.
    """

    def __init__(self, nt, nx, flavour="compat", m=1.0, g=1.0, mu=0.0, seed=None, nsteps=10):
        src = _ref_path(nt, nx, flavour, nsteps)
        if not os.path.exists(src):
            raise FileNotFoundError(src + " (run oracle/build_ref.sh where /root/reference exists)")
        fd, self._tmp = tempfile.mkstemp(prefix="hmcref_", suffix=".so")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        self.lib = L = C.CDLL(self._tmp)
        os.unlink(self._tmp)
        self.nt, self.nx, self.flavour = nt, nx, flavour
        self.mode = MODE_ADJOINT if flavour.startswith("adjoint") else MODE_REF_COMPAT
        for name, val in (("m", m), ("g", g), ("mu", mu)):
            C.c_double.in_dll(L, name).value = val
        self.m, self.g, self.mu = m, g, mu
        # neighbour tables and eta exactly as main() fills them (hmc.c:899-925)
        self._tup = np.array([(i + 1) % nt for i in range(nt)], dtype=np.int32)
        self._tdn = np.array([(i - 1 + nt) % nt for i in range(nt)], dtype=np.int32)
        self._xup = np.array([(i + 1) % nx for i in range(nx + 1)], dtype=np.int32)
        self._xdn = np.array([(i - 1 + nx) % nx for i in range(nx + 1)], dtype=np.int32)
        for name, arr in (("tup", self._tup), ("tdn", self._tdn), ("xup", self._xup), ("xdn", self._xdn)):
            C.c_void_p.in_dll(L, name).value = arr.ctypes.data
        eta = np.zeros((nt, nx + 1, 2), dtype=np.int32)
        eta[:, :nx, 1] = 1
        eta[:, 0:nx:2, 0] = 1
        eta[:, 1:nx:2, 0] = -1
        self._eta, self._eta_rows, self._eta_top = self._triple(eta)
        C.c_void_p.in_dll(L, "eta").value = self._eta_top.ctypes.data
        self._keep = []
        vp = C.c_void_p
        for f in ("fm_mul", "fm_conjugate_mul", "fmdm_invert_cg", "fm_invert_cg"):
            getattr(L, f).argtypes = [vp, vp, vp]
            getattr(L, f).restype = None
        L.fermion_matrix.argtypes = [vp]
        L.update_puregauge_hb.argtypes = [vp]
        L.stochastic_vector.argtypes = [vp]
        L.random_pseudofermion.argtypes = [vp, vp]
        L.random_pseudofermion.restype = C.c_double
        L.pseudofermion_action.argtypes = [vp, vp]
        L.pseudofermion_action.restype = C.c_double
        L.seed_mersenne.argtypes = [C.c_long]
        L.mersenne_generate.restype = C.c_double
        if seed is not None:
            self.seed(seed)

    # --- marshalling -------------------------------------------------------------------------
    @staticmethod
    def _triple(arr):
        """arr[t][x][:] contiguous -> (arr, per-row pointer tables, top table) mimicking T***."""
        nt, nxs, k = arr.shape
        base = arr.ctypes.data
        item = arr.itemsize * k
        rows = np.empty((nt, nxs), dtype=np.uint64)
        for t in range(nt):
            rows[t] = base + (t * nxs + np.arange(nxs, dtype=np.uint64)) * item
        top = rows.ctypes.data + np.arange(nt, dtype=np.uint64) * (nxs * 8)
        return arr, rows, np.ascontiguousarray(top, dtype=np.uint64)

    def _vec(self, v):
        """complex128 (NT,NX) -> row-pointer table as alloc_vector() lays it out (hmc.c:105-112)."""
        assert v.dtype == np.complex128 and v.flags.c_contiguous and v.shape == (self.nt, self.nx)
        rows = v.ctypes.data + np.arange(self.nt, dtype=np.uint64) * (self.nx * 16)
        return np.ascontiguousarray(rows, dtype=np.uint64)

    def gauge(self, A=None):
        """Wrap (or allocate, zero-filled) a gauge field as double*** with NX+1 x-slots (hmc.c:889-897)."""
        full = np.zeros((self.nt, self.nx + 1, 2), dtype=np.float64)
        if A is not None:
            full[:, : self.nx, :] = A
        return _Gauge(*self._triple(full), self.nx)

    # --- RNG -----------------------------------------------------------------------------------
    def seed(self, seed, warmup=543210):
        """seed_mersenne + the 543210 warm-up draws of main() (hmc.c:874-877)."""
        self.lib.seed_mersenne(seed)
        for _ in range(warmup):
            self.mersenne()

    def mersenne(self):
        """The mersenne() macro of mersenne.h:11."""
        i = C.c_int.in_dll(self.lib, "mersenne_i")
        if i.value > 0:
            i.value -= 1
            return (C.c_double * 624).in_dll(self.lib, "mersenne_array")[i.value]
        return self.lib.mersenne_generate()

    # --- reference functions ---------------------------------------------------------------------
    def heatbath(self, G, sweeps=1):
        for _ in range(sweeps):
            self.lib.update_puregauge_hb(G.top.ctypes.data)

    def stochastic_vector(self):
        v = np.empty((self.nt, self.nx), dtype=np.complex128)
        self.lib.stochastic_vector(self._vec(v).ctypes.data)
        return v

    def _apply(self, fname, v, G):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.empty_like(v)
        vi, vo = self._vec(v), self._vec(out)
        getattr(self.lib, fname)(vi.ctypes.data, vo.ctypes.data, G.top.ctypes.data)
        return out

    def fm_mul(self, v, G):
        return self._apply("fm_mul", v, G)

    def fm_conjugate_mul(self, v, G):
        return self._apply("fm_conjugate_mul", v, G)

    def fmdm_invert_cg(self, b, G):
        return self._apply("fmdm_invert_cg", b, G)

    def fm_invert_cg(self, b, G):
        return self._apply("fm_invert_cg", b, G)

    def fermion_matrix(self, G):
        """Dense matrix (hmc.c:269-310 reads the GLOBAL A) returned as Mdense[row, col]."""
        C.c_void_p.in_dll(self.lib, "A").value = G.top.ctypes.data
        V = self.nt * self.nx
        M = np.empty((V, V), dtype=np.complex128)
        self.lib.fermion_matrix(M.ctypes.data)
        return M.T.copy()


class _Gauge:
    def __init__(self, arr, rows, top, nx):
        self.arr, self.rows, self.top, self.nx = arr, rows, top, nx

    @property
    def A(self):
        """(NT, NX, 2) float64 view of the physical links."""
        return self.arr[:, : self.nx, :]


def ref_b_available(nt, nx, bc=BC_ANTISYMMETRIC):
    return os.path.exists(os.path.join(REF_DIR, f"libvecopsref_{nt}x{nx}" + ("_symmetric" if bc == BC_SYMMETRIC else "") + ".so"))


class RefLibB:
    """The reference's own vec_ops.c (family B), compiled unmodified with the EXTERN globals of Thirring.h
    defined in the same object (oracle/build_ref.sh).  A private RTLD_DEEPBIND copy: its internal calls can
    never be interposed, so it stays a pure CPU checker even when the GPU shim is loaded globally."""

    def __init__(self, nt, nx, m=1.0, mu=0.0, deepbind=True, bc=BC_ANTISYMMETRIC):
        # SYMMETRIC is another build (the #define of Thirring.h:27-28 swapped); OPENX is the ANTISYMMETRIC object with
        # the neighbour tables and the phantom column the driver sets up for it (fermionbag.c:713-717,761-765)
        src = os.path.join(REF_DIR, f"libvecopsref_{nt}x{nx}" + ("_symmetric" if bc == BC_SYMMETRIC else "") + ".so")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        fd, tmp = tempfile.mkstemp(prefix="vecopsref_", suffix=".so")
        os.close(fd)
        shutil.copyfile(src, tmp)
        mode = (os.RTLD_LOCAL | os.RTLD_DEEPBIND | os.RTLD_NOW) if deepbind else (os.RTLD_GLOBAL | os.RTLD_NOW)
        self.lib = L = C.CDLL(tmp, mode=mode)
        os.unlink(tmp)
        self.nt, self.nx = nt, nx
        C.c_double.in_dll(L, "m").value = m
        C.c_double.in_dll(L, "mu").value = mu
        self.m, self.mu = m, mu
        self._tup = np.array([(i + 1) % nt for i in range(nt)], dtype=np.int32)
        self._tdn = np.array([(i - 1 + nt) % nt for i in range(nt)], dtype=np.int32)
        self._xup = np.array([(i + 1) % nx for i in range(nx + 1)], dtype=np.int32)
        self._xdn = np.array([(i - 1 + nx) % nx for i in range(nx + 1)], dtype=np.int32)
        if bc == BC_OPENX:
            self._xdn[0] = nx
            self._xup[nx - 1] = nx
            self._xup[nx] = nx
        for name, arr in (("tup", self._tup), ("tdn", self._tdn), ("xup", self._xup), ("xdn", self._xdn)):
            C.c_void_p.in_dll(L, name).value = arr.ctypes.data
        eta = np.zeros((nt, nx + 1, 2), dtype=np.int32)   # fermionbag.c:770-781
        eta[:, :nx, 1] = 1
        eta[:, 0:nx:2, 0] = 1
        eta[:, 1:nx:2, 0] = -1
        self._eta, self._eta_rows, self._eta_top = RefLib._triple(eta)
        C.c_void_p.in_dll(L, "eta").value = self._eta_top.ctypes.data
        self.field = np.zeros((nt, nx + 1), dtype=np.int32)   # fermionbag.c:697,710-712
        if bc == BC_OPENX:
            self.field[:, nx] = -100   # EMPTY, Thirring.h:41
        self._field_rows = np.ascontiguousarray(
            self.field.ctypes.data + np.arange(nt, dtype=np.uint64) * ((nx + 1) * 4), dtype=np.uint64)
        C.c_void_p.in_dll(L, "field").value = self._field_rows.ctypes.data
        vp = C.c_void_p
        for f in ("fM", "fM_transpose", "cg_MdM", "cg_propagator"):
            getattr(L, f).argtypes = [vp, vp]
            getattr(L, f).restype = None

    def set_field(self, field):
        self.field[:, : self.nx] = field

    def _rows(self, v):
        assert v.dtype == np.float64 and v.flags.c_contiguous and v.shape == (self.nt, self.nx)
        return np.ascontiguousarray(v.ctypes.data + np.arange(self.nt, dtype=np.uint64) * (self.nx * 8), dtype=np.uint64)

    def call(self, fname, src):
        """All four functions take (out, in) (vec_ops.c:96,135,261,311)."""
        src = np.ascontiguousarray(src, dtype=np.float64)
        out = np.zeros_like(src)
        o, i = self._rows(out), self._rows(src)
        getattr(self.lib, fname)(o.ctypes.data, i.ctypes.data)
        return out

    def call_flat(self, fname, src):
        """The flat-array family (vec_ops.c:345-461): fM_occupied, fM_occupied_sq take (out, in) as double[NT*NX];
        cg_MdM_occupied returns (psi or None, ret)."""
        src = np.ascontiguousarray(src, dtype=np.float64)
        out = np.full_like(src, np.nan)
        fn = getattr(self.lib, fname)
        fn.argtypes = [C.c_void_p, C.c_void_p]
        fn.restype = C.c_int if fname == "cg_MdM_occupied" else None
        ret = fn(out.ctypes.data, src.ctypes.data)
        if fname == "cg_MdM_occupied":
            return (out if ret == 0 else None), ret
        return out

    def set_mu(self, mu):
        C.c_double.in_dll(self.lib, "mu").value = mu
        self.mu = mu
