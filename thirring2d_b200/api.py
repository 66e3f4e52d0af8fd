"""Host-side mirror of the reference interface for the hot path (hmc.c:123-414), batched over chains.

``Context`` owns one tb_ctx.  Methods carry the reference's names and argument meaning:

    fm_mul(v)            hmc.c:123   M v
    fm_conjugate_mul(v)  hmc.c:188   M~ v  (== M v in REF_COMPAT, M^dagger v in ADJOINT)
    fmdm_invert_cg(b)    hmc.c:341   (M~ M)^-1 b by CG from x0 = 0, absolute stop ||r||^2 < 1e-30
    fm_invert_cg(v)      hmc.c:408   (M~ M)^-1 M~ v

Host arrays: vectors complex128 (nchains, NT, NX) (a single (NT, NX) array is accepted for nchains == 1),
gauge fields float64 (nchains, NT, NX, 2).  The ``*_dev`` methods take raw device pointers (e.g.
``torch.Tensor.data_ptr()``) in the library's device layout and move no data across PCIe.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .lib import check, load_library

MODE_REF_COMPAT, MODE_ADJOINT = 0, 1
OP_M, OP_MDAG, OP_MCONJ, OP_MDM = 0, 1, 2, 3
CG_CONVERGED, CG_MAXITER, CG_DIVERGED, CG_ZERO_SOURCE = 0, 1, 2, 3
BC_ANTISYMMETRIC, BC_SYMMETRIC, BC_OPENX = 0, 1, 2   # family B boundary variants, Thirring.h:27-29

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class CGInfo:
    """Per-chain outcome of a batched solve."""

    def __init__(self, status, iters, rr):
        self.status, self.iters, self.rr = status, iters, rr

    def __repr__(self):
        return f"CGInfo(status={self.status.tolist()}, iters={self.iters.tolist()})"


class Context:
    def __init__(self, nt, nx, nchains=1, mode=MODE_ADJOINT, device=0, m=1.0, mu=0.0, stream=None,
                 slab_rank=None, slab_nranks=1):
        """nt is the GLOBAL number of t-rows; with slab_nranks > 1 this rank owns nt // slab_nranks of them
        and every array passed to the methods is the local slab (nchains, nt_local, nx[, 2])."""
        self.lib = load_library()
        self.nt_global, self.nranks, self.rank = nt, slab_nranks, slab_rank or 0
        h = C.c_void_p()
        if slab_nranks > 1:
            check(self.lib.tb_create_slab(C.byref(h), nt, nx, nchains, mode, device, self.rank, slab_nranks),
                  "tb_create_slab")
            nt = nt // slab_nranks
        else:
            check(self.lib.tb_create(C.byref(h), nt, nx, nchains, mode, device), "tb_create")
        self.nt, self.nx, self.nchains, self.mode, self.device = nt, nx, nchains, mode, device
        self._h = h
        if stream is not None:
            self.set_stream(stream)
        self.set_params(m, mu)

    # -- life cycle ----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.tb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- slab decomposition (one process per GPU) ---------------------------------------------------------
    def slab_export(self) -> bytes:
        buf = C.create_string_buffer(self.lib.tb_slab_handle_bytes())
        check(self.lib.tb_slab_export(self._h, buf), "tb_slab_export")
        return buf.raw

    def slab_connect(self, handles):
        """handles: list of the nranks exported handles in rank order."""
        blob = b"".join(handles)
        check(self.lib.tb_slab_connect(self._h, blob), "tb_slab_connect")

    def slab_setup(self, dist):
        """Exchange the IPC handles over an initialised torch.distributed group and connect."""
        handles = [None] * self.nranks
        dist.all_gather_object(handles, self.slab_export())
        self.slab_connect(handles)
        dist.barrier()

    # -- configuration ---------------------------------------------------------------------------------
    def set_stream(self, stream_ptr: int):
        check(self.lib.tb_set_stream(self._h, C.c_void_p(stream_ptr)), "tb_set_stream")

    def synchronize(self):
        check(self.lib.tb_synchronize(self._h), "tb_synchronize")

    def set_params(self, m, mu):
        m = np.atleast_1d(np.asarray(m, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        n = max(m.size, mu.size)
        if n > 1:
            m = np.ascontiguousarray(np.broadcast_to(m, (n,)))
            mu = np.ascontiguousarray(np.broadcast_to(mu, (n,)))
        check(self.lib.tb_set_params(self._h, m.ctypes.data_as(_dp), mu.ctypes.data_as(_dp), n), "tb_set_params")

    def set_cg(self, accuracy=1e-30, max_iter=100000):
        check(self.lib.tb_set_cg(self._h, accuracy, max_iter), "tb_set_cg")

    def set_tuning(self, rows_per_thread=0, iters_per_launch=0, solver=0):
        check(self.lib.tb_set_tuning(self._h, rows_per_thread, iters_per_launch, solver), "tb_set_tuning")

    def solver_info(self):
        """(kind, chains_in_flight): kind 0 streaming, 1 on-chip CTA per chain, 2 on-chip cluster per chain."""
        k, n = C.c_int(0), C.c_int(0)
        check(self.lib.tb_solver_info(self._h, C.byref(k), C.byref(n)), "tb_solver_info")
        return k.value, n.value

    def streaming_info(self):
        """(kernels, tile_chains, tile_sites, rows_per_block) of the streaming solver: kernels 0 register-marching,
        1 TMA-staged whole-batch tiles, 2 TMA-staged 16-chain tiles through tensor maps."""
        v = [C.c_int(0) for _ in range(4)]
        check(self.lib.tb_streaming_info(self._h, *[C.byref(q) for q in v]), "tb_streaming_info")
        return tuple(q.value for q in v)

    # -- host-buffer path (reference-facing) -----------------------------------------------------------------
    def _vec(self, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        if v.ndim == 2:
            v = v[None]
        assert v.shape == (self.nchains, self.nt, self.nx), (v.shape, (self.nchains, self.nt, self.nx))
        return v

    def set_gauge(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        if A.ndim == 3:
            A = A[None]
        assert A.shape == (self.nchains, self.nt, self.nx, 2), A.shape
        check(self.lib.tb_set_gauge(self._h, A.ctypes.data), "tb_set_gauge")

    def set_gauge_shared(self, A):
        """One gauge field (NT, NX, 2) for every chain of the context: the chains are the right-hand sides of a multi-RHS
        solve on that field (fermion_phase's sources, hmc.c:794-815)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        assert A.shape == (self.nt, self.nx, 2), A.shape
        check(self.lib.tb_set_gauge_shared(self._h, A.ctypes.data), "tb_set_gauge_shared")

    def set_links_trig(self, trig_t, trig_x):
        """Links from the caller's cos / sin: float64 (nchains, NT, NX, 2) = (cos A, sin A) per direction."""
        tt = np.ascontiguousarray(trig_t, dtype=np.float64)
        tx = np.ascontiguousarray(trig_x, dtype=np.float64)
        if tt.ndim == 3:
            tt, tx = tt[None], tx[None]
        assert tt.shape == tx.shape == (self.nchains, self.nt, self.nx, 2), tt.shape
        check(self.lib.tb_set_links_trig(self._h, tt.ctypes.data, tx.ctypes.data), "tb_set_links_trig")

    def set_occupancy(self, field, bc=BC_ANTISYMMETRIC):
        """Family B (vec_ops.c): occupation field, ints, 0 = free.  (nchains, NT, NX): one field per source; (NT, NX)
        with nchains > 1: ONE field shared by every source of the batch (multi-RHS).  bc: boundary variant."""
        field = np.ascontiguousarray(field, dtype=np.int32)
        shared = field.ndim == 2 and self.nchains > 1
        if field.ndim == 2 and not shared:
            field = field[None]
        assert field.shape == ((self.nt, self.nx) if shared else (self.nchains, self.nt, self.nx)), field.shape
        check(self.lib.tb_set_occupancy_bc(self._h, field.ctypes.data, int(bc), 1 if shared else 0), "tb_set_occupancy_bc")

    # family B names (vec_ops.c:96,135,261,311): real vectors in, real vectors out, 8 bytes per site end to end
    def _real(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        squeeze = v.ndim == 2
        if squeeze:
            v = v[None]
        assert v.shape == (self.nchains, self.nt, self.nx), (v.shape, (self.nchains, self.nt, self.nx))
        return v, squeeze

    def _apply_real(self, op, psi):
        psi, squeeze = self._real(psi)
        out = np.empty_like(psi)
        check(self.lib.tb_apply_real(self._h, op, psi.ctypes.data, out.ctypes.data), "tb_apply_real")
        return out[0] if squeeze else out

    def fM(self, psi):
        return self._apply_real(OP_M, psi)

    def fM_transpose(self, psi):
        return self._apply_real(OP_MDAG, psi)

    def _solve_real(self, source, propagator):
        b, squeeze = self._real(source)
        x = np.empty_like(b)
        st = np.empty(self.nchains, dtype=np.int32)
        it = np.empty(self.nchains, dtype=np.int32)
        rr = np.empty(self.nchains, dtype=np.float64)
        check(self.lib.tb_cg_real(self._h, b.ctypes.data, x.ctypes.data, 1 if propagator else 0, st.ctypes.data_as(_ip),
                                  it.ctypes.data_as(_ip), rr.ctypes.data_as(_dp)), "tb_cg_real")
        return (x[0] if squeeze else x), CGInfo(st, it, rr)

    def cg_MdM(self, source):
        return self._solve_real(source, False)

    def cg_propagator(self, source):
        return self._solve_real(source, True)

    def cg_real_host_ptr(self, b_ptr: int, x_ptr: int, propagator=True):
        """Raw host pointers (e.g. pinned buffers) of real batches double[nchains][NT][NX]."""
        check(self.lib.tb_cg_real(self._h, C.c_void_p(b_ptr), C.c_void_p(x_ptr), 1 if propagator else 0, None, None, None),
              "tb_cg_real")

    def vec_dot_dev(self, d_a: int, d_b: int):
        """vec_dot (vec_ops.c:56) for every vector of a device-resident real batch."""
        out = np.empty(self.nchains, dtype=np.float64)
        check(self.lib.tb_vec_dot_real_dev(self._h, C.c_void_p(d_a), C.c_void_p(d_b), out.ctypes.data_as(_dp)),
              "tb_vec_dot_real_dev")
        return out

    def vec_dmul_add_dev(self, d_a: int, d_b: int, d_d: int, e):
        """vec_dmul_add (vec_ops.c:51): a = b + e[chain] * d on device-resident real batches."""
        e = np.ascontiguousarray(np.broadcast_to(np.asarray(e, dtype=np.float64), (self.nchains,)))
        check(self.lib.tb_vec_dmul_add_real_dev(self._h, C.c_void_p(d_a), C.c_void_p(d_b), C.c_void_p(d_d),
                                                e.ctypes.data_as(_dp)), "tb_vec_dmul_add_real_dev")

    def apply(self, op, v):
        squeeze = np.ndim(v) == 2
        v = self._vec(v)
        out = np.empty_like(v)
        check(self.lib.tb_apply(self._h, op, v.ctypes.data, out.ctypes.data), "tb_apply")
        return out[0] if squeeze else out

    def fm_mul(self, v):
        return self.apply(OP_M, v)

    def fm_conjugate_mul(self, v):
        return self.apply(OP_MCONJ, v)

    def fm_dagger_mul(self, v):
        return self.apply(OP_MDAG, v)

    def fmdm_mul(self, v):
        return self.apply(OP_MDM, v)

    def _solve(self, fn, name, b):
        squeeze = np.ndim(b) == 2
        b = self._vec(b)
        x = np.empty_like(b)
        st = np.empty(self.nchains, dtype=np.int32)
        it = np.empty(self.nchains, dtype=np.int32)
        rr = np.empty(self.nchains, dtype=np.float64)
        check(fn(self._h, b.ctypes.data, x.ctypes.data, st.ctypes.data_as(_ip), it.ctypes.data_as(_ip),
                 rr.ctypes.data_as(_dp)), name)
        return (x[0] if squeeze else x), CGInfo(st, it, rr)

    def fmdm_invert_cg(self, b):
        return self._solve(self.lib.tb_cg, "tb_cg", b)

    def fm_invert_cg(self, v):
        return self._solve(self.lib.tb_invert, "tb_invert", v)

    # raw host pointers (e.g. pinned torch tensors) — used by bench.py's e2e leg
    def cg_host_ptr(self, b_ptr: int, x_ptr: int):
        check(self.lib.tb_cg(self._h, C.c_void_p(b_ptr), C.c_void_p(x_ptr), None, None, None), "tb_cg")

    def cg_gauge_host_ptr(self, a_ptr: int, b_ptr: int, x_ptr: int):
        check(self.lib.tb_cg_gauge(self._h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(x_ptr), None, None, None),
              "tb_cg_gauge")

    def fmdm_invert_cg_with_gauge(self, A, b):
        """set_gauge(A) followed by fmdm_invert_cg(b), uploads interleaved per sub-batch of chains."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        if A.ndim == 3:
            A = A[None]
        assert A.shape == (self.nchains, self.nt, self.nx, 2), A.shape
        squeeze = np.ndim(b) == 2
        b = self._vec(b)
        x = np.empty_like(b)
        st = np.empty(self.nchains, dtype=np.int32)
        it = np.empty(self.nchains, dtype=np.int32)
        rr = np.empty(self.nchains, dtype=np.float64)
        check(self.lib.tb_cg_gauge(self._h, A.ctypes.data, b.ctypes.data, x.ctypes.data, st.ctypes.data_as(_ip),
                                   it.ctypes.data_as(_ip), rr.ctypes.data_as(_dp)), "tb_cg_gauge")
        return (x[0] if squeeze else x), CGInfo(st, it, rr)

    def set_gauge_host_ptr(self, a_ptr: int):
        check(self.lib.tb_set_gauge(self._h, C.c_void_p(a_ptr)), "tb_set_gauge")

    # -- device-resident path ------------------------------------------------------------------------
    @property
    def vec_doubles(self):
        return int(self.lib.tb_vec_doubles(self._h))

    def pack_dev(self, d_canonical: int, d_vec: int):
        check(self.lib.tb_pack_dev(self._h, C.c_void_p(d_canonical), C.c_void_p(d_vec)), "tb_pack_dev")

    def unpack_dev(self, d_vec: int, d_canonical: int):
        check(self.lib.tb_unpack_dev(self._h, C.c_void_p(d_vec), C.c_void_p(d_canonical)), "tb_unpack_dev")

    def set_gauge_dev(self, d_A_canonical: int):
        check(self.lib.tb_set_gauge_dev(self._h, C.c_void_p(d_A_canonical)), "tb_set_gauge_dev")

    def set_gauge_shared_dev(self, d_A_one_field: int):
        check(self.lib.tb_set_gauge_shared_dev(self._h, C.c_void_p(d_A_one_field)), "tb_set_gauge_shared_dev")

    def apply_dev(self, op, d_in: int, d_out: int):
        check(self.lib.tb_apply_dev(self._h, op, C.c_void_p(d_in), C.c_void_p(d_out)), "tb_apply_dev")

    def cg_dev(self, d_b: int, d_x: int):
        check(self.lib.tb_cg_dev(self._h, C.c_void_p(d_b), C.c_void_p(d_x)), "tb_cg_dev")

    def invert_dev(self, d_v: int, d_x: int):
        check(self.lib.tb_invert_dev(self._h, C.c_void_p(d_v), C.c_void_p(d_x)), "tb_invert_dev")

    def cg_result(self):
        st = np.empty(self.nchains, dtype=np.int32)
        it = np.empty(self.nchains, dtype=np.int32)
        rr = np.empty(self.nchains, dtype=np.float64)
        check(self.lib.tb_cg_result(self._h, st.ctypes.data_as(_ip), it.ctypes.data_as(_ip),
                                    rr.ctypes.data_as(_dp)), "tb_cg_result")
        return CGInfo(st, it, rr)

    def re_dot_dev(self, d_a: int, d_b: int):
        out = np.empty(self.nchains, dtype=np.float64)
        check(self.lib.tb_re_dot_dev(self._h, C.c_void_p(d_a), C.c_void_p(d_b), out.ctypes.data_as(_dp)),
              "tb_re_dot_dev")
        return out

    # -- device-resident HMC trajectory (hmc.c:671-746, batched) ------------------------------------------
    def hmc_set_coupling(self, g):
        g = np.ascontiguousarray(np.atleast_1d(np.asarray(g, dtype=np.float64)))
        check(self.lib.tb_hmc_set_coupling(self._h, g.ctypes.data_as(_dp), g.size), "tb_hmc_set_coupling")

    def hmc_set_chain_offset(self, first_chain: int):
        """Global index of this context's first chain (enters the device RNG key; see shard.chain_range)."""
        check(self.lib.tb_hmc_set_chain_offset(self._h, int(first_chain)), "tb_hmc_set_chain_offset")

    def hmc_heatbath(self, sweeps=100, seed=1):
        check(self.lib.tb_hmc_heatbath(self._h, sweeps, seed), "tb_hmc_heatbath")

    def get_gauge(self):
        A = np.empty((self.nchains, self.nt, self.nx, 2), dtype=np.float64)
        check(self.lib.tb_get_gauge(self._h, A.ctypes.data), "tb_get_gauge")
        return A

    def hmc_trajectory(self, nsteps=10, traj_length=1.0, seed=1, traj_index=0, xi=None, mom=None, st=None, u=None):
        """update_gauge for every chain.  Returns (obs, accepted, cg_iterations): obs[c] = Sg, Smdm, Smd, Smom,
        Sg', Smdm', Smd', Smom', dS, accepted.  xi/mom/st/u: host random inputs for parity tests (else Philox)."""
        keep = []

        def ptr(arr, dtype, shape):
            if arr is None:
                return None
            arr = np.ascontiguousarray(arr, dtype=dtype)
            assert arr.shape == shape, (arr.shape, shape)
            keep.append(arr)
            return arr.ctypes.data

        vs = (self.nchains, self.nt, self.nx)
        obs = np.empty((self.nchains, 10), dtype=np.float64)
        acc = np.empty(self.nchains, dtype=np.int32)
        its = C.c_longlong(0)
        check(self.lib.tb_hmc_trajectory(self._h, nsteps, traj_length, seed, traj_index,
                                         ptr(xi, np.complex128, vs), ptr(mom, np.float64, vs + (2,)),
                                         ptr(st, np.complex128, vs), ptr(u, np.float64, (self.nchains,)),
                                         obs.ctypes.data, acc.ctypes.data_as(_ip), C.byref(its)),
              "tb_hmc_trajectory")
        return obs, acc, its.value

    def hmc_force(self, psi, st=None):
        """dS/dA of momentum_step (hmc.c:504-661) for every link: real (nchains, NT, NX, 2)."""
        psi = self._vec(psi)
        stp = None
        if st is not None:
            st = self._vec(st)
            stp = st.ctypes.data
        out = np.empty((self.nchains, self.nt, self.nx, 2), dtype=np.float64)
        check(self.lib.tb_hmc_force(self._h, psi.ctypes.data, stp, out.ctypes.data), "tb_hmc_force")
        return out

    def hmc_cg_failures(self):
        """Per chain: bit (1 << CG_MAXITER) / (1 << CG_DIVERGED) set if a solve of the last trajectory ended so."""
        mask = np.empty(self.nchains, dtype=np.int32)
        check(self.lib.tb_hmc_cg_failures(self._h, mask.ctypes.data_as(_ip)), "tb_hmc_cg_failures")
        return mask

    def hmc_measure(self, nsrc=20, seed=1, meas_index=0, sources=None):
        mag = np.empty(self.nchains, dtype=np.float64)
        ph = np.empty(self.nchains, dtype=np.float64)
        src = None
        if sources is not None:
            sources = np.ascontiguousarray(sources, dtype=np.complex128)
            assert sources.shape == (nsrc, self.nchains, self.nt, self.nx)
            src = sources.ctypes.data
        check(self.lib.tb_hmc_measure(self._h, nsrc, seed, meas_index, src, mag.ctypes.data_as(_dp),
                                      ph.ctypes.data_as(_dp)), "tb_hmc_measure")
        return mag, ph

    def hmc_condensate(self, nsrc=20, seed=1, meas_index=0, sources=None):
        """Per-chain stochastic estimate of (1/V) Tr M^-1 through fm_invert_cg.  Returns (condensate, cg_iterations)."""
        cond = np.empty(self.nchains, dtype=np.float64)
        src = None
        if sources is not None:
            sources = np.ascontiguousarray(sources, dtype=np.complex128)
            assert sources.shape == (nsrc, self.nchains, self.nt, self.nx)
            src = sources.ctypes.data
        its = C.c_longlong(0)
        check(self.lib.tb_hmc_condensate(self._h, nsrc, seed, meas_index, src, cond.ctypes.data_as(_dp),
                                         C.byref(its)), "tb_hmc_condensate")
        return cond, its.value

    def checkpoint_write(self, path, next_trajectory=None):
        """next_trajectory: index the next trajectory will use (keys the device random stream); stored in the header."""
        if next_trajectory is not None:
            check(self.lib.tb_checkpoint_set_next_trajectory(self._h, int(next_trajectory)),
                  "tb_checkpoint_set_next_trajectory")
        check(self.lib.tb_checkpoint_write(self._h, os.fsencode(path)), "tb_checkpoint_write")

    def checkpoint_read(self, path):
        """Returns the index of the next trajectory recorded in the file (0: not recorded)."""
        check(self.lib.tb_checkpoint_read(self._h, os.fsencode(path)), "tb_checkpoint_read")
        n = C.c_uint(0)
        check(self.lib.tb_checkpoint_next_trajectory(self._h, C.byref(n)), "tb_checkpoint_next_trajectory")
        return n.value

    @property
    def launch_count(self):
        return int(self.lib.tb_launch_count(self._h))

    def reset_launch_count(self):
        self.lib.tb_reset_launch_count(self._h)

    def measure_fp64_peak(self, repeats=5):
        """FP64 FMA rate of the device in TFLOP/s (independent DFMA chains): the on-chip solvers' roofline denominator."""
        tf = C.c_double(0.0)
        check(self.lib.tb_measure_fp64_peak(self._h, repeats, C.byref(tf)), "tb_measure_fp64_peak")
        return tf.value

    def measure_fp64_rate(self, kind, repeats=5):
        """FP64 FMA rate in TFLOP/s for an instruction mix: kind 0 = measure_fp64_peak, 1 = every source operand of every
        FMA in its own register (a stencil's FMAs)."""
        tf = C.c_double(0.0)
        check(self.lib.tb_measure_fp64_rate(self._h, kind, repeats, C.byref(tf)), "tb_measure_fp64_rate")
        return tf.value

    @property
    def last_solve_ms(self):
        return float(self.lib.tb_last_solve_ms(self._h))
