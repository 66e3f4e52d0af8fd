"""Child process of tests/test_gpu_family_b.py::test_vec_ops_replacement_object_under_reference_code: the replacement
object libthirring_vecops.so under the reference's own calling code (vec_ops.c), checked against the oracle."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Oracle, RefLibB  # noqa: E402  (the checker)
from tests.util import APPLY_TOL, CG_SOL_TOL, assert_close  # noqa: E402


def main():
    oracle = Oracle()
    nt = nx = 32
    shim = ctypes.CDLL(os.path.join(ROOT, "thirring2d_b200", "libthirring_vecops.so"), mode=os.RTLD_GLOBAL | os.RTLD_NOW)
    shim.tb_vecops_configure(nt, nx, 0)
    m, mu = 0.2, 0.1
    drv = RefLibB(nt, nx, m=m, mu=mu, deepbind=False)   # plays fermionbag.c: owns the globals, calls through the PLT
    rng = np.random.default_rng(8)
    field = (rng.random((nt, nx)) < 0.15).astype(np.int32)
    drv.set_field(field)
    psi = rng.normal(size=(nt, nx))
    before = shim.tb_vecops_gpu_calls()
    prop = drv.call("cg_propagator", psi)          # reference code, GPU hot path
    assert shim.tb_vecops_gpu_calls() - before >= 2  # fM_transpose + cg_MdM went to the GPU
    xo, st, it, rr = oracle.cg_MdM(psi, field, m, mu, propagator=True)
    assert_close(prop, xo, CG_SOL_TOL, "interposed cg_propagator")
    # the exported symbols called directly, (out, in) order
    out = np.zeros_like(psi)
    rows = lambda v: np.ascontiguousarray(v.ctypes.data + np.arange(nt, dtype=np.uint64) * (nx * 8), dtype=np.uint64)
    o, i = rows(out), rows(psi)
    shim.fM(ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(i.ctypes.data))
    assert_close(out, oracle.fM(psi, field, m, mu), APPLY_TOL, "shim fM")
    field[3, 4] = 1 - field[3, 4]                    # the driver changes the configuration between calls
    drv.set_field(field)
    shim.fM_transpose(ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(i.ctypes.data))
    assert_close(out, oracle.fM(psi, field, m, mu, transpose=True), APPLY_TOL, "shim fM_transpose after update")
    flat_array_family(shim, drv, oracle, rng, nt, nx, m)
    shim.tb_vecops_shutdown()
    print("vecops child ok")


def flat_array_family(shim, drv, oracle, rng, nt, nx, m):
    """vec_ops.c:345-461 through the replacement object: fM_occupied / fM_occupied_sq / cg_MdM_occupied on the GPU
    (a second, massless context), action and vec_gaussian on the host with the driver's Mersenne state."""
    vp, dbl = ctypes.c_void_p, ctypes.c_double
    shim.action.restype = dbl
    shim.alloc_field.restype = vp
    chk = RefLibB(nt, nx, m=m, mu=0.0)                       # private CPU copy of the reference: the checker
    for mu, ndimers in ((0.1, 0), (0.0, 0), (0.0, 10)):
        field = np.zeros((nt, nx), dtype=np.int32)
        for _ in range(ndimers):
            t, x = rng.integers(nt), rng.integers(nx - 1)
            field[t, x] = field[t, x + 1] = 1
        if mu:
            field = (rng.random((nt, nx)) < 0.15).astype(np.int32)
        drv.set_field(field)
        drv.set_mu(mu)
        chk.set_field(field)
        chk.set_mu(mu)
        psi = rng.normal(size=(nt, nx))
        out = np.full_like(psi, np.nan)
        before = shim.tb_vecops_gpu_calls()
        shim.fM_occupied(vp(out.ctypes.data), vp(psi.ctypes.data))
        assert_close(out, chk.call_flat("fM_occupied", psi), APPLY_TOL, "fM_occupied")
        assert np.array_equal(out[field != 0], psi[field != 0])
        shim.fM_occupied_sq(vp(out.ctypes.data), vp(psi.ctypes.data))
        assert_close(out, chk.call_flat("fM_occupied_sq", psi), APPLY_TOL, "fM_occupied_sq")
        assert shim.tb_vecops_gpu_calls() - before == 2
        assert shim.action(vp(psi.ctypes.data)) == oracle.action(psi)
        if mu == 0.0:
            src = np.where(field == 0, psi, 0.0)
            want, ret_ref, it = oracle.cg_MdM_occupied(src, field, mu)
            assert ret_ref == 0
            got = np.full_like(src, np.nan)
            ret = shim.cg_MdM_occupied(vp(got.ctypes.data), vp(src.ctypes.data))
            assert ret == 0
            assert_close(got, want, CG_SOL_TOL, "cg_MdM_occupied")
    # zero source: the reference divides 0/0 and bails out with 1, psi untouched
    got = np.full((nt, nx), 7.0)
    zero = np.zeros((nt, nx))
    assert shim.cg_MdM_occupied(vp(got.ctypes.data), vp(zero.ctypes.data)) == 1 and np.all(got == 7.0)
    # vec_gaussian draws from the DRIVER's generator: same state, same numbers as the reference's own routine
    drv.lib.seed_mersenne.argtypes = [ctypes.c_long]
    chk.lib.seed_mersenne.argtypes = [ctypes.c_long]
    drv.lib.seed_mersenne(4354365264)
    chk.lib.seed_mersenne(4354365264)
    a = np.zeros((nt, nx))
    b = np.zeros((nt, nx))
    shim.vec_gaussian(vp(a.ctypes.data))
    chk.lib.vec_gaussian(vp(b.ctypes.data))
    assert np.array_equal(a, b) and abs(a.std() - 1) < 0.1
    fld = shim.alloc_field()
    assert fld
    ctypes.CDLL(None).free(vp(fld))


if __name__ == "__main__":
    main()
