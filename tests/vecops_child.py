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
    shim.tb_vecops_shutdown()
    print("vecops child ok")


if __name__ == "__main__":
    main()
