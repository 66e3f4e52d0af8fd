"""ctypes loader for libthirring_b200.so.  Fails loudly when the CUDA extension is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class TBError(RuntimeError):
    pass


def library_path() -> str:
    return os.path.join(_HERE, "libthirring_b200.so")


# name -> (restype, argtypes); every symbol include/thirring_b200.h declares
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "tb_last_error": (C.c_char_p, []),
    "tb_device_count": (_i, []),
    "tb_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "tb_destroy": (_i, [_vp]),
    "tb_set_stream": (_i, [_vp, _vp]),
    "tb_synchronize": (_i, [_vp]),
    "tb_set_params": (_i, [_vp, _dp, _dp, _i]),
    "tb_set_cg": (_i, [_vp, _d, _i]),
    "tb_set_tuning": (_i, [_vp, _i, _i, _i]),
    "tb_solver_info": (_i, [_vp, _ip, _ip]),
    "tb_streaming_info": (_i, [_vp, _ip, _ip, _ip, _ip]),
    "tb_plan_schedule": (_i, [_ip, _ip, _i, _i, _ip, _ip, _ip]),
    "tb_set_gauge": (_i, [_vp, _vp]),
    "tb_set_links_trig": (_i, [_vp, _vp, _vp]),
    "tb_set_gauge_shared": (_i, [_vp, _vp]),
    "tb_set_gauge_shared_dev": (_i, [_vp, _vp]),
    "tb_set_occupancy": (_i, [_vp, _vp]),
    "tb_set_occupancy_bc": (_i, [_vp, _vp, _i, _i]),
    "tb_apply_real": (_i, [_vp, _i, _vp, _vp]),
    "tb_cg_real": (_i, [_vp, _vp, _vp, _i, _ip, _ip, _dp]),
    "tb_vec_dot_real_dev": (_i, [_vp, _vp, _vp, _dp]),
    "tb_vec_dmul_add_real_dev": (_i, [_vp, _vp, _vp, _vp, _dp]),
    "tb_apply": (_i, [_vp, _i, _vp, _vp]),
    "tb_cg": (_i, [_vp, _vp, _vp, _ip, _ip, _dp]),
    "tb_invert": (_i, [_vp, _vp, _vp, _ip, _ip, _dp]),
    "tb_cg_gauge": (_i, [_vp, _vp, _vp, _vp, _ip, _ip, _dp]),
    "tb_vec_doubles": (C.c_size_t, [_vp]),
    "tb_pack_dev": (_i, [_vp, _vp, _vp]),
    "tb_unpack_dev": (_i, [_vp, _vp, _vp]),
    "tb_set_gauge_dev": (_i, [_vp, _vp]),
    "tb_apply_dev": (_i, [_vp, _i, _vp, _vp]),
    "tb_cg_dev": (_i, [_vp, _vp, _vp]),
    "tb_invert_dev": (_i, [_vp, _vp, _vp]),
    "tb_cg_result": (_i, [_vp, _ip, _ip, _dp]),
    "tb_re_dot_dev": (_i, [_vp, _vp, _vp, _dp]),
    "tb_create_slab": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i]),
    "tb_slab_handle_bytes": (_i, []),
    "tb_slab_export": (_i, [_vp, _vp]),
    "tb_slab_connect": (_i, [_vp, _vp]),
    "tb_hmc_set_coupling": (_i, [_vp, _dp, _i]),
    "tb_hmc_set_chain_offset": (_i, [_vp, C.c_uint]),
    "tb_hmc_heatbath": (_i, [_vp, _i, C.c_ulonglong]),
    "tb_hmc_trajectory": (_i, [_vp, _i, _d, C.c_ulonglong, C.c_uint, _vp, _vp, _vp, _vp, _vp, _ip,
                               C.POINTER(C.c_longlong)]),
    "tb_hmc_cg_failures": (_i, [_vp, _ip]),
    "tb_hmc_force": (_i, [_vp, _vp, _vp, _vp]),
    "tb_hmc_measure": (_i, [_vp, _i, C.c_ulonglong, C.c_uint, _vp, _dp, _dp]),
    "tb_hmc_condensate": (_i, [_vp, _i, C.c_ulonglong, C.c_uint, _vp, _dp, C.POINTER(C.c_longlong)]),
    "tb_get_gauge": (_i, [_vp, _vp]),
    "tb_checkpoint_write": (_i, [_vp, C.c_char_p]),
    "tb_checkpoint_read": (_i, [_vp, C.c_char_p]),
    "tb_checkpoint_set_next_trajectory": (_i, [_vp, C.c_uint]),
    "tb_checkpoint_next_trajectory": (_i, [_vp, C.POINTER(C.c_uint)]),
    "tb_launch_count": (C.c_longlong, [_vp]),
    "tb_reset_launch_count": (_i, [_vp]),
    "tb_last_solve_ms": (_d, [_vp]),
    "tb_measure_fp64_peak": (_i, [_vp, _i, _dp]),
    "tb_measure_fp64_rate": (_i, [_vp, _i, _i, _dp]),
}


def load_library():
    """dlopen the in-tree CUDA library and type every entry point.  No fallback of any kind."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise TBError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C thirring2d_b200/csrc`).  thirring2d_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load_library().tb_last_error()
        raise TBError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
