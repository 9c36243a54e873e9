"""ctypes binding of libfnm_b200.so (C ABI declared in include/fnm_b200.h).

The CUDA library is the product: there is NO Python / PyTorch / CPU fallback.  If the shared
object is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfnm_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

FNM_W_REFERENCE, FNM_W_MODE_MAJOR = 0, 1
FNM_PLAN_ZOUT_IS_GELU_GRAD, FNM_PLAN_PRE_IS_GELU_GRAD = 1, 2
FNM_PLAN_BWD_WEIGHTS_ONLY, FNM_PLAN_BWD_INPUT_ONLY = 4, 8
FNM_PW_MUL_IS_GRAD, FNM_PW_OUT_IS_GRAD = 1, 2
FNM_MODE_FP32 = 0
FNM_MODE_TF32 = 1
ADAM_MAX_TENSORS = 48

_vp, _i32, _i64, _sz, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double


class FnmPlan(C.Structure):
    _fields_ = [("B", C.c_int32), ("P1", C.c_int32), ("P2", C.c_int32), ("S1", C.c_int32), ("S2", C.c_int32),
                ("R", C.c_int32), ("NY", C.c_int32), ("Ci", C.c_int32), ("Co", C.c_int32),
                ("rows_per_w", C.c_int32), ("mode", C.c_int32), ("w_layout", C.c_int32),
                ("ay", _vp), ("ayT", _vp), ("by", _vp), ("byT", _vp),
                ("ex", _vp), ("fx", _vp), ("exb", _vp), ("fxb", _vp), ("w_shadow", _vp), ("flags", C.c_int32),
                ("by16", _vp), ("ayT16", _vp), ("ay4", _vp), ("byT4", _vp)]


_PP = C.POINTER(FnmPlan)

# name -> (restype, argtypes); mirrors include/fnm_b200.h one to one
PROTOTYPES = {
    "fnm_version": (C.c_char_p, []),
    "fnm_plan_sizeof": (C.c_size_t, []),
    "fnm_last_error": (C.c_char_p, []),
    "fnm_has_tcgen05": (_i32, []),
    "fnm_launch_count": (_i64, []),
    "fnm_set_sm_limit": (_i32, [_i32]),
    "fnm_profile_begin": (_i32, []),
    "fnm_profile_end": (_sz, [C.c_char_p, _sz]),
    "fnm_fourier_layer_workspace_bytes": (_sz, [_PP]),
    "fnm_fourier_layer_shadow_bytes": (_sz, [_PP]),
    "fnm_fourier_layer_fwd": (_i32, [_PP, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_fourier_layer_bwd": (_i32, [_PP, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_lfunc_workspace_bytes": (_sz, [_PP]),
    "fnm_lfunc_fwd": (_i32, [_PP, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_lfunc_bwd": (_i32, [_PP, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_ldec_workspace_bytes": (_sz, [_PP]),
    "fnm_ldec_fwd": (_i32, [_PP, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_ldec_bwd": (_i32, [_PP, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fnm_gelu_fwd": (_i32, [_vp, _vp, _i64, _vp]),
    "fnm_gelu_bwd": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "fnm_adam_multi_tensor": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _dbl,
                                     _i32, _i32, _vp, _vp]),
    "fnm_counter_inc": (_i32, [_vp, _vp]),
    "fnm_ydft": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fnm_xdft": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fnm_xsyn": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fnm_resample": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fnm_mix_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "fnm_mix_fwd": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp]),
    "fnm_mix_bwd_x": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp]),
    "fnm_mix_bwd_w": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp]),
    "fnm_pointwise_ysyn": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                  _i32, _vp]),
    "fnm_pointwise_ysyn_ex": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                     _i32, _i32, _vp, _vp]),
    "fnm_lift_fwd": (_i32, [_vp, _vp, _vp, _vp] + [_i32] * 8 + [_vp]),
    "fnm_lift_bwd_workspace_bytes": (_sz, [_i32, _i32]),
    "fnm_lift_bwd": (_i32, [_vp, _vp, _vp, _vp] + [_i32] * 8 + [_vp, _sz, _vp]),
    "fnm_mlp_head_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "fnm_mlp_head_bwd_workspace_bytes": (_sz, [_i32, _i32]),
    "fnm_mlp_head_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz, _vp]),
    "fnm_rel_l2_workspace_bytes": (_sz, [_i32]),
    "fnm_rel_l2_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, C.c_float, _vp, _sz, _vp]),
    "fnm_rel_l2_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "fnm_local_functional_workspace_bytes": (_sz, [_i32] * 5),
    "fnm_local_functional_fwd": (_i32, [_vp] * 6 + [_i32] * 5 + [_vp, _sz, _i32, _vp]),
    "fnm_local_functional_bwd": (_i32, [_vp] * 7 + [_i32] * 5 + [_i32, _vp]),
    "fnm_wsum_gelu_workspace_bytes": (_sz, [_i32, _i32]),
    "fnm_wsum_gelu_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "fnm_wsum_gelu_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fnm_conv_wgrad_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "fnm_conv_wgrad": (_i32, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _sz, _i32, _vp]),
    "fnm_ydft_conv_wgrad": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _sz, _i32, _vp]),
}

_lib = None


class FnmError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libfnm_b200.so in-tree for sm_100a (nvcc cross-compiles; no GPU needed)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise FnmError("building libfnm_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FnmError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
                "There is no CPU / PyTorch fallback for the Fourier-layer hot path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        if handle.fnm_plan_sizeof() != C.sizeof(FnmPlan):
            raise FnmError(f"{LIB_PATH}: struct fnm_plan has {handle.fnm_plan_sizeof()} bytes, this binding's mirror "
                           f"{C.sizeof(FnmPlan)} -- rebuild the library (make -C {CSRC_DIR}) or update _lib.FnmPlan")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().fnm_last_error().decode(errors="replace")
        raise FnmError(f"{what or 'libfnm_b200'} failed (code {rc}): {msg}")


def version():
    return lib().fnm_version().decode()


def set_sm_limit(n):
    """Cap the grid of the persistent kernels at ``n`` CTAs (0 = every SM); returns the previous cap."""
    return int(lib().fnm_set_sm_limit(int(n)))


def launch_count():
    return int(lib().fnm_launch_count())


def profile_begin():
    check(lib().fnm_profile_begin(), "fnm_profile_begin")


def profile_end():
    """{kernel name: (launches, total_ms)} of everything launched since profile_begin()."""
    buf = C.create_string_buffer(1 << 16)
    lib().fnm_profile_end(buf, len(buf))
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, cnt, ms = item.rsplit(":", 2)
            out[name] = (int(cnt), float(ms))
    return out
