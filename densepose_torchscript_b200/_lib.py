"""ctypes binding of the C-ABI library (include/dpb200.h).

The CUDA library is the product: there is no Python / CPU fallback.  Importing this module
without a built ``libdpb200.so`` raises, and calling any op without an sm_100 device raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpb200.so")


class DPB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise DPB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no fallback path."
        )
    return C.CDLL(LIB_PATH)


lib = _load()

c_i32, c_i64, c_vp, c_fp = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p


class Conv2dArgs(C.Structure):
    _fields_ = [
        ("x", c_vp), ("n", c_i32), ("h", c_i32), ("w", c_i32), ("cin", c_i32),
        ("x_sn", c_i64), ("x_sh", c_i64), ("x_sw", c_i64),
        ("wgt", c_vp), ("cin_pad", c_i32), ("cout_pad", c_i32), ("bias", c_vp),
        ("kh", c_i32), ("kw", c_i32), ("sy", c_i32), ("sx", c_i32),
        ("pad_y", c_i32), ("pad_x", c_i32), ("dil", c_i32),
        ("h_out", c_i32), ("w_out", c_i32), ("relu", c_i32),
        ("res", c_vp), ("res_sn", c_i64), ("res_sy", c_i64), ("res_sx", c_i64), ("res_shift", c_i32),
        ("y", c_vp), ("y_fp32", c_i32), ("y_sn", c_i64), ("y_sy", c_i64), ("y_sx", c_i64),
        ("n_valid", c_vp), ("block_n", c_i32), ("stages", c_i32), ("tiled", c_i32),
    ]


lib.dpb200_last_error.restype = C.c_char_p
lib.dpb200_abi_version.restype = C.c_int
lib.dpb200_device_ok.restype = C.c_int
lib.dpb200_conv2d.argtypes = [C.POINTER(Conv2dArgs), c_vp]
lib.dpb200_conv2d.restype = C.c_int


def last_error() -> str:
    return lib.dpb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise DPB200Error(f"{what} failed ({rc}): {last_error()}")


def require_device():
    if not lib.dpb200_device_ok():
        raise DPB200Error("dpb200 needs a CUDA device of compute capability 10.x (B200); none is current")
