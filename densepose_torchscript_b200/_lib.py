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
            f"{LIB_PATH} is missing: build it with `python -m densepose_torchscript_b200.build` "
            "(nvcc, sm_100a). There is no fallback path."
        )
    return C.CDLL(LIB_PATH)


lib = _load()

i32, i64, vp, f32, u8p = C.c_int32, C.c_int64, C.c_void_p, C.c_float, C.c_void_p


class Conv2dArgs(C.Structure):
    _fields_ = [
        ("x", vp), ("n", i32), ("h", i32), ("w", i32), ("cin", i32),
        ("x_sn", i64), ("x_sh", i64), ("x_sw", i64),
        ("wgt", vp), ("cin_pad", i32), ("cout_pad", i32), ("bias", vp),
        ("kh", i32), ("kw", i32), ("sy", i32), ("sx", i32),
        ("pad_y", i32), ("pad_x", i32), ("dil", i32),
        ("h_out", i32), ("w_out", i32), ("relu", i32),
        ("res", vp), ("res_sn", i64), ("res_sy", i64), ("res_sx", i64), ("res_shift", i32),
        ("y", vp), ("y_fp32", i32), ("y_sn", i64), ("y_sy", i64), ("y_sx", i64),
        ("n_valid", vp), ("block_n", i32), ("stages", i32), ("tiled", i32),
        ("y_sc", i64), ("epilogue", i32), ("ks", i32), ("phase_taps", i32), ("pair", i32),
        ("x_lo", vp), ("res_lo", vp), ("y_lo", vp),
    ]


class PreprocessArgs(C.Structure):
    _fields_ = [
        ("src", vp), ("src_u8", i32), ("b", i32), ("h0", i32), ("w0", i32),
        ("hr", i32), ("wr", i32), ("inv_scale", f32), ("flip_rgb", i32),
        ("mean", f32 * 3), ("std", f32 * 3), ("dst", vp), ("hp", i32), ("wx", i32),
        ("tables", vp), ("variant", i32), ("dst_lo", vp),
    ]


class RpnArgs(C.Structure):
    _fields_ = [
        ("head", vp * 5), ("h", i32 * 5), ("w", i32 * 5), ("stride", f32 * 5), ("anchors", (f32 * 12) * 5),
        ("b", i32), ("pre_topk", i32), ("post_topk", i32), ("nms_thresh", f32), ("clip_x", f32), ("clip_y", f32),
        ("cand_boxes", vp), ("cand_scores", vp), ("cand_count", vp), ("cand_keep", vp),
        ("prop_boxes", vp), ("prop_scores", vp), ("prop_count", vp),
    ]


class RoiAlignArgs(C.Structure):
    _fields_ = [
        ("feat", vp * 4), ("h", i32 * 4), ("w", i32 * 4), ("scale", f32 * 4), ("n_levels", i32), ("c", i32),
        ("rois", vp), ("n_rois", vp), ("r", i32), ("p", i32), ("out", vp), ("out_fp32", i32),
    ]


class BoxPredictArgs(C.Structure):
    _fields_ = [
        ("head", vp), ("prop_boxes", vp), ("prop_count", vp), ("b", i32), ("r", i32),
        ("score_thresh", f32), ("nms_thresh", f32), ("topk", i32),
        ("scale_x", f32), ("scale_y", f32), ("out_w", f32), ("out_h", f32),
        ("ws_boxes", vp), ("ws_keep", vp),
        ("det_boxes_raw", vp), ("det_boxes", vp), ("det_scores", vp), ("det_count", vp),
    ]


class ResampleArgs(C.Structure):
    _fields_ = [
        ("coarse", vp), ("fine", vp), ("u", vp), ("v", vp), ("d", i32), ("kc", i32), ("s", i32),
        ("box_wh", vp), ("offsets", vp), ("labels", vp), ("uv", vp), ("total_pixels", i64), ("labels_u8", i32),
    ]


class ModelConfig(C.Structure):
    _fields_ = [
        ("depth", i32), ("head", i32), ("decoder_on", i32), ("pooler_res", i32), ("coarse_ch", i32),
        ("score_thresh", f32), ("nms_test", f32), ("rpn_nms", f32),
        ("dets_per_image", i32), ("rpn_pre_topk", i32), ("rpn_post_topk", i32),
        ("min_size", i32), ("max_size", i32),
        ("pixel_mean", f32 * 3), ("pixel_std", f32 * 3), ("input_rgb", i32), ("extra_ch", i32 * 5), ("resize_variant", i32), ("strict", i32),
    ]


class Weight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data0", vp), ("data1", vp), ("cin_pad", i32), ("cout_pad", i32)]


class ForwardIO(C.Structure):
    _fields_ = [
        ("images", vp), ("bgr", i32), ("pred_boxes", vp), ("scores", vp), ("det_count", vp), ("det_offsets", vp),
        ("coarse", vp), ("fine", vp), ("u", vp), ("v", vp), ("out_half", i32), ("extra", vp * 5),
    ]


def _proto(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_proto("dpb200_last_error", C.c_char_p, [])
_proto("dpb200_abi_version", C.c_int, [])
_proto("dpb200_device_ok", C.c_int, [])
_proto("dpb200_conv2d", C.c_int, [C.POINTER(Conv2dArgs), vp])
_proto("dpb200_preprocess", C.c_int, [C.POINTER(PreprocessArgs), vp])
_proto("dpb200_u8_resize_tables", C.c_int, [vp, i32, i32, i32, i32, C.c_double, vp])
_proto("dpb200_maxpool3x3s2", C.c_int, [vp, vp, i32, i32, i32, i32, vp])
_proto("dpb200_upsample2x", C.c_int, [vp, vp, i32, i32, i32, i32, vp])
_proto("dpb200_decoder_merge", C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp])
_proto("dpb200_rpn_proposals", C.c_int, [C.POINTER(RpnArgs), vp])
_proto("dpb200_nms_sorted", C.c_int, [vp, i32, f32, vp, vp])
_proto("dpb200_roi_align", C.c_int, [C.POINTER(RoiAlignArgs), vp])
_proto("dpb200_box_predict", C.c_int, [C.POINTER(BoxPredictArgs), vp])
_proto("dpb200_groupnorm_relu", C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp])
_proto("dpb200_avgpool", C.c_int, [vp, vp, i32, i32, i32, vp, vp])
_proto("dpb200_predictor_upsample", C.c_int, [vp, i32, i32, i32, vp, C.POINTER(vp), C.POINTER(i32), i32, vp])
_proto("dpb200_dp_resample", C.c_int, [C.POINTER(ResampleArgs), vp])
_proto("dpb200_model_create", C.c_int, [C.POINTER(ModelConfig), C.POINTER(Weight), i32, C.POINTER(vp)])
_proto("dpb200_model_destroy", None, [vp])
_proto("dpb200_session_workspace_bytes", C.c_size_t, [vp, i32, i32, i32])
_proto("dpb200_session_create", C.c_int, [vp, i32, i32, i32, i32, vp, C.c_size_t, C.POINTER(vp)])
_proto("dpb200_session_destroy", None, [vp])
_proto("dpb200_session_run", C.c_int, [vp, C.POINTER(ForwardIO), vp])
_proto("dpb200_session_set_graph", C.c_int, [vp, i32])
_proto("dpb200_session_launch_count", C.c_int, [vp])
_proto("dpb200_session_flops", C.c_double, [vp])
_proto("dpb200_session_op_info", C.c_int, [vp, i32, C.c_char_p, i32, C.POINTER(C.c_double)])
_proto("dpb200_session_op_bytes", C.c_int, [vp, i32, C.POINTER(C.c_double)])
_proto("dpb200_session_profile", C.c_int, [vp, C.POINTER(ForwardIO), vp, C.POINTER(f32), i32])
_proto("dpb200_session_geometry", None, [vp, C.POINTER(i32 * 4)])
_proto("dpb200_session_tap", C.c_int, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i64 * 4), C.POINTER(i32)])

EXPORTS = [
    "dpb200_last_error", "dpb200_abi_version", "dpb200_device_ok", "dpb200_conv2d", "dpb200_preprocess",
    "dpb200_u8_resize_tables",
    "dpb200_maxpool3x3s2", "dpb200_upsample2x", "dpb200_decoder_merge", "dpb200_rpn_proposals",
    "dpb200_nms_sorted", "dpb200_roi_align", "dpb200_box_predict", "dpb200_groupnorm_relu", "dpb200_avgpool",
    "dpb200_predictor_upsample", "dpb200_dp_resample", "dpb200_model_create", "dpb200_model_destroy",
    "dpb200_session_workspace_bytes", "dpb200_session_create", "dpb200_session_destroy", "dpb200_session_run", "dpb200_session_set_graph",
    "dpb200_session_launch_count", "dpb200_session_flops", "dpb200_session_op_info", "dpb200_session_op_bytes", "dpb200_session_profile", "dpb200_session_geometry", "dpb200_session_tap",
]


def last_error() -> str:
    return lib.dpb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise DPB200Error(f"{what} failed ({rc}): {last_error()}")


def require_device():
    if not lib.dpb200_device_ok():
        raise DPB200Error("dpb200 needs a CUDA device of compute capability 10.x (B200); none is current")
