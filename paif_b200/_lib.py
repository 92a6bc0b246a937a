"""ctypes binding of the C-ABI library ``libpaif_b200.so`` (include/paif_b200.h).

There is no CPU or PyTorch fallback: if the library is missing it is built in-tree with
nvcc; if that fails the import raises.
"""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpaif_b200.so")

ABI_VERSION = 5                                             # PAIF_ABI_VERSION of include/paif_b200.h
ENGINE_AUTO, ENGINE_DIRECT, ENGINE_TCGEN05 = 0, 1, 2
STORAGE_F32, STORAGE_BF16, STORAGE_F32_BF16 = 0, 1, 2      # PaifConvDesc.storage

_f = C.c_void_p      # device pointers travel as integers
_i = C.c_int
_ll = C.c_longlong


class ConvDesc(C.Structure):
    """Mirror of ``PaifConvDesc`` (include/paif_b200.h)."""
    _fields_ = [
        ("B", _i), ("H", _i), ("W", _i),
        ("nsrc", _i), ("cin_per_src", _i), ("cout", _i),
        ("kh", _i), ("kw", _i), ("dil", _i), ("engine", _i),
        ("src", _f * 3),
        ("weight", _f), ("weight_mma", _f),
        ("ch_scale", _f), ("ch_shift", _f),
        ("pre_res", _f * 2),
        ("out_pre", _f),
        ("mask_src", _f), ("mask_slope", _f),
        ("slope", _f),
        ("post_scale", C.c_float),
        ("post_res", _f * 3),
        ("out", _f), ("out_act2", _f), ("slope2", _f),
        ("chan_partials", _f),
        ("storage", _i),
    ]


class FusionConv(C.Structure):
    """Mirror of ``PaifFusionConv``."""
    _fields_ = [("direct", _f), ("mma_tf32", _f), ("mma_bf16", _f)]


class FusionRDB(C.Structure):
    """Mirror of ``PaifFusionRDB``."""
    _fields_ = [("conv", FusionConv * 3), ("slope", _f)]


class FusionWeights(C.Structure):
    """Mirror of ``PaifFusionWeights`` (include/paif_b200.h)."""
    _fields_ = [
        ("stem_w", _f * 2), ("stem_a", _f * 2), ("gfmix_w", _f * 2), ("c1x1_b", _f * 2),
        ("rdb", FusionRDB * 3),
        ("dil_dense", FusionConv), ("dil_scale", _f), ("dil_shift", _f),
        ("spa_w", _f), ("spa_k", _i),
        ("eca_conv1", FusionConv), ("eca_conv2", FusionConv), ("eca_w1d", _f), ("eca_a", _f),
        ("res_conv7", FusionConv), ("res_merged", FusionConv), ("res_scale", _f), ("res_shift", _f), ("res_a", _f),
        ("out_mma_tf32", _f), ("out_mma_bf16", _f), ("out_wm", _f), ("out_a", _f),
    ]


class FusionRDBGrad(C.Structure):
    """Mirror of ``PaifFusionRDBGrad``."""
    _fields_ = [("c3", FusionConv * 3), ("c2", FusionConv * 2), ("c1", FusionConv)]


class FusionGradWeights(C.Structure):
    """Mirror of ``PaifFusionGradWeights`` (include/paif_b200.h)."""
    _fields_ = [
        ("rdb", FusionRDBGrad * 3),
        ("dil_dense_d", FusionConv), ("eca_conv1_d", FusionConv), ("eca_conv2_d", FusionConv),
        ("res_conv7_d", FusionConv), ("res_merged_d", FusionConv),
        ("c1x1_d", (FusionConv * 3) * 2),
    ]


# name -> argtypes (restype is int unless noted); must list EVERY symbol of the header.
SIGNATURES = {
    "paif_abi_version": [],
    "paif_last_error_string": [],
    "paif_stem_forward": [_f, _ll, _ll, _ll, _f, _f, _f, _f, _i, _i, _i, _f],
    "paif_stem_forward_bf16copy": [_f, _ll, _ll, _ll, _f, _f, _f, _f, _f, _i, _i, _i, _f],
    "paif_dilconv_forward_bf16": [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _f],
    "paif_spa_fused_forward_bf16": [_f, _i, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_eca_apply_bf16": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_out_forward_bf16": [_f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_out_forward_tc": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f],
    "paif_gf_guide_stats": [_f, _f, _i, _i, _i, _f],
    "paif_gf_decomp_forward": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_gf_mix_supported": [_i, _i, _i],
    "paif_gf_mix_forward": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f],
    "paif_gf_mix_forward_save": [_f, _f, _f, _f, _f, _f, _i, _f, _i, _i, _i, _i, _f],
    "paif_conv_forward": [C.POINTER(ConvDesc), _f],
    "paif_conv_num_tiles": [_i, _i, _i],
    "paif_conv_tc_kq": [_i, _i, _i],
    "paif_conv_tc_kq_bf16": [_i, _i, _i],
    "paif_conv_set_persistent": [_i],
    "paif_dwconv_forward": [_f, _f, _i, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f],
    "paif_dilconv_forward": [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _f],
    "paif_add_act": [_f, _f, _f, _f, _f, _ll, _f],
    "paif_channel_pool": [_f, _f, _f, _i, _i, _i, _i, _f],
    "paif_spa_blend_forward": [_f, _f, _i, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_spa_fused_forward": [_f, _i, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_eca_scale": [_f, _i, _f, _i, _f, _i, _i, _i, _i, _f],
    "paif_eca_apply": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_out_forward": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_out_backward": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_mask_scale": [_f, _f, _f, C.c_float, _f, _i, _i, _i, _i, _f],
    "paif_add_maps": [_f, _f, _f, _f, _ll, _f],
    "paif_eca_bwd_pass1": [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_eca_bwd_tiles": [_i, _i],
    "paif_eca_bwd_scale": [_f, _i, _f, _f, _i, _f, _i, _i, _i, _i, _f],
    "paif_eca_bwd_pass2": [_f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_spa_blend_backward_pre": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_spa_blend_backward": [_f, _f, _f, _f, _f, _f, _i, _f, _f, _i, _i, _i, _i, _f],
    "paif_gf_backward_work_floats": [_i, _i, _i, _i],
    "paif_gf_guide_parts": [_i],
    "paif_gf_decomp_backward": [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_gf_decomp_backward_saved": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_stem_backward_pre": [_f, _f, _f, _f, _f, _f, _f, _i, _f, _i, _i, _i, _i, _f],
    "paif_stem_backward": [_f, _f, _f, _i, _i, _i, _i, _f],
    "paif_confusion_accumulate": [_f, _f, _ll, _i, _f, _f],
    "paif_pgd_step": [_f, _f, _f, C.c_float, C.c_float, _ll, _f],
    "paif_fusion_workspace_bytes": [_i, _i, _i, _i],
    "paif_fusion_forward": [C.POINTER(FusionWeights), _f, _ll, _ll, _ll, _f, _ll, _ll, _ll, _f, _f, _ll, _i, _i, _i, _i, _f],
    "paif_fusion_train_workspace_bytes": [_i, _i, _i],
    "paif_fusion_forward_save": [C.POINTER(FusionWeights), _f, _ll, _ll, _ll, _f, _ll, _ll, _ll, _f, _f, _ll, _i, _i, _i, _f],
    "paif_fusion_backward_input": [C.POINTER(FusionWeights), C.POINTER(FusionGradWeights), _f, _f, _f, _f, _ll, _i, _i, _i, _f],
    "paif_widen_bf16_map": [_f, _f, _i, _i, _i, _i, _f],
    "paif_segloss_forward": [_f, _f, _f, _f, _ll, C.c_float, _i, _i, _i, _i, _i, _i, _f],
    "paif_segloss_backward": [_f, _f, _f, _i, _i, _i, _i, _i, _i, _f],
    "paif_glue_blocks": [_i, _i],
    "paif_glue_forward": [_f, _f, C.POINTER(C.c_float), C.POINTER(C.c_float), _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_glue_backward": [_f, _f, _f, C.POINTER(C.c_float), _f, _f, _f, _f, _f, _i, _i, _i, _i, _f],
    "paif_stem_forward_rgb": [_f, _ll, _ll, _ll, _ll, _f, _f, _f, _f, _f, _i, _i, _i, _f],
}

_lib = None


class PaifError(RuntimeError):
    pass


def load():
    """Load (building first if needed) the C-ABI library; raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("PAIF_B200_LIB"):
        path = os.environ["PAIF_B200_LIB"]             # development only: an explicitly chosen build of the library
    elif os.environ.get("PAIF_B200_PROFILE_LIB") == "1":
        # development only: same sources compiled with -DPAIF_TC_PROFILE (role timeline counters in the conv engine)
        path = _build.build(profile=True)
    elif not os.path.exists(LIB_PATH) or (_build.needs_build() and os.environ.get("PAIF_NO_REBUILD") != "1"):
        _build.build()
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = (C.c_char_p if name == "paif_last_error_string" else
                      _ll if name in ("paif_gf_backward_work_floats", "paif_fusion_workspace_bytes",
                                  "paif_fusion_train_workspace_bytes") else _i)
    if lib.paif_abi_version() != ABI_VERSION:
        raise PaifError("libpaif_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().paif_last_error_string()
        raise PaifError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def call(name, *args):
    check(getattr(load(), name)(*args), name)
