"""ctypes binding of libophelia_sm100.so (declared in include/ophelia_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a call fails, we raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OPH_LIB_PATH") or os.path.join(_HERE, "libophelia_sm100.so")      # (override: A/B of two builds)

P, LL, I, F, D, U64, SZ = (ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                          ctypes.c_uint64, ctypes.c_size_t)

class Act(ctypes.Structure):
    """`oph_act`: fp32 view and/or split-bf16 planes of one activation (include/ophelia_b200.h)."""
    _fields_ = [("f32", P), ("ld", LL), ("hi", P), ("lo", P), ("ldp", LL)]


AP = ctypes.POINTER(Act)


class Guide(ctypes.Structure):
    """`oph_guide`: per-utterance attention targets of the batch (include/ophelia_b200.h)."""
    _fields_ = [("w", P), ("item_stride", LL), ("ld", LL), ("Ng", I), ("Tg", I), ("pad", F), ("mse", I),
                ("col_g", P), ("col_h", P), ("c_aout", F)]


GP = ctypes.POINTER(Guide)


class ArLayer(ctypes.Structure):
    """`oph_ar_layer`: one AudioEnc layer of the fused frame step (include/ophelia_b200.h)."""
    _fields_ = [("w", P), ("bias", P), ("g1", P), ("b1", P), ("g2", P), ("b2", P), ("x", P), ("y", P),
                ("x_item", LL), ("ldx", LL), ("y_item", LL), ("ldy", LL),
                ("Cin", I), ("C", I), ("k", I), ("rate", I), ("kind", I), ("act", I), ("in_shift", I)]

# name -> (restype, argtypes); must list every symbol of include/ophelia_b200.h
SIGNATURES = {
    "oph_version": (I, []),
    "oph_last_error": (ctypes.c_char_p, []),
    "oph_crc32c": (ctypes.c_uint, [P, ctypes.c_ulonglong, ctypes.c_uint]),
    "oph_launch_count": (LL, []),
    "oph_gemm_debug_buffer": (I, [P]),
    "oph_gemm_debug_ring": (I, [P, I]),
    "oph_gemm_debug_ring_desc": (I, [I, P, I]),
    "oph_gemm_debug_flags": (I, [I]),
    "oph_last_trap": (I, [P, P, I]),
    "oph_debug_trap_selftest": (I, [P]),
    "oph_wgrad_stream": (I, [P, I]),
    "oph_cache_config": (I, [I]),
    "oph_profile_begin": (I, []),
    "oph_profile_end": (I, [P]),
    "oph_conv_pack_bytes": (SZ, [I, I, I, I, I]),
    "oph_conv_pack": (I, [P, I, I, I, I, P, P, P]),
    "oph_pack_job_bytes": (SZ, []),
    "oph_pack_plan_add": (I, [P, I, P, P, P, I, I, I, I, P, P]),
    "oph_pack_run": (I, [P, I, LL, P]),
    "oph_conv1d_fwd": (I, [AP, P, P, P, P, P, LL, P, AP, P, LL, I, I, I, I, I, I, I, I, I, I, F, U64, P, P]),
    "oph_conv1d_bwd": (I, [P, LL, AP, P, LL, P, P, P, P, P, LL, P, LL, P, P, P, P,
                           I, I, I, I, I, I, I, I, I, I, F, U64, P, P]),
    "oph_normalize_fwd": (I, [P, LL, P, P, AP, P, LL, I, P]),
    "oph_normalize_bwd": (I, [P, LL, P, LL, P, P, P, P, LL, P, P, LL, I, P]),
    "oph_hc_fwd": (I, [AP, P, P, P, P, P, P, P, LL, P, AP, I, I, I, I, I, I, I, F, U64, P, P]),
    "oph_hc_bwd": (I, [P, LL, AP, P, LL, P, P, P, P, P, P, P, LL, P, LL, P, LL, P, P, P, P, P, P,
                       I, I, I, I, I, I, I, F, U64, P, P]),
    "oph_lcc_context": (I, [P, P, P]),
    "oph_lcc_fwd": (I, [P, LL, P, P, AP, P, LL, I, I, I, P]),
    "oph_lcc_bwd": (I, [P, LL, P, LL, P, P, P, LL, P, I, I, I, P]),
    "oph_lcc_reduce": (I, [P, P, P, I, I, I, P]),
    "oph_deconv_fwd": (I, [AP, P, P, P, P, P, LL, P, AP, I, I, I, F, U64, P, P]),
    "oph_deconv_bwd": (I, [P, LL, AP, P, LL, P, P, P, P, P, LL, P, LL, P, P, P, P, I, I, I, F, U64, P, P]),
    "oph_embed_fwd": (I, [P, P, P, LL, I, I, P]),
    "oph_embed_bwd": (I, [P, P, LL, P, I, I, P]),
    "oph_attention_fwd": (I, [AP, AP, AP, AP, AP, P, P, P, I, P, I, I, F, I, I, I, I, GP, P]),
    "oph_attention_bwd": (I, [AP, AP, AP, AP, AP, AP, P, LL, P, LL, P, LL, P, LL, F, I, I, F, I, I, I, I, GP, P]),
    "oph_attention_guide_sum": (I, [P, LL, I, I, I, P, I, I, F, GP, P]),
    "oph_attention_extra_fwd": (I, [P, LL, I, I, I, F, F, P, P, P, P]),
    "oph_attention_extra_finalize": (I, [P, P, I, I, I, F, F, F, I, P]),
    "oph_split_planes": (I, [P, LL, LL, I, P, P, LL, P]),
    "oph_recon_loss": (I, [P, LL, P, LL, P, LL, LL, I, I, F, F, F, P, P]),
    "oph_loss_finalize": (I, [P, P, D, D, F, F, F, F, I, I, P]),
    "oph_adam_prepare": (I, [P, P, F, F, F, I, F, P]),
    "oph_adam_clip": (I, [P, P, P, P, LL, P, F, F, F, F, F, P]),
    "oph_step_inc": (I, [P, P]),
    "oph_ar_encoder_step": (I, [ctypes.POINTER(ArLayer), I, I, P, P]),
    "oph_ar_scratch_floats": (SZ, []),
    "oph_ar_conv_step": (I, [P, LL, LL, P, P, P, P, P, LL, LL, P, LL, LL, P, I, I, I, I, I, I, I, P, P]),
    "oph_ar_hc_step": (I, [P, LL, LL, P, P, P, P, P, P, P, LL, LL, P, I, I, I, I, P, P]),
    "oph_ar_window_gather": (I, [P, LL, LL, P, LL, LL, I, I, I, I, I, P, P]),
    "oph_ar_window_scatter": (I, [P, LL, LL, P, LL, LL, I, P, P, P, P, P, I, I, I, I, I, P, P]),
    "oph_ar_advance": (I, [P, P]),
    "oph_gemm_nt": (I, [P, LL, P, LL, P, LL, P, I, I, I, I, F, I, LL, LL, LL, P]),
    "oph_gemm_tn": (I, [P, LL, P, LL, P, LL, I, I, I, I, P]),
}

_lib = None


class OpheliaError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OpheliaError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA path has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("OPH_LIB_PATH"):
            continue                                    # an older build under A/B comparison lacks the newer entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    if os.environ.get("OPH_DEBUG_FLAGS"):            # diagnostics only (A/B timing of scheduling features)
        set_debug_flags(int(os.environ["OPH_DEBUG_FLAGS"], 0))
    return lib


_debug_flags = 0


def set_debug_flags(flags):
    """oph_gemm_debug_flags, remembered on the host (a few wrappers size their buffers by the path the flags select)."""
    global _debug_flags
    _debug_flags = int(flags)
    load().oph_gemm_debug_flags(_debug_flags)


def debug_flags():
    return _debug_flags


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise OpheliaError("%s failed (%d): %s" % (name, rc, lib.oph_last_error().decode()))


def pack_bytes(k, cin, cout, deconv, backward):
    return int(load().oph_conv_pack_bytes(k, cin, cout, int(deconv), int(backward)))
