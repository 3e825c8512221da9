"""ctypes binding of libscn_b200.so (the C-ABI in include/scn_b200.h).

The product path is CUDA only: if the shared library is missing this module raises at import
time - there is no CPU or eager-PyTorch fallback behind it.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8,
                    c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCN_B200_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "libscn_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "sparseconvnet (B200): %s not found. Build it with "
        "`python __graft_entry__.py build` (or csrc/build.sh); there is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

# include/scn_b200.h SCN_PRECISION_*.  "fp32" is the library's fp32 mode: 3xTF32 on the tcgen05 tensor
# cores (hi/lo operand split, fp32 accumulation, error ~2^-20 relative - inside the 1e-4 parity bound),
# with the exact FFMA tiles for the shapes the tensor path does not take (Cin = 9 stem).
# "bf16": forward / input-gradient contractions with bf16 operands (fp32 accumulation), weight gradient in tf32.
PRECISIONS = {"fp32_ffma": 0, "tf32": 1, "fp32": 2, "bf16": 3}
_precision = PRECISIONS[os.environ.get("SCN_B200_PRECISION", "fp32")]


def set_conv_precision(name):
    """'fp32' (3xTF32 on the tensor cores, fp32-accurate), 'fp32_ffma' (exact FFMA tiles everywhere),
    'tf32' (single-pass tf32 operands, fp32 accumulate) or 'bf16' (bf16 operands, fp32 accumulate)."""
    global _precision
    _precision = PRECISIONS[name]


def get_conv_precision():
    return {v: k for k, v in PRECISIONS.items()}[_precision]


def precision():
    return _precision


I64P = POINTER(c_int64)
_sig = {
    "scn_last_error": (c_char_p, []),
    "scn_version": (c_int, []),
    "scn_n_rulebook_bits": (c_int, []),
    "scn_metadata_create": (c_int, [c_int, POINTER(c_void_p)]),
    "scn_metadata_destroy": (None, [c_void_p]),
    "scn_metadata_clear": (c_int, [c_void_p, c_void_p]),
    "scn_get_nactive": (c_int, [c_void_p, I64P, I64P]),
    "scn_get_batch_size": (c_int, [c_void_p, I64P]),
    "scn_get_spatial_locations": (c_int, [c_void_p, I64P, c_void_p, c_void_p]),
    "scn_get_spatial_locations_device": (c_int, [c_void_p, I64P, c_void_p, c_void_p]),
    "scn_quantize_points": (c_int, [c_void_p, c_int64, c_double, I64P, c_int64, c_void_p, c_void_p,
                                    I64P, c_void_p]),
    "scn_voxelize_batch": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, POINTER(c_double), c_double, I64P,
                                   c_int, c_void_p, c_void_p, I64P, c_void_p]),
    "scn_input_layer_prepare": (c_int, [c_void_p, I64P, c_void_p, c_int64, c_int, c_int, c_int64,
                                        c_int, c_void_p, I64P]),
    "scn_build_plan": (c_int, [c_void_p, I64P, c_void_p, c_int64, c_int, c_int, c_int64, c_int, c_void_p, c_int,
                               c_void_p, I64P]),
    "scn_input_layer_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "scn_input_layer_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "scn_output_layer_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "scn_output_layer_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "scn_input_rulebook_header": (c_int, [c_void_p, I64P, c_void_p]),
    "scn_input_rulebook_copy": (c_int, [c_void_p, c_void_p, c_void_p]),
    "scn_submanifold_rulebook_prepare": (c_int, [c_void_p, I64P, I64P, c_void_p, I64P]),
    "scn_conv_rulebook_prepare": (c_int, [c_void_p, I64P, I64P, I64P, I64P, c_void_p, I64P, I64P]),
    "scn_submanifold_rulebook_copy": (c_int, [c_void_p, I64P, I64P, c_int64, c_void_p, c_void_p]),
    "scn_conv_rulebook_copy": (c_int, [c_void_p, I64P, I64P, I64P, c_int64, c_void_p, c_void_p]),
    "scn_sparse_to_dense_rules_copy": (c_int, [c_void_p, I64P, c_void_p, c_void_p, c_void_p]),
    "scn_submanifold_conv_forward": (c_int, [c_void_p, I64P, I64P, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int64, c_int64, c_int, c_void_p,
                                             POINTER(c_double), I64P]),
    "scn_submanifold_conv_backward": (c_int, [c_void_p, I64P, I64P, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                              c_void_p, I64P]),
    "scn_nin_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                c_int, c_void_p, POINTER(c_double)]),
    "scn_nin_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int64, c_int64, c_int, c_void_p]),
    "scn_batchnorm_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_float, c_float, c_int, c_float, c_int64,
                                      c_int64, c_void_p]),
    "scn_batchnorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int64,
                                       c_void_p]),
    "scn_batchnorm_backward_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int64,
                                           c_void_p, c_void_p]),
    "scn_batchnorm_backward_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int, c_void_p, c_void_p, c_float, c_int64, c_int64, c_void_p,
                                             c_void_p]),
    "scn_graph_forward": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_void_p, POINTER(c_double)]),
    "scn_graph_backward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p]),
    "scn_graph_backward_marked": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_int, c_void_p, c_int32, c_void_p, c_void_p]),
    "scn_set_graph_overlap": (c_int, [c_int]),
    "scn_set_pdl": (c_int, [c_int]),
    "scn_event_create": (c_int, [POINTER(c_void_p)]),
    "scn_event_destroy": (c_int, [c_void_p]),
    "scn_event_record": (c_int, [c_void_p, c_void_p]),
    "scn_stream_wait_event": (c_int, [c_void_p, c_void_p]),
    "scn_sparse_to_dense_forward": (c_int, [c_void_p, I64P, c_void_p, c_void_p, c_int64, c_int64,
                                            c_void_p]),
    "scn_grid_extent": (c_int, [c_void_p, I64P, I64P, c_void_p]),
    "scn_sparse_to_dense_cropped_forward": (c_int, [c_void_p, I64P, I64P, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "scn_sparse_to_dense_cropped_backward": (c_int, [c_void_p, I64P, I64P, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "scn_sparse_to_dense_backward": (c_int, [c_void_p, I64P, c_void_p, c_void_p, c_int64, c_int64,
                                             c_void_p]),
    "scn_scale_inplace": (c_int, [c_void_p, c_float, c_int64, c_void_p]),
    "scn_rotate_iou": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "scn_rotate_nms": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_int64, c_int64, c_void_p, I64P, c_void_p]),
    "scn_grid_anchors": (c_int, [c_void_p, I64P, c_void_p, c_int64, c_float, POINTER(c_float), c_void_p, c_void_p,
                                 c_int64, c_void_p]),
    "scn_rpn_head_forward": (c_int, [c_void_p, c_int64, c_int64] + [c_void_p] * 4 + [c_int64, c_void_p, c_void_p,
                                     c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "scn_rpn_head_backward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_int64] + [c_void_p] * 10 + [c_int, c_void_p]),
    "scn_roi_align_rotated_3d_forward": (c_int, [c_void_p, I64P, c_void_p, c_int64, c_void_p, c_int64, c_float,
                                                 I64P, c_int, c_void_p, c_void_p]),
    "scn_roi_align_rotated_3d_backward": (c_int, [c_void_p, I64P, c_void_p, c_int64, c_void_p, c_int64, c_float,
                                                  I64P, c_int, c_void_p, c_void_p]),
    "scn_launch_count": (c_int64, []),
    "scn_rulebook_stats": (c_int, [c_void_p, c_int, I64P, I64P, I64P, I64P]),
    "scn_set_tile_grouping": (c_int, [c_int]),
    "scn_set_gemm_grid_limit": (c_int, [c_int]),
    "scn_prof_enable": (c_int, [c_int]),
    "scn_prof_read": (c_int, [c_int, POINTER(c_double)]),
}
for _name in ("scn_conv_forward", "scn_deconv_forward"):
    _sig[_name] = (c_int, [c_void_p, I64P, I64P, I64P, I64P, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_int64, c_int64, c_int, c_void_p, POINTER(c_double), I64P])
for _name in ("scn_conv_backward", "scn_deconv_backward"):
    _sig[_name] = (c_int, [c_void_p, I64P, I64P, I64P, I64P, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, I64P])

EXPORTS = sorted(_sig)
for _name, (_res, _args) in _sig.items():
    _f = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _f.restype = _res
    _f.argtypes = _args


def check(status):
    if status != 0:
        raise RuntimeError("libscn_b200: " + lib.scn_last_error().decode("utf-8", "replace"))


def i64x3(t):
    """torch LongTensor / list of 3 ints -> ctypes int64[3].  Size tensors are shared by all layers of a
    scale and never mutated in place, so the converted array is cached on the tensor object."""
    if isinstance(t, torch.Tensor):
        c = getattr(t, "_scn_i64x3", None)
        if c is None:
            v = t.tolist()
            c = (c_int64 * 3)(int(v[0]), int(v[1]), int(v[2]))
            t._scn_i64x3 = c
        return c
    return (c_int64 * 3)(int(t[0]), int(t[1]), int(t[2]))


_next_token = [1]


def weight_tag(w):
    """{identity token, in-place version} of a weight tensor for the library's packed-operand cache.  The
    token is minted once per tensor OBJECT (never reused, unlike data_ptr / id()); optimizers and
    load_state_dict update parameters in place, which bumps `_version`."""
    t = getattr(w, "_scn_token", None)
    if t is None:
        t = _next_token[0]
        _next_token[0] += 1
        try:
            w._scn_token = t
        except AttributeError:
            return None
    return (c_int64 * 2)(t, w._version)


def stream():
    """torch's current CUDA stream as a cudaStream_t (raw getter: torch.cuda.current_stream() builds a
    Stream object and costs ~15 us per call, several hundred calls per step)"""
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def ptr(t):
    """device/host pointer of an optional tensor (None or empty -> NULL)"""
    if t is None or t.numel() == 0:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def require_cuda_f32(t, what):
    if not t.is_cuda:
        raise RuntimeError("sparseconvnet (B200): %s must be a CUDA tensor - this build has no "
                           "CPU path (got device %s)" % (what, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("sparseconvnet (B200): %s must be float32, got %s" % (what, t.dtype))
    return t if t.is_contiguous() else t.contiguous()
