"""ctypes binding of libtinyfusers_b200.so — the B200 replacement for the reference's
`tinyfusers/native/{cuda,cublas,nvrtc}/ops.py` singletons (reference: native/cublas/ops.py:3-70).

Same style as the reference: one class owning `self.dll = ctypes.CDLL(...)`, explicit
`restype` / `argtypes`, methods that return the raw integer status, a module-level singleton
(`b200`) and status constants on it. Callers turn a non-zero status into
`RuntimeError("<fn> failed with status <n>")` exactly like `ff/linear.py:100-103`; `check()` does that
and appends the library's error string.

There is deliberately no fallback: if the shared library is missing, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtinyfusers_b200.so")

c_int, c_size_t, c_void_p, c_float, c_longlong = (ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                                  ctypes.c_float, ctypes.c_longlong)
_P = c_void_p

class GemmExtras(ctypes.Structure):
    """tf_gemm_extras (include/tinyfusers_b200.h): optional fused extras of tf_gemm_ex_f16."""
    _fields_ = [("gn_stats", c_void_p), ("gn_unit", c_int), ("gn_rows_per_image", c_int), ("row_stats_out", c_void_p),
                ("ln_stats", c_void_p), ("ln_chunks", c_int), ("ln_c1", c_void_p), ("ln_eps", c_float)]


# name -> (restype, argtypes). Keep in sync with include/tinyfusers_b200.h (tests/test_abi.py checks it).
_SIGNATURES = {
    "tf_version": (c_int, []),
    "tf_init": (c_int, [c_int]),
    "tf_last_error": (ctypes.c_char_p, []),
    "tf_set_pdl": (c_int, [c_int]),
    "tf_launch_count": (c_longlong, []),
    "tf_launch_count_reset": (None, []),
    "tf_gemm_set_tuning": (c_int, [c_int, c_int]),
    "tf_gemm_set_ctas": (c_int, [c_int]),
    "tf_gemm_set_cluster_splitk": (c_int, [c_int]),
    "tf_graph_begin_capture": (c_int, [_P]),
    "tf_graph_end_capture": (c_int, [_P, _P]),
    "tf_graph_launch": (c_int, [_P, _P]),
    "tf_graph_destroy": (c_int, [_P]),
    "tf_gemm_set_timeline": (c_int, [_P]),
    "tf_weight_prefetch_mode": (c_int, [c_int]),
    "tf_weight_prefetch_stats": (c_int, [_P, _P, _P]),
    "tf_weight_prefetch_limits": (c_int, [ctypes.c_longlong, ctypes.c_longlong]),
    "tf_gemm_set_max_stages": (c_int, [c_int]),
    "tf_gemm_tuning_add": (c_int, [c_int] * 8),
    "tf_gemm_tuning_clear": (c_int, []),
    "tf_gemm_last_choice": (c_int, [_P, _P, _P]),
    "tf_gemm_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int,
                            _P, c_size_t, _P]),
    "tf_conv2d_nhwc_f16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, c_int,
                                   _P, _P, c_int, c_int, _P, c_size_t, _P]),
    "tf_gemm_ex_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int,
                               _P, c_size_t, ctypes.POINTER(GemmExtras), _P]),
    "tf_gemm_gn_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int,
                               _P, c_size_t, _P, c_int, c_int, _P]),
    "tf_conv2d_nhwc_gn_f16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, c_int,
                                      _P, _P, c_int, c_int, _P, c_size_t, _P, c_int, _P]),
    "tf_gn_stats_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tf_conv2d_up2x_nhwc_f16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, c_int, _P, c_int, _P]),
    "tf_conv2d_nhwc_skip_f16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, c_int, _P, c_int,
                                        _P, c_size_t, _P, c_int, _P]),
    "tf_groupnorm_fused_nhwc_f16": (c_int, [_P, c_int, c_int, _P, c_int, _P, c_int, c_int, _P, c_int, _P, c_int, c_int,
                                            c_int, c_int, _P, _P, c_float, c_int, _P]),
    "tf_groupnorm_nhwc_f16": (c_int, [_P, c_int, c_int, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, _P,
                                      c_float, c_int, _P, _P]),
    "tf_groupnorm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tf_layernorm_f16": (c_int, [_P, _P, c_int, c_int, _P, _P, c_float, c_int, _P]),
    "tf_attention_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_longlong, c_longlong, c_longlong, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "tf_attention_v_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_longlong, c_longlong, c_longlong, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "tf_attention_set_tuning": (c_int, [c_int]),
    "tf_attention_set_variant": (c_int, [c_int, c_int]),
    "tf_attention_set_timeline": (c_int, [_P]),
    "tf_attention_set_debug": (c_int, [_P]),
    "tf_attention_causal_f16": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_longlong, c_longlong, c_longlong, c_int,
                                        c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "tf_plane_attention_f16": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "tf_softmax_rows_f32_to_f16": (c_int, [_P, c_longlong, _P, c_longlong, c_longlong, c_int, c_float, _P]),
    "tf_conv1x1_small_f32nchw": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    "tf_image_to_u8": (c_int, [_P, c_int, _P, c_longlong, c_int, _P]),
    "tf_embedding_f16": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "tf_cast_f16_to_f32": (c_int, [_P, _P, c_longlong, _P]),
    "tf_timestep_embedding_f32": (c_int, [_P, _P, c_int, c_float, _P, _P]),
    "tf_gemv_f16w": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "tf_conv3x3_smallcin_f32nchw": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "tf_conv3x3_smallcin_gn_f32nchw": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "tf_upsample_nearest2x_nhwc_f16": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "tf_nchw_to_nhwc_f16": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "tf_nhwc_to_nchw": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "tf_pad_tokens_f32_to_f16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "tf_cfg_ddim_step_f32": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "tf_p2p_blocks": (c_int, [c_int, c_int]),
    "tf_p2p_alloc": (c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), _P]),
    "tf_p2p_open": (c_int, [_P, ctypes.POINTER(ctypes.c_void_p)]),
    "tf_p2p_close": (c_int, [_P]),
    "tf_p2p_free": (c_int, [_P]),
    "tf_cfg_ddim_step_split_f32": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                           c_int, _P]),
    "tf_add_int": (c_int, [_P, c_int, _P]),
    "tf_unary": (c_int, [_P, _P, c_longlong, c_int, c_int, _P]),
    "tf_nhwc_f32_to_nchw_f32": (c_int, [_P, c_int, _P, c_int, c_int, c_int, _P]),
    "tf_ddim_step_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_longlong, _P]),
    # fp32 parity mode (csrc/tf_fp32.cu)
    "tf_gemm_f32": (c_int, [_P, c_longlong, _P, c_longlong, c_int, _P, _P, c_longlong, _P, c_longlong, c_int, c_int, c_int,
                            c_float, c_int, c_int, c_int, _P, _P]),
    "tf_conv2d_nchw_f32": (c_int, [_P, _P, _P, _P, _P, _P] + [c_int] * 10 + [_P]),
    "tf_groupnorm_nchw_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "tf_layernorm_f32": (c_int, [_P, _P, _P, _P, c_longlong, c_int, c_float, _P]),
    "tf_softmax_rows_f32": (c_int, [_P, c_longlong, c_int, c_int, _P]),
    "tf_embedding_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "tf_unary_f32": (c_int, [_P, _P, c_longlong, c_int, _P]),
    "tf_geglu_f32": (c_int, [_P, c_longlong, _P, c_longlong, c_int, _P]),
    "tf_cfg_combine_f32": (c_int, [_P, _P, c_float, _P, c_longlong, _P]),
}


class B200:
    def __init__(self, path=_LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python -m tinyfusers_b200.csrc.build` "
                "(tinyfusers_b200 has no CPU / eager fallback)")
        self.path = path
        self.dll = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)
        self._initialised = False

    # -- helpers ---------------------------------------------------------------------------------
    def last_error(self):
        msg = self.dll.tf_last_error()
        return msg.decode() if msg else ""

    def check(self, status, name):
        if status != 0:
            raise RuntimeError(f"{name} failed with status {status}: {self.last_error()}")

    def init(self, device=0):
        """One process per GPU: the first call binds the library to `device`; a different device later raises."""
        device = int(device)
        if not self._initialised:
            self.check(self.tf_init(device), "tf_init")
            self._initialised = True
            self._device = device
            self.load_tuning()
        elif device != self._device:
            raise RuntimeError(f"tinyfusers_b200 serves one GPU per process (bound to cuda:{self._device}, asked for cuda:{device}); "
                               "launch one process per GPU (torchrun / example.sd1 --gpus N)")

    def load_tuning(self, path=None):
        """Measured GEMM / conv tile choices (tools/autotune_gemm.py, run on a B200). Optional: shapes without an entry
        use the library's cost model. TINYFUSERS_B200_TUNING=0 disables the table."""
        import json
        path = os.path.join(_HERE, "gemm_tuning.json") if path is None else path
        self.tf_gemm_tuning_clear()
        if os.environ.get("TINYFUSERS_B200_TUNING", "1") == "0" or not os.path.exists(path):
            return 0
        with open(path) as fh:
            entries = json.load(fh)["entries"]
        for e in entries:
            self.check(self.tf_gemm_tuning_add(*[int(v) for v in e[:8]]), "tf_gemm_tuning_add")
        return len(entries)


b200 = B200()
b200.TF_OK = 0
b200.TF_ERR_ARG = -1
b200.TF_ERR_UNSUPPORTED = -2
b200.TF_ERR_DEVICE = -3
b200.TF_EPI_NONE = 0
b200.TF_EPI_OUT_F32 = 1
b200.TF_EPI_GEGLU = 2
b200.TF_GEMM_W_STATIC = 4
