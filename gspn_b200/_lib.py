"""ctypes binding of libgspn_b200.so (include/gspn_b200.h).

The product has no CPU or PyTorch fallback: if the CUDA library is missing, or an
entry point returns an error code, the call raises.  `build()` compiles the library
in-tree for sm_100a with nvcc (cross-compiles without a GPU).
"""
import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_long, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgspn_b200.so")

GSPN_DT_F32 = 0
GSPN_DT_BF16 = 1
GSPN_DT_BF16X2 = 2  # split (hi | lo) tile image of the bf16x3 arithmetic
GSPN_DT_F16 = 3
GSPN_MLP_BF16 = 0
GSPN_MLP_BF16X3 = 1

GSPN_E_BAD_SHAPE, GSPN_E_NULL_PTR, GSPN_E_BAD_DTYPE, GSPN_E_WORKSPACE, GSPN_E_CUDA, GSPN_E_UNSUPPORTED = -1, -2, -3, -4, -5, -6

P = c_void_p  # every tensor / stream argument is a raw address

# name -> (restype, argtypes); mirrors include/gspn_b200.h declaration by declaration
SIGNATURES = {
    "gspn_error_string": (ctypes.c_char_p, [c_int]),
    "gspn_version": (c_int, []),
    "gspn_last_cuda_error": (ctypes.c_char_p, []),
    "gspn_fp32_peak_probe": (c_int, [c_int, c_int, P, P, P]),
    "gspn_farthest_point_sample_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gspn_fps_max_resident_points": (c_int, []),
    "gspn_farthest_point_sample": (c_int, [c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "gspn_farthest_point_sample_cfg": (c_int, [c_int, c_int, c_int, P, P, c_int, c_int, c_int, P]),
    "gspn_fps_tune": (None, [c_int]),
    "gspn_fps_tune_mapping": (None, [c_int, c_int, c_int]),
    "gspn_fps_tune_pack": (None, [c_int]),
    "gspn_fps_bucket_profile": (c_int, [c_int, c_int, c_int, P, P, P, c_size_t, P, P]),
    "gspn_fps_pruned_profile": (c_int, [c_int, c_int, c_int, P, P, P, c_size_t, P, P]),
    "gspn_fps_profile": (c_int, [c_int, c_int, c_int, P, P, c_int, c_int, c_int, P, P]),
    "gspn_gather_point": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P]),
    "gspn_gather_point_grad": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P]),
    "gspn_grid_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gspn_grid_query_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gspn_query_ball_point": (c_int, [c_int, c_int, c_int, c_float, c_int, P, P, P, P, P, c_size_t, P]),
    "gspn_query_ball_point_multi": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "gspn_ballquery_tune": (None, [c_int, c_int]),
    "gspn_group_point": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "gspn_group_point_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "gspn_grouped_bytes": (c_size_t, [c_long, c_int, c_int]),
    "gspn_ballquery_group": (c_int, [c_int, c_int, c_int, c_int, c_float, c_int, P, P, P, P, c_int, P, P, P, c_int, c_int, P, c_size_t, P]),
    "gspn_three_nn": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, c_size_t, P]),
    "gspn_three_interpolate": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "gspn_three_interpolate_grad": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "gspn_scatter_det_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gspn_gather_point_grad_det": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    "gspn_group_point_grad_det": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    "gspn_three_interpolate_grad_det": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, c_size_t, P]),
    "gspn_nn_distance": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, c_int, P, c_size_t, P]),
    "gspn_nn_distance_grad": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, P]),
    "gspn_mlp_layer_f32": (c_int, [c_long, c_int, c_int, P, c_int, P, P, P, c_int, c_int, P, P]),
    "gspn_max_pool_rows": (c_int, [c_long, c_int, c_int, P, P, P]),
    "gspn_mlp_weight_image_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gspn_mlp_pack_weights": (c_int, [c_int, c_int, c_int, P, P, P, c_int, P]),
    "gspn_mlp_chain": (c_int, [c_long, c_int, P, c_int, P, P, P, P, P, c_int, P, P, c_int, c_int, P]),
    "gspn_mlp_chain_gather": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, P, P, P, P, P, c_int, P, P, c_int, c_int, P]),
    "gspn_mlp_chain_fp": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, P, P, P, P, P, P, P, c_int, c_int, P]),
    "gspn_mlp_chain_set_profile": (None, [P]),
    "gspn_mlp_chain_tune": (None, [c_int, c_int, c_int]),
    "gspn_mlp_chain_tune_fp": (None, [c_int]),
    "gspn_mlp_chain_plan": (c_int, [c_int, c_long, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "gspn_mlp_chain_tune_sched": (None, [c_int]),
    "gspn_col_moments_f32": (c_int, [c_long, c_int, P, P, P, P]),
    "gspn_bn_act_f32": (c_int, [c_long, c_int, P, P, P, P, P, c_int, P, P]),
    "gspn_maxpool_argmax_f32": (c_int, [c_long, c_int, c_int, P, P, P, P]),
    "gspn_bn_act_pool_bwd_f32": (c_int, [c_long, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    "gspn_bn_bwd_sums_f32": (c_int, [c_long, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P]),
    "gspn_bn_bwd_apply_f32": (c_int, [c_long, c_long, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P]),
    "gspn_mlp_wgrad_f32": (c_int, [c_long, c_int, c_int, P, c_int, P, P, P, P]),
    "gspn_group_rows_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "gspn_fp_assemble": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, c_int, P]),
    "gspn_p2p_mailbox_bytes": (c_size_t, [c_int, c_int]),
    "gspn_p2p_mailbox_create": (c_int, [c_int, c_int, P, P]),
    "gspn_p2p_mailbox_open": (c_int, [P, P]),
    "gspn_p2p_mailbox_close": (c_int, [P]),
    "gspn_p2p_mailbox_destroy": (c_int, [P]),
    "gspn_p2p_allreduce_f64": (c_int, [c_int, c_int, c_int, P, c_int, P, P]),
    "gspn_nearest_point": (c_int, [c_int, c_int, c_int, P, P, P, P, c_int, P, c_size_t, P]),
    "gspn_box_shrink": (c_int, [c_int, c_int, c_int, P, P, P, P]),
}


class GspnError(RuntimeError):
    pass


def build(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> gspn_b200/libgspn_b200.so"""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise GspnError("building libgspn_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None
CALLS = [0]  # C-ABI compute calls that returned GSPN_OK (each enqueues >= 1 kernel); bench.py reports the delta


def lib():
    """The loaded library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GspnError(
                "libgspn_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gspn_b200/csrc`; there is no CPU/PyTorch fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code, what):
    """Map a GSPN_E_* return code to the exception the reference raises for the same mistake
    (errors::InvalidArgument -> ValueError; anything CUDA -> RuntimeError)."""
    if code == 0:
        CALLS[0] += 1
        return
    l = lib()
    msg = l.gspn_error_string(code).decode()
    if code == GSPN_E_CUDA:
        raise GspnError("%s: %s: %s" % (what, msg, l.gspn_last_cuda_error().decode()))
    if code in (GSPN_E_BAD_SHAPE, GSPN_E_NULL_PTR, GSPN_E_BAD_DTYPE):
        raise ValueError("%s: %s" % (what, msg))
    raise GspnError("%s: %s" % (what, msg))
