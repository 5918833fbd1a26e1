"""The C-ABI library loads and exports exactly what include/gspn_b200.h declares (no compute)."""
import ctypes
import os
import re

from gspn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gspn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gspn_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 20
    dll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_strings():
    L = _lib.lib()
    assert L.gspn_version() >= 1000
    assert L.gspn_error_string(0) == b"ok"
    for code in range(-6, 0):
        assert L.gspn_error_string(code) != b"unknown error code"


def test_argument_checks_need_no_gpu():
    """Shape/attr validation happens before any CUDA call (OP_REQUIRES analogue)."""
    L = _lib.lib()
    assert L.gspn_farthest_point_sample(1, 0, 4, None, None, None, 0, None) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_farthest_point_sample(1, 8, 4, None, None, None, 0, None) == _lib.GSPN_E_NULL_PTR
    assert L.gspn_query_ball_point(1, 8, 4, -1.0, 4, None, None, None, None, None, 0, None) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_query_ball_point(1, 8, 4, 0.5, 0, None, None, None, None, None, 0, None) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_nn_distance(1, 8, 8, None, None, None, None, None, None, 7, None, 0, None) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_grouped_bytes(256, 6, _lib.GSPN_DT_BF16) == 2 * 16384
    assert L.gspn_grouped_bytes(129, 67, _lib.GSPN_DT_BF16) == 2 * 2 * 16384


def test_glue_entry_points_validate_without_a_gpu():
    """gspn_nearest_point / gspn_box_shrink / the query workspace size: argument checks come before any CUDA call."""
    L = _lib.lib()
    assert L.gspn_nearest_point(1, 8, 0, None, None, None, None, 0, None, 0, None) == _lib.GSPN_E_BAD_SHAPE     # no reference points
    assert L.gspn_nearest_point(1, 8, 4, None, None, None, None, 2, None, 0, None) == _lib.GSPN_E_BAD_SHAPE     # rounding in {0,1}
    assert L.gspn_nearest_point(1, 8, 4, None, None, None, None, 0, None, 0, None) == _lib.GSPN_E_NULL_PTR
    assert L.gspn_nearest_point(0, 8, 4, None, None, None, None, 0, None, 0, None) == 0                          # empty batch: nothing to do
    assert L.gspn_box_shrink(1, 4, 0, None, None, None, None) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_box_shrink(1, 4, 16, None, None, None, None) == _lib.GSPN_E_NULL_PTR
    assert L.gspn_box_shrink(1, 0, 16, None, None, None, None) == 0
    scanned = L.gspn_grid_workspace_bytes(2, 5000)
    assert L.gspn_grid_query_workspace_bytes(2, 1000, 5000) == scanned            # small query set: the scanned set's grid only
    assert L.gspn_grid_query_workspace_bytes(2, 70000, 5000) >= scanned + L.gspn_grid_workspace_bytes(2, 70000)
    assert L.gspn_grid_query_workspace_bytes(0, 70000, 5000) == 0


def test_round2_entry_points_validate_without_a_gpu():
    """The chain's arithmetic / image dtypes, the fused multi-radius query, the deterministic backward and the SyncBN split:
    argument checks come before any CUDA call; the library reads no environment variables."""
    import ctypes
    L = _lib.lib()
    assert L.gspn_mlp_weight_image_bytes(64, 32, _lib.GSPN_MLP_BF16) == 32 * 128
    assert L.gspn_mlp_weight_image_bytes(64, 32, _lib.GSPN_MLP_BF16X3) == 2 * 32 * 128   # [hi rows | lo rows]
    assert L.gspn_mlp_weight_image_bytes(64, 32, 7) == 0
    assert L.gspn_grouped_bytes(256, 6, _lib.GSPN_DT_BF16X2) == 2 * 2 * 16384               # every block a [hi | lo] pair
    dims = (ctypes.c_int * 2)(64, 32)
    one = (ctypes.c_void_p * 1)(None)
    rel = (ctypes.c_int * 1)(1)
    c = lambda a: ctypes.cast(a, ctypes.c_void_p)
    # unknown arithmetic -> bad dtype; rows == 0 -> nothing to do; missing image -> null pointer
    assert L.gspn_mlp_chain(128, 1, c(dims), 0, None, c(one), c(one), c(one), c(rel), 1, None, None, _lib.GSPN_DT_BF16, 9, None) in (
        _lib.GSPN_E_NULL_PTR, _lib.GSPN_E_BAD_DTYPE)
    assert L.gspn_mlp_chain(0, 1, c(dims), 0, None, c(one), c(one), c(one), c(rel), 1, None, None, _lib.GSPN_DT_BF16, _lib.GSPN_MLP_BF16X3, None) == 0
    assert L.gspn_fp_assemble(1, 8, 4, 3, 5, None, None, None, None, None, 64, 9, None) == _lib.GSPN_E_BAD_DTYPE
    assert L.gspn_fp_assemble(1, 8, 4, 0, 0, None, None, None, None, None, 64, _lib.GSPN_DT_BF16X2, None) == _lib.GSPN_E_BAD_SHAPE
    rad = (ctypes.c_float * 5)(0.5, 1.0, 1.5, 2.0, 3.0)
    ns = (ctypes.c_int * 5)(8, 8, 8, 8, 8)
    ptrs = (ctypes.c_void_p * 5)(*([None] * 5))
    assert L.gspn_query_ball_point_multi(1, 8, 4, 5, c(rad), c(ns), None, None, c(ptrs), c(ptrs), None) == _lib.GSPN_E_BAD_SHAPE  # > 4 balls
    assert L.gspn_query_ball_point_multi(1, 8, 4, 2, c(rad), c(ns), None, None, c(ptrs), c(ptrs), None) == _lib.GSPN_E_NULL_PTR
    assert L.gspn_query_ball_point_multi(0, 8, 4, 2, c(rad), c(ns), None, None, c(ptrs), c(ptrs), None) == 0
    assert L.gspn_scatter_det_workspace_bytes(2, 100, 16) == 256 + 8 * 2 * 100 * 16
    assert L.gspn_group_point_grad_det(1, 8, 4, 2, 2, None, None, None, None, 0, None) == _lib.GSPN_E_NULL_PTR
    assert L.gspn_bn_bwd_apply_f32(8, 4, 4, 1, 1, None, None, None, None, None, None, None, None, None, None, None) == _lib.GSPN_E_BAD_SHAPE  # total < rows
    # the curve-ordered copy of the pruned FPS kernels: both opt-in, none for the full scan
    assert L.gspn_farthest_point_sample_workspace_bytes(8, 32768, 2048) == 0
    assert L.gspn_fps_pruned_profile(1, 8192, 300, None, None, None, 0, None, None) == _lib.GSPN_E_NULL_PTR
    try:
        for mode in (1, 2):
            L.gspn_fps_tune(mode)
            assert L.gspn_farthest_point_sample_workspace_bytes(8, 32768, 2048) == 8 * 32768 * 16
            assert L.gspn_farthest_point_sample_workspace_bytes(8, 8192, 2048) == 0
    finally:
        L.gspn_fps_tune(0)
    for path in ("mlp_tc.cu", "fps.cu", "fps_bucket.cu", "fps_pruned.cu", "ballquery_group.cu", "grid_search.cu", "gather_ops.cu", "nn_search.cu"):
        assert "getenv" not in open(os.path.join(ROOT, "gspn_b200", "csrc", path)).read(), path


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import gspn_b200
    with pytest.raises(RuntimeError, match="no CPU path"):
        gspn_b200.farthest_point_sample(4, torch.zeros(1, 8, 3))
    with pytest.raises(ValueError, match="FarthestPointSample expects positive npoint"):
        gspn_b200.farthest_point_sample(0, torch.zeros(1, 8, 3))
    with pytest.raises(ValueError, match="QueryBallPoint expects positive radius"):
        gspn_b200.query_ball_point(0.0, 4, torch.zeros(1, 8, 3), torch.zeros(1, 2, 3))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU / PyTorch fallback: without the CUDA library the product path raises instead of computing elsewhere."""
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libgspn_b200.so"))
    with pytest.raises(_lib.GspnError, match="no CPU/PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gspn_b200/ may import it."""
    import glob
    for path in glob.glob(os.path.join(ROOT, "gspn_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "import oracle" not in src and "from oracle" not in src, path
    for path in glob.glob(os.path.join(ROOT, "gspn_b200", "csrc", "*")):
        if os.path.isfile(path) and not path.endswith(".so"):
            assert "oracle" not in open(path, errors="ignore").read().lower() or path.endswith(".log"), path


def test_chain_planner_fits_the_sm_for_every_model_shape():
    """gspn_mlp_chain_plan: what the launcher would do, without a device.  For every chain of the model (SA1-4, FP1-4 incl. the commuted
    fa_layer4 and its pre-multiplication, config 3's three balls) and both arithmetics: tensor memory, shared memory and the register
    file hold the planned number of co-resident CTAs, and the rings are deep enough to be rings."""
    L = _lib.lib()
    c = lambda a: ctypes.cast(a, ctypes.c_void_p)

    def plan(mode, rows, dims, k0, pool, want_h, arith):
        d, out = (ctypes.c_int * len(dims))(*dims), (ctypes.c_int * 12)()
        rc = L.gspn_mlp_chain_plan(mode, rows, len(dims) - 1, c(d), k0, pool, 1, want_h, arith, c(out))
        assert rc == 0, (rc, dims)
        return dict(zip(("occ", "epi", "threads", "nch", "dcols", "bufs", "tmem", "a_col", "a_stages", "w_stages", "smem", "passes"), out))

    shapes = [("sa1", 1, 524288, [64, 32, 32, 64], 6, 32, 0), ("sa2", 0, 131072, [128, 64, 64, 128], 67, 32, 0),
              ("sa3", 0, 32768, [192, 128, 128, 256], 131, 32, 0), ("sa4", 0, 8192, [320, 256, 256, 512], 259, 32, 0),
              ("fp1", 0, 1024, [768, 256, 256], 768, 1, 0), ("fp2", 0, 4096, [384, 256, 256], 384, 1, 0),
              ("fp3", 0, 16384, [320, 256, 128], 320, 1, 0), ("fp4", 2, 262144, [128, 128, 128], 0, 1, 1),
              ("fp4_y2", 0, 16384, [128, 128], 128, 1, 0), ("cfg3_r0.5", 1, 16 * 128 * 256, [64, 64, 128, 256], 6, 256, 0),
              ("cfg3_r1.5", 1, 16 * 128 * 512, [64, 64, 128, 256], 6, 512, 0)]
    for name, mode, rows, dims, k0, pool, want_h in shapes:
        for arith in (_lib.GSPN_MLP_BF16, _lib.GSPN_MLP_BF16X3):
            p = plan(mode, rows, dims, k0, pool, want_h, arith)
            split = 2 if arith == _lib.GSPN_MLP_BF16X3 else 1
            mid = max(dims[1:-1]) if len(dims) > 2 else 0
            assert p["occ"] in (1, 2) and p["occ"] * p["tmem"] <= 512, (name, p)               # tensor memory: 512 columns per SM
            assert p["tmem"] & (p["tmem"] - 1) == 0 and p["tmem"] >= 32                           # allocations are powers of two
            assert p["a_col"] == p["bufs"] * p["dcols"] and p["a_col"] + split * mid // 2 <= p["tmem"], (name, p)
            assert p["dcols"] >= mid and p["dcols"] % p["nch"] == 0                               # a middle layer never runs in passes
            assert p["passes"] == -(-dims[-1] // p["dcols"])
            assert p["occ"] * (p["smem"] + 1024) <= 228 * 1024 and p["smem"] <= 226 * 1024, (name, p)
            assert (p["occ"] + 1) * (p["smem"] + 1024) > 228 * 1024, (name, p)                  # ... and no more CTAs than planned
            assert p["threads"] == (p["epi"] + {0: 1, 1: 2, 2: 8}[mode] + 2) * 32
            assert p["occ"] * p["threads"] * (128 if p["threads"] != 352 else 168) <= 65536, (name, p)  # registers (launch bounds)
            assert p["w_stages"] >= 2 and p["a_stages"] >= (2 if mode == 2 else 1)
    # nothing the kernel cannot run is planned: widths must be multiples of 32 up to 512, at most four layers
    d, out = (ctypes.c_int * 3)(64, 48, 520), (ctypes.c_int * 12)()
    assert L.gspn_mlp_chain_plan(0, 128, 2, c(d), 0, 1, 1, 0, _lib.GSPN_MLP_BF16X3, c(out)) == _lib.GSPN_E_UNSUPPORTED
    assert L.gspn_mlp_chain_plan(3, 128, 2, c(d), 0, 1, 1, 0, _lib.GSPN_MLP_BF16X3, c(out)) == _lib.GSPN_E_BAD_SHAPE
    assert L.gspn_mlp_chain_plan(0, 0, 2, c(d), 0, 1, 1, 0, _lib.GSPN_MLP_BF16X3, c(out)) == _lib.GSPN_E_BAD_SHAPE
