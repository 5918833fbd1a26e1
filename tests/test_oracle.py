"""The CPU oracle against (a) independent brute-force definitions, (b) the reference's own CPU
loops compiled into oracle/_ref, (c) the golden vectors produced by the reference's own CUDA
kernels on a B200 (tests/golden/, see tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from gspn_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _d2(a, b):  # float64 squared distances (n,m)
    return ((a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)) ** 2).sum(-1)


def test_fps_is_farthest_point_sampling(oracle):
    xyz = scenes.uniform_cube(2, 700, seed=3)
    idx = oracle.farthest_point_sample(64, xyz)
    assert idx.shape == (2, 64) and (idx[:, 0] == 0).all()
    for b in range(2):
        mind = np.full(700, np.inf)
        for j in range(1, 64):
            mind = np.minimum(mind, _d2(xyz[b], xyz[b, idx[b, j - 1]][None])[:, 0])
            # the chosen point attains the maximum min-distance (up to float rounding)
            assert mind[idx[b, j]] >= mind.max() * (1 - 1e-5)
        assert len(set(idx[b])) == 64


def test_fps_tie_break_is_lowest_kmod512_then_k(oracle):
    # all points identical -> every distance ties at 0 -> key (k mod 512, k) minimal = 0
    xyz = np.ones((1, 1500, 3), np.float32)
    assert (oracle.farthest_point_sample(5, xyz) == 0).all()
    # two farthest candidates at k=600 (600 mod 512 = 88) and k=100: the reference picks k=600
    xyz = np.zeros((1, 1500, 3), np.float32)
    xyz[0, 100] = [1, 0, 0]
    xyz[0, 600] = [-1, 0, 0]
    assert oracle.farthest_point_sample(2, xyz)[0, 1] == 600
    xyz[0, 600] = 0
    xyz[0, 612] = [-1, 0, 0]  # 612 mod 512 = 100, same lane as k=100 -> lower k wins
    assert oracle.farthest_point_sample(2, xyz)[0, 1] == 100


def test_ball_query_semantics(oracle):
    xyz = scenes.uniform_cube(2, 600, seed=5)
    q = xyz[:, :50]
    r, K = 0.2, 16
    idx, cnt = oracle.query_ball_point(r, K, xyz, q)
    for b in range(2):
        inside = np.sqrt(_d2(q[b], xyz[b])) < r
        for j in range(50):
            hits = np.nonzero(inside[j])[0]
            # borderline float cases aside (none for this seed), first-K-in-index-order + back-fill
            assert cnt[b, j] == min(K, len(hits))
            exp = list(hits[:K]) + [hits[0]] * (K - min(K, len(hits)))
            assert list(idx[b, j]) == exp


def test_ball_query_zero_hit_rows_are_zero(oracle):
    xyz = scenes.uniform_cube(1, 100, seed=1)
    q = np.full((1, 3, 3), 50.0, np.float32)
    idx, cnt = oracle.query_ball_point(0.1, 8, xyz, q)
    assert (idx == 0).all() and (cnt == 0).all()


def test_three_nn_and_nn_distance_definitions(oracle):
    rng = np.random.RandomState(0)
    a = rng.randn(2, 300, 3).astype(np.float32)
    b = rng.randn(2, 77, 3).astype(np.float32)
    dist, idx = oracle.three_nn(a, b)
    d1, i1, d2, i2 = oracle.nn_distance(a, b)
    for k in range(2):
        D = _d2(a[k], b[k])
        order = np.argsort(D, axis=1, kind="stable")[:, :3]
        assert (idx[k] == order).all()
        np.testing.assert_allclose(dist[k], np.take_along_axis(D, order, 1), rtol=1e-5)
        assert (i1[k] == D.argmin(1)).all() and (i2[k] == D.argmin(0)).all()
        np.testing.assert_allclose(d1[k], D.min(1), rtol=1e-5)


def test_three_nn_fewer_than_three_known(oracle):
    a = scenes.uniform_cube(1, 10, seed=2)
    dist, idx = oracle.three_nn(a, a[:, :2])
    assert np.isinf(dist[..., 2]).all() and (idx[..., 2] == 0).all()


@pytest.mark.parametrize("dup", [False, True])
def test_oracle_matches_reference_cpu_loops_bit_for_bit(oracle, dup):
    if oracle.ref_cpu() is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    xyz1 = scenes.uniform_cube(3, 1000, seed=11)
    xyz2 = scenes.uniform_cube(3, 257, seed=12)
    if dup:
        xyz1 = scenes.with_duplicates(xyz1)
        xyz2 = scenes.with_duplicates(xyz2)
        xyz1[:, :40] = xyz2[:, :40]
    d, i = oracle.three_nn(xyz1, xyz2)
    dr, ir = oracle.ref_three_nn(xyz1, xyz2)
    assert np.array_equal(i, ir) and np.array_equal(d.view(np.int32), dr.view(np.int32))
    pts = np.random.RandomState(1).randn(3, 257, 37).astype(np.float32)
    w = oracle.fp_weights(d)
    assert np.array_equal(oracle.three_interpolate(pts, i, w).view(np.int32), oracle.ref_three_interpolate(pts, i, w).view(np.int32))
    go = np.random.RandomState(2).randn(3, 1000, 37).astype(np.float32)
    assert np.array_equal(oracle.three_interpolate_grad(pts, i, w, go), oracle.ref_three_interpolate_grad(pts, i, w, go))
    for x, y in zip(oracle.nn_distance(xyz1, xyz2), oracle.ref_nn_distance(xyz1, xyz2)):
        assert np.array_equal(x.view(np.int32), y.view(np.int32))


def test_group_and_gather_and_grads(oracle):
    rng = np.random.RandomState(4)
    pts = rng.randn(2, 50, 5).astype(np.float32)
    idx = rng.randint(0, 50, size=(2, 7, 4)).astype(np.int32)
    g = oracle.group_point(pts, idx)
    assert np.array_equal(g, np.stack([pts[b][idx[b]] for b in range(2)]))
    go = rng.randn(2, 7, 4, 5).astype(np.float32)
    gp = oracle.group_point_grad(pts, idx, go)
    exp = np.zeros_like(pts)
    for b in range(2):
        np.add.at(exp[b], idx[b].reshape(-1), go[b].reshape(-1, 5))
    np.testing.assert_allclose(gp, exp, rtol=1e-5, atol=1e-6)
    i2 = rng.randint(0, 50, size=(2, 9)).astype(np.int32)
    assert np.array_equal(oracle.gather_point(pts, i2), np.stack([pts[b][i2[b]] for b in range(2)]))


def test_mlp_layer_and_pool(oracle):
    rng = np.random.RandomState(6)
    x = rng.randn(2, 5, 8, 7).astype(np.float32)
    layer = dict(weights=rng.randn(7, 11).astype(np.float32), biases=rng.randn(11).astype(np.float32),
                 gamma=rng.rand(11).astype(np.float32) + 0.5, beta=rng.randn(11).astype(np.float32),
                 moving_mean=rng.randn(11).astype(np.float32), moving_variance=rng.rand(11).astype(np.float32) + 0.1)
    y = oracle.mlp_layer(x, layer)
    z = x.astype(np.float64) @ layer["weights"] + layer["biases"]
    z = (z - layer["moving_mean"]) / np.sqrt(layer["moving_variance"].astype(np.float64) + 1e-3) * layer["gamma"] + layer["beta"]
    np.testing.assert_allclose(y, np.maximum(z, 0), rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(oracle.max_over_k(y), y.max(axis=2))


def _golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else []


@pytest.mark.parametrize("name", _golden_files() or ["<none>"])
def test_oracle_matches_reference_cuda_kernels_golden(oracle, name):
    """tests/golden/*.npz hold outputs of the reference's OWN .cu kernels (oracle/_ref/libref_gpu.so)
    run on a B200 by tests/golden/make_golden.py."""
    if name == "<none>":
        pytest.skip("no golden vectors committed yet")
    g = np.load(os.path.join(GOLDEN, name))
    kind = str(g["kind"])
    if kind == "fps":
        assert np.array_equal(oracle.farthest_point_sample(int(g["npoint"]), g["xyz"]), g["idx"])
    elif kind == "ball":
        idx, cnt = oracle.query_ball_point(float(g["radius"]), int(g["nsample"]), g["xyz"], g["new_xyz"])
        assert np.array_equal(cnt, g["cnt"])
        full = g["cnt"] > 0  # zero-hit rows are uninitialised memory in the reference
        assert np.array_equal(idx[full], g["idx"][full])
    elif kind == "nnd":
        d1, i1, d2, i2 = oracle.nn_distance(g["xyz1"], g["xyz2"], gpu_variant=True)
        assert np.array_equal(i1, g["i1"]) and np.array_equal(i2, g["i2"])
        assert np.array_equal(d1.view(np.int32), g["d1"].view(np.int32)) and np.array_equal(d2.view(np.int32), g["d2"].view(np.int32))
    else:
        raise AssertionError(kind)


def test_nearest_point_is_the_dense_argmin(oracle):
    """models/model_rpointnet.py:1136: argmin over reduce_sum(square(pc[:, :, None] - seed[:, None]), -1); ties -> first index."""
    rng = np.random.RandomState(11)
    pc = rng.rand(2, 300, 3).astype(np.float32)
    seed = rng.rand(2, 17, 3).astype(np.float32)
    seed[:, 5] = seed[:, 3]  # an exact tie: the lower index must win, as tf.argmin returns the first minimum
    pc[:, 7] = seed[:, 3]
    d, i = oracle.nearest_point(pc, seed)
    diff = pc[:, :, None, :] - seed[:, None, :, :]
    sq = diff * diff
    dense = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # float32, every operation rounded
    assert np.array_equal(i, dense.argmin(-1).astype(np.int32))
    assert np.array_equal(d, dense.min(-1))
    assert (i[:, 7] == 3).all()


def test_box_shrink_restatement(oracle):
    """models/model_rpointnet.py:529-551: tight boxes, empty boxes zeroed."""
    rng = np.random.RandomState(12)
    pc = rng.rand(2, 500, 3).astype(np.float32)
    box = np.concatenate([rng.rand(2, 6, 3), 0.2 + 0.4 * rng.rand(2, 6, 3)], -1).astype(np.float32)
    box[:, 0, :3] = 5.0  # far away: holds no point
    out = oracle.box_shrink(box, pc)
    assert out.shape == (2, 6, 6) and (out[:, 0] == 0).all()
    for b in range(2):
        for s in range(1, 6):
            lo, hi = box[b, s, :3] - box[b, s, 3:] / 2, box[b, s, :3] + box[b, s, 3:] / 2
            inside = pc[b][np.all((pc[b] >= lo) & (pc[b] <= hi), -1)]
            if len(inside) < 2:
                continue
            np.testing.assert_allclose(out[b, s, :3], (inside.max(0) + inside.min(0)) / 2, rtol=1e-6)
            np.testing.assert_allclose(out[b, s, 3:], inside.max(0) - inside.min(0) + 1e-3, rtol=1e-5)



def test_blas_mlp_layer_of_the_reference_arm_matches_the_checker(oracle):
    """bench.py's reference arm runs the shared MLP through a BLAS-class GEMM (TensorFlow-CPU stand-in); same layer, same numbers
    up to summation order."""
    rng = np.random.RandomState(5)
    x = rng.randn(3, 50, 8, 67).astype(np.float32)
    layer = {"weights": (rng.randn(67, 96) * 0.2).astype(np.float32), "biases": rng.randn(96).astype(np.float32) * 0.1,
             "gamma": (0.75 + 0.5 * rng.rand(96)).astype(np.float32), "beta": rng.randn(96).astype(np.float32) * 0.1,
             "moving_mean": rng.randn(96).astype(np.float32) * 0.1, "moving_variance": (0.75 + 0.5 * rng.rand(96)).astype(np.float32)}
    a, b = oracle.mlp_layer(x, layer), oracle.mlp_layer_blas(x, layer)
    assert a.shape == b.shape == (3, 50, 8, 96)
    np.testing.assert_allclose(b, a, rtol=1e-4, atol=1e-5)
    nobn = {"weights": layer["weights"], "biases": layer["biases"]}
    np.testing.assert_allclose(oracle.mlp_layer_blas(x, nobn, relu=False), oracle.mlp_layer(x, nobn, relu=False), rtol=1e-4, atol=1e-5)
