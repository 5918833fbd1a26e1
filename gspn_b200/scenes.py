"""Deterministic synthetic inputs for tests and bench.

ScanNet is not available offline, so scenes are generated: a room of U[4,10] x U[4,10] x
U[2.4,3.2] metres; half of the points on floor / ceiling / walls, half on the faces of 10-30
axis-aligned boxes ("furniture", 0.3-2 m); 5 mm Gaussian jitter; colour U[0,1)^3; random point
order; float32; metres, so the reference radii 0.2/0.4/0.8/1.6 (model_rpointnet.py:168-171)
mean what they mean on real scans.  Seed = 1000 + scene_id.
"""
import numpy as np


def _surface_points(rng, lo, hi, count):
    """count points uniformly on the 6 faces of the box [lo,hi], area-weighted."""
    ext = hi - lo
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1], ext[0] * ext[1]])
    face = rng.choice(6, size=count, p=areas / areas.sum())
    pts = lo + rng.random((count, 3)) * ext
    axis = face // 2
    side = face % 2
    pts[np.arange(count), axis] = np.where(side == 0, lo[axis], hi[axis])
    return pts


def scannet_like_scene(scene_id, npoints=32768):
    """-> xyz (npoints,3) float32 metres, colour (npoints,3) float32 in [0,1)."""
    rng = np.random.Generator(np.random.PCG64(1000 + int(scene_id)))
    room = np.array([rng.uniform(4, 10), rng.uniform(4, 10), rng.uniform(2.4, 3.2)])
    n_room = npoints // 2
    n_furn = npoints - n_room
    pts = [_surface_points(rng, np.zeros(3), room, n_room)]
    nbox = int(rng.integers(10, 31))
    sizes = rng.uniform(0.3, 2.0, size=(nbox, 3))
    sizes[:, 2] = np.minimum(sizes[:, 2], room[2] * 0.8)
    lo = rng.random((nbox, 3)) * np.maximum(room - sizes, 0.1)
    lo[:, 2] = 0.0
    share = sizes.prod(axis=1) ** (2.0 / 3.0)
    counts = np.floor(share / share.sum() * n_furn).astype(int)
    counts[0] += n_furn - counts.sum()
    for i in range(nbox):
        if counts[i] > 0:
            pts.append(_surface_points(rng, lo[i], lo[i] + sizes[i], counts[i]))
    xyz = np.concatenate(pts, axis=0)
    xyz += rng.normal(0.0, 0.005, size=xyz.shape)
    xyz = xyz[rng.permutation(npoints)]
    colour = rng.random((npoints, 3))
    return xyz.astype(np.float32), colour.astype(np.float32)


def scannet_like_batch(first_scene, batch, npoints=32768):
    xs, cs = zip(*(scannet_like_scene(first_scene + i, npoints) for i in range(batch)))
    return np.stack(xs), np.stack(cs)


def uniform_cube(batch, npoints, seed=100, channels=0):
    """np.random.random clouds as in the reference's own op tests (tf_grouping_op_test.py:11-16)."""
    rng = np.random.RandomState(seed)
    xyz = rng.random_sample((batch, npoints, 3)).astype(np.float32)
    if channels:
        return xyz, rng.random_sample((batch, npoints, channels)).astype(np.float32)
    return xyz


def with_duplicates(xyz, frac=0.25, seed=7):
    """The reference dataset pads clouds by repeating random points (dataset.py:100-105,116-118):
    overwrite a fraction of each cloud with copies of other points -> exact zero-distance ties."""
    rng = np.random.RandomState(seed)
    out = xyz.copy()
    b, n, _ = xyz.shape
    k = int(n * frac)
    for i in range(b):
        dst = rng.choice(n, k, replace=False)
        src = rng.choice(n, k, replace=True)
        out[i, dst] = xyz[i, src]
    return out


def shard_scenes(total_scenes, rank, world_size):
    """Contiguous scene range [lo,hi) owned by `rank` (scenes are independent, so
    multi-GPU is batch sharding with no data-path collective)."""
    base, rem = divmod(total_scenes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
