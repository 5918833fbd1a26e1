"""Generates tests/golden/*.npz: outputs of the reference's OWN CUDA kernels (oracle/_ref/libref_gpu.so,
built unmodified from /root/reference by oracle/Makefile) on seeded inputs, run on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden     # then copy *.npz into tests/golden/

The CPU test tests/test_oracle.py::test_oracle_matches_reference_cuda_kernels_golden pins the oracle
restatement to these vectors; they are what makes FPS / ball-query / NmDistance parity "pinned".
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refgpu  # noqa: E402
from gspn_b200 import scenes  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # FPS: n not a multiple of 512, > 3072 (smem/global split of the reference), duplicates -> exact ties
    for name, xyz, m in [
        ("fps_cube_3500", scenes.uniform_cube(2, 3500, seed=21), 96),
        ("fps_dups_1500", scenes.with_duplicates(scenes.uniform_cube(2, 1500, seed=22), 0.5), 128),
        ("fps_scene_4096", scenes.scannet_like_batch(0, 1, 4096)[0], 256),
        ("fps_tiny_100", scenes.uniform_cube(3, 100, seed=23), 100),
    ]:
        idx = refgpu.fps(m, t(xyz)).cpu().numpy()
        np.savez_compressed(os.path.join(out, name + ".npz"), kind="fps", xyz=xyz, npoint=m, idx=idx)
    # ball query: under-filled rows, duplicates, radius edge
    for name, xyz, nq, r, k in [
        ("ball_cube_1500", scenes.uniform_cube(2, 1500, seed=31), 64, 0.15, 16),
        ("ball_dups_1200", scenes.with_duplicates(scenes.uniform_cube(2, 1200, seed=32), 0.5), 48, 0.2, 32),
        ("ball_scene_4096", scenes.scannet_like_batch(3, 1, 4096)[0], 128, 0.4, 32),
    ]:
        q = np.ascontiguousarray(xyz[:, :nq])
        idx, cnt = refgpu.query_ball_point(r, k, t(xyz), t(q))
        np.savez_compressed(os.path.join(out, name + ".npz"), kind="ball", xyz=xyz, new_xyz=q, radius=r, nsample=k,
                            idx=idx.cpu().numpy(), cnt=cnt.cpu().numpy())
    # NmDistance (GPU rounding): tails not multiple of 512 / 4
    rng = np.random.RandomState(41)
    a = rng.randn(2, 701, 3).astype(np.float32)
    b = rng.randn(2, 1027, 3).astype(np.float32)
    b[:, :50] = a[:, :50]
    d1, i1, d2, i2 = [x.cpu().numpy() for x in refgpu.nn_distance(t(a), t(b))]
    np.savez_compressed(os.path.join(out, "nnd_randn_701_1027.npz"), kind="nnd", xyz1=a, xyz2=b, d1=d1, i1=i1, d2=d2, i2=i2)
    print("golden vectors written to", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
