"""Generates tests/golden/c1_small.npz from the CPU oracle (NOT from the reference: it cannot be
built or run here).  These fixtures freeze the restatement against regressions and
travel to the GPU box; they are oracle-generated, so they do not lift the "parity unpinned" status
of the stages past the sort key.   Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bevy_gaussian_splatting_b200 as B  # noqa: E402
from oracle import oracle as O  # noqa: E402

n, seed, w, h, scale = 1000, 0, 256, 256, 1.0   # config C1
cloud = B.random_gaussians_3d_seeded(n, seed)
s = B.CloudSettings(global_scale=scale)
view = B.headless_view(w, h)
u = B.GaussianSplattingPlugin.cloud_uniform(s)
keys = O.keygen(cloud.position_visibility, view.to_abi(), u, 32)
sk, si = O.radix_sort(keys, 32)
til = O.render_tiles(cloud, view.to_abi(), u, s.to_abi())
np.savez_compressed(os.path.join(os.path.dirname(__file__), "c1_small.npz"), n=n, seed=seed, w=w, h=h, scale=scale,
                    keys=keys, order=si, tile_ranges=til["tile_ranges"], image=til["image"][::4, ::4].astype(np.float32))
print("wrote c1_small.npz", til["n_vis"], til["n_pairs"])


def make_gcloud_fixture():
    """tests/golden/c64_seed5.gcloud: `encode_gcloud(random_gaussians_3d_seeded(64, 5))`, frozen so that a change of the
    writer's layout or of the reader shows up (tests/test_io_gcloud.py::test_committed_fixture)."""
    from bevy_gaussian_splatting_b200 import gcloud as G

    with open(os.path.join(os.path.dirname(__file__), "c64_seed5.gcloud"), "wb") as f:
        f.write(G.encode_gcloud(B.random_gaussians_3d_seeded(64, 5)))


make_gcloud_fixture()
print("wrote c64_seed5.gcloud")
