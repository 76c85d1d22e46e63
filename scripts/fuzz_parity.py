"""Randomised CUDA-vs-oracle sweep (run on the GPU box; not part of the pytest suite)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
from oracle import oracle as O

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
pl = B.GaussianSplattingPlugin(0)
t0 = time.time(); it = 0; worst = 0.0
while time.time() - t0 < budget:
    it += 1
    # (the larger counts reach the sort's bigger tile variants: ITEMS 2..16 per thread, one and two CTAs per SM)
    n = int(rng.choice([1, 2, 31, 33, 1000, 4096, 4097, 20000, 50000, 300000, 1200000, 2500000], p=[.08, .05, .05, .05, .12, .1, .1, .15, .12, .08, .06, .04]))
    w, h = int(rng.integers(8, 700)), int(rng.integers(8, 400))
    scale = float(10 ** rng.uniform(-2, 0.5)) if n <= 50000 else float(10 ** rng.uniform(-2.2, -1.3))
    gm = B.GaussianMode(int(rng.integers(0, 2))); aabb = bool(rng.integers(0, 2))
    rm = B.RasterizeMode(int(rng.integers(0, 4))); dm = B.DrawMode(int(rng.integers(0, 3)))
    bits = B.RadixSortDepthBits(int(rng.choice([16, 24, 32])))
    f16 = bool(rng.integers(0, 2)); sort_all = bool(rng.integers(0, 4) == 0)
    s = B.CloudSettings(global_scale=scale, gaussian_mode=gm, aabb=aabb, rasterize_mode=rm, draw_mode=dm, radix_sort_depth_bits=bits,
                        opacity_adaptive_radius=bool(rng.integers(0, 2)), global_opacity=float(rng.uniform(0.3, 2.0)),
                        color_space=B.GaussianColorSpace(int(rng.integers(0, 2))), sort_all=sort_all)
    cloud = B.random_gaussians_3d_seeded(n, int(rng.integers(0, 1000)))
    cloud.position_visibility[:, 3] = (rng.random(n) > 0.3).astype(np.float32)
    eye = rng.uniform(-8, 8, 3); tgt = rng.uniform(-3, 3, 3)
    view = B.perspective_view(tuple(eye), tuple(tgt), w, h, fov_y=float(rng.uniform(0.4, 1.4)))
    m = np.eye(4, dtype=np.float32)
    if rng.integers(0, 2):
        a = rng.uniform(0, 6.28); sc = rng.uniform(0.5, 2.0)
        m[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32) * np.float32(sc)
        m[:3, 3] = rng.uniform(-2, 2, 3)
    tr = B.CloudTransform(m)
    hd = pl.add_cloud(cloud, f16=f16)
    desc = f"it={it} n={n} {w}x{h} scale={scale:.3g} {gm.name} aabb={aabb} {rm.name} {dm.name} bits={int(bits)} f16={f16} sort_all={sort_all}"
    try:
        img = pl.render_view(hd, s, view, transform=tr)
        oc = cloud.rounded_to_f16() if f16 else cloud
        u = pl.cloud_uniform(s, tr, hd.aabb)
        keys = O.keygen(oc.position_visibility, view.to_abi(), u, int(bits))
        sk, si = O.radix_sort(keys, int(bits))
        got = pl.sorted_entries()
        assert np.array_equal(got[:, 0], sk) and np.array_equal(got[:, 1], si), "sort " + desc
        til = O.render_tiles(oc, view.to_abi(), u, s.to_abi())
        assert np.array_equal(pl.tile_ranges(), til["tile_ranges"]), "ranges " + desc
        assert np.array_equal(pl.tile_entries(), til["tile_entries"]), "entries " + desc
        fin = np.isfinite(til["image"])
        err = float(np.abs(img - til["image"])[fin].max()) if fin.any() else 0.0
        assert np.array_equal(np.isfinite(img), fin), "nan pattern " + desc
        worst = max(worst, err)
        assert err <= 1e-3, f"pixels {err} " + desc
        img2 = pl.render_view(hd, s, view, transform=tr)     # hinted second frame
        assert np.array_equal(np.isfinite(img2), fin) and float(np.abs(img2 - til["image"])[fin].max() if fin.any() else 0) <= 1e-3, "hinted " + desc
        if not aabb:   # front-to-back binning rounds (quad-uv records): the one-round frame, bit for bit
            import dataclasses
            img3 = pl.render_view(hd, dataclasses.replace(s, binning_rounds=True), view, transform=tr)
            assert pl.frame_stats().rounds > 1, "rounds " + desc
            assert np.array_equal(img3.view(np.uint32), img.view(np.uint32)), "rounds differ " + desc
    finally:
        hd.destroy()
print(f"fuzz ok: {it} random configurations, worst pixel L-inf {worst:.2e}")
