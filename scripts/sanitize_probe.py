"""Small frames through every kernel variant, for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B

pl = B.GaussianSplattingPlugin(0)
cloud = B.random_gaussians_3d_seeded(6000, 1)
view = B.headless_view(208, 120)
n = 0
for f16 in (False, True):
    hd = pl.add_cloud(cloud, f16=f16)
    for kw in (dict(), dict(sort_all=True), dict(aabb=True), dict(gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True),
               dict(gaussian_mode=B.GaussianMode.Gaussian2d), dict(rasterize_mode=B.RasterizeMode.Depth),
               dict(rasterize_mode=B.RasterizeMode.Normal), dict(rasterize_mode=B.RasterizeMode.Position),
               dict(radix_sort_depth_bits=B.RadixSortDepthBits.Bits16), dict(binning_rounds=True),
               dict(binning_rounds=True, global_scale=1.0), dict(global_scale=1.0)):
        s = B.CloudSettings(**{"global_scale": 0.3, **kw})
        for fmt in ("rgba32f", "rgba8_srgb"):
            img = pl.render_view(hd, s, view, fmt=fmt)
            img = pl.render_view(hd, s, view, fmt=fmt)      # hinted frame (other kernel variants)
            n += 2
    # round-2 paths: compositing output modes, aux frames
    s = B.CloudSettings(global_scale=0.3)
    pl.render_view(hd, s, view, fmt="rgba32f")
    pl.render_view(hd, s, view, fmt="rgba32f", blend_over=True)
    pl.render_view(hd, s, view, fmt="rgba8_srgb", premultiplied=True)
    pl.render_view_aux(hd, s, view, fmt="rgba32f")
    pl.render_view_aux(hd, B.CloudSettings(global_scale=0.3, gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True), view, fmt="rgba8_srgb")
    n += 5
    out = np.empty((120, 208, 4), np.float32)
    for _ in range(3):
        pl.render_view(hd, B.CloudSettings(global_scale=0.3), view, fmt="rgba32f", out=out, asynchronous=True)
    pl.sync()
    hd.destroy()
hc = pl.add_cloud(cloud, precompute_covariance=True)
pl.render_view(hc, B.CloudSettings(), view, fmt="rgba32f"); n += 1
hc.destroy()
print("sanitize probe ok:", n, "frames")
