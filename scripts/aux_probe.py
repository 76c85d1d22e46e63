"""Config C4 (2M surfels, 2DGS + USE_AABB, 1080p): colour + depth + normal in one pass (bgs_render_aux) vs three frames."""
import os, sys, time, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
cloud = B.random_gaussians_3d_seeded(2_000_000, 4)
pl = B.GaussianSplattingPlugin(0); h = pl.add_cloud(cloud)
s = B.CloudSettings(global_scale=0.02, gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True); v = B.headless_view(1920, 1080)
def med(fn, reps=25):
    rows = []
    for _ in range(reps):
        fn(); rows.append(pl.stage_times_us())
    return np.median(np.array(rows[5:]), 0)
one = {}
for m in (B.RasterizeMode.Color, B.RasterizeMode.Depth, B.RasterizeMode.Normal):
    sm = dataclasses.replace(s, rasterize_mode=m)
    one[m.name] = med(lambda: pl.render_view(h, sm, v, fmt="rgba8_srgb", to_host=False))
import ctypes as C, torch
outs = [torch.empty(1920 * 1080 * 4, dtype=torch.uint8, device="cuda") for _ in range(3)]
vv, uu, ss = v.to_abi(), pl.cloud_uniform(s, None, h.aabb), s.to_abi()
aux = med(lambda: pl._check(pl._lib.bgs_render_aux(pl._ctx, h._h, C.byref(vv), C.byref(uu), C.byref(ss), C.c_void_p(outs[0].data_ptr()),
                                                   C.c_void_p(outs[1].data_ptr()), C.c_void_p(outs[2].data_ptr()), 0, 1)))
print("| frame | keygen | sort | project | bin | raster | total us |"); print("|---|---|---|---|---|---|---|")
for k, r in one.items(): print(f"| {k} only | " + " | ".join(f"{x:.0f}" for x in r) + " |")
print("| colour + depth + normal, one pass (bgs_render_aux) | " + " | ".join(f"{x:.0f}" for x in aux) + " |")
print(f"three frames {sum(r[5] for r in one.values()):.0f} us; one pass {aux[5]:.0f} us = {aux[5] / one['Color'][5]:.2f} x a colour frame")
