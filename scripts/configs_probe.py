"""Stage timings for the other BASELINE.json configs (C1, C2, C4) -- informational, not bench lines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B

def run(name, n, f16, w, h, frames=30, **kw):
    cloud = B.random_gaussians_3d_seeded(n, 0)
    pl = B.GaussianSplattingPlugin(0); hd = pl.add_cloud(cloud, f16=f16)
    s = B.CloudSettings(**kw); v = B.headless_view(w, h)
    rows = []
    for _ in range(frames):
        pl.render_view(hd, s, v, fmt="rgba8_srgb", to_host=False); rows.append(pl.stage_times_us())
    med = np.median(np.array(rows[5:]), 0); fs = pl.frame_stats()
    print(f"| {name} | {n} | {'f16' if f16 else 'f32'} | {w}x{h} | {fs.n_visible} | {fs.n_pairs} | "
          + " | ".join(f"{x:.0f}" for x in med) + f" | {n/med[5]:.0f} |", flush=True)
    hd.destroy(); pl.destroy()

print("| config | N | layout | frame | n_vis | pairs | keygen us | sort us | project us | bin us | raster us | frame us | Msplats/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
run("C1 1k @256^2", 1000, False, 256, 256)
run("C2 1M f32, global_scale 1 (raw generator)", 1_000_000, False, 1920, 1080, frames=12)
run("C2 1M f32, global_scale 0.02", 1_000_000, False, 1920, 1080, global_scale=0.02)
run("C3 6M f16, global_scale 0.02 (bench)", 6_000_000, True, 1920, 1080, global_scale=0.02)
G2 = dict(gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True, global_scale=0.02)
run("C4 2M surfels (2DGS, aabb) colour", 2_000_000, False, 1920, 1080, **G2)
run("C4 2M surfels depth", 2_000_000, False, 1920, 1080, rasterize_mode=B.RasterizeMode.Depth, **G2)
run("C4 2M surfels normal", 2_000_000, False, 1920, 1080, rasterize_mode=B.RasterizeMode.Normal, **G2)
