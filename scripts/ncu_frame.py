"""A few synchronous C3 frames for ncu (one context, one frame at a time): 2 upload kernels, then 6 kernels per frame
(key-gen, depth sort, projection, binning, pair sort + ranges, blend)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevy_gaussian_splatting_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cloud = B.random_gaussians_3d_seeded(6_000_000, 0)
pl = B.GaussianSplattingPlugin(0); h = pl.add_cloud(cloud, f16=True)
s = B.CloudSettings(global_scale=0.02); v = B.headless_view(1920, 1080)
for _ in range(n): pl.render_view(h, s, v, fmt="rgba8_srgb", to_host=False)
print(pl.stage_times_us(), pl.last_launch_count)
