"""One multi-round frame for an ncu launch list: python scripts/rounds_launches.py [n] [f16] (frames: 2 warm + 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevy_gaussian_splatting_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
f16 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
pl = B.GaussianSplattingPlugin(0); hd = pl.add_cloud(B.random_gaussians_3d_seeded(n, 0), f16=f16)
s = B.CloudSettings(global_scale=1.0, binning_rounds=True); v = B.headless_view(1920, 1080)
for i in range(3):
    pl.render_view(hd, s, v, fmt="rgba8_srgb", to_host=False)
print(pl.frame_stats().n_pairs, pl.stage_times_us())
