"""Heavy-footprint probe: C2 / C3 clouds at global_scale 1.0 (hundreds of tiles per splat), one-round frames vs
front-to-back binning rounds.  Usage: python scripts/raw6m.py [n] [f16:0|1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
f16 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cloud = B.random_gaussians_3d_seeded(n, 0)
pl = B.GaussianSplattingPlugin(0); hd = pl.add_cloud(cloud, f16=f16)
v = B.headless_view(1920, 1080)
imgs = {}
for name, rounds in (("rounds", True), ("auto", None), ("one", False)):
    s = B.CloudSettings(global_scale=scale, binning_rounds=rounds)
    for i in range(4):
        t0 = time.perf_counter(); pl.render_view(hd, s, v, fmt="rgba8_srgb", to_host=False); dt = time.perf_counter() - t0
        st = pl.stage_times_us(); fs = pl.frame_stats()
        print(f"{name} frame {i}: wall {dt*1e3:.2f} ms rounds={fs.rounds} sat={fs.tiles_saturated}/{fs.tiles_x*fs.tiles_y} "
              f"n_vis={fs.n_visible} pairs={fs.n_pairs} stage_us={st.round(0).tolist()}", flush=True)
    imgs[name] = pl.render_view(hd, s, v, fmt="rgba32f")
print("identical:", np.array_equal(imgs["rounds"], imgs["one"]), np.array_equal(imgs["auto"], imgs["one"]))
