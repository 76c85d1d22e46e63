"""Experiment: two contexts on one GPU, frames alternating (2 frames in flight)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bevy_gaussian_splatting_b200 as B
cloud = B.random_gaussians_3d_seeded(6_000_000, 0)
p = [B.GaussianSplattingPlugin(0) for _ in range(2)]
h = p[0].add_cloud(cloud, f16=True)
s = B.CloudSettings(global_scale=0.02); v = B.headless_view(1920, 1080)
for q in p:
    q.render_view(h, s, v, fmt="rgba8_srgb", to_host=False)
for nctx in (1, 2):
    for _ in range(10):
        for i in range(nctx): p[i].render_view(h, s, v, fmt="rgba8_srgb", to_host=False, asynchronous=True)
    for q in p: q.sync()
    torch.cuda.synchronize()
    K = 200
    t0 = time.perf_counter()
    for i in range(K):
        p[i % nctx].render_view(h, s, v, fmt="rgba8_srgb", to_host=False, asynchronous=True)
    for q in p: q.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print(f"contexts={nctx}: {dt*1e6:.1f} us/frame -> {6e6/dt/1e6:.0f} Msplats/s", flush=True)
