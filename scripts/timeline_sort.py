import os, sys, ctypes as C
os.environ["BGS_TIMELINE"] = "1"; os.environ["BGS_TIMELINE_SORT"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
cloud = B.random_gaussians_3d_seeded(6_000_000, 0)
pl = B.GaussianSplattingPlugin(0); h = pl.add_cloud(cloud, f16=True)
s = B.CloudSettings(global_scale=0.02, rasterize_mode=B.RasterizeMode.Depth)   # Depth: projection waits, the sort runs alone
v = B.headless_view(1920, 1080)
for _ in range(5): pl.render_view(h, s, v, fmt="rgba8_srgb", to_host=False)
buf = np.zeros((4096, 8), np.uint64); g = C.c_uint32()
lib = pl._lib; lib.bgs_debug_timeline_.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
lib.bgs_debug_timeline_(pl._ctx, buf.ctypes.data_as(C.c_void_p), C.byref(g))
nt = (721340 + 2047) // 2048
t = buf[:nt, :5].astype(np.int64); t = (t - t[:, :1]) / 1965.0   # clock64 cycles -> us at 1.965 GHz, per-tile origin
for i, nm in enumerate(["tile start", "ranked", "pre-lookback", "post-lookback", "written"]):
    col = t[:, i]; print(f"{nm:14s} min {col.min():7.1f} median {np.median(col):7.1f} max {col.max():7.1f} us")
print("per-phase medians:", np.median(np.diff(t, axis=1), axis=0).round(2))
print("look-back duration by tile idx (every 32):", (t[::32, 3] - t[::32, 2]).round(1))
print(pl.stage_times_us())
