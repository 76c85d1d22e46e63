import os, sys, ctypes as C
os.environ["BGS_TIMELINE"] = "1"; os.environ["BGS_TIMELINE_KEYGEN"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
cloud = B.random_gaussians_3d_seeded(6_000_000, 0)
pl = B.GaussianSplattingPlugin(0); h = pl.add_cloud(cloud, f16=True)
s = B.CloudSettings(global_scale=0.02); v = B.headless_view(1920, 1080)
for _ in range(5): pl.render_view(h, s, v, fmt="rgba8_srgb", to_host=False)
buf = np.zeros((4096, 8), np.uint64); g = C.c_uint32()
lib = pl._lib; lib.bgs_debug_timeline_.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
lib.bgs_debug_timeline_(pl._ctx, buf.ctypes.data_as(C.c_void_p), C.byref(g))
nb = int((buf[:, 0] != 0).sum())
t = buf[:nb, :6].astype(np.int64); t0 = t[:, 0].min(); t = (t - t0) / 1000.0
for i, nm in enumerate(["start", "phase1 done", "barrier passed", "prefix done", "phase2 done", "end"]):
    col = t[:, i]; print(f"{nm:16s} min {col.min():7.1f} median {np.median(col):7.1f} max {col.max():7.1f} us")
print("blocks", nb, pl.stage_times_us())
