"""Ad-hoc GPU probe: parity diagnostics + stage timing at a few sizes (not a test, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bevy_gaussian_splatting_b200 as B
from oracle import oracle as O


def parity(n, w, h, scale, f16=False, bits=32, sort_all=False):
    cloud = B.random_gaussians_3d_seeded(n, 1)
    view = B.headless_view(w, h)
    s = B.CloudSettings(global_scale=scale, radix_sort_depth_bits=B.RadixSortDepthBits(bits), sort_all=sort_all)
    pl = B.GaussianSplattingPlugin(0)
    hnd = pl.add_cloud(cloud, f16=f16)
    img = pl.render_view(hnd, s, view)
    oc = cloud.rounded_to_f16() if f16 else cloud
    u = pl.cloud_uniform(s)
    keys = O.keygen(oc.position_visibility, view.to_abi(), u, bits)
    sk, si = O.radix_sort(keys, bits)
    got = pl.sorted_entries()
    ok_sort = np.array_equal(got[:, 0], sk) and np.array_equal(got[:, 1], si)
    til = O.render_tiles(oc, view.to_abi(), u, s.to_abi())
    rng_ok = np.array_equal(pl.tile_ranges(), til["tile_ranges"])
    ent_ok = np.array_equal(pl.tile_entries(), til["tile_entries"])
    rec, ids = pl.projected()
    orec = O.project(oc, view.to_abi(), u, s.to_abi(), til["rank_to_id"])
    geo = np.stack([orec[k] for k in ("cx", "cy", "ux", "uy", "vx", "vy")], 1)
    drawn = orec["xlo"] <= orec["xhi"]
    geo_ok = np.array_equal(rec[drawn, :6].view(np.uint32), geo[drawn].view(np.uint32))
    col = np.stack([orec[k] for k in ("r", "g", "b", "op")], 1)
    col_err = np.abs(rec[drawn, 8:12] - col[drawn]).max() if drawn.any() else 0
    err = np.abs(img - til["image"]).max()
    fs = pl.frame_stats()
    print(f"parity n={n} {w}x{h} scale={scale} f16={f16} bits={bits} sort_all={sort_all}: sort={ok_sort} ranges={rng_ok} "
          f"entries={ent_ok} ids={np.array_equal(ids, til['rank_to_id'])} geo={geo_ok} col_err={col_err:.2e} Linf={err:.2e} "
          f"n_vis={fs.n_visible} pairs={fs.n_pairs} stage_us={pl.stage_times_us().round(1).tolist()}", flush=True)


def timing(n, scale, f16, frames=20, sort_all=False):
    t0 = time.time()
    cloud = B.random_gaussians_3d_seeded(n, 0)
    view = B.headless_view(1920, 1080)
    s = B.CloudSettings(global_scale=scale, sort_all=sort_all)
    pl = B.GaussianSplattingPlugin(0)
    hnd = pl.add_cloud(cloud, f16=f16)
    t1 = time.time()
    rows = []
    for i in range(frames):
        pl.render_view(hnd, s, view, fmt="rgba8_srgb", to_host=False)
        rows.append(pl.stage_times_us())
    rows = np.array(rows[3:])
    fs = pl.frame_stats()
    med = np.median(rows, 0)
    print(f"timing n={n} scale={scale} f16={f16} sort_all={sort_all}: gen+upload {t1-t0:.1f}s n_vis={fs.n_visible} pairs={fs.n_pairs} "
          f"median stage_us [keygen, sort, project, bin, raster, frame] = {med.round(1).tolist()} -> {n/med[5]:.1f} Msplats/s", flush=True)


if __name__ == "__main__":
    import __graft_entry__ as g
    g.smoke()
    parity(100_000, 640, 360, 0.1)
    parity(100_000, 640, 360, 0.1, f16=True)
    parity(50_000, 320, 200, 1.0)
    parity(200_000, 640, 360, 0.05, bits=24)
    parity(200_000, 640, 360, 0.05, bits=16, sort_all=True)
    timing(1_000_000, 0.02, False)
    timing(6_000_000, 0.02, True)
    timing(6_000_000, 0.02, True, sort_all=True)
    timing(1_000_000, 1.0, False, frames=8)
