"""The oracle against an INDEPENDENT second restatement (tests/wgsl_emu.py: WGSL value semantics + the reference
functions transliterated line by line, float64).  Off-axis camera, non-identity (rotated, non-uniformly scaled,
translated) model matrix, so a misread matrix convention, eigenvector sign, SH direction or homography shows up.

CPU only: this pins oracle/bgs_oracle.cpp; the CUDA path is pinned to the oracle by tests/test_gpu_parity.py."""
import math

import numpy as np
import pytest

import bevy_gaussian_splatting_b200 as B
from bevy_gaussian_splatting_b200.plugin import CloudTransform, GaussianSplattingPlugin

from wgsl_emu import CloudU, Shader, Vec, ViewU, ndc_to_pixel, render_reference_semantics

W, H = 640, 360


def _model():
    a, b = 0.7, -0.4
    ry = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = (ry @ rx @ np.diag([1.3, 0.8, 1.1])).astype(np.float32)
    m[:3, 3] = [0.5, -0.25, 1.0]
    return m


def _view(w=W, h=H):
    # off-axis: the eye is off every axis, the target is not the origin, the up vector is tilted
    return B.perspective_view((3.0, 2.5, 6.0), (-1.0, 0.5, -2.0), w, h, up=(0.15, 1.0, -0.05))


def _cloud(n, seed, spread=6.0):
    c = B.random_gaussians_3d_seeded(n, seed)
    c.position_visibility[:, :3] *= np.float32(spread / 20.0)
    c.scale_opacity[:, 3] = np.maximum(c.scale_opacity[:, 3], 0.02)
    return c


def _shader(view, u, s):
    mode = {B.RasterizeMode.Color: "color", B.RasterizeMode.Depth: "depth", B.RasterizeMode.Normal: "normal",
            B.RasterizeMode.Position: "position"}[s.rasterize_mode]
    return Shader(ViewU(view.to_abi()), CloudU(u), use_obb=not s.aabb, gaussian_2d=s.gaussian_mode == B.GaussianMode.Gaussian2d,
                  adaptive=s.opacity_adaptive_radius, rasterize=mode)


def _vs(sh, cloud, i, **kw):
    pv = cloud.position_visibility[i]
    return sh.vs_points(Vec(pv[0], pv[1], pv[2]), [float(x) for x in cloud.spherical_harmonic[i]],
                        Vec(*[float(x) for x in cloud.rotation[i]]), [float(x) for x in cloud.scale_opacity[i]],
                        visibility=float(pv[3]), **kw)


def _visible_ids(oracle, cloud, view, u):
    keys = oracle.keygen(cloud.position_visibility, view.to_abi(), u, 32)
    return np.nonzero(keys != 0xFFFFFFFF)[0].astype(np.uint32), keys


SETTINGS = [
    dict(),                                                         # 3DGS, USE_OBB, adaptive cutoff, sRGB decode
    dict(opacity_adaptive_radius=False, color_space=B.GaussianColorSpace.LinRec709Display, global_opacity=1.7),
    dict(aabb=True),                                                # 3DGS USE_AABB conic
    dict(gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True),       # 2DGS ray-splat
    dict(gaussian_mode=B.GaussianMode.Gaussian2d),                  # 2DGS quad-uv
    dict(rasterize_mode=B.RasterizeMode.Normal),
    dict(rasterize_mode=B.RasterizeMode.Position),
]


@pytest.mark.parametrize("kw", SETTINGS, ids=[",".join(f"{k}={getattr(v, 'name', v)}" for k, v in kw.items()) or "default" for kw in SETTINGS])
def test_oracle_projection_matches_wgsl_literal(oracle, kw):
    cloud = _cloud(260, 17)
    view = _view()
    s = B.CloudSettings(global_scale=0.35, **kw)
    tr = CloudTransform(_model())
    u = GaussianSplattingPlugin.cloud_uniform(s, tr, cloud.compute_aabb())
    ids, keys = _visible_ids(oracle, cloud, view, u)
    assert 60 <= len(ids) <= 250, len(ids)
    orec = oracle.project(cloud, view.to_abi(), u, s.to_abi(), ids)
    sh = _shader(view, u, s)
    checked = obb_checked = 0
    for k, gi in enumerate(ids):
        vs = _vs(sh, cloud, int(gi))
        ndc = sh.world_to_clip((sh.gu.transform * Vec(*cloud.position_visibility[gi, :3].tolist(), 1.0)).xyz)
        margin = min(abs(abs(ndc.x) - 1.1), abs(abs(ndc.y) - 1.1), abs(ndc.z), abs(ndc.z - 1.0))
        if margin < 1e-4:
            continue                      # f32-vs-f64 rounding may decide such a point either way
        assert vs is not None, f"gaussian {gi}: the oracle draws it, the WGSL restatement culls it"
        o = orec[k]
        # key: 0xFFFFFFFF - bits(|p_w - cam|^2)   (radix.wgsl:86-101)
        d = vs["transformed_position"] - sh.view.world_position
        d2 = np.float32(d.x * d.x + d.y * d.y + d.z * d.z)
        assert abs(int(0xFFFFFFFF - int(keys[gi])) - int(d2.view(np.uint32))) <= 64      # a few ulp of f32 d^2
        # centre in pixels
        c = ndc_to_pixel(vs["projected_position"].xy.v, W, H)
        assert abs(c[0] - o["cx"]) <= 2e-3 and abs(c[1] - o["cy"]) <= 2e-3, (gi, c, o["cx"], o["cy"])
        # colour + opacity
        col = vs["color"]
        assert np.allclose([o["r"], o["g"], o["b"]], col.v[:3], rtol=2e-5, atol=2e-5), (gi, col.v, o["r"], o["g"], o["b"])
        assert abs(o["op"] - col[3]) <= 1e-6 * max(1.0, abs(col[3]))
        drawn = o["xlo"] <= o["xhi"]
        if s.gaussian_mode == B.GaussianMode.Gaussian3d and not s.aabb:
            # USE_OBB: the oracle's pixel-offset -> uv map must send each emitted quad corner to its own uv (+-1, +-1).
            # Near-isotropic footprints have an ill-conditioned eigenvector; skip those (the map, not the maths, is unstable).
            cv = vs["cov2d"]
            mid = 0.5 * (cv.x + cv.z)
            lam1 = mid + math.sqrt(max(0.0, mid * mid - (cv.x * cv.z - cv.y * cv.y)))
            if math.hypot(-cv.y, lam1 - cv.x) > 2e-2 * mid and np.isfinite([o["ux"], o["uy"], o["vx"], o["vy"]]).all():
                for v in vs["vertices"]:
                    p = ndc_to_pixel(v["position"].xy.v, W, H)
                    du, dv = p[0] - c[0], p[1] - c[1]
                    uu = o["ux"] * du + o["uy"] * dv
                    vv = o["vx"] * du + o["vy"] * dv
                    assert abs(uu - v["uv"].x) <= 2e-3 and abs(vv - v["uv"].y) <= 2e-3, (gi, uu, vv, v["uv"].v)
                obb_checked += 1
                # and the conservative pixel bbox must contain every emitted corner that lies on screen
                if drawn:
                    P = np.array([ndc_to_pixel(v["position"].xy.v, W, H) for v in vs["vertices"]])
                    assert o["xlo"] <= max(0, math.ceil(P[:, 0].min() - 0.5)) and o["xhi"] >= min(W - 1, math.floor(P[:, 0].max() - 0.5))
                    assert o["ylo"] <= max(0, math.ceil(P[:, 1].min() - 0.5)) and o["yhi"] >= min(H - 1, math.floor(P[:, 1].max() - 0.5))
        elif s.gaussian_mode == B.GaussianMode.Gaussian3d:
            conic, rq = vs["conic"], abs(vs["vertices"][3]["bb"].z)
            assert np.allclose(o["extra"][:3], conic.v, rtol=3e-4, atol=1e-7), (gi, o["extra"][:4], conic.v)
            assert abs(o["extra"][3] - rq) <= 3e-4 * rq
        else:
            sf = vs["surfel"]
            if sf["local_to_pixel"] is not None and sf["extent"].x >= 1e-3 and sf["extent"].y >= 1e-3:
                T = sf["local_to_pixel"]
                rq = max(max(math.sqrt(sf["extent"].x), math.sqrt(sf["extent"].y)), vs["cutoff"] * 0.707106)
                e = o["extra"]
                scale_t = max(np.abs(T[0].v).max(), np.abs(T[1].v).max(), np.abs(T[2].v).max())
                # the homography is ill-conditioned when d = test . T2^2 is small against its terms; scale the tolerance
                t2 = T[2].v
                cond = (vs["cutoff"] ** 2 * (t2[0] ** 2 + t2[1] ** 2) + t2[2] ** 2) / abs(vs["cutoff"] ** 2 * (t2[0] ** 2 + t2[1] ** 2) - t2[2] ** 2)
                tol = 3e-5 * cond
                # oracle layout: extra[3] = quad half-side, [4:6] = mean_2d, [6] = W / H, [7:10] / [10:13] / [13:16] = T[0] / T[1] / T[2]
                # extent = mean^2 - t cancels (mean ~ hundreds of pixels, extent ~ tens): the f32 oracle carries that
                # relative error into the quad half-side sqrt(extent)
                m2 = sf["mean_2d"]
                canc = max((m2.x ** 2 + abs(m2.x ** 2 - sf["extent"].x)) / sf["extent"].x, (m2.y ** 2 + abs(m2.y ** 2 - sf["extent"].y)) / sf["extent"].y)
                if tol < 5e-2:
                    assert abs(e[3] - rq) <= (tol + 1e-6 * cond * canc) * rq + 1e-4, (gi, e[3], rq)
                    assert abs(e[4] - m2.x) <= tol * (abs(m2.x) + rq) and abs(e[5] - m2.y) <= tol * (abs(m2.y) + rq)
                assert abs(e[6] - W / H) < 1e-6
                assert np.allclose(e[7:10], T[0].v, rtol=0, atol=1e-4 * scale_t) and np.allclose(e[10:13], T[1].v, rtol=0, atol=1e-4 * scale_t)
                assert np.allclose(e[13:16], T[2].v, rtol=0, atol=1e-4 * scale_t)
        checked += 1
    assert checked >= 50
    if s.gaussian_mode == B.GaussianMode.Gaussian3d and not s.aabb:
        assert obb_checked >= 30


@pytest.mark.parametrize("kw,n,scale", [(dict(), 48, 0.45), (dict(aabb=True), 40, 0.4),
                                        (dict(gaussian_mode=B.GaussianMode.Gaussian2d, aabb=True), 40, 0.5),
                                        (dict(gaussian_mode=B.GaussianMode.Gaussian2d, global_opacity=1.5), 36, 0.5),
                                        (dict(draw_mode=B.DrawMode.HighlightSelected), 30, 0.4),
                                        (dict(draw_mode=B.DrawMode.Selected), 45, 0.45),
                                        (dict(rasterize_mode=B.RasterizeMode.Depth), 40, 0.45),
                                        (dict(rasterize_mode=B.RasterizeMode.Normal), 36, 0.45),
                                        (dict(rasterize_mode=B.RasterizeMode.Position, gaussian_mode=B.GaussianMode.Gaussian2d), 36, 0.5)],
                         ids=["3dgs-obb", "3dgs-aabb", "2dgs-aabb", "2dgs-obb", "highlight", "selected", "depth", "normal", "2dgs-position"])
def test_oracle_ref_mode_matches_wgsl_literal_frame(oracle, kw, n, scale):
    """Whole frames: the emulator rasterises the emitted quads (vs_points -> affine patch -> fs_main -> premultiplied
    "over", far -> near) and must reproduce the oracle's ref_mode image."""
    w, h = 112, 72
    cloud = _cloud(n, 23, spread=3.0)
    cloud.position_visibility[::3, 3] = 0.0
    view = _view(w, h)
    s = B.CloudSettings(global_scale=scale, **kw)
    tr = CloudTransform(_model())
    u = GaussianSplattingPlugin.cloud_uniform(s, tr, cloud.compute_aabb())
    keys = oracle.keygen(cloud.position_visibility, view.to_abi(), u, 32)
    _, full_order = oracle.radix_sort(keys, 32)         # ascending key = far -> near, culled (all-ones) last
    order = [int(i) for i in full_order if keys[i] != 0xFFFFFFFF]
    assert len(order) >= 12
    got = oracle.render_ref(cloud, view.to_abi(), u, s.to_abi())
    shader = _shader(view, u, s)
    # RasterizeMode::Depth reads its range from entries 1 and count - 1 of the sorted buffer (gaussian.wgsl:331-332):
    # with culled gaussians in the cloud the last entry IS a culled one -- the literal behaviour, restated as such
    shader.depth_entries = (Vec(*cloud.position_visibility[int(full_order[1]), :3].tolist()),
                            Vec(*cloud.position_visibility[int(full_order[len(cloud) - 1]), :3].tolist()))
    want = render_reference_semantics(shader, cloud, order, w, h,
                                      draw_selected=s.draw_mode == B.DrawMode.Selected,
                                      highlight_selected=s.draw_mode == B.DrawMode.HighlightSelected)
    assert (want[..., :3].max(axis=2) > 0.02).sum() >= 200, "the scene must actually cover pixels"
    diff = np.abs(got.astype(np.float64) - want)
    # a pixel centre within rounding of a quad edge may be covered on one side only: allow a handful of such pixels
    bad = diff.max(axis=2) > 2e-4
    assert bad.sum() <= 3, (int(bad.sum()), float(diff.max()))
