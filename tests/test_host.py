"""Host-side mirror of the reference interface: settings, pass plan, cameras, generator."""
import math

import numpy as np

import bevy_gaussian_splatting_b200 as B
from bevy_gaussian_splatting_b200 import abi


def test_cloud_settings_defaults_match_reference():
    """src/gaussian/settings.rs:110-133."""
    s = B.CloudSettings()
    assert (s.aabb, s.global_opacity, s.global_scale, s.opacity_adaptive_radius) == (False, 1.0, 1.0, True)
    assert s.radix_sort_depth_bits == B.RadixSortDepthBits.Bits32
    assert s.gaussian_mode == B.GaussianMode.Gaussian3d and s.rasterize_mode == B.RasterizeMode.Color
    assert s.draw_mode == B.DrawMode.All and s.color_space == B.GaussianColorSpace.SrgbRec709Display
    a = s.to_abi()
    assert (a.gaussian_mode, a.rasterize_mode, a.aabb, a.opacity_adaptive_radius, a.draw_mode,
            a.radix_sort_depth_bits, a.flags) == (1, 0, 0, 1, 0, 32, 0)
    assert B.CloudSettings(sort_all=True).to_abi().flags == abi.BGS_FLAG_SORT_ALL


def test_headless_camera_matrices():
    """examples/headless.rs:177-184 + glam perspective_infinite_reverse_rh (fov pi/4, near 0.1)."""
    v = B.headless_view(1920, 1080)
    f = 1.0 / math.tan(math.pi / 8)
    assert np.isclose(v.clip_from_view[1, 1], f) and np.isclose(v.clip_from_view[0, 0], f / (1920 / 1080))
    assert v.clip_from_view[3, 2] == -1.0 and np.isclose(v.clip_from_view[2, 3], 0.1)
    assert np.allclose(v.view_from_world[:3, :3], np.eye(3)) and np.allclose(v.view_from_world[:3, 3], [0, -1.5, -5])
    # a point 10 units in front of the camera: ndc.z = near / depth (reverse-Z), inside (0, 1)
    p = np.array([0, 1.5, -5, 1], np.float32)
    c = v.clip_from_world @ p
    assert np.isclose(c[2] / c[3], 0.1 / 10.0) and abs(c[0]) < 1e-6
    a = v.to_abi()
    assert list(a.viewport) == [0.0, 0.0, 1920.0, 1080.0]
    assert np.allclose(np.array(list(a.clip_from_world)).reshape(4, 4).T, v.clip_from_world)   # column-major
    assert np.allclose(B.orbit_view(0, 8).view_from_world, v.view_from_world, atol=1e-6)


def test_random_gaussians_distributions_and_determinism():
    """planar_3d.rs:120-168: ranges per field; same (n, seed) -> same cloud; prefix-stable in n."""
    c = B.random_gaussians_3d_seeded(50_000, 0)
    assert c.rotation.min() >= -1 and c.rotation.max() < 1 and abs(c.rotation.mean()) < 0.02
    assert c.position_visibility[:, :3].min() >= -20 and c.position_visibility[:, :3].max() < 20
    assert np.all(c.position_visibility[:, 3] == 1.0)
    assert c.scale_opacity[:, :3].min() >= 0 and c.scale_opacity[:, :3].max() < 1
    assert c.scale_opacity[:, 3].min() >= 0 and c.scale_opacity[:, 3].max() < 0.8
    assert c.spherical_harmonic.shape == (50_000, 48) and c.spherical_harmonic.min() >= -1
    d = B.random_gaussians_3d_seeded(50_000, 0)
    assert np.array_equal(c.spherical_harmonic, d.spherical_harmonic)
    e = B.random_gaussians_3d_seeded(300_000, 0)
    assert np.array_equal(e.rotation[:50_000], c.rotation)
    assert not np.array_equal(B.random_gaussians_3d_seeded(1000, 1).rotation, c.rotation[:1000])


def test_planar_layout_sizes():
    """a1: 240 B/gaussian f32, 128 B/gaussian f16 (position stays f32)."""
    c = B.random_gaussians_3d_seeded(10, 0)
    f32_bytes = sum(a.nbytes for a in (c.position_visibility, c.spherical_harmonic, c.rotation, c.scale_opacity)) // 10
    shp, rso = c.pack_f16()
    f16_bytes = (c.position_visibility.nbytes + shp.nbytes + rso.nbytes) // 10
    assert (f32_bytes, f16_bytes) == (240, 128)
