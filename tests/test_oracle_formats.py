"""Oracle: f16 planar layout (a1/A.9) and the fixed-series ln()."""
import numpy as np

import bevy_gaussian_splatting_b200 as B


def test_f16_pack_matches_ieee_rne(oracle):
    """f16.rs:244-252 pack(upper, lower) = f16(upper) << 16 | f16(lower), half::from_f32 = IEEE RNE
    (checked against numpy's float16 conversion, incl. subnormals / overflow / ties)."""
    rng = np.random.default_rng(0)
    n = 4096
    vals = np.concatenate([
        rng.uniform(-1, 1, n * 56 - 64).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, 6.103e-5,
                  1.0009766, 1.0004883, 1.0014648, np.inf, -np.inf, 1e10, -1e10, 3.0e-8, 8.9e-8] + [0.33325195] * 44,
                 np.float32)])
    sh = vals[: n * 48].reshape(n, 48)
    rot = vals[n * 48: n * 52].reshape(n, 4)
    so = vals[n * 52: n * 56].reshape(n, 4)
    shp, rso = oracle.pack_f16(sh, rot, so)
    cloud = B.PlanarGaussian3d(np.zeros((n, 4), np.float32), sh, rot, so)
    shp2, rso2 = cloud.pack_f16()
    assert np.array_equal(shp, shp2) and np.array_equal(rso, rso2)


def test_f16_decode_roundtrip_and_field_order(oracle):
    """planar.wgsl:117-176: sh[2k] = low half, sh[2k+1] = high half; rotation = (hi w0, lo w0, hi w1, lo w1);
    scale = (hi w2, lo w2, hi w3); opacity = lo w3."""
    rng = np.random.default_rng(1)
    n = 257
    c = B.PlanarGaussian3d(np.zeros((n, 4), np.float32), rng.uniform(-1, 1, (n, 48)), rng.uniform(-1, 1, (n, 4)),
                           rng.uniform(0, 1, (n, 4)))
    shp, rso = oracle.pack_f16(c.spherical_harmonic, c.rotation, c.scale_opacity)
    sh, rot, so = oracle.decode_f16(shp, rso)
    r = c.rounded_to_f16()
    assert np.array_equal(sh, r.spherical_harmonic) and np.array_equal(rot, r.rotation) and np.array_equal(so, r.scale_opacity)
    h = lambda x: int(np.float16(x).view(np.uint16))
    assert int(rso[0, 0]) == (h(c.rotation[0, 0]) << 16) | h(c.rotation[0, 1])
    assert int(rso[0, 3]) == (h(c.scale_opacity[0, 2]) << 16) | h(c.scale_opacity[0, 3])
    assert int(shp[0, 5]) == (h(c.spherical_harmonic[0, 11]) << 16) | h(c.spherical_harmonic[0, 10])


def test_fixed_series_ln(oracle):
    """The policy ln(): within 1 ulp of the exact value over the opacity range, exact specials."""
    lib = oracle.load()
    xs = np.concatenate([np.geomspace(1e-38, 1e38, 2000), np.linspace(1e-6, 1.0, 2000), [0.8, 0.5, 1.0, 2.0, np.e]]).astype(np.float32)
    got = np.array([lib.orc_ln(float(x)) for x in xs], np.float32)
    want = np.log(xs.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - want) <= 0.5000001 * ulp + 1e-45)
    assert lib.orc_ln(1.0) == 0.0
    assert lib.orc_ln(0.0) == -np.inf
    assert np.isnan(lib.orc_ln(-1.0)) and np.isnan(lib.orc_ln(float("nan")))
    assert lib.orc_ln(float("inf")) == np.inf
    sub = np.float32(1e-42)
    assert abs(lib.orc_ln(float(sub)) - np.log(np.float64(sub))) < 1e-4


def test_precomputed_covariance_form_matches_covariance_rs(oracle):
    """Row f3: the host's `Covariance3dOpacity` (PlanarGaussian3d.precomputed_covariance, numpy f32) against the oracle's
    restatement of src/gaussian/covariance.rs:4-41, bit for bit, and its f16 packing (Covariance3dOpacityPacked128,
    f16.rs:131-170: pack(upper, lower) words, opacity in both halves)."""
    import numpy as np
    import bevy_gaussian_splatting_b200 as B

    cloud = B.random_gaussians_3d_seeded(5000, 3)
    cov = cloud.precomputed_covariance()
    want = oracle.covariance_3d(cloud.rotation, cloud.scale_opacity)
    got = np.concatenate([cov.rotation, cov.scale_opacity[:, :2]], axis=1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(cov.scale_opacity[:, 2], cloud.scale_opacity[:, 3]) and np.array_equal(cov.scale_opacity[:, 3], cloud.scale_opacity[:, 3])
    # symmetric positive semi-definite, and equal to an independent float64 evaluation
    q, s = cloud.rotation.astype(np.float64), cloud.scale_opacity[:, :3].astype(np.float64)
    r, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)], 1),
                  np.stack([2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)], 1),
                  np.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], 1)], 1)   # R[:, row, col]
    M = s[:, :, None] * R
    S64 = np.einsum("nki,nkj->nij", M, M)
    ref = np.stack([S64[:, 0, 0], S64[:, 0, 1], S64[:, 0, 2], S64[:, 1, 1], S64[:, 1, 2], S64[:, 2, 2]], 1)
    assert np.abs(want - ref).max() <= 1e-5 * np.abs(ref).max()
    sh_p, words = cov.pack_f16()
    h = cov.rounded_to_f16()
    dec = oracle.decode_f16(sh_p, words)
    assert np.array_equal(dec[1], h.rotation) and np.array_equal(dec[2], h.scale_opacity)
    assert np.array_equal(words[:, 3] >> 16, words[:, 3] & 0xFFFF)        # opacity packed in both halves
