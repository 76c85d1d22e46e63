"""Host-side cloud container and synthetic generators.

`PlanarGaussian3d` mirrors the reference's struct-of-Vecs (src/gaussian/formats/planar_3d.rs:45-54;
planes in binding order position_visibility, spherical_harmonic, rotation, scale_opacity) as four
C-contiguous float32 numpy arrays -- exactly the four host pointers `bgs_cloud_upload_f32` borrows.

`random_gaussians_3d_seeded` follows the reference generator's distributions and field order
(planar_3d.rs:120-168: rotation 4xU(-1,1) unnormalised; position 3xU(-20,20), visibility 1;
scale 3xU(0,1); opacity U(0,0.8); SH 48xU(-1,1)) with this repo's own counter-based PRNG (numpy
Philox): the reference's ChaCha12 stream is not reproduced bit-for-bit.
"""
from __future__ import annotations

import dataclasses

import numpy as np

SH_COEFF_COUNT = 48  # src/material/spherical_harmonics.rs:46-47 (sh3 default)
HALF_SH_COEFF_COUNT = 24


@dataclasses.dataclass
class PlanarGaussian3d:
    position_visibility: np.ndarray  # (n, 4) f32: x, y, z, visibility      f32.rs:53-56
    spherical_harmonic: np.ndarray   # (n, 48) f32: sh[3k + c]              spherical_harmonics.rs:114-120
    rotation: np.ndarray             # (n, 4) f32: w, x, y, z               f32.rs:95-97
    scale_opacity: np.ndarray        # (n, 4) f32: sx, sy, sz, opacity      f32.rs:172-175

    def __post_init__(self):
        for name, width in (("position_visibility", 4), ("spherical_harmonic", SH_COEFF_COUNT), ("rotation", 4),
                            ("scale_opacity", 4)):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float32)
            if a.ndim != 2 or a.shape[1] != width:
                raise ValueError(f"{name} must have shape (n, {width})")
            setattr(self, name, a)
        n = len(self.position_visibility)
        if not (len(self.spherical_harmonic) == len(self.rotation) == len(self.scale_opacity) == n):
            raise ValueError("planes disagree on n")

    def __len__(self) -> int:
        return len(self.position_visibility)

    def subset(self, n: int) -> "PlanarGaussian3d":
        return PlanarGaussian3d(self.position_visibility[:n], self.spherical_harmonic[:n], self.rotation[:n],
                                self.scale_opacity[:n])

    # ---- f16 planar layout (src/gaussian/f16.rs:30-56,244-263; planar.wgsl:117-176) -----------------
    def compute_aabb(self) -> tuple[np.ndarray, np.ndarray]:
        """The entity `Aabb` as the render world sees it: `compute_aabb` (src/gaussian/interface.rs:22-66: positions
        +- 0.1) -> `Aabb {center, half_extents}` (src/gaussian/cloud.rs:45-62) -> `aabb.min()` / `aabb.max()`
        (center -+ half_extents, src/render/mod.rs:1070-1071), every step in f32 like glam."""
        p = self.position_visibility[:, :3]
        off = np.float32(0.1)
        # (min over (p - 0.1) == min(p) - 0.1: f32 subtraction of a constant is monotone)
        lo = (p.min(0) - off).astype(np.float32) if len(p) else np.full(3, np.inf, np.float32)
        hi = (p.max(0) + off).astype(np.float32) if len(p) else np.full(3, -np.inf, np.float32)
        two = np.float32(2.0)
        center = ((lo + hi).astype(np.float32) / two).astype(np.float32)
        half = ((hi - lo).astype(np.float32) / two).astype(np.float32)
        return (center - half).astype(np.float32), (center + half).astype(np.float32)

    def pack_f16(self) -> tuple[np.ndarray, np.ndarray]:
        """-> (sh_packed (n,24) u32, rot_scale_opacity (n,4) u32); pack(upper, lower) = upper<<16 | lower."""
        def bits(a):
            return np.ascontiguousarray(a, dtype=np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)

        sh = bits(self.spherical_harmonic)
        sh_packed = (sh[:, 1::2] << 16) | sh[:, 0::2]  # even coefficient in the low half
        r, s = bits(self.rotation), bits(self.scale_opacity)
        rso = np.stack([(r[:, 0] << 16) | r[:, 1], (r[:, 2] << 16) | r[:, 3], (s[:, 0] << 16) | s[:, 1],
                        (s[:, 2] << 16) | s[:, 3]], axis=1)
        return np.ascontiguousarray(sh_packed, dtype=np.uint32), np.ascontiguousarray(rso, dtype=np.uint32)

    def precomputed_covariance(self) -> "PlanarGaussian3d":
        """`Covariance3dOpacity` per gaussian (src/gaussian/f32.rs:238-251 <- covariance.rs:4-41, every product in f32, glam's
        accumulation order), laid out in the plane slots `Covariance3dOpacityPacked128` occupies (f16.rs:131-170): rotation
        = (c0, c1, c2, c3), scale_opacity = (c4, c5, opacity, opacity).  `pack_f16()` of the result is the record the
        reference's PRECOMPUTE_COVARIANCE_3D layout uploads; `rounded_to_f16()` is what its shader decodes."""
        f = np.float32
        q, so = self.rotation.astype(f), self.scale_opacity.astype(f)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        two, one = f(2.0), f(1.0)
        R = np.empty((len(q), 3, 3), f)     # R[:, i, j]: row i, column j (columns = the triplets covariance.rs writes)
        R[:, 0, 0] = one - two * (y * y + z * z); R[:, 1, 0] = two * (x * y - r * z); R[:, 2, 0] = two * (x * z + r * y)
        R[:, 0, 1] = two * (x * y + r * z); R[:, 1, 1] = one - two * (x * x + z * z); R[:, 2, 1] = two * (y * z - r * x)
        R[:, 0, 2] = two * (x * z - r * y); R[:, 1, 2] = two * (y * z + r * x); R[:, 2, 2] = one - two * (x * x + y * y)
        M = (so[:, :3, None] * R).astype(f)                      # M = S R: row i scaled by s_i
        def sg(i, j):
            return ((M[:, 0, i] * M[:, 0, j] + M[:, 1, i] * M[:, 1, j]).astype(f) + M[:, 2, i] * M[:, 2, j]).astype(f)
        cov_rot = np.stack([sg(0, 0), sg(0, 1), sg(0, 2), sg(1, 1)], axis=1)
        cov_so = np.stack([sg(1, 2), sg(2, 2), so[:, 3], so[:, 3]], axis=1)
        return PlanarGaussian3d(self.position_visibility, self.spherical_harmonic, cov_rot, cov_so)

    def rounded_to_f16(self) -> "PlanarGaussian3d":
        """The f32 cloud the f16 layout decodes to (position stays f32: bindings.wgsl:104-106)."""
        def rt(a):
            return a.astype(np.float16).astype(np.float32)

        return PlanarGaussian3d(self.position_visibility, rt(self.spherical_harmonic), rt(self.rotation),
                                rt(self.scale_opacity))


def random_gaussians_3d_seeded(n: int, seed: int = 0, chunk: int = 1 << 18) -> PlanarGaussian3d:
    """planar_3d.rs:182-191 (distributions + field order), Philox stream, deterministic in (n, seed)."""
    rng = np.random.Generator(np.random.Philox(seed))
    pos = np.empty((n, 4), np.float32)
    sh = np.empty((n, SH_COEFF_COUNT), np.float32)
    rot = np.empty((n, 4), np.float32)
    so = np.empty((n, 4), np.float32)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        u = rng.random((chunk, 59), dtype=np.float32)[: hi - lo]  # always draw whole chunks: stream is n-independent
        rot[lo:hi] = u[:, 0:4] * 2.0 - 1.0
        pos[lo:hi, 0:3] = u[:, 4:7] * 40.0 - 20.0
        pos[lo:hi, 3] = 1.0
        so[lo:hi, 0:3] = u[:, 7:10]
        so[lo:hi, 3] = u[:, 10] * 0.8
        sh[lo:hi] = u[:, 11:59] * 2.0 - 1.0
    return PlanarGaussian3d(pos, sh, rot, so)


def random_gaussians_3d(n: int) -> PlanarGaussian3d:
    """planar_3d.rs:170-180 (unseeded in the reference; seed 0 here so runs are reproducible)."""
    return random_gaussians_3d_seeded(n, 0)


def test_model(seed: int = 0) -> PlanarGaussian3d:
    """planar_3d.rs:193-251: 8 corner gaussians at (+-0.5)^3 + a repeat of the first, shuffled SH."""
    rng = np.random.Generator(np.random.Philox(seed))
    base = rng.random(SH_COEFF_COUNT, dtype=np.float32) * 2.0 - 1.0
    pos, sh = [], []
    for x in (-0.5, 0.5):
        for y in (-0.5, 0.5):
            for z in (-0.5, 0.5):
                pos.append([x, y, z, 1.0])
                sh.append(rng.permutation(base))
    pos.append(pos[0]); sh.append(sh[0])
    n = len(pos)
    return PlanarGaussian3d(np.array(pos, np.float32), np.array(sh, np.float32),
                            np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
                            np.tile(np.array([0.125] * 4, np.float32), (n, 1)))
