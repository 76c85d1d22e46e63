"""Row f1: the INRIA `.ply` input format, restating src/io/ply.rs:23-132 incl. its quirks."""
import os

import numpy as np

import bevy_gaussian_splatting_b200 as B
from bevy_gaussian_splatting_b200 import io as bio


def test_ply_transformations_and_quirks(tmp_path):
    rng = np.random.default_rng(0)
    n = 100
    c = B.random_gaussians_3d_seeded(n, 4)
    c.scale_opacity[:, :3] = rng.uniform(0.01, 0.5, (n, 3))
    c.scale_opacity[0, :3] = [1e-6, 0.5, 0.5]          # log-scale spread > 4 around the mean -> clamped
    c.scale_opacity[:, 3] = rng.uniform(0.05, 0.95, n)
    p = tmp_path / "cloud.ply"
    bio.write_ply_3d(p, c)
    got = bio.parse_ply_3d(str(p))
    assert len(got) == n + (32 - n % 32) == 128                       # padded to a multiple of 32 (ply.rs:127-129)
    assert np.allclose(got.position_visibility[:n, :3], c.position_visibility[:, :3])
    assert np.all(got.position_visibility[:, 3] == 1.0)
    assert np.allclose(got.scale_opacity[:n, 3], c.scale_opacity[:, 3], atol=2e-6)      # sigmoid(logit(o))
    assert np.allclose(got.scale_opacity[1:n, :3], c.scale_opacity[1:, :3], rtol=2e-5)   # exp(log(s))
    raw = np.log(c.scale_opacity[0, :3].astype(np.float32)); mean = raw.sum() / 3
    assert np.allclose(got.scale_opacity[0, :3], np.exp(np.clip(raw, mean - 4, mean + 4)), rtol=1e-5)
    assert got.scale_opacity[0, 0] > c.scale_opacity[0, 0] * 10           # the tiny axis was clamped up
    assert np.allclose(np.linalg.norm(got.rotation[:n], axis=1), 1.0, atol=1e-5)          # normalised
    assert np.allclose(got.rotation[:n], c.rotation / np.linalg.norm(c.rotation, axis=1, keepdims=True), atol=1e-5)
    # dc band
    assert np.allclose(got.spherical_harmonic[:n, :3], c.spherical_harmonic[:, :3])
    # f_rest quirk (ply.rs:49-69): channel = i // 16 (not // 15), coefficient = (i % 15) + 1, later properties
    # overwrite earlier ones, indices >= 48 are dropped.  Expected planes from a literal replay:
    f_rest = lambda i: c.spherical_harmonic[:, ((i % 15) + 1) * 3 + i // 15]      # what write_ply_3d stored
    want = np.zeros((n, 48), np.float32)
    want[:, :3] = c.spherical_harmonic[:, :3]
    for i in range(45):
        idx = ((i % 15) + 1) * 3 + i // 16
        if idx < 48:
            want[:, idx] = f_rest(i)
    assert np.allclose(got.spherical_harmonic[:n], want)
    assert np.allclose(got.spherical_harmonic[:n, 1 * 3 + 0], f_rest(15))   # f_rest_15 overwrote f_rest_0 on channel 0
    assert not np.allclose(want, c.spherical_harmonic)                      # i.e. NOT the INRIA mapping
    # padding = Gaussian3d::default(): zeros, visibility 1
    assert np.all(got.scale_opacity[n:] == 0) and np.all(got.rotation[n:] == 0) and np.all(got.spherical_harmonic[n:] == 0)
    # a multiple of 32 still gets a full block of padding
    bio.write_ply_3d(p, c, 64)
    assert len(bio.parse_ply_3d(p.read_bytes())) == 96


def test_ply_missing_required_property_and_ascii(tmp_path):
    import pytest

    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nend_header\n0 0\n")
    with pytest.raises(ValueError, match="missing required"):
        bio.parse_ply_3d(str(p))
    props = bio.REQUIRED + ["scale_2"]
    body = " ".join(["0.5"] * len(props))
    q = tmp_path / "ok.ply"
    q.write_text("ply\nformat ascii 1.0\nelement vertex 2\n" + "".join(f"property float {k}\n" for k in props) + "end_header\n" + body + "\n" + body + "\n")
    c = bio.parse_ply_3d(str(q))
    assert len(c) == 32 and np.allclose(c.scale_opacity[:2, 3], 1 / (1 + np.exp(-0.5)))


def test_ply_cloud_renders_through_the_oracle(oracle, tmp_path):
    c = B.random_gaussians_3d_seeded(500, 2)
    c.scale_opacity[:, :3] *= 0.2
    c.scale_opacity[:, 3] = np.clip(c.scale_opacity[:, 3], 0.05, 0.75)
    p = tmp_path / "scene.ply"
    bio.write_ply_3d(p, c)
    cloud = bio.parse_ply_3d(str(p))
    s = B.CloudSettings()
    view = B.headless_view(96, 64)
    u = B.GaussianSplattingPlugin.cloud_uniform(s)
    ref = oracle.render_ref(cloud, view.to_abi(), u, s.to_abi())
    til = oracle.render_tiles(cloud, view.to_abi(), u, s.to_abi())["image"]
    assert np.isfinite(ref).all() and ref[..., :3].max() > 0.01 and np.abs(ref - til).max() <= 1e-3


def test_cpp_host_mirror_reads_the_same_planes(tmp_path):
    """include/bgs_io.hpp (`bgs::io::parse_ply_3d`, the compiled-language host mirror of src/io/ply.rs) against the
    Python mirror on the same file: identical planes (exp / sigmoid within 2 ulp of each other's libm), same padding,
    same error on a missing required property.  CPU only (examples/cloud_tool links nothing from libbgs.so)."""
    import subprocess

    root = os.path.join(os.path.dirname(__file__), "..")
    subprocess.run(["make", "-C", os.path.join(root, "examples"), "-s", "cloud_tool"], check=True)
    tool = os.path.join(root, "examples", "cloud_tool")
    cloud = B.random_gaussians_3d_seeded(333, 6)
    bio.write_ply_3d(tmp_path / "a.ply", cloud)
    # an ascii file with an extra non-float property and an extra element, like exporters write
    rows = np.random.default_rng(0).normal(size=(5, 14)).astype(np.float32)
    names = ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "f_rest_0", "f_rest_31", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1"]
    with open(tmp_path / "b.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by a test\nelement vertex 5\n")
        for nm in names:
            f.write(f"property float {nm}\n")
        f.write("property float rot_2\nproperty float rot_3\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
        for r in rows:
            f.write(" ".join(repr(float(v)) for v in r) + " 0.5 -0.25\n")
        f.write("3 0 1 2\n")
    for name in ("a.ply", "b.ply"):
        r = subprocess.run([tool, str(tmp_path / name), str(tmp_path / "out.bin")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        raw = (tmp_path / "out.bin").read_bytes()
        n = int(np.frombuffer(raw, "<u8", 1)[0])
        py = bio.parse_ply_3d(tmp_path / name)
        assert n == len(py) and n % 32 == 0
        off = 8
        for w, want, exact in ((4, py.position_visibility, True), (48, py.spherical_harmonic, True), (4, py.rotation, False),
                               (4, py.scale_opacity, False)):
            got = np.frombuffer(raw, "<f4", n * w, off).reshape(n, w); off += n * w * 4
            if exact:
                assert np.array_equal(got, want)
            else:
                assert np.allclose(got, want, rtol=3e-7, atol=1e-12, equal_nan=True)
    bad = (tmp_path / "a.ply").read_bytes().replace(b"property float opacity\n", b"property float opacitx\n")
    (tmp_path / "c.ply").write_bytes(bad)
    r = subprocess.run([tool, str(tmp_path / "c.ply"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 2 and "missing required properties" in r.stderr
