"""Python host mirror of the plugin surface for the forward splat path.

Reference: `GaussianSplattingPlugin` (src/lib.rs:48-80) wires, for this path, the render-world
system `run_radix_sort` (src/sort/radix.rs:616-756) and the `DrawGaussians` render command
(src/render/mod.rs:986-992,1501-1569), both invoked once per `GaussianCamera` view per frame over
entities holding `(PlanarGaussian3dHandle, CloudSettings)`.  Here the same roles exist with the
same names, but the per-view work is ONE call across the C ABI (`bgs_render`).

This module is test/bench plumbing above the ABI (the production host stays Rust, see
INTEGRATION.md); it never computes anything itself and has no fallback path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import abi
from .camera import GaussianCamera, View
from .gaussian import PlanarGaussian3d
from .settings import CloudSettings


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclasses.dataclass
class CloudTransform:
    """GlobalTransform of the cloud entity -> CloudUniform.transform (render/mod.rs:1056-1072)."""

    matrix: np.ndarray = dataclasses.field(default_factory=lambda: np.eye(4, dtype=np.float32))


class PlanarGaussian3dHandle:
    """A cloud resident in HBM (the role of `Handle<PlanarGaussian3d>` + its prepared GPU planes)."""

    _next_serial = 0

    def __init__(self, plugin: "GaussianSplattingPlugin", cloud: PlanarGaussian3d, f16: bool = False,
                 precompute_covariance: bool = False):
        PlanarGaussian3dHandle._next_serial += 1
        self.serial = PlanarGaussian3dHandle._next_serial   # never reused (id() is, once a handle is collected)
        self._plugin = plugin
        self._lib = plugin._lib
        self.n = len(cloud)
        self.f16 = f16
        self.aabb = cloud.compute_aabb()        # the entity's Aabb (calculate_bounds, src/gaussian/cloud.rs:45-62)
        self._h = C.c_void_p()
        self.precompute_covariance = precompute_covariance
        if precompute_covariance:
            # the reference's `precompute_covariance_3d` feature: Covariance3dOpacityPacked128 in the second plane
            sh_p, cov_op = cloud.precomputed_covariance().pack_f16()
            st = self._lib.bgs_cloud_upload_f16_cov(plugin._ctx, self.n, _ptr(cloud.position_visibility), _ptr(sh_p),
                                                    _ptr(cov_op), C.byref(self._h))
        elif f16:
            sh_p, rso = cloud.pack_f16()
            st = self._lib.bgs_cloud_upload_f16(plugin._ctx, self.n, _ptr(cloud.position_visibility), _ptr(sh_p),
                                                _ptr(rso), C.byref(self._h))
        else:
            st = self._lib.bgs_cloud_upload_f32(plugin._ctx, self.n, _ptr(cloud.position_visibility),
                                                _ptr(cloud.spherical_harmonic), _ptr(cloud.rotation),
                                                _ptr(cloud.scale_opacity), C.byref(self._h))
        plugin._check(st)

    def destroy(self):
        if self._h:
            self._lib.bgs_cloud_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class GaussianSplattingPlugin:
    """Owns one `bgs_context` (one GPU).  `render_view` = run_radix_sort + DrawGaussians for one view."""

    FORMATS = {"rgba8_srgb": (abi.BGS_FORMAT_RGBA8_SRGB, np.uint8, 4), "rgba16f": (abi.BGS_FORMAT_RGBA16F, np.float16, 4),
               "rgba32f": (abi.BGS_FORMAT_RGBA32F, np.float32, 4)}

    def __init__(self, cuda_device: int = 0):
        self._lib = abi.load()
        self._ctx = C.c_void_p()
        st = self._lib.bgs_context_create(cuda_device, C.byref(self._ctx))
        if st != abi.BGS_OK:
            raise abi.BgsError(st, f"bgs_context_create(device={cuda_device}) failed (no usable CUDA device?)")
        self.device = cuda_device

    # -- resources
    def add_cloud(self, cloud: PlanarGaussian3d, f16: bool = False, precompute_covariance: bool = False) -> PlanarGaussian3dHandle:
        return PlanarGaussian3dHandle(self, cloud, f16, precompute_covariance)

    def _check(self, st: int):
        if st != abi.BGS_OK:
            raise abi.BgsError(st, (self._lib.bgs_last_error(self._ctx) or b"").decode())

    @staticmethod
    def cloud_uniform(settings: CloudSettings, transform: CloudTransform | None = None, aabb=None) -> abi.bgs_cloud_uniform:
        """extract_gaussians (src/render/mod.rs:1056-1072); `aabb` = (min, max) of the entity's Aabb."""
        u = abi.bgs_cloud_uniform()
        lo, hi = aabb if aabb is not None else (np.zeros(3, np.float32), np.ones(3, np.float32))
        u.aabb_min[:] = [float(lo[0]), float(lo[1]), float(lo[2]), 1.0]
        u.aabb_max[:] = [float(hi[0]), float(hi[1]), float(hi[2]), 1.0]
        m = (transform.matrix if transform is not None else np.eye(4, dtype=np.float32)).astype(np.float32)
        u.transform[:] = m.T.reshape(-1).tolist()
        u.global_opacity = settings.global_opacity
        u.global_scale = settings.global_scale
        u.color_space = int(settings.color_space)
        u.time = settings.time
        return u

    # -- the per-view, per-frame call
    def render_view(self, handle: PlanarGaussian3dHandle, settings: CloudSettings, view: View,
                    camera: GaussianCamera | None = None, transform: CloudTransform | None = None,
                    fmt: str = "rgba32f", out: np.ndarray | None = None, to_host: bool = True,
                    asynchronous: bool = False, premultiplied: bool = False, blend_over: bool = False):
        """Returns the (H, W, 4) frame (host) or None when `to_host` is False / the camera is warming up.
        `asynchronous`: only enqueue the frame (BGS_FLAG_ASYNC); call `sync()` before reading anything.
        `premultiplied`: the splat layer alone, (C, 1 - T) (BGS_FLAG_PREMULTIPLIED_OUT).  `blend_over`: blend over what
        the context's frame already holds -- the previous call's result (BGS_FLAG_BLEND_OVER_TARGET): one call per
        cloud, far cloud first, like the reference's Transparent3d items (render/mod.rs:398-452, :944-948)."""
        if camera is not None and camera.warmup:   # queue_gaussians skips warm-up cameras (render/mod.rs:361-371)
            return None
        code, dtype, ch = self.FORMATS[fmt]
        v = view.to_abi()
        u, s = self._uniform_and_settings(handle, settings, transform, asynchronous, premultiplied, blend_over)
        if to_host:
            if out is None:
                out = np.empty((view.height, view.width, ch), dtype)
            assert out.dtype == dtype and out.size == view.height * view.width * ch and out.flags.c_contiguous
            st = self._lib.bgs_render(self._ctx, handle._h, C.byref(v), C.byref(u), C.byref(s), _ptr(out), code, 0)
        else:
            st = self._lib.bgs_render(self._ctx, handle._h, C.byref(v), C.byref(u), C.byref(s), None, code, 0)
        self._check(st)
        return out if to_host else None

    def _uniform_and_settings(self, handle, settings, transform, asynchronous=False, premultiplied=False, blend_over=False):
        """The two ABI structs of a call; cached while (settings, transform, cloud, flags) repeat from frame to frame."""
        key = (dataclasses.astuple(settings), None if transform is None else transform.matrix.tobytes(), asynchronous, handle.serial,
               premultiplied, blend_over)
        if getattr(self, "_us_cache", (None,))[0] != key:
            s_ = settings.to_abi()
            if asynchronous:
                s_.flags |= abi.BGS_FLAG_ASYNC
            if premultiplied:
                s_.flags |= abi.BGS_FLAG_PREMULTIPLIED_OUT
            if blend_over:
                s_.flags |= abi.BGS_FLAG_BLEND_OVER_TARGET
            self._us_cache = (key, self.cloud_uniform(settings, transform, handle.aabb), s_)
        return self._us_cache[1], self._us_cache[2]

    def render_view_aux(self, handle: PlanarGaussian3dHandle, settings: CloudSettings, view: View,
                        transform: CloudTransform | None = None, fmt: str = "rgba32f"):
        """Colour, depth and normal frames of one view in ONE pass (`bgs_render_aux`, BASELINE.json config 4)."""
        code, dtype, ch = self.FORMATS[fmt]
        v = view.to_abi()
        u = self.cloud_uniform(settings, transform, handle.aabb)
        s = settings.to_abi()
        outs = [np.empty((view.height, view.width, ch), dtype) for _ in range(3)]
        st = self._lib.bgs_render_aux(self._ctx, handle._h, C.byref(v), C.byref(u), C.byref(s), _ptr(outs[0]), _ptr(outs[1]),
                                      _ptr(outs[2]), code, 0)
        self._check(st)
        return outs

    def render_view_to_device(self, handle: PlanarGaussian3dHandle, settings: CloudSettings, view: View, device_ptr: int,
                              transform: CloudTransform | None = None, fmt: str = "rgba8_srgb", asynchronous: bool = False) -> None:
        """Render straight into caller-owned device memory: an exported frame target (`bgs_frame_export_create`), or
        another GPU's memory mapped into this process (`bgs_peer_buffer_open`) -- the blend kernel's pixel stores then
        travel over NVLink themselves."""
        code, _, _ = self.FORMATS[fmt]
        v = view.to_abi()
        u, s = self._uniform_and_settings(handle, settings, transform, asynchronous)
        self._check(self._lib.bgs_render(self._ctx, handle._h, C.byref(v), C.byref(u), C.byref(s), C.c_void_p(device_ptr), code, 1))

    def sync(self) -> bool:
        """Complete the frames enqueued with `asynchronous=True`.  False = the last frame must be rendered again
        (its pair list outgrew the buffer, which has been grown)."""
        st = self._lib.bgs_sync(self._ctx)
        if st == abi.BGS_NOT_READY:
            return False
        self._check(st)
        return True

    # -- parity / measurement hooks
    def frame_stats(self) -> abi.bgs_frame_stats:
        fs = abi.bgs_frame_stats()
        self._check(self._lib.bgs_frame_stats_get(self._ctx, C.byref(fs)))
        return fs

    def stage_times_us(self) -> np.ndarray:
        arr = (C.c_float * 6)()
        self._check(self._lib.bgs_stage_times_us(self._ctx, C.byref(arr)))
        return np.array(list(arr), np.float32)

    def sorted_entries(self) -> np.ndarray:
        n = self.frame_stats().n
        out = np.empty((n, 2), np.uint32)
        self._check(self._lib.bgs_debug_sorted_entries(self._ctx, _ptr(out)))
        return out

    def tile_ranges(self) -> np.ndarray:
        fs = self.frame_stats()
        out = np.empty((fs.tiles_x * fs.tiles_y, 2), np.uint32)
        self._check(self._lib.bgs_debug_tile_ranges(self._ctx, _ptr(out)))
        return out

    def tile_entries(self) -> np.ndarray:
        fs = self.frame_stats()
        out = np.empty((fs.n_pairs,), np.uint32)
        if fs.n_pairs:
            self._check(self._lib.bgs_debug_tile_entries(self._ctx, _ptr(out), fs.n_pairs))
        return out

    def projected(self):
        fs = self.frame_stats()
        rec = np.empty((fs.n_visible, 12), np.float32)
        ids = np.empty((fs.n_visible,), np.uint32)
        if fs.n_visible:
            self._check(self._lib.bgs_debug_projected(self._ctx, _ptr(rec), _ptr(ids)))
        return rec, ids

    @property
    def stream_ptr(self) -> int:
        return int(self._lib.bgs_context_stream(self._ctx) or 0)

    @property
    def copy_stream_ptr(self) -> int:
        return int(self._lib.bgs_context_copy_stream(self._ctx) or 0)

    @property
    def frame_device_ptr(self) -> int:
        return int(self._lib.bgs_frame_device_ptr(self._ctx) or 0)

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.bgs_last_launch_count(self._ctx))

    def destroy(self):
        if self._ctx:
            self._lib.bgs_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
