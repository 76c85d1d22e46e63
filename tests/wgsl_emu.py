"""A WGSL-literal emulator: an INDEPENDENT second restatement of the reference's per-gaussian maths.

TEST INFRASTRUCTURE.  The CUDA kernels and oracle/bgs_oracle.cpp share one author and one reading of the
WGSL; a shared misreading would pass every CUDA-vs-oracle test.  This module does
not read either of them: it gives Python the WGSL value semantics -- vecN, matCxR with the COLUMN-major
constructor, `M * v`, `v * M`, `M * M`, `transpose`, `M[i]` = column i -- and then transliterates the
reference functions line by line, keeping the reference's own variable names and expression order.
Everything is evaluated in float64, so agreement with the f32 oracle is expected to ~1e-5 relative.

Transliterated (reference file:line):
  world_to_clip, in_frustum             src/render/transform.wgsl:5-14
  get_rotation_matrix, get_scale_matrix src/render/helpers.wgsl:137-168
  compute_cov3d                         src/render/gaussian_3d.wgsl:49-72
  cov2d                                 src/render/helpers.wgsl:8-47
  get_bounding_box_clip (OBB and AABB)  src/render/helpers.wgsl:49-119
  intrinsic_matrix                      src/render/helpers.wgsl:122-135
  world_to_local_direction              src/render/gaussian.wgsl:166-183
  spherical_harmonics_lookup, srgb      src/material/spherical_harmonics.wgsl:3-68
  compute_cov2d_surfel, get_bounding_box_cov2d, surfel_fragment_power   src/render/gaussian_2d.wgsl:49-156
  vs_points / fs_main                   src/render/gaussian.wgsl:185-505
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------- WGSL value types
class Vec:
    __array_priority__ = 100

    def __init__(self, *c):
        flat = []
        for x in c:
            if isinstance(x, Vec):
                flat.extend(x.v.tolist())
            elif isinstance(x, (list, tuple, np.ndarray)):
                flat.extend([float(y) for y in x])
            else:
                flat.append(float(x))
        self.v = np.array(flat, np.float64)

    # swizzles used by the shaders
    x = property(lambda s: s.v[0]); y = property(lambda s: s.v[1]); z = property(lambda s: s.v[2]); w = property(lambda s: s.v[3])
    xy = property(lambda s: Vec(s.v[0], s.v[1])); xyz = property(lambda s: Vec(s.v[0], s.v[1], s.v[2]))
    zw = property(lambda s: Vec(s.v[2], s.v[3]))

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i]

    def _bin(self, o, f):
        if isinstance(o, Mat):
            return NotImplemented
        ov = o.v if isinstance(o, Vec) else float(o)
        return Vec(f(self.v, ov))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return Vec(float(o) - self.v)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return Vec(float(o) / self.v)
    def __neg__(self): return Vec(-self.v)

    def __mul__(self, o):
        if isinstance(o, Mat):                      # v * M : row vector times matrix -> component i = dot(v, M[i])
            assert len(self.v) == o.rows
            return Vec([float(np.dot(self.v, o.col(i).v)) for i in range(o.cols)])
        return self._bin(o, np.multiply)            # component-wise / scalar

    def __rmul__(self, o):
        return Vec(float(o) * self.v)


class Mat:
    """matCxR: C columns of R-vectors.  Mat(cols=C, rows=R, values...) takes scalars in WGSL constructor order
    (column by column) or C column vectors."""
    __array_priority__ = 200

    def __init__(self, cols, rows, *vals):
        self.cols, self.rows = cols, rows
        if len(vals) == cols and all(isinstance(v, Vec) for v in vals):
            cs = [v.v for v in vals]
        else:
            flat = [float(v) for v in vals]
            assert len(flat) == cols * rows
            cs = [np.array(flat[i * rows:(i + 1) * rows], np.float64) for i in range(cols)]
        for c in cs:
            assert len(c) == rows
        self.c = [np.array(c, np.float64) for c in cs]

    def col(self, i):
        return Vec(self.c[i])

    def __getitem__(self, i):       # M[i] is column i; M[i][j] is column i, row j
        return Vec(self.c[i])

    def _a(self):                   # ordinary (row, col) array
        return np.stack(self.c, axis=1)

    @staticmethod
    def _from_a(a):
        return Mat(a.shape[1], a.shape[0], *[Vec(a[:, i]) for i in range(a.shape[1])])

    def __mul__(self, o):
        if isinstance(o, Mat):      # (R x C) * (C x K)
            assert self.cols == o.rows
            return Mat._from_a(self._a() @ o._a())
        if isinstance(o, Vec):      # M * v : column vector
            assert self.cols == len(o.v)
            return Vec(self._a() @ o.v)
        return Mat._from_a(self._a() * float(o))

    def __rmul__(self, o):
        return Mat._from_a(self._a() * float(o))

    def set(self, i, j, val):       # cov[i][j] = ...
        self.c[i][j] = val


def mat3x3(*v): return Mat(3, 3, *v)
def mat2x2(*v): return Mat(2, 2, *v)
def mat3x4(*v): return Mat(3, 4, *v)
def mat4x4(*v): return Mat(4, 4, *v)
def transpose(m): return Mat._from_a(m._a().T)
def dot(a, b): return float(np.dot(a.v, b.v))
def length(a): return math.sqrt(dot(a, a))
def normalize(a): return a / length(a)
def cross(a, b): return Vec(np.cross(a.v, b.v))
def vmax(a, b): return Vec(np.maximum(a.v, b.v if isinstance(b, Vec) else b))
def vsqrt(a): return Vec(np.sqrt(a.v))


# ---------------------------------------------------------------- uniforms
class ViewU:
    """The Bevy `View` uniform fields the shaders read (bindings.wgsl:3-9), from the column-major float[16]s of bgs_view."""

    def __init__(self, abi_view):
        def m(a):
            f = [float(x) for x in a]
            return mat4x4(*f)       # float[16] column-major == the WGSL constructor order
        self.view_from_world = m(abi_view.view_from_world)
        self.clip_from_view = m(abi_view.clip_from_view)
        self.clip_from_world = m(abi_view.clip_from_world)
        self.unjittered_clip_from_world = self.clip_from_world
        self.world_position = Vec(*[float(x) for x in abi_view.world_position])
        self.viewport = Vec(*[float(x) for x in abi_view.viewport])


class CloudU:
    def __init__(self, abi_uniform):
        self.transform = mat4x4(*[float(x) for x in abi_uniform.transform])
        self.global_opacity = float(abi_uniform.global_opacity)
        self.global_scale = float(abi_uniform.global_scale)
        self.color_space = int(abi_uniform.color_space)
        self.min = Vec(*[float(x) for x in abi_uniform.aabb_min])
        self.max = Vec(*[float(x) for x in abi_uniform.aabb_max])


class Shader:
    """One pipeline specialisation (shader defs) bound to a view + cloud uniform."""

    def __init__(self, view: ViewU, gaussian_uniforms: CloudU, use_obb=True, gaussian_2d=False, adaptive=True,
                 rasterize="color"):
        self.view, self.gu = view, gaussian_uniforms
        self.USE_OBB, self.USE_AABB = use_obb, not use_obb
        self.GAUSSIAN_2D = gaussian_2d
        self.OPACITY_ADAPTIVE_RADIUS = adaptive
        self.rasterize = rasterize

    # transform.wgsl:5-14
    def world_to_clip(self, world_pos):
        homogenous_pos = self.view.unjittered_clip_from_world * Vec(world_pos, 1.0)
        return homogenous_pos / (homogenous_pos.w + 0.000000001)

    @staticmethod
    def in_frustum(clip_space_pos):
        return abs(clip_space_pos.x) < 1.1 and abs(clip_space_pos.y) < 1.1 and abs(clip_space_pos.z - 0.5) < 0.5

    # helpers.wgsl:137-158
    @staticmethod
    def get_rotation_matrix(rotation):
        r = rotation.x; x = rotation.y; y = rotation.z; z = rotation.w
        return mat3x3(
            1.0 - 2.0 * (y * y + z * z),
            2.0 * (x * y - r * z),
            2.0 * (x * z + r * y),

            2.0 * (x * y + r * z),
            1.0 - 2.0 * (x * x + z * z),
            2.0 * (y * z - r * x),

            2.0 * (x * z - r * y),
            2.0 * (y * z + r * x),
            1.0 - 2.0 * (x * x + y * y),
        )

    # helpers.wgsl:160-168
    def get_scale_matrix(self, scale):
        gs = self.gu.global_scale
        return mat3x3(
            scale.x * gs, 0.0, 0.0,
            0.0, scale.y * gs, 0.0,
            0.0, 0.0, scale.z * gs,
        )

    # gaussian_3d.wgsl:49-72
    def compute_cov3d(self, scale, rotation):
        S = self.get_scale_matrix(scale)
        T = mat3x3(self.gu.transform[0].xyz, self.gu.transform[1].xyz, self.gu.transform[2].xyz)
        R = self.get_rotation_matrix(rotation)
        M = S * R
        Sigma = transpose(M) * M
        TS = T * Sigma * transpose(T)
        return [TS[0][0], TS[0][1], TS[0][2], TS[1][1], TS[1][2], TS[2][2]]

    # helpers.wgsl:8-47
    def cov2d(self, position, cov3d):
        view = self.view
        Vrk = mat3x3(
            cov3d[0], cov3d[1], cov3d[2],
            cov3d[1], cov3d[3], cov3d[4],
            cov3d[2], cov3d[4], cov3d[5],
        )
        t = view.view_from_world * Vec(position, 1.0)
        focal = Vec(view.clip_from_view[0].x * view.viewport.z, view.clip_from_view[1].y * view.viewport.w)
        s = 1.0 / (t.z * t.z)
        J = mat3x3(
            focal.x / t.z, 0.0, -(focal.x * t.x) * s,
            0.0, -focal.y / t.z, (focal.y * t.y) * s,
            0.0, 0.0, 0.0,
        )
        W = transpose(mat3x3(view.view_from_world[0].xyz, view.view_from_world[1].xyz, view.view_from_world[2].xyz))
        T = W * J
        cov = transpose(T) * transpose(Vrk) * T
        cov.set(0, 0, cov[0][0] + 0.3)
        cov.set(1, 1, cov[1][1] + 0.3)
        return Vec(cov[0][0], cov[0][1], cov[1][1])

    # helpers.wgsl:49-119
    def get_bounding_box_clip(self, cov2d, direction, cutoff):
        view = self.view
        det = cov2d.x * cov2d.z - cov2d.y * cov2d.y
        trace = cov2d.x + cov2d.z
        mid = 0.5 * trace
        discriminant = max(0.0, mid * mid - det)
        term = math.sqrt(discriminant)
        lambda1 = mid + term
        lambda2 = max(mid - term, 0.0)
        x_axis_length = math.sqrt(lambda1)
        y_axis_length = math.sqrt(lambda2)
        if self.USE_AABB:
            radius_px = cutoff * max(x_axis_length, y_axis_length)
            radius_ndc = Vec(radius_px / view.viewport.zw)
            return Vec(radius_ndc * direction, radius_px * direction)
        a = (cov2d.x - cov2d.z) * (cov2d.x - cov2d.z)
        b = math.sqrt(a + 4.0 * cov2d.y * cov2d.y)
        major_radius = math.sqrt((cov2d.x + cov2d.z + b) * 0.5)
        minor_radius = math.sqrt((cov2d.x + cov2d.z - b) * 0.5)
        bounds = cutoff * Vec(major_radius, minor_radius)
        eigvec1 = normalize(Vec(-cov2d.y, lambda1 - cov2d.x))
        eigvec2 = Vec(eigvec1.y, -eigvec1.x)
        rotation_matrix = transpose(mat2x2(eigvec1, eigvec2))
        scaled_vertex = direction * bounds
        rotated_vertex = scaled_vertex * rotation_matrix
        scaling_factor = 1.0 / view.viewport.zw
        ndc_vertex = rotated_vertex * scaling_factor
        return Vec(ndc_vertex, rotated_vertex)

    # helpers.wgsl:122-135
    def intrinsic_matrix(self):
        view = self.view
        focal = Vec(view.clip_from_view[0].x * view.viewport.z / 2.0, view.clip_from_view[1].y * view.viewport.w / 2.0)
        return mat3x4(
            Vec(focal.x, 0.0, 0.0, (view.viewport.z - 1.0) / 2.0),
            Vec(0.0, focal.y, 0.0, (view.viewport.w - 1.0) / 2.0),
            Vec(0.0, 0.0, 0.0, 1.0),
        )

    # gaussian.wgsl:166-183
    @staticmethod
    def world_to_local_direction(ray_direction_world, transform):
        basis = mat3x3(transform[0].xyz, transform[1].xyz, transform[2].xyz)
        basis_x = normalize(basis[0]); basis_y = normalize(basis[1]); basis_z = normalize(basis[2])
        local = Vec(dot(basis_x, ray_direction_world), dot(basis_y, ray_direction_world), dot(basis_z, ray_direction_world))
        return normalize(local)

    # spherical_harmonics.wgsl:3-68
    shc = [0.28209479177387814, -0.4886025119029199, 0.4886025119029199, -0.4886025119029199, 1.0925484305920792,
           -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396, -0.5900435899266435,
           2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
           -0.5900435899266435]

    @staticmethod
    def srgb_to_linear(srgb_color):
        out = []
        for i in range(3):
            if srgb_color[i] <= 0.04045:
                out.append(srgb_color[i] / 12.92)
            else:
                out.append(math.pow((srgb_color[i] + 0.055) / 1.055, 2.4))
        return Vec(out)

    def spherical_harmonics_lookup(self, ray_direction, sh):
        shc = self.shc
        rds = ray_direction * ray_direction
        color = Vec(0.5, 0.5, 0.5)
        color = color + shc[0] * Vec(sh[0], sh[1], sh[2])
        color = color + shc[1] * Vec(sh[3], sh[4], sh[5]) * ray_direction.y
        color = color + shc[2] * Vec(sh[6], sh[7], sh[8]) * ray_direction.z
        color = color + shc[3] * Vec(sh[9], sh[10], sh[11]) * ray_direction.x
        color = color + shc[4] * Vec(sh[12], sh[13], sh[14]) * ray_direction.x * ray_direction.y
        color = color + shc[5] * Vec(sh[15], sh[16], sh[17]) * ray_direction.y * ray_direction.z
        color = color + shc[6] * Vec(sh[18], sh[19], sh[20]) * (2.0 * rds.z - rds.x - rds.y)
        color = color + shc[7] * Vec(sh[21], sh[22], sh[23]) * ray_direction.x * ray_direction.z
        color = color + shc[8] * Vec(sh[24], sh[25], sh[26]) * (rds.x - rds.y)
        color = color + shc[9] * Vec(sh[27], sh[28], sh[29]) * ray_direction.y * (3.0 * rds.x - rds.y)
        color = color + shc[10] * Vec(sh[30], sh[31], sh[32]) * ray_direction.x * ray_direction.y * ray_direction.z
        color = color + shc[11] * Vec(sh[33], sh[34], sh[35]) * ray_direction.y * (4.0 * rds.z - rds.x - rds.y)
        color = color + shc[12] * Vec(sh[36], sh[37], sh[38]) * ray_direction.z * (2.0 * rds.z - 3.0 * rds.x - 3.0 * rds.y)
        color = color + shc[13] * Vec(sh[39], sh[40], sh[41]) * ray_direction.x * (4.0 * rds.z - rds.x - rds.y)
        color = color + shc[14] * Vec(sh[42], sh[43], sh[44]) * ray_direction.z * (rds.x - rds.y)
        color = color + shc[15] * Vec(sh[45], sh[46], sh[47]) * ray_direction.x * (rds.x - 3.0 * rds.y)
        return color

    # planar.wgsl:91-106
    def get_color(self, sh, ray_direction):
        color = self.spherical_harmonics_lookup(ray_direction, sh)
        if self.gu.color_space == 1:
            return color
        return self.srgb_to_linear(color)

    # gaussian_2d.wgsl:49-75
    def get_bounding_box_cov2d(self, extent, direction, cutoff):
        filter_size = 0.707106
        if extent.x < 1.0e-4 or extent.y < 1.0e-4:
            return Vec(0.0, 0.0, 0.0, 0.0)
        radius = vsqrt(extent)
        m = max(max(radius.x, radius.y), cutoff * filter_size)
        max_radius = Vec(m, m)
        radius_ndc = Vec(max_radius / self.view.viewport.zw)
        return Vec(radius_ndc * direction, max_radius)

    # gaussian_2d.wgsl:77-132
    def compute_cov2d_surfel(self, gaussian_position, rotation, scale, cutoff):
        gu, view = self.gu, self.view
        T_r = mat3x3(gu.transform[0].xyz, gu.transform[1].xyz, gu.transform[2].xyz)
        S = self.get_scale_matrix(scale)
        R = self.get_rotation_matrix(rotation)
        L = T_r * transpose(R) * S
        world_from_local = mat3x4(Vec(L[0], 0.0), Vec(L[1], 0.0), Vec(gaussian_position, 1.0))
        ndc_from_world = transpose(view.clip_from_world)
        pixels_from_ndc = self.intrinsic_matrix()
        T = transpose(world_from_local) * ndc_from_world * pixels_from_ndc
        test = Vec(cutoff * cutoff, cutoff * cutoff, -1.0)
        d = dot(test * T[2], T[2])
        if abs(d) < 1.0e-4:
            return dict(extent=Vec(0.0, 0.0), local_to_pixel=None, mean_2d=None)
        f = (1.0 / d) * test
        mean_2d = Vec(dot(f, T[0] * T[2]), dot(f, T[1] * T[2]))
        t = Vec(dot(f * T[0], T[0]), dot(f * T[1], T[1]))
        extent = mean_2d * mean_2d - t
        return dict(local_to_pixel=T, mean_2d=mean_2d, extent=extent)

    # gaussian_2d.wgsl:134-156
    @staticmethod
    def surfel_fragment_power(local_to_pixel, pixel_coord, mean_2d):
        deltas = mean_2d - pixel_coord
        hu = pixel_coord.x * local_to_pixel[2] - local_to_pixel[0]
        hv = pixel_coord.y * local_to_pixel[2] - local_to_pixel[1]
        p = cross(hu, hv)
        us = p.x / p.z
        vs = p.y / p.z
        sigmas_3d = us * us + vs * vs
        sigmas_2d = 2.0 * (deltas.x * deltas.x + deltas.y * deltas.y)
        sigmas = 0.5 * min(sigmas_3d, sigmas_2d)
        return -sigmas

    # ---- vs_points, gaussian.wgsl:185-436, for one gaussian: the four emitted vertices + the flat varyings
    def vs_points(self, position3, sh, rotation, scale_opacity, key_is_culled=False, visibility=1.0,
                  draw_selected=False, highlight_selected=False):
        gu, view = self.gu, self.view
        position = Vec(position3, 1.0)
        transformed_position = (gu.transform * position).xyz
        discard_quad = key_is_culled
        if draw_selected:
            discard_quad = discard_quad or visibility < 0.5
        projected_position = self.world_to_clip(transformed_position)
        discard_quad = discard_quad or not self.in_frustum(projected_position.xyz)
        if discard_quad:
            return None
        quad_vertices = [Vec(-1.0, -1.0), Vec(-1.0, 1.0), Vec(1.0, -1.0), Vec(1.0, 1.0)]
        opacity = float(scale_opacity[3])
        scale = Vec(scale_opacity[0], scale_opacity[1], scale_opacity[2])
        if self.OPACITY_ADAPTIVE_RADIUS:
            lg = math.log(opacity) if opacity > 0.0 else -math.inf
            cutoff = math.sqrt(max(9.0 + 2.0 * lg, 0.000001))
        else:
            cutoff = 3.0
        out = dict(cutoff=cutoff, projected_position=projected_position, transformed_position=transformed_position)
        verts = []
        if self.GAUSSIAN_2D:
            surfel = self.compute_cov2d_surfel(transformed_position, rotation, scale, cutoff)
            out["surfel"] = surfel
            for q in quad_vertices:
                bb = self.get_bounding_box_cov2d(surfel["extent"], q, cutoff)
                verts.append((q, bb))
            out["radius"] = verts[0][1].zw
        else:
            cov3d = self.compute_cov3d(scale, rotation)
            gaussian_cov2d = self.cov2d(transformed_position, cov3d)
            out["cov2d"] = gaussian_cov2d
            for q in quad_vertices:
                bb = self.get_bounding_box_clip(gaussian_cov2d, q, cutoff)
                verts.append((q, bb))
            if self.USE_AABB:
                det = gaussian_cov2d.x * gaussian_cov2d.z - gaussian_cov2d.y * gaussian_cov2d.y
                det_inv = 1.0 / det
                out["conic"] = Vec(gaussian_cov2d.z * det_inv, -gaussian_cov2d.y * det_inv, gaussian_cov2d.x * det_inv)
        rgb = Vec(0.0, 0.0, 0.0)
        if self.rasterize == "color":
            ray_direction_world = normalize(transformed_position - view.world_position)
            ray_direction_local = self.world_to_local_direction(ray_direction_world, gu.transform)
            rgb = self.get_color(sh, ray_direction_local)
        elif self.rasterize == "normal":
            R = self.get_rotation_matrix(rotation)
            S = self.get_scale_matrix(scale)
            T = mat3x3(gu.transform[0].xyz, gu.transform[1].xyz, gu.transform[2].xyz)
            L = T * S * R
            local_normal = Vec(L[2], 0.0)
            world_normal = view.view_from_world * local_normal
            t = normalize(world_normal)
            rgb = Vec(0.5 * (t.x + 1.0), 0.5 * (t.y + 1.0), 0.5 * (t.z + 1.0))
        elif self.rasterize == "depth":
            # gaussian.wgsl:329-349 -- the range comes from two entries of the SORTED buffer, literally entry 1 and entry
            # count - 1 (`self.depth_entries` = their untransformed positions, set by the caller from the sort it feeds in)
            first_position, last_position = self.depth_entries
            min_position = (gu.transform * Vec(last_position, 1.0)).xyz
            max_position = (gu.transform * Vec(first_position, 1.0)).xyz
            camera_position = view.world_position
            min_distance = length(min_position - camera_position)
            max_distance = length(max_position - camera_position)
            depth = length(transformed_position - camera_position)
            rgb = self.depth_to_rgb(depth, min_distance, max_distance)
        elif self.rasterize == "position":
            rgb = (transformed_position - gu.min.xyz) / (gu.max.xyz - gu.min.xyz)
        color = Vec(rgb, opacity * gu.global_opacity)
        if highlight_selected and visibility > 0.5:
            color = Vec(0.3, 1.0, 0.1, 1.0)
        out["color"] = color
        out["vertices"] = [dict(uv=q, position=Vec(projected_position.xy + bb.xy, projected_position.zw), bb=bb) for q, bb in verts]
        return out

    # ---- material/depth.wgsl:3-11 (clamp, smoothstep: WGSL built-ins; smoothstep(e0, e1, x) = t*t*(3 - 2t), t = clamp((x-e0)/(e1-e0), 0, 1))
    @staticmethod
    def depth_to_rgb(depth, min_depth, max_depth):
        def clamp(x, lo, hi):
            return min(max(x, lo), hi)

        def smoothstep(e0, e1, x):
            t = clamp((x - e0) / (e1 - e0), 0.0, 1.0)
            return t * t * (3.0 - 2.0 * t)

        normalized_depth = clamp((depth - min_depth) / (max_depth - min_depth), 0.0, 1.0)
        r = smoothstep(0.5, 1.0, normalized_depth)
        g = 1.0 - abs(normalized_depth - 0.5) * 2.0
        b = 1.0 - smoothstep(0.0, 0.5, normalized_depth)
        return Vec(r, g, b)

    # ---- fs_main, gaussian.wgsl:438-505: the fragment's premultiplied colour, or None for `discard`
    def fs_main(self, vs, uv, major_minor=None):
        color = vs["color"]
        if self.USE_AABB:
            if self.GAUSSIAN_2D:
                radius = vs["radius"]
                mean_2d = vs["surfel"]["mean_2d"]
                aspect = Vec(1.0, self.view.viewport.z / self.view.viewport.w)
                pixel_coord = uv * radius * aspect + mean_2d
                power = self.surfel_fragment_power(vs["surfel"]["local_to_pixel"], pixel_coord, mean_2d)
            else:
                d = -major_minor
                conic = vs["conic"]
                power = -0.5 * (conic.x * d.x * d.x + conic.z * d.y * d.y) + conic.y * d.x * d.y
            if power > 0.0:
                return None
        if self.USE_OBB:
            sigma = 1.0 / 3.0
            sigma_squared = 2.0 * sigma * sigma
            distance_squared = dot(uv, uv)
            power = -distance_squared / sigma_squared
            if distance_squared > 3.0 * 3.0:
                return None
        alpha = min(math.exp(power) * color[3], 0.999)
        return Vec(color[0] * alpha, color[1] * alpha, color[2] * alpha, alpha)


# ---------------------------------------------------------------- a tiny rasteriser of the emitted quads
def ndc_to_pixel(ndc_xy, W, H):
    """Viewport transform of the fixed-function pipeline: x right, y DOWN, pixel centres at +0.5."""
    return np.array([(ndc_xy[0] * 0.5 + 0.5) * W, (1.0 - (ndc_xy[1] * 0.5 + 0.5)) * H])


def render_reference_semantics(shader: Shader, cloud, order, W, H, draw_selected=False, highlight_selected=False):
    """Draw the instanced quads of `order` (far -> near, the sorted entry buffer) the way the GPU pipeline would:
    each quad is an affine patch (triangle strip of 4 vertices whose uv / major_minor varyings interpolate linearly),
    a fragment is generated for every pixel centre inside it, fs_main gives its premultiplied colour, and the target
    blends dst = src + (1 - src.a) * dst (render/mod.rs:944-948) over a cleared (0, 0, 0, 1) target."""
    img = np.zeros((H, W, 4), np.float64)
    img[..., 3] = 1.0
    for gi in order:
        pv = cloud.position_visibility[gi]
        vs = shader.vs_points(Vec(pv[0], pv[1], pv[2]), [float(x) for x in cloud.spherical_harmonic[gi]],
                              Vec(*[float(x) for x in cloud.rotation[gi]]), [float(x) for x in cloud.scale_opacity[gi]],
                              visibility=float(pv[3]), draw_selected=draw_selected, highlight_selected=highlight_selected)
        if vs is None:
            continue
        P = [ndc_to_pixel(v["position"].xy.v, W, H) for v in vs["vertices"]]      # uv = (-1,-1), (-1,1), (1,-1), (1,1)
        # affine patch: pixel(uv) = c + A uv, from three of the corners
        c = 0.25 * (P[0] + P[1] + P[2] + P[3])
        A = np.stack([0.5 * (P[2] - P[0]), 0.5 * (P[1] - P[0])], axis=1)
        if not np.all(np.isfinite(A)) or abs(np.linalg.det(A)) < 1e-300:
            continue
        Ainv = np.linalg.inv(A)
        lo = np.floor(np.min(P, axis=0)).astype(int) - 1
        hi = np.ceil(np.max(P, axis=0)).astype(int) + 1
        mm_corner = [v["bb"].zw for v in vs["vertices"]]
        for y in range(max(lo[1], 0), min(hi[1], H - 1) + 1):
            for x in range(max(lo[0], 0), min(hi[0], W - 1) + 1):
                uv = Ainv @ (np.array([x + 0.5, y + 0.5]) - c)
                if abs(uv[0]) > 1.0 or abs(uv[1]) > 1.0:
                    continue                                                     # outside the quad: no fragment
                mm = None
                if shader.USE_AABB and not shader.GAUSSIAN_2D:
                    # major_minor = bb.zw is linear in uv (radius_px * direction)
                    mm = Vec(mm_corner[3].x * uv[0], mm_corner[3].y * uv[1])
                src = shader.fs_main(vs, Vec(uv[0], uv[1]), mm)
                if src is None:
                    continue
                img[y, x] = src.v + (1.0 - src.v[3]) * img[y, x]
    return img
