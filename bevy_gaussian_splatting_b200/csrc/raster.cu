// raster.cu -- stage 5: per-16x16-tile front-to-back alpha blend.
//
// Replaces the reference's instanced-quad draw + fixed-function ROP blending
// (vs_points/fs_main, src/render/gaussian.wgsl:185-505; PREMULTIPLIED_ALPHA_BLENDING applied
// far->near, src/render/mod.rs:944-948).  Per pixel, the same coverage rule (pixel centre inside
// the splat's OBB quad, |u|<=1 and |v|<=1), the same falloff (alpha = min(exp(-4.5|uv|^2) * o *
// g_o, 0.999), gaussian.wgsl:474-504) and the same "over" operator, evaluated front-to-back:
//   C = sum_j rgb_j a_j T_j,  T_j = prod_{k nearer}(1 - a_k);   out = C + T*background(=0), a=1.
// A pixel stops once T < 1e-4; a tile stops when all its pixels have stopped (block vote).
//
// One CTA per tile, 256 threads = 8 warps, each warp owning an 8x4-pixel sub-rectangle so a
// warp-uniform bbox test skips splats that cannot touch any of its 32 pixels.  The tile's slice
// of the sorted pair list is staged through shared memory in chunks of 256 records.
// Coverage maths uses explicit __fmul_rn/__fmaf_rn so u,v are bit-identical to the oracle.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace bgs {

constexpr int RT_THREADS = 256;
constexpr int RT_CHUNK = 256;

// ---- TMA (cp.async.bulk) staging of a tile's slice of the sorted pair list ------------------------------
// The slice [range.x, range.y) of tile_entries is contiguous, so each 256-entry chunk is brought into shared
// memory by ONE bulk async copy (UBLKCP) issued by one thread and tracked by an mbarrier; the next chunk's copy
// is issued before the current chunk is rasterised (double buffer).  Bulk copies need 16 B alignment: the copy
// starts at the slice address rounded down to 16 B and `lead` skips the extra leading entries.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
constexpr uint32_t ENT_WORDS = RT_CHUNK + 8;   // a chunk plus up to 3 leading + rounding words

// issue the bulk copy of entries [base, base + cnt) into buffer `buf`; returns nothing (thread 0 only)
__device__ __forceinline__ void issue_entries(const uint32_t* tile_entries, uint32_t base, uint32_t cnt, uint32_t a_ent,
                                              uint32_t a_bar, int buf) {
    const uint32_t lead = base & 3u;
    const uint32_t bytes = ((lead + cnt) * 4u + 15u) & ~15u;
    mbar_expect_tx(a_bar + 8u * buf, bytes);
    tma_bulk_g2s(a_ent + (uint32_t)buf * ENT_WORDS * 4u, tile_entries + (base - lead), bytes, a_bar + 8u * buf);
}

// CTA -> tile: rows are visited from the middle row outwards (mid, mid-1, mid+1, ...), so the tiles launched last --
// the ones that form the kernel's tail -- are the top / bottom rows, usually the lightest.
__device__ __forceinline__ void centre_out_tile(int b, int tiles_x, int tiles_y, int& tile_x, int& tile_y) {
    const int k = b / tiles_x, mid = tiles_y / 2;
    tile_x = b - k * tiles_x;
    tile_y = (k & 1) ? mid - (k + 1) / 2 : mid + k / 2;
}

__device__ __forceinline__ float linear_to_srgb(float c) {
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    // __powf = ex2.approx(lg2.approx(c) / 2.4): ~1e-6 relative, far below the 8-bit quantisation step (the full
    // powf was 7 % of the kernel's instructions)
    return c <= 0.0031308f ? 12.92f * c : 1.055f * __powf(c, 1.0f / 2.4f) - 0.055f;
}

// ---- frame output.  `format` = BGS_FORMAT_* | output mode << 8:
//   mode 0: the splat layer over an opaque black clear (examples/headless.rs:70): (C, 1)
//   mode 1 (BGS_FLAG_PREMULTIPLIED_OUT): the layer alone, premultiplied: (C, 1 - T)
//   mode 2 (BGS_FLAG_BLEND_OVER_TARGET): blended over what the target holds, dst = src + (1 - src.a) dst on all four
//           channels (PREMULTIPLIED_ALPHA_BLENDING, render/mod.rs:944-948): (C + T dst.rgb, (1 - T) + T dst.a)
constexpr uint32_t OUT_PREMUL = 1u, OUT_OVER = 2u;
__device__ __forceinline__ float srgb_decode(float c) {
    return c <= 0.04045f ? c * (1.0f / 12.92f) : __powf((c + 0.055f) * (1.0f / 1.055f), 2.4f);
}
__device__ __forceinline__ float4 read_pixel(const void* out, uint32_t fmt, size_t pix) {
    if (fmt == BGS_FORMAT_RGBA32F) return reinterpret_cast<const float4*>(out)[pix];
    if (fmt == BGS_FORMAT_RGBA16F) {
        const uint2 v = reinterpret_cast<const uint2*>(out)[pix];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    const uint32_t v = reinterpret_cast<const uint32_t*>(out)[pix];
    return make_float4(srgb_decode((float)(v & 255u) * (1.0f / 255.0f)), srgb_decode((float)((v >> 8) & 255u) * (1.0f / 255.0f)),
                       srgb_decode((float)((v >> 16) & 255u) * (1.0f / 255.0f)), (float)(v >> 24) * (1.0f / 255.0f));
}
// one pixel's accumulated premultiplied colour (r, g, b) and remaining transmittance T -> the frame
__device__ __forceinline__ void write_pixel(void* out, uint32_t format, size_t pix, float r, float g, float b, float T) {
    const uint32_t fmt = format & 0xFFu, mode = format >> 8;
    float a = 1.0f;
    if (mode & OUT_OVER) {
        const float4 d = read_pixel(out, fmt, pix);
        r = fmaf(T, d.x, r); g = fmaf(T, d.y, g); b = fmaf(T, d.z, b);
        a = fmaf(T, d.w, 1.0f - T);
    } else if (mode & OUT_PREMUL) {
        a = 1.0f - T;
    }
    if (fmt == BGS_FORMAT_RGBA32F) {
        reinterpret_cast<float4*>(out)[pix] = make_float4(r, g, b, a);
    } else if (fmt == BGS_FORMAT_RGBA16F) {
        const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(out)[pix] = o;
    } else {
        const uint32_t r8 = (uint32_t)(linear_to_srgb(r) * 255.0f + 0.5f);
        const uint32_t g8 = (uint32_t)(linear_to_srgb(g) * 255.0f + 0.5f);
        const uint32_t b8 = (uint32_t)(linear_to_srgb(b) * 255.0f + 0.5f);
        const uint32_t a8 = mode ? (uint32_t)(fminf(fmaxf(a, 0.0f), 1.0f) * 255.0f + 0.5f) : 255u;
        reinterpret_cast<uint32_t*>(out)[pix] = r8 | (g8 << 8) | (b8 << 16) | (a8 << 24);
    }
}

// shared-memory layout (byte offsets from one base so the hot loop needs a single address register)
constexpr uint32_t SM_Q0 = 0;                         // float4 [256]: cx, cy, ux, uy
constexpr uint32_t SM_UV = SM_Q0 + RT_CHUNK * 16;     // float4 [256]: vx, vy, bbox x, bbox y
constexpr uint32_t SM_Q2 = SM_UV + RT_CHUNK * 16;     // float4 [256]: r, g, b, opacity
constexpr uint32_t SM_LIST = SM_Q2 + RT_CHUNK * 16;   // u16 [8][256]: per-warp candidates, stored as index * 16
constexpr uint32_t SM_EXTRA = SM_LIST + (RT_THREADS / 32) * RT_CHUNK * 2;   // MODE 2 only: 4 x float4 [256]
constexpr uint32_t SM_BYTES = SM_EXTRA;
constexpr uint32_t SM_BYTES_2D = SM_EXTRA + 4 * RT_CHUNK * 16;
// AUX (bgs_render_aux): 2 x float4 [256] after the mode's own arrays: depth rgb, normal rgb of the staged splats


// MODE 0: USE_OBB quad-uv falloff (3DGS, and 2DGS without aabb)   gaussian.wgsl:474-504
// MODE 1: 3DGS USE_AABB conic falloff                              gaussian.wgsl:459-471
// MODE 2: 2DGS USE_AABB ray-splat intersection                     gaussian.wgsl:441-458, gaussian_2d.wgsl:134-156
// AUX: the same pass also blends the splats' Depth and Normal colour sources (aux records, 2 x float4 per splat) into two
// more frames with the very same alphas: config C4's colour + depth + normal outputs cost one pass, not three.
template <int MODE, bool AUX>
__global__ void __launch_bounds__(RT_THREADS, (MODE == 0 && !AUX) ? 6 : 5)
raster_kernel(const SplatRec* __restrict__ recs, const float4* __restrict__ extra, const uint32_t* __restrict__ tile_entries,
              const uint2* __restrict__ ranges, int W, int H, int tiles_x, void* __restrict__ out, uint32_t format,
              const float4* __restrict__ aux, void* __restrict__ out_depth, void* __restrict__ out_normal) {
    __shared__ __align__(16) unsigned char s_mem[(MODE == 2 ? SM_BYTES_2D : SM_BYTES) + (AUX ? 2 * RT_CHUNK * 16 : 0)];
    constexpr uint32_t SM_AUX = MODE == 2 ? SM_BYTES_2D : SM_BYTES;
    __shared__ __align__(16) uint32_t s_ent[2][ENT_WORDS];    // TMA destination: the tile's pair-list chunks
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(8) float2 s_thr[MODE == 0 ? RT_CHUNK : 1];   // MODE 0: per staged splat cull thresholds (u, v)
    float4* s_q0 = reinterpret_cast<float4*>(s_mem + SM_Q0);
    float4* s_uv = reinterpret_cast<float4*>(s_mem + SM_UV);
    float4* s_q2 = reinterpret_cast<float4*>(s_mem + SM_Q2);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_mem + SM_LIST) + warp * RT_CHUNK;
    const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(s_mem);
    const uint32_t a_list = a_base + SM_LIST + (uint32_t)warp * RT_CHUNK * 2u;
    int tile_x, tile_y;
    centre_out_tile((int)blockIdx.x, tiles_x, (int)gridDim.x / tiles_x, tile_x, tile_y);
    const int tile = tile_y * tiles_x + tile_x;
    // warp w covers the 8x4 rectangle at ((w & 1) * 8, (w >> 1) * 4) of the tile
    const int wx0 = tile_x * TILE_PX + (warp & 1) * 8, wy0 = tile_y * TILE_PX + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    const float rcx = (float)wx0 + 4.0f, rcy = (float)wy0 + 2.0f;   // centre of the warp's pixel centres
    const float tcx = (float)(tile_x * TILE_PX) + 8.0f, tcy = (float)(tile_y * TILE_PX) + 8.0f;   // tile centre
    uint2 range = ranges[tile];
    range.x = ~range.x;                  // stored as (~start, end): the sort's last pass builds it with atomicMax (radix.cu)

    const uint32_t a_ent = (uint32_t)__cvta_generic_to_shared(&s_ent[0][0]);
    const uint32_t a_bar = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
    if (range.x >= range.y) {            // empty tile: nothing to stage (uniform across the CTA)
        // (blend-over mode leaves the target's pixels as they are)
        if (inside && !((format >> 8) & OUT_OVER)) {
            write_pixel(out, format, (size_t)py * W + px, 0.f, 0.f, 0.f, 1.0f);
            if (AUX) {
                write_pixel(out_depth, format, (size_t)py * W + px, 0.f, 0.f, 0.f, 1.0f);
                write_pixel(out_normal, format, (size_t)py * W + px, 0.f, 0.f, 0.f, 1.0f);
            }
        }
        return;
    }
    // tiles with more than one chunk stream their pair list through the TMA double buffer (the next chunk's
    // copy overlaps this chunk's blending); single-chunk tiles read it directly (no barrier set-up on their path)
    const bool use_tma = range.y - range.x > (uint32_t)RT_CHUNK;
    uint32_t issued = 0u, chunk = 0u;
    if (use_tma) {
        if (t == 0) {
            mbar_init(a_bar, 1u); mbar_init(a_bar + 8u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (t == 0) issue_entries(tile_entries, range.x, (uint32_t)RT_CHUNK, a_ent, a_bar, 0);
        issued = 1u;
    }

    float T = inside ? 1.0f : 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;   // T < T_STOP <=> this pixel is done
    float dr = 0.0f, dg = 0.0f, db = 0.0f, nr = 0.0f, ng = 0.0f, nb = 0.0f;   // AUX: depth / normal frames
    for (uint32_t base = range.x; base < range.y; base += RT_CHUNK, ++chunk) {
        if (__syncthreads_count(T < T_STOP ? 0 : 1) == 0) break;   // also fences reuse of the staging buffers
        const uint32_t cnt = min((uint32_t)RT_CHUNK, range.y - base);
        const int buf = (int)(chunk & 1u);
        // prefetch the NEXT chunk's entries (its buffer was last read two iterations ago: the vote above fenced it)
        if (use_tma) {
            if (base + RT_CHUNK < range.y) {
                if (t == 0) issue_entries(tile_entries, base + RT_CHUNK, min((uint32_t)RT_CHUNK, range.y - base - RT_CHUNK), a_ent, a_bar, buf ^ 1);
                ++issued;
            }
            mbar_wait(a_bar + 8u * buf, (chunk >> 1) & 1u);
        }
        if ((uint32_t)t < cnt) {
            const uint32_t r = use_tma ? s_ent[buf][(base & 3u) + t] : __ldg(tile_entries + base + t);
            const float4* rp = reinterpret_cast<const float4*>(recs + r);
            const float4 p0 = __ldg(rp), p1 = __ldg(rp + 1);
            s_q0[t] = p0;
            s_uv[t] = p1;
            s_q2[t] = __ldg(rp + 2);
            if (MODE == 0) {
                // thresholds of the per-warp separating-axis cull below, once per splat: the quad |u| <= 1, |v| <= 1
                // misses a warp rectangle (pixel centres within +-3.5 x +-1.5 of its centre) when |u(centre)| exceeds
                // 1 + |ux| 3.5 + |uy| 1.5 (same for v).  Slack: 1e-5 of the largest magnitude the terms of u can take
                // anywhere in the tile, ~100x the rounding error of the per-pixel u, v.  NaN/inf never cull.
                const float ax = fabsf(p0.x - tcx) + 4.0f, ay = fabsf(p0.y - tcy) + 6.0f;
                const float ur = fabsf(p0.z) * 3.5f + fabsf(p0.w) * 1.5f, vr = fabsf(p1.x) * 3.5f + fabsf(p1.y) * 1.5f;
                const float um = fabsf(p0.z) * ax + fabsf(p0.w) * ay + ur, vm = fabsf(p1.x) * ax + fabsf(p1.y) * ay + vr;
                s_thr[t] = make_float2(ur + 1.0f + 1e-5f * um, vr + 1.0f + 1e-5f * vm);
            }
            if (MODE == 2) {
                float4* s_ex = reinterpret_cast<float4*>(s_mem + SM_EXTRA);
                const float4* ep = extra + (size_t)r * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) s_ex[q * RT_CHUNK + t] = __ldg(ep + q);
            }
            if (AUX) {
                float4* s_ax = reinterpret_cast<float4*>(s_mem + SM_AUX);
                s_ax[t] = __ldg(aux + (size_t)r * 2);
                s_ax[RT_CHUNK + t] = __ldg(aux + (size_t)r * 2 + 1);
            }
        }
        __syncthreads();
        // each warp compacts the chunk to the splats whose bbox touches its 8x4 pixels (order kept)
        uint32_t nl = 0;
        if (__any_sync(0xffffffffu, !(T < T_STOP))) {
            for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool hit = false;
                if (j < cnt) {
                    const float4 q = s_uv[j];
                    const uint32_t bx = __float_as_uint(q.z), by = __float_as_uint(q.w);
                    hit = !((int)(bx >> 16) < wx0 || (int)(bx & 0xFFFFu) > wx0 + 7 || (int)(by >> 16) < wy0 ||
                            (int)(by & 0xFFFFu) > wy0 + 3);
                    if (MODE == 0 && hit) {
                        // separating-axis test of the splat's quad against this warp's pixel centres
                        // [wx0 + .5, wx0 + 7.5] x [wy0 + .5, wy0 + 3.5] (thresholds staged per splat above): the bbox
                        // of a slanted quad passes many warps none of whose pixels it covers
                        const float4 p = s_q0[j];
                        const float2 th = s_thr[j];
                        const float dxc = rcx - p.x, dyc = rcy - p.y;
                        if (fabsf(p.z * dxc + p.w * dyc) > th.x || fabsf(q.x * dxc + q.y * dyc) > th.y) hit = false;
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (hit) s_list[nl + __popc(m & lanemask_lt())] = (unsigned short)(a_base + j * 16u);   // shared address of q0[j]
                nl += __popc(m);
            }
            __syncwarp();
        }
        if (MODE == 0 && !AUX) {
            // two candidates per iteration: one 32-bit load brings both list entries, the four record loads and both
            // coverage tests are independent (ILP), loop control is paid once; blending stays strictly in list order
            // The blend is PREDICATED, not branched: 14 predicated instructions per candidate instead of a divergent block
            // with its BSSY / BRA / BREAK / BSYNC bookkeeping (~23 issue slots; 98 % of the candidates cover some pixel of
            // the warp anyway).  `lim` = 1 while the pixel is alive, -1 once it has stopped (T < T_STOP) or lies outside the
            // frame, so "covered" and "alive" are one comparison and a stopped pixel skips every later splat exactly as an
            // early exit would.
            {
                const uint32_t a_end = a_list + nl * 2u;
                uint32_t a_it = a_list;
                float lim = (T < T_STOP) ? -1.0f : 1.0f;
                auto blend_if_covered = [&](uint32_t a_rec, float u, float v) {
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p, q;\n\t"
                        ".reg .f32 au, av, x, y, z, o, qd, e, a, w, na;\n\t"
                        "abs.f32 au, %5;\n\t"
                        "abs.f32 av, %6;\n\t"
                        "setp.le.f32 p, au, %4;\n\t"
                        "setp.le.and.f32 p, av, %4, p;\n\t"
                        // (the temporaries are computed unconditionally -- a predicated definition would keep their old values
                        // alive across iterations -- only the four accumulations are predicated)
                        "ld.shared.v4.f32 {x, y, z, o}, [%7+8192];\n\t"
                        "mul.rn.f32 qd, %5, %5;\n\t"
                        "fma.rn.f32 qd, %6, %6, qd;\n\t"
                        "mul.rn.f32 qd, qd, 0fC0CFBF83;\n\t"             // -6.492127684f: exp(-4.5 qd) = 2^(qd * -4.5 log2 e)
                        "ex2.approx.ftz.f32 e, qd;\n\t"
                        "mul.rn.f32 a, e, o;\n\t"
                        "min.f32 a, a, 0f3F7FBE77;\n\t"                 // 0.999f
                        "mul.rn.f32 w, a, %0;\n\t"
                        "neg.f32 na, a;\n\t"
                        "@p fma.rn.f32 %1, w, x, %1;\n\t"
                        "@p fma.rn.f32 %2, w, y, %2;\n\t"
                        "@p fma.rn.f32 %3, w, z, %3;\n\t"
                        "@p fma.rn.f32 %0, na, %0, %0;\n\t"
                        "setp.lt.and.f32 q, %0, 0f38D1B717, p;\n\t"     // T < T_STOP (1e-4f) after a blend: the pixel stops
                        "@q mov.f32 %4, 0fBF800000;\n\t"
                        "}"
                        : "+f"(T), "+f"(cr), "+f"(cg), "+f"(cb), "+f"(lim)
                        : "f"(u), "f"(v), "r"(a_rec));
                };
                for (; a_it + 2u < a_end; a_it += 4u) {
                    uint32_t two;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(two) : "r"(a_it));
                    const uint32_t ra = two & 0xFFFFu, rb = two >> 16;
                    float4 pa, pb; float2 sa, sb;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pa.x), "=f"(pa.y), "=f"(pa.z), "=f"(pa.w) : "r"(ra));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pb.x), "=f"(pb.y), "=f"(pb.z), "=f"(pb.w) : "r"(rb));
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+4096];" : "=f"(sa.x), "=f"(sa.y) : "r"(ra));
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+4096];" : "=f"(sb.x), "=f"(sb.y) : "r"(rb));
                    const float dxa = __fsub_rn(fx, pa.x), dya = __fsub_rn(fy, pa.y);
                    const float dxb = __fsub_rn(fx, pb.x), dyb = __fsub_rn(fy, pb.y);
                    const float ua = __fmaf_rn(pa.w, dya, __fmul_rn(pa.z, dxa)), va = __fmaf_rn(sa.y, dya, __fmul_rn(sa.x, dxa));
                    const float ub = __fmaf_rn(pb.w, dyb, __fmul_rn(pb.z, dxb)), vb = __fmaf_rn(sb.y, dyb, __fmul_rn(sb.x, dxb));
                    blend_if_covered(ra, ua, va);
                    blend_if_covered(rb, ub, vb);
                }
                if (a_it != a_end) {   // odd tail
                    uint32_t ra;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ra) : "r"(a_it));
                    float4 pa; float2 sa;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pa.x), "=f"(pa.y), "=f"(pa.z), "=f"(pa.w) : "r"(ra));
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+4096];" : "=f"(sa.x), "=f"(sa.y) : "r"(ra));
                    const float dxa = __fsub_rn(fx, pa.x), dya = __fsub_rn(fy, pa.y);
                    const float ua = __fmaf_rn(pa.w, dya, __fmul_rn(pa.z, dxa)), va = __fmaf_rn(sa.y, dya, __fmul_rn(sa.x, dxa));
                    blend_if_covered(ra, ua, va);
                }
            }
        } else if (!(T < T_STOP)) {
            const uint32_t a_end = a_list + nl * 2u;
            for (uint32_t a_it = a_list; a_it != a_end; a_it += 2u) {
                uint32_t a_rec;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(a_rec) : "r"(a_it));
                float4 q0; float2 q1;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(a_rec));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+4096];" : "=f"(q1.x), "=f"(q1.y) : "r"(a_rec));
                const float dx = __fsub_rn(fx, q0.x), dy = __fsub_rn(fy, q0.y);
                float e, opac;
                float4 q2;
                if (MODE == 0) {
                    const float u = __fmaf_rn(q0.w, dy, __fmul_rn(q0.z, dx));
                    const float v = __fmaf_rn(q1.y, dy, __fmul_rn(q1.x, dx));
                    if (!(fabsf(u) <= 1.0f && fabsf(v) <= 1.0f)) continue;
                    const float qd = __fmaf_rn(v, v, __fmul_rn(u, u));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+8192];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(a_rec));
                    // exp(-4.5 qd) = 2^(qd * -4.5 log2 e); qd <= 2 so the argument stays >= -13 (no range fix-up)
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(qd * -6.492127684f));
                } else {
                    // quad-space offset in half-pixels (x right, y up); the quad is the square |m| <= Rq
                    const float mx = __fadd_rn(dx, dx), my = -__fadd_rn(dy, dy);
                    const float Rq = q1.y;   // MODE 1: quad half-side in half-pixels
                    float power;
                    if (MODE == 1) {
                        if (!(fabsf(mx) <= Rq && fabsf(my) <= Rq)) continue;
                        // q0.z, q0.w, q1.x = conic x, y, z;  d = -m  (gaussian.wgsl:459-462)
                        const float ddx = -mx, ddy = -my;
                        const float t1 = __fmul_rn(__fmul_rn(q0.z, ddx), ddx), t2 = __fmul_rn(__fmul_rn(q1.x, ddy), ddy);
                        power = __fadd_rn(__fmul_rn(-0.5f, __fadd_rn(t1, t2)), __fmul_rn(__fmul_rn(q0.w, ddx), ddy));
                    } else {
                        float4 e0, e1, e2, e3;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e0.x), "=f"(e0.y), "=f"(e0.z), "=f"(e0.w) : "r"(a_rec + SM_EXTRA));
                        if (!(fabsf(mx) <= e0.x && fabsf(my) <= e0.x)) continue;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e1.x), "=f"(e1.y), "=f"(e1.z), "=f"(e1.w) : "r"(a_rec + SM_EXTRA + RT_CHUNK * 16));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e2.x), "=f"(e2.y), "=f"(e2.z), "=f"(e2.w) : "r"(a_rec + SM_EXTRA + 2 * RT_CHUNK * 16));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e3.x), "=f"(e3.y), "=f"(e3.z), "=f"(e3.w) : "r"(a_rec + SM_EXTRA + 3 * RT_CHUNK * 16));
                        // pixel_coord = uv * radius * (1, W/H) + mean   (gaussian.wgsl:441-447; uv * radius == m)
                        const float pcx = __fadd_rn(mx, e0.y), pcy = __fadd_rn(__fmul_rn(my, e0.w), e0.z);
                        // gaussian_2d.wgsl:134-156: hu = px*T2 - T0, hv = py*T2 - T1, p = hu x hv
                        const float hux = __fsub_rn(__fmul_rn(pcx, e3.x), e1.x), huy = __fsub_rn(__fmul_rn(pcx, e3.y), e1.y),
                                    huz = __fsub_rn(__fmul_rn(pcx, e3.z), e1.z);
                        const float hvx = __fsub_rn(__fmul_rn(pcy, e3.x), e2.x), hvy = __fsub_rn(__fmul_rn(pcy, e3.y), e2.y),
                                    hvz = __fsub_rn(__fmul_rn(pcy, e3.z), e2.z);
                        const float cpx = __fsub_rn(__fmul_rn(huy, hvz), __fmul_rn(huz, hvy));
                        const float cpy = __fsub_rn(__fmul_rn(huz, hvx), __fmul_rn(hux, hvz));
                        const float cpz = __fsub_rn(__fmul_rn(hux, hvy), __fmul_rn(huy, hvx));
                        const float us = __fdiv_rn(cpx, cpz), vs = __fdiv_rn(cpy, cpz);
                        const float s3 = __fadd_rn(__fmul_rn(us, us), __fmul_rn(vs, vs));
                        const float ex = __fsub_rn(e0.y, pcx), ey = __fsub_rn(e0.z, pcy);
                        const float s2 = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
                        power = -__fmul_rn(0.5f, fminf(s3, s2));
                    }
                    if (power > 0.0f) continue;                      // gaussian.wgsl:468-470
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+8192];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(a_rec));
                    e = __expf(power);
                }
                opac = q2.w;
                const float a = fminf(e * opac, 0.999f);
                const float w = a * T;
                cr = fmaf(w, q2.x, cr); cg = fmaf(w, q2.y, cg); cb = fmaf(w, q2.z, cb);
                if (AUX) {
                    float4 ad, an;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ad.x), "=f"(ad.y), "=f"(ad.z), "=f"(ad.w) : "r"(a_rec + SM_AUX));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(an.x), "=f"(an.y), "=f"(an.z), "=f"(an.w) : "r"(a_rec + SM_AUX + RT_CHUNK * 16));
                    dr = fmaf(w, ad.x, dr); dg = fmaf(w, ad.y, dg); db = fmaf(w, ad.z, db);
                    nr = fmaf(w, an.x, nr); ng = fmaf(w, an.y, ng); nb = fmaf(w, an.z, nb);
                }
                T = fmaf(-a, T, T);
                if (T < T_STOP) break;
            }
        }
    }
    // an early exit (all pixels saturated) may leave one prefetch in flight: keep the CTA alive until it lands
    if (t == 0 && issued > chunk)             // chunks [0, chunk) were waited for inside the loop
        mbar_wait(a_bar + 8u * (chunk & 1u), (chunk >> 1) & 1u);
    if (!inside) return;
    write_pixel(out, format, (size_t)py * W + px, cr, cg, cb, T);
    if (AUX) {
        write_pixel(out_depth, format, (size_t)py * W + px, dr, dg, db, T);
        write_pixel(out_normal, format, (size_t)py * W + px, nr, ng, nb, T);
    }
}

// ---- MODE 0 fast path: 2 horizontally adjacent pixels per thread -----------------------------------------
// CTA = tile, 4 warps, warp w = rows 4w..4w+3 of the tile (16x4 pixels), lane = (column pair, row).  Per
// candidate splat the record loads, the loop and dy are shared by the two pixels, and a 16-wide warp rectangle
// halves the number of (warp, splat) candidates.  Same per-pixel formulas (bit-identical coverage).
constexpr int R2_THREADS = 128;
constexpr uint32_t R2_LIST = SM_Q2 + RT_CHUNK * 16;                      // u16 [4][256]
constexpr uint32_t R2_BYTES = R2_LIST + (R2_THREADS / 32) * RT_CHUNK * 2;

__device__ __forceinline__ void store_pixel2(void* out, uint32_t format, size_t pix, bool in0, bool in1, float r0, float g0,
                                             float b0, float r1, float g1, float b1, float T0, float T1) {
    if (format >> 8) {            // premultiplied / blend-over output: the generic per-pixel path
        if (in0) write_pixel(out, format, pix, r0, g0, b0, T0);
        if (in1) write_pixel(out, format, pix + 1, r1, g1, b1, T1);
        return;
    }
    if (format == BGS_FORMAT_RGBA32F) {
        float4* o = reinterpret_cast<float4*>(out) + pix;
        if (in0) o[0] = make_float4(r0, g0, b0, 1.0f);
        if (in1) o[1] = make_float4(r1, g1, b1, 1.0f);
    } else if (format == BGS_FORMAT_RGBA16F) {
        uint2* o = reinterpret_cast<uint2*>(out) + pix;
        const __half2 l0 = __floats2half2_rn(r0, g0), h0 = __floats2half2_rn(b0, 1.0f);
        const __half2 l1 = __floats2half2_rn(r1, g1), h1 = __floats2half2_rn(b1, 1.0f);
        if (in0) o[0] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&h0));
        if (in1) o[1] = make_uint2(*reinterpret_cast<const uint32_t*>(&l1), *reinterpret_cast<const uint32_t*>(&h1));
    } else {
        uint32_t* o = reinterpret_cast<uint32_t*>(out) + pix;
        const uint32_t p0 = (uint32_t)(linear_to_srgb(r0) * 255.0f + 0.5f) | ((uint32_t)(linear_to_srgb(g0) * 255.0f + 0.5f) << 8) |
                            ((uint32_t)(linear_to_srgb(b0) * 255.0f + 0.5f) << 16) | 0xFF000000u;
        const uint32_t p1 = (uint32_t)(linear_to_srgb(r1) * 255.0f + 0.5f) | ((uint32_t)(linear_to_srgb(g1) * 255.0f + 0.5f) << 8) |
                            ((uint32_t)(linear_to_srgb(b1) * 255.0f + 0.5f) << 16) | 0xFF000000u;
        if (in0 && in1 && (pix & 1) == 0) *reinterpret_cast<uint2*>(o) = make_uint2(p0, p1);
        else { if (in0) o[0] = p0; if (in1) o[1] = p1; }
    }
}

template <bool CHUNKED>
__global__ void __launch_bounds__(R2_THREADS)
raster2_kernel(const SplatRec* __restrict__ recs, const uint32_t* __restrict__ tile_entries, const uint2* __restrict__ ranges,
               int W, int H, int tiles_x, void* __restrict__ out, uint32_t format, float4* __restrict__ state,
               unsigned char* __restrict__ tile_done, uint32_t* __restrict__ tiles_done, int first, int last) {
    __shared__ __align__(16) unsigned char s_mem[R2_BYTES];
    float4* s_q0 = reinterpret_cast<float4*>(s_mem + SM_Q0);
    float4* s_uv = reinterpret_cast<float4*>(s_mem + SM_UV);
    float4* s_q2 = reinterpret_cast<float4*>(s_mem + SM_Q2);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_mem + R2_LIST) + warp * RT_CHUNK;
    const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(s_mem);
    const uint32_t a_list = a_base + R2_LIST + (uint32_t)warp * RT_CHUNK * 2u;
    int tile_x, tile_y;
    centre_out_tile((int)blockIdx.x, tiles_x, (int)gridDim.x / tiles_x, tile_x, tile_y);
    const int tile = tile_y * tiles_x + tile_x;
    const int wy0 = tile_y * TILE_PX + warp * 4;                    // this warp's 4 rows
    const int px0 = tile_x * TILE_PX + 2 * (lane & 7), py = wy0 + (lane >> 3);
    const bool in0 = px0 < W && py < H, in1 = px0 + 1 < W && py < H;
    const float fx0 = (float)px0 + 0.5f, fx1 = (float)px0 + 1.5f, fy = (float)py + 0.5f;
    uint2 range = ranges[tile];
    range.x = ~range.x;                  // stored as (~start, end), (0, 0) = empty (radix.cu)

    // T < T_STOP <=> pixel done; lim = 1 while alive, -1 once done (folds the "alive" test into |u| <= lim)
    float T0 = in0 ? 1.0f : 0.0f, T1 = in1 ? 1.0f : 0.0f;
    float lim0 = in0 ? 1.0f : -1.0f, lim1 = in1 ? 1.0f : -1.0f;
    float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
    float4* st = nullptr;
    if (CHUNKED) {
        // one of several front-to-back rounds of a frame: the blend state (premultiplied rgb, transmittance) of
        // every pixel lives in `state` (tile-major, so a warp's accesses are contiguous) between rounds; since a
        // round resumes each pixel exactly where the previous one stopped, the frame is bit-identical to one round
        st = state + ((size_t)tile * R2_THREADS + t) * 2;
        if (!first) {
            const bool done = tile_done[tile] != 0;             // every pixel saturated in an earlier round
            if (!last && (done || range.x >= range.y)) return;  // nothing to blend, state unchanged
            const float4 s0 = st[0], s1 = st[1];
            r0 = s0.x; g0 = s0.y; b0 = s0.z; T0 = s0.w;
            r1 = s1.x; g1 = s1.y; b1 = s1.z; T1 = s1.w;
            lim0 = (in0 && !(T0 < T_STOP)) ? 1.0f : -1.0f;
            lim1 = (in1 && !(T1 < T_STOP)) ? 1.0f : -1.0f;
            if (done) range.y = range.x;
        }
    }
    // multi-chunk tiles (the norm on this path: heavy footprints put thousands of entries in a tile) stream their slice of
    // the sorted pair list through the TMA double buffer, like raster_kernel: chunk k + 1's bulk copy (UBLKCP, mbarrier
    // complete_tx) is in flight while chunk k is blended
    __shared__ __align__(16) uint32_t s_ent[2][ENT_WORDS];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const uint32_t a_ent = (uint32_t)__cvta_generic_to_shared(&s_ent[0][0]);
    const uint32_t a_bar = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
    const bool use_tma = range.y > range.x && range.y - range.x > (uint32_t)RT_CHUNK;
    uint32_t issued = 0u, chunk = 0u;
    if (use_tma) {
        if (t == 0) {
            mbar_init(a_bar, 1u); mbar_init(a_bar + 8u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (t == 0) issue_entries(tile_entries, range.x, (uint32_t)RT_CHUNK, a_ent, a_bar, 0);
        issued = 1u;
    }
    for (uint32_t base = range.x; base < range.y; base += RT_CHUNK, ++chunk) {
        if (__syncthreads_count((lim0 > 0.f || lim1 > 0.f) ? 1 : 0) == 0) break;   // also fences smem reuse
        const uint32_t cnt = min((uint32_t)RT_CHUNK, range.y - base);
        const int buf = (int)(chunk & 1u);
        if (use_tma) {
            // prefetch the NEXT chunk's entries (its buffer was last read two iterations ago: the vote above fenced it)
            if (base + RT_CHUNK < range.y) {
                if (t == 0) issue_entries(tile_entries, base + RT_CHUNK, min((uint32_t)RT_CHUNK, range.y - base - RT_CHUNK), a_ent, a_bar, buf ^ 1);
                ++issued;
            }
            mbar_wait(a_bar + 8u * buf, (chunk >> 1) & 1u);
        }
#pragma unroll
        for (int k = 0; k < RT_CHUNK / R2_THREADS; ++k) {
            const uint32_t j = t + k * R2_THREADS;
            if (j < cnt) {
                const uint32_t r = use_tma ? s_ent[buf][(base & 3u) + j] : __ldg(tile_entries + base + j);
                const float4* rp = reinterpret_cast<const float4*>(recs + r);
                s_q0[j] = __ldg(rp);
                s_uv[j] = __ldg(rp + 1);
                s_q2[j] = __ldg(rp + 2);
            }
        }
        __syncthreads();
        // per-warp candidate list: splats whose bbox reaches this warp's rows (the x extent already meets the tile)
        uint32_t nl = 0;
        if (__any_sync(0xffffffffu, lim0 > 0.f || lim1 > 0.f)) {
            for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool hit = false;
                if (j < cnt) {
                    const uint32_t by = __float_as_uint(s_uv[j].w);
                    hit = !((int)(by >> 16) < wy0 || (int)(by & 0xFFFFu) > wy0 + 3);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (hit) s_list[nl + __popc(m & lanemask_lt())] = (unsigned short)(j * 16u);
                nl += __popc(m);
            }
            __syncwarp();
        }
        if (lim0 > 0.f || lim1 > 0.f) {
            const uint32_t a_end = a_list + nl * 2u;
            for (uint32_t a_it = a_list; a_it != a_end; a_it += 2u) {
                uint32_t off;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(off) : "r"(a_it));
                const uint32_t a_rec = a_base + off;
                float4 q0; float2 q1;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(a_rec));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+4096];" : "=f"(q1.x), "=f"(q1.y) : "r"(a_rec));
                const float dy = __fsub_rn(fy, q0.y);
                const float dxa = __fsub_rn(fx0, q0.x), dxb = __fsub_rn(fx1, q0.x);
                const float ua = __fmaf_rn(q0.w, dy, __fmul_rn(q0.z, dxa)), va = __fmaf_rn(q1.y, dy, __fmul_rn(q1.x, dxa));
                const float ub = __fmaf_rn(q0.w, dy, __fmul_rn(q0.z, dxb)), vb = __fmaf_rn(q1.y, dy, __fmul_rn(q1.x, dxb));
                const bool ca = fabsf(ua) <= lim0 && fabsf(va) <= lim0;
                const bool cb = fabsf(ub) <= lim1 && fabsf(vb) <= lim1;
                if (!(ca || cb)) continue;
                float4 q2;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+8192];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(a_rec));
                if (ca) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmaf_rn(va, va, __fmul_rn(ua, ua)) * -6.492127684f));
                    const float a = fminf(e * q2.w, 0.999f);
                    const float w = a * T0;
                    r0 = fmaf(w, q2.x, r0); g0 = fmaf(w, q2.y, g0); b0 = fmaf(w, q2.z, b0);
                    T0 = fmaf(-a, T0, T0);
                    if (T0 < T_STOP) lim0 = -1.0f;
                }
                if (cb) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmaf_rn(vb, vb, __fmul_rn(ub, ub)) * -6.492127684f));
                    const float a = fminf(e * q2.w, 0.999f);
                    const float w = a * T1;
                    r1 = fmaf(w, q2.x, r1); g1 = fmaf(w, q2.y, g1); b1 = fmaf(w, q2.z, b1);
                    T1 = fmaf(-a, T1, T1);
                    if (T1 < T_STOP) lim1 = -1.0f;
                }
                if (!(lim0 > 0.f || lim1 > 0.f)) break;
            }
        }
    }
    // an early exit (all pixels saturated) may leave one prefetch in flight: keep the CTA alive until it lands
    if (t == 0 && issued > chunk) mbar_wait(a_bar + 8u * (chunk & 1u), (chunk >> 1) & 1u);
    if (CHUNKED && !last) {
        st[0] = make_float4(r0, g0, b0, T0);
        st[1] = make_float4(r1, g1, b1, T1);
        if (__syncthreads_count((lim0 > 0.f || lim1 > 0.f) ? 1 : 0) == 0 && t == 0) {
            tile_done[tile] = 1;
            atomicAdd(tiles_done, 1u);
        }
        return;
    }
    if (!(in0 || in1)) return;
    store_pixel2(out, format, (size_t)py * W + px0, in0, in1, r0, g0, b0, r1, g1, b1, T0, T1);
}

void launch_raster(int mode, bool large_footprints, const SplatRec* recs, const float4* extra, const uint32_t* tile_entries,
                   const uint2* ranges, int W, int H, int tiles_x, int tiles_y, void* out, uint32_t format,
                   const float4* aux, void* out_depth, void* out_normal, cudaStream_t stream) {
    const int grid = tiles_x * tiles_y;
    if (aux != nullptr) {          // colour + depth + normal in one pass (bgs_render_aux)
        if (mode == 0) raster_kernel<0, true><<<grid, RT_THREADS, 0, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, aux, out_depth, out_normal);
        else if (mode == 1) raster_kernel<1, true><<<grid, RT_THREADS, 0, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, aux, out_depth, out_normal);
        else raster_kernel<2, true><<<grid, RT_THREADS, 0, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, aux, out_depth, out_normal);
        return;
    }
    // measured on B200: the 2-pixels-per-thread variant wins when splats cover many tiles (C2 raw, scale 1:
    // 131 -> 117 us) and loses when most splats are a few pixels (C3, scale 0.02: 178 -> 217 us)
    if (mode == 0 && large_footprints)
        raster2_kernel<false><<<grid, R2_THREADS, 0, stream>>>(recs, tile_entries, ranges, W, H, tiles_x, out, format,
                                                               nullptr, nullptr, nullptr, 1, 1);
    else if (mode == 0) {
        // experiment knob: pad the CTA's shared memory so fewer raster CTAs fit per SM and kernels of another
        // in-flight frame can co-run (BGS_RASTER_PAD = bytes of dynamic shared memory, default 0)
        static int pad = -1;
        if (pad < 0) { const char* e = getenv("BGS_RASTER_PAD"); pad = e ? atoi(e) : 0; }
        raster_kernel<0, false><<<grid, RT_THREADS, pad, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, nullptr, nullptr, nullptr);
    }
    else if (mode == 1)
        raster_kernel<1, false><<<grid, RT_THREADS, 0, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, nullptr, nullptr, nullptr);
    else
        raster_kernel<2, false><<<grid, RT_THREADS, 0, stream>>>(recs, extra, tile_entries, ranges, W, H, tiles_x, out, format, nullptr, nullptr, nullptr);
}

// One front-to-back round of a chunked frame (quad-uv records only); see raster2_kernel.
void launch_raster_round(const SplatRec* recs, const uint32_t* tile_entries, const uint2* ranges, int W, int H, int tiles_x,
                         int tiles_y, void* out, uint32_t format, float4* state, unsigned char* tile_done,
                         uint32_t* tiles_done, int first, int last, cudaStream_t stream) {
    raster2_kernel<true><<<tiles_x * tiles_y, R2_THREADS, 0, stream>>>(recs, tile_entries, ranges, W, H, tiles_x, out, format,
                                                                       state, tile_done, tiles_done, first, last);
}

}  // namespace bgs
