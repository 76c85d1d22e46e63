// common.cuh -- shared device-side types and helpers of libbgs (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bgs.h"

namespace bgs {

constexpr int TILE_PX = 16;           // 16x16-pixel raster tiles (a7)
constexpr float T_STOP = 1.0e-4f;     // a pixel stops once its transmittance drops below this

// Look-back status word: bits 31:30 flag, bits 29:0 value (n and n_pairs are < 2^30).
constexpr uint32_t LB_EMPTY = 0u, LB_AGG = 1u << 30, LB_INC = 2u << 30, LB_VMASK = (1u << 30) - 1u;

// Per-frame constants handed to the kernels by value (column-major matrices, as Bevy).
struct FrameConsts {
    float model[16];            // CloudUniform.transform
    float view_from_world[16];
    float clip_from_world[16];
    float cam[3];
    float W, H;                 // viewport.zw
    float p00, p11;             // clip_from_view[0].x, clip_from_view[1].y
    float global_opacity, global_scale;
    uint32_t color_space;
    uint32_t key_shift;         // 32 - radix_sort_depth_bits
    uint32_t gaussian_mode, rasterize_mode, aabb, adaptive, draw_mode;
    int Wi, Hi, tiles_x, tiles_y;
    uint32_t n_cloud;           // gaussians in the cloud
    uint32_t model_identity;    // CloudUniform.transform is exactly the identity (key-gen skips the multiply)
    float aabb_min[3], aabb_max[3];   // CloudUniform.min / .max (RasterizeMode::Position)
    uint32_t cov_pre;           // the cloud carries Covariance3dOpacityPacked128 records (f16.rs:131-170) instead of rotation + scale
    uint32_t aux;               // bgs_render_aux: the projection also emits the Depth and Normal colour sources
};

// Projected splat record, 48 B, stored by front-to-back rank.
struct __align__(16) SplatRec {
    float cx, cy, ux, uy;       // centre (px), first row of the pixel-offset -> uv map
    float vx, vy;               // second row
    uint32_t bx, by;            // pixel bbox: lo | hi << 16 (inclusive); lo > hi = empty
    float r, g, b, op;          // linear rgb (unclamped), opacity * global_opacity
};
static_assert(sizeof(SplatRec) == 48, "SplatRec must be 48 bytes");

// Counters of one binning round.  A frame normally has ONE round (chunk 0 = all visible splats); frames whose
// splats cover many tiles each are binned / sorted / rasterised in several front-to-back rank chunks so that the
// rounds after every tile has saturated emit nothing (saturation-aware binning, see api.cu).
struct ChunkCounters {
    uint32_t n_pairs;           // (splat, tile) pairs emitted (clamped to capacity)
    uint32_t n_pairs_needed;    // pairs the round needs (may exceed capacity -> host regrows and redoes the frame)
    uint32_t barrier;           // grid barrier of bin_emit_coop (two uses per launch)
    uint32_t big_count, big_head;   // queue of large footprints (grows from the back of the queue arrays)
    uint32_t med_count, med_head;   // queue of medium footprints (grows from the front), drained 32 per warp
    uint32_t tile_ctr_bin;      // tile tickets of the fallback (chained look-back) bin kernel
    uint32_t tile_ctr_sort[4];  // tile tickets of the pair sort's passes
    uint32_t skipped;           // 1 = the round emitted nothing because every tile had already saturated
    uint32_t pad[3];
};
static_assert(sizeof(ChunkCounters) == 64, "ChunkCounters is 64 bytes");
constexpr int MAX_CHUNKS = 8;

// Device-resident per-frame counters (cleared at frame start).
struct FrameCounters {
    uint32_t n_sort;            // entries the depth sort runs over (n_vis, or N with SORT_ALL)
    uint32_t n_vis;             // in-frustum gaussians
    uint32_t tiles_done;        // tiles whose pixels have all saturated (chunked frames)
    uint32_t pad0;
    uint32_t tile_ctr[8];       // dynamic tile tickets: [0] keygen (fallback kernel), [1..4] depth sort passes
    uint32_t culled_min_inv;    // RasterizeMode::Depth: max over culled of (0xFFFFFFFF - index); 0 = none culled
    uint32_t culled_max_p1;     //                       max over culled of (index + 1);          0 = none culled
    float depth_min, depth_max; //                       distances of sorted[N-1] / sorted[1] (gaussian.wgsl:329-349)
    uint32_t barrier[4];        // grid barrier of keygen_coop: [0]
    ChunkCounters chunk[MAX_CHUNKS];
};

// flag/counter words exchanged between CTAs of one kernel: relaxed, GPU scope (L2 is the coherence point;
// ld/st.volatile would be SYSTEM scope)
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
#ifdef BGS_SYS_SCOPE
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
#ifdef BGS_SYS_SCOPE
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t lanemask_le() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// Optional per-CTA timeline (debug builds of a run: BGS_TIMELINE=1): %globaltimer stamps at phase boundaries.
__device__ __forceinline__ void timeline_stamp(unsigned long long* tl, int slot) {
    if (tl != nullptr && threadIdx.x == 0) {
        unsigned long long t;
#ifdef BGS_TIMELINE_CLOCK64
        t = (unsigned long long)clock64();              // SM cycles: exact intra-CTA deltas
#else
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#endif
        tl[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

// Grid-wide barrier for kernels launched with cudaLaunchCooperativeKernel (all CTAs co-resident).
// `bar` is zeroed by the per-frame clear; each use passes the cumulative arrival target.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_volatile(bar) < target) { __nanosleep(32); }
        __threadfence();
    }
    __syncthreads();
}

// Sum of counts[0 .. upto) by the whole CTA (counts were published before a grid barrier).
template <int THREADS>
__device__ __forceinline__ uint32_t block_sum_prefix(const uint32_t* counts, uint32_t upto, uint32_t* s_red /*[THREADS/32]*/) {
    uint32_t v = 0u;
    for (uint32_t p = threadIdx.x; p < upto; p += THREADS) v += ld_volatile(counts + p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t tot = 0u;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) tot += s_red[w];
    __syncthreads();
    return tot;
}

// Decoupled look-back over single-word tile status, executed by ONE full warp.
// Publishes this tile's aggregate, sums predecessors' aggregates back to the nearest inclusive
// prefix, publishes the inclusive prefix, returns the exclusive prefix (same value on all lanes).
__device__ __forceinline__ uint32_t warp_lookback(uint32_t* status, int tile, uint32_t aggregate) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_volatile(status, LB_INC | aggregate);
        return 0u;
    }
    if (lane == 0) st_volatile(status + tile, LB_AGG | aggregate);
    uint32_t excl = 0u;
    for (int base = tile - 1; base >= 0; base -= 32) {
        const int idx = base - lane;
        uint32_t w;
        do {
            w = (idx >= 0) ? ld_volatile(status + idx) : (LB_INC | 0u);
        } while (__any_sync(0xffffffffu, (w >> 30) == 0u));
        const uint32_t inc_mask = __ballot_sync(0xffffffffu, (w >> 30) == 2u);
        uint32_t v = w & LB_VMASK;
        if (inc_mask) {
            const int first = __ffs(inc_mask) - 1;   // nearest predecessor holding an inclusive prefix
            if (lane > first) v = 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (inc_mask) break;
    }
    if (lane == 0) st_volatile(status + tile, LB_INC | ((excl + aggregate) & LB_VMASK));
    return excl;
}

}  // namespace bgs
