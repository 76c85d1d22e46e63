// keygen.cu -- stage 1: depth-key generation + stable stream compaction of the visible set.
//
// Replaces radix_sort_a's key half (src/sort/radix.wgsl:86-106): key = 0xFFFFFFFF - bits(|Mp-cam|^2)
// for in-frustum gaussians, 0xFFFFFFFF otherwise, shifted by 32 - depth_bits.  Instead of sorting
// the culled (all-ones) keys with everything else, the visible (key, index) pairs are compacted
// IN INDEX ORDER (single-pass chained scan with decoupled look-back), so the stable LSD sort that
// follows sees the same tie order as the reference; the culled tail of the reference's
// sorted_entry_buffer is "ascending index" and is reconstructed only by the debug hook.
//
// HBM-bound streaming kernel: 16 B read per gaussian (coalesced float4), 8 B written per
// visible gaussian.  Compiled with -fmad=false (see project_math.cuh).
#include "project_math.cuh"

namespace bgs {

constexpr int KG_THREADS = 256;
constexpr int KG_ITEMS = 8;
constexpr int KG_TILE = KG_THREADS * KG_ITEMS;

__global__ void __launch_bounds__(KG_THREADS)
keygen_compact_kernel(const float4* __restrict__ pos, uint32_t n, FrameConsts fc, int sort_all,
                      uint32_t* __restrict__ keys_out, uint32_t* __restrict__ ids_out, uint32_t* __restrict__ slots_out,
                      uint32_t* __restrict__ status, FrameCounters* __restrict__ ctr) {
    __shared__ uint32_t s_cnt[KG_ITEMS * (KG_THREADS / 32)];
    __shared__ uint32_t s_off[KG_ITEMS * (KG_THREADS / 32)];
    __shared__ uint32_t s_base;
    __shared__ int s_tile;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // dynamic tile ticket: look-back only ever waits on tiles that already started
    if (t == 0) s_tile = (int)atomicAdd(&ctr->tile_ctr[0], 1u);
    __syncthreads();
    const int tile = s_tile;
    const uint32_t tile_base = (uint32_t)tile * KG_TILE;

    float4 p[KG_ITEMS];
#pragma unroll
    for (int j = 0; j < KG_ITEMS; ++j) {
        const uint32_t i = tile_base + j * KG_THREADS + t;
        p[j] = (i < n) ? __ldcs(pos + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t key[KG_ITEMS];
    uint32_t prefix[KG_ITEMS];   // rank among the visible items of (item j, this warp)
    uint32_t vis_bits = 0;
#pragma unroll
    for (int j = 0; j < KG_ITEMS; ++j) {
        const uint32_t i = tile_base + j * KG_THREADS + t;
        bool kvis;
        const uint32_t kkey = key_of_fast(fc, p[j].x, p[j].y, p[j].z, kvis);
        const bool v = (i < n) && kvis;
        key[j] = kkey;
        if (fc.rasterize_mode == BGS_RASTERIZE_DEPTH && i < n && !kvis) {
            atomicMax(&ctr->culled_min_inv, 0xFFFFFFFFu - i);
            atomicMax(&ctr->culled_max_p1, i + 1u);
        }
        const uint32_t b = __ballot_sync(0xffffffffu, v);
        prefix[j] = __popc(b & lanemask_lt());
        if (v) vis_bits |= 1u << j;
        if (lane == 0) s_cnt[j * (KG_THREADS / 32) + warp] = __popc(b);
    }
    __syncthreads();
    if (sort_all) {
        // reference-literal mode: every entry goes to the sort; just count the visible ones
        if (warp == 0) {
            uint32_t v = s_cnt[lane] + s_cnt[lane + 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(&ctr->n_vis, v);
            if (lane == 0 && tile == 0) ctr->n_sort = n;
        }
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            const uint32_t i = tile_base + j * KG_THREADS + t;
            if (i < n) { keys_out[i] = key[j]; ids_out[i] = i; }
        }
        return;
    }
    if (warp == 0) {
        // 64 (item, warp) counts in index order: exclusive scan, two per lane
        const uint32_t a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
        uint32_t incl = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t excl = incl - (a + b);
        s_off[2 * lane] = excl;
        s_off[2 * lane + 1] = excl + a;
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t base = warp_lookback(status, tile, total);
        if (lane == 0) {
            s_base = base;
            // the last tile (in index order) knows the final count
            if (tile_base + KG_TILE >= n) { ctr->n_vis = base + total; ctr->n_sort = base + total; }
        }
    }
    __syncthreads();
    const uint32_t base = s_base;
#pragma unroll
    for (int j = 0; j < KG_ITEMS; ++j) {
        if (vis_bits & (1u << j)) {
            const uint32_t dst = base + s_off[j * (KG_THREADS / 32) + warp] + prefix[j];
            keys_out[dst] = key[j];
            ids_out[dst] = tile_base + j * KG_THREADS + t;
            slots_out[dst] = dst;
        }
    }
}

// Cooperative variant (all CTAs co-resident, launched with cudaLaunchCooperativeKernel): each CTA owns
// a CONTIGUOUS range of tiles.  Phase 1 streams the positions once (16 B/gaussian), writes the
// uncompacted keys (4 B/gaussian, L2-sized scratch) and publishes the CTA's visible count.  One grid
// barrier.  Phase 2 sums the counts of all earlier CTAs in parallel (no chained look-back), re-reads
// its own keys from L2 and writes the visible (key, index) pairs compacted in index order.
// SMEM_KEYS: the CTA's keys stay in (dynamic) shared memory across the barrier instead of taking a round trip
// through the global scratch (used whenever the CTA's range fits: <= 48 KB of keys).
template <bool SMEM_KEYS>
__global__ void __launch_bounds__(KG_THREADS)
keygen_coop_kernel(const float4* __restrict__ pos, uint32_t n, FrameConsts fc, uint32_t* __restrict__ keys_tmp,
                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ ids_out, uint32_t* __restrict__ slots_out,
                   uint32_t* __restrict__ block_cnt, FrameCounters* __restrict__ ctr, uint32_t* __restrict__ hist,
                   int hist_passes) {
    __shared__ uint32_t s_cnt[KG_ITEMS * (KG_THREADS / 32)];
    __shared__ uint32_t s_hist[4 * 256];   // digit histograms of the visible keys (the depth sort's pre-pass, fused)
    __shared__ uint32_t s_off[KG_ITEMS * (KG_THREADS / 32)];
    __shared__ uint32_t s_red[KG_THREADS / 32];
    __shared__ uint32_t s_total;
    extern __shared__ uint32_t s_keys_dyn[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t G = gridDim.x, b = blockIdx.x;
    const uint32_t tiles_total = (n + KG_TILE - 1) / KG_TILE;
    const uint32_t t0 = (uint32_t)((uint64_t)b * tiles_total / G), t1 = (uint32_t)((uint64_t)(b + 1) * tiles_total / G);
    const uint32_t culled = 0xFFFFFFFFu >> fc.key_shift;

    // ---- phase 1: keys for every gaussian of this CTA's range, visible count
    uint32_t mine = 0u;
    uint32_t cmin_inv = 0u, cmax_p1 = 0u;   // Depth mode only: extremes of the culled indices
    for (uint32_t tile = t0; tile < t1; ++tile) {
        const uint32_t tile_base = tile * KG_TILE;
        float4 p[KG_ITEMS];
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            const uint32_t i = tile_base + j * KG_THREADS + t;
            p[j] = (i < n) ? __ldcs(pos + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            const uint32_t i = tile_base + j * KG_THREADS + t;
            bool kvis;
            const uint32_t kkey = key_of_fast(fc, p[j].x, p[j].y, p[j].z, kvis);
            if (i < n) {
                const uint32_t key = kvis ? kkey : culled;
                if (SMEM_KEYS) s_keys_dyn[(tile - t0) * KG_TILE + j * KG_THREADS + t] = key;
                else __stcg(keys_tmp + i, key);
                mine += kvis ? 1u : 0u;
                if (!kvis) { cmin_inv = max(cmin_inv, 0xFFFFFFFFu - i); cmax_p1 = max(cmax_p1, i + 1u); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_red[warp] = mine;
    if (fc.rasterize_mode == BGS_RASTERIZE_DEPTH) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cmin_inv = max(cmin_inv, __shfl_xor_sync(0xffffffffu, cmin_inv, o));
            cmax_p1 = max(cmax_p1, __shfl_xor_sync(0xffffffffu, cmax_p1, o));
        }
        if (lane == 0 && cmax_p1) { atomicMax(&ctr->culled_min_inv, cmin_inv); atomicMax(&ctr->culled_max_p1, cmax_p1); }
    }
    __syncthreads();
    if (t == 0) {
        uint32_t tot = 0u;
#pragma unroll
        for (int w = 0; w < KG_THREADS / 32; ++w) tot += s_red[w];
        s_total = tot;
        st_volatile(block_cnt + b, tot);
    }
    grid_barrier(&ctr->barrier[0], G);

    // ---- phase 2: exclusive prefix over CTAs, then ordered compaction of this CTA's range
    for (int i = t; i < hist_passes * 256; i += KG_THREADS) s_hist[i] = 0u;
    uint32_t run = block_sum_prefix<KG_THREADS>(block_cnt, b, s_red);
    if (b == G - 1 && t == 0) { ctr->n_vis = run + s_total; ctr->n_sort = run + s_total; }
    for (uint32_t tile = t0; tile < t1; ++tile) {
        const uint32_t tile_base = tile * KG_TILE;
        uint32_t key[KG_ITEMS], prefix[KG_ITEMS];
        uint32_t vis_bits = 0u;
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            const uint32_t i = tile_base + j * KG_THREADS + t;
            key[j] = (i < n) ? (SMEM_KEYS ? s_keys_dyn[(tile - t0) * KG_TILE + j * KG_THREADS + t] : __ldcg(keys_tmp + i)) : culled;
        }
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            const bool v = key[j] != culled;
            const uint32_t bal = __ballot_sync(0xffffffffu, v);
            prefix[j] = __popc(bal & lanemask_lt());
            if (v) vis_bits |= 1u << j;
            if (lane == 0) s_cnt[j * (KG_THREADS / 32) + warp] = __popc(bal);
        }
        __syncthreads();
        if (warp == 0) {
            const uint32_t a = s_cnt[2 * lane], c2 = s_cnt[2 * lane + 1];
            uint32_t incl = a + c2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const uint32_t excl = incl - (a + c2);
            s_off[2 * lane] = excl;
            s_off[2 * lane + 1] = excl + a;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KG_ITEMS; ++j) {
            if (vis_bits & (1u << j)) {
                const uint32_t dst = run + s_off[j * (KG_THREADS / 32) + warp] + prefix[j];
                keys_out[dst] = key[j];
                ids_out[dst] = tile_base + j * KG_THREADS + t;   // compact slot -> gaussian index
                slots_out[dst] = dst;                            // the sort's payload: the compact slot
                // (warp-aggregating the clustered upper digits with match.any was measured slower: +12 us on C3)
                for (int p = 0; p < hist_passes; ++p) atomicAdd(&s_hist[p * 256 + ((key[j] >> (8 * p)) & 255u)], 1u);
            }
        }
        run += s_total;
        __syncthreads();
    }
    for (int i = t; i < hist_passes * 256; i += KG_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// Debug hook: rebuild the reference's full sorted_entry_buffer (sort/mod.rs:323-329) from the
// compacted result: [0, n_vis) = sorted visible entries, then every culled index ascending
// with key 0xFFFFFFFF >> shift.  Single block per call chunk; not on the hot path.
__global__ void culled_flags_kernel(const float4* __restrict__ pos, uint32_t n, FrameConsts fc,
                                    uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    flags[i] = key_of(fc, p.x, p.y, p.z).visible ? 0u : 1u;
}

void launch_keygen(const float4* pos, uint32_t n, const FrameConsts& fc, int sort_all, uint32_t* keys_out,
                   uint32_t* ids_out, uint32_t* slots_out, uint32_t* status, FrameCounters* ctr, cudaStream_t stream) {
    const uint32_t tiles = (n + KG_TILE - 1) / KG_TILE;
    keygen_compact_kernel<<<tiles, KG_THREADS, 0, stream>>>(pos, n, fc, sort_all, keys_out, ids_out, slots_out, status, ctr);
}
uint32_t keygen_num_tiles(uint32_t n) { return (n + KG_TILE - 1) / KG_TILE; }

constexpr size_t KG_SMEM_MAX = 5 * KG_TILE * 4;   // 40 KB of keys per CTA (static + dynamic stays under the 48 KB default limit)
int keygen_coop_blocks_per_sm() {
    // sized for the shared-memory variant at its largest footprint, so either variant is co-resident
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, keygen_coop_kernel<true>, KG_THREADS, KG_SMEM_MAX) != cudaSuccess) return 0;
    return b;
}
cudaError_t launch_keygen_coop(const float4* pos, uint32_t n, const FrameConsts& fc, uint32_t* keys_tmp, uint32_t* keys_out,
                               uint32_t* ids_out, uint32_t* slots_out, uint32_t* block_cnt, FrameCounters* ctr,
                               uint32_t* hist, int hist_passes, uint32_t grid, cudaStream_t stream) {
    FrameConsts fcc = fc;
    void* args[] = {(void*)&pos, (void*)&n, (void*)&fcc, (void*)&keys_tmp, (void*)&keys_out, (void*)&ids_out,
                    (void*)&slots_out, (void*)&block_cnt, (void*)&ctr, (void*)&hist, (void*)&hist_passes};
    const uint32_t tiles_total = (n + KG_TILE - 1) / KG_TILE;
    const size_t need = (size_t)((tiles_total + grid - 1) / grid) * KG_TILE * 4;   // keys of the largest CTA range
    if (need <= KG_SMEM_MAX)
        return cudaLaunchCooperativeKernel((const void*)keygen_coop_kernel<true>, dim3(grid), dim3(KG_THREADS), args, need, stream);
    return cudaLaunchCooperativeKernel((const void*)keygen_coop_kernel<false>, dim3(grid), dim3(KG_THREADS), args, 0, stream);
}

void launch_culled_flags(const float4* pos, uint32_t n, const FrameConsts& fc, uint32_t* flags, cudaStream_t stream) {
    culled_flags_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pos, n, fc, flags);
}

}  // namespace bgs
