// radix.cu -- stage 2 (and the tile-id sort of stage 4): stable LSD radix sort, 8-bit digits,
// ONE global sweep per digit place ("onesweep": per-tile warp-level ranking + decoupled
// look-back across tiles; no separate count / scan-over-tiles dispatches).
//
// Replaces radix_sort_b / radix_sort_c_count_tiles / radix_sort_c_scan_tiles /
// radix_sort_c_scatter (src/sort/radix.wgsl:110-279) and the 3P+3 dispatches of run_radix_sort
// (src/sort/radix.rs:672-754).  Same contract: ascending by key, stable (ties keep input order),
// P = depth_bits / 8 passes (src/render/mod.rs:715-745).  Entry count comes from device memory
// (n_ptr) so no host round-trip sits between key-gen and the sort.
//
// HBM-bound: per pass 8 B read + 8 B written per entry; histogram pre-pass reads 4 B per entry.
#include "common.cuh"

namespace bgs {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS_MIN = 8;                  // small sorts: 2048-entry tiles (latency), large: 4096 (throughput)
constexpr int LB_BATCH = 8;                       // look-back loads in flight per thread

// ---- digit histograms for all passes in one read of the keys -----------------------------
constexpr int HS_THREADS = 256;
constexpr int HS_ITEMS = 16;

__global__ void __launch_bounds__(HS_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, int passes,
                  uint32_t* __restrict__ hist /* [passes][256] */) {
    __shared__ uint32_t s_hist[4 * 256];
    const uint32_t n = *n_ptr;
    for (int i = threadIdx.x; i < passes * 256; i += HS_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const uint32_t chunk = HS_THREADS * HS_ITEMS;
    for (uint32_t base = blockIdx.x * chunk; base < n; base += gridDim.x * chunk) {
        uint32_t k[HS_ITEMS];
#pragma unroll
        for (int j = 0; j < HS_ITEMS; ++j) {
            const uint32_t i = base + j * HS_THREADS + threadIdx.x;
            k[j] = (i < n) ? __ldcs(keys + i) : 0u;
        }
        for (int p = 0; p < passes; ++p) {
            // a thread's consecutive keys often share the high digits (depth keys): merge runs
            // before touching shared memory to keep same-bin atomic contention low
            uint32_t run_d = 0xFFFFFFFFu, run_c = 0u;
#pragma unroll
            for (int j = 0; j < HS_ITEMS; ++j) {
                const uint32_t i = base + j * HS_THREADS + threadIdx.x;
                if (i >= n) break;
                const uint32_t d = (k[j] >> (8 * p)) & 255u;
                if (d == run_d) { ++run_c; }
                else {
                    if (run_c) atomicAdd(&s_hist[p * 256 + run_d], run_c);
                    run_d = d; run_c = 1u;
                }
            }
            if (run_c) atomicAdd(&s_hist[p * 256 + run_d], run_c);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += HS_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---- one digit place ------------------------------------------------------------------------
template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, 4)
onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ hist /* raw counts [256] */,
                uint32_t* __restrict__ status /* [tiles][256], zeroed */, uint32_t* __restrict__ tile_ctr,
                int shift, unsigned long long* __restrict__ tl) {
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    __shared__ uint32_t s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];
    __shared__ uint32_t s_whist[RS_WARPS][256];   // per-warp digit counts -> per-warp exclusive offsets
    __shared__ uint32_t s_binstart[256];          // tile-local exclusive digit offsets
    __shared__ uint32_t s_gbase[256];             // global destination of digit d's run, minus s_binstart[d]
    __shared__ uint32_t s_wtot[RS_WARPS];
    __shared__ uint32_t s_tile;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n = *n_ptr;
    const uint32_t num_tiles = (n + RS_TILE - 1) / RS_TILE;

    while (true) {
        if (t == 0) s_tile = atomicAdd(tile_ctr, 1u);
#pragma unroll
        for (int i = 0; i < RS_WARPS; ++i) s_whist[i][t] = 0u;
        if (RS_ITEMS == RS_ITEMS_MIN) {
#pragma unroll
            for (int i = 0; i < RS_WARPS; ++i) s_vals[i * 256 + t] = 0u;   // the peer-mask table
        }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;
        const uint32_t tile_base = tile * RS_TILE;
        unsigned long long* tlt = tl ? tl + (size_t)(tile < 4096u ? tile : 4095u) * 8 - (size_t)blockIdx.x * 8 : nullptr;
        timeline_stamp(tlt, 0);

        // warp-striped load: warp w owns [w*512, (w+1)*512) of the tile; item j = 32 consecutive entries
        uint32_t k[RS_ITEMS];
        const uint32_t my_base = tile_base + warp * (32 * RS_ITEMS) + lane;
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint32_t i = my_base + j * 32;
            k[j] = (i < n) ? __ldcs(keys_in + i) : 0xFFFFFFFFu;   // padding sorts to the tile's tail
        }
        // stable in-warp ranking: entries of one digit are ranked in (item, lane) order
        uint32_t rank[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint32_t d = (k[j] >> shift) & 255u;
#ifdef RS_MATCH_BALLOT
            // 8 ballots + mask intersections instead of one MATCH.ANY (which issues at a fraction of the ALU rate)
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const bool on = (d >> bit) & 1u;
                const uint32_t m = __ballot_sync(0xffffffffu, on);
                peers &= on ? m : ~m;
            }
#else
            uint32_t peers;
            if (RS_ITEMS == RS_ITEMS_MIN) {
                // small (latency-bound) sorts: peers via a per-warp mask table in shared memory (aliases s_vals,
                // unused until the scatter): one ATOMS.OR per lane, conflicts only among lanes sharing the digit.
                // Measured on B200 (C3): -8 us on the depth sort, -12 us on the pair sort vs MATCH.ANY.
                uint32_t* mm = s_vals + warp * 256;
                atomicOr(&mm[d], 1u << lane);
                __syncwarp();
                peers = mm[d];
                __syncwarp();
                if (lane == 31 - __clz(peers)) mm[d] = 0u;
            } else {
                // large (throughput-bound) sorts: MATCH.ANY (the mask table loses there: 232 vs 190 us at 6 M entries)
                peers = __match_any_sync(0xffffffffu, d);
            }
#endif
            const int leader = 31 - __clz(peers);
            uint32_t old = 0u;
            if (lane == leader) {
                old = s_whist[warp][d];
                s_whist[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[j] = old + __popc(peers & lanemask_lt());
            __syncwarp();
        }
        timeline_stamp(tlt, 1);
        __syncthreads();

        // thread t owns digit t: exclusive scan across warps, tile totals
        uint32_t cnt = 0u;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = s_whist[w][t];
            s_whist[w][t] = cnt;
            cnt += c;
        }
        const uint32_t tile_end = tile_base + RS_TILE;
        const uint32_t pad = (tile_end > n) ? (tile_end - n) : 0u;
        const uint32_t cnt_valid = (t == 255) ? cnt - pad : cnt;   // padding is all digit 255

        // tile-local exclusive scan over digits (padding included: it defines smem positions)
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_wtot[warp] = incl;
        // global exclusive scan of the raw histogram (tile 0 seeds the look-back chain with it)
        uint32_t gh_incl = 0u, gh = 0u;
        if (tile == 0) {
            gh = hist[t];
            gh_incl = gh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, gh_incl, o);
                if (lane >= o) gh_incl += y;
            }
        }
        __syncthreads();
        uint32_t wprefix = 0u;
        for (int w = 0; w < warp; ++w) wprefix += s_wtot[w];
        const uint32_t binstart = wprefix + incl - cnt;
        s_binstart[t] = binstart;

        timeline_stamp(tlt, 2);
        // decoupled look-back, one digit per thread
        uint32_t excl;
        uint32_t* my_status = status + (size_t)tile * 256 + t;
        if (tile == 0) {
            __syncthreads();                      // s_wtot reuse below
            if (lane == 31) s_wtot[warp] = gh_incl;
            __syncthreads();
            uint32_t gp = 0u;
            for (int w = 0; w < warp; ++w) gp += s_wtot[w];
            excl = gp + gh_incl - gh;
            st_volatile(my_status, LB_INC | ((excl + cnt_valid) & LB_VMASK));
        } else {
            st_volatile(my_status, LB_AGG | cnt_valid);
            excl = 0u;
#ifdef BGS_EXP_NOLOOKBACK
            if (false)
#endif
            {
            // look back over the predecessors' per-digit words, LB_BATCH independent loads in flight
            const uint32_t* ps = my_status;
            uint32_t back = tile;
            bool found = false;
            while (!found) {
                uint32_t w[LB_BATCH];
#pragma unroll
                for (int q = 0; q < LB_BATCH; ++q)
                    w[q] = ((uint32_t)q < back) ? ld_volatile(ps - 256 * (q + 1)) : (LB_INC | 0u);
#pragma unroll
                for (int q = 0; q < LB_BATCH; ++q) {
                    if (!found) {
                        uint32_t x = w[q];
                        while ((x >> 30) == 0u) x = ld_volatile(ps - 256 * (q + 1));
                        excl += x & LB_VMASK;
                        found = (x >> 30) == 2u;
                    }
                }
                ps -= 256 * LB_BATCH;
                back = back > (uint32_t)LB_BATCH ? back - LB_BATCH : 0u;
            }
            }
            st_volatile(my_status, LB_INC | ((excl + cnt_valid) & LB_VMASK));
        }
        s_gbase[t] = excl - binstart;
        __syncthreads();
        timeline_stamp(tlt, 3);

        // scatter into tile-sorted order in shared memory
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint32_t d = (k[j] >> shift) & 255u;
            const uint32_t pos = s_binstart[d] + s_whist[warp][d] + rank[j];
            const uint32_t i = my_base + j * 32;
            s_keys[pos] = k[j];
            s_vals[pos] = (i < n) ? __ldcs(vals_in + i) : 0u;
        }
        __syncthreads();
        // coalesced write-out: consecutive positions of one digit land on consecutive addresses
        const uint32_t valid = RS_TILE - pad;
#pragma unroll 4
        for (uint32_t p = t; p < valid; p += RS_THREADS) {
            const uint32_t kk = s_keys[p];
            const uint32_t dst = s_gbase[(kk >> shift) & 255u] + p;
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[p];
        }
        timeline_stamp(tlt, 4);
        __syncthreads();
    }
}

// Clears the look-back status rows a sort over *n_ptr entries has used (passes rows of `stride` words), so the same
// rows can serve the next sort of the frame (chunked frames run one pair sort per round).
__global__ void status_clear_kernel(uint32_t* __restrict__ status, size_t stride, int passes, const uint32_t* __restrict__ n_ptr) {
    const uint32_t n = *n_ptr;
    const uint32_t t = RS_THREADS * RS_ITEMS_MIN;
    const size_t words = (size_t)((n + t - 1) / t) * 256u / 4u;   // uint4 stores
    for (int p = 0; p < passes; ++p) {
        uint4* row = reinterpret_cast<uint4*>(status + (size_t)p * stride);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
            row[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- host-side launch helpers ------------------------------------------------------------------
// status rows are sized for the smallest tile so either variant fits
uint32_t radix_num_tiles(uint32_t capacity) {
    const uint32_t t = RS_THREADS * RS_ITEMS_MIN;
    return (capacity + t - 1) / t;
}

void launch_radix_hist(const uint32_t* keys, const uint32_t* n_ptr, uint32_t capacity, int passes, uint32_t* hist,
                       int sm_count, cudaStream_t stream) {
    const uint32_t chunk = HS_THREADS * HS_ITEMS;
    uint32_t blocks = (capacity + chunk - 1) / chunk;
    const uint32_t cap_blocks = (uint32_t)sm_count * 4u;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks == 0) blocks = 1;
    radix_hist_kernel<<<blocks, HS_THREADS, 0, stream>>>(keys, n_ptr, passes, hist);
}

void launch_onesweep(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                     const uint32_t* n_ptr, uint32_t capacity, uint32_t n_hint, const uint32_t* hist, uint32_t* status,
                     uint32_t* tile_ctr, int shift, int sm_count, cudaStream_t stream, unsigned long long* tl) {
    // n_hint (expected entry count, e.g. last frame's) only picks the tile size; correctness never depends on it
    const bool small = n_hint <= (uint32_t)sm_count * 4u * (RS_THREADS * 16u);
    const uint32_t tile = RS_THREADS * (small ? 8u : 16u);
    uint32_t blocks = (capacity + tile - 1) / tile;
    const uint32_t cap_blocks = (uint32_t)sm_count * 4u;   // persistent: blocks pull tiles from the ticket counter
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks == 0) blocks = 1;
    if (small)
        onesweep_kernel<8><<<blocks, RS_THREADS, 0, stream>>>(keys_in, vals_in, keys_out, vals_out, n_ptr, hist, status,
                                                               tile_ctr, shift, tl);
    else
        onesweep_kernel<16><<<blocks, RS_THREADS, 0, stream>>>(keys_in, vals_in, keys_out, vals_out, n_ptr, hist, status,
                                                                tile_ctr, shift, tl);
}

void launch_status_clear(uint32_t* status, size_t stride, int passes, const uint32_t* n_ptr, int sm_count, cudaStream_t stream) {
    status_clear_kernel<<<sm_count * 2, 256, 0, stream>>>(status, stride, passes, n_ptr);
}

}  // namespace bgs
