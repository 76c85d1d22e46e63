// bin.cu -- stage 4: tile binning / range build (no counterpart in the
// reference, which rasterises one instanced quad per gaussian: src/render/mod.rs:1562-1566).
//
// For every visible splat in front-to-back rank order, emit one (tile id, rank) pair per 16x16
// tile its conservative pixel bbox touches (payload = the splat's record index).  Pair offsets come from a single-pass chained scan
// (decoupled look-back) over the per-splat tile counts, so pairs are emitted in rank order and
// the stable tile-id radix sort that follows yields, per tile, a slice of the GLOBAL depth order.
// range build: boundaries of equal tile ids in the sorted pair keys.
#include "common.cuh"

namespace bgs {

constexpr int BIN_THREADS = 256;
constexpr int BIN_ITEMS = 4;
constexpr int BIN_TILE = BIN_THREADS * BIN_ITEMS;
constexpr int COOP_ITEMS = 8;        // cooperative kernel: up to 8 ranks per thread per round
constexpr uint32_t BIN_TINY = 4u;   // footprints up to this many tiles: written by the owning thread
constexpr uint32_t BIN_BIG = 128u;   // larger than this: global queue, drained by the whole grid   // splats touching more tiles than this are emitted by the whole block

__global__ void __launch_bounds__(BIN_THREADS)
bin_emit_kernel(const SplatRec* __restrict__ recs, const uint32_t* __restrict__ perm, FrameCounters* __restrict__ ctr,
                ChunkCounters* __restrict__ cc, uint32_t* __restrict__ status, int tiles_x, uint32_t capacity,
                uint32_t* __restrict__ pair_keys, uint32_t* __restrict__ pair_vals, uint32_t* __restrict__ sticky_need) {
    __shared__ uint32_t s_wtot[BIN_THREADS / 32];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_nbig;
    __shared__ uint4 s_big[BIN_TILE];   // (rank, pair offset, bbox x, bbox y) of large-footprint splats
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n_vis = ctr->n_vis;
    const uint32_t num_tiles = (n_vis + BIN_TILE - 1) / BIN_TILE;
    while (true) {
        if (t == 0) { s_tile = atomicAdd(&cc->tile_ctr_bin, 1u); s_nbig = 0u; }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;
        const uint32_t r0 = tile * BIN_TILE + t * BIN_ITEMS;   // blocked: 4 consecutive ranks per thread
        uint32_t bx[BIN_ITEMS], by[BIN_ITEMS], cnt[BIN_ITEMS], ri[BIN_ITEMS];
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < BIN_ITEMS; ++j) {
            const uint32_t r = r0 + j;
            cnt[j] = 0u; ri[j] = 0u;
            if (r < n_vis) {
                ri[j] = perm ? __ldg(perm + (n_vis - 1u - r)) : r;   // rank -> record index
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(recs + ri[j]) + 24));
                bx[j] = b.x; by[j] = b.y;
                const uint32_t xlo = b.x & 0xFFFFu, xhi = b.x >> 16, ylo = b.y & 0xFFFFu, yhi = b.y >> 16;
                if (xlo <= xhi && ylo <= yhi)
                    cnt[j] = ((xhi >> 4) - (xlo >> 4) + 1u) * ((yhi >> 4) - (ylo >> 4) + 1u);
            }
            mine += cnt[j];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = (lane < BIN_THREADS / 32) ? s_wtot[lane] : 0u;
            uint32_t total = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            // a frame needing >= 2^30 pairs cannot be represented in the status word: saturate
            const uint32_t agg = total > LB_VMASK ? LB_VMASK : total;
            const uint32_t base = warp_lookback(status, (int)tile, agg);
            if (lane == 0) {
                s_base = base;
                if (tile == num_tiles - 1) {
                    const uint32_t need = base + agg;
                    cc->n_pairs_needed = need;
                    cc->n_pairs = need < capacity ? need : capacity;
                    atomicMax(sticky_need, need);   // survives the per-frame clear: bgs_sync sees every queued frame's need
                }
            }
        }
        __syncthreads();
        uint32_t wprefix = 0u;
        for (int w = 0; w < warp; ++w) wprefix += s_wtot[w];
        uint32_t off = s_base + wprefix + incl - mine;
#pragma unroll
        for (int j = 0; j < BIN_ITEMS; ++j) {
            if (cnt[j] == 0u) continue;
            const uint32_t r = ri[j];
            if (cnt[j] > BIN_BIG) {
                // large footprint: hand it to the whole block (coalesced, parallel emission below)
                const uint32_t q = atomicAdd(&s_nbig, 1u);
                s_big[q] = make_uint4(r, off, bx[j], by[j]);
            } else {
                const uint32_t txlo = (bx[j] & 0xFFFFu) >> 4, txhi = (bx[j] >> 16) >> 4;
                const uint32_t tylo = (by[j] & 0xFFFFu) >> 4, tyhi = (by[j] >> 16) >> 4;
                uint32_t o = off;
                for (uint32_t ty = tylo; ty <= tyhi; ++ty)
                    for (uint32_t tx = txlo; tx <= txhi; ++tx) {
                        if (o < capacity) {
                            pair_keys[o] = ty * (uint32_t)tiles_x + tx;
                            pair_vals[o] = r;
                        }
                        ++o;
                    }
            }
            off += cnt[j];
        }
        __syncthreads();
        const uint32_t nbig = s_nbig;
        for (uint32_t q = 0; q < nbig; ++q) {
            const uint4 b = s_big[q];
            const uint32_t txlo = (b.z & 0xFFFFu) >> 4, txhi = (b.z >> 16) >> 4;
            const uint32_t tylo = (b.w & 0xFFFFu) >> 4, tyhi = (b.w >> 16) >> 4;
            const uint32_t w = txhi - txlo + 1u, total = w * (tyhi - tylo + 1u);
            for (uint32_t i = t; i < total; i += BIN_THREADS) {
                const uint32_t o = b.y + i;
                if (o < capacity) {
                    pair_keys[o] = (tylo + i / w) * (uint32_t)tiles_x + (txlo + i % w);
                    pair_vals[o] = b.x;
                }
            }
        }
        __syncthreads();
    }
    // n_vis == 0: nothing was published; counters stay zero from the per-frame clear
}

// Cooperative variant (all CTAs co-resident).
//   phase 1: count the tiles touched by this CTA's contiguous rank range, publish the CTA total
//   -- grid barrier --
//   phase 2: sum the earlier CTAs' totals in parallel (no chained look-back); emit small footprints
//            directly; push large footprints (rank, pair offset) to a global queue
//   -- grid barrier --
//   phase 3: all warps of the grid drain the queue (front-most splats cover hundreds of tiles and all
//            sit in the first CTAs' ranges: without this the frame waits on a handful of CTAs)
__global__ void __launch_bounds__(BIN_THREADS)
bin_emit_coop_kernel(const SplatRec* __restrict__ recs, const uint32_t* __restrict__ perm, FrameCounters* __restrict__ ctr,
                     ChunkCounters* __restrict__ cc, uint32_t frac_a, uint32_t frac_b, uint32_t num_tiles_total,
                     uint32_t* __restrict__ block_cnt, int tiles_x, uint32_t capacity, uint32_t* __restrict__ pair_keys,
                     uint32_t* __restrict__ pair_vals, uint32_t* __restrict__ q_rank, uint32_t* __restrict__ q_off,
                     uint32_t q_cap, unsigned long long* __restrict__ tl, uint32_t* __restrict__ sticky_need) {
    timeline_stamp(tl, 0);
    __shared__ uint32_t s_wtot[BIN_THREADS / 32];
    __shared__ uint32_t s_wbig[BIN_THREADS / 32];
    __shared__ uint32_t s_wmed[BIN_THREADS / 32];
    __shared__ uint32_t s_mbase;
    __shared__ uint32_t s_red[BIN_THREADS / 32];
    __shared__ unsigned long long s_red64[BIN_THREADS / 32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_qbase;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t G = gridDim.x, b = blockIdx.x;
    const uint32_t n_vis = ctr->n_vis;
    // this round's front-to-back rank range [ra, rb) = n_vis * [frac_a, frac_b) / 65536 (the whole visible set in a
    // one-round frame); empty once every tile has saturated in the earlier rounds (nothing left to blend into)
    uint32_t ra = (uint32_t)((uint64_t)n_vis * frac_a >> 16), rb = (uint32_t)((uint64_t)n_vis * frac_b >> 16);
    if (frac_a != 0u && ld_volatile(&ctr->tiles_done) >= num_tiles_total) {
        if (b == 0 && t == 0) cc->skipped = 1u;
        rb = ra;
    }
    const uint32_t n_rng = rb - ra;
    if (n_rng == 0u) return;   // (grid-uniform) nothing to emit: the round's counters stay zero
    // this CTA's contiguous rank range, cut into sub-tiles of 256 * ipt ranks (ipt chosen so that the whole
    // range is ONE sub-tile whenever it fits 8 items per thread: every CTA then does the same number of rounds)
    const uint32_t rlo = ra + (uint32_t)((uint64_t)b * n_rng / G), rhi = ra + (uint32_t)((uint64_t)(b + 1) * n_rng / G);
    const uint32_t chunk = (n_rng + G - 1) / G;
    uint32_t ipt = (chunk + BIN_THREADS - 1) / BIN_THREADS;
    if (ipt > COOP_ITEMS) ipt = COOP_ITEMS;
    if (ipt == 0) ipt = 1;
    const uint32_t sub = BIN_THREADS * ipt;

    auto tiles_of = [&](uint32_t r, uint32_t& bx, uint32_t& by, uint32_t& ri) -> uint32_t {
        if (r >= rhi) return 0u;
        ri = perm ? __ldg(perm + (n_vis - 1u - r)) : r;   // rank -> record index
        const uint2 bb = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(recs + ri) + 24));
        bx = bb.x; by = bb.y;
        const uint32_t xlo = bb.x & 0xFFFFu, xhi = bb.x >> 16, ylo = bb.y & 0xFFFFu, yhi = bb.y >> 16;
        if (xlo <= xhi && ylo <= yhi) return ((xhi >> 4) - (xlo >> 4) + 1u) * ((yhi >> 4) - (ylo >> 4) + 1u);
        return 0u;
    };

    // ---- phase 1
    uint32_t mine = 0u;
    for (uint32_t base = rlo; base < rhi; base += sub) {
#pragma unroll
        for (int j = 0; j < COOP_ITEMS; ++j) {
            uint32_t bx, by, ri;
            if ((uint32_t)j < ipt) mine += tiles_of(base + t * ipt + j, bx, by, ri);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_red[warp] = mine;
    __syncthreads();
    if (t == 0) {
        uint32_t tot = 0u;
#pragma unroll
        for (int w = 0; w < BIN_THREADS / 32; ++w) tot += s_red[w];
        s_total = tot > LB_VMASK ? LB_VMASK : tot;
        st_volatile(block_cnt + b, s_total);
    }
    timeline_stamp(tl, 1);
    grid_barrier(&cc->barrier, G);
    timeline_stamp(tl, 2);

    // ---- phase 2 (sums saturate at 2^30 - 1: such a frame is rejected by the host)
    uint64_t run64 = 0;
    {
        uint64_t v = 0;
        for (uint32_t p = t; p < b; p += BIN_THREADS) v += ld_volatile(block_cnt + p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red64[warp] = v;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < BIN_THREADS / 32; ++w) run64 += s_red64[w];
    }
    if (b == G - 1 && t == 0) {
        const uint64_t need64 = run64 + s_total;
        const uint32_t need = need64 > LB_VMASK ? LB_VMASK : (uint32_t)need64;
        cc->n_pairs_needed = need;
        cc->n_pairs = need < capacity ? need : capacity;
        atomicMax(sticky_need, need);
    }
    uint32_t run = run64 > LB_VMASK ? LB_VMASK : (uint32_t)run64;
    for (uint32_t base = rlo; base < rhi; base += sub) {
        const uint32_t r0 = base + t * ipt;
        uint32_t bx[COOP_ITEMS], by[COOP_ITEMS], cnt[COOP_ITEMS], ri[COOP_ITEMS];
        uint32_t tmine = 0u, nbig = 0u;
#pragma unroll
        for (int j = 0; j < COOP_ITEMS; ++j) {
            bx[j] = 0u; by[j] = 0u; ri[j] = 0u;
            cnt[j] = ((uint32_t)j < ipt) ? tiles_of(r0 + j, bx[j], by[j], ri[j]) : 0u;
            tmine += cnt[j];
            nbig += cnt[j] > BIN_BIG ? 1u : 0u;
        }
        uint32_t nmed = 0u;
#pragma unroll
        for (int j = 0; j < COOP_ITEMS; ++j) nmed += (cnt[j] > BIN_TINY && cnt[j] <= BIN_BIG) ? 1u : 0u;
        uint32_t incl = tmine, bincl = nbig, mincl = nmed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            const uint32_t z = __shfl_up_sync(0xffffffffu, bincl, o);
            const uint32_t x = __shfl_up_sync(0xffffffffu, mincl, o);
            if (lane >= o) { incl += y; bincl += z; mincl += x; }
        }
        if (lane == 31) { s_wtot[warp] = incl; s_wbig[warp] = bincl; s_wmed[warp] = mincl; }
        __syncthreads();
        uint32_t wprefix = 0u, ttotal = 0u, bprefix = 0u, btotal = 0u, mprefix = 0u, mtotal = 0u;
#pragma unroll
        for (int w = 0; w < BIN_THREADS / 32; ++w) {
            const uint32_t c = s_wtot[w], d = s_wbig[w], e = s_wmed[w];
            if (w < warp) { wprefix += c; bprefix += d; mprefix += e; }
            ttotal += c; btotal += d; mtotal += e;
        }
        if (t == 0) {   // one reservation per round in each global queue
            if (btotal) s_qbase = atomicAdd(&cc->big_count, btotal);
            if (mtotal) s_mbase = atomicAdd(&cc->med_count, mtotal);
        }
        __syncthreads();
        // footprint classes:  <= BIN_TINY tiles: written right here by the owning thread;
        //   <= BIN_BIG: medium queue (front of the queue arrays), drained 32 splats per warp in phase 3;
        //   larger: big queue (back of the queue arrays), one splat per warp in phase 3
        uint32_t off = run + wprefix + incl - tmine;
        uint32_t qat = s_qbase + bprefix + bincl - nbig;
        uint32_t mat = s_mbase + mprefix + mincl - nmed;
#pragma unroll
        for (int j = 0; j < COOP_ITEMS; ++j) {
            if (cnt[j] == 0u) continue;
            if (cnt[j] > BIN_BIG) {
                q_rank[q_cap - 1u - qat] = ri[j]; q_off[q_cap - 1u - qat] = off; ++qat;
            } else if (cnt[j] > BIN_TINY) {
                q_rank[mat] = ri[j]; q_off[mat] = off; ++mat;
            } else {
                const uint32_t txlo = (bx[j] & 0xFFFFu) >> 4, txhi = (bx[j] >> 16) >> 4;
                const uint32_t tylo = (by[j] & 0xFFFFu) >> 4, tyhi = (by[j] >> 16) >> 4;
                uint32_t o = off;
                for (uint32_t ty = tylo; ty <= tyhi; ++ty)
                    for (uint32_t tx = txlo; tx <= txhi; ++tx) {
                        if (o < capacity) {
                            pair_keys[o] = ty * (uint32_t)tiles_x + tx;
                            pair_vals[o] = ri[j];
                        }
                        ++o;
                    }
            }
            off += cnt[j];
        }
        run += ttotal;
        __syncthreads();
    }
    timeline_stamp(tl, 3);
    grid_barrier(&cc->barrier, 2u * G);
    timeline_stamp(tl, 4);

    // ---- phase 3a: medium footprints, 32 per warp: each lane fetches one splat's (record, offset, bbox)
    //      so the memory latency is paid once per 32 splats; then the warp writes them one after another
    const uint32_t nm = ld_volatile(&cc->med_count);
    while (true) {
        uint32_t mb = 0u;
        if (lane == 0) mb = atomicAdd(&cc->med_head, 32u);
        mb = __shfl_sync(0xffffffffu, mb, 0);
        if (mb >= nm) break;
        const uint32_t i = mb + lane;
        uint32_t m_ri = 0u, m_off = 0u, m_txlo = 0u, m_tylo = 0u, m_w = 1u, m_total = 0u;
        if (i < nm) {
            m_ri = __ldcg(q_rank + i); m_off = __ldcg(q_off + i);
            const uint2 bb = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(recs + m_ri) + 24));
            m_txlo = (bb.x & 0xFFFFu) >> 4; m_tylo = (bb.y & 0xFFFFu) >> 4;
            m_w = ((bb.x >> 16) >> 4) - m_txlo + 1u;
            m_total = m_w * (((bb.y >> 16) >> 4) - m_tylo + 1u);
        }
        for (int sI = 0; sI < 32; ++sI) {
            const uint32_t total = __shfl_sync(0xffffffffu, m_total, sI);
            if (total == 0u) continue;
            const uint32_t r = __shfl_sync(0xffffffffu, m_ri, sI), off = __shfl_sync(0xffffffffu, m_off, sI);
            const uint32_t txlo = __shfl_sync(0xffffffffu, m_txlo, sI), tylo = __shfl_sync(0xffffffffu, m_tylo, sI);
            const uint32_t w = __shfl_sync(0xffffffffu, m_w, sI);
            for (uint32_t k = lane; k < total; k += 32) {
                const uint32_t o = off + k;
                if (o < capacity) {
                    const uint32_t qy = k / w;
                    pair_keys[o] = (tylo + qy) * (uint32_t)tiles_x + (txlo + (k - qy * w));
                    pair_vals[o] = r;
                }
            }
        }
    }
    // ---- phase 3b: warps pull large-footprint splats from the back queue, one at a time
    //      (a few splats that each cover thousands of tiles -- the front of a heavy scene -- are cut into up to
    //      16 parts so the whole grid shares them)
    const uint32_t nq = ld_volatile(&cc->big_count);
    uint32_t part_shift = 0u;
    while (part_shift < 4u && ((uint64_t)nq << (part_shift + 2u)) <= (uint64_t)G * (BIN_THREADS / 32)) ++part_shift;
    const uint32_t n_tickets = nq << part_shift;
    while (true) {
        uint32_t q = 0u;
        if (lane == 0) q = atomicAdd(&cc->big_head, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_tickets) break;
        const uint32_t part = q & ((1u << part_shift) - 1u);
        q >>= part_shift;
        const uint32_t r = __ldcg(q_rank + (q_cap - 1u - q)), off = __ldcg(q_off + (q_cap - 1u - q));
        const uint2 bb = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(recs + r) + 24));
        const uint32_t txlo = (bb.x & 0xFFFFu) >> 4, txhi = (bb.x >> 16) >> 4;
        const uint32_t tylo = (bb.y & 0xFFFFu) >> 4, tyhi = (bb.y >> 16) >> 4;
        const uint32_t w = txhi - txlo + 1u, total_all = w * (tyhi - tylo + 1u);
        // this part's slice [i0, total) of the footprint, in multiples of 32 pairs
        const uint32_t per = (((total_all + (1u << part_shift) - 1u) >> part_shift) + 31u) & ~31u;
        const uint32_t i0 = part * per;
        if (i0 >= total_all) continue;
        const uint32_t total = min(total_all, i0 + per);
        uint32_t ty = tylo + (i0 + (uint32_t)lane) / w, tx = txlo + (i0 + (uint32_t)lane) % w;
        const uint32_t dy = 32u / w, dxr = 32u % w;
        for (uint32_t i = i0 + lane; i < total; i += 32) {
            const uint32_t o = off + i;
            if (o < capacity) {
                pair_keys[o] = ty * (uint32_t)tiles_x + tx;
                pair_vals[o] = r;
            }
            ty += dy; tx += dxr;
            if (tx > txhi) { tx -= w; ++ty; }
        }
    }
    timeline_stamp(tl, 5);
}

__global__ void tile_ranges_kernel(const uint32_t* __restrict__ sorted_tile_ids, const uint32_t* __restrict__ n_ptr,
                                   uint2* __restrict__ ranges) {
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t t = sorted_tile_ids[i];
        if (i == 0 || sorted_tile_ids[i - 1] != t) ranges[t].x = i;
        if (i == n - 1 || sorted_tile_ids[i + 1] != t) ranges[t].y = i + 1;
    }
}

void launch_bin_emit(const SplatRec* recs, const uint32_t* perm, FrameCounters* ctr, ChunkCounters* cc, uint32_t* status,
                     int tiles_x, uint32_t capacity, uint32_t* pair_keys, uint32_t* pair_vals, uint32_t n_upper,
                     int sm_count, uint32_t* sticky_need, cudaStream_t stream) {
    uint32_t blocks = (n_upper + BIN_TILE - 1) / BIN_TILE;
    const uint32_t cap_blocks = (uint32_t)sm_count * 4u;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks == 0) blocks = 1;
    bin_emit_kernel<<<blocks, BIN_THREADS, 0, stream>>>(recs, perm, ctr, cc, status, tiles_x, capacity, pair_keys, pair_vals, sticky_need);
}
uint32_t bin_num_tiles(uint32_t n) { return (n + BIN_TILE - 1) / BIN_TILE; }

int bin_coop_blocks_per_sm() {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, bin_emit_coop_kernel, BIN_THREADS, 0) != cudaSuccess) return 0;
    return b;
}
cudaError_t launch_bin_emit_coop(const SplatRec* recs, const uint32_t* perm, FrameCounters* ctr, ChunkCounters* cc,
                                 uint32_t frac_a, uint32_t frac_b, uint32_t num_tiles_total, uint32_t* block_cnt,
                                 int tiles_x, uint32_t capacity, uint32_t* pair_keys, uint32_t* pair_vals,
                                 uint32_t* q_rank, uint32_t* q_off, uint32_t q_cap, unsigned long long* timeline,
                                 uint32_t grid, uint32_t* sticky_need, cudaStream_t stream) {
    void* args[] = {(void*)&recs, (void*)&perm, (void*)&ctr, (void*)&cc, (void*)&frac_a, (void*)&frac_b, (void*)&num_tiles_total,
                    (void*)&block_cnt, (void*)&tiles_x, (void*)&capacity, (void*)&pair_keys,
                    (void*)&pair_vals, (void*)&q_rank, (void*)&q_off, (void*)&q_cap, (void*)&timeline, (void*)&sticky_need};
    return cudaLaunchCooperativeKernel((const void*)bin_emit_coop_kernel, dim3(grid), dim3(BIN_THREADS), args, 0, stream);
}

void launch_tile_ranges(const uint32_t* sorted_tile_ids, const uint32_t* n_ptr, uint2* ranges, uint32_t capacity,
                        int sm_count, cudaStream_t stream) {
    uint32_t blocks = (capacity + 255) / 256;
    const uint32_t cap_blocks = (uint32_t)sm_count * 8u;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks == 0) blocks = 1;
    tile_ranges_kernel<<<blocks, 256, 0, stream>>>(sorted_tile_ids, n_ptr, ranges);
}

}  // namespace bgs
