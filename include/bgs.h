/*
 * bgs.h -- C ABI of the B200-native forward splat path (libbgs.so).
 *
 * Drop-in boundary for mosure/bevy_gaussian_splatting's per-view, per-frame GPU work.
 * One Rust render-world system calls bgs_render() in place of BOTH reference call sites:
 *   - system  run_radix_sort::<R>                 (src/sort/radix.rs:616-756, scheduled :97-118)
 *   - command DrawGaussians<R> / DrawGaussianInstanced::render
 *                                                 (src/render/mod.rs:986-992, :1501-1569)
 * The Bevy plugin surface (GaussianSplattingPlugin, PlanarGaussian3dHandle, CloudSettings,
 * GaussianCamera) stays Rust; see INTEGRATION.md for the binding and the system that calls this.
 *
 * Conventions: plain pointers and sizes only; host arrays are borrowed for the duration of the
 * call; device memory is library-owned; every entry point returns a bgs_status and never
 * aborts or throws across the boundary.  A context is single-threaded (Bevy's render thread);
 * distinct contexts (one per GPU) may be used concurrently.  Matrices are column-major f32,
 * exactly as Bevy's `View` uniform / `CloudUniform` hold them.
 */
#ifndef BGS_H
#define BGS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgs_context bgs_context; /* opaque, library-owned */
typedef struct bgs_cloud bgs_cloud;     /* opaque, library-owned */

typedef enum {
    BGS_OK = 0,
    BGS_NOT_READY = 1, /* maps to the reference's silent skip-frame (radix.rs:645-658, mod.rs:1533-1539) */
    BGS_EINVAL = 2,
    BGS_ECUDA = 3,
    BGS_ENOMEM = 4,
    BGS_ENCCL = 5
} bgs_status;

/* The Bevy `View` uniform fields the path reads (src/render/bindings.wgsl:3-9;
 * helpers.wgsl:18-38, transform.wgsl:6, gaussian_2d.wgsl:104, radix.wgsl:92). */
typedef struct {
    float view_from_world[16];
    float clip_from_view[16];
    float clip_from_world[16]; /* used as unjittered_clip_from_world and clip_from_world */
    float world_position[3];
    float viewport[4]; /* x, y, w, h (px); the frame is w x h */
} bgs_view;

/* CloudUniform (src/render/mod.rs:995-1009, bindings.wgsl:13-26), fields this path reads. */
typedef struct {
    float transform[16]; /* model -> world */
    float global_opacity;
    float global_scale;
    uint32_t color_space; /* GaussianColorSpace: 0 SrgbRec709Display, 1 LinRec709Display */
    float time;
    float aabb_min[4];    /* CloudUniform.min / .max = the entity's Aabb.min()/.max() extended with 1.0 */
    float aabb_max[4];    /* (render/mod.rs:1070-1071); read by RasterizeMode::Position only */
} bgs_cloud_uniform;

/* CloudSettings / CloudPipelineKey (src/gaussian/settings.rs:90-133, render/mod.rs:898-909). */
enum { BGS_GAUSSIAN_2D = 0, BGS_GAUSSIAN_3D = 1 };
enum { BGS_RASTERIZE_COLOR = 0, BGS_RASTERIZE_DEPTH = 1, BGS_RASTERIZE_NORMAL = 2, BGS_RASTERIZE_POSITION = 3 };
enum { BGS_DRAW_ALL = 0, BGS_DRAW_SELECTED = 1, BGS_DRAW_HIGHLIGHT_SELECTED = 2 };
enum {
    BGS_FLAG_SORT_ALL = 1u, /* sort all N entries like the reference (culled keyed 0xFFFFFFFF)
                               instead of stream-compacting the visible ones first; same output */
    BGS_FLAG_ASYNC = 2u,    /* bgs_render only enqueues the frame on the context stream and returns;
                               bgs_sync() completes it (frames may be queued back to back, like the
                               reference's command-buffer submission: radix.rs / mod.rs never read back) */
    BGS_FLAG_NO_CHUNKS = 4u,/* never split the frame into front-to-back binning rounds (see BGS_FLAG_CHUNKS);
                               the tile debug hooks need a one-round frame */
    BGS_FLAG_PREMULTIPLIED_OUT = 16u, /* frame = the splat layer alone, premultiplied: (C, 1 - T) -- no implicit black clear,
                               alpha = coverage.  What a compositor needs to put the layer over anything later. */
    BGS_FLAG_BLEND_OVER_TARGET = 32u, /* blend over what the target already holds, dst = src + (1 - src.a) * dst on all four
                               channels -- the reference's PREMULTIPLIED_ALPHA_BLENDING on the view target
                               (render/mod.rs:944-948): several clouds per view (one bgs_render each, far cloud first,
                               mod.rs:398-452) and a scene behind the splats.  Target = out_rgba when it is a device
                               pointer, else the context's frame (which keeps the previous call's result; frames
                               delivered to host memory are copied out after blending). */
    BGS_FLAG_CHUNKS = 8u    /* always bin / tile-sort / blend in front-to-back rank rounds that stop emitting
                               (splat, tile) pairs once every tile has saturated.  Same pixels, bit for bit.
                               Without either flag the library picks rounds when the previous frame had
                               >= 32 (splat, tile) pairs per visible splat and >= 2^24 pairs (USE_OBB records only). */
};
typedef struct {
    uint32_t gaussian_mode;           /* BGS_GAUSSIAN_* */
    uint32_t rasterize_mode;          /* BGS_RASTERIZE_* */
    uint32_t aabb;                    /* 0 = USE_OBB (default), 1 = USE_AABB */
    uint32_t opacity_adaptive_radius; /* default 1 */
    uint32_t draw_mode;               /* BGS_DRAW_* */
    uint32_t radix_sort_depth_bits;   /* 16 | 24 | 32 (RadixSortDepthBits) */
    uint32_t flags;                   /* BGS_FLAG_* */
    uint32_t reserved;
} bgs_settings;

/* Output frame formats.  RGBA8_SRGB / RGBA16F mirror the reference targets
 * (Rgba8UnormSrgb / Rgba16Float, render/mod.rs:917-921); RGBA32F is the premultiplied linear
 * accumulator parity is judged on. */
enum { BGS_FORMAT_RGBA8_SRGB = 0, BGS_FORMAT_RGBA16F = 1, BGS_FORMAT_RGBA32F = 2 };

typedef struct {
    uint32_t n;          /* gaussians in the cloud */
    uint32_t n_visible;  /* in-frustum gaussians this frame */
    uint64_t n_pairs;    /* (splat, tile) pairs emitted this frame (multi-round frames: summed over the rounds;
                            fewer than a one-round frame's when the frame saturated early) */
    uint32_t tiles_x, tiles_y;
    uint32_t width, height;
    uint32_t rounds;          /* binning rounds of the frame: 1, or > 1 on chunked frames (BGS_FLAG_CHUNKS) */
    uint32_t tiles_saturated; /* chunked frames: tiles whose every pixel saturated before the last round */
} bgs_frame_stats;

bgs_status bgs_context_create(int cuda_device, bgs_context** out);
void bgs_context_destroy(bgs_context* ctx);

/* Planar SoA upload, f32 layout (240 B/gaussian): planar_3d.rs:45-54, f32.rs:53-175.
 * pos_vis n*4 (x,y,z,visibility); sh n*48 (interleaved RGB x16); rot n*4 (w,x,y,z);
 * scale_opacity n*4 (linear scale xyz, linear opacity). */
bgs_status bgs_cloud_upload_f32(bgs_context* ctx, uint32_t n, const float* pos_vis, const float* sh,
                                const float* rot_wxyz, const float* scale_opacity, bgs_cloud** out);
/* f16 planar layout (128 B/gaussian): f16.rs:30-56,244-263; planar.wgsl:117-176.
 * sh_packed n*24 words (even coefficient in the low half); rot_scale_opacity n*4 words. */
bgs_status bgs_cloud_upload_f16(bgs_context* ctx, uint32_t n, const float* pos_vis, const uint32_t* sh_packed,
                                const uint32_t* rot_scale_opacity, bgs_cloud** out);
/* f16 planar layout with PRECOMPUTED 3D covariance (the reference's `precompute_covariance_3d` feature): the second plane
 * holds Covariance3dOpacityPacked128 {cov3d: [u32; 3], opacity: u32} (f16.rs:131-170; decode planar.wgsl:133-152)
 * instead of rotation + scale.  Projection then skips quat/scale -> Sigma3D; as in the reference shader
 * (gaussian_3d.wgsl:78-79) neither global_scale nor the model 3x3 touch the stored covariance.  Gaussian3d with
 * RasterizeMode Color / Depth / Position only (Normal and 2DGS need the rotation). */
bgs_status bgs_cloud_upload_f16_cov(bgs_context* ctx, uint32_t n, const float* pos_vis, const uint32_t* sh_packed,
                                    const uint32_t* cov3d_opacity, bgs_cloud** out);
void bgs_cloud_destroy(bgs_cloud* cloud);

/* One view of one cloud: key-gen -> depth radix sort -> projection + SH colour -> tile
 * binning -> per-tile front-to-back blend.  out_rgba is caller-owned (host pointer, or a
 * device pointer when out_is_device_ptr != 0); may be NULL to keep the frame on the device. */
bgs_status bgs_render(bgs_context* ctx, const bgs_cloud* cloud, const bgs_view* view,
                      const bgs_cloud_uniform* uniform, const bgs_settings* settings, void* out_rgba,
                      uint32_t out_format, int out_is_device_ptr);

/* Colour + depth + normal frames of one view in ONE pass (BASELINE.json config 4: "2M-surfel 2dgs cloud with depth+normal
 * outputs").  out_rgba gets the frame bgs_render would produce with `settings` as given; out_depth / out_normal get, bit
 * for bit, the frames bgs_render would produce with rasterize_mode = Depth / Normal (gaussian.wgsl:329-368,
 * material/depth.wgsl:3-11): the splats' geometry, order and alpha do not depend on the colour source, so the extra
 * colours ride along (two more float3 per projected record, six more FMAs per blend).  All three frames share
 * out_format and the host/device kind of the pointers.  Synchronous only. */
bgs_status bgs_render_aux(bgs_context* ctx, const bgs_cloud* cloud, const bgs_view* view,
                          const bgs_cloud_uniform* uniform, const bgs_settings* settings, void* out_rgba,
                          void* out_depth, void* out_normal, uint32_t out_format, int out_is_device_ptr);

/* Wait for every frame enqueued with BGS_FLAG_ASYNC.  BGS_OK: the last frame is complete and valid.
 * BGS_NOT_READY: the last frame's (splat, tile) pair list outgrew its buffer (scene/camera changed a
 * lot); the buffer has been grown -- render that frame again.  A no-op after a synchronous render. */
bgs_status bgs_sync(bgs_context* ctx);

/* Parity / debug hooks (valid after a completed bgs_render on this context). */
/* n*2 words (key, index): the reference's sorted_entry_buffer (sort/mod.rs:323-329). */
bgs_status bgs_debug_sorted_entries(bgs_context* ctx, uint32_t* key_index_pairs);
/* tiles*2 words (start, end) into the per-tile entry list.  The two tile hooks need a one-round frame
 * (BGS_NOT_READY after a multi-round one: render with BGS_FLAG_NO_CHUNKS). */
bgs_status bgs_debug_tile_ranges(bgs_context* ctx, uint32_t* start_end);
/* n_pairs words: front-to-back rank of each (tile, splat) pair, tile-major. */
bgs_status bgs_debug_tile_entries(bgs_context* ctx, uint32_t* ranks, uint64_t capacity);
/* n_visible records of 12 floats: cx, cy, ux, uy, vx, vy, bbox(2 words), r, g, b, opacity;
 * and n_visible gaussian indices (front-to-back rank -> index). */
bgs_status bgs_debug_projected(bgs_context* ctx, float* records, uint32_t* rank_to_index);
bgs_status bgs_frame_stats_get(bgs_context* ctx, bgs_frame_stats* out);

/* Last frame's stage times (CUDA events on the context stream), microseconds:
 * [0] key-gen, [1] depth sort, [2] projection, [3] binning + tile sort + ranges,
 * [4] raster, [5] whole frame.  (Multi-round frames: [3] also holds the earlier rounds' blends,
 * [4] the last round's.) */
bgs_status bgs_stage_times_us(bgs_context* ctx, float out[6]);
const char* bgs_last_error(const bgs_context* ctx);
/* The CUDA stream (cudaStream_t) all of this context's work is launched on. */
void* bgs_context_stream(bgs_context* ctx);
/* The copy/comm stream (cudaStream_t): async frames delivered to host memory or gathered over NCCL are consumed here,
 * so device-side timing of a multi-frame window must cover this stream as well as the render stream. */
void* bgs_context_copy_stream(bgs_context* ctx);
/* Device pointer of the last frame in `out_format` layout (valid until the next render). */
const void* bgs_frame_device_ptr(bgs_context* ctx);
/* Kernel launches issued by the last bgs_render. */
uint32_t bgs_last_launch_count(const bgs_context* ctx);

/* Frame hand-back without a copy.  bgs_frame_export_create allocates a frame target in memory that is
 * exportable as a POSIX file descriptor: pass *out_device_ptr to bgs_render as out_rgba with out_is_device_ptr = 1, and
 * hand *out_fd (+ *out_alloc_bytes) to the graphics API -- Vulkan / wgpu-hal import it with VK_KHR_external_memory_fd
 * (VkImportMemoryFdInfoKHR, handle type OPAQUE_FD) as the memory behind the view-target image or a staging buffer
 * (INTEGRATION.md §6).  The fd is owned by the caller (importing into Vulkan transfers that ownership).
 * bgs_frame_export_import is the consumer side in CUDA terms (another process / library maps the same allocation); the
 * tests use it to prove the exported handle carries the rendered frame.  bgs_frame_export_destroy unmaps and releases a
 * pointer returned by either call. */
bgs_status bgs_frame_export_create(int cuda_device, size_t bytes, void** out_device_ptr, int* out_fd, size_t* out_alloc_bytes);
bgs_status bgs_frame_export_import(int cuda_device, int fd, size_t alloc_bytes, void** out_device_ptr);
void bgs_frame_export_destroy(void* device_ptr);

/* Multi-GPU (one view per GPU, replicated cloud): gather every rank's frame to `root`.
 * nccl_comm is an ncclComm_t.  Enqueued on the context's streams (a frame produced by an async
 * bgs_render is gathered on the copy/comm stream so the next frame overlaps the transfer);
 * bgs_sync() completes it. */
bgs_status bgs_nccl_unique_id(void* out_id128 /* 128 bytes */);
bgs_status bgs_nccl_comm_init(bgs_context* ctx, int nranks, int rank, const void* id128, void** out_comm);
void bgs_nccl_comm_destroy(void* nccl_comm);
bgs_status bgs_gather_frames(bgs_context* ctx, void* nccl_comm, int root, const void* local_frame,
                             void* all_frames, size_t bytes);

/* Copy-engine alternative to bgs_gather_frames (same contract: rank r's frame lands at all_frames + r * bytes on the
 * root), reported beside the NCCL gather by bench.py (`gather_ce`).  The root creates its frame array exportable
 * (bgs_peer_buffer_create -> 64-byte CUDA IPC handle, shipped to the other ranks by the host's own channel); every other
 * rank -- a separate process -- opens it (bgs_peer_buffer_open) and pushes each finished frame with bgs_push_frame: a
 * peer-to-peer copy over NVLink on the sender's copy/comm stream, no SM on either side.  The root itself pushes into its
 * own array (index = its rank).  Completion on the root is the host's to establish (the bench: sync + barrier).
 * bgs_peer_buffer_release(ptr, opened): opened != 0 for pointers from bgs_peer_buffer_open. */
bgs_status bgs_peer_buffer_create(int cuda_device, size_t bytes, void** out_ptr, void* out_handle64);
bgs_status bgs_peer_buffer_open(int cuda_device, const void* handle64, void** out_ptr);
void bgs_peer_buffer_release(void* ptr, int opened);
bgs_status bgs_push_frame(bgs_context* ctx, const void* local_frame, void* remote_frames, int index, size_t bytes);

/* Device-side completion for the copy-engine gather.  bgs_push_frame_signal = bgs_push_frame, then the 32-bit word
 * remote_flags[index] is set to `sequence` by the same stream (ordered after the copy; the words live in the peer buffer,
 * e.g. behind the frames, and start at 0 -- bgs_peer_buffer_create clears the allocation).  bgs_wait_frames makes
 * `cuda_stream` (a CUstream / cudaStream_t of the consumer, NULL = the legacy default stream) wait until
 * flags[0..count) have all reached `sequence` (cyclic >=, so sequences may wrap): work queued behind it sees every
 * frame of that step.  Senders use increasing sequences (frame number + 1).  When local_frame IS the slot
 * (remote_frames + index * bytes) -- the frame was rendered straight into the peer buffer by passing that address to
 * bgs_render as a device target, so the blend kernel's own stores crossed NVLink -- no copy is queued, only the word.  No host round-trip and no cross-process
 * event is involved; the caller must make sure every awaited push is eventually queued, or the stream never resumes.
 * Replaces: the queue-submission order that makes a finished view target visible to its consumer in the reference
 * (src/render/mod.rs:1501-1569 draws inside the view's render pass; here the producer is another process / GPU). */
bgs_status bgs_push_frame_signal(bgs_context* ctx, const void* local_frame, void* remote_frames, int index, size_t bytes,
                                 void* remote_flags, uint32_t sequence);
bgs_status bgs_wait_frames(void* cuda_stream, const void* flags, int count, uint32_t sequence);

#ifdef __cplusplus
}
#endif
#endif
