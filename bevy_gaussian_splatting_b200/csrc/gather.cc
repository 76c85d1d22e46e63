// gather.cc -- multi-GPU frame gather: one view per GPU, replicated cloud, and
// ONE collective per frame: every rank's finished frame is sent to `root` over NCCL (NVLink 5 /
// NVSwitch); frames of async renders are gathered on the copy/comm stream so the next frame overlaps the
// transfer.  The reference has no multi-GPU path at all.
//
// NCCL is resolved lazily with dlopen so single-GPU users of libbgs.so never need it, and so a
// host process that already loaded an NCCL (e.g. the torch-bundled one) shares that instance.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <cstring>

#include "../../include/bgs.h"

extern "C" cudaStream_t bgs_internal_gather_begin_(bgs_context* ctx, const void* local_frame, int* slot);   // api.cu
extern "C" void bgs_internal_gather_end_(bgs_context* ctx, int slot);

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.lib || a.ok) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
#define SYM(field, name) *(void**)(&a.field) = dlsym(a.lib, name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(CommCount, "ncclCommCount");
    SYM(CommUserRank, "ncclCommUserRank");
#undef SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv &&
           a.CommCount && a.CommUserRank;
    return a;
}

}  // namespace

extern "C" {

bgs_status bgs_nccl_unique_id(void* out_id128) {
    if (!out_id128) return BGS_EINVAL;
    NcclApi& a = api();
    if (!a.ok) return BGS_ENCCL;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != ncclSuccess) return BGS_ENCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out_id128, &id, 128);
    return BGS_OK;
}

bgs_status bgs_nccl_comm_init(bgs_context* ctx, int nranks, int rank, const void* id128, void** out_comm) {
    if (!ctx || !id128 || !out_comm || nranks < 1 || rank < 0 || rank >= nranks) return BGS_EINVAL;
    NcclApi& a = api();
    if (!a.ok) return BGS_ENCCL;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    // the caller's current device must be the context's device; bgs_context_create set it
    if (a.CommInitRank(&comm, nranks, id, rank) != ncclSuccess) return BGS_ENCCL;
    *out_comm = comm;
    return BGS_OK;
}

void bgs_nccl_comm_destroy(void* nccl_comm) {
    NcclApi& a = api();
    if (a.ok && nccl_comm) a.CommDestroy((ncclComm_t)nccl_comm);
}

bgs_status bgs_gather_frames(bgs_context* ctx, void* nccl_comm, int root, const void* local_frame, void* all_frames,
                             size_t bytes) {
    if (!ctx || !nccl_comm || !local_frame || bytes == 0) return BGS_EINVAL;
    NcclApi& a = api();
    if (!a.ok) return BGS_ENCCL;
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    int nranks = 0, rank = -1;
    if (a.CommCount(comm, &nranks) != ncclSuccess || a.CommUserRank(comm, &rank) != ncclSuccess) return BGS_ENCCL;
    if (root < 0 || root >= nranks) return BGS_EINVAL;
    if (rank == root && !all_frames) return BGS_EINVAL;
    int slot = -1;
    cudaStream_t q = bgs_internal_gather_begin_(ctx, local_frame, &slot);   // comm stream for async library frames
    ncclResult_t r = a.GroupStart();
    if (r != ncclSuccess) return BGS_ENCCL;
    if (rank == root) {
        for (int p = 0; p < nranks && r == ncclSuccess; ++p) {
            char* dst = (char*)all_frames + (size_t)p * bytes;
            if (p == root) {
                if (cudaMemcpyAsync(dst, local_frame, bytes, cudaMemcpyDeviceToDevice, q) != cudaSuccess) r = ncclUnhandledCudaError;
            } else {
                r = a.Recv(dst, bytes, ncclUint8, p, comm, q);
            }
        }
    } else {
        r = a.Send(local_frame, bytes, ncclUint8, root, comm, q);
    }
    const ncclResult_t e = a.GroupEnd();
    bgs_internal_gather_end_(ctx, slot);
    if (r != ncclSuccess || e != ncclSuccess) return BGS_ENCCL;
    return BGS_OK;
}

// ---- copy-engine variant of the gather (measured beside the NCCL one, bench.py `gather_ce`): the root's frame array is a
// cudaMalloc allocation exported with CUDA IPC; every other rank (a separate process) opens it and PUSHES its finished
// frame with a peer-to-peer cudaMemcpyAsync on its own copy/comm stream -- NVLink through the sender's copy engine, no SM
// on either side, nothing queued on the root.  Completion is signalled on the devices as well (bgs_push_frame_signal /
// bgs_wait_frames below): a 32-bit sequence word per slot, stored by the sender's stream after its copy and awaited by a
// stream memory operation on the root -- no cross-process event, no host round-trip.  NCCL stays the default path
// (north_star).
bgs_status bgs_peer_buffer_create(int cuda_device, size_t bytes, void** out_ptr, void* out_handle64) {
    if (!out_ptr || !out_handle64 || bytes == 0) return BGS_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (cudaSetDevice(cuda_device) != cudaSuccess) return BGS_ECUDA;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return BGS_ENOMEM;
    cudaIpcMemHandle_t h;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return BGS_ECUDA; }
    memcpy(out_handle64, &h, 64);
    *out_ptr = p;
    return BGS_OK;
}

bgs_status bgs_peer_buffer_open(int cuda_device, const void* handle64, void** out_ptr) {
    if (!out_ptr || !handle64) return BGS_EINVAL;
    if (cudaSetDevice(cuda_device) != cudaSuccess) return BGS_ECUDA;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return BGS_ECUDA; }
    *out_ptr = p;
    return BGS_OK;
}

void bgs_peer_buffer_release(void* ptr, int opened) {
    if (!ptr) return;
    if (opened) cudaIpcCloseMemHandle(ptr); else cudaFree(ptr);
}

bgs_status bgs_push_frame(bgs_context* ctx, const void* local_frame, void* remote_frames, int index, size_t bytes) {
    if (!ctx || !local_frame || !remote_frames || index < 0 || bytes == 0) return BGS_EINVAL;
    int slot = -1;
    cudaStream_t q = bgs_internal_gather_begin_(ctx, local_frame, &slot);   // copy/comm stream for async library frames
    char* const dst = (char*)remote_frames + (size_t)index * bytes;
    const cudaError_t e = dst == (const char*)local_frame ? cudaSuccess : cudaMemcpyAsync(dst, local_frame, bytes, cudaMemcpyDefault, q);
    bgs_internal_gather_end_(ctx, slot);
    return e == cudaSuccess ? BGS_OK : BGS_ECUDA;
}

// ---- per-slot completion words.  The sender stores `sequence` into remote_flags[index] with a 32-bit memset on the SAME
// stream as the frame copy (stream order: the copy has completed, i.e. its peer writes have landed, before the store is
// issued); the consumer queues cuStreamWaitValue32(flags[i] >= sequence, cyclic compare) on the stream that reads the
// frames.  Both are driver-API calls, resolved with dlopen like NCCL (libbgs.so links cudart only).
namespace {
struct SigApi {
    CUresult (*MemsetD32Async)(CUdeviceptr, unsigned int, size_t, CUstream) = nullptr;
    CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    bool ok = false;
};
SigApi& sig_api() {
    static SigApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return;
        *(void**)(&a.MemsetD32Async) = dlsym(lib, "cuMemsetD32Async");
        *(void**)(&a.StreamWaitValue32) = dlsym(lib, "cuStreamWaitValue32_v2");
        if (!a.StreamWaitValue32) *(void**)(&a.StreamWaitValue32) = dlsym(lib, "cuStreamWaitValue32");
        a.ok = a.MemsetD32Async && a.StreamWaitValue32;
    });
    return a;
}
}  // namespace

bgs_status bgs_push_frame_signal(bgs_context* ctx, const void* local_frame, void* remote_frames, int index, size_t bytes,
                                 void* remote_flags, uint32_t sequence) {
    if (!ctx || !local_frame || !remote_frames || !remote_flags || index < 0 || bytes == 0) return BGS_EINVAL;
    SigApi& a = sig_api();
    if (!a.ok) return BGS_ECUDA;
    int slot = -1;
    cudaStream_t q = bgs_internal_gather_begin_(ctx, local_frame, &slot);
    // local_frame == the slot itself: the frame was RENDERED into the peer buffer (bgs_render with that device target:
    // the blend kernel's stores crossed NVLink), so only the completion word remains to be stored after it
    char* const dst = (char*)remote_frames + (size_t)index * bytes;
    const cudaError_t e = dst == (const char*)local_frame ? cudaSuccess : cudaMemcpyAsync(dst, local_frame, bytes, cudaMemcpyDefault, q);
    CUresult r = CUDA_ERROR_UNKNOWN;
    if (e == cudaSuccess) r = a.MemsetD32Async((CUdeviceptr)((uint32_t*)remote_flags + index), sequence, 1, (CUstream)q);
    bgs_internal_gather_end_(ctx, slot);
    return (e == cudaSuccess && r == CUDA_SUCCESS) ? BGS_OK : BGS_ECUDA;
}

bgs_status bgs_wait_frames(void* cuda_stream, const void* flags, int count, uint32_t sequence) {
    if (!flags || count <= 0) return BGS_EINVAL;
    SigApi& a = sig_api();
    if (!a.ok) return BGS_ECUDA;
    for (int i = 0; i < count; ++i)
        if (a.StreamWaitValue32((CUstream)cuda_stream, (CUdeviceptr)((const uint32_t*)flags + i), sequence, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return BGS_ECUDA;
    return BGS_OK;
}

}  // extern "C"
