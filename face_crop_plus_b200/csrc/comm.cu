// comm.cu — the ONE collective of the path: an all-gather of fixed-size per-face metadata records
// (landmarks, global image index, crop matrix, valid) over NCCL / NVLink.  Images are independent end to end, so nothing
// else ever crosses devices: crops, labels and masks stay on the GPU that owns the image.
//
// The records are packed on the device (pack_records_kernel), all-gathered on a side stream and — inside fcp_pipeline —
// overlapped with the parser, which does not depend on them.  No host round trip: the send buffer is filled by a kernel
// that reads the device-side face count, the receive buffer is the caller's.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries — torch's — or the system one): the
// library itself links only libcudart, and single-GPU users never touch NCCL.
#include <dlfcn.h>

#include <cstring>

#include "graphs.h"

namespace fcp {

namespace {

// the slice of nccl.h this file needs (stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int NCCL_FLOAT64 = 8;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return GetUniqueId && CommInitRank && AllGather && CommDestroy; }
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* env = getenv("FCP_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    return api;
}

int nccl_fail(fcp_ctx* ctx, const char* what, ncclResult_t r) {
    const char* msg = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
    return fail(ctx, FCP_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(r) + " (" + msg + ")");
}

}  // namespace

int gather_meta_async(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const double* matrices, const uint8_t* valid,
                      const int32_t* face_count, int cap, int index_base, double* out_records) {
    const size_t rec_doubles = (size_t)(cap + 1) * 20;
    if (ctx->gather_send_cap < rec_doubles) {
        if (ctx->gather_send) cudaFree(ctx->gather_send);
        FCP_CUDA(ctx, cudaMalloc(&ctx->gather_send, rec_doubles * sizeof(double)));
        ctx->gather_send_cap = rec_doubles;
    }
    if (!ctx->comm_stream) {
        FCP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        FCP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_ready, cudaEventDisableTiming));
        FCP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
    }
    // the records depend on the align stage (matrices / valid): fork the side stream here
    FCP_CUDA(ctx, cudaEventRecord(ctx->comm_ready, ctx->stream));
    FCP_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ready, 0));
    {
        StageScope st(ctx, ST_GATHER, ctx->comm_stream, true);
        const bool single = ctx->comm_world <= 1 || !ctx->nccl_comm;
        double* send = single ? out_records : ctx->gather_send;
        FCP_TRY(launch_pack_records(ctx, landmarks, face_img, matrices, valid, face_count, cap, index_base, send, ctx->comm_stream));
        if (!single) {
            ncclResult_t r = nccl().AllGather(send, out_records, rec_doubles, NCCL_FLOAT64, static_cast<ncclComm_t>(ctx->nccl_comm),
                                              ctx->comm_stream);
            if (r != 0) return nccl_fail(ctx, "ncclAllGather", r);
        }
    }
    FCP_CUDA(ctx, cudaEventRecord(ctx->comm_done, ctx->comm_stream));
    return FCP_OK;
}

int gather_meta_join(fcp_ctx* ctx) {
    if (ctx->comm_stream && ctx->comm_done) FCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0));
    return FCP_OK;
}

}  // namespace fcp

using namespace fcp;

extern "C" {

int fcp_comm_unique_id(fcp_ctx* ctx, void* out_id128) {
    if (!ctx || !out_id128) return fail(ctx, FCP_ERR_INVALID, "fcp_comm_unique_id: bad argument");
    if (!nccl().ok()) return fail(ctx, FCP_ERR_STATE, "libnccl.so.2 not found (set FCP_NCCL_LIB)");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != 0) return nccl_fail(ctx, "ncclGetUniqueId", r);
    std::memcpy(out_id128, &id, sizeof id);
    return FCP_OK;
}

int fcp_comm_init(fcp_ctx* ctx, int rank, int world, const void* id128) {
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return fail(ctx, FCP_ERR_INVALID, "fcp_comm_init: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    fcp_comm_destroy(ctx);
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (world == 1) return FCP_OK;
    if (!nccl().ok()) return fail(ctx, FCP_ERR_STATE, "libnccl.so.2 not found (set FCP_NCCL_LIB)");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    ncclResult_t r = nccl().CommInitRank(&comm, world, id, rank);
    if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = comm;
    return FCP_OK;
}

void fcp_comm_destroy(fcp_ctx* ctx) {
    if (!ctx) return;
    if (ctx->nccl_comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
}

int fcp_set_gather(fcp_ctx* ctx, double* out_records, int cap, int index_base) {
    if (!ctx || cap < 0) return fail(ctx, FCP_ERR_INVALID, "fcp_set_gather: bad argument");
    if (out_records && !is_device_ptr(out_records)) return fail(ctx, FCP_ERR_INVALID, "fcp_set_gather: out_records must be device memory");
    ctx->gather_out = out_records;
    ctx->gather_cap = cap;
    ctx->gather_base = index_base;
    return FCP_OK;
}

int fcp_allgather_meta(fcp_ctx* ctx, const float* landmarks, const int32_t* indices, const double* matrices, const uint8_t* valid,
                       int count, int cap, int index_base, double* out_records) {
    if (!ctx || count < 0 || cap < count || !out_records || (count && (!landmarks || !indices)))
        return fail(ctx, FCP_ERR_INVALID, "fcp_allgather_meta: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int world = ctx->nccl_comm ? ctx->comm_world : 1;
    DevIn lms, idx, mats, val, cnt;
    FCP_TRY(lms.init(ctx, landmarks, sizeof(float) * 10 * count));
    FCP_TRY(idx.init(ctx, indices, sizeof(int32_t) * count));
    FCP_TRY(mats.init(ctx, matrices, sizeof(double) * 6 * count));
    FCP_TRY(val.init(ctx, valid, count));
    const int32_t c = count;
    FCP_TRY(cnt.init(ctx, &c, sizeof c));
    DevOut out;
    FCP_TRY(out.init(ctx, out_records, sizeof(double) * 20 * (size_t)(cap + 1) * world));
    FCP_TRY(gather_meta_async(ctx, lms.as<float>(), idx.as<int32_t>(), mats.as<double>(), val.as<uint8_t>(), cnt.as<int32_t>(), cap,
                              index_base, out.as<double>()));
    FCP_TRY(gather_meta_join(ctx));
    FCP_TRY(out.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

}  // extern "C"
