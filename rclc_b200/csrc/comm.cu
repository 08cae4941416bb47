// comm.cu — the multi-GPU exchanges of the path, behind the C ABI: NCCL called from C++ on the context's compute stream.
//
// The path shards by FRAME (one rank per GPU, frames f0 .. f1 of the reference's `omp parallel for`, apps/Calibrate.cpp:90-107);
// two places exchange data:
//
//   * the extrinsic solve (calib_opt, apps/Calibrate.cpp:12-67 -> pose_reprj_opt, include/opt/opt_reprj_calib.h:21-66) needs
//     every frame's corners: ONE ncclAllGather of (35 x 3 + 35 x 2) doubles per frame, then every rank runs the same
//     deterministic solve on the same data in the same order — identical bits on every rank, zero per-iteration collectives
//     (the alternative, an all-reduce of the 6 x 6 system in each of the ~1000 LM iterations, costs ~1000 collective latencies);
//   * ONE frame whose points are split across ranks: the 82 x 16 accumulator table of a sweep is all-reduced before the
//     non-linear residual (include/opt/board_fitting_cost.h:211-235). The accumulators are exact sums of 24-bit values and
//     integer counts held in doubles, so the fp64 sum over ranks is exact in any order: split == unsplit, bit for bit.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, or the copy the process has already loaded — PyTorch's): the library has
// no link-time dependency on it, and a caller that never creates a communicator never needs it.
#include "gridfit.cuh"
#include "ops_dev.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

struct rclc_comm {
    rclc_ctx* ctx = nullptr;
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1;
    rclc::Buf d_send, d_recv, d_cnt;
    int64_t collectives = 0;
};

namespace rclc {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GetErrorString;
    });
    return api;
}

#define RCLC_NCCL(expr)                                                                                   \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) {                                                                          \
            ::rclc::set_last_error(std::string(#expr) + ": " + nccl_api().GetErrorString(_r));            \
            return RCLC_ERR_CUDA;                                                                         \
        }                                                                                                 \
    } while (0)

// gathered [world][Fmax] blocks -> the frames of all ranks back to back, rank-major
__global__ void k_compact_frames(const double* __restrict__ gathered, const int* __restrict__ counts, int world, int Fmax, int per_frame, double* __restrict__ out)
{
    int f0 = 0;
    for (int r = 0; r < world; ++r) {
        const int fr = counts[r];
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)fr * per_frame; i += (int64_t)gridDim.x * blockDim.x)
            out[(int64_t)f0 * per_frame + i] = gathered[(int64_t)r * Fmax * per_frame + i];
        f0 += fr;
    }
}

} // namespace
} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_comm_get_unique_id(void* id_out_128)
{
    RCLC_CHECK(id_out_128, RCLC_ERR_INVALID, "null argument");
    static_assert(sizeof(ncclUniqueId) == RCLC_COMM_ID_BYTES, "RCLC_COMM_ID_BYTES");
    NcclApi& api = nccl_api();
    RCLC_CHECK(api.ok, RCLC_ERR_UNSUPPORTED, "libnccl.so.2 not found: multi-GPU exchanges need NCCL");
    ncclUniqueId id;
    RCLC_NCCL(api.GetUniqueId(&id));
    std::memcpy(id_out_128, &id, sizeof(id));
    return RCLC_OK;
}

int rclc_comm_create(rclc_ctx* ctx, const void* id_128, int32_t rank, int32_t world, rclc_comm** out)
{
    RCLC_CHECK(ctx && id_128 && out && world >= 1 && rank >= 0 && rank < world, RCLC_ERR_INVALID, "bad argument");
    NcclApi& api = nccl_api();
    RCLC_CHECK(api.ok, RCLC_ERR_UNSUPPORTED, "libnccl.so.2 not found: multi-GPU exchanges need NCCL");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id_128, sizeof(id));
    rclc_comm* c = new rclc_comm();
    c->ctx = ctx; c->rank = rank; c->world = world;
    const ncclResult_t r = api.CommInitRank(&c->nccl, world, id, rank);
    if (r != ncclSuccess) { set_last_error(std::string("ncclCommInitRank: ") + api.GetErrorString(r)); delete c; return RCLC_ERR_CUDA; }
    *out = c;
    return RCLC_OK;
}

int rclc_comm_destroy(rclc_comm* c)
{
    if (!c) return RCLC_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->nccl) nccl_api().CommDestroy(c->nccl);
    c->d_send.release(); c->d_recv.release(); c->d_cnt.release();
    delete c;
    return RCLC_OK;
}

int64_t rclc_comm_collectives(rclc_comm* c) { return c ? c->collectives : 0; }

int rclc_pose_refine_sharded(rclc_comm* comm, const rclc_config* cfg, const double* pc_local, const double* img_local, int32_t F_local, int32_t C,
                             double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations, int32_t* F_total_out)
{
    RCLC_CHECK(comm && cfg && Tcl_io && F_local >= 0 && C > 0 && (F_local == 0 || (pc_local && img_local)), RCLC_ERR_INVALID, "bad argument");
    rclc_ctx* ctx = comm->ctx;
    NcclApi& api = nccl_api();
    RCLC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int W = comm->world, per_frame = C * 5;                     // 3 doubles per corner + 2 per image point
    // the ranks' frame counts (ragged shards), then the payload padded to the largest shard
    RCLC_TRY(comm->d_cnt.reserve((size_t)(W + 1) * 4));
    RCLC_TRY(ctx->h_chain.reserve(256 + (size_t)W * 4));
    int* h_cnt = ctx->h_chain.as<int>() + 48;
    int* d_cnt = comm->d_cnt.as<int>();
    h_cnt[0] = F_local;
    RCLC_CUDA(cudaMemcpyAsync(d_cnt + W, h_cnt, 4, cudaMemcpyHostToDevice, st));
    RCLC_NCCL(api.AllGather(d_cnt + W, d_cnt, 1, ncclInt32, comm->nccl, st));
    RCLC_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)W * 4, cudaMemcpyDeviceToHost, st));
    RCLC_CUDA(cudaStreamSynchronize(st));
    comm->collectives += 1;
    int Fmax = 0, F = 0;
    for (int r = 0; r < W; ++r) { Fmax = std::max(Fmax, h_cnt[r]); F += h_cnt[r]; }
    if (F_total_out) *F_total_out = F;
    RCLC_CHECK(F > 0, RCLC_ERR_INVALID, "no rank has a frame");
    const size_t blk = (size_t)Fmax * per_frame;
    RCLC_TRY(comm->d_send.reserve(blk * 8 + 64));
    RCLC_TRY(comm->d_recv.reserve(blk * 8 * W + 64));
    double* d_send = comm->d_send.as<double>();
    // one frame = its corners then its image points, so that a rank's shard is one contiguous block of F_local * per_frame doubles
    std::vector<double> packed((size_t)F_local * per_frame);
    for (int f = 0; f < F_local; ++f) {
        std::memcpy(&packed[(size_t)f * per_frame], pc_local + (size_t)f * C * 3, (size_t)C * 24);
        std::memcpy(&packed[(size_t)f * per_frame + (size_t)C * 3], img_local + (size_t)f * C * 2, (size_t)C * 16);
    }
    if (F_local > 0) RCLC_CUDA(cudaMemcpyAsync(d_send, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice, st));
    RCLC_NCCL(api.AllGather(d_send, comm->d_recv.p, blk, ncclDouble, comm->nccl, st));
    comm->collectives += 1;
    // rank-major compaction, then the split into the solver's two arrays (corners | image points), all on the device
    RCLC_TRY(ctx->d_tmp.reserve(((size_t)F * per_frame * 2 + 16) * 8));
    double* d_all = ctx->d_tmp.as<double>();
    double* d_pc = d_all + (size_t)F * per_frame;
    double* d_img = d_pc + (size_t)F * C * 3;
    double* d_T = d_img + (size_t)F * C * 2;
    k_compact_frames<<<64, 256, 0, st>>>(comm->d_recv.as<double>(), d_cnt, W, Fmax, per_frame, d_all);
    ctx->launches += 1;
    RCLC_CUDA(cudaMemcpy2DAsync(d_pc, (size_t)C * 24, d_all, (size_t)per_frame * 8, (size_t)C * 24, (size_t)F, cudaMemcpyDeviceToDevice, st));
    RCLC_CUDA(cudaMemcpy2DAsync(d_img, (size_t)C * 16, d_all + (size_t)C * 3, (size_t)per_frame * 8, (size_t)C * 16, (size_t)F, cudaMemcpyDeviceToDevice, st));
    RCLC_CUDA(cudaMemcpyAsync(d_T, Tcl_io, 7 * 8, cudaMemcpyHostToDevice, st));
    RCLC_CUDA(cudaStreamSynchronize(st));                              // (packed and Tcl_io are pageable)
    return pose_refine_dev(ctx, cfg, d_pc, d_img, F, C, d_T, Tcl_io, initial_cost, final_cost, iterations);
}

int rclc_batch_allreduce_acc(rclc_batch* b, rclc_comm* comm)
{
    RCLC_CHECK(b && comm && b->ctx == comm->ctx, RCLC_ERR_INVALID, "the batch and the communicator must share a context");
    NcclApi& api = nccl_api();
    RCLC_CUDA(cudaSetDevice(b->ctx->device));
    // every frame's accumulator tables (acc_stride doubles per frame; a plain sweep fills pose slot 0): exact sums, so the
    // order in which NCCL adds the ranks' shares does not matter
    RCLC_NCCL(api.AllReduce(b->v.acc, b->v.acc, (size_t)b->n_frames * (size_t)b->v.acc_stride, ncclDouble, ncclSum, comm->nccl, b->ctx->stream));
    comm->collectives += 1;
    return RCLC_OK;
}

} // extern "C"
