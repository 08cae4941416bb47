// common.cuh — context, error handling and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rclc_b200.h"

namespace rclc {

void set_last_error(const std::string& s);

#define RCLC_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::rclc::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));          \
            return RCLC_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

#define RCLC_CHECK(cond, code, msg)                                                              \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ::rclc::set_last_error(msg);                                                         \
            return (code);                                                                       \
        }                                                                                        \
    } while (0)

#define RCLC_TRY(expr)                                                                           \
    do {                                                                                         \
        int _rc = (expr);                                                                        \
        if (_rc != RCLC_OK) return _rc;                                                          \
    } while (0)

// A growable device (or pinned-host) scratch buffer.
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return RCLC_OK;
        release();
        size_t want = bytes + bytes / 8 + 256;
        if (pinned) RCLC_CUDA(cudaMallocHost(&p, want));
        else RCLC_CUDA(cudaMalloc(&p, want));
        cap = want;
        return RCLC_OK;
    }
    void release()
    {
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

} // namespace rclc

struct rclc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evk0 = nullptr, evk1 = nullptr;       // bracket the last k_sweep launched by rclc_batch_sweep
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    int64_t l2_bytes = 0;
    int64_t launches = 0;
    int flush_parity = 0;
    int solve_groups = 10;                             // independent frame groups (streams) of the grid-fit solve
    std::vector<cudaStream_t> aux_streams;             // extra streams of the grouped grid-fit solve
    cudaEvent_t ev_sync = nullptr;
    std::vector<cudaEvent_t> ev_groups;                // one per solve group: "this group's last batch of rounds has finished"
    rclc::Buf d_pts, d_tmp, d_tmp2, d_out, d_flush;   // device scratch
    rclc::Buf h_pin, h_pin2;                           // pinned staging
    rclc::Buf h_up[2];                                 // pinned double buffer of the cloud uploads
    cudaEvent_t ev_up[2] = {nullptr, nullptr};         // "the copy out of h_up[i] has finished"
    int up_flip = 0;
    rclc_ctx() { h_pin.pinned = true; h_pin2.pinned = true; h_up[0].pinned = true; h_up[1].pinned = true; }
};

namespace rclc {

// Packs a strided host cloud into pinned float4 {x,y,z,I} and uploads it to `dst` (device).
// `st`: the stream the copy is enqueued on (null: ctx->stream).
int upload_cloud_f4(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, float4* dst, cudaStream_t st = nullptr);

// ---- device helpers ----
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pcl::transformPointCloud(Matrix4f) restated: ((m0*x + m1*y) + m2*z) + m3 in float, no FMA contraction.
__device__ __forceinline__ float affine_row_f(float m0, float m1, float m2, float m3, float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z)), m3);
}

} // namespace rclc
