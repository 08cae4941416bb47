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
    int solve_groups = 0;                              // independent frame groups (streams) of the grid-fit solve; 0 = chosen per batch (gridfit.cu: solve_groups)
    std::vector<cudaStream_t> aux_streams;             // extra streams of the grouped grid-fit solve
    cudaEvent_t ev_sync = nullptr;
    cudaEvent_t ev_rank = nullptr;                     // "the cluster-rank list has left h_pin" (cluster.cu)
    std::vector<cudaEvent_t> ev_groups;                // one per solve group: "this group's last batch of rounds has finished"
    rclc::Buf d_pts, d_tmp, d_tmp2, d_out, d_flush;   // device scratch
    rclc::Buf h_pin, h_pin2;                           // pinned staging
    rclc::Buf h_up[2];                                 // pinned double buffer of the cloud uploads
    cudaEvent_t ev_up[2] = {nullptr, nullptr};         // "the copy out of h_up[i] has finished"
    int up_flip = 0;
    // the calibration chains (calib.cu): clouds that live across operators, the worker contexts of the batched calibration and its
    // cached grid-fit batch
    rclc::Buf d_chain[10], d_boards, h_chain;
    struct rclc_batch* calib_batch = nullptr;
    void* calib_pool = nullptr;
    int pose_shape = 0, pose_debug = 0, pose_eval_split = 0;      // rclc_ctx_set_pose_test_hooks
    int calib_workers = 0;                             // 0: one per host core (at most one per frame)
    rclc_ctx() { h_pin.pinned = true; h_pin2.pinned = true; h_up[0].pinned = true; h_up[1].pinned = true; h_chain.pinned = true; }
};

namespace rclc {

// Packs a strided host cloud into pinned float4 {x,y,z,I} and uploads it to `dst` (device).
// `st`: the stream the copy is enqueued on (null: ctx->stream).
int upload_cloud_f4(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, float4* dst, cudaStream_t st = nullptr);
void destroy_calib_state(rclc_ctx* ctx);      // calib.cu: the worker pool and the cached batch

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

// ---- exact, order-free sums ------------------------------------------------------------------------------------------------
// A sum that must not depend on the order of its terms (grid shape, warp schedule, CPU vs GPU) is kept in 128-bit fixed point:
// value = (hi * 2^64 + lo) * 2^-80 (sums below 2^47). A term (a double) is floored onto that grid — exact for everything a float
// coordinate or a product of two of them can hold, and below that every term loses the same bits whatever the order — and added
// as an integer; the total is rounded to a double once, to nearest even (fix128_to_double).
struct Fix128 { long long hi; unsigned long long lo; };

__device__ __forceinline__ void fix128_add(Fix128& a, long long hi, unsigned long long lo)
{
    const unsigned long long nl = a.lo + lo;
    a.hi += hi + (nl < lo ? 1 : 0);
    a.lo = nl;
}
__device__ __forceinline__ void fix128_add_double(Fix128& a, double t)
{
    const double s = t * 65536.0;                                 // exact
    const double fl = floor(s);
    const long long hi = __double2ll_rd(s);                       // floor(t * 2^16)
    const unsigned long long lo = __double2ull_rz((s - fl) * 18446744073709551616.0);     // [0, 1) * 2^64, exact
    fix128_add(a, hi, lo);
}
__device__ __forceinline__ void fix128_warp_sum(Fix128& a)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fix128_add(a, __shfl_xor_sync(0xffffffffu, a.hi, o), __shfl_xor_sync(0xffffffffu, a.lo, o));
}
// (hi * 2^64 + lo) * 2^-80 rounded to the nearest double, ties to even — what a conversion of the 128-bit integer does.
__host__ __device__ inline double fix128_to_double(long long hi, unsigned long long lo)
{
    const bool neg = hi < 0;
    unsigned long long h = (unsigned long long)hi, l = lo;
    if (neg) { l = ~l + 1ull; h = ~h + (l == 0ull ? 1ull : 0ull); }
    if (h == 0ull && l == 0ull) return 0.0;
    int p = 0;                                                    // index of the leading one
    {
        unsigned long long w = h ? h : l;
        int b = 0;
        while (w >>= 1) ++b;
        p = h ? 64 + b : b;
    }
    double r;
    if (p <= 52) r = ldexp((double)l, -80);
    else {
        const int s = p - 52;                                     // bits to drop: 1 .. 75
        unsigned long long mant;
        bool above_half, is_half;
        if (s < 64) {
            mant = (h << (64 - s)) | (l >> s);
            const unsigned long long rem = l & ((1ull << s) - 1ull), half = 1ull << (s - 1);
            above_half = rem > half; is_half = rem == half;
        } else {
            const int t = s - 64;
            mant = h >> t;
            const unsigned long long rem_hi = t ? (h & ((1ull << t) - 1ull)) : 0ull;
            const unsigned long long half_hi = t ? (1ull << (t - 1)) : 0ull, half_lo = t ? 0ull : (1ull << 63);
            above_half = rem_hi > half_hi || (rem_hi == half_hi && l > half_lo);
            is_half = rem_hi == half_hi && l == half_lo;
        }
        if (above_half || (is_half && (mant & 1ull))) ++mant;
        r = ldexp((double)mant, s - 80);
    }
    return neg ? -r : r;
}

// A double whose arithmetic is spelled with the _rn intrinsics: nvcc never contracts it into FMAs, so an expression written with
// xd gives the bits the same expression gives on a CPU built without FMA contraction (the checker of the tests is).
struct xd { double v; };
__device__ __forceinline__ xd operator+(xd a, xd b) { return xd{__dadd_rn(a.v, b.v)}; }
__device__ __forceinline__ xd operator-(xd a, xd b) { return xd{__dsub_rn(a.v, b.v)}; }
__device__ __forceinline__ xd operator*(xd a, xd b) { return xd{__dmul_rn(a.v, b.v)}; }
__device__ __forceinline__ xd operator/(xd a, xd b) { return xd{__ddiv_rn(a.v, b.v)}; }
__device__ __forceinline__ xd operator-(xd a) { return xd{-a.v}; }
__device__ __forceinline__ xd xsqrt(xd a) { return xd{__dsqrt_rn(a.v)}; }
__device__ __forceinline__ xd xabs(xd a) { return xd{fabs(a.v)}; }

// pcl::transformPointCloud(Matrix4f) restated: ((m0*x + m1*y) + m2*z) + m3 in float, no FMA contraction.
__device__ __forceinline__ float affine_row_f(float m0, float m1, float m2, float m3, float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z)), m3);
}

} // namespace rclc
