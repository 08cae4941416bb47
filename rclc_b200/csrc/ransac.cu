// ransac.cu — K1/K2: batched plane-RANSAC hypothesis scoring for chessboard segmentation.
//
// Replaces the arithmetic inside pcl::SACSegmentation<PointXYZI>::segment as configured at
// apps/BoardSegmentation.cpp:64-72 and called at :90 (PCL is not vendored; semantics per SURVEY.md
// Appendix A): SampleConsensusModelPlane::computeModelCoefficients for every sample triplet,
// countWithinDistance for every hypothesis (K1), then selectWithinDistance + the centroid/covariance
// reduction of optimizeModelCoefficients for the chosen model (K2). PCL's adaptive, sequential stopping
// rule is replayed on the host over the device-computed counts (include/rclc_ops.hpp, PlaneSegmentation::replay).
//
// Everything that decides an inlier is float arithmetic in PCL's (Eigen SSE2) association order with
// no FMA contraction, so masks and counts are bit-exact for a given hypothesis set.
#include "ops_dev.cuh"

#include <cmath>

namespace rclc {

constexpr int kScoreThreads = 256;
constexpr int kScoreTile = 2048;        // points staged in shared memory per CTA (24 KB as SoA x,y,z)

__device__ __forceinline__ float plane_dot(float a, float b, float c, float d, float x, float y, float z)
{
    // Eigen Vector4f dot, SSE2 predux: (c0*x + c2*z) + (c1*y + c3*1)
    return __fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(c, z)), __fadd_rn(__fmul_rn(b, y), __fmul_rn(d, 1.0f)));
}

// sac_model_plane.hpp computeModelCoefficients
__global__ void k_planes(const float4* __restrict__ pts, const int* __restrict__ trip, int H, float4* __restrict__ planes,
                         int* __restrict__ valid, int* __restrict__ counts)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    counts[h] = 0;                                              // (k_score, next on the stream, adds into it)
    const float4 p0 = pts[trip[3 * h]], p1 = pts[trip[3 * h + 1]], p2 = pts[trip[3 * h + 2]];
    const float a0 = __fsub_rn(p1.x, p0.x), a1 = __fsub_rn(p1.y, p0.y), a2 = __fsub_rn(p1.z, p0.z);
    const float b0 = __fsub_rn(p2.x, p0.x), b1 = __fsub_rn(p2.y, p0.y), b2 = __fsub_rn(p2.z, p0.z);
    const float d0 = __fdiv_rn(a0, b0), d1 = __fdiv_rn(a1, b1), d2 = __fdiv_rn(a2, b2);
    if ((d0 == d1) && (d2 == d1)) { planes[h] = make_float4(0, 0, 0, 0); valid[h] = 0; return; }   // collinear
    float c0 = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
    float c1 = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
    float c2 = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c2, c2)), __fadd_rn(__fmul_rn(c1, c1), 0.0f));
    const float nrm = __fsqrt_rn(n2);
    c0 = __fdiv_rn(c0, nrm); c1 = __fdiv_rn(c1, nrm); c2 = __fdiv_rn(c2, nrm);
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(c0, p0.x), __fmul_rn(c2, p0.z)), __fadd_rn(__fmul_rn(c1, p0.y), 0.0f));
    planes[h] = make_float4(c0, c1, c2, __fmul_rn(-1.0f, dot));
    valid[h] = 1;
}

// countWithinDistance for all hypotheses: a CTA stages a tile of points in shared memory (SoA, broadcast
// reads), each thread owns hypotheses h = tid, tid+T, ... and keeps their coefficients in registers.
__global__ void __launch_bounds__(kScoreThreads) k_score(const float4* __restrict__ pts, int64_t n, int tile, const float4* __restrict__ planes,
                                                         const int* __restrict__ valid, int H, float thr_le, int* __restrict__ counts)
{
    __shared__ float sx[kScoreTile], sy[kScoreTile], sz[kScoreTile];
    const int64_t base = (int64_t)blockIdx.x * tile;                          // tile <= kScoreTile: small clouds take small tiles (more CTAs, shorter loops)
    const int m = (int)min((int64_t)tile, n - base);
    for (int i = threadIdx.x; i < m; i += kScoreThreads) {
        const float4 q = __ldg(pts + base + i);
        sx[i] = q.x; sy[i] = q.y; sz[i] = q.z;
    }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += kScoreThreads) {
        if (!valid[h]) continue;
        const float4 pl = planes[h];
        int cnt = 0;
#pragma unroll 8
        for (int i = 0; i < m; ++i)
            cnt += (fabsf(plane_dot(pl.x, pl.y, pl.z, pl.w, sx[i], sy[i], sz[i])) <= thr_le) ? 1 : 0;
        if (cnt) atomicAdd(&counts[h], cnt);
    }
}

// selectWithinDistance mask + first/second moments of the inliers, per-CTA partials.
// The moments are EXACT, order-free sums: every term (x, y, z and the products of two floats, which a double holds exactly) is
// put on the fixed-point grid of 2^-80 (floor; nothing is lost for coordinates of 8 um and more, and what is lost below is lost
// the same way for every order of the points) and added as a 128-bit integer. The CPU checker of the tests defines the sums the
// same way, so the refined plane and the re-selected mask agree with it bit for bit — the reference's own accumulation order
// (PCL's computeMeanAndCovarianceMatrix) is unpinned.
constexpr int kSelThreads = 256;
constexpr int kSelTerms = 9;            // xx xy xz yy yz zz x y z

__global__ void __launch_bounds__(kSelThreads) k_select(const float4* __restrict__ pts, int64_t n, float4 pl, float thr_le,
                                                        uint8_t* __restrict__ mask, unsigned long long* __restrict__ partials /* [grid][2 * kSelTerms + 1] */)
{
    Fix128 a[kSelTerms];
#pragma unroll
    for (int k = 0; k < kSelTerms; ++k) { a[k].hi = 0; a[k].lo = 0; }
    unsigned long long cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads) {
        const float4 q = __ldg(pts + i);
        const bool in = fabsf(plane_dot(pl.x, pl.y, pl.z, pl.w, q.x, q.y, q.z)) <= thr_le;
        mask[i] = in ? 1 : 0;
        if (in) {
            const double x = q.x, y = q.y, z = q.z;
            fix128_add_double(a[0], __dmul_rn(x, x)); fix128_add_double(a[1], __dmul_rn(x, y)); fix128_add_double(a[2], __dmul_rn(x, z));
            fix128_add_double(a[3], __dmul_rn(y, y)); fix128_add_double(a[4], __dmul_rn(y, z)); fix128_add_double(a[5], __dmul_rn(z, z));
            fix128_add_double(a[6], x); fix128_add_double(a[7], y); fix128_add_double(a[8], z);
            ++cnt;
        }
    }
    __shared__ unsigned long long s_red[kSelThreads / 32][2 * kSelTerms + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kSelTerms; ++k) fix128_warp_sum(a[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        for (int k = 0; k < kSelTerms; ++k) { s_red[warp][2 * k] = (unsigned long long)a[k].hi; s_red[warp][2 * k + 1] = a[k].lo; }
        s_red[warp][2 * kSelTerms] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < kSelTerms) {
        Fix128 t;
        t.hi = 0; t.lo = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) fix128_add(t, (long long)s_red[w][2 * threadIdx.x], s_red[w][2 * threadIdx.x + 1]);
        partials[(size_t)blockIdx.x * (2 * kSelTerms + 1) + 2 * threadIdx.x] = (unsigned long long)t.hi;
        partials[(size_t)blockIdx.x * (2 * kSelTerms + 1) + 2 * threadIdx.x + 1] = t.lo;
    }
    if (threadIdx.x == kSelTerms) {
        unsigned long long c = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) c += s_red[w][2 * kSelTerms];
        partials[(size_t)blockIdx.x * (2 * kSelTerms + 1) + 2 * kSelTerms] = c;
    }
}

// `(double)|dot| < thr` (PCL compares std::abs(float) with a double threshold) as a float `<=` test
static float threshold_le(double thr)
{
    float tf = (float)thr;
    if (!((double)tf < thr)) tf = std::nextafterf(tf, -INFINITY);
    return tf;
}

// smallest eigenvector of a symmetric 3x3 (cyclic Jacobi, double) — pcl::eigen33's job
static void smallest_eigvec3(const double cov[9], double ev[3])
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = cov[r * 3 + c];
    for (int sweep = 0; sweep < 64; ++sweep) {
        if (a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2] < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
            }
    }
    int best = 0;
    for (int k = 1; k < 3; ++k) if (a[k][k] < a[best][best]) best = k;
    for (int k = 0; k < 3; ++k) ev[k] = v[k][best];
}

int dev_ransac_score(rclc_ctx* ctx, const float4* d_pts, int64_t n, const int32_t* triplets, int32_t H, double thr, float* planes_out,
                     int32_t* valid_out, int32_t* counts_out)
{
    const size_t o_pl = ((size_t)H * 12 + 255) & ~(size_t)255, o_va = o_pl + (((size_t)H * 16 + 255) & ~(size_t)255);
    const size_t o_cn = o_va + (((size_t)H * 4 + 255) & ~(size_t)255), total = o_cn + (size_t)H * 4;
    RCLC_TRY(ctx->d_tmp.reserve(total));
    char* base = static_cast<char*>(ctx->d_tmp.p);
    int* d_trip = reinterpret_cast<int*>(base);
    float4* d_planes = reinterpret_cast<float4*>(base + o_pl);
    int* d_valid = reinterpret_cast<int*>(base + o_va);
    int* d_counts = reinterpret_cast<int*>(base + o_cn);
    // page-locked staging both ways: a copy to or from pageable memory goes through the driver's own staging buffer under its lock,
    // which serialises the frames of the batched chain
    RCLC_TRY(ctx->h_pin2.reserve(total + 256));
    char* hb = ctx->h_pin2.as<char>();
    std::memcpy(hb, triplets, (size_t)H * 12);
    RCLC_CUDA(cudaMemcpyAsync(d_trip, hb, (size_t)H * 12, cudaMemcpyHostToDevice, ctx->stream));
    const int tile = n <= 262144 ? 512 : kScoreTile;
    k_planes<<<(H + 127) / 128, 128, 0, ctx->stream>>>(d_pts, d_trip, H, d_planes, d_valid, d_counts);
    k_score<<<(unsigned)((n + tile - 1) / tile), kScoreThreads, 0, ctx->stream>>>(d_pts, n, tile, d_planes, d_valid, H, threshold_le(thr), d_counts);
    ctx->launches += 2;
    RCLC_CUDA(cudaGetLastError());
    RCLC_CUDA(cudaMemcpyAsync(hb + o_pl, base + o_pl, total - o_pl, cudaMemcpyDeviceToHost, ctx->stream));      // planes | valid | counts
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (planes_out) std::memcpy(planes_out, hb + o_pl, (size_t)H * 16);
    if (valid_out) std::memcpy(valid_out, hb + o_va, (size_t)H * 4);
    std::memcpy(counts_out, hb + o_cn, (size_t)H * 4);
    return RCLC_OK;
}

int dev_ransac_select(rclc_ctx* ctx, const float4* d_pts, int64_t n, const float plane[4], double thr, const uint8_t** d_mask_out, float refined_out[4],
                      float centroid_out[3], int64_t* n_inliers)
{
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + kSelThreads * 4 - 1) / (kSelThreads * 4), (int64_t)ctx->sm_count * 4));
    constexpr int kW = 2 * kSelTerms + 1;
    const size_t o_part = ((size_t)n + 255) & ~(size_t)255;
    RCLC_TRY(ctx->d_tmp.reserve(o_part + (size_t)grid * kW * 8));
    uint8_t* d_mask = ctx->d_tmp.as<uint8_t>();
    unsigned long long* d_part = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx->d_tmp.p) + o_part);
    k_select<<<grid, kSelThreads, 0, ctx->stream>>>(d_pts, n, make_float4(plane[0], plane[1], plane[2], plane[3]), threshold_le(thr), d_mask, d_part);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    if (d_mask_out) *d_mask_out = d_mask;
    RCLC_TRY(ctx->h_pin2.reserve((size_t)grid * kW * 8));
    unsigned long long* part = ctx->h_pin2.as<unsigned long long>();
    RCLC_CUDA(cudaMemcpyAsync(part, d_part, (size_t)grid * kW * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    // fold the 128-bit partial sums (integer addition: any order) and put them back on the double grid, rounded to nearest once
    double accu[9];
    int64_t cnt = 0;
    for (int k = 0; k < kSelTerms; ++k) {
        __int128 t = 0;
        for (int c = 0; c < grid; ++c)
            t += (__int128)(((unsigned __int128)part[(size_t)c * kW + 2 * k] << 64) | (unsigned __int128)part[(size_t)c * kW + 2 * k + 1]);
        accu[k] = fix128_to_double((long long)(t >> 64), (unsigned long long)t);
    }
    for (int c = 0; c < grid; ++c) cnt += (int64_t)part[(size_t)c * kW + 2 * kSelTerms];
    if (n_inliers) *n_inliers = cnt;
    if (refined_out) {
        // optimizeModelCoefficients (sac_model_plane.hpp): needs more than 3 inliers
        if (cnt <= 3) {
            for (int k = 0; k < 4; ++k) refined_out[k] = plane[k];
            if (centroid_out) centroid_out[0] = centroid_out[1] = centroid_out[2] = 0.f;
        } else {
            for (int k = 0; k < 9; ++k) accu[k] /= (double)cnt;
            double cov[9];
            cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
            cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
            cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
            double ev[3];
            smallest_eigvec3(cov, ev);
            if (ev[0] * plane[0] + ev[1] * plane[1] + ev[2] * plane[2] < 0) { ev[0] = -ev[0]; ev[1] = -ev[1]; ev[2] = -ev[2]; }
            const float c0 = (float)ev[0], c1 = (float)ev[1], c2 = (float)ev[2];
            const float m0 = (float)accu[6], m1 = (float)accu[7], m2 = (float)accu[8];
            const float dot = (c0 * m0 + c2 * m2) + (c1 * m1 + 0.0f);
            refined_out[0] = c0; refined_out[1] = c1; refined_out[2] = c2; refined_out[3] = -1.0f * dot;
            if (centroid_out) { centroid_out[0] = m0; centroid_out[1] = m1; centroid_out[2] = m2; }
        }
    }
    return RCLC_OK;
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_ransac_plane_score(rclc_ctx* ctx, const void* pts, int stride, int64_t n, const int32_t* triplets, int32_t H,
                            double thr, float* planes_out, int32_t* valid_out, int32_t* counts_out)
{
    RCLC_CHECK(ctx && pts && triplets && counts_out && n > 0 && H > 0 && stride >= 16, RCLC_ERR_INVALID, "bad argument");
    for (int k = 0; k < 3 * H; ++k) RCLC_CHECK(triplets[k] >= 0 && triplets[k] < n, RCLC_ERR_INVALID, "triplet index out of range");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, 12, n, ctx->d_pts.as<float4>()));
    return dev_ransac_score(ctx, ctx->d_pts.as<float4>(), n, triplets, H, thr, planes_out, valid_out, counts_out);
}

int rclc_ransac_plane_select(rclc_ctx* ctx, const void* pts, int stride, int64_t n, const float plane[4], double thr,
                             uint8_t* mask_out, float refined_out[4], float centroid_out[3], int64_t* n_inliers)
{
    RCLC_CHECK(ctx && pts && plane && mask_out && n > 0 && stride >= 16, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, 12, n, ctx->d_pts.as<float4>()));
    const uint8_t* d_mask = nullptr;
    RCLC_TRY(dev_ransac_select(ctx, ctx->d_pts.as<float4>(), n, plane, thr, &d_mask, refined_out, centroid_out, n_inliers));
    RCLC_CUDA(cudaMemcpyAsync(mask_out, d_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    return RCLC_OK;
}

} // extern "C"
