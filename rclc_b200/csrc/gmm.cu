// gmm.cu — K3: 1-D, 2-component Gaussian-mixture EM on laser reflectance.
//
// Replaces GMMFit::fit_1d (src/GMMFit.cpp:7-108) + GaussianFunction::eval (src/GaussianFunction.cpp:7-23).
// The reference makes six passes and allocates two N x K matrices per iteration; here an iteration of a large cloud is ONE
// streaming pass (k_gmm_iter: posterior + moments about the old mean), and the whole fit of a chain-sized cloud is one
// cluster launch that keeps points and posteriors in registers (k_gmm_fit_cluster: the reference's two passes, second moment
// about the NEW mean as GMMFit.cpp:79-85). Reductions are fixed-order, so results are deterministic.
#include "gridfit.cuh"
#include "ops_dev.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace rclc {

constexpr int kGmmThreads = 256;
constexpr int kGmmMaxCtas = 1184;      // 148 SMs x 8

struct GmmState {          // per frame
    double mu[2], sigma[2], prior[2];
    double mu_new[2], denom[2];
    unsigned int counter;
    unsigned int pad;
};

struct GmmView {
    const float* x;            // intensity base
    int64_t frame_stride;      // in floats
    int elem_stride;           // in floats (1 or 4)
    const int64_t* n_pts;      // [F]
    GmmState* st;              // [F]
    double* partials;          // [F][kGmmMaxCtas][6]
};

// posterior P(k | x) exactly as GMMFit.cpp:45-65 evaluates it (no FMA contraction)
__device__ __forceinline__ void posterior(double x, const double* mu, const double* inv2s2, const double* factor,
                                          const double* prior, double& p0, double& p1)
{
    const double d0 = __dsub_rn(x, mu[0]), d1 = __dsub_rn(x, mu[1]);
    const double l0 = __dmul_rn(factor[0], exp(__dmul_rn(__dmul_rn(inv2s2[0], d0), d0)));
    const double l1 = __dmul_rn(factor[1], exp(__dmul_rn(__dmul_rn(inv2s2[1], d1), d1)));
    const double a0 = __dmul_rn(l0, prior[0]), a1 = __dmul_rn(l1, prior[1]);
    const double px = __dadd_rn(a0, a1);
    p0 = __ddiv_rn(a0, px);
    p1 = __ddiv_rn(a1, px);
}

// One EM iteration of the large-cloud path in ONE pass. The reference's second pass (the variance about the NEW mean,
// GMMFit.cpp:79-85) repeats the posterior of the first — two exp and two divisions per point — only because the new mean is not
// known yet; the same number follows from moments about the OLD mean: with e = x - mu_old,
//   mu_new = mu_old + S(p e) / S(p),      S(p (x - mu_new)^2) / S(p) = S(p e^2) / S(p) - (mu_new - mu_old)^2.
// Shifting by the old mean keeps the subtraction harmless (the mean moves by a fraction of sigma per iteration): against the
// two-pass kernel the parameters agree to 1e-13 relative after ten iterations, far inside the 1e-9 this path is held to against
// the oracle, and the labels are the same (tests/test_gpu_parity.py). Reductions are two-stage and fixed-order (per-CTA partials, the last CTA folds them): deterministic.
__global__ void __launch_bounds__(kGmmThreads) k_gmm_iter(GmmView v)
{
    const int f = blockIdx.y;
    const int64_t n = v.n_pts[f];
    GmmState& st = v.st[f];
    if (n == 0) return;
    __shared__ double s_par[8];
    __shared__ double s_red[kGmmThreads / 32][6];
    __shared__ bool s_last;
    if (threadIdx.x < 2) {
        const int k = threadIdx.x;
        const double sg = st.sigma[k];
        s_par[k] = st.mu[k];
        s_par[2 + k] = __ddiv_rn(-1.0, __dmul_rn(2.0, __dmul_rn(sg, sg)));                 // -1.0 / (2*sigma_sqr)
        s_par[4 + k] = __ddiv_rn(1.0, __dmul_rn(sg, sqrt(2.0 * 3.14159265358979323846)));  // 1/(sigma*sqrt(2 pi))
        s_par[6 + k] = st.prior[k];
    }
    __syncthreads();
    const float* x = v.x + (size_t)f * v.frame_stride;
    const double m0 = s_par[0], m1 = s_par[1];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)kGmmThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kGmmThreads) {
        const double xi = (double)__ldg(x + i * v.elem_stride);
        double p0, p1;
        posterior(xi, s_par, s_par + 2, s_par + 4, s_par + 6, p0, p1);
        const double e0 = xi - m0, e1 = xi - m1;
        a[0] += p0; a[1] += p1;
        a[2] = fma(p0, e0, a[2]); a[3] = fma(p1, e1, a[3]);
        a[4] = fma(p0 * e0, e0, a[4]); a[5] = fma(p1 * e1, e1, a[5]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) a[q] = warp_sum(a[q]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int q = 0; q < 6; ++q) s_red[warp][q] = a[q];
    __syncthreads();
    double* part = v.partials + ((size_t)f * kGmmMaxCtas + blockIdx.x) * 6;
    if (threadIdx.x < 6) {
        double t = 0;
        for (int w = 0; w < kGmmThreads / 32; ++w) t += s_red[w][threadIdx.x];
        part[threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) {                    // last CTA: fold the per-CTA partials in a fixed order
        double t[6] = {0, 0, 0, 0, 0, 0};
        const double* all = v.partials + (size_t)f * kGmmMaxCtas * 6;
        for (unsigned int c = lane; c < gridDim.x; c += 32)
            for (int q = 0; q < 6; ++q) t[q] += __ldcg(all + (size_t)c * 6 + q);
        for (int q = 0; q < 6; ++q) t[q] = warp_sum(t[q]);
        if (lane == 0) {
            for (int k = 0; k < 2; ++k) {
                const double shift = t[2 + k] / t[k];                          // mu_new - mu_old        (GMMFit.cpp:75-79)
                st.denom[k] = t[k];
                st.mu_new[k] = st.mu[k] + shift;
                st.sigma[k] = sqrt(t[4 + k] / t[k] - shift * shift);           // :85-92
                st.mu[k] = st.mu_new[k];
                st.prior[k] = t[k] / double(n);                                // :98-101
            }
            st.counter = 0;
        }
    }
}

// The shape for the clouds of the calibration chain (tens of thousands of points): the WHOLE fit in one launch on a thread-block
// cluster. A thread keeps its points — and, up to kGmmCache of them, the posteriors of the current iteration — in registers, so the
// cloud is read once and the second pass of an iteration (the variance about the new mean, GMMFit.cpp:79-85) costs no exp and no
// division. The partial sums of a pass are exchanged through distributed shared memory and added in rank order by every CTA:
// all CTAs hold the same state, nothing has to be published. Four cluster barriers per EM iteration.
constexpr int kGmmCluster = 8;
constexpr int kGmmCache = 24;
constexpr int64_t kGmmClusterMaxPts = 200000;

__device__ __forceinline__ void gmm_cluster_sum(cg::cluster_group& cluster, double (&a)[4], double (*s_red)[4], double* s_part, double* s_tot)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = warp_sum(a[q]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int q = 0; q < 4; ++q) s_red[warp][q] = a[q];
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < kGmmThreads / 32; ++w) t += s_red[w][threadIdx.x];
        s_part[threadIdx.x] = t;
    }
    cluster.sync();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int r = 0; r < kGmmCluster; ++r) t += cluster.map_shared_rank(s_part, r)[threadIdx.x];
        s_tot[threadIdx.x] = t;
    }
    cluster.sync();                                   // the partial sums are read: the next pass may overwrite them (and s_tot is visible)
}

template <bool CACHE>
__global__ void __cluster_dims__(kGmmCluster, 1, 1) __launch_bounds__(kGmmThreads)
k_gmm_fit_cluster(const float* __restrict__ x, int elem_stride, int64_t n, int iters, double mu0, double mu1, double sg0, double sg1, GmmState* __restrict__ out)
{
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_par[12], s_red[kGmmThreads / 32][4], s_part[4], s_tot[4];
    __shared__ GmmState s_st;
    if (threadIdx.x == 0) {
        s_st.mu[0] = mu0; s_st.mu[1] = mu1; s_st.sigma[0] = sg0; s_st.sigma[1] = sg1; s_st.prior[0] = 0.5; s_st.prior[1] = 0.5;
        s_st.mu_new[0] = s_st.mu_new[1] = s_st.denom[0] = s_st.denom[1] = 0.0; s_st.counter = 0; s_st.pad = 0;
    }
    const int64_t stride = (int64_t)kGmmCluster * kGmmThreads, first = (int64_t)cluster.block_rank() * kGmmThreads + threadIdx.x;
    const int mine = first < n ? (int)((n - first + stride - 1) / stride) : 0;
    float xs[CACHE ? kGmmCache : 1];
    double p0s[CACHE ? kGmmCache : 1], p1s[CACHE ? kGmmCache : 1];
    if (CACHE) {
#pragma unroll
        for (int k = 0; k < kGmmCache; ++k) xs[k] = k < mine ? __ldg(x + (first + k * stride) * elem_stride) : 0.f;
    }
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x < 2) {
            const int k = threadIdx.x;
            const double sg = s_st.sigma[k];
            s_par[k] = s_st.mu[k];
            s_par[2 + k] = __ddiv_rn(-1.0, __dmul_rn(2.0, __dmul_rn(sg, sg)));                 // -1.0 / (2*sigma_sqr)
            s_par[4 + k] = __ddiv_rn(1.0, __dmul_rn(sg, sqrt(2.0 * 3.14159265358979323846)));  // 1/(sigma*sqrt(2 pi))
            s_par[6 + k] = s_st.prior[k];
        }
        __syncthreads();
        double a[4] = {0, 0, 0, 0};
        if (CACHE) {
#pragma unroll
            for (int k = 0; k < kGmmCache; ++k)
                if (k < mine) {
                    const double xi = (double)xs[k];
                    posterior(xi, s_par, s_par + 2, s_par + 4, s_par + 6, p0s[k], p1s[k]);
                    a[0] += p0s[k]; a[1] += p1s[k];
                    a[2] += __dmul_rn(p0s[k], xi); a[3] += __dmul_rn(p1s[k], xi);
                }
        } else {
            for (int64_t i = first; i < n; i += stride) {
                const double xi = (double)__ldg(x + i * elem_stride);
                double p0, p1;
                posterior(xi, s_par, s_par + 2, s_par + 4, s_par + 6, p0, p1);
                a[0] += p0; a[1] += p1;
                a[2] += __dmul_rn(p0, xi); a[3] += __dmul_rn(p1, xi);
            }
        }
        gmm_cluster_sum(cluster, a, s_red, s_part, s_tot);
        if (threadIdx.x < 2) { s_st.denom[threadIdx.x] = s_tot[threadIdx.x]; s_st.mu_new[threadIdx.x] = s_tot[2 + threadIdx.x] / s_tot[threadIdx.x]; }      // GMMFit.cpp:75-79
        __syncthreads();
        const double m0 = s_st.mu_new[0], m1 = s_st.mu_new[1];
        a[0] = a[1] = a[2] = a[3] = 0;
        if (CACHE) {
#pragma unroll
            for (int k = 0; k < kGmmCache; ++k)
                if (k < mine) {
                    const double xi = (double)xs[k];
                    const double e0 = __dsub_rn(xi, m0), e1 = __dsub_rn(xi, m1);
                    a[0] += __dmul_rn(p0s[k], __dmul_rn(e0, e0));
                    a[1] += __dmul_rn(p1s[k], __dmul_rn(e1, e1));
                }
        } else {
            for (int64_t i = first; i < n; i += stride) {
                const double xi = (double)__ldg(x + i * elem_stride);
                double p0, p1;
                posterior(xi, s_par, s_par + 2, s_par + 4, s_par + 6, p0, p1);
                const double e0 = __dsub_rn(xi, m0), e1 = __dsub_rn(xi, m1);
                a[0] += __dmul_rn(p0, __dmul_rn(e0, e0));
                a[1] += __dmul_rn(p1, __dmul_rn(e1, e1));
            }
        }
        gmm_cluster_sum(cluster, a, s_red, s_part, s_tot);
        if (threadIdx.x < 2) {
            const int k = threadIdx.x;
            s_st.mu[k] = s_st.mu_new[k];
            s_st.sigma[k] = sqrt(s_tot[k] / s_st.denom[k]);                                      // :85-92
            s_st.prior[k] = s_st.denom[k] / double(n);                                           // :98-101
        }
    }
    __syncthreads();
    if (cluster.block_rank() == 0 && threadIdx.x == 0) *out = s_st;
}

// intensity plane of the resident float4 clouds: the 20 EM passes then read 4 B per point from a 40 MB array that stays in L2,
// instead of striding through 16 B records in DRAM
__global__ void __launch_bounds__(256) k_pack_intensity(const float4* __restrict__ cloud, int64_t frame_stride, const int64_t* __restrict__ n_pts,
                                                        float* __restrict__ out, int64_t out_stride)
{
    const int f = blockIdx.y;
    const int64_t n = n_pts[f];
    const float4* src = cloud + (size_t)f * frame_stride;
    float* dst = out + (size_t)f * out_stride;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) dst[i] = __ldg(src + i).w;
}

static int gmm_run(rclc_ctx* ctx, GmmView v, int F, int64_t max_n, int iters)
{
    const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>((max_n + kGmmThreads * 8 - 1) / (kGmmThreads * 8), kGmmMaxCtas));
    for (int it = 0; it < iters; ++it) {
        k_gmm_iter<<<dim3(ctas, F), kGmmThreads, 0, ctx->stream>>>(v);
        ctx->launches += 1;
    }
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

static int gmm_solve(rclc_ctx* ctx, const float* d_x, int64_t frame_stride, int elem_stride, const int64_t* d_npts, int F,
                     int64_t max_n, int iters, double* mu_io, double* sigma_io, double* priors_out)
{
    const size_t st_bytes = sizeof(GmmState) * F, part_bytes = (size_t)F * kGmmMaxCtas * 6 * 8;
    RCLC_TRY(ctx->d_out.reserve(st_bytes + part_bytes + 256));
    std::vector<GmmState> h(F);
    for (int f = 0; f < F; ++f) {
        std::memset(&h[f], 0, sizeof(GmmState));
        for (int k = 0; k < 2; ++k) { h[f].mu[k] = mu_io[2 * f + k]; h[f].sigma[k] = sigma_io[2 * f + k]; h[f].prior[k] = 0.5; }
    }
    GmmView v;
    v.x = d_x; v.frame_stride = frame_stride; v.elem_stride = elem_stride; v.n_pts = d_npts;
    v.st = ctx->d_out.as<GmmState>();
    v.partials = reinterpret_cast<double*>(static_cast<char*>(ctx->d_out.p) + ((st_bytes + 255) & ~(size_t)255));
    RCLC_CUDA(cudaMemcpyAsync(v.st, h.data(), st_bytes, cudaMemcpyHostToDevice, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));          // h is pageable
    RCLC_TRY(gmm_run(ctx, v, F, max_n, iters));
    RCLC_CUDA(cudaMemcpyAsync(h.data(), v.st, st_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 2; ++k) {
            mu_io[2 * f + k] = h[f].mu[k];
            sigma_io[2 * f + k] = h[f].sigma[k];
            if (priors_out) priors_out[2 * f + k] = h[f].prior[k];
        }
    return RCLC_OK;
}

int dev_gmm_fit(rclc_ctx* ctx, const float* d_x, int elem_stride, int64_t n, int iters, double* mu_io, double* sigma_io, double* priors_out)
{
    if (n > 0 && n <= kGmmClusterMaxPts) {
        RCLC_TRY(ctx->d_out.reserve(sizeof(GmmState) + 256));
        RCLC_TRY(ctx->h_pin2.reserve(sizeof(GmmState) + 64));
        GmmState* d_st = ctx->d_out.as<GmmState>();
        GmmState* h = ctx->h_pin2.as<GmmState>();
        if (n <= (int64_t)kGmmCache * kGmmCluster * kGmmThreads)
            k_gmm_fit_cluster<true><<<kGmmCluster, kGmmThreads, 0, ctx->stream>>>(d_x, elem_stride, n, iters, mu_io[0], mu_io[1], sigma_io[0], sigma_io[1], d_st);
        else
            k_gmm_fit_cluster<false><<<kGmmCluster, kGmmThreads, 0, ctx->stream>>>(d_x, elem_stride, n, iters, mu_io[0], mu_io[1], sigma_io[0], sigma_io[1], d_st);
        ctx->launches += 1;
        RCLC_CUDA(cudaGetLastError());
        RCLC_CUDA(cudaMemcpyAsync(h, d_st, sizeof(GmmState), cudaMemcpyDeviceToHost, ctx->stream));
        RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 2; ++k) { mu_io[k] = h->mu[k]; sigma_io[k] = h->sigma[k]; if (priors_out) priors_out[k] = h->prior[k]; }
        return RCLC_OK;
    }
    // the point count lives next to the solver state (gmm_solve reserves d_out first; the count goes behind everything it uses)
    const size_t o_n = sizeof(GmmState) + 256 + (size_t)kGmmMaxCtas * 6 * 8 + 512;
    RCLC_TRY(ctx->d_out.reserve(o_n + 64));
    RCLC_TRY(ctx->h_pin2.reserve(64));
    int64_t* h_n = ctx->h_pin2.as<int64_t>();
    *h_n = n;
    int64_t* d_n = reinterpret_cast<int64_t*>(ctx->d_out.as<char>() + o_n);
    RCLC_CUDA(cudaMemcpyAsync(d_n, h_n, 8, cudaMemcpyHostToDevice, ctx->stream));
    return gmm_solve(ctx, d_x, 0, elem_stride, d_n, 1, n, iters, mu_io, sigma_io, priors_out);
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_gmm_fit_1d(rclc_ctx* ctx, const void* intensity, int stride, int64_t n, int K, int iters,
                    double* mu_io, double* sigma_io, double* priors_out)
{
    RCLC_CHECK(ctx && intensity && mu_io && sigma_io && n > 0 && stride >= 4, RCLC_ERR_INVALID, "bad argument");
    RCLC_CHECK(K == 2, RCLC_ERR_UNSUPPORTED, "only K == 2 is implemented (every call site in the reference uses 2)");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->h_pin.reserve((size_t)n * 4 + 64));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 4 + 64));
    float* hp = ctx->h_pin.as<float>();
    const char* src = static_cast<const char*>(intensity);
    if (stride == 4) std::memcpy(hp, src, (size_t)n * 4);
    else for (int64_t i = 0; i < n; ++i) std::memcpy(hp + i, src + i * (int64_t)stride, 4);
    RCLC_CUDA(cudaMemcpyAsync(ctx->d_pts.p, hp, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    return dev_gmm_fit(ctx, ctx->d_pts.as<float>(), 1, n, iters, mu_io, sigma_io, priors_out);
}

int rclc_batch_gmm_fit(rclc_batch* b, double* mu_io, double* sigma_io, double* priors_out)
{
    RCLC_CHECK(b && mu_io && sigma_io, RCLC_ERR_INVALID, "null argument");
    RCLC_CUDA(cudaSetDevice(b->ctx->device));
    RCLC_TRY(batch_join_uploads(b));
    int64_t max_n = 0;
    for (int64_t n : b->h_npts) max_n = std::max(max_n, n);
    if (max_n == 0) return gmm_solve(b->ctx, reinterpret_cast<const float*>(b->d_pristine) + 3, b->v.frame_stride * 4, 4, b->d_npts, b->n_frames,
                                     max_n, b->cfg.gmm_iters, mu_io, sigma_io, priors_out);
    // intensity is the .w lane of the resident float4 cloud (pc_process.h:603-606 copies it out into an Eigen vector; so does this)
    rclc_ctx* ctx = b->ctx;
    const int64_t out_stride = (max_n + 63) & ~(int64_t)63;
    RCLC_TRY(ctx->d_pts.reserve((size_t)b->n_frames * (size_t)out_stride * 4));
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((max_n + 256 * 8 - 1) / (256 * 8), 1184));
    k_pack_intensity<<<dim3((unsigned)gx, (unsigned)b->n_frames), 256, 0, ctx->stream>>>(b->d_pristine, b->v.frame_stride, b->d_npts, ctx->d_pts.as<float>(),
                                                                                            out_stride);
    ctx->launches += 1;
    return gmm_solve(ctx, ctx->d_pts.as<float>(), out_stride, 1, b->d_npts, b->n_frames, max_n, b->cfg.gmm_iters, mu_io, sigma_io, priors_out);
}

} // extern "C"
