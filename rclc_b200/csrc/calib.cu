// calib.cu — the calibration's per-frame chains on clouds that stay in HBM, and the batched calibration over many frames.
//
//   seg_board_dev          the per-cloud body of apps/BoardSegmentation.cpp:45-165 (stage 1)
//   board_frame_dev        include/pc_process.h:583-621: est_board_corner up to the cloud in its board frame (stage 2, per frame)
//   rclc_calibrate_frames  stage 1 + stage 2 (apps/Calibrate.cpp:90-107) + calib_opt (apps/Calibrate.cpp:12-67) for F raw scans
//
// A scan is uploaded ONCE. Every operator of the chain is the resident-cloud function of ops_dev.cuh; between them only control
// scalars cross the bus (counts, the winning hypothesis, a plane, 24 block centroids). The reference's control flow is sequential
// per frame and data dependent (the RANSAC peeling loop, the reflectivity rounds of eular_cluster_iter), so a frame's chain is a
// latency chain of ~100 short launches and ~100 scalar read-backs: frames are therefore run CONCURRENTLY, one host worker with
// its own stream and scratch per frame in flight (the reference's `#pragma omp parallel for` over frames,
// BoardSegmentation.cpp:31-33, Calibrate.cpp:90-92), and the grid fit — the one heavy stage — runs for all frames together
// through the batched solve (gridfit.cu).
#include "gridfit.cuh"
#include "ops_dev.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <functional>
#include <mutex>
#include <random>
#include <thread>

namespace rclc {

#ifdef RCLC_LAB
// lab build only (-DRCLC_LAB): wall time per stage of the chains, summed over the frames of a call (scripts/chain_timing.py)
enum LabStage { LS_UPLOAD, LS_ROI_US, LS_PEEL, LS_SEG_CLUSTERS, LS_PLANE, LS_PROJECT_GMM, LS_BLOCKS, LS_FRAME, LS_COUNT };
static std::atomic<long long> g_lab_ns[LS_COUNT];
struct LabTimer {
    int stage; std::chrono::steady_clock::time_point t0;
    explicit LabTimer(int s) : stage(s), t0(std::chrono::steady_clock::now()) {}
    void next(int s) { const auto t = std::chrono::steady_clock::now(); g_lab_ns[stage] += std::chrono::duration_cast<std::chrono::nanoseconds>(t - t0).count(); stage = s; t0 = t; }
    ~LabTimer() { next(stage); }
};
#define LAB_TIMER(var, s) LabTimer var(s)
#define LAB_NEXT(var, s) var.next(s)
#else
#define LAB_TIMER(var, s)
#define LAB_NEXT(var, s)
#endif

// ---------------------------------------------------------------------------------------------------------------------------
// small device pieces
// ---------------------------------------------------------------------------------------------------------------------------

// pcl::PassThrough (inclusive limits, finite points only) on one field, or the conjunction over x, y, z of cut_roi
// (pc_process.h:81-101: three PassThrough filters in a row keep exactly the points inside the box, in their order).
struct RangePred {
    float lo[4], hi[4];
    int use;                                    // bit d: field d (x, y, z, intensity) is limited to [lo[d], hi[d]]
    __device__ __forceinline__ bool operator()(const float4& p) const
    {
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;
        const float v[4] = {p.x, p.y, p.z, p.w};
        bool in = true;
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if ((use >> d) & 1) in = in && isfinite(v[d]) && v[d] >= lo[d] && v[d] <= hi[d];
        return in;
    }
};
static RangePred field_range(int field, float lo, float hi)
{
    RangePred r;
    for (int d = 0; d < 4; ++d) { r.lo[d] = 0.f; r.hi[d] = 0.f; }
    r.lo[field] = lo; r.hi[field] = hi; r.use = 1 << field;
    return r;
}
struct IsSet { __device__ __forceinline__ bool operator()(uint8_t m) const { return m != 0; } };
struct IsClear { __device__ __forceinline__ bool operator()(uint8_t m) const { return m == 0; } };
struct IsFirst { __device__ __forceinline__ bool operator()(int label) const { return label == 0; } };

// pcl::compute3DCentroid into a Vector4f: x, y, z are summed in FLOAT, point after point. Every addition rounds, so the sum is
// the reference's only in the reference's order: one thread adds, the CTA stages the tiles.
constexpr int kSeqTile = 2048;
__global__ void __launch_bounds__(256) k_seq_sum3(const float4* __restrict__ pts, int n, float* __restrict__ out)
{
    __shared__ float sx[kSeqTile], sy[kSeqTile], sz[kSeqTile];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int base = 0; base < n; base += kSeqTile) {
        for (int k = threadIdx.x; k < kSeqTile; k += 256) {
            const int i = base + k;
            const float4 p = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            sx[k] = p.x; sy[k] = p.y; sz[k] = p.z;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = min(kSeqTile, n - base);
#pragma unroll 8
            for (int k = 0; k < m; ++k) { ax = __fadd_rn(ax, sx[k]); ay = __fadd_rn(ay, sy[k]); az = __fadd_rn(az, sz[k]); }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = ax; out[1] = ay; out[2] = az; }
}

// the same sums for every cluster of a labelled cloud at once: thread s owns cluster s and walks the cloud in point order
__global__ void __launch_bounds__(256) k_cluster_sum3(const float4* __restrict__ pts, const int* __restrict__ label, int n, int nc, float* __restrict__ out)
{
    __shared__ float4 sp[256];
    __shared__ int sl[256];
    const int s = blockIdx.x * 256 + threadIdx.x;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        sl[threadIdx.x] = i < n ? label[i] : -1;
        sp[threadIdx.x] = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const int m = min(256, n - base);
        for (int k = 0; k < m; ++k)
            if (sl[k] == s) { const float4 p = sp[k]; ax = __fadd_rn(ax, p.x); ay = __fadd_rn(ay, p.y); az = __fadd_rn(az, p.z); }
        __syncthreads();
    }
    if (s < nc) { out[3 * s] = ax; out[3 * s + 1] = ay; out[3 * s + 2] = az; }
}

// pcl::getMinMax3D: float min / max are exact in any order (ordered-integer atomics)
__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
static float o2f(int i) { const int v = i >= 0 ? i : i ^ 0x7fffffff; float f; std::memcpy(&f, &v, 4); return f; }
__global__ void k_minmax_init(int* bb) { if (threadIdx.x < 3) bb[threadIdx.x] = INT_MAX; else if (threadIdx.x < 6) bb[threadIdx.x] = INT_MIN; }
__global__ void __launch_bounds__(256) k_minmax(const float4* __restrict__ pts, int64_t n, int* __restrict__ bb)
{
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const float4 p = __ldg(pts + i);
        const int o[3] = {f2o(p.x), f2o(p.y), f2o(p.z)};
#pragma unroll
        for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], o[d]); hi[d] = max(hi[d], o[d]); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
    }
    if ((threadIdx.x & 31) == 0)
        for (int d = 0; d < 3; ++d) { atomicMin(bb + d, lo[d]); atomicMax(bb + 3 + d, hi[d]); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// chain scratch of one context: clouds that live across operators (the operators' own scratch is d_tmp / d_out / h_pin2)
// ---------------------------------------------------------------------------------------------------------------------------
enum ChainBuf { CB_SCAN = 0, CB_ROI, CB_ROI2, CB_VOX_A, CB_VOX_B, CB_EXT, CB_TARGET, CB_TC, CB_BOARD, CB_COUNT };
static_assert(CB_COUNT <= 10, "rclc_ctx::d_chain");

static float4* chain_buf(rclc_ctx* ctx, int which, int64_t n, int* rc)
{
    *rc = ctx->d_chain[which].reserve((size_t)std::max<int64_t>(n, 1) * 16);
    return ctx->d_chain[which].as<float4>();
}
#define CHAIN_BUF(var, which, n)                                  \
    float4* var;                                                   \
    { int cb_rc_ = RCLC_OK; var = chain_buf(ctx, which, n, &cb_rc_); if (cb_rc_ != RCLC_OK) return cb_rc_; }

// order-preserving compaction; *m == nullptr: the caller knows the count, nothing is read back
template <class InIt, class FlagIt>
static int compact_flagged(rclc_ctx* ctx, InIt src, FlagIt flags, int64_t n, float4* dst, int64_t* m)
{
    if (n == 0) { if (m) *m = 0; return RCLC_OK; }
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, src, flags, dst, (int*)nullptr, (int)n, ctx->stream);
    RCLC_TRY(ctx->d_tmp2.reserve(256 + bytes));
    int* d_cnt = ctx->d_tmp2.as<int>();
    RCLC_CUDA(cub::DeviceSelect::Flagged(ctx->d_tmp2.as<char>() + 256, bytes, src, flags, dst, d_cnt, (int)n, ctx->stream));
    ctx->launches += 1;
    if (m) {
        RCLC_TRY(ctx->h_chain.reserve(256));
        RCLC_CUDA(cudaMemcpyAsync(ctx->h_chain.p, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
        RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
        *m = *ctx->h_chain.as<int>();
    }
    return RCLC_OK;
}
static int compact_range(rclc_ctx* ctx, const float4* src, int64_t n, const RangePred& pred, float4* dst, int64_t* m)
{
    if (n == 0) { *m = 0; return RCLC_OK; }
    size_t bytes = 0;
    cub::DeviceSelect::If(nullptr, bytes, src, dst, (int*)nullptr, (int)n, pred, ctx->stream);
    RCLC_TRY(ctx->d_tmp2.reserve(256 + bytes));
    int* d_cnt = ctx->d_tmp2.as<int>();
    RCLC_CUDA(cub::DeviceSelect::If(ctx->d_tmp2.as<char>() + 256, bytes, src, dst, d_cnt, (int)n, pred, ctx->stream));
    ctx->launches += 1;
    RCLC_TRY(ctx->h_chain.reserve(256));
    RCLC_CUDA(cudaMemcpyAsync(ctx->h_chain.p, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    *m = *ctx->h_chain.as<int>();
    return RCLC_OK;
}

// EulerCluster (pc_process.h:59-79): the largest cluster (tolerance 3 cm, 500 .. 2.5 M points) of src, in src's order
static int largest_cluster(rclc_ctx* ctx, const float4* src, int64_t n, float4* dst, int64_t* m, const float* bbox = nullptr)
{
    *m = 0;
    if (n == 0) return RCLC_OK;
    ClusterSet cs;
    if (bbox) { cs.have_bbox = true; for (int d = 0; d < 6; ++d) cs.bbox[d] = bbox[d]; }
    RCLC_TRY(dev_euclidean_cluster(ctx, src, n, 0.03, 500, 2500000, &cs));
    if (cs.sizes.empty()) return RCLC_OK;
    auto first = thrust::make_transform_iterator(cs.d_label, IsFirst());
    RCLC_TRY(compact_flagged(ctx, src, first, n, dst, nullptr));
    *m = cs.sizes[0];
    return RCLC_OK;
}

// PCL SampleConsensusModel::drawIndexSample: mt19937 seeded 12345 per model object, partial Fisher-Yates on a persistent index
// list. The hypotheses are an input of the parity definition; this is the list the reference would draw.
static void draw_triplets(int64_t n, int H, std::vector<int32_t>& idx, std::vector<int32_t>& tri)
{
    std::mt19937 rng(12345u);
    idx.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) idx[(size_t)i] = (int32_t)i;
    tri.resize((size_t)H * 3);
    for (int h = 0; h < H; ++h) {
        for (int i = 0; i < 3; ++i) { const uint32_t r = rng() >> 1; std::swap(idx[(size_t)i], idx[(size_t)(i + (r % (uint64_t)(n - i)))]); }
        for (int i = 0; i < 3; ++i) tri[(size_t)h * 3 + i] = idx[(size_t)i];
    }
}
// pcl::RandomSampleConsensus::computeModel replayed over the per-hypothesis inlier counts: the winner, or -1
static int ransac_replay(const std::vector<int32_t>& counts, const std::vector<int32_t>& valid, int64_t n, int max_iterations, double probability)
{
    const int H = (int)counts.size();
    int best = -INT_MAX, winner = -1, iter = 0, skipped = 0, h = 0;
    double k = 1.0;
    const double log_probability = std::log(1.0 - probability);
    const int max_skip = max_iterations * 10;
    while (iter < k && skipped < max_skip && h < H) {
        const int cur = h++;
        if (!valid[(size_t)cur]) { ++skipped; continue; }
        if (counts[(size_t)cur] > best) {
            best = counts[(size_t)cur]; winner = cur;
            const double w = (double)best / (double)n;
            double p_no_outliers = 1.0 - w * w * w;
            p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
            p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
            k = log_probability / std::log(p_no_outliers);
        }
        ++iter;
        if (iter > max_iterations) break;
    }
    return winner;
}

// ---------------------------------------------------------------------------------------------------------------------------
// stage 1: the chessboard's points out of one raw scan (lidar frame, x forward)
// ---------------------------------------------------------------------------------------------------------------------------
// d_scan [n] -> *d_board [*m] (ctx scratch: white cluster, then black cluster, BoardSegmentation.cpp:160). *m == 0: no plane
// qualifies (the reference goes on with an empty cloud).
int seg_board_dev(rclc_ctx* ctx, const rclc_config* cfg, const float4* d_scan, int64_t n, double x_min, double x_max, const float4** d_board, int64_t* m)
{
    *m = 0; *d_board = nullptr;
    LAB_TIMER(lab, LS_ROI_US);
    CHAIN_BUF(d_roi, CB_ROI, n);
    int64_t n_roi = 0;
    RCLC_TRY(compact_range(ctx, d_scan, n, field_range(0, (float)x_min, (float)x_max), d_roi, &n_roi));        // :48
    if (n_roi == 0) return RCLC_OK;
    // 8 mm uniform sampling (:58-61)
    CHAIN_BUF(d_vox_a, CB_VOX_A, n_roi);
    CHAIN_BUF(d_vox_b, CB_VOX_B, n_roi);
    CHAIN_BUF(d_ext, CB_EXT, n_roi);
    CHAIN_BUF(d_target, CB_TARGET, n_roi);
    const int* d_idx = nullptr;
    int64_t n_vox = 0;
    RCLC_TRY(dev_uniform_sample(ctx, d_roi, n_roi, (double)0.008f, &d_idx, &n_vox));
    RCLC_TRY(dev_gather(ctx, d_roi, d_idx, n_vox, d_vox_a));
    const size_t pt_size = (size_t)n_vox;
    LAB_NEXT(lab, LS_PEEL);
    float4 *vox = d_vox_a, *vox_next = d_vox_b;
    // planes are peeled off until 5 % of the points are left (:81-113); the nearest large plane that faces the lidar is the candidate
    const int max_iterations = 800, H = max_iterations + 1;
    const double thr = 0.08;
    std::vector<int32_t> idx, tri, valid((size_t)H), counts((size_t)H);
    std::vector<float> planes((size_t)H * 4);
    int64_t n_target = 0;
    double closest = 1000;
    RCLC_TRY(ctx->h_chain.reserve(256));
    float* h_sum = ctx->h_chain.as<float>() + 16;
    float* d_sum = nullptr;
    while ((double)n_vox > pt_size * 0.05) {
        if (n_vox < 3) break;
        draw_triplets(n_vox, H, idx, tri);
        RCLC_TRY(dev_ransac_score(ctx, vox, n_vox, tri.data(), H, thr, planes.data(), valid.data(), counts.data()));
        const int best = ransac_replay(counts, valid, n_vox, max_iterations, 0.99);
        if (best < 0) break;
        const uint8_t* d_mask = nullptr;
        float refined[4], centroid[3];
        int64_t n_in = 0;
        RCLC_TRY(dev_ransac_select(ctx, vox, n_vox, &planes[(size_t)best * 4], thr, &d_mask, refined, centroid, &n_in));
        RCLC_TRY(dev_ransac_select(ctx, vox, n_vox, refined, thr, &d_mask, nullptr, nullptr, &n_in));      // SACSegmentation re-selects with the optimised model
        if (n_in == 0) break;
        auto in_it = thrust::make_transform_iterator(d_mask, IsSet());
        auto out_it = thrust::make_transform_iterator(d_mask, IsClear());
        RCLC_TRY(compact_flagged(ctx, vox, in_it, n_vox, d_ext, nullptr));
        RCLC_TRY(compact_flagged(ctx, vox, out_it, n_vox, vox_next, nullptr));
        std::swap(vox, vox_next);
        n_vox -= n_in;
        const double a = refined[0], b = refined[1], c = refined[2];
        const double nx = a / std::sqrt(a * a + b * b + c * c);
        if ((double)n_in < pt_size * 0.02 || std::fabs(nx) <= cfg->angle_thd) continue;                       // :101
        if (!d_sum) { RCLC_TRY(ctx->d_tmp2.reserve(4096)); }
        d_sum = reinterpret_cast<float*>(ctx->d_tmp2.as<char>() + 64);                                          // (the compaction's counter lives in the first word)
        k_seq_sum3<<<1, 256, 0, ctx->stream>>>(d_ext, (int)n_in, d_sum);
        ctx->launches += 1;
        RCLC_CUDA(cudaMemcpyAsync(h_sum, d_sum, 12, cudaMemcpyDeviceToHost, ctx->stream));
        RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
        const float inv = 1.0f / (float)n_in;
        const float cx = h_sum[0] * inv, cy = h_sum[1] * inv, cz = h_sum[2] * inv;
        const double dist = std::sqrt((double)cx * cx + (double)cy * cy + (double)cz * cz);
        if ((closest - dist) > 0.3) {                                                                           // :107
            closest = dist;
            RCLC_CUDA(cudaMemcpyAsync(d_target, d_ext, (size_t)n_in * 16, cudaMemcpyDeviceToDevice, ctx->stream));
            n_target = n_in;
        }
    }
    if (n_target == 0) return RCLC_OK;
    LAB_NEXT(lab, LS_SEG_CLUSTERS);
    // the candidate's largest cluster gives the ROI box of the raw scan (:116-128)
    CHAIN_BUF(d_tc, CB_TC, n_target);
    int64_t n_tc = 0;
    RCLC_TRY(largest_cluster(ctx, d_target, n_target, d_tc, &n_tc));
    if (n_tc == 0) return RCLC_OK;
    RCLC_TRY(ctx->d_tmp2.reserve(4096));
    int* d_bb = reinterpret_cast<int*>(ctx->d_tmp2.as<char>() + 128);
    int* h_bb = ctx->h_chain.as<int>() + 32;
    k_minmax_init<<<1, 32, 0, ctx->stream>>>(d_bb);
    k_minmax<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n_tc + 255) / 256, ctx->sm_count * 4)), 256, 0, ctx->stream>>>(d_tc, n_tc, d_bb);
    ctx->launches += 2;
    RCLC_CUDA(cudaMemcpyAsync(h_bb, d_bb, 24, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    RangePred box;
    box.use = 7;
    for (int d = 0; d < 3; ++d) { box.lo[d] = o2f(h_bb[d]); box.hi[d] = o2f(h_bb[3 + d]); }
    const float roi_box[6] = {box.lo[0], box.lo[1], box.lo[2], box.hi[0], box.hi[1], box.hi[2]};      // holds every point of the cut scan
    box.lo[3] = box.hi[3] = 0.f;
    CHAIN_BUF(d_roi2, CB_ROI2, n_roi);
    int64_t n_roi2 = 0;
    RCLC_TRY(compact_range(ctx, d_roi, n_roi, box, d_roi2, &n_roi2));
    // a GMM on the candidate's reflectance splits black from white (:130-158); each side keeps its largest cluster
    double mu[2] = {15, 100}, sigma[2] = {15, 15}, pri[2];
    RCLC_TRY(dev_gmm_fit(ctx, reinterpret_cast<const float*>(d_tc) + 3, 4, n_tc, cfg->gmm_iters, mu, sigma, pri));
    const double thd = mu[0] + 1.7 * sigma[0];                                                                  // :134
    CHAIN_BUF(d_board_buf, CB_BOARD, n_roi2);
    float4* d_side = d_vox_a;                                                                                   // (the peeling buffers are free again)
    int64_t n_side = 0, n_white = 0, n_black = 0;
    RCLC_TRY(compact_range(ctx, d_roi2, n_roi2, field_range(3, (float)thd, 150.f), d_side, &n_side));
    RCLC_TRY(largest_cluster(ctx, d_side, n_side, d_board_buf, &n_white, roi_box));
    RCLC_TRY(compact_range(ctx, d_roi2, n_roi2, field_range(3, (float)cfg->noise_reflactivity_thd, (float)thd), d_side, &n_side));
    RCLC_TRY(largest_cluster(ctx, d_side, n_side, d_board_buf + n_white, &n_black, roi_box));
    *m = n_white + n_black;                                                                                     // *final_plane = *white + *black (:160)
    *d_board = d_board_buf;
    return RCLC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// stage 2, per frame: est_board_corner up to the cloud in its (approximate) board frame
// ---------------------------------------------------------------------------------------------------------------------------
// eular_cluster_iter (pc_process.h:104-211) on a resident cloud: centroids of the board's black blocks
static int block_centroids_dev(rclc_ctx* ctx, const rclc_config* cfg, const float4* d_cloud, int64_t n, double reflactive_thd, std::vector<double>& centroids,
                               int max_rounds)
{
    centroids.clear();
    CHAIN_BUF(d_a, CB_VOX_A, n);
    CHAIN_BUF(d_b, CB_VOX_B, n);
    const int* d_idx = nullptr;
    int64_t nb = 0;
    RCLC_TRY(dev_uniform_sample(ctx, d_cloud, n, cfg->block_us_radius, &d_idx, &nb));
    RCLC_TRY(dev_gather(ctx, d_cloud, d_idx, nb, d_a));
    float4 *blocks = d_a, *next = d_b;
    const int block_num = cfg->board_width * cfg->board_height / 2;
    const double diag = cfg->gride_side_length * std::sqrt(2.0f);                                               // src/global.cpp:47
    double curr = reflactive_thd / cfg->reflactivity_dec_step;
    std::vector<size_t> block_size;
    std::vector<int> block_num_pts;
    std::vector<float> sums;
    bool have_box = false;
    float blocks_box[6];                                 // of the first round: the later rounds cluster subsets of it
    for (int round = 0; round < max_rounds; ++round) {
        curr *= cfg->reflactivity_dec_step;
        int64_t kept = 0;
        RCLC_TRY(compact_range(ctx, blocks, nb, field_range(3, 3.0f, (float)curr), next, &kept));             // :127-130
        std::swap(blocks, next);
        nb = kept;
        if (nb == 0) return RCLC_OK;
        ClusterSet cs;
        if (have_box) { cs.have_bbox = true; for (int d = 0; d < 6; ++d) cs.bbox[d] = blocks_box[d]; }
        RCLC_TRY(dev_euclidean_cluster(ctx, blocks, nb, cfg->black_pt_cluster_dst_thd, 3, 25000, &cs));
        if (!have_box) { have_box = true; for (int d = 0; d < 6; ++d) blocks_box[d] = cs.bbox[d]; }
        const int nc = (int)cs.sizes.size();
        if (nc < block_num) continue;
        // pcl::compute3DCentroid of every cluster: float sums in point order
        RCLC_TRY(ctx->d_out.reserve((size_t)nc * 12 + 256));
        RCLC_TRY(ctx->h_pin2.reserve((size_t)nc * 12 + 256));
        float* d_sums = ctx->d_out.as<float>();
        k_cluster_sum3<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(blocks, cs.d_label, (int)nb, nc, d_sums);
        ctx->launches += 1;
        RCLC_CUDA(cudaMemcpyAsync(ctx->h_pin2.p, d_sums, (size_t)nc * 12, cudaMemcpyDeviceToHost, ctx->stream));
        RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
        sums.assign(ctx->h_pin2.as<float>(), ctx->h_pin2.as<float>() + (size_t)nc * 3);
        std::vector<double>& bc = centroids;
        for (int s = 0; s < nc; ++s) {
            const size_t sz = (size_t)cs.sizes[(size_t)s];
            const float inv = 1.0f / (float)sz;
            const double c[3] = {(double)(sums[3 * (size_t)s] * inv), (double)(sums[3 * (size_t)s + 1] * inv), (double)(sums[3 * (size_t)s + 2] * inv)};
            if (s < block_num) {
                if (s > 0 && (double(sz) / double(block_size[(size_t)s - 1])) < 0.65) break;                    // :160
                block_size.push_back(sz);
                bc.insert(bc.end(), c, c + 3);
                block_num_pts.push_back((int)sz);
            } else {
                int nearest = -1;
                double nearest_dst = 10e3;
                for (int i = 0; i < (int)(bc.size() / 3); ++i) {
                    const double dx = bc[3 * (size_t)i] - c[0], dy = bc[3 * (size_t)i + 1] - c[1], dz = bc[3 * (size_t)i + 2] - c[2];
                    const float dst = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
                    if (dst < nearest_dst) { nearest_dst = dst; nearest = i; }
                }
                if (nearest >= 0 && nearest_dst < 0.5 * diag) {                                                 // :180-195
                    double* old = &bc[3 * (size_t)nearest];
                    const double a = block_num_pts[(size_t)nearest], b = (double)sz;
                    for (int d = 0; d < 3; ++d) old[d] = (old[d] * a + c[d] * b) / (a + b);
                    block_num_pts[(size_t)nearest] += (int)sz;
                    block_size[(size_t)nearest] += sz;
                }
            }
        }
        if ((int)(bc.size() / 3) < block_num) { block_size.clear(); bc.clear(); block_num_pts.clear(); continue; }
        return RCLC_OK;
    }
    centroids.clear();
    return RCLC_OK;
}

// d_cloud [n] (the board-only cloud in the T_base'd laser frame) is projected onto its plane and moved into the board frame in
// place, like the reference's chess_board_processe; Tbl = board <- laser. *stage: 0 done, else the stage that gave up
// (1 plane fit, 2 block centroids, 3 board frame).
int board_frame_dev(rclc_ctx* ctx, const rclc_config* cfg, float4* d_cloud, int64_t n, double Tbl[16], int* stage)
{
    *stage = 0;
    if (cfg->board_order == 1) {                                                                                // pc_process.h:590-595
        const float T[16] = {-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        RCLC_TRY(dev_transform(ctx, d_cloud, n, T));
    }
    LAB_TIMER(lab, LS_PLANE);
    double plane[4];
    int32_t its = 0;
    int64_t used = 0;
    const int rc = dev_plane_fit(ctx, cfg, d_cloud, n, plane, &its, &used);                                     // :598-600
    if (rc == RCLC_ERR_NUMERIC) { *stage = 1; return RCLC_OK; }
    RCLC_TRY(rc);
    LAB_NEXT(lab, LS_PROJECT_GMM);
    RCLC_TRY(dev_plane_project(ctx, d_cloud, n, plane));                                                        // :601
    double mu[2] = {15, 100}, sigma[2] = {15, 15}, pri[2];
    RCLC_TRY(dev_gmm_fit(ctx, reinterpret_cast<const float*>(d_cloud) + 3, 4, n, cfg->gmm_iters, mu, sigma, pri));     // :604-611
    const double thd = mu[0] + 2.0 * sigma[0];                                                                  // :612
    LAB_NEXT(lab, LS_BLOCKS);
    std::vector<double> cen;
    RCLC_TRY(block_centroids_dev(ctx, cfg, d_cloud, n, thd, cen, 200));                                         // :614
    LAB_NEXT(lab, LS_FRAME);
    if ((int)(cen.size() / 3) < cfg->board_width * cfg->board_height / 2) { *stage = 2; return RCLC_OK; }
    if (rclc_calc_laser_coor(cfg, cen.data(), (int32_t)(cen.size() / 3), plane, Tbl) != RCLC_OK) { *stage = 3; return RCLC_OK; }      // :616
    float Tf[16];
    for (int k = 0; k < 16; ++k) Tf[k] = (float)Tbl[k];
    RCLC_TRY(dev_transform(ctx, d_cloud, n, Tf));                                                               // :617-618
    return RCLC_OK;
}

// est_board_corner's tail for board_order == 1 (pc_process.h:627-643): the rows run the other way
static void flip_corners(const rclc_config* cfg, double* corners)
{
    const int per_row = cfg->board_width - 1, rows = cfg->board_height - 1;
    std::vector<double> tmp((size_t)per_row * rows * 3);
    for (int row = 0; row < rows; ++row)
        for (int col = 0; col < per_row; ++col) {
            const double* p = corners + 3 * (row * per_row + col);
            double* q = tmp.data() + 3 * (row * per_row + per_row - col - 1);
            q[0] = -p[0]; q[1] = p[1]; q[2] = p[2];
        }
    std::memcpy(corners, tmp.data(), tmp.size() * 8);
}

// ---------------------------------------------------------------------------------------------------------------------------
// host workers: one per frame in flight, each with its own context (stream + scratch)
// ---------------------------------------------------------------------------------------------------------------------------
struct WorkerPool {
    std::vector<std::thread> threads;
    std::vector<rclc_ctx*> ctxs;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(int, rclc_ctx*)> job;     // (worker index, worker context)
    uint64_t generation = 0;
    int running = 0, active = 0;
    bool quit = false;

    void loop(int wi)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return quit || (generation != seen && wi < active); });
                if (quit) return;
                seen = generation;
            }
            cudaSetDevice(ctxs[(size_t)wi]->device);
            job(wi, ctxs[(size_t)wi]);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
    int ensure(int device, int n)
    {
        while ((int)ctxs.size() < n) {
            rclc_ctx* c = nullptr;
            RCLC_TRY(rclc_ctx_create(device, &c));
            ctxs.push_back(c);
            const int wi = (int)ctxs.size() - 1;
            threads.emplace_back([this, wi] { loop(wi); });
        }
        return RCLC_OK;
    }
    void run(int n, std::function<void(int, rclc_ctx*)> fn)
    {
        std::unique_lock<std::mutex> lk(mu);
        job = std::move(fn);
        active = n; running = n;
        ++generation;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return running == 0; });
        active = 0;
    }
    ~WorkerPool()
    {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv_go.notify_all();
        for (std::thread& t : threads) t.join();
        for (rclc_ctx* c : ctxs) rclc_ctx_destroy(c);
    }
};

void destroy_calib_state(rclc_ctx* ctx)
{
    if (ctx->calib_batch) { rclc_batch_destroy(ctx->calib_batch); ctx->calib_batch = nullptr; }
    if (ctx->calib_pool) { delete static_cast<WorkerPool*>(ctx->calib_pool); ctx->calib_pool = nullptr; }
}

static int default_workers(int F)
{
    static const int want = [] {
        if (const char* e = std::getenv("RCLC_CALIB_WORKERS")) return std::max(1, atoi(e));
        const char* lw = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
        return std::max(1, std::min(32, hw / ranks));
    }();
    return std::max(1, std::min(want, F));
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_ctx_set_calib_workers(rclc_ctx* ctx, int32_t workers)
{
    RCLC_CHECK(ctx && workers >= 0 && workers <= 64, RCLC_ERR_INVALID, "calibration workers must be 0 (automatic) .. 64");
    ctx->calib_workers = workers;
    return RCLC_OK;
}

int rclc_segment_board(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double x_min, double x_max, void* out_xyzi,
                       int64_t* n_out)
{
    RCLC_CHECK(ctx && cfg && pts && out_xyzi && n_out && n > 0 && n < (1ll << 31), RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    CHAIN_BUF(d_scan, CB_SCAN, n);
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, ioff, n, d_scan));
    const float4* d_board = nullptr;
    RCLC_TRY(seg_board_dev(ctx, cfg, d_scan, n, x_min, x_max, &d_board, n_out));
    if (*n_out > 0) RCLC_CUDA(cudaMemcpyAsync(out_xyzi, d_board, (size_t)*n_out * 16, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    return RCLC_OK;
}

int rclc_est_board_corner(rclc_ctx* ctx, const rclc_config* cfg, void* pts_io_xyzi, int64_t n, double* corners_out, double* T_bl_out,
                          rclc_gridfit_report* report, int32_t* stage_failed)
{
    RCLC_CHECK(ctx && cfg && pts_io_xyzi && corners_out && n > 0 && n < (1ll << 31), RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    if (stage_failed) *stage_failed = 0;
    CHAIN_BUF(d_cloud, CB_BOARD, n);
    RCLC_TRY(upload_cloud_f4(ctx, pts_io_xyzi, 16, 12, n, d_cloud));
    double Tbl[16];
    int stage = 0;
    RCLC_TRY(board_frame_dev(ctx, cfg, d_cloud, n, Tbl, &stage));
    if (stage == 0) {
        // opt_grid_fitting on the resident cloud (pc_process.h:620-621)
        if (!ctx->calib_batch || ctx->calib_batch->n_frames != 1 || ctx->calib_batch->max_pts < n || std::memcmp(&ctx->calib_batch->cfg, cfg, sizeof(*cfg)) != 0) {
            if (ctx->calib_batch) { rclc_batch_destroy(ctx->calib_batch); ctx->calib_batch = nullptr; }
            RCLC_TRY(rclc_batch_create(ctx, cfg, 1, n + n / 4 + 1024, &ctx->calib_batch));
        }
        RCLC_TRY(rclc_batch_set_frame_dev(ctx->calib_batch, 0, d_cloud, n));
        rclc_gridfit_report rep;
        RCLC_TRY(rclc_batch_gridfit_solve(ctx->calib_batch, Tbl, T_bl_out, corners_out, &rep));
        if (report) *report = rep;
        // A Ceres solve that ended in FAILURE (report.status 1: a NaN residual) is NOT a failed stage: opt_grid_fitting.h:96-131 never
        // looks at the summary and goes on to the corners of the pose it has. Only a frame without points or without an inverse is.
        if (rep.status >= 2) stage = 4;
        else if (cfg->board_order == 1) flip_corners(cfg, corners_out);
    }
    RCLC_CUDA(cudaMemcpyAsync(pts_io_xyzi, d_cloud, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (stage_failed) *stage_failed = stage;
    return RCLC_OK;
}

int rclc_calibrate_frames(rclc_ctx* ctx, const rclc_config* cfg, const void* const* scans, const int64_t* n_pts, int32_t F, int stride, int ioff, double x_min,
                          double x_max, const double* img_norm, int32_t flags, double* corners_out, int32_t* frame_status, int64_t* n_board,
                          rclc_gridfit_report* reports, double* Tcl_io, double* pose_costs, int32_t* pose_iterations, double* stage_ms)
{
    RCLC_CHECK(ctx && cfg && scans && n_pts && corners_out && frame_status && F > 0 && F <= 65535, RCLC_ERR_INVALID, "bad argument");
    for (int f = 0; f < F; ++f) RCLC_CHECK(scans[f] && n_pts[f] > 0 && n_pts[f] < (1ll << 30), RCLC_ERR_INVALID, "bad scan");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    const auto t0 = std::chrono::steady_clock::now();
    const int nc = (cfg->board_width - 1) * (cfg->board_height - 1);
    if (!ctx->calib_pool) ctx->calib_pool = new WorkerPool();
    WorkerPool* pool = static_cast<WorkerPool*>(ctx->calib_pool);
    const int W = ctx->calib_workers > 0 ? std::min<int>(ctx->calib_workers, F) : default_workers(F);
    RCLC_TRY(pool->ensure(ctx->device, W));
    // the moved board clouds of all frames, side by side: what the batched grid fit takes over
    std::vector<int64_t> off((size_t)F + 1, 0);
    for (int f = 0; f < F; ++f) off[(size_t)f + 1] = off[(size_t)f] + ((n_pts[f] + 63) & ~(int64_t)63);
    RCLC_TRY(ctx->d_boards.reserve((size_t)off[(size_t)F] * 16));
    float4* d_boards = ctx->d_boards.as<float4>();
    std::vector<int64_t> m((size_t)F, 0);
    std::vector<double> Tbl((size_t)F * 16, 0.0);
    std::vector<int> rcs((size_t)F, RCLC_OK);
    for (int f = 0; f < F; ++f) frame_status[f] = 0;
    std::vector<std::string> errs((size_t)W);
    std::atomic<int> next{0};
    std::vector<int64_t> launches((size_t)W, 0);
    pool->run(W, [&](int wi, rclc_ctx* wc) {
        const int64_t l0 = wc->launches;
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= F) break;
            frame_status[f] = 0;
            auto body = [&]() -> int {
                int rc_buf = RCLC_OK;
                float4* d_scan = chain_buf(wc, CB_SCAN, n_pts[f], &rc_buf);
                if (rc_buf != RCLC_OK) return rc_buf;
                {
                    LAB_TIMER(lab_up, LS_UPLOAD);
                    RCLC_TRY(upload_cloud_f4(wc, scans[f], stride, ioff, n_pts[f], d_scan));
                    RCLC_CUDA(cudaStreamSynchronize(wc->stream));
                }
                const float4* d_board = nullptr;
                int64_t mb = 0;
                RCLC_TRY(seg_board_dev(wc, cfg, d_scan, n_pts[f], x_min, x_max, &d_board, &mb));
                if (mb == 0) { frame_status[f] = 10; return RCLC_OK; }
                // the hand-off between the two programs (a PCD file in the reference) and T_base (Calibrate.cpp:98-99, src/global.cpp:116-122)
                float4* dst = d_boards + off[(size_t)f];
                RCLC_CUDA(cudaMemcpyAsync(dst, d_board, (size_t)mb * 16, cudaMemcpyDeviceToDevice, wc->stream));
                const float Tb[16] = {0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0, 0, 1};
                RCLC_TRY(dev_transform(wc, dst, mb, Tb));
                int stage = 0;
                RCLC_TRY(board_frame_dev(wc, cfg, dst, mb, &Tbl[(size_t)f * 16], &stage));
                RCLC_CUDA(cudaStreamSynchronize(wc->stream));
                if (stage != 0) { frame_status[f] = 20 + stage; return RCLC_OK; }
                m[(size_t)f] = mb;
                return RCLC_OK;
            };
            rcs[(size_t)f] = body();
            if (rcs[(size_t)f] != RCLC_OK) { errs[(size_t)wi] = rclc_last_error_string(); frame_status[f] = 1; }
        }
        launches[(size_t)wi] = wc->launches - l0;
    });
    for (int w = 0; w < W; ++w) ctx->launches += launches[(size_t)w];
    for (int f = 0; f < F; ++f)
        if (rcs[(size_t)f] != RCLC_OK) {
            for (const std::string& e : errs) if (!e.empty()) { set_last_error(e); break; }
            return rcs[(size_t)f];
        }
    const auto t1 = std::chrono::steady_clock::now();
#ifdef RCLC_LAB
    {
        const char* names[LS_COUNT] = {"upload", "roi+us", "peel", "seg clusters+gmm", "plane fit", "project+gmm", "block centroids", "board frame+move"};
        fprintf(stderr, "[rclc lab] %d frames, %d workers, chains %.2f ms wall; per frame (ms):", F, W, std::chrono::duration<double, std::milli>(t1 - t0).count());
        for (int k = 0; k < LS_COUNT; ++k) fprintf(stderr, " %s %.3f", names[k], g_lab_ns[k].exchange(0) * 1e-6 / F);
        fprintf(stderr, "\n");
    }
#endif
    // ---- opt_grid_fitting for every frame that came through, together ----
    std::vector<int> ok_frames;
    int64_t max_m = 0;
    for (int f = 0; f < F; ++f) { if (n_board) n_board[f] = m[(size_t)f]; if (m[(size_t)f] > 0) { ok_frames.push_back(f); max_m = std::max(max_m, m[(size_t)f]); } }
    const int K = (int)ok_frames.size();
    for (int f = 0; f < F; ++f) for (int k = 0; k < nc * 3; ++k) corners_out[(size_t)f * nc * 3 + k] = 0.0;
    if (K > 0) {
        rclc_batch*& b = ctx->calib_batch;
        if (!b || b->n_frames != K || b->max_pts < max_m || std::memcmp(&b->cfg, cfg, sizeof(*cfg)) != 0) {
            if (b) { rclc_batch_destroy(b); b = nullptr; }
            RCLC_TRY(rclc_batch_create(ctx, cfg, K, max_m + max_m / 4 + 1024, &b));
        }
        std::vector<double> Tk((size_t)K * 16), ck((size_t)K * nc * 3);
        std::vector<rclc_gridfit_report> rk((size_t)K);
        for (int k = 0; k < K; ++k) {
            const int f = ok_frames[(size_t)k];
            RCLC_TRY(rclc_batch_set_frame_dev(b, k, d_boards + off[(size_t)f], m[(size_t)f]));
            std::memcpy(&Tk[(size_t)k * 16], &Tbl[(size_t)f * 16], 128);
        }
        RCLC_TRY(rclc_batch_gridfit_solve(b, Tk.data(), nullptr, ck.data(), rk.data()));
        for (int k = 0; k < K; ++k) {
            const int f = ok_frames[(size_t)k];
            if (reports) reports[f] = rk[(size_t)k];
            if (rk[(size_t)k].status >= 2) { frame_status[f] = 24; continue; }     // (a Ceres FAILURE, status 1, is not: see rclc_est_board_corner)
            if (cfg->board_order == 1) flip_corners(cfg, &ck[(size_t)k * nc * 3]);
            std::memcpy(&corners_out[(size_t)f * nc * 3], &ck[(size_t)k * nc * 3], (size_t)nc * 24);
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    // ---- calib_opt (Calibrate.cpp:12-67): PnP start + pose_reprj_opt over the frames that have corners ----
    auto t3 = t2;
    if (img_norm && Tcl_io) {
        std::vector<double> pc, im;
        int Fk = 0;
        for (int f = 0; f < F; ++f) {
            if (frame_status[f] != 0) continue;
            pc.insert(pc.end(), &corners_out[(size_t)f * nc * 3], &corners_out[(size_t)f * nc * 3] + (size_t)nc * 3);
            im.insert(im.end(), &img_norm[(size_t)f * nc * 2], &img_norm[(size_t)f * nc * 2] + (size_t)nc * 2);
            ++Fk;
        }
        RCLC_CHECK(Fk >= 1, RCLC_ERR_NUMERIC, "no frame produced board corners");
        if (!(flags & RCLC_CALIB_KEEP_TCL)) {
            double R[9], t[3];
            const double focal = 0.5 * (cfg->fx + cfg->fy);
            RCLC_TRY(rclc_pnp_ransac(pc.data(), im.data(), (int64_t)Fk * nc, 8.0 / focal, 100, 0.99, 1u, R, t, nullptr, nullptr));
            rclc_quat_from_R(R, Tcl_io);
            Tcl_io[4] = t[0]; Tcl_io[5] = t[1]; Tcl_io[6] = t[2];
        }
        double c0 = 0, c1 = 0;
        int32_t its = 0;
        const int rc = rclc_pose_refine(ctx, cfg, pc.data(), im.data(), Fk, nc, Tcl_io, &c0, &c1, &its);
        if (pose_costs) { pose_costs[0] = c0; pose_costs[1] = c1; }
        if (pose_iterations) *pose_iterations = its;
        if (rc != RCLC_OK && rc != RCLC_ERR_NUMERIC) return rc;
        t3 = std::chrono::steady_clock::now();
    }
    if (stage_ms) {
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        stage_ms[0] = ms(t0, t1); stage_ms[1] = ms(t1, t2); stage_ms[2] = ms(t2, t3); stage_ms[3] = ms(t0, t3);
    }
    return RCLC_OK;
}

} // extern "C"
