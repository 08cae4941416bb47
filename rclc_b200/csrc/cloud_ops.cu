// cloud_ops.cu — per-point streaming operators (HBM-bound, float4 loads/stores):
//   K8 plane_project     pc_plane_prj                         include/pc_process.h:41-57
//      transform_cloud   pcl::transformPointCloud(Matrix4f)   (called at pc_process.h:594,618; opt_grid_fitting.h:54,107; Calibrate.cpp:99,134)
//   K7 project_colorize  colourise loop                       apps/Calibrate.cpp:134-148 + include/PinholeCam.h:93-110
#include "ops_dev.cuh"

#include <climits>

namespace rclc {

__global__ void k_plane_project(float4* __restrict__ p, int64_t n, double a, double b, double c, double d)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 q = p[i];
        // Vector2d(pt.x / pt.z, pt.y / pt.z): float divisions widened to double
        const double d0 = (double)__fdiv_rn(q.x, q.z), d1 = (double)__fdiv_rn(q.y, q.z);
        const double den = __dadd_rn(__dadd_rn(__dmul_rn(a, d0), __dmul_rn(b, d1)), c);
        const float z = (float)__ddiv_rn(-d, den);
        q.z = z;
        q.x = (float)__dmul_rn((double)z, d0);
        q.y = (float)__dmul_rn((double)z, d1);
        p[i] = q;
    }
}

struct Mat34f { float m[12]; };

__global__ void k_transform(float4* __restrict__ p, int64_t n, Mat34f T)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 q = p[i];
        const float x = affine_row_f(T.m[0], T.m[1], T.m[2], T.m[3], q.x, q.y, q.z);
        const float y = affine_row_f(T.m[4], T.m[5], T.m[6], T.m[7], q.x, q.y, q.z);
        const float z = affine_row_f(T.m[8], T.m[9], T.m[10], T.m[11], q.x, q.y, q.z);
        p[i] = make_float4(x, y, z, q.w);
    }
}

// cvRound(double): SSE2 cvtsd2si, round-half-even, INT_MIN when out of range / NaN
__device__ __forceinline__ int cv_round_dev(double v)
{
    if (!(fabs(v) < 2147483648.0)) return INT_MIN;
    return (int)rint(v);
}

// out: uchar4 {r,g,b,in_fov}, int2 pixel
__global__ void k_project_colorize(const float4* __restrict__ p, int64_t n, Mat34f T, double fx, double fy, double cx, double cy,
                                   int W, int H, const uint8_t* __restrict__ bgr, uchar4* __restrict__ rgba, int2* __restrict__ pix)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 q = __ldg(p + i);
        const double X = (double)affine_row_f(T.m[0], T.m[1], T.m[2], T.m[3], q.x, q.y, q.z);
        const double Y = (double)affine_row_f(T.m[4], T.m[5], T.m[6], T.m[7], q.x, q.y, q.z);
        const double Z = (double)affine_row_f(T.m[8], T.m[9], T.m[10], T.m[11], q.x, q.y, q.z);
        // PinholeCam.h:96-98: cvRound((X/Z)*fx + cx)
        const int u = cv_round_dev(__dadd_rn(__dmul_rn(__ddiv_rn(X, Z), fx), cx));
        const int v = cv_round_dev(__dadd_rn(__dmul_rn(__ddiv_rn(Y, Z), fy), cy));
        const bool in = !(u < 0 || v < 0 || u >= W || v >= H);          // :101-110 (no Z > 0 test, SURVEY C-9)
        uchar4 o = make_uchar4(0, 0, 0, 0);
        if (in) {
            const uint8_t* px = bgr + ((size_t)v * W + u) * 3;
            o = make_uchar4(__ldg(px + 2), __ldg(px + 1), __ldg(px + 0), 1);
        }
        rgba[i] = o;
        if (pix) pix[i] = make_int2(u, v);
    }
}

static int grid_for(rclc_ctx* ctx, int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16)); }

// download float4 xyz back into the caller's strided records (intensity untouched)
static int download_xyz(rclc_ctx* ctx, const float4* d, void* pts, int stride, int64_t n)
{
    RCLC_TRY(ctx->h_pin.reserve((size_t)n * 16));
    RCLC_CUDA(cudaMemcpyAsync(ctx->h_pin.p, d, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    const float* h = ctx->h_pin.as<float>();
    char* dst = static_cast<char*>(pts);
    for (int64_t k = 0; k < n; ++k) std::memcpy(dst + k * (int64_t)stride, h + 4 * k, 12);
    return RCLC_OK;
}

int dev_plane_project(rclc_ctx* ctx, float4* d_pts, int64_t n, const double plane[4])
{
    if (n == 0) return RCLC_OK;
    k_plane_project<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(d_pts, n, plane[0], plane[1], plane[2], plane[3]);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

int dev_transform(rclc_ctx* ctx, float4* d_pts, int64_t n, const float T[16])
{
    if (n == 0) return RCLC_OK;
    Mat34f M;
    for (int k = 0; k < 12; ++k) M.m[k] = T[k];
    k_transform<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(d_pts, n, M);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_plane_project(rclc_ctx* ctx, void* pts_io, int stride, int64_t n, const double plane[4])
{
    RCLC_CHECK(ctx && pts_io && plane && n > 0 && stride >= 16, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(upload_cloud_f4(ctx, pts_io, stride, 12, n, ctx->d_pts.as<float4>()));
    RCLC_TRY(dev_plane_project(ctx, ctx->d_pts.as<float4>(), n, plane));
    return download_xyz(ctx, ctx->d_pts.as<float4>(), pts_io, stride, n);
}

int rclc_transform_cloud(rclc_ctx* ctx, void* pts_io, int stride, int64_t n, const float T[16])
{
    RCLC_CHECK(ctx && pts_io && T && n > 0 && stride >= 16, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(upload_cloud_f4(ctx, pts_io, stride, 12, n, ctx->d_pts.as<float4>()));
    RCLC_TRY(dev_transform(ctx, ctx->d_pts.as<float4>(), n, T));
    return download_xyz(ctx, ctx->d_pts.as<float4>(), pts_io, stride, n);
}

int rclc_project_colorize(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int64_t n,
                          const float T_cl[16], const uint8_t* bgr, uint8_t* rgb_out, int32_t* pix_out, uint8_t* infov_out)
{
    RCLC_CHECK(ctx && cfg && pts && T_cl && bgr && n > 0 && stride >= 16, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    const size_t img_bytes = (size_t)cfg->cam_width * cfg->cam_height * 3;
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(ctx->d_tmp.reserve(img_bytes));
    RCLC_TRY(ctx->d_out.reserve((size_t)n * 12 + 256));
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, 12, n, ctx->d_pts.as<float4>()));
    RCLC_CUDA(cudaMemcpyAsync(ctx->d_tmp.p, bgr, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
    Mat34f M;
    for (int k = 0; k < 12; ++k) M.m[k] = T_cl[k];
    uchar4* d_rgba = ctx->d_out.as<uchar4>();
    int2* d_pix = reinterpret_cast<int2*>(static_cast<char*>(ctx->d_out.p) + (((size_t)n * 4 + 255) & ~(size_t)255));
    k_project_colorize<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(ctx->d_pts.as<float4>(), n, M, cfg->fx, cfg->fy, cfg->cx, cfg->cy,
                                                                 cfg->cam_width, cfg->cam_height, ctx->d_tmp.as<uint8_t>(), d_rgba,
                                                                 pix_out ? d_pix : nullptr);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    RCLC_TRY(ctx->h_pin.reserve((size_t)n * 12 + 256));
    uint8_t* h = ctx->h_pin.as<uint8_t>();
    RCLC_CUDA(cudaMemcpyAsync(h, d_rgba, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (pix_out) RCLC_CUDA(cudaMemcpyAsync(pix_out, d_pix, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t k = 0; k < n; ++k) {
        const bool in = h[4 * k + 3] != 0;
        if (infov_out) infov_out[k] = in ? 1 : 0;
        if (in && rgb_out) { rgb_out[3 * k] = h[4 * k]; rgb_out[3 * k + 1] = h[4 * k + 1]; rgb_out[3 * k + 2] = h[4 * k + 2]; }
    }
    return RCLC_OK;
}

} // extern "C"
