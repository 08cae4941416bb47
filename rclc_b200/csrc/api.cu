// api.cu — context lifecycle, timing helpers and small shared host utilities of the C ABI.
#define RCLC_CONFIG_IMPLEMENTATION
#include "common.cuh"
#include <omp.h>
#include <thread>
#include <algorithm>

namespace rclc {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }



// Host side of an upload: the strided records are packed to float4 {x,y,z,I} into one of two pinned buffers by a few
// threads, then copied asynchronously; the copy of frame f overlaps the packing of frame f + 1 (the buffers are guarded by
// events, the stream orders the kernels that follow behind the copy).
static void pack_range(const char* src, int stride, int ioff, int64_t k0, int64_t k1, float* o)
{
    if (stride == 16 && ioff == 12) {
        std::memcpy(o + 4 * k0, src + k0 * 16, (size_t)(k1 - k0) * 16);
        return;
    }
    for (int64_t k = k0; k < k1; ++k) {
        const char* r = src + k * (int64_t)stride;
        std::memcpy(o + 4 * k, r, 12);
        std::memcpy(o + 4 * k + 3, r + ioff, 4);
    }
}

int upload_cloud_f4(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, float4* dst, cudaStream_t st)
{
    if (!st) st = ctx->stream;
    RCLC_CHECK(stride >= 16 && ioff + 4 <= stride && ioff >= 12, RCLC_ERR_INVALID, "bad stride/ioff");
    if (n == 0) return RCLC_OK;
    if (stride == 16 && ioff == 12) {
        // {x, y, z, I} records that already sit in page-locked host memory need no staging copy: the DMA reads them in place.
        // (The copy is asynchronous: the caller keeps the buffer unchanged until the next call that returns results.)
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, pts) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            RCLC_CUDA(cudaMemcpyAsync(dst, pts, (size_t)n * 16, cudaMemcpyHostToDevice, st));
            return RCLC_OK;
        }
        cudaGetLastError();                                         // (older drivers flag unregistered host pointers as an error)
    }
    const int bi = ctx->up_flip;
    ctx->up_flip ^= 1;
    if (!ctx->ev_up[bi]) RCLC_CUDA(cudaEventCreateWithFlags(&ctx->ev_up[bi], cudaEventDisableTiming));
    RCLC_CUDA(cudaEventSynchronize(ctx->ev_up[bi]));               // previous copy out of this buffer (no-op the first time)
    RCLC_TRY(ctx->h_up[bi].reserve((size_t)n * 16));
    float* o = ctx->h_up[bi].as<float>();
    const char* src = static_cast<const char*>(pts);
    // OpenMP keeps its worker threads between calls (spawning std::threads per frame cost more than the copy itself)
    const int64_t chunk = 65536;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    // up to 8 threads per context, sized from the machine and the ranks sharing it rather than from OMP_NUM_THREADS: torchrun
    // exports OMP_NUM_THREADS=1 to every rank, which made the multi-GPU uploads single-threaded (RCLC_UPLOAD_THREADS overrides)
    static const int want = [] {
        if (const char* e = std::getenv("RCLC_UPLOAD_THREADS")) return std::max(1, atoi(e));
        const char* lw = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
        return std::max(1, std::min(8, hw / ranks));
    }();
    const int nth = (int)std::min<int64_t>(want, nchunks);
#pragma omp parallel for num_threads(nth) schedule(static)
    for (int64_t c = 0; c < nchunks; ++c) pack_range(src, stride, ioff, c * chunk, std::min(n, (c + 1) * chunk), o);
    RCLC_CUDA(cudaMemcpyAsync(dst, o, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    RCLC_CUDA(cudaEventRecord(ctx->ev_up[bi], st));
    return RCLC_OK;
}

__global__ void k_fill(float4* p, size_t n, float v)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) p[i] = make_float4(v, v, v, v);
}

// reads n float4 and keeps the loads alive through a (never taken) store
__global__ void k_read_sink(const float4* __restrict__ p, size_t n, float* sink)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    float s = 0.f;
    for (; i < n; i += step) { const float4 v = __ldcs(p + i); s += v.x + v.y + v.z + v.w; }
    if (s == 123.456f) *sink = s;
}

} // namespace rclc

extern "C" {

const char* rclc_last_error_string(void) { return rclc::g_last_error.c_str(); }
int rclc_version(void) { return 100; }
uint64_t rclc_config_sizeof(void) { return sizeof(rclc_config); }

int rclc_ctx_create(int device, rclc_ctx** out)
{
    RCLC_CHECK(out != nullptr, RCLC_ERR_INVALID, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        rclc::set_last_error(std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return RCLC_ERR_CUDA;
    }
    RCLC_CHECK(device >= 0 && device < count, RCLC_ERR_INVALID, "device index out of range");
    RCLC_CUDA(cudaSetDevice(device));
    rclc_ctx* c = new rclc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    RCLC_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->l2_bytes = prop.l2CacheSize;
    RCLC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RCLC_CUDA(cudaEventCreate(&c->ev0));
    RCLC_CUDA(cudaEventCreate(&c->ev1));
    RCLC_CUDA(cudaEventCreate(&c->evk0));
    RCLC_CUDA(cudaEventCreate(&c->evk1));
    *out = c;
    return RCLC_OK;
}

int rclc_ctx_destroy(rclc_ctx* c)
{
    if (!c) return RCLC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    rclc::destroy_calib_state(c);
    for (rclc::Buf& b : c->d_chain) b.release();
    c->d_boards.release(); c->h_chain.release();
    c->d_pts.release(); c->d_tmp.release(); c->d_tmp2.release(); c->d_out.release(); c->d_flush.release();
    c->h_pin.release(); c->h_pin2.release();
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->evk0);
    cudaEventDestroy(c->evk1);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    if (c->ev_rank) cudaEventDestroy(c->ev_rank);
    for (cudaEvent_t e : c->ev_groups) cudaEventDestroy(e);
    for (cudaStream_t st : c->aux_streams) cudaStreamDestroy(st);
    for (int i = 0; i < 2; ++i) { if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]); c->h_up[i].release(); }
    cudaStreamDestroy(c->stream);
    delete c;
    return RCLC_OK;
}

int rclc_ctx_synchronize(rclc_ctx* c)
{
    RCLC_CHECK(c, RCLC_ERR_INVALID, "ctx is null");
    RCLC_CUDA(cudaStreamSynchronize(c->stream));
    for (cudaStream_t st : c->aux_streams) RCLC_CUDA(cudaStreamSynchronize(st));      // (uploads of the later solve groups live there)
    return RCLC_OK;
}

void* rclc_ctx_stream(rclc_ctx* c) { return c ? (void*)c->stream : nullptr; }

int rclc_ctx_device_info(rclc_ctx* c, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes)
{
    RCLC_CHECK(c, RCLC_ERR_INVALID, "ctx is null");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (l2_bytes) *l2_bytes = c->l2_bytes;
    return RCLC_OK;
}

int rclc_ctx_timer_start(rclc_ctx* c)
{
    RCLC_CHECK(c, RCLC_ERR_INVALID, "ctx is null");
    RCLC_CUDA(cudaEventRecord(c->ev0, c->stream));
    return RCLC_OK;
}

int rclc_ctx_timer_stop_ms(rclc_ctx* c, float* ms_out)
{
    RCLC_CHECK(c && ms_out, RCLC_ERR_INVALID, "null argument");
    RCLC_CUDA(cudaEventRecord(c->ev1, c->stream));
    RCLC_CUDA(cudaEventSynchronize(c->ev1));
    RCLC_CUDA(cudaEventElapsedTime(ms_out, c->ev0, c->ev1));
    return RCLC_OK;
}

int rclc_ctx_flush_l2(rclc_ctx* c)
{
    RCLC_CHECK(c, RCLC_ERR_INVALID, "ctx is null");
    // Two regions of 2 x L2 each, READ alternately: every flush misses everywhere and leaves L2 full of CLEAN foreign
    // lines. (A write flush leaves it full of dirty lines, and the kernel timed next would pay for their write-back.)
    const size_t half = (size_t)std::max<int64_t>(c->l2_bytes * 2, 256ll << 20);
    if (c->d_flush.cap < 2 * half + 256) {
        RCLC_TRY(c->d_flush.reserve(2 * half + 256));
        rclc::k_fill<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_flush.as<float4>(), (2 * half + 256) / 16, 1.0f);
    }
    c->flush_parity ^= 1;
    const float4* region = c->d_flush.as<float4>() + (c->flush_parity ? half / 16 : 0);
    rclc::k_read_sink<<<c->sm_count * 8, 256, 0, c->stream>>>(region, half / 16, reinterpret_cast<float*>(c->d_flush.as<char>() + 2 * half));
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

int64_t rclc_ctx_kernel_launches(rclc_ctx* c) { return c ? c->launches : 0; }

int rclc_ctx_set_solve_groups(rclc_ctx* c, int32_t groups)
{
    RCLC_CHECK(c && groups >= 1 && groups <= 32, RCLC_ERR_INVALID, "solve groups must be 1..32");
    c->solve_groups = groups;
    return RCLC_OK;
}

int rclc_ctx_last_sweep_ms(rclc_ctx* c, float* ms_out)
{
    RCLC_CHECK(c && ms_out, RCLC_ERR_INVALID, "null argument");
    RCLC_CUDA(cudaEventSynchronize(c->evk1));
    RCLC_CUDA(cudaEventElapsedTime(ms_out, c->evk0, c->evk1));
    return RCLC_OK;
}

} // extern "C"
