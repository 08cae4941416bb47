// cluster.cu — voxel uniform sampling and Euclidean clustering on device (SURVEY.md §8f "next #2").
//
// Replaces pcl::UniformSampling (pc_process.h:111-115, BoardSegmentation.cpp:58-61) and pcl::EuclideanClusterExtraction
// (pc_process.h:62-72, 131-139) with the semantics SURVEY.md Appendix A lists:
//   * uniform sampling: one hash-table slot per occupied voxel, the winner of a voxel decided by a 64-bit atomicMin over
//     (float distance to the voxel centre, point index) — the same float expressions as PCL, so the same point wins — and the
//     kept indices sorted by voxel key (cub radix sort of the survivors);
//   * clustering: the host kd-tree region growing becomes connected components over a uniform grid of cells of 0.57 tolerance:
//     a cell's points are within the tolerance of each other by construction, so the union-find runs over CELLS: points sorted
//     by cell, a hash table cell -> range, every pair of cells within the 5 x 5 x 5 neighbourhood is joined as soon as one point
//     pair passes FLANN's L2_Simple test; roots are the smallest member index. Sizes and the final ranking (size descending,
//     smallest member first) follow on the few roots.
#include "ops_dev.cuh"

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

namespace rclc {

constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ uint32_t hash64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// slot of `key` in an open-addressing table of `cap` (power of two) keys, inserting it if absent
__device__ __forceinline__ uint32_t table_insert(unsigned long long* keys, uint32_t cap, unsigned long long key)
{
    uint32_t s = hash64(key) & (cap - 1);
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[s], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) return s;
        s = (s + 1) & (cap - 1);
    }
}
__device__ __forceinline__ int table_find(const unsigned long long* keys, uint32_t cap, unsigned long long key)
{
    uint32_t s = hash64(key) & (cap - 1);
    for (;;) {
        const unsigned long long k = keys[s];
        if (k == key) return (int)s;
        if (k == kEmptyKey) return -1;
        s = (s + 1) & (cap - 1);
    }
}

__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// ---- bounding box of the finite points: out[0..2] min, out[3..5] max as ordered-int bits ----
__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_bbox_init(int* bb) { if (threadIdx.x < 3) bb[threadIdx.x] = f2o(3.0e38f); else if (threadIdx.x < 6) bb[threadIdx.x] = f2o(-3.0e38f); }

__global__ void __launch_bounds__(256) k_bbox(const float4* __restrict__ pts, int64_t n, int* __restrict__ bb)
{
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const float4 p = pts[i];
        if (!finite3(p)) continue;
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bb[d], f2o(mn[d])); atomicMax(&bb[3 + d], f2o(mx[d])); }
    }
}

// ---- uniform sampling ----
struct UsGeom { float leaf, inv; long long mb[3], div[3]; };

__device__ __forceinline__ bool us_key(const UsGeom& g, const float4& p, unsigned long long& key, float& dist)
{
    if (!finite3(p)) return false;
    const float c[3] = {p.x, p.y, p.z};
    long long ijk[3];
    dist = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float fl = floorf(__fmul_rn(c[d], g.inv));
        ijk[d] = (long long)fl;
        const float ctr = __fmul_rn(__fadd_rn(fl, 0.5f), g.leaf);
        const float e = __fsub_rn(c[d], ctr);
        const float e2 = __fmul_rn(e, e);
        dist = d == 0 ? e2 : __fadd_rn(dist, e2);
    }
    key = (unsigned long long)((ijk[0] - g.mb[0]) + (ijk[1] - g.mb[1]) * g.div[0] + (ijk[2] - g.mb[2]) * g.div[0] * g.div[1]);
    return true;
}

__global__ void __launch_bounds__(256) k_us_vote(const float4* __restrict__ pts, int64_t n, UsGeom g, unsigned long long* __restrict__ keys,
                                                 unsigned long long* __restrict__ best, uint32_t cap)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        unsigned long long key;
        float dist;
        if (!us_key(g, pts[i], key, dist)) continue;
        const uint32_t s = table_insert(keys, cap, key);
        atomicMin(&best[s], ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(uint32_t)i);      // dist >= 0: bits order as values
    }
}

__global__ void __launch_bounds__(256) k_us_collect(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ best, uint32_t cap,
                                                    unsigned long long* __restrict__ out_keys, int* __restrict__ out_idx, int* __restrict__ n_out)
{
    for (uint32_t s = blockIdx.x * 256u + threadIdx.x; s < cap; s += gridDim.x * 256u) {
        if (keys[s] == kEmptyKey) continue;
        const int o = atomicAdd(n_out, 1);
        out_keys[o] = keys[s];
        out_idx[o] = (int)(best[s] & 0xffffffffull);
    }
}

// ---- clustering ----
// The grid of the clustering: cells of side 0.57 tol, so that two points of one cell are within tol of each other whatever they
// are (the cell's diagonal is 0.987 tol) — a cell is connected by construction — and a point's neighbours within tol lie in the
// 5 x 5 x 5 cells around its own. Keys pack the three cell coordinates (biased by 2) into bx + by + bz bits, sized from the cloud's
// extent, so the radix sort runs over the bits in use only.
struct EcGeom { double inv_cell; long long mb[3]; float r2; int bx, by, bz; };
constexpr double kEcCellPerTol = 0.57;

__device__ __forceinline__ bool ec_cell(const EcGeom& g, const float4& p, long long (&c)[3])
{
    if (!finite3(p)) return false;
    c[0] = (long long)floor((double)p.x * g.inv_cell) - g.mb[0];
    c[1] = (long long)floor((double)p.y * g.inv_cell) - g.mb[1];
    c[2] = (long long)floor((double)p.z * g.inv_cell) - g.mb[2];
    return true;
}
__device__ __forceinline__ unsigned long long ec_pack(const EcGeom& g, long long x, long long y, long long z)
{
    return (unsigned long long)(x + 2) | ((unsigned long long)(y + 2) << g.bx) | ((unsigned long long)(z + 2) << (g.bx + g.by));
}

__global__ void __launch_bounds__(256) k_ec_keys(const float4* __restrict__ pts, int64_t n, EcGeom g, unsigned long long* __restrict__ keys, int* __restrict__ idx,
                                                 int* __restrict__ parent)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        long long c[3];
        parent[i] = (int)i;
        keys[i] = ec_cell(g, pts[i], c) ? ec_pack(g, c[0], c[1], c[2]) : kEmptyKey;   // non-finite points sort to the end
        idx[i] = (int)i;
    }
}

// after the sort: one table entry per occupied cell -> first sorted position; the cell ends where the key changes
__global__ void __launch_bounds__(256) k_ec_cells(const unsigned long long* __restrict__ skeys, int64_t n, unsigned long long* __restrict__ tkeys,
                                                  int* __restrict__ tstart, uint32_t cap)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const unsigned long long k = skeys[i];
        if (k == kEmptyKey) continue;
        if (i == 0 || skeys[i - 1] != k) tstart[table_insert(tkeys, cap, k)] = (int)i;
    }
}

__device__ __forceinline__ int uf_find(int* parent, int a)
{
    int p = parent[a];
    while (p != a) { const int gp = parent[p]; if (gp != p) parent[a] = gp; a = p; p = parent[a]; }
    return a;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // a > b: hang the larger root under the smaller
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;                                    // a was still a root: linked
        a = old;                                                 // somebody re-rooted a meanwhile: retry from there
    }
}

// every point hangs under its cell's first point — the smallest index of the cell, the sort being stable
__global__ void __launch_bounds__(256) k_ec_attach(const unsigned long long* __restrict__ skeys, const int* __restrict__ sidx, int64_t n,
                                                   const unsigned long long* __restrict__ tkeys, const int* __restrict__ tstart, uint32_t cap, int* __restrict__ parent,
                                                   int* __restrict__ root_counter)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) root_counter[0] = 0;       // (the list of roots starts empty; its space was the sort's input until now)
    for (int64_t s = blockIdx.x * 256ll + threadIdx.x; s < n; s += gridDim.x * 256ll) {
        const unsigned long long k = skeys[s];
        if (k == kEmptyKey) continue;
        parent[sidx[s]] = sidx[tstart[table_find(tkeys, cap, k)]];
    }
}

// Cells are joined pairwise: one thread per (occupied cell A, one of the 62 cell offsets that come after the centre in the
// 5 x 5 x 5 block, so every unordered pair of cells is looked at once). The pair is joined as soon as ONE point of A and one of
// B pass FLANN's test (L2_Simple in float, in order; RadiusResultSet: dist < radius^2) — the partition is that of the full
// radius graph, with a union-find over cells instead of points and a handful of distance tests per pair instead of all of them.
constexpr int kEcOffsets = 62;
__global__ void __launch_bounds__(128) k_ec_link(const float4* __restrict__ pts, const unsigned long long* __restrict__ skeys, const int* __restrict__ sidx,
                                                 int64_t n, EcGeom g, const unsigned long long* __restrict__ tkeys, const int* __restrict__ tstart,
                                                 uint32_t cap, int* __restrict__ parent)
{
    const unsigned long long mx = (1ull << g.bx) - 1, my = (1ull << g.by) - 1, mz = (1ull << g.bz) - 1;
    for (int64_t t = blockIdx.x * 128ll + threadIdx.x; t < n * kEcOffsets; t += gridDim.x * 128ll) {
        const int64_t a0 = t / kEcOffsets;
        const int o = (int)(t - a0 * kEcOffsets) + 63;                                // linear offset index 63 .. 124 (62 is the centre)
        const unsigned long long kA = skeys[a0];
        if (kA == kEmptyKey || (a0 > 0 && skeys[a0 - 1] == kA)) continue;             // cell heads only
        const long long cx = (long long)(kA & mx) - 2, cy = (long long)((kA >> g.bx) & my) - 2, cz = (long long)((kA >> (g.bx + g.by)) & mz) - 2;
        const long long x = cx + (o % 5 - 2), y = cy + ((o / 5) % 5 - 2), z = cz + (o / 25 - 2);
        if (x < -2 || y < -2 || z < -2 || (unsigned long long)(x + 2) > mx || (unsigned long long)(y + 2) > my || (unsigned long long)(z + 2) > mz) continue;
        const unsigned long long kB = ec_pack(g, x, y, z);
        const int slot = table_find(tkeys, cap, kB);
        if (slot < 0) continue;
        const int b0 = tstart[slot];
        const int ra = sidx[a0], rb = sidx[b0];                                        // the cells' representatives
        if (uf_find(parent, ra) == uf_find(parent, rb)) continue;                     // joined through other cells already
        bool hit = false;
        for (int64_t a = a0; !hit && a < n && skeys[a] == kA; ++a) {
            const float4 p = pts[sidx[a]];
            for (int64_t b = b0; b < n && skeys[b] == kB; ++b) {
                const float4 q = pts[sidx[b]];
                const float ex = __fsub_rn(p.x, q.x), ey = __fsub_rn(p.y, q.y), ez = __fsub_rn(p.z, q.z);
                float d = __fmul_rn(ex, ex);                                         // flann L2_Simple, in order
                d = __fadd_rn(d, __fmul_rn(ey, ey));
                d = __fadd_rn(d, __fmul_rn(ez, ez));
                if (d < g.r2) { hit = true; break; }
            }
        }
        if (hit) uf_union(parent, ra, rb);
    }
}

__global__ void __launch_bounds__(256) k_iota(int* a, int64_t n) { for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) a[i] = (int)i; }

// the root of every point (read-only walk: a compressing find running next to the store of a point's final root could leave that
// point hanging under an inner node) and the size of every component
__global__ void __launch_bounds__(256) k_ec_roots(const float4* __restrict__ pts, int64_t n, const int* __restrict__ parent, int* __restrict__ root,
                                                  int* __restrict__ size)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        if (!finite3(pts[i])) { root[i] = -1; continue; }
        int r = (int)i;
        for (int p = parent[r]; p != r; p = parent[r]) r = p;
        root[i] = r;
        atomicAdd(&size[r], 1);
    }
}

static uint32_t pow2_at_least(uint64_t v) { uint32_t c = 1024; while (c < v && c < (1u << 31)) c <<= 1; return c; }
static int grid_for(int64_t n, int threads, int sms) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)sms * 8)); }

// roots of components: (root index, size) pairs appended in no particular order (the host ranks them)
__global__ void __launch_bounds__(256) k_ec_collect_roots(const int* __restrict__ size, int64_t n, int2* __restrict__ out, int* __restrict__ n_out)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const int sz = size[i];
        if (sz > 0) out[atomicAdd(n_out, 1)] = make_int2((int)i, sz);
    }
}
__global__ void __launch_bounds__(256) k_fill_int(int* a, int64_t n, int v) { for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) a[i] = v; }
__global__ void __launch_bounds__(256) k_ec_set_rank(const int2* __restrict__ ranked, int n_ranked, int* __restrict__ rank_of)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < n_ranked) rank_of[ranked[k].x] = ranked[k].y;
}
__global__ void __launch_bounds__(256) k_ec_labels(const int* root, const int* __restrict__ rank_of, int64_t n, int* label)      // (in place)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) { const int r = root[i]; label[i] = r >= 0 ? rank_of[r] : -1; }
}
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ src, const int* __restrict__ idx, int64_t n, float4* __restrict__ dst)
{
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) dst[i] = src[idx[i]];
}

int dev_gather(rclc_ctx* ctx, const float4* d_src, const int* d_idx, int64_t n, float4* d_dst)
{
    if (n == 0) return RCLC_OK;
    k_gather<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(d_src, d_idx, n, d_dst);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

int dev_uniform_sample(rclc_ctx* ctx, const float4* d_pts, int64_t n, double leaf_d, const int** d_idx, int64_t* n_out)
{
    cudaStream_t st = ctx->stream;
    const uint32_t cap = pow2_at_least((uint64_t)n * 2);
    // scratch: bbox | table keys | table best | out keys (x2 for the sort) | out idx (x2) | counter | cub temp
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, 0, 64, st);
    const size_t o_bb = 0, o_cnt = 64, o_keys = 256, o_best = o_keys + (size_t)cap * 8, o_ok = o_best + (size_t)cap * 8, o_ok2 = o_ok + (size_t)n * 8,
                 o_oi = o_ok2 + (size_t)n * 8, o_oi2 = o_oi + (((size_t)n * 4 + 255) & ~(size_t)255), o_tmp = o_oi2 + (((size_t)n * 4 + 255) & ~(size_t)255);
    RCLC_TRY(ctx->d_tmp.reserve(o_tmp + temp_bytes + 256));
    char* base = ctx->d_tmp.as<char>();
    int* d_bb = reinterpret_cast<int*>(base + o_bb);
    int* d_cnt = reinterpret_cast<int*>(base + o_cnt);
    unsigned long long* d_keys = reinterpret_cast<unsigned long long*>(base + o_keys);
    unsigned long long* d_best = reinterpret_cast<unsigned long long*>(base + o_best);
    const int sms = ctx->sm_count;
    k_bbox_init<<<1, 32, 0, st>>>(d_bb);
    k_bbox<<<grid_for(n, 256, sms), 256, 0, st>>>(d_pts, n, d_bb);
    ctx->launches += 2;
    RCLC_TRY(ctx->h_pin2.reserve(64));
    int* h_bb = ctx->h_pin2.as<int>();
    RCLC_CUDA(cudaMemcpyAsync(h_bb, d_bb, 24, cudaMemcpyDeviceToHost, st));
    RCLC_CUDA(cudaStreamSynchronize(st));
    auto o2f_h = [](int i) { const int v = i >= 0 ? i : i ^ 0x7fffffff; float f; std::memcpy(&f, &v, 4); return f; };
    *n_out = 0;
    *d_idx = nullptr;
    if (o2f_h(h_bb[0]) > o2f_h(h_bb[3])) return RCLC_OK;                                     // no finite point
    UsGeom g;
    g.leaf = (float)leaf_d;
    g.inv = 1.0f / g.leaf;
    for (int d = 0; d < 3; ++d) {
        g.mb[d] = (long long)std::floor(o2f_h(h_bb[d]) * g.inv);
        g.div[d] = (long long)std::floor(o2f_h(h_bb[3 + d]) * g.inv) - g.mb[d] + 1;
    }
    RCLC_CHECK((double)g.div[0] * (double)g.div[1] * (double)g.div[2] < 9.0e18, RCLC_ERR_UNSUPPORTED, "voxel grid too large for 64-bit keys");
    RCLC_CUDA(cudaMemsetAsync(d_keys, 0xff, (size_t)cap * 16, st));                            // keys and best
    RCLC_CUDA(cudaMemsetAsync(d_cnt, 0, 4, st));
    k_us_vote<<<grid_for(n, 256, sms), 256, 0, st>>>(d_pts, n, g, d_keys, d_best, cap);
    unsigned long long* d_ok = reinterpret_cast<unsigned long long*>(base + o_ok);
    unsigned long long* d_ok2 = reinterpret_cast<unsigned long long*>(base + o_ok2);
    int* d_oi = reinterpret_cast<int*>(base + o_oi);
    int* d_oi2 = reinterpret_cast<int*>(base + o_oi2);
    k_us_collect<<<grid_for(cap, 256, sms), 256, 0, st>>>(d_keys, d_best, cap, d_ok, d_oi, d_cnt);
    ctx->launches += 2;
    RCLC_CUDA(cudaMemcpyAsync(h_bb, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    RCLC_CUDA(cudaStreamSynchronize(st));
    const int kept = h_bb[0];
    int key_bits = 1;                                                                          // voxel keys are below div0 * div1 * div2: sort those bits only
    while (key_bits < 64 && (double)(1ull << key_bits) < (double)g.div[0] * (double)g.div[1] * (double)g.div[2]) ++key_bits;
    RCLC_CUDA(cub::DeviceRadixSort::SortPairs(base + o_tmp, temp_bytes, d_ok, d_ok2, d_oi, d_oi2, kept, 0, key_bits, st));
    ctx->launches += 1;
    *n_out = kept;
    *d_idx = d_oi2;
    return RCLC_OK;
}

int dev_euclidean_cluster(rclc_ctx* ctx, const float4* d_pts, int64_t n, double tol, int64_t min_size, int64_t max_size, ClusterSet* out)
{
    cudaStream_t st = ctx->stream;
    const uint32_t cap = pow2_at_least((uint64_t)n * 2);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, 0, 64, st);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_bb = 0, o_k = 256, o_k2 = o_k + al((size_t)(n + 1) * 8), o_i = o_k2 + al((size_t)n * 8), o_i2 = o_i + al((size_t)n * 4), o_tk = o_i2 + al((size_t)n * 4),
                 o_ts = o_tk + (size_t)cap * 8, o_par = o_ts + (size_t)cap * 4, o_size = o_par + al((size_t)n * 4), o_lab = o_size + al((size_t)n * 4),
                 o_tmp = o_lab + al((size_t)n * 4);
    RCLC_TRY(ctx->d_tmp.reserve(o_tmp + temp_bytes + 256));
    char* base = ctx->d_tmp.as<char>();
    int* d_bb = reinterpret_cast<int*>(base + o_bb);
    int* d_label = reinterpret_cast<int*>(base + o_lab);
    const int sms = ctx->sm_count;
    out->d_label = d_label;
    out->sizes.clear();
    constexpr int kRootsFirst = 2047;                                                         // roots that come back with the count in one copy
    RCLC_TRY(ctx->h_pin2.reserve(std::max<size_t>((size_t)(n + 1) * 8, (size_t)(kRootsFirst + 1) * 8) + 64));
    if (!out->have_bbox) {
        k_bbox_init<<<1, 32, 0, st>>>(d_bb);
        k_bbox<<<grid_for(n, 256, sms), 256, 0, st>>>(d_pts, n, d_bb);
        ctx->launches += 2;
        int* h_bb = ctx->h_pin2.as<int>();
        RCLC_CUDA(cudaMemcpyAsync(h_bb, d_bb, 24, cudaMemcpyDeviceToHost, st));
        RCLC_CUDA(cudaStreamSynchronize(st));
        auto o2f_h = [](int i) { const int v = i >= 0 ? i : i ^ 0x7fffffff; float f; std::memcpy(&f, &v, 4); return f; };
        for (int d = 0; d < 6; ++d) out->bbox[d] = o2f_h(h_bb[d]);
        out->have_bbox = true;
    }
    if (out->bbox[0] > out->bbox[3]) {                                                        // no finite point
        k_fill_int<<<grid_for(n, 256, sms), 256, 0, st>>>(d_label, n, -1);
        ctx->launches += 1;
        return RCLC_OK;
    }
    EcGeom g;
    g.inv_cell = 1.0 / (tol * kEcCellPerTol);
    g.r2 = (float)(tol * tol);
    int bits[3];
    for (int d = 0; d < 3; ++d) {
        g.mb[d] = (long long)std::floor((double)out->bbox[d] * g.inv_cell);
        const long long ext = (long long)std::floor((double)out->bbox[3 + d] * g.inv_cell) - g.mb[d] + 6;      // + the bias of 2 and the two cells beyond
        bits[d] = 1;
        while ((1ll << bits[d]) <= ext) ++bits[d];
    }
    g.bx = bits[0]; g.by = bits[1]; g.bz = bits[2];
    const int key_bits = g.bx + g.by + g.bz;
    RCLC_CHECK(key_bits <= 63, RCLC_ERR_UNSUPPORTED, "cloud extent / tolerance exceeds 2^63 cells");
    unsigned long long* d_k = reinterpret_cast<unsigned long long*>(base + o_k);
    unsigned long long* d_k2 = reinterpret_cast<unsigned long long*>(base + o_k2);
    int* d_i = reinterpret_cast<int*>(base + o_i);
    int* d_i2 = reinterpret_cast<int*>(base + o_i2);
    unsigned long long* d_tk = reinterpret_cast<unsigned long long*>(base + o_tk);
    int* d_ts = reinterpret_cast<int*>(base + o_ts);
    int* d_par = reinterpret_cast<int*>(base + o_par);
    int* d_size = reinterpret_cast<int*>(base + o_size);
    // the roots (a few, against n points) go to the host for the ranking: slot 0 of the list is its counter, and the list takes
    // the space of the unsorted keys, which are dead after the sort
    int2* d_roots = reinterpret_cast<int2*>(d_k);
    k_ec_keys<<<grid_for(n, 256, sms), 256, 0, st>>>(d_pts, n, g, d_k, d_i, d_par);
    RCLC_CUDA(cub::DeviceRadixSort::SortPairs(base + o_tmp, temp_bytes, d_k, d_k2, d_i, d_i2, (int)n, 0, key_bits, st));     // (the empty key is all ones: it sorts last on any prefix)
    RCLC_CUDA(cudaMemsetAsync(d_tk, 0xff, (size_t)cap * 8, st));
    RCLC_CUDA(cudaMemsetAsync(d_size, 0, (size_t)n * 4, st));
    k_ec_cells<<<grid_for(n, 256, sms), 256, 0, st>>>(d_k2, n, d_tk, d_ts, cap);
    k_ec_attach<<<grid_for(n, 256, sms), 256, 0, st>>>(d_k2, d_i2, n, d_tk, d_ts, cap, d_par, reinterpret_cast<int*>(d_roots));
    k_ec_link<<<grid_for(n * kEcOffsets, 128, sms), 128, 0, st>>>(d_pts, d_k2, d_i2, n, g, d_tk, d_ts, cap, d_par);
    k_ec_roots<<<grid_for(n, 256, sms), 256, 0, st>>>(d_pts, n, d_par, d_label, d_size);
    k_ec_collect_roots<<<grid_for(n, 256, sms), 256, 0, st>>>(d_size, n, d_roots + 1, reinterpret_cast<int*>(d_roots));
    ctx->launches += 7;
    RCLC_CUDA(cudaGetLastError());
    int2* h_list = reinterpret_cast<int2*>(ctx->h_pin2.as<char>() + 64);
    const size_t first = (size_t)std::min<int64_t>(n, kRootsFirst) + 1;
    RCLC_CUDA(cudaMemcpyAsync(h_list, d_roots, first * 8, cudaMemcpyDeviceToHost, st));
    RCLC_CUDA(cudaStreamSynchronize(st));
    const int n_roots = h_list[0].x;
    if ((size_t)n_roots + 1 > first) {
        RCLC_CUDA(cudaMemcpyAsync(h_list + first, d_roots + first, ((size_t)n_roots + 1 - first) * 8, cudaMemcpyDeviceToHost, st));
        RCLC_CUDA(cudaStreamSynchronize(st));
    }
    const int2* h_roots = h_list + 1;
    std::vector<int2> kept;
    for (int k = 0; k < n_roots; ++k) if (h_roots[k].y >= min_size && h_roots[k].y <= max_size) kept.push_back(h_roots[k]);
    std::sort(kept.begin(), kept.end(), [](const int2& a, const int2& b) { return a.y != b.y ? a.y > b.y : a.x < b.x; });
    // rank of every root (the size array is dead: it becomes the root -> rank table), then of every point. The list goes up
    // from its own page-locked buffer: nobody waits for that copy, and the next operator is free to reuse the read-back buffer.
    k_fill_int<<<grid_for(n, 256, sms), 256, 0, st>>>(d_size, n, -1);
    if (!kept.empty()) {
        RCLC_TRY(ctx->h_pin.reserve(kept.size() * 8 + 64));
        if (!ctx->ev_rank) RCLC_CUDA(cudaEventCreateWithFlags(&ctx->ev_rank, cudaEventDisableTiming));
        RCLC_CUDA(cudaEventSynchronize(ctx->ev_rank));                                        // the previous list has left the buffer
        int2* h_rank = ctx->h_pin.as<int2>();
        for (size_t k = 0; k < kept.size(); ++k) { out->sizes.push_back(kept[k].y); h_rank[k] = make_int2(kept[k].x, (int)k); }
        RCLC_CUDA(cudaMemcpyAsync(d_roots, h_rank, kept.size() * 8, cudaMemcpyHostToDevice, st));
        RCLC_CUDA(cudaEventRecord(ctx->ev_rank, st));
        k_ec_set_rank<<<(unsigned)((kept.size() + 255) / 256), 256, 0, st>>>(d_roots, (int)kept.size(), d_size);
    }
    k_ec_labels<<<grid_for(n, 256, sms), 256, 0, st>>>(d_label, d_size, n, d_label);
    ctx->launches += 3;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

} // namespace rclc

using namespace rclc;

extern "C" int rclc_uniform_sample(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, double leaf_d, int32_t* idx_out, int64_t* n_out)
{
    RCLC_CHECK(ctx && pts && idx_out && n_out && n > 0 && n < (1ll << 31) && leaf_d > 0, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    float4* d_pts = ctx->d_pts.as<float4>();
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, ioff, n, d_pts));
    const int* d_idx = nullptr;
    RCLC_TRY(dev_uniform_sample(ctx, d_pts, n, leaf_d, &d_idx, n_out));
    if (*n_out > 0) RCLC_CUDA(cudaMemcpyAsync(idx_out, d_idx, (size_t)*n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    return RCLC_OK;
}

extern "C" int rclc_euclidean_cluster(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, double tol, int64_t min_size, int64_t max_size,
                                      int32_t* labels_out, int32_t* n_clusters, int32_t* sizes_out)
{
    RCLC_CHECK(ctx && pts && labels_out && n_clusters && n > 0 && n < (1ll << 31) && tol > 0, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    float4* d_pts = ctx->d_pts.as<float4>();
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, ioff, n, d_pts));
    ClusterSet cs;
    RCLC_TRY(dev_euclidean_cluster(ctx, d_pts, n, tol, min_size, max_size, &cs));
    RCLC_CUDA(cudaMemcpyAsync(labels_out, cs.d_label, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_clusters = (int32_t)cs.sizes.size();
    if (sizes_out) for (size_t k = 0; k < cs.sizes.size(); ++k) sizes_out[k] = cs.sizes[k];
    return RCLC_OK;
}
