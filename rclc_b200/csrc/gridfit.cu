// gridfit.cu — K4: reflectance-guided grid fitting on device.
//
// Replaces, for a batch of frames resident in HBM:
//   BOARD_FITTING_COST::Evaluate x 82 targets        include/opt/board_fitting_cost.h:74-262
//   BOARD_FITTING_COST::update_neighbour_points      include/opt/board_fitting_cost.h:264-280 (kd-tree radius search)
//   ceres::Solve as configured                        include/opt/opt_grid_fitting.h:96-103 (trust-region LM)
//   opt_grid_fitting outer re-linearisation loop     include/opt/opt_grid_fitting.h:60-131
//   generate_std_corners / SE2_to_SE3 / calc_transfered_standard_corners   include/utils.h:286-316, 273-284, 47-66
//
// Kernels (one launch covers every frame of a group):
//   k_count / k_scatter / k_rowbox   once per solve: counting sort of the cloud into lattice cells and sub-cells (warp-
//                        aggregated atomics), cell table, sentinels, bounding box of every 32-point row
//   k_move               every later outer iteration: moves the SORTED points in place by the solved SE(2) (float, like
//                        pcl::transformPointCloud) and refreshes the row boxes
//   k_sweep              the hot loop: one warp per (frame, target), gathers the rows its discs can touch, accumulates the
//                        target's 16 sums (16 B per point and pose); its last warps turn the sums into residuals, heuristic
//                        Jacobian, Huber corrector and 3x3 normal equations and run the Ceres LM accept/reject/radius state
//                        machine plus the outer loop bookkeeping (lm_frame_step / lm_advance)
//
// Arithmetic that decides class membership is done exactly as the reference does it wherever the fast fp32 evaluation is
// within its error bound of a decision boundary (double maths on float-stored points, explicit _rn intrinsics so nvcc cannot
// contract a*b+c into an FMA the x86 reference build does not use). Sums of float intensities in double are exact for
// realistic data, so the accumulator tables do not depend on summation order.
#include "gridfit.cuh"

#include <thread>
#include <type_traits>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

namespace rclc {

// ------------------------------------------------------------------------------------------------
// Host geometry
// ------------------------------------------------------------------------------------------------
struct HostGeom {
    GridGeom g;
    std::vector<double2> tgt_xy;
    std::vector<float2> tgt_xyf;
    std::vector<uint8_t> is_row;
};

// include/utils.h:286-316 — same expressions, same order, so the doubles are bit-identical.
static void make_targets(const rclc_config& c, std::vector<double2>& xy, std::vector<uint8_t>& is_row)
{
    const int bx = c.board_width, by = c.board_height;
    const int cxid = bx / 2, cyid = by / 2;
    const double s = c.gride_side_length;
    const double col_shift_y = double(cyid) * s;
    const double col_shift_x = double(cxid - 1) * s;
    for (int cc = 0; cc < bx - 1; ++cc)
        for (int r = 0; r < by; ++r) {
            xy.push_back(make_double2(cc * s - col_shift_x, s / 2.0 + r * s - col_shift_y));
            is_row.push_back(0);
        }
    const double row_shift_x = cxid * s;
    const double row_shift_y = double(cyid - 1) * s;
    for (int r = 0; r < by - 1; ++r)
        for (int cc = 0; cc < bx; ++cc) {
            xy.push_back(make_double2(s / 2.0 + cc * s - row_shift_x, r * s - row_shift_y));
            is_row.push_back(1);
        }
}

static int build_geom(const rclc_config& c, HostGeom& hg)
{
    RCLC_CHECK(c.board_width >= 2 && c.board_height >= 2 && c.gride_side_length > 0, RCLC_ERR_INVALID, "bad board size");
    make_targets(c, hg.tgt_xy, hg.is_row);
    GridGeom& g = hg.g;
    g.nt = (int)hg.tgt_xy.size();
    g.nc = (c.board_width - 1) * (c.board_height - 1);
    RCLC_CHECK(g.nt <= kMaxTargets, RCLC_ERR_UNSUPPORTED, "too many targets");
    g.board_w = c.board_width;
    g.board_h = c.board_height;
    g.side = c.gride_side_length;
    // board_fitting_cost.h:41-59
    const double factor = std::sqrt(2.0);
    const double neighbour_radius = c.neighbour_radius_rate * c.gride_side_length;
    double cost_radius = c.cost_radius_rate * c.gride_side_length;
    double cost_gradient_radius = cost_radius * factor;
    if (!(neighbour_radius > cost_gradient_radius)) {
        cost_gradient_radius = 0.9 * neighbour_radius;
        cost_radius = cost_gradient_radius / factor;
    }
    g.rc2 = cost_radius * cost_radius;
    g.rg2 = cost_gradient_radius * cost_gradient_radius;
    g.rn2f = static_cast<float>(neighbour_radius * neighbour_radius);
    g.huber_a = c.gridfit_huber_a;
    g.function_tol = c.gridfit_function_tol;
    g.max_lm_iters = c.gridfit_max_lm_iters;
    g.max_iter_time = c.max_iter_time;
    // Lattice: all targets sit on sites of the half-square lattice (unit g = side/2); its cells are the coarse sort key.
    g.g = c.gride_side_length / 2.0;
    g.inv_g = 1.0 / g.g;
    // The lattice box keeps a margin of 2 cells around the outermost targets; points outside it are filed in the border cells
    // (fine_of), so the window of an outer target reaches them once they have moved into its neighbourhood (board_fitting_cost.h:273).
    RCLC_CHECK(neighbour_radius < 1.98 * g.g, RCLC_ERR_UNSUPPORTED, "neighbour_radius_rate must be < 0.99 (lattice margin)");
    RCLC_CHECK(c.uniform_sampling_radius == 0.0, RCLC_ERR_UNSUPPORTED,
               "uniform_sampling_radius != 0 is not implemented on device (shipped config uses 0)");
    int imin = INT32_MAX, imax = INT32_MIN, jmin = INT32_MAX, jmax = INT32_MIN;
    std::vector<int> si(g.nt), sj(g.nt);
    for (int t = 0; t < g.nt; ++t) {
        si[t] = (int)std::lround(hg.tgt_xy[t].x * g.inv_g);
        sj[t] = (int)std::lround(hg.tgt_xy[t].y * g.inv_g);
        RCLC_CHECK(std::fabs(hg.tgt_xy[t].x - si[t] * g.g) < 1e-9 && std::fabs(hg.tgt_xy[t].y - sj[t] * g.g) < 1e-9,
                   RCLC_ERR_UNSUPPORTED, "targets are not on the half-square lattice");
        imin = std::min(imin, si[t]); imax = std::max(imax, si[t]);
        jmin = std::min(jmin, sj[t]); jmax = std::max(jmax, sj[t]);
        hg.tgt_xyf.push_back(make_float2((float)hg.tgt_xy[t].x, (float)hg.tgt_xy[t].y));
    }
    g.i0 = imin - 2;
    g.j0 = jmin - 2;
    g.ncx = imax - imin + 4;
    g.ncy = jmax - jmin + 4;
    g.ncells = g.ncx * g.ncy;
    RCLC_CHECK(g.ncells <= kMaxCoarse, RCLC_ERR_UNSUPPORTED, "board too large for the cell table");
    // sub-cells per axis inside a cell (sort key only): the finer, the tighter the row boxes
    g.sub = 4;
    while (g.sub > 1 && g.ncells * g.sub * g.sub > kMaxCells) g.sub >>= 1;
    g.nfine = g.ncells * g.sub * g.sub;
    return RCLC_OK;
}

static PoseCoef host_pose_coef(const double pose[3])
{
    // board_fitting_cost.h:80-82,113 with Eigen's toRotationMatrix for q = (w,0,0,z)
    PoseCoef pc;
    const double theta = (pose[2] * double(M_PI)) / (180.0);
    const double qw = std::cos(theta / 2.0), qz = std::sin(theta / 2.0);
    const double tz = 2.0 * qz, twz = tz * qw, tzz = tz * qz;
    pc.r00 = 1.0 - (0.0 + tzz);
    pc.r01 = 0.0 - twz;
    pc.r10 = 0.0 + twz;
    pc.r11 = 1.0 - (0.0 + tzz);
    pc.tx = pose[0];
    pc.ty = pose[1];
    pc.yaw_deg = pose[2];
    return pc;
}

// ------------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------------
// Sort key of a point: lattice cell (the unit the candidate tables are built on) and, inside it, one of sub x sub
// sub-cells, so that consecutive points of a cell's segment are spatial neighbours. Returns cell * sub^2 + sub-cell.
// A point outside the lattice box (the outermost targets plus two cells: holder, surroundings in the board's plane) is
// kept in the BORDER cell next to it, with its own coordinates: the reference rebuilds its kd-tree on the full cloud in
// every outer iteration (opt_grid_fitting.h:75-76, 107), so a point that starts out of every target's reach can come
// into reach once the accumulated motion exceeds the 2 cm of slack between the box and the neighbour radius. The border
// cell is the right place for it: a target's cell window is its disc widened by the motion since the sort, clamped to
// the grid, so it covers the border cell of every outside point that has moved into the disc; row boxes come from the
// coordinates themselves. -1 only for non-finite coordinates.
__device__ __forceinline__ int fine_of(const GridGeom& g, float x, float y)
{
    // (fp32: a point within rounding of a sub-cell border may fall on either side — every use of the cell tolerates that:
    //  row boxes come from the coordinates themselves, candidate windows have 2 cm of slack)
    if (!(isfinite(x) && isfinite(y))) return -1;
    const float sc = (float)(g.inv_g * g.sub);
    const int fi = min(max(__float2int_rd(x * sc) - g.i0 * g.sub, 0), g.ncx * g.sub - 1);
    const int fj = min(max(__float2int_rd(y * sc) - g.j0 * g.sub, 0), g.ncy * g.sub - 1);
    const int ci = fi / g.sub, cj = fj / g.sub;
    return (cj * g.ncx + ci) * (g.sub * g.sub) + (fj - cj * g.sub) * g.sub + (fi - ci * g.sub);
}

// Is the stored point a kd-tree neighbour of the target? FLANN L2_Simple<float> over (x,y,z), kept when
// dist < radius^2 (board_fitting_cost.h:273); tq = (float)target, the query point.
__device__ __forceinline__ bool is_neighbour(float tqx, float tqy, float rn2f, float x, float y, float z)
{
    const float dx = __fsub_rn(tqx, x), dy = __fsub_rn(tqy, y), dz = __fsub_rn(0.0f, z);
    float r = __fmul_rn(dx, dx);
    r = __fadd_rn(r, __fmul_rn(dy, dy));
    r = __fadd_rn(r, __fmul_rn(dz, dz));
    return r < rn2f;
}

__device__ __forceinline__ void pose_coef_dev(double x, double y, double yaw, PoseCoef& pc)
{
    const double theta = __ddiv_rn(__dmul_rn(yaw, 3.14159265358979323846), 180.0);
    const double half = __ddiv_rn(theta, 2.0);
    const double qw = cos(half), qz = sin(half);
    const double tz = __dmul_rn(2.0, qz), twz = __dmul_rn(tz, qw), tzz = __dmul_rn(tz, qz);
    pc.r00 = __dsub_rn(1.0, tzz);
    pc.r01 = __dsub_rn(0.0, twz);
    pc.r10 = twz;
    pc.r11 = __dsub_rn(1.0, tzz);
    pc.tx = x;
    pc.ty = y;
    pc.yaw_deg = yaw;
}

// Block-wide exclusive scan of data[0..n) in shared memory (n <= kMaxCells). Returns the total.
template <int T>
__device__ int block_exclusive_scan(int* data, int n, int* s_part /* [T] */)
{
    const int tid = threadIdx.x;
    const int per = (n + T - 1) / T;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += data[i];
    s_part[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over T partials
    for (int off = 1; off < T; off <<= 1) {
        int v = (tid >= off) ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    const int total = s_part[T - 1];
    int run = s_part[tid] - sum;
    for (int i = lo; i < hi; ++i) {
        const int v = data[i];
        data[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------------------------------------
// k_init: working copy + control/LM state for a fresh solve (or a fresh binning)
// ------------------------------------------------------------------------------------------------
__device__ void reset_lm_for_outer(LmState& s, FrameCtl& c)
{
    s.core.phase = PH_EVAL_INIT;
    for (int k = 0; k < 3; ++k) { s.core.x[k] = 0; s.core.cand[k] = 0; s.core.best[k] = 0; s.spec[0][k] = 0; }
    pose_coef_dev(0.0, 0.0, 0.0, c.pose[0]);
    c.n_pose = 1;                    // T_se2_wl = {0, 0, 0} (opt_grid_fitting.h:70): nothing to speculate on before the first Jacobian
    c.sweep_active = 1;
}

__global__ void k_init_state(BatchView bv, const double* __restrict__ T_bl /* F*16 or null */)
{
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int nc = bv.geom.ncells, nf = bv.geom.nfine;
    for (int c = tid; c < nf; c += blockDim.x) { bv.hist[(size_t)f * nf + c] = 0; bv.cursor[(size_t)f * nf + c] = 0; }
    for (int c = tid; c < nc; c += blockDim.x) { bv.cellz2[(size_t)f * nc + c] = 0u; bv.cellcnt[(size_t)f * nc + c] = 0; }
    for (int k = tid; k < kSpec * bv.geom.nt * 16; k += blockDim.x) bv.acc[(size_t)f * bv.acc_stride + k] = 0.0;
    for (int k = tid; k < kSpec * bv.geom.nt; k += blockDim.x) bv.tgt_done[(size_t)f * kSpec * bv.geom.nt + k] = 0;
    if (tid < kSpec) bv.slot_done[(size_t)f * kSpec + tid] = 0;
    if (tid == 0) {
        FrameCtl& c = bv.ctl[f];
        LmState& s = bv.lm[f];
        c.rebin = 1;
        c.n_rows = 0;
        c.n_binned = 0;
        c.use_exact = 0;
        c.drift = 0.f;
        bv.sweep_done[f] = 0;
        for (int k = 0; k < 8; ++k) c.move[k] = 0.f;
        c.move[0] = 1.f;
        c.move[5] = 1.f;
        s.outer_iter = 0;
        s.status = 0;
        s.chain = 0;
        s.n_rejected = 0;
        for (int k = 0; k < 16; ++k) {
            s.T_base_laser[k] = (k % 5 == 0) ? 1.f : 0.f;
            s.T_bl[k] = T_bl ? T_bl[(size_t)f * 16 + k] : ((k % 5 == 0) ? 1.0 : 0.0);
        }
        s.rep.outer_iterations = 0; s.rep.status = 0; s.rep.pose_evals = 0; s.rep.lm_iterations = 0;
        s.rep.first_initial_cost = 0; s.rep.last_final_cost = 0;
        s.rep.poses_swept = 0; s.rep.rounds = 0; s.rep.spec_mismatch = 0;
#ifdef RCLC_LAB
        for (int k = 0; k < 8; ++k) s.dbg[k] = 0;
#endif
        for (int k = 0; k < 3; ++k) s.rep.se2_first[k] = 0;
        reset_lm_for_outer(s, c);
        if (bv.n_pts[f] != 0) bv.rebin_list[bv.list_parity * bv.n_frames + atomicAdd(&bv.rebin_cnt[bv.list_parity], 1)] = f;
        if (bv.n_pts[f] == 0) { s.core.phase = PH_DONE; c.sweep_active = 0; c.rebin = 0; s.status = 2; s.rep.status = 2; atomicAdd(bv.n_done, 1); }
    }
}

// ------------------------------------------------------------------------------------------------
// k_count / k_scatter: move (float) + counting sort by lattice cell. Grid-stride over chunks so that the
// (frequent) launches in which no frame needs a re-bin cost only a few hundred early-exit CTAs.
// ------------------------------------------------------------------------------------------------

// A cell's segment of `binned` is padded to a multiple of 32 points with sentinels that are far outside every
// disc for every pose, so the sweep needs no tail predicate and a 32-point row never mixes two cells.
__device__ __forceinline__ int pad32(int v) { return (v + 31) & ~31; }
constexpr float kSentinelXY = 3.0e18f;      // (3e18)^2 * 2 is finite in fp32

// Release / acquire fence at device scope. (__threadfence() is the sequentially-consistent fence: MEMBAR.SC.GPU plus an L1
// invalidate, at the end of every one of thousands of tasks; the last-arriver protocols here only need their own earlier
// writes and atomics ordered before the counter increment, and the reader reads through L2.)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// warp-aggregated atomicAdd(&counter[bin], 1): lanes with the same bin elect a leader; returns this lane's rank
__device__ __forceinline__ int bin_rank(int* counter, int bin, bool valid)
{
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    int rank = 0;
    if (valid) {
        const unsigned peers = __match_any_sync(act, bin);
        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(counter + bin, __popc(peers));
        base = __shfl_sync(peers, base, leader);
        rank = base + __popc(peers & ((1u << lane) - 1u));
    }
    return rank;
}

__device__ __forceinline__ void k_count_frame(const BatchView& bv, int f);
__device__ __forceinline__ void k_scatter_frame(const BatchView& bv, int f);

// warp-aggregated atomicAdd(&counter[bin], 1) without a return value
__device__ __forceinline__ void bin_count(int* counter, int bin, bool valid)
{
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counter + bin, __popc(peers));
    }
}

// Pass 1 of the counting sort: histogram of (cell, sub-cell) bins of the moved points; the last CTA of the frame then
// turns the histogram into bin offsets (cells padded to whole rows), the cell table and the sentinels.
__global__ void __launch_bounds__(kBinThreads) k_count(const __grid_constant__ BatchView bv)
{
    // the list the LM steps of this round will fill was consumed by the previous round: empty it
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) bv.rebin_cnt[bv.list_parity ^ 1] = 0;
    const int n_list = bv.rebin_cnt[bv.list_parity];
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) k_count_frame(bv, bv.rebin_list[bv.list_parity * bv.n_frames + li]);
}

__device__ __forceinline__ void k_count_frame(const BatchView& bv, int f)
{
    FrameCtl& c = bv.ctl[f];
    const int64_t n = bv.n_pts[f];
    const int nf = bv.geom.nfine, nc = bv.geom.ncells, s2 = bv.geom.sub * bv.geom.sub;
    const int tid = threadIdx.x;
    int* hist = bv.hist + (size_t)f * nf;
    {
        const float m00 = c.move[0], m01 = c.move[1], m02 = c.move[2], m03 = c.move[3];
        const float m10 = c.move[4], m11 = c.move[5], m12 = c.move[6], m13 = c.move[7];
        const float4* src = bv.raw + (size_t)f * bv.frame_stride;
        // four independent loads in flight per thread; whole warps stay converged for the aggregated atomics
        const int64_t stride = (int64_t)gridDim.x * kBinThreads, n_up = (n + 31) & ~(int64_t)31;
        for (int64_t idx0 = (int64_t)blockIdx.x * kBinThreads + tid; idx0 < n_up; idx0 += 4 * stride) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t idx = idx0 + u * stride;
                q[u] = idx < n ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t idx = idx0 + u * stride;
                if (idx >= n_up) break;                       // warp-uniform
                int fine = -1;
                if (idx < n) {
                    const float x = affine_row_f(m00, m01, m02, m03, q[u].x, q[u].y, q[u].z);
                    const float y = affine_row_f(m10, m11, m12, m13, q[u].x, q[u].y, q[u].z);
                    fine = fine_of(bv.geom, x, y);
                }
                bin_count(hist, max(fine, 0), fine >= 0);
            }
        }
    }
    // ---- last CTA of the frame: offsets ----
    __shared__ int s_cell[kMaxCoarse];
    __shared__ int s_part[kBinThreads];
    __shared__ int s_last;
    fence_acq_rel_gpu();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&bv.bin_done[f], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) { __syncthreads(); return; }
    fence_acq_rel_gpu();
    if (tid == 0) bv.bin_done[f] = 0;
    int my_total = 0;
    for (int cc = tid; cc < nc; cc += kBinThreads) {
        int tot = 0;
        for (int k = 0; k < s2; ++k) tot += __ldcg(hist + cc * s2 + k);
        s_cell[cc] = pad32(tot);
        my_total += tot;
        bv.cellcnt[(size_t)f * nc + cc] = tot;
    }
    __syncthreads();
    const int n_padded = block_exclusive_scan<kBinThreads>(s_cell, nc, s_part);      // s_cell: first point of every cell
    float4* dstb = bv.binned + (size_t)f * bv.frame_stride;
    int* cursor = bv.cursor + (size_t)f * nf;                                        // becomes: next free slot of every bin
    for (int cc = tid; cc < nc; cc += kBinThreads) {
        int run = s_cell[cc];
        bv.cellstart[(size_t)f * nc + cc] = run;
        for (int k = 0; k < s2; ++k) {
            cursor[cc * s2 + k] = run;
            run += __ldcg(hist + cc * s2 + k);
        }
        for (int j = run; j < s_cell[cc] + pad32(run - s_cell[cc]); ++j)
            dstb[j] = make_float4(kSentinelXY, kSentinelXY, 0.f, 0.f);
    }
    __syncthreads();
    s_part[tid] = my_total;
    __syncthreads();
    for (int off = kBinThreads / 2; off > 0; off >>= 1) {
        if (tid < off) s_part[tid] += s_part[tid + off];
        __syncthreads();
    }
    if (tid == 0) { c.n_rows = n_padded / 32; c.n_binned = s_part[0]; }
    __syncthreads();                 // shared arrays are reused by the next frame of this CTA
}

// Pass 2: move the points for good (pcl::transformPointCloud restated in float), drop each one into the next free slot of
// its bin (warp-aggregated atomics on the bin cursors) together with its z (the kd-tree neighbour test needs it).
__global__ void __launch_bounds__(kBinThreads) k_scatter(const __grid_constant__ BatchView bv)
{
    const int n_list = bv.rebin_cnt[bv.list_parity];
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) k_scatter_frame(bv, bv.rebin_list[bv.list_parity * bv.n_frames + li]);
}

__device__ __forceinline__ void k_scatter_frame(const BatchView& bv, int f)
{
    FrameCtl& c = bv.ctl[f];
    const int64_t n = bv.n_pts[f];
    const int nf = bv.geom.nfine, nc = bv.geom.ncells, s2 = bv.geom.sub * bv.geom.sub;
    const int tid = threadIdx.x;
    const float m00 = c.move[0], m01 = c.move[1], m02 = c.move[2], m03 = c.move[3];
    const float m10 = c.move[4], m11 = c.move[5], m12 = c.move[6], m13 = c.move[7];
    const float4* raw = bv.raw + (size_t)f * bv.frame_stride;
    float4* dstb = bv.binned + (size_t)f * bv.frame_stride;
    int* cursor = bv.cursor + (size_t)f * nf;
    unsigned* cellz2 = bv.cellz2 + (size_t)f * nc;
    bool bad = false;
    // four points per thread and iteration: the loads, then the (warp-aggregated) cursor atomics, then the stores — three
    // memory round trips per four points instead of twelve
    const int64_t stride = (int64_t)gridDim.x * kBinThreads, n_up = (n + 31) & ~(int64_t)31;
    for (int64_t idx0 = (int64_t)blockIdx.x * kBinThreads + tid; idx0 < n_up; idx0 += 4 * stride) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = idx0 + u * stride;
            q[u] = idx < n ? __ldg(raw + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        int fine[4], dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = idx0 + u * stride;
            fine[u] = -1;
            if (idx < n) {
                const float x = affine_row_f(m00, m01, m02, m03, q[u].x, q[u].y, q[u].z);
                const float y = affine_row_f(m10, m11, m12, m13, q[u].x, q[u].y, q[u].z);
                q[u].x = x; q[u].y = y;
                fine[u] = fine_of(bv.geom, x, y);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dst[u] = 0;
            if (idx0 + u * stride < n_up) dst[u] = bin_rank(cursor, max(fine[u], 0), fine[u] >= 0);      // warp-uniform guard
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (fine[u] >= 0) {
                const int cell = fine[u] / s2;
                const float in = q[u].w;
                dstb[dst[u]] = make_float4(q[u].x, q[u].y, in, q[u].z);
                // the fast sweep carries counts inside its fp64 sums: exact for I in {0} U [2^-3, 256)
                bad |= !(in == 0.f || (in >= 0.125f && in < 256.f));
                const unsigned z2 = __float_as_uint(__fmul_rn(q[u].z, q[u].z));          // >= 0: float order == uint order
                if (z2 > cellz2[cell]) atomicMax(&cellz2[cell], z2);
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) atomicOr(&c.use_exact, 1);
}

// Bounding box of every 32-point row (sentinels excluded; an all-sentinel row gets an empty box that no disc touches).
// MOVE = false: right after the initial sort. MOVE = true: the outer loop's pcl::transformPointCloud(pc_iter, T_w1_w0)
// (opt_grid_fitting.h:105-107, float) applied IN PLACE to the sorted points — a rigid motion of a few millimetres keeps
// rows spatially compact, so the sort is not repeated; the frame only remembers a bound on the accumulated drift, by
// which the sweep widens its cell look-up.
constexpr int kBoxThreads = 256;
template <bool MOVE>
__device__ __forceinline__ void rowbox_frame(const BatchView& bv, int f)
{
    FrameCtl& c = bv.ctl[f];
    const int n_rows = c.n_rows;
    const int lane = threadIdx.x & 31;
    float4* pts = bv.binned + (size_t)f * bv.frame_stride;
    float4* box = bv.rowbox + (size_t)f * (bv.frame_stride / 32);
    const float m00 = c.move[0], m01 = c.move[1], m02 = c.move[2], m03 = c.move[3];
    const float m10 = c.move[4], m11 = c.move[5], m12 = c.move[6], m13 = c.move[7];
    for (int r = blockIdx.x * (kBoxThreads / 32) + (threadIdx.x >> 5); r < n_rows; r += gridDim.x * (kBoxThreads / 32)) {
        float4 q = pts[(size_t)r * 32 + lane];
        const bool real = q.x != kSentinelXY;
        if (MOVE && real) {
            const float x = affine_row_f(m00, m01, m02, m03, q.x, q.y, q.w);       // q.w is z; row 2 of an SE(2) leaves it alone
            const float y = affine_row_f(m10, m11, m12, m13, q.x, q.y, q.w);
            q.x = x; q.y = y;
            pts[(size_t)r * 32 + lane] = q;
        }
        float x0 = real ? q.x : 3.0e38f, y0 = real ? q.y : 3.0e38f, x1 = real ? q.x : -3.0e38f, y1 = real ? q.y : -3.0e38f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o));
            y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0) box[r] = make_float4(x0, y0, x1, y1);
    }
    if (MOVE && blockIdx.x == 0 && threadIdx.x == 0) {
        // |p' - p| <= |R - I| |p| + |t| over the lattice the points live in (+ the drift they already have)
        const GridGeom& g = bv.geom;
        const float ex = (float)(fmax(fabs(g.i0 * g.g), fabs((g.i0 + g.ncx) * g.g))), ey = (float)(fmax(fabs(g.j0 * g.g), fabs((g.j0 + g.ncy) * g.g)));
        const float pmax = sqrtf(ex * ex + ey * ey) + c.drift;
        const float dr = sqrtf((m00 - 1.f) * (m00 - 1.f) + m01 * m01 + m10 * m10 + (m11 - 1.f) * (m11 - 1.f));     // Frobenius >= spectral
        c.drift += 1.0001f * (dr * pmax + fabsf(m03) + fabsf(m13)) + 1e-7f;
    }
}

__global__ void __launch_bounds__(kBoxThreads) k_rowbox(const __grid_constant__ BatchView bv)
{
    const int n_list = bv.rebin_cnt[bv.list_parity];
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) rowbox_frame<false>(bv, bv.rebin_list[bv.list_parity * bv.n_frames + li]);
}

// One launch per round: frames whose LM step asked for the next outer iteration are moved in place.
__global__ void __launch_bounds__(kBoxThreads) k_move(const __grid_constant__ BatchView bv)
{
    // the list the LM steps of this round will fill was consumed by the previous round: empty it
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) bv.rebin_cnt[bv.list_parity ^ 1] = 0;
    const int n_list = bv.rebin_cnt[bv.list_parity];
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) rowbox_frame<true>(bv, bv.rebin_list[bv.list_parity * bv.n_frames + li]);
}

// ------------------------------------------------------------------------------------------------
// Residual / Jacobian of one target from its 16 accumulators (board_fitting_cost.h:211-260)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void target_residual(const double* a, bool is_row, double tx, double ty, double cth, double sth,
                                                double& r, double& j0, double& j1, double& j2)
{
    const double C_d = a[0] / a[4], C_u = a[1] / a[5], C_l = a[2] / a[6], C_r = a[3] / a[7];
    const double dC_d = a[8] / a[12], dC_u = a[9] / a[13], dC_l = a[10] / a[14], dC_r = a[11] / a[15];
    const double cost_ud = fabs(C_d - C_u);
    const double cost_lr = fabs(C_r - C_l);
    const double gradient_x = fabs(C_r - dC_r) - fabs(C_l - dC_l);
    const double gradient_y = fabs(C_d - dC_d) - fabs(C_u - dC_u);
    r = is_row ? 1.0 - tanh(cost_ud / 100.0) : 1.0 - tanh(cost_lr / 100.0);
    const double dp1 = __dsub_rn(__dmul_rn(-tx, cth), __dmul_rn(ty, sth));
    const double dp2 = __dsub_rn(__dmul_rn(tx, sth), __dmul_rn(ty, cth));
    j0 = gradient_x;
    j1 = gradient_y;
    j2 = __dadd_rn(__dmul_rn(gradient_x, dp1), __dmul_rn(gradient_y, dp2));
}

__global__ void k_eval_out(BatchView bv)
{
    const int f = blockIdx.x;
    const PoseCoef pc = bv.ctl[f].pose[0];
    const double theta = (pc.yaw_deg * 3.14159265358979323846) / 180.0;
    const double cth = cos(theta), sth = sin(theta);
    for (int t = threadIdx.x; t < bv.geom.nt; t += blockDim.x) {
        double r, j0, j1, j2;
        target_residual(bv.acc + (size_t)f * bv.acc_stride + (size_t)t * 16, bv.tgt_is_row[t] != 0, pc.tx, pc.ty, cth, sth, r, j0, j1, j2);
        double* o = bv.eval_out + ((size_t)f * bv.geom.nt + t) * 4;
        o[0] = r; o[1] = j0; o[2] = j1; o[3] = j2;
    }
}

// ------------------------------------------------------------------------------------------------
// k_lm: Ceres trust-region LM as a resumable state machine (one CTA per frame)
// ------------------------------------------------------------------------------------------------
struct EvalSums { double cost; double A[6]; double b[3]; int res_valid; int jac_valid; };

__device__ void se2_to_se3f_dev(double x, double y, double yaw, float T[16])
{
    PoseCoef pc;
    pose_coef_dev(x, y, yaw, pc);       // same quaternion matrix as utils.h:273-284
    for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.f : 0.f;
    T[0] = (float)pc.r00; T[1] = (float)pc.r01; T[4] = (float)pc.r10; T[5] = (float)pc.r11;
    T[3] = (float)x;
    T[7] = (float)y;
}

__device__ void mat4f_mul_dev(const float* A, const float* B, float* C)
{
    float out[16];
    for (int r = 0; r < 4; ++r)
        for (int cc = 0; cc < 4; ++cc) {
            float s = __fmul_rn(A[r * 4 + 0], B[0 * 4 + cc]);
            s = __fadd_rn(s, __fmul_rn(A[r * 4 + 1], B[1 * 4 + cc]));
            s = __fadd_rn(s, __fmul_rn(A[r * 4 + 2], B[2 * 4 + cc]));
            s = __fadd_rn(s, __fmul_rn(A[r * 4 + 3], B[3 * 4 + cc]));
            out[r * 4 + cc] = s;
        }
    for (int k = 0; k < 16; ++k) C[k] = out[k];
}

// A double whose products, sums and differences are rounded one by one (no multiply-add is fused): the final composition of the
// board frame and the corners (opt_grid_fitting.h:129-131, utils.h:47-66) are written with it, in the oracle's text, and come
// out with the oracle's bits.
struct Xd { double v; };
__device__ __forceinline__ Xd operator*(Xd a, Xd b) { return {__dmul_rn(a.v, b.v)}; }
__device__ __forceinline__ Xd operator+(Xd a, Xd b) { return {__dadd_rn(a.v, b.v)}; }
__device__ __forceinline__ Xd operator-(Xd a, Xd b) { return {__dsub_rn(a.v, b.v)}; }
__device__ __forceinline__ Xd operator-(Xd a) { return {-a.v}; }

__device__ bool inverse4_dev(const double* m_in, double* o)
{
    Xd m[16], inv[16];
    for (int i = 0; i < 16; ++i) m[i].v = m_in[i];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const Xd det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det.v == 0) return false;
    const Xd id = {1.0 / det.v};
    for (int i = 0; i < 16; i++) o[i] = (inv[i] * id).v;
    return true;
}

// Solve (As + diag(l^2)) y = bs for a 3x3 SPD system by Cholesky. Returns false on a non-positive pivot.
__device__ bool chol3_solve(const double* A6 /*00,01,02,11,12,22*/, const double* l, const double* b, double* y)
{
    const double a00 = A6[0] + l[0] * l[0], a01 = A6[1], a02 = A6[2];
    const double a11 = A6[3] + l[1] * l[1], a12 = A6[4], a22 = A6[5] + l[2] * l[2];
    if (!(a00 > 0.0)) return false;
    const double L00 = sqrt(a00), L10 = a01 / L00, L20 = a02 / L00;
    const double d1 = a11 - L10 * L10;
    if (!(d1 > 0.0)) return false;
    const double L11 = sqrt(d1), L21 = (a12 - L20 * L10) / L11;
    const double d2 = a22 - L20 * L20 - L21 * L21;
    if (!(d2 > 0.0)) return false;
    const double L22 = sqrt(d2);
    const double z0 = b[0] / L00;
    const double z1 = (b[1] - L10 * z0) / L11;
    const double z2 = (b[2] - L20 * z0 - L21 * z1) / L22;
    y[2] = z2 / L22;
    y[1] = (z1 - L21 * y[2]) / L11;
    y[0] = (z0 - L10 * y[1] - L20 * y[2]) / L00;
    return isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]);
}

__device__ void accept_eval(LmCore& s, const EvalSums& e)
{
    s.x_cost = e.cost;
    for (int k = 0; k < 6; ++k) s.Au[k] = e.A[k];
    for (int k = 0; k < 3; ++k) s.bu[k] = e.b[k];
    // gradient_max_norm = |x - Plus(x, -g)|_inf  (trust_region_minimizer.cc EvaluateGradientAndJacobian)
    double gm = 0;
    for (int k = 0; k < 3; ++k) {
        const double xp = s.x[k] + (-e.b[k]);
        gm = fmax(gm, fabs(s.x[k] - xp));
    }
    s.gmax = gm;
}

enum LmTerm { T_CONV = 0, T_NOCONV = 1, T_FAIL = 2, T_RUNNING = 3, T_FAIL_INITIAL = 4 };
enum LmVerdict { V_ACCEPTED = 0, V_REJECTED = 1, V_ENDED = 2 };

// ceres::TrustRegionMinimizer, resumable, in two halves.
// lm_judge: consumes the evaluation of the pose the minimiser asked for last (x = 0 at the start of a solve, then s.cand):
// parameter / function tolerance, step quality, accept or reject and the radius update. `term` is set when the solve ends here.
__device__ int lm_judge(const GridGeom& g, LmCore& s, const EvalSums& e, int& term)
{
    term = T_RUNNING;
    if (s.phase == PH_EVAL_INIT) {
        s.n_lm_iterations_this_solve = 0;
        if (!e.res_valid || !e.jac_valid) {
            s.initial_cost = nan("");
            s.min_iter_cost = nan("");
            term = T_FAIL_INITIAL;
            return V_ENDED;
        }
        accept_eval(s, e);
        for (int k = 0; k < 3; ++k) {
            const double cn = (k == 0) ? e.A[0] : (k == 1) ? e.A[3] : e.A[5];
            s.scale[k] = 1.0 / (1.0 + sqrt(cn));                              // jacobi scaling, computed once
        }
        s.initial_cost = s.x_cost;
        s.min_iter_cost = s.x_cost;
        s.minimum_cost = DBL_MAX;
        s.step_is_successful = 1;
        s.iteration = 0;
        s.radius = 1e4;
        s.decrease_factor = 2.0;
        s.reuse_diagonal = 0;
        s.num_consecutive_invalid = 0;
        s.x_norm = 0.0;
        return V_ACCEPTED;
    }
    const double cand_cost = e.res_valid ? e.cost : DBL_MAX;
    double sn = 0;
    for (int k = 0; k < 3; ++k) sn += (s.x[k] - s.cand[k]) * (s.x[k] - s.cand[k]);
    if (sqrt(sn) <= 1e-8 * (s.x_norm + 1e-8)) { term = T_CONV; return V_ENDED; }                              // parameter tolerance
    if (fabs(s.x_cost - cand_cost) <= g.function_tol * s.x_cost) { term = T_CONV; return V_ENDED; }           // function tolerance
    const double rel = (cand_cost >= DBL_MAX) ? -DBL_MAX : (s.x_cost - cand_cost) / s.model_cost_change;
    if (rel > 1e-3) {
        for (int k = 0; k < 3; ++k) s.x[k] = s.cand[k];
        s.x_norm = sqrt(s.x[0] * s.x[0] + s.x[1] * s.x[1] + s.x[2] * s.x[2]);
        if (!e.jac_valid) { term = T_FAIL; return V_ENDED; }
        accept_eval(s, e);
        s.step_is_successful = 1;
        const double q = 2.0 * rel - 1.0;
        s.radius = s.radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
        s.radius = fmin(1e16, s.radius);
        s.decrease_factor = 2.0;
        s.reuse_diagonal = 0;
        s.min_iter_cost = fmin(s.min_iter_cost, s.x_cost);
        return V_ACCEPTED;
    }
    s.min_iter_cost = fmin(s.min_iter_cost, cand_cost);
    s.radius = s.radius / s.decrease_factor;            // StepRejected: the rule does not look at the cost
    s.decrease_factor *= 2.0;
    s.reuse_diagonal = 1;
    s.step_is_successful = 0;
    return V_REJECTED;
}

// lm_propose: from the judged state to the next candidate (FinalizeIterationAndCheckIfMinimizerCanContinue, then
// LevenbergMarquardtStrategy::ComputeStep on the jacobi-scaled normal equations, invalid steps handled like Ceres). Returns
// T_RUNNING with s.cand / s.model_cost_change set, or how the solve ended. Out of line on purpose: the real minimiser and the
// speculating lanes (lm_frame_step) must execute the very same instructions, because their candidates are compared bit for bit.
__device__ __noinline__ int lm_propose(const GridGeom& g, LmCore& s)
{
    for (;;) {
        if (s.step_is_successful && s.x_cost < s.minimum_cost) {
            s.minimum_cost = s.x_cost;
            for (int k = 0; k < 3; ++k) s.best[k] = s.x[k];
        }
        s.n_lm_iterations_this_solve++;
        if (s.iteration >= g.max_lm_iters) return T_NOCONV;
        if (s.step_is_successful && s.gmax <= 1e-10) return T_CONV;
        if (s.radius <= 1e-32) return T_CONV;
        s.iteration++;
        s.step_is_successful = 0;
        double As[6], bs[3], l[3], y[3];
        const double* sc = s.scale;
        As[0] = s.Au[0] * sc[0] * sc[0]; As[1] = s.Au[1] * sc[0] * sc[1]; As[2] = s.Au[2] * sc[0] * sc[2];
        As[3] = s.Au[3] * sc[1] * sc[1]; As[4] = s.Au[4] * sc[1] * sc[2]; As[5] = s.Au[5] * sc[2] * sc[2];
        for (int k = 0; k < 3; ++k) bs[k] = s.bu[k] * sc[k];
        if (!s.reuse_diagonal) {
            s.diag[0] = fmin(fmax(As[0], 1e-6), 1e32);
            s.diag[1] = fmin(fmax(As[3], 1e-6), 1e32);
            s.diag[2] = fmin(fmax(As[5], 1e-6), 1e32);
        }
        for (int k = 0; k < 3; ++k) l[k] = sqrt(s.diag[k] / s.radius);
        bool valid = chol3_solve(As, l, bs, y);
        s.reuse_diagonal = 1;
        double mcc = 0;
        double step[3];
        if (valid) {
            for (int k = 0; k < 3; ++k) step[k] = -y[k];
            // model_cost_change = -(J step)'(r + J step / 2) = -(step'bs + step' As step / 2)
            const double sAs = step[0] * (As[0] * step[0] + As[1] * step[1] + As[2] * step[2]) +
                               step[1] * (As[1] * step[0] + As[3] * step[1] + As[4] * step[2]) +
                               step[2] * (As[2] * step[0] + As[4] * step[1] + As[5] * step[2]);
            mcc = -((step[0] * bs[0] + step[1] * bs[1] + step[2] * bs[2]) + 0.5 * sAs);
            valid = mcc > 0.0;
        }
        if (!valid) {
            if (++s.num_consecutive_invalid >= 5) return T_FAIL;
            s.radius *= 0.5;
            s.min_iter_cost = fmin(s.min_iter_cost, s.x_cost);
            continue;
        }
        s.num_consecutive_invalid = 0;
        s.model_cost_change = mcc;
        for (int k = 0; k < 3; ++k) s.cand[k] = s.x[k] + step[k] * sc[k];
        s.phase = PH_EVAL_CAND;
        return T_RUNNING;
    }
}

// The inner solve has ended: opt_grid_fitting.h:105-125 (move the cloud, decide whether to go on, compose the transforms) and,
// at the end of the outer loop, :129-131 + utils.h:47-66 (the corners).
__device__ void outer_finish(const GridGeom& g, LmState& st, FrameCtl& c, int term, double* corners, int* n_done)
{
    LmCore& s = st.core;
    const bool initial_failure = term == T_FAIL_INITIAL;
    if (initial_failure) term = T_FAIL;
    const double final_cost = fmin(s.initial_cost, s.min_iter_cost);
    double xo[3] = {0.0, 0.0, 0.0};
    if (term != T_FAIL) for (int k = 0; k < 3; ++k) xo[k] = s.best[k];
    float Tw[16];
    se2_to_se3f_dev(xo[0], xo[1], xo[2], Tw);
    const double rate = (s.initial_cost - final_cost) / s.initial_cost;
    st.outer_iter++;
    if (st.outer_iter == 1) {
        st.rep.first_initial_cost = s.initial_cost;
        for (int k = 0; k < 3; ++k) st.rep.se2_first[k] = xo[k];
    }
    st.rep.last_final_cost = final_cost;
    st.rep.lm_iterations += s.n_lm_iterations_this_solve;
    if (term == T_FAIL) { st.status = 1; st.rep.status = 1; }
    if (initial_failure) {
        // x stays (0,0,0) -> identity move -> every remaining outer iteration repeats this failure.
        const int remaining = g.max_iter_time + 1 - st.outer_iter;
        if (remaining > 0) { st.rep.pose_evals += remaining; st.outer_iter += remaining; }
    }
    const bool quit = (fabs(rate) < 1e-6) || (st.outer_iter > g.max_iter_time);
    if (quit) {
        if (rate > 0) mat4f_mul_dev(Tw, st.T_base_laser, st.T_base_laser);
        double Tb[16], Fm[16], Fi[16];
        for (int k = 0; k < 16; ++k) Tb[k] = (double)st.T_base_laser[k];
        for (int r = 0; r < 4; ++r)
            for (int cc = 0; cc < 4; ++cc) {
                Xd v = {0.0};
                for (int k = 0; k < 4; ++k) v = v + Xd{Tb[r * 4 + k]} * Xd{st.T_bl[k * 4 + cc]};
                Fm[r * 4 + cc] = v.v;
            }
        if (!inverse4_dev(Fm, Fi)) { st.status = 3; st.rep.status = 3; for (int k = 0; k < 16; ++k) Fi[k] = nan(""); }
        // utils.h:47-66
        const double sd = g.side;
        const Xd fcx = Xd{double(int((g.board_w - 1) / 2))} * Xd{sd}, fcy = Xd{double(int((g.board_h - 1) / 2))} * Xd{sd};
        int kk = 0;
        for (int row = 0; row < g.board_h - 1; ++row)
            for (int col = 0; col < g.board_w - 1; ++col) {
                const Xd p[4] = {fcx - Xd{double(col)} * Xd{sd}, fcy - Xd{double(row)} * Xd{sd}, {0.0}, {1.0}};
                for (int r = 0; r < 3; ++r) {
                    Xd v = {0.0};
                    for (int q = 0; q < 4; ++q) v = v + Xd{Fi[r * 4 + q]} * p[q];
                    corners[kk * 3 + r] = v.v;
                }
                ++kk;
            }
        st.rep.outer_iterations = st.outer_iter;
        for (int k = 0; k < 16; ++k) st.rep.T_base_laser[k] = st.T_base_laser[k];
        s.phase = PH_DONE;
        c.sweep_active = 0;
        c.n_pose = 0;
        c.rebin = 0;
        atomicAdd(n_done, 1);
    } else {
        mat4f_mul_dev(Tw, st.T_base_laser, st.T_base_laser);
        c.move[0] = Tw[0]; c.move[1] = Tw[1]; c.move[2] = Tw[2]; c.move[3] = Tw[3];
        c.move[4] = Tw[4]; c.move[5] = Tw[5]; c.move[6] = Tw[6]; c.move[7] = Tw[7];
        c.rebin = 1;
        reset_lm_for_outer(st, c);
    }
}

// With the heuristic Jacobian of board_fitting_cost.h:245-260 Ceres rejects most steps (typically six in a row, then one accept,
// then rejections until the function tolerance ends the solve). A rejection leaves x, J and r alone and shrinks the radius by a
// rule that does not look at the cost (radius /= decrease_factor; decrease_factor *= 2), so the poses Ceres will ask for next,
// IF it keeps rejecting, are known before the first of them has been evaluated. One sweep round evaluates a batch of them; the
// real minimiser then consumes the evaluations in order for as long as its state is, bit for bit, the state the speculation
// assumed (lm_frame_step).
//
// How many poses the next round should carry is a performance heuristic only (whatever it answers, the minimiser consumes
// exactly the evaluations Ceres would have made). With initial_trust_region_radius = 1e4 the first accepted step of a solve
// typically comes after six rejections, and later chains end after 1 .. 7 evaluations: aim each batch at the likely end of the
// chain instead of always sweeping kSpec poses (measured on the 20 x 500 k frames of BASELINE configs[1]: 1.44 -> 1.12 poses
// swept per pose consumed).
__device__ __forceinline__ int speculation_depth(int chain, int n_rejected, int cap)
{
    int k;
    if (chain == 0) k = n_rejected < 7 ? 7 - n_rejected : (n_rejected == 7 ? 2 : 1);
    else k = n_rejected < 4 ? 4 - n_rejected : (n_rejected < 7 ? 7 - n_rejected : 1);
    return max(1, min(k, cap));
}

constexpr int kLmThreads = 128;         // k_ms_cost block size

struct LmStage { LmState lm; FrameCtl ctl; double ev[kSpec][12]; LmCore lane_core[kSpec]; };

// The ten LM terms of one target from its 16 accumulators: residual + heuristic Jacobian (board_fitting_cost.h:211-260),
// HuberLoss + Corrector (rho'' <= 0 branch): cost 0.5 rho0, r *= sqrt(rho1), J *= sqrt(rho1); then the products the
// 3x3 normal equations sum. out[0] cost, out[1..6] J'J (00,01,02,11,12,22), out[7..9] J'r, out[10] bad residual,
// out[11] bad Jacobian.
__device__ __forceinline__ void target_terms(const double* a16, bool is_row, const PoseCoef& pc, double cth, double sth, double huber_a,
                                             double (&out)[12])
{
    double r, j0, j1, j2;
    target_residual(a16, is_row, pc.tx, pc.ty, cth, sth, r, j0, j1, j2);
    out[10] = isfinite(r) ? 0.0 : 1.0;
    out[11] = (isfinite(j0) && isfinite(j1) && isfinite(j2)) ? 0.0 : 1.0;
    const double sq = r * r, bb = huber_a * huber_a;
    double rho0, rho1;
    if (sq > bb) { const double rr = sqrt(sq); rho0 = 2.0 * huber_a * rr - bb; rho1 = fmax(DBL_MIN, huber_a / rr); }
    else { rho0 = sq; rho1 = 1.0; }
    const double w = sqrt(rho1);
    r *= w; j0 *= w; j1 *= w; j2 *= w;
    out[0] = 0.5 * rho0;
    out[1] = j0 * j0; out[2] = j0 * j1; out[3] = j0 * j2;
    out[4] = j1 * j1; out[5] = j1 * j2; out[6] = j2 * j2;
    out[7] = j0 * r;  out[8] = j1 * r;  out[9] = j2 * r;
}

template <typename T>
__device__ __forceinline__ void stage_copy16(T* dst, const T* src, int lane, bool from_global)
{
    static_assert(sizeof(T) % 16 == 0, "staged with 16-byte accesses");
    constexpr int n16 = (int)(sizeof(T) / 16);
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    constexpr int per = (n16 + 31) / 32;
    int4 v[per];
#pragma unroll
    for (int u = 0; u < per; ++u) if (lane + 32 * u < n16) v[u] = from_global ? __ldcg(s4 + lane + 32 * u) : s4[lane + 32 * u];
#pragma unroll
    for (int u = 0; u < per; ++u) if (lane + 32 * u < n16) d4[lane + 32 * u] = v[u];
}

// One WARP, once per (frame, pose slot) and round — the warp that writes the slot's last target: sums the slot's per-target LM
// terms into the 3x3 normal equations. Fixed order (lane-strided over the targets, then a butterfly), so a pose's sums do not
// depend on the slot or the round it was evaluated in. All loads are in flight together: one L2 round trip.
__device__ __forceinline__ void slot_sums(const BatchView& bv, int f, int slot, int lane)
{
    const int nt = bv.geom.nt;
    const double* tm = bv.terms + ((size_t)f * kSpec + slot) * 12 * nt;
    double sums[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) sums[k] = 0.0;
    for (int t0 = 0; t0 < nt; t0 += 96) {
        double v[3][12];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int t = t0 + lane + 32 * u;
#pragma unroll
            for (int k = 0; k < 12; ++k) v[u][k] = t < nt ? __ldcg(tm + (size_t)k * nt + t) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int k = 0; k < 12; ++k) sums[k] += v[u][k];
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) sums[k] = warp_sum(sums[k]);
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) if (lane == k) mine = sums[k];
    if (lane < 12) bv.evsum[((size_t)f * kSpec + slot) * 12 + lane] = mine;
}

// One WARP, once per frame and round (the warp that finishes the frame's last pose slot): replays Ceres on the evaluations of
// the round (lane 0), runs the opt_grid_fitting outer-loop bookkeeping, and proposes the poses of the next round — lane k
// computes the candidate Ceres will ask for after k further rejections, all lanes at once. The frame's state is staged through
// shared memory (one LmStage per CTA: all warps of a CTA work on the same frame, and only one warp per frame and launch gets
// here).
__device__ __noinline__ void lm_frame_step(const BatchView& bv, int f, LmStage* stage)
{
    const int lane = threadIdx.x & 31;
    const GridGeom& g = bv.geom;
    LmState& s_lm = stage->lm;
    FrameCtl& s_ctl = stage->ctl;
#ifdef RCLC_LAB
    long long tm[5];
    auto now = [] { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    tm[0] = now();
#endif
    stage_copy16(&s_lm, &bv.lm[f], lane, true);
    stage_copy16(&s_ctl, &bv.ctl[f], lane, true);
    if (lane < kSpec * 12 / 4) {            // evsum[f]: kSpec x 12 doubles, 32 bytes per lane
        const int4* src = reinterpret_cast<const int4*>(bv.evsum + (size_t)f * kSpec * 12) + 2 * lane;
        const int4 a = __ldcg(src), b = __ldcg(src + 1);
        int4* dst = reinterpret_cast<int4*>(&stage->ev[0][0]) + 2 * lane;
        dst[0] = a; dst[1] = b;
    }
    __syncwarp();
    const int n_pose = s_ctl.n_pose;
#ifdef RCLC_LAB
    tm[1] = now();
#endif
    // ---- replay (lane 0) ----
    int need = 0;                   // 0: no proposal wanted (frame done, or next outer iteration); 1: propose from the judged state
    if (lane == 0) {
        LmCore& s = s_lm.core;
        s_lm.rep.rounds++;
        s_lm.rep.poses_swept += n_pose;
        for (int i = 0; i < n_pose; ++i) {
            if (i > 0) {
                // adopt the proposal lane i made in the last round: valid iff the state it assumed is the state we are in
                if (!(s.radius == s_lm.spec_radius[i] && s.decrease_factor == s_lm.spec_factor[i])) { s_lm.rep.spec_mismatch++; need = 1; break; }
                s.n_lm_iterations_this_solve++;             // what lm_propose does on a clean pass (no invalid step, no end)
                s.iteration++;
                s.step_is_successful = 0;
                s.reuse_diagonal = 1;
                s.num_consecutive_invalid = 0;
                s.model_cost_change = s_lm.spec_mcc[i];
                for (int k = 0; k < 3; ++k) s.cand[k] = s_lm.spec[i][k];
            }
            EvalSums e;
            e.cost = stage->ev[i][0];
            for (int k = 0; k < 6; ++k) e.A[k] = stage->ev[i][1 + k];
            for (int k = 0; k < 3; ++k) e.b[k] = stage->ev[i][7 + k];
            e.res_valid = stage->ev[i][10] == 0.0;
            e.jac_valid = e.res_valid && stage->ev[i][11] == 0.0;
            const bool was_init = s.phase == PH_EVAL_INIT;
            if (was_init) s_ctl.rebin = 0;
            s_lm.rep.pose_evals++;
            int term;
            const int verdict = lm_judge(g, s, e, term);
            if (verdict == V_ENDED) {
                outer_finish(g, s_lm, s_ctl, term, bv.corners + (size_t)f * g.nc * 3, bv.n_done);
                need = 0;
                break;
            }
            need = 1;
            if (verdict == V_ACCEPTED) { s_lm.chain = was_init ? 0 : 1; s_lm.n_rejected = 0; break; }       // new J: the rest of the batch is void
            s_lm.n_rejected++;
        }
#ifdef RCLC_LAB
        tm[2] = now();
#endif
        if (need) stage->lane_core[0] = s;
    }
    need = __shfl_sync(0xffffffffu, need, 0);
    __syncwarp();
    // ---- proposals: lane k's copy of the judged state is rejected k more times, then proposes ----
    if (need) {
        const int depth = speculation_depth(s_lm.chain, s_lm.n_rejected, bv.npose_max);
        int term = T_RUNNING;
        bool clean = false;
        double r_in = 0.0, f_in = 0.0;
        if (lane < depth) {
            LmCore& c = stage->lane_core[lane];
            if (lane > 0) {
                c = stage->lane_core[0];
                for (int j = 0; j < lane; ++j) {
                    if (j > 0) { c.n_lm_iterations_this_solve++; c.iteration++; }      // the clean lm_propose pass before rejection j + 1 ...
                    c.radius = c.radius / c.decrease_factor;                           // ... and StepRejected
                    c.decrease_factor *= 2.0;
                }
                // (the first pass, on the judged state, is lane 0's own: counted below once lane 0 has shown to be clean)
                c.n_lm_iterations_this_solve++; c.iteration++;
                c.step_is_successful = 0;
                c.num_consecutive_invalid = 0;
                // (reuse_diagonal stays as judged: after an accepted step every lane derives the same diagonal from the same J'J)
            }
            r_in = c.radius; f_in = c.decrease_factor;
        }
        __syncwarp();
        if (lane < depth) {
            LmCore& c = stage->lane_core[lane];
            const int iter_in = c.iteration;
            term = lm_propose(g, c);
            clean = term == T_RUNNING && c.radius == r_in && c.iteration == iter_in + 1;      // no invalid step inside, no end
        }
        // lane k is usable iff lanes 0 .. k are all clean; lane 0's own result is the real minimiser's whatever it is
        const unsigned dirty = __ballot_sync(0xffffffffu, lane < depth && !clean) | (depth < 32 ? ~0u << depth : 0u);
        const int n_next = max(1, __ffs(dirty) - 1);
        if (lane == 0) {
            LmCore& s = s_lm.core;
            s = stage->lane_core[0];
            if (term != T_RUNNING) {
                outer_finish(g, s_lm, s_ctl, term, bv.corners + (size_t)f * g.nc * 3, bv.n_done);
            } else {
                s_ctl.sweep_active = 1;
                s_ctl.n_pose = n_next;
            }
        }
        __syncwarp();
        if (s_lm.core.phase == PH_EVAL_CAND && lane < s_ctl.n_pose) {
            const LmCore& c = stage->lane_core[lane];
            for (int k = 0; k < 3; ++k) s_lm.spec[lane][k] = c.cand[k];
            s_lm.spec_mcc[lane] = c.model_cost_change;
            s_lm.spec_radius[lane] = r_in;
            s_lm.spec_factor[lane] = f_in;
            pose_coef_dev(c.cand[0], c.cand[1], c.cand[2], s_ctl.pose[lane]);
        }
    }
    __syncwarp();
#ifdef RCLC_LAB
    if (lane == 0) {
        tm[3] = now();
        tm[4] = tm[3];
        for (int k = 0; k < 3; ++k) s_lm.dbg[k] += tm[k + 1] - tm[k];
        s_lm.dbg[4] += 1;
        if (s_lm.dbg[5] == 0) s_lm.dbg[5] = tm[0];          // first LM step of the solve
        s_lm.dbg[6] = tm[4];                                 // last
    }
    __syncwarp();
#endif
    stage_copy16(&bv.lm[f], &s_lm, lane, false);
    stage_copy16(&bv.ctl[f], &s_ctl, lane, false);
    // the next outer iteration starts with the points moved in place (k_move walks this list)
    if (s_ctl.rebin && lane == 0) bv.rebin_list[(bv.list_parity ^ 1) * bv.n_frames + atomicAdd(&bv.rebin_cnt[bv.list_parity ^ 1], 1)] = f;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// k_sweep: the hot loop, a GATHER. One WARP = one target of one frame (blockIdx.y) — or one of `parts` interleaved
// shares of it.
//
//  * the target's discs, pulled back through the trial pose, have centre c = R^T ref - t in the frame the points are
//    stored in; the warp walks the (at most 3 x 3) cells the gradient disc can touch, tests the bounding boxes of
//    their 32-point rows 32 at a time (one box per lane) and collects the rows that can touch the disc;
//  * the collected rows are streamed with one coalesced float4 load per lane and row (512 B per warp instruction,
//    two rows in flight per warp, ~24 warps per SM);
//  * per point: e = p - c, d^2 = |e|^2, and the rotated offsets -(R e) whose signs give the down / right classes —
//    identical, in exact arithmetic, to the reference's p' = R (p + t), ref - p' (board_fitting_cost.h:132-140);
//    two compares, two selects and four sign-xors build multipliers +-1 / +-0 for six DFMAs that accumulate the six
//    nested totals as signed sums (accumulate6). Each addend is I + 2^17, so the COUNT rides in the high bits of the
//    same exact sum (valid for I in {0} U [2^-3, 256) — k_scatter checks; other frames take the all-fp64 body);
//  * exactness: predicates are evaluated in fp32; every lane tracks the minimum distance of its points to any
//    decision boundary (two circles in d^2, two lines). Only when a row brings that minimum inside the rigorous fp32
//    error bound (about 1e-7 m; ~1 % of the rows) the lanes concerned re-evaluate that point exactly like the
//    reference (fp64, no FMA) and correct the sums — so the accumulator tables equal the oracle's bit for bit;
//  * the target's 16 accumulators are written once per warp; the last warp of a frame runs the LM step.
// ------------------------------------------------------------------------------------------------
constexpr double kCntUnit = 131072.0;           // 2^17 > 448 rows per lane x 256
static_assert(kFlushRows * 256.0 <= kCntUnit, "per-lane intensity sum must stay below the count unit");

struct ExactClass { bool in, cost, down, right; };

__device__ __forceinline__ ExactClass classify_exact(float qx, float qy, const PoseCoef& pc, double ref_x, double ref_y,
                                                     double rg2, double rc2)
{
    const double vx = __dadd_rn((double)qx, pc.tx), vy = __dadd_rn((double)qy, pc.ty);      // board_fitting_cost.h:132-133
    const double xt = __dadd_rn(__dmul_rn(pc.r00, vx), __dmul_rn(pc.r01, vy));
    const double yt = __dadd_rn(__dmul_rn(pc.r10, vx), __dmul_rn(pc.r11, vy));
    const double dx = fmax(1e-14, fabs(__dsub_rn(ref_x, xt)));                                // :138-139
    const double dy = fmax(1e-14, fabs(__dsub_rn(ref_y, yt)));
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));                        // :140
    ExactClass e;
    e.in = !(d2 > rg2);                                                                       // :142
    e.cost = d2 < rc2;                                                                        // :151
    e.down = yt > ref_y;                                                                      // :147
    e.right = xt > ref_x;                                                                     // :178
    return e;
}

// expands six nested totals into the reference's 16 accumulators
__device__ __forceinline__ void store_totals(double* o, double tI, double dI, double rI, double cI, double cdI, double crI,
                                             double tN, double dN, double rN, double cN, double cdN, double crN)
{
    o[0] = cdI; o[1] = cI - cdI; o[2] = cI - crI; o[3] = crI;                                  // cost disc: down, up, left, right
    o[4] = cdN; o[5] = cN - cdN; o[6] = cN - crN; o[7] = crN;
    o[8] = dI - cdI; o[9] = (tI - dI) - (cI - cdI); o[10] = (tI - rI) - (cI - crI); o[11] = rI - crI;   // interval annulus
    o[12] = dN - cdN; o[13] = (tN - dN) - (cN - cdN); o[14] = (tN - rN) - (cN - crN); o[15] = rN - crN;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

struct SweepConst {            // fp32 constants of one (target, pose)
    float cx, cy;                         // disc centre in the stored frame (row-box tests)
    float nkx, nky;                       // -(ref - R t): w = R p - k = x' - ref
    float r00, r01, r10, r11;             // rotation
    float rg2, rc2;                       // radii^2
    float E, E2;                          // error bounds: rotated offsets, squared distance
    float tqx, tqy, rn2;                  // (float)target and neighbour radius^2: the kd-tree membership test
};

// fp32 geometry of one point — the SAME intrinsics in the hot loop and in the correction, so both see identical values.
// wx, wy = R p - k = x' - ref_x, y' - ref_y (k = ref - R t, folded into one constant per target): the point is right / down of
// the target when wx / wy > 0; d^2 = |w|^2 (a rotation keeps the norm). Six instructions: two FFMA chains, FMUL, FFMA.
__device__ __forceinline__ void point_geom(const SweepConst& k, float qx, float qy, float& wx, float& wy, float& d2)
{
    wx = __fmaf_rn(k.r00, qx, __fmaf_rn(k.r01, qy, k.nkx));
    wy = __fmaf_rn(k.r10, qx, __fmaf_rn(k.r11, qy, k.nky));
    d2 = __fmaf_rn(wx, wx, __fmul_rn(wy, wy));
}

// One point update. ptxas turns predicated fp64 adds into add + two selects (12 ALU-pipe selects per point), so the
// class is applied as a MULTIPLIER instead: m = +-2^-7 / +-0.0, whose high word is what FSET writes for a float compare
// (1.0f = 0x3f800000 read as the high word of a double is 2^-7) and whose sign is mixed in by ONE LOP3, h ^ (~w & signbit);
// then six DFMAs:
//   a[0] += [in] v          a[1] += [in] s_y v          a[2] += [in] s_x v        (s = -1 when down / right)
//   a[3] += [cost] v        a[4] += [cost] s_y v        a[5] += [cost] s_x v      (all times 2^-7)
// All products and sums are exact, so down = 64 (a0 - a1) etc. are exact too (to_nested_totals).
__device__ __forceinline__ uint32_t fset_le_bits(float a, float b)      // 0x3f800000 when a <= b, else 0
{
    uint32_t r;
    asm("set.le.f32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint32_t fset_lt_bits(float a, float b)
{
    uint32_t r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint32_t sign_mix_bits(uint32_t h, float w)   // h ^ (~w & 0x80000000)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD2;" : "=r"(r) : "r"(h), "r"(__float_as_uint(w)));
    return r;
}
__device__ __forceinline__ void accumulate6(double (&a)[6], float d2, float wx, float wy, float rg2, float rc2, double v, bool member, int zlo)
{
    uint32_t hin = fset_le_bits(d2, rg2), hc = fset_lt_bits(d2, rc2);
    hin = member ? hin : 0u;              // (member is a compile-time true outside the masked body)
    hc = member ? hc : 0u;
    // zlo: a zero the compiler cannot see through (stream_rows), so the six multipliers share ONE low word held in a register
    // instead of six freshly zeroed ones per point
    a[0] = fma(__hiloint2double((int)hin, zlo), v, a[0]);
    a[1] = fma(__hiloint2double((int)sign_mix_bits(hin, wy), zlo), v, a[1]);
    a[2] = fma(__hiloint2double((int)sign_mix_bits(hin, wx), zlo), v, a[2]);
    a[3] = fma(__hiloint2double((int)hc, zlo), v, a[3]);
    a[4] = fma(__hiloint2double((int)sign_mix_bits(hc, wy), zlo), v, a[4]);
    a[5] = fma(__hiloint2double((int)sign_mix_bits(hc, wx), zlo), v, a[5]);
}

// signed sums (times 2^-7) -> the six nested totals (in, in&down, in&right, cost, cost&down, cost&right); exact
__device__ __forceinline__ void to_nested_totals(double (&a)[6])
{
    a[1] = 64.0 * (a[0] - a[1]);
    a[2] = 64.0 * (a[0] - a[2]);
    a[0] = 128.0 * a[0];
    a[4] = 64.0 * (a[3] - a[4]);
    a[5] = 64.0 * (a[3] - a[5]);
    a[3] = 128.0 * a[3];
}

// Everything the cold paths need about the warp's task, in the warp's slice of shared memory.
struct WarpCtx {
    SweepConst k;
    PoseCoef pc;
    double2 ref;
    double rg2, rc2;
};

struct Corr6 { double d[6]; };

// Rare (about 1e-4 of the points in reach): the point is within the fp32 error bound of a decision boundary. Its class
// is re-evaluated with the reference's own arithmetic (fp64, no FMA) and, if it differs from the fp32 class the hot
// loop has just used, the signed sums are corrected — every term is exact, so the corrected sums equal what an
// all-fp64 sweep accumulates.
__device__ __noinline__ Corr6 correct_point(float4 q, const WarpCtx* wc, bool masked)
{
    Corr6 a;
#pragma unroll
    for (int j = 0; j < 6; ++j) a.d[j] = 0.0;
    const SweepConst k = wc->k;
    float wx, wy, d2;
    point_geom(k, q.x, q.y, wx, wy, d2);
    const bool near = fabsf(__fsub_rn(d2, k.rg2)) <= k.E2 || fabsf(__fsub_rn(d2, k.rc2)) <= k.E2 || fabsf(wx) <= k.E || fabsf(wy) <= k.E;
    if (!near) return a;
    if (masked && !is_neighbour(k.tqx, k.tqy, k.rn2, q.x, q.y, q.w)) return a;      // not a kd-tree neighbour: contributes nothing either way
    // what accumulate6 added
    const double f_in = d2 <= k.rg2 ? 1.0 : 0.0, f_c = d2 < k.rc2 ? 1.0 : 0.0;
    const double f_sy = (__float_as_uint(wy) & 0x80000000u) ? 1.0 : -1.0, f_sx = (__float_as_uint(wx) & 0x80000000u) ? 1.0 : -1.0;
    // what the reference adds
    const PoseCoef pc = wc->pc;
    const ExactClass e = classify_exact(q.x, q.y, pc, wc->ref.x, wc->ref.y, wc->rg2, wc->rc2);
    const double e_in = e.in ? 1.0 : 0.0, e_c = (e.in && e.cost) ? 1.0 : 0.0;
    const double e_sy = e.down ? -1.0 : 1.0, e_sx = e.right ? -1.0 : 1.0;
    const double v = ((double)q.z + kCntUnit) * 0.0078125;        // the accumulators carry 2^-7 of the sums
    a.d[0] = (e_in - f_in) * v;
    a.d[1] = (e_in * e_sy - f_in * f_sy) * v;
    a.d[2] = (e_in * e_sx - f_in * f_sx) * v;
    a.d[3] = (e_c - f_c) * v;
    a.d[4] = (e_c * e_sy - f_c * f_sy) * v;
    a.d[5] = (e_c * e_sx - f_c * f_sx) * v;
    return a;
}

// Splits the count off the per-lane signed sums and reduces both over the warp into the warp's shared totals:
// tot[0..5] intensity sums, tot[6..11] counts (in, in&down, in&right, cost, cost&down, cost&right).
__device__ __forceinline__ void reduce_totals(double (&a)[6], double* tot, int lane)
{
    to_nested_totals(a);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        // a = n * 2^17 + S, 0 <= S < 2^17, everything exact
        int n = (int)(a[j] * (1.0 / kCntUnit));
        double S = fma(-(double)n, kCntUnit, a[j]);
        n = __reduce_add_sync(0xffffffffu, n);
        S = warp_sum(S);
        if (lane == 0) { tot[j] += S; tot[6 + j] += (double)n; }
        a[j] = 0.0;
    }
    __syncwarp();
}


// all-fp64 body (frames whose intensities cannot carry the count, or RCLC_SWEEP_EXACT=1): every predicate exactly as
// the reference, plain sums and integer counts. list entries as in stream_rows.
__device__ __noinline__ void stream_rows_exact(const float4* __restrict__ src, const uint32_t* list, int n, const WarpCtx* wc, double* tot)
{
    const int lane = threadIdx.x & 31;
    const PoseCoef pc = wc->pc;
    const double2 ref = wc->ref;
    const double rg2 = wc->rg2, rc2 = wc->rc2;
    double sI[6] = {0, 0, 0, 0, 0, 0};
    int sN[6] = {0, 0, 0, 0, 0, 0};
    const SweepConst k = wc->k;
    for (int i = 0; i < n; ++i) {
        const float4 q = __ldg(src + (size_t)(list[i] << 5) + lane);
        if (q.x == kSentinelXY) continue;
        if (!is_neighbour(k.tqx, k.tqy, k.rn2, q.x, q.y, q.w)) continue;
        const ExactClass c = classify_exact(q.x, q.y, pc, ref.x, ref.y, rg2, rc2);
        if (!c.in) continue;
        const double I = (double)q.z;
        sI[0] += I; sN[0]++;
        if (c.down) { sI[1] += I; sN[1]++; }
        if (c.right) { sI[2] += I; sN[2]++; }
        if (c.cost) {
            sI[3] += I; sN[3]++;
            if (c.down) { sI[4] += I; sN[4]++; }
            if (c.right) { sI[5] += I; sN[5]++; }
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double S = warp_sum(sI[j]);
        const int N = warp_sum(sN[j]);
        if (lane == 0) { tot[j] += S; tot[6 + j] += (double)N; }
    }
    __syncwarp();
}

// Streams the collected rows through a per-warp shared-memory ring of kRingGroups x kGroupRows rows filled with
// cp.async (LDGSTS, 16 B per lane = 512 B per warp instruction, L2-only caching): up to 12 rows are in flight per warp
// without holding registers, and since every lane later reads exactly the 16 bytes it copied itself the ring needs
// no barrier at all. list entries: row index | (bit index of the target in the row's cell candidate list) << 27; the
// caller pads the list with copies of its last entry, so the copy window never needs a bounds check.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool MASKED>
__device__ __forceinline__ void stream_rows(const float4* __restrict__ src, const uint32_t* list, int n, int lane, uint32_t ring,
                                            const SweepConst& k, const WarpCtx* wc, double (&acc)[6], float& mA, float& mB)
{
    const char* lsrc = reinterpret_cast<const char*>(src + lane);
    const uint32_t lring = ring + lane * 16u;
    int zlo;
    asm volatile("mov.b32 %0, 0;" : "=r"(zlo));
    auto one_row = [&](const float4& qq) {
        float wx, wy, d2;
        point_geom(k, qq.x, qq.y, wx, wy, d2);
        const double v = (double)qq.z + kCntUnit;
        mA = fminf(mA, fminf(fabsf(__fsub_rn(d2, k.rg2)), fabsf(__fsub_rn(d2, k.rc2))));
        mB = fminf(mB, fminf(fabsf(wx), fabsf(wy)));
        accumulate6(acc, d2, wx, wy, k.rg2, k.rc2, v, !MASKED || is_neighbour(k.tqx, k.tqy, k.rn2, qq.x, qq.y, qq.w), zlo);
    };
    auto issue_group = [&](int g) {           // rows g*4 .. g*4+3 of the (padded) list -> slot g % kRingGroups
        const uint32_t slot = lring + (uint32_t)(g % kRingGroups) * (kGroupRows * 512u);
        // the group's four list entries with ONE 16-byte load: ptxas pads every shared load that precedes an LDGSTS with three
        // dummy loads, so four scalar loads cost sixteen issue slots per group
        static_assert(kGroupRows == 4, "one 16-byte load per group");
        const uint4 e = *reinterpret_cast<const uint4*>(list + g * kGroupRows);
        cp_async16(slot, lsrc + (uint64_t)e.x * 512u);
        cp_async16(slot + 512u, lsrc + (uint64_t)e.y * 512u);
        cp_async16(slot + 1024u, lsrc + (uint64_t)e.z * 512u);
        cp_async16(slot + 1536u, lsrc + (uint64_t)e.w * 512u);
        cp_async_commit();
    };
    const int ngroups = (n + kGroupRows - 1) / kGroupRows;
#pragma unroll
    for (int g = 0; g < kRingGroups - 1; ++g) issue_group(g);        // beyond the list: harmless copies of its last row
    for (int g = 0; g < ngroups; ++g) {
        issue_group(g + kRingGroups - 1);                             // its slot was read in the previous iteration
        cp_async_wait<kRingGroups - 1>();                             // group g has landed
        const uint32_t slot = lring + (uint32_t)(g % kRingGroups) * (kGroupRows * 512u);
        const int rows = min(kGroupRows, n - g * kGroupRows);
        if (rows == kGroupRows) {
#pragma unroll
            for (int h = 0; h < kGroupRows; ++h) one_row(lds_f4(slot + h * 512u));
        } else {
            for (int h = 0; h < rows; ++h) one_row(lds_f4(slot + h * 512u));
        }
        if (__any_sync(0xffffffffu, mA <= k.E2 || mB <= k.E)) {       // cold: a few % of the 4-row groups
            if (mA <= k.E2 || mB <= k.E) {
                for (int h = 0; h < rows; ++h) {
                    const Corr6 d = correct_point(lds_f4(slot + h * 512u), wc, MASKED);
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc[j] += d.d[j];
                }
            }
            mA = 3.0e38f;
            mB = 3.0e38f;
            __syncwarp();
        }
    }
    cp_async_wait<0>();
}

constexpr int kListPad = kRingGroups * kGroupRows;    // padding for the copy window
constexpr int kListCap = kRowList + kListPad;
constexpr size_t kRingBytes = (size_t)kRingGroups * kGroupRows * 512;

// mode: 0 fast, 1 fast with neighbour mask, 2 all-fp64
__device__ __forceinline__ void drain_list(int mode, const float4* __restrict__ src, uint32_t* list, int& n_list, int lane, uint32_t ring,
                                           const SweepConst& k, const WarpCtx* wc, double (&acc)[6], float& mA, float& mB, double* tot,
                                           int& rows_since_split)
{
    if (n_list == 0) return;
    if (mode != 2 && rows_since_split + n_list > kFlushRows) {        // keep the riding count exact: split it off BEFORE a list that
        reduce_totals(acc, tot, lane);                                // would take a lane past kFlushRows points (a task of a 500 k
        rows_since_split = 0;                                         // frame has ~330 rows: one reduction at its end, not three)
    }
    if (lane < kListPad) list[n_list + lane] = list[n_list - 1];      // padding for the copy window
    __syncwarp();
    if (mode == 2) stream_rows_exact(src, list, n_list, wc, tot);
    else if (mode == 1) stream_rows<true>(src, list, n_list, lane, ring, k, wc, acc, mA, mB);
    else stream_rows<false>(src, list, n_list, lane, ring, k, wc, acc, mA, mB);
    rows_since_split += n_list;
    n_list = 0;
    __syncwarp();
}

constexpr size_t kSweepSmem = (size_t)kSweepWarps * (kRingBytes + kListCap * 4 + 16 * 8 + 12 * 8 + sizeof(WarpCtx)) + sizeof(LmStage);
static_assert(kListPad <= 32, "one lane per padding entry");

// Three CTAs (12 warps) per SM: with 160 registers the hot loop and the cold paths around it keep everything in registers.
// Measured (20 x 500 k points, one pose): 5 CTAs / 96 registers 63.5 us, 4 / 128 59.4 us, 3 / 160 57.3 us — the single-pose
// sweep needs only 11 warps per SM, and the 25 MB of spill traffic per launch that ncu showed at 96 registers is gone.
template <int CTAS>
__device__ __forceinline__ void sweep_body(const BatchView& bv);

__global__ void __launch_bounds__(kSweepThreads, 3) k_sweep(const __grid_constant__ BatchView bv) { sweep_body<3>(bv); }
// The multi-start search reads one frame (L2-resident) under thousands of poses: it is bound by issue slots and L2 latency,
// not by DRAM, and wants more resident warps — the same body at 4 CTAs per SM (128 registers, no spills; 5 CTAs / 96 registers
// spilled 316 bytes per thread into the hot loop and was 4 % slower: 4.27 vs 4.12 ms per 4096 x 200 k search).
__global__ void __launch_bounds__(kSweepThreads, 4) k_sweep_dense(const __grid_constant__ BatchView bv) { sweep_body<4>(bv); }

#ifdef RCLC_SWEEP_TIMING
// lab build only (scripts/sweep_timeline.py): stage clocks of every task of the last sweep
__device__ long long g_sw_dbg[8192 * 8];
#define SW_STAMP(k) do { if (lane == 0) sw_t[k] = clock64(); } while (0)
#else
#define SW_STAMP(k)
#endif

template <int CTAS>
__device__ __forceinline__ void sweep_body(const BatchView& bv)
{
#ifdef RCLC_SWEEP_TIMING
    long long sw_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); sw_t[7] = gt; sw_t[0] = clock64(); }
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int parts = bv.parts;
    // a frame's tasks live in grid column blockIdx.y (the solve: a CTA's warps then share the frame's LM staging area), or — plain
    // sweeps — the tasks of all frames are packed into consecutive warps
    const int per_frame = bv.geom.nt * parts * bv.npose_max;
    const long long flat = (long long)blockIdx.x * kSweepWarps + warp;
    const int f = bv.flat_tasks ? (int)(flat / per_frame) : (int)blockIdx.y;
    const int task = bv.flat_tasks ? (int)(flat - (long long)f * per_frame) : (int)flat;
    if (f >= bv.n_frames) return;
    // tasks of a frame: pose slot (slowest), target, part — the slots a round does not use are the CTAs at the end of the grid
    const int slot = task / (bv.geom.nt * parts), t = (task - slot * bv.geom.nt * parts) / parts, part = task % parts;
    if (slot >= bv.npose_max) return;
    const int fd = bv.fixed_frame >= 0 ? bv.fixed_frame : f;      // frame whose points are read
    // everything the task needs from global memory, requested before the first use
    const FrameCtl& c = bv.ctl[f];
    const int active = c.sweep_active;
    const int n_pose = c.n_pose;
    const PoseCoef pc = c.pose[slot];
    const int use_exact = c.use_exact;
    const float drift = bv.ctl[fd].drift;
    const double2 ref = bv.tgt_xy[t];
    const float2 tq = bv.tgt_xyf[t];
    if (!active || slot >= n_pose) return;

    extern __shared__ __align__(128) unsigned char s_dyn[];
    unsigned char* sp = s_dyn;
    const uint32_t ring = smem_addr(sp + warp * kRingBytes);                     sp += kSweepWarps * kRingBytes;
    uint32_t* list = reinterpret_cast<uint32_t*>(sp) + warp * kListCap;           sp += kSweepWarps * kListCap * 4;
    double* scratch16 = reinterpret_cast<double*>(sp) + warp * 16;                sp += kSweepWarps * 16 * 8;
    double* tot = reinterpret_cast<double*>(sp) + warp * 12;                      sp += kSweepWarps * 12 * 8;
    WarpCtx* wc = reinterpret_cast<WarpCtx*>(sp) + warp;                          sp += kSweepWarps * sizeof(WarpCtx);
    LmStage* stage = reinterpret_cast<LmStage*>(sp);

    const GridGeom& g = bv.geom;
    const bool exact_body = bv.force_exact || use_exact;
    // p' = R (p + t)  =>  ref - p' = -R (p - c),  c = R^T ref - t
    const double cxd = pc.r00 * ref.x + pc.r10 * ref.y - pc.tx, cyd = pc.r01 * ref.x + pc.r11 * ref.y - pc.ty;
    const double rgd = sqrt(g.rg2);
    SweepConst k;
    k.cx = (float)cxd; k.cy = (float)cyd;
    k.nkx = (float)(-(ref.x - (pc.r00 * pc.tx + pc.r01 * pc.ty)));
    k.nky = (float)(-(ref.y - (pc.r10 * pc.tx + pc.r11 * pc.ty)));
    k.r00 = (float)pc.r00; k.r01 = (float)pc.r01; k.r10 = (float)pc.r10; k.r11 = (float)pc.r11;
    k.rg2 = (float)g.rg2; k.rc2 = (float)g.rc2;
    k.tqx = tq.x; k.tqy = tq.y; k.rn2 = g.rn2f;
    {
        // |fp32 offsets - reference's| <= E for every point that can matter (|p - c| <= 2 rg). w = fl(r00 x + fl(r01 y - k)) with
        // r00, r01, k rounded to fp32: |error| <= u (|x| + |y| + |k| + |r01 y - k| + |w|), u = 2^-24, and |x| + |y| <= |c| + 4 rg,
        // |k| <= |ref| + |t|: at most 4.1 u Bc, plus the reference's own (double) rounding of R (p + t) at magnitude Bc
        const double Bc = fabs(cxd) + fabs(cyd) + fabs(pc.tx) + fabs(pc.ty) + fabs(ref.x) + fabs(ref.y) + 0.2;
        k.E = (float)(6.0 * 5.9604644775390625e-8 * Bc + 1e-9);
        k.E2 = 1.05f * (2.9f * (float)rgd * k.E + 3e-7f * k.rg2 + 2.f * k.E * k.E + 1e-9f);
    }
    if (lane == 0) { wc->k = k; wc->pc = pc; wc->ref = ref; wc->rg2 = g.rg2; wc->rc2 = g.rc2; }
    if (lane < 12) tot[lane] = 0.0;
    __syncwarp();
    SW_STAMP(1);                 // control block, pose and target have arrived; constants computed

    // kd-tree neighbour mask (3-D distance < neighbour_radius at pose 0, board_fitting_cost.h:273): it can exclude a point
    // that the pose moves into the gradient disc only if (rg + displacement)^2 + z^2 >= rn^2
    const float rgf = (float)rgd, reach = rgf + 1e-5f;
    const float pmax = (float)(sqrt(cxd * cxd + cyd * cyd) + 2.0 * rgd);
    const float disp = sqrtf(fmaxf(0.f, 2.f - 2.f * k.r00)) * pmax + (float)(fabs(pc.tx) + fabs(pc.ty)) + 1e-5f;

    double acc[6] = {0, 0, 0, 0, 0, 0};
    float mA = 3.0e38f, mB = 3.0e38f;
    int n_list = 0, rows_since_split = 0;
    const float4* src = bv.binned + (size_t)fd * bv.frame_stride;
    const float4* box = bv.rowbox + (size_t)fd * (bv.frame_stride / 32);

    // ---- cells that can hold a point of the gradient disc: the disc widened by how far points have drifted out of the
    //      cells they were sorted into (in-place moves). Window of nci x ncj cells, normally 3 x 3. ----
    const double rsel = rgd + (double)drift + 1e-5;
    const int ci0 = max((int)floor((cxd - rsel) * g.inv_g) - g.i0, 0), ci1 = min((int)floor((cxd + rsel) * g.inv_g) - g.i0, g.ncx - 1);
    const int cj0 = max((int)floor((cyd - rsel) * g.inv_g) - g.j0, 0), cj1 = min((int)floor((cyd + rsel) * g.inv_g) - g.j0, g.ncy - 1);
    const int nci = min(max(ci1 - ci0 + 1, 0), 32), ncj = max(cj1 - cj0 + 1, 0);      // (a lattice wider than 32 cells under > 0.7 m of drift: not a case)
    // kd-tree neighbour test needed? (max z^2 over all cells of the window; reduced once the first cell look-ups are on their way)
    float z2max = 0.f;
    for (int cc = lane; cc < nci * ncj; cc += 32)
        z2max = fmaxf(z2max, __uint_as_float(bv.cellz2[(size_t)fd * g.ncells + (cj0 + cc / nci) * g.ncx + ci0 + cc % nci]));
    int mode = -1;                       // 0 fast, 1 fast with neighbour mask, 2 all-fp64: decided below, once the first cell look-ups are on their way

    // lane c looks up cell c of (a slab of) the window: at most 32 cells per pass, one pass for windows up to 5 x 5
    const int rows_per_pass = nci > 0 ? max(32 / nci, 1) : 1;
    int gbase = 0;                       // running index of 32-row groups over the window: dealt round-robin to the parts
    for (int cjb = cj0; cjb <= cj1 && nci > 0; cjb += rows_per_pass) {
        const int ncj_p = min(rows_per_pass, cj1 - cjb + 1);
        int my_row0 = 0, my_nrows = 0, my_ngr = 0;
        if (lane < nci * ncj_p) {
            const int cell = (cjb + lane / nci) * g.ncx + ci0 + lane % nci;
            const int cnt = bv.cellcnt[(size_t)fd * g.ncells + cell];
            const int start = bv.cellstart[(size_t)fd * g.ncells + cell];
            if (cnt > 0) {
                my_row0 = start >> 5;
                my_nrows = pad32(cnt) >> 5;
                my_ngr = (my_nrows + 31) >> 5;
            }
        }
        if (mode < 0) {
            SW_STAMP(2);         // cell table entries have arrived
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) z2max = fmaxf(z2max, __shfl_xor_sync(0xffffffffu, z2max, o));
            const bool masked = !((rgf + disp) * (rgf + disp) + z2max < g.rn2f * 0.9999f);
            mode = exact_body ? 2 : (masked ? 1 : 0);
        }
        // the window's boxes on their way into L2 (16 B per row): the batches below then wait for L2, not for DRAM
        for (int i = 0; i * 8 < my_nrows; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(box + my_row0 + i * 8));
        // Order of the cells = the four parity classes of (ci, cj), one after the other. A cell is wanted by the two targets
        // on its corners (and grazed by a few more); every one of them reaches it in the same quarter of its own walk, so the
        // second reader finds the rows in L2 instead of DRAM (measured: DRAM reads 239 -> 166 MB per sweep, L2 hit rate 16 -> 36 %).
        const int slot = (lane < nci * ncj_p) ? (((ci0 + lane % nci) & 1) | (((cjb + lane / nci) & 1) << 1)) : 4;
        // exclusive prefix of my_ngr in that order: one scan over four 16-bit fields (a cell has far fewer than 2047 groups)
        unsigned long long inc = (slot < 4) ? (unsigned long long)my_ngr << (16 * slot) : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        const unsigned long long all = __shfl_sync(0xffffffffu, inc, 31);
        int my_pre = (slot < 4) ? (int)((inc >> (16 * slot)) & 0xffffu) - my_ngr : 0, n_groups = 0;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
            const int tot_sl = (int)((all >> (16 * sl)) & 0xffffu);
            if (sl < slot && slot < 4) my_pre += tot_sl;
            n_groups += tot_sl;
        }

        // ---- this part's 32-row groups, four box loads in flight per lane ----
        for (int v0 = (part - gbase % parts + parts) % parts; v0 < n_groups; v0 += 4 * parts) {
            // a batch appends at most 128 rows: make room first, so that nothing of the batch is live across the streaming
            if (n_list > kRowList - 128) drain_list(mode, src, list, n_list, lane, ring, k, wc, acc, mA, mB, tot, rows_since_split);
            float4 b[4];
            int rbase[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = v0 + u * parts;
                const unsigned own = __ballot_sync(0xffffffffu, v >= my_pre && v < my_pre + my_ngr);      // v >= n_groups: nobody
                const int o = own ? __ffs(own) - 1 : 0;
                const int row0 = __shfl_sync(0xffffffffu, my_row0, o), nrows = __shfl_sync(0xffffffffu, my_nrows, o);
                const int j = v - __shfl_sync(0xffffffffu, my_pre, o);
                rbase[u] = row0 + j * 32;
                b[u] = (own && j * 32 + lane < nrows) ? __ldg(box + rbase[u] + lane) : make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float ex = fmaxf(fmaxf(b[u].x - k.cx, k.cx - b[u].z), 0.f), ey = fmaxf(fmaxf(b[u].y - k.cy, k.cy - b[u].w), 0.f);
                const bool live = ex * ex + ey * ey <= reach * reach;          // an empty box (+big, -big) is never live
                const unsigned m = __ballot_sync(0xffffffffu, live);
                if (live) list[n_list + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(rbase[u] + lane);
                n_list += __popc(m);
            }
        }
        gbase += n_groups;
    }
    SW_STAMP(3);                 // last row-box batch tested (streaming of earlier lists included)
    drain_list(mode, src, list, n_list, lane, ring, k, wc, acc, mA, mB, tot, rows_since_split);
    SW_STAMP(4);                 // last list streamed

    // ---- the target's 16 accumulators ----
    if (mode != 2 && rows_since_split > 0) reduce_totals(acc, tot, lane);
    if (lane == 0) store_totals(scratch16, tot[0], tot[1], tot[2], tot[3], tot[4], tot[5], tot[6], tot[7], tot[8], tot[9], tot[10], tot[11]);
    __syncwarp();
    double* acc_t = bv.acc + (size_t)f * bv.acc_stride + ((size_t)slot * g.nt + t) * 16;
    if (lane < 16) {
        const double v = scratch16[lane];
        if (parts > 1) { if (v != 0.0) atomicAdd(acc_t + lane, v); }
        else acc_t[lane] = v;
    }
#ifdef RCLC_SWEEP_TIMING
    if (lane == 0) {
        sw_t[5] = clock64();
        const int w = bv.flat_tasks ? (int)flat : (int)((blockIdx.y * gridDim.x + blockIdx.x) * kSweepWarps + warp);
        if (w < 8192) for (int q = 0; q < 8; ++q) g_sw_dbg[w * 8 + q] = sw_t[q];
    }
#endif
    if (!bv.fuse_lm) return;
    // ---- fused LM step, part 1: the warp that adds the (pose, target)'s last share turns its 16 accumulators into the LM terms
    //      (in parallel over the targets, overlapped with the tasks still sweeping) and clears them for the next sweep ----
    fence_acq_rel_gpu();
    int last = 1;
    if (parts > 1) {
        int* td = bv.tgt_done + ((size_t)f * kSpec + slot) * g.nt + t;
        if (lane == 0) last = (atomicAdd(td, 1) == parts - 1);
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return;
        fence_acq_rel_gpu();
        if (lane == 0) *td = 0;
    }
    {
        double a16[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) a16[q] = __ldcg(acc_t + q);               // same addresses in every lane: one transaction each
        const PoseCoef pcc = wc->pc;
        const double theta = (pcc.yaw_deg * 3.14159265358979323846) / 180.0;
        const double cth = cos(theta), sth = sin(theta);
        double tt[12];
        target_terms(a16, bv.tgt_is_row[t] != 0, pcc, cth, sth, g.huber_a, tt);
        double* tm = bv.terms + ((size_t)f * kSpec + slot) * 12 * g.nt + t;
        double mine = 0.0;
#pragma unroll
        for (int q = 0; q < 12; ++q) if (lane == q) mine = tt[q];
        if (lane < 12) tm[(size_t)lane * g.nt] = mine;
        if (lane >= 16) acc_t[lane - 16] = 0.0;
    }
    // ---- part 2: the warp that writes a pose slot's last target sums the slot's terms (all lanes, one L2 round trip) ----
    fence_acq_rel_gpu();
    last = 0;
    if (lane == 0) last = (atomicAdd(&bv.slot_done[(size_t)f * kSpec + slot], 1) == g.nt - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    fence_acq_rel_gpu();
    if (lane == 0) bv.slot_done[(size_t)f * kSpec + slot] = 0;
    slot_sums(bv, f, slot, lane);
    // ---- part 3: the warp that finishes the frame's last pose slot runs the LM state machine ----
    fence_acq_rel_gpu();
    last = 0;
    if (lane == 0) last = (atomicAdd(&bv.sweep_done[f], 1) == n_pose - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    fence_acq_rel_gpu();
    if (lane == 0) bv.sweep_done[f] = 0;
    lm_frame_step(bv, f, stage);
}

#include "sweep_stream.cuh"

__global__ void k_set_pose(BatchView bv, const PoseCoef* __restrict__ poses)
{
    const int f = blockIdx.x;
    for (int k = threadIdx.x; k < bv.geom.nt * 16; k += blockDim.x) bv.acc[(size_t)f * bv.acc_stride + k] = 0.0;
    if (threadIdx.x == 0) {
        bv.ctl[f].pose[0] = poses[f];
        bv.ctl[f].n_pose = 1;
        bv.ctl[f].sweep_active = 1;
        bv.ctl[f].rebin = 0;
    }
}


// ------------------------------------------------------------------------------------------------
// K5: multi-start grid-pose search — P poses x one frame's points -> robustified cost per pose.
// Not in the reference: cost(p) = sum_k 0.5 * rho_Huber(r_k(p)^2) with r_k from
// BOARD_FITTING_COST::Evaluate, i.e. what ceres::Problem::Evaluate would report at pose p.
// ------------------------------------------------------------------------------------------------
__global__ void k_ms_setup(BatchView bv, FrameCtl* ctl_ms, const PoseCoef* __restrict__ poses, int P, int frame)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    FrameCtl& c = ctl_ms[p];
    c.sweep_active = 1;
    c.rebin = 0;
    c.n_rows = bv.ctl[frame].n_rows;
    c.n_binned = bv.ctl[frame].n_binned;
    c.use_exact = bv.ctl[frame].use_exact;
    c.drift = bv.ctl[frame].drift;
    for (int k = 0; k < 8; ++k) c.move[k] = 0.f;
    c.n_pose = 1;
    c.pose[0] = poses[p];
}

__global__ void __launch_bounds__(kLmThreads) k_ms_cost(BatchView bv, double* __restrict__ cost_out)
{
    const int p = blockIdx.x;
    const GridGeom& g = bv.geom;
    const PoseCoef pc = bv.ctl[p].pose[0];
    const double theta = (pc.yaw_deg * 3.14159265358979323846) / 180.0;
    const double cth = cos(theta), sth = sin(theta);
    __shared__ double s_red[kLmThreads / 32][2];
    double cost = 0, bad = 0;
    const double* acc = bv.acc + (size_t)p * bv.acc_stride;
    for (int t = threadIdx.x; t < g.nt; t += kLmThreads) {
        double r, j0, j1, j2;
        target_residual(acc + (size_t)t * 16, bv.tgt_is_row[t] != 0, pc.tx, pc.ty, cth, sth, r, j0, j1, j2);
        if (!isfinite(r)) bad = 1.0;
        const double sq = r * r, bb = g.huber_a * g.huber_a;
        cost += 0.5 * ((sq > bb) ? (2.0 * g.huber_a * sqrt(sq) - bb) : sq);
    }
    cost = warp_sum(cost);
    bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = cost; s_red[threadIdx.x >> 5][1] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0, b = 0;
        for (int w = 0; w < kLmThreads / 32; ++w) { c += s_red[w][0]; b += s_red[w][1]; }
        cost_out[p] = (b != 0.0) ? nan("") : c;
    }
}

// ------------------------------------------------------------------------------------------------
// Host side of the batch
// ------------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

constexpr int kRoundsPerGraph = 12;      // rounds per graph launch: a solve takes about three rounds per outer iteration

// The binning kernels walk the re-bin list with gridDim.y CTA columns: most rounds have nothing (or one frame) to
// move and must cost next to nothing, while one frame alone still gets hundreds of CTAs.
constexpr int kBinGridY = 2;

// A contiguous range of frames of the batch, run as a unit on its own stream.
struct GroupRun {
    BatchView v;
    int f0 = 0, nf = 0;
    int round = 0;            // parity of the re-bin list the next round consumes
    cudaStream_t st = nullptr;
};

static BatchView slice_view(const BatchView& v, int f0, int nf, int group)
{
    BatchView s = v;
    const size_t nt = v.geom.nt, nc = v.geom.ncells, nfine = v.geom.nfine, ncorn = v.geom.nc;
    s.n_frames = nf;
    s.n_pts = v.n_pts + f0;
    s.raw = v.raw + (size_t)f0 * v.frame_stride;
    s.binned = v.binned + (size_t)f0 * v.frame_stride;
    s.hist = v.hist + (size_t)f0 * nfine;
    s.cursor = v.cursor + (size_t)f0 * nfine;
    s.cellstart = v.cellstart + (size_t)f0 * nc;
    s.cellcnt = v.cellcnt + (size_t)f0 * nc;
    s.cellz2 = v.cellz2 + (size_t)f0 * nc;
    s.rowbox = v.rowbox + (size_t)f0 * (v.frame_stride / 32);
    s.acc = v.acc + (size_t)f0 * v.acc_stride;
    s.ctl = v.ctl + f0;
    s.lm = v.lm + f0;
    s.bin_done = v.bin_done + f0;
    s.terms = v.terms + (size_t)f0 * kSpec * nt * 12;
    s.tgt_done = v.tgt_done + (size_t)f0 * kSpec * nt;
    s.slot_done = v.slot_done + (size_t)f0 * kSpec;
    s.evsum = v.evsum + (size_t)f0 * kSpec * 12;
    s.sweep_done = v.sweep_done + f0;
    s.eval_out = v.eval_out + (size_t)f0 * nt * 4;
    s.corners = v.corners + (size_t)f0 * ncorn * 3;
    s.rebin_list = v.rebin_list + 2 * f0;          // [2][nf] inside the batch's [2 * F]
    s.rebin_cnt = v.rebin_cnt + 2 * group;
    s.n_done = v.n_done + group;                   // finished frames are counted per group: the groups advance independently
    return s;
}

static dim3 bin_grid(int64_t max_pts, int nf)
{
    const int64_t ctas = (max_pts + kBinThreads * 4 - 1) / (kBinThreads * 4);
    return dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>(ctas, 512)), (unsigned)std::min(kBinGridY, nf));
}

// the one counting sort of a solve (or of rclc_batch_bin): count + offsets, scatter, row boxes
static int launch_rebin(rclc_batch* b, GroupRun& gr)
{
    gr.v.list_parity = gr.round & 1;
    k_count<<<bin_grid(b->max_pts, gr.nf), kBinThreads, 0, gr.st>>>(gr.v);
    k_scatter<<<bin_grid(b->max_pts, gr.nf), kBinThreads, 0, gr.st>>>(gr.v);
    const int64_t rows = gr.v.frame_stride / 32;
    k_rowbox<<<dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, 256)), (unsigned)std::min(kBinGridY, gr.nf)), kBoxThreads, 0, gr.st>>>(gr.v);
    b->ctx->launches += 3;
    gr.round++;              // the sweep that follows keeps this round's parity (its LM steps fill the other list)
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

// rounds after the sort: frames that start their next outer iteration are moved in place
static int launch_move(rclc_batch* b, GroupRun& gr)
{
    gr.v.list_parity = gr.round & 1;
    const int64_t rows = gr.v.frame_stride / 32;
    // (most rounds move nothing: a small grid keeps the idle launch cheap; a real move still gets ~2400 warps)
    k_move<<<dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, b->ctx->sm_count)), (unsigned)std::min(kBinGridY, gr.nf)), kBoxThreads, 0, gr.st>>>(gr.v);
    b->ctx->launches += 1;
    gr.round++;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

// One warp per (pose slot, target, part). Tasks are latency-heavy at both ends (cell and box look-ups, the final reduction), so
// fewer and longer tasks win: measured on B200 the best split of a plain sweep is parts = 1 from 20 active 8x6 frames up and 8
// for a single one (profiles/r01_sweep_history.md).
// Inside the solve (fused LM step) a finer split is better: a frame's LM step starts when its last task ends, and short
// tasks let the first frames' LM steps overlap the sweeping of the later ones. Most rounds of a solve carry kSpec poses per frame.
static int sweep_parts(const rclc_batch* b, int n_frames, bool fused)
{
    const double slots = (double)b->ctx->sm_count * 12.0;              // resident warps (3 CTAs of 4 warps per SM)
    const double tasks = (double)std::max(n_frames, 1) * b->v.geom.nt * (fused ? 0.75 * kSpec : 1.0);
    if (fused) return (int)std::max(1.0, std::min(8.0, std::ceil(2.0 * slots / tasks)));
    return (int)std::max(1.0, std::min(8.0, std::floor(0.9 * slots / tasks + 0.5)));
}

// k_sweep_stream: CTAs per (frame, pose slot), each a contiguous share of the frame's rows. A plain sweep fills the machine once
// (3 CTAs per SM); inside the solve most pose slots of a round are in use and the groups' launches overlap, so a launch asks for
// about twice its share.
static int stream_ctas(const rclc_batch* b, int n_frames, bool fused)
{
    const double slots = (double)b->ctx->sm_count * 3.0;
    const double units = (double)std::max(n_frames, 1) * (fused ? 0.75 * kSpec : 1.0);
    const int64_t rows = b->v.frame_stride / 32;
    const double k = fused ? std::ceil(2.0 * slots / units) : std::floor(slots / units);
    return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)k, std::max<int64_t>(1, rows / 64)));
}

static int launch_sweep_view(rclc_batch* b, BatchView& v, int n_frames, cudaStream_t st, int frames_in_flight)
{
    if (v.fixed_frame < 0 && !v.force_exact && b->stream_sweep) {
        v.npose_max = v.fuse_lm ? kSpec : 1;
        const int k = stream_ctas(b, frames_in_flight, v.fuse_lm != 0);
        k_sweep_stream<<<dim3((unsigned)k, (unsigned)n_frames, (unsigned)v.npose_max), kStThreads, st_smem_bytes(v.geom.nt, v.geom.ncells), st>>>(v);
        b->ctx->launches += 1;
        RCLC_CUDA(cudaGetLastError());
        return RCLC_OK;
    }
    v.parts = sweep_parts(b, frames_in_flight, v.fuse_lm != 0);
    v.npose_max = v.fuse_lm ? kSpec : 1;
    v.flat_tasks = v.fuse_lm ? 0 : 1;
    v.n_frames = n_frames;
    const int per_frame = v.geom.nt * v.parts * v.npose_max;
    dim3 grid((unsigned)((per_frame + kSweepWarps - 1) / kSweepWarps), (unsigned)n_frames);
    if (v.flat_tasks) grid = dim3((unsigned)(((long long)per_frame * n_frames + kSweepWarps - 1) / kSweepWarps), 1u);
    if (v.fixed_frame >= 0) k_sweep_dense<<<grid, kSweepThreads, kSweepSmem, st>>>(v);
    else k_sweep<<<grid, kSweepThreads, kSweepSmem, st>>>(v);
    b->ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

static GroupRun whole_batch(rclc_batch* b)
{
    GroupRun gr;
    gr.v = b->v;
    gr.f0 = 0; gr.nf = b->n_frames; gr.round = b->round; gr.st = b->ctx->stream;
    return gr;
}

static int launch_sweep(rclc_batch* b, bool fuse_lm)
{
    BatchView v = b->v;
    v.fuse_lm = fuse_lm ? 1 : 0;
    return launch_sweep_view(b, v, b->n_frames, b->ctx->stream, b->n_frames);
}

static int launch_init(rclc_batch* b, GroupRun& gr, const double* d_Tbl)
{
    RCLC_CUDA(cudaMemsetAsync(gr.v.rebin_cnt, 0, 2 * sizeof(int), gr.st));
    gr.v.list_parity = gr.round & 1;
    k_init_state<<<gr.nf, 256, 0, gr.st>>>(gr.v, d_Tbl ? d_Tbl + (size_t)gr.f0 * 16 : nullptr);
    b->ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    return RCLC_OK;
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_generate_targets(const rclc_config* cfg, double* targets_xy, int32_t* is_row, int32_t* n_targets)
{
    RCLC_CHECK(cfg && targets_xy && is_row && n_targets, RCLC_ERR_INVALID, "null argument");
    std::vector<double2> xy;
    std::vector<uint8_t> ir;
    make_targets(*cfg, xy, ir);
    for (size_t k = 0; k < xy.size(); ++k) { targets_xy[2 * k] = xy[k].x; targets_xy[2 * k + 1] = xy[k].y; is_row[k] = ir[k]; }
    *n_targets = (int32_t)xy.size();
    return RCLC_OK;
}

int rclc_batch_create(rclc_ctx* ctx, const rclc_config* cfg, int32_t n_frames, int64_t max_pts, rclc_batch** out)
{
    RCLC_CHECK(ctx && cfg && out, RCLC_ERR_INVALID, "null argument");
    RCLC_CHECK(n_frames > 0 && n_frames <= 65535 && max_pts > 0 && max_pts < (1ll << 30), RCLC_ERR_INVALID, "bad batch shape");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    HostGeom hg;
    RCLC_TRY(build_geom(*cfg, hg));
    rclc_batch* b = new rclc_batch();
    b->ctx = ctx;
    b->cfg = *cfg;
    b->n_frames = n_frames;
    b->max_pts = max_pts;
    b->h_npts.assign(n_frames, 0);
    BatchView& v = b->v;
    v.geom = hg.g;
    v.n_frames = n_frames;
    v.fixed_frame = -1;
    v.fuse_lm = 0;
    v.force_exact = 0;
    v.npose_max = 1;
    // every cell's segment of `binned` is padded to a multiple of 32 points
    v.frame_stride = (int64_t)align256(((size_t)max_pts + 32 * (size_t)hg.g.ncells) * 16) / 16;
    v.parts = 1;
    RCLC_CUDA(cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
    RCLC_CUDA(cudaFuncSetAttribute(k_sweep_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
    // (the streaming kernel keeps a per-target table and the LM staging area in shared memory: boards beyond ~400 targets only get the
    //  gather kernel — rclc_batch_set_stream_sweep then reports RCLC_ERR_UNSUPPORTED)
    if (st_smem_bytes(hg.g.nt, hg.g.ncells) <= 72 * 1024)
        RCLC_CUDA(cudaFuncSetAttribute(k_sweep_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem_bytes(hg.g.nt, hg.g.ncells)));
    const size_t F = n_frames, nt = hg.g.nt, ncell = hg.g.ncells, nfine = hg.g.nfine;
    auto fail = [&](const char* what) { set_last_error(what); rclc_batch_destroy(b); return RCLC_ERR_CUDA; };
    const size_t big = F * (size_t)v.frame_stride * 16;
    if (cudaMalloc(&b->d_pristine, big) != cudaSuccess) return fail("cudaMalloc pristine");
    v.raw = b->d_pristine;
    if (cudaMalloc(&v.binned, big) != cudaSuccess) return fail("cudaMalloc binned");
    // small arrays in one block
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
    const size_t o_npts = take(F * 8), o_hist = take(F * nfine * 4), o_cur = take(F * nfine * 4), o_cmask = take(F * ncell * 4);
    const size_t o_cstart = take(F * ncell * 4), o_ccnt = take(F * ncell * 4), o_box = take(F * (size_t)(v.frame_stride / 32) * 16);
    v.acc_stride = (int64_t)kSpec * nt * 16;
    const size_t o_acc = take(F * (size_t)v.acc_stride * 8), o_ctl = take(F * sizeof(FrameCtl));
    const size_t o_lm = take(F * sizeof(LmState)), o_done = take(256), o_sdone = take(F * 4), o_bdone = take(F * 4), o_rlist = take(2 * F * 4), o_rcnt = take(256), o_terms = take(F * kSpec * nt * 12 * 8), o_tdone = take(F * kSpec * nt * 4), o_sldone = take(F * kSpec * 4), o_evsum = take(F * kSpec * 12 * 8), o_txy = take(nt * 16), o_txyf = take(nt * 8);
    const size_t o_row = take(nt), o_eval = take(F * nt * 4 * 8);
    const size_t o_corn = take(F * (size_t)hg.g.nc * 3 * 8), o_gmm = take(F * 64 * 8 + 4096 * 8 * 8);
    if (cudaMalloc(&b->d_block, off) != cudaSuccess) return fail("cudaMalloc block");
    char* base = static_cast<char*>(b->d_block);
    cudaMemsetAsync(base, 0, off, ctx->stream);
    b->d_npts = reinterpret_cast<int64_t*>(base + o_npts);
    v.n_pts = b->d_npts;
    v.hist = reinterpret_cast<int*>(base + o_hist);
    v.cursor = reinterpret_cast<int*>(base + o_cur);
    v.cellz2 = reinterpret_cast<unsigned*>(base + o_cmask);
    v.cellstart = reinterpret_cast<int*>(base + o_cstart);
    v.cellcnt = reinterpret_cast<int*>(base + o_ccnt);
    v.rowbox = reinterpret_cast<float4*>(base + o_box);
    v.acc = reinterpret_cast<double*>(base + o_acc);
    v.ctl = reinterpret_cast<FrameCtl*>(base + o_ctl);
    v.lm = reinterpret_cast<LmState*>(base + o_lm);
    v.n_done = reinterpret_cast<int*>(base + o_done);
    v.sweep_done = reinterpret_cast<int*>(base + o_sdone);
    v.bin_done = reinterpret_cast<int*>(base + o_bdone);
    v.terms = reinterpret_cast<double*>(base + o_terms);
    v.tgt_done = reinterpret_cast<int*>(base + o_tdone);
    v.slot_done = reinterpret_cast<int*>(base + o_sldone);
    v.evsum = reinterpret_cast<double*>(base + o_evsum);
    v.rebin_list = reinterpret_cast<int*>(base + o_rlist);
    v.rebin_cnt = reinterpret_cast<int*>(base + o_rcnt);
    v.list_parity = 0;
    v.tgt_xy = reinterpret_cast<double2*>(base + o_txy);
    v.tgt_xyf = reinterpret_cast<float2*>(base + o_txyf);
    v.tgt_is_row = reinterpret_cast<uint8_t*>(base + o_row);
    v.eval_out = reinterpret_cast<double*>(base + o_eval);
    v.corners = reinterpret_cast<double*>(base + o_corn);
    b->d_gmm = reinterpret_cast<double*>(base + o_gmm);
    cudaMemcpyAsync(base + o_txy, hg.tgt_xy.data(), nt * 16, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(base + o_txyf, hg.tgt_xyf.data(), nt * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(base + o_row, hg.is_row.data(), nt, cudaMemcpyHostToDevice, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail("batch setup copies");
    *out = b;
    return RCLC_OK;
}

int rclc_batch_destroy(rclc_batch* b)
{
    if (!b) return RCLC_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (cudaStream_t st : b->ctx->aux_streams) cudaStreamSynchronize(st);
    if (b->d_pristine) cudaFree(b->d_pristine);
    if (b->v.binned) cudaFree(b->v.binned);
    if (b->d_block) cudaFree(b->d_block);
    for (cudaGraphExec_t ge : b->round_graphs) cudaGraphExecDestroy(ge);
    delete b;
    return RCLC_OK;
}

// The solve runs the batch as G groups of consecutive frames on G streams (ctx->stream and ctx->aux_streams). A frame is
// uploaded on the stream of ITS group, so that a solve which follows the uploads starts on the first group as soon as that
// group's frames have arrived, while the copy engine is still busy with the later groups.
static int solve_groups(const rclc_batch* b)
{
    // 0 = automatic. Large frames (the 500 k-point workload): 20 groups, i.e. a frame starts solving when IT has arrived over PCIe
    // (end-to-end step 11.9 -> 11.0 ms; resident steps are indifferent between 10 and 20). Board-sized frames of a calibration
    // (tens of thousands of points) are latency-bound from the first round on: fewer groups, fewer launches (grid fit of 20 scans
    // 3.3 ms with 10 groups, 4.1 ms with 20).
    int G = b->ctx->solve_groups;
    if (G <= 0) G = b->max_pts >= 200000 ? 20 : 10;
    return std::max(1, std::min(G, b->n_frames));
}
static int ensure_group_streams(rclc_ctx* ctx, int G)
{
    if (!ctx->ev_sync) RCLC_CUDA(cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming));
    while ((int)ctx->aux_streams.size() < G - 1) {
        cudaStream_t st;
        RCLC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->aux_streams.push_back(st);
    }
    return RCLC_OK;
}
static int frame_stream(rclc_batch* b, int32_t frame, cudaStream_t* st)
{
    const int G = solve_groups(b);
    RCLC_TRY(ensure_group_streams(b->ctx, G));
    int gi = 0;
    while (gi + 1 < G && (int)((int64_t)b->n_frames * (gi + 1) / G) <= frame) ++gi;      // group gi holds frames [F gi / G, F (gi + 1) / G)
    *st = gi == 0 ? b->ctx->stream : b->ctx->aux_streams[(size_t)gi - 1];
    if (gi > 0) b->uploads_pending = true;
    return RCLC_OK;
}
// ctx->stream waits for the uploads that went to the other group streams (every entry point that reads the frames on ctx->stream)
static int join_uploads(rclc_batch* b)
{
    if (!b->uploads_pending) return RCLC_OK;
    rclc_ctx* ctx = b->ctx;
    for (cudaStream_t st : ctx->aux_streams) {
        RCLC_CUDA(cudaEventRecord(ctx->ev_sync, st));
        RCLC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_sync, 0));
    }
    b->uploads_pending = false;
    return RCLC_OK;
}
} // extern "C"
int rclc::batch_join_uploads(rclc_batch* b) { return join_uploads(b); }
extern "C" {

static int set_npts(rclc_batch* b, int32_t frame, int64_t n, cudaStream_t st)
{
    b->h_npts[frame] = n;
    b->binned_valid = false;
    RCLC_CUDA(cudaMemcpyAsync(b->d_npts + frame, &b->h_npts[frame], 8, cudaMemcpyHostToDevice, st));
    return RCLC_OK;
}

int rclc_batch_upload(rclc_batch* b, int32_t frame, const void* pts, int stride, int ioff, int64_t n)
{
    RCLC_CHECK(b && (pts || n == 0), RCLC_ERR_INVALID, "null argument");
    RCLC_CHECK(frame >= 0 && frame < b->n_frames && n >= 0 && n <= b->max_pts, RCLC_ERR_INVALID, "frame/n out of range");
    cudaStream_t st;
    RCLC_TRY(frame_stream(b, frame, &st));
    RCLC_TRY(upload_cloud_f4(b->ctx, pts, stride, ioff, n, b->d_pristine + (size_t)frame * b->v.frame_stride, st));
    return set_npts(b, frame, n, st);
}

int rclc_batch_set_frame_dev(rclc_batch* b, int32_t frame, const void* pts_dev, int64_t n)
{
    RCLC_CHECK(b && (pts_dev || n == 0), RCLC_ERR_INVALID, "null argument");
    RCLC_CHECK(frame >= 0 && frame < b->n_frames && n >= 0 && n <= b->max_pts, RCLC_ERR_INVALID, "frame/n out of range");
    cudaStream_t st;
    RCLC_TRY(frame_stream(b, frame, &st));
    RCLC_CUDA(cudaMemcpyAsync(b->d_pristine + (size_t)frame * b->v.frame_stride, pts_dev, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
    return set_npts(b, frame, n, st);
}

int64_t rclc_batch_total_points(rclc_batch* b)
{
    int64_t t = 0;
    if (b) for (int64_t n : b->h_npts) t += n;
    return t;
}

int rclc_batch_bin(rclc_batch* b)
{
    RCLC_CHECK(b, RCLC_ERR_INVALID, "null batch");
    RCLC_CUDA(cudaSetDevice(b->ctx->device));
    RCLC_TRY(join_uploads(b));
    RCLC_CUDA(cudaMemsetAsync(b->v.n_done, 0, sizeof(int), b->ctx->stream));
    GroupRun gr = whole_batch(b);
    RCLC_TRY(launch_init(b, gr, nullptr));
    RCLC_TRY(launch_rebin(b, gr));
    b->round = gr.round;
    b->v.list_parity = gr.v.list_parity;
    b->binned_valid = true;
    return RCLC_OK;
}

int rclc_batch_sweep(rclc_batch* b, const double* poses)
{
    RCLC_CHECK(b && poses, RCLC_ERR_INVALID, "null argument");
    RCLC_CHECK(b->binned_valid, RCLC_ERR_INVALID, "rclc_batch_bin must be called first");
    rclc_ctx* ctx = b->ctx;
    RCLC_TRY(ctx->h_pin2.reserve(sizeof(PoseCoef) * b->n_frames));
    RCLC_TRY(ctx->d_tmp2.reserve(sizeof(PoseCoef) * b->n_frames));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));        // pinned staging reuse
    PoseCoef* hp = ctx->h_pin2.as<PoseCoef>();
    for (int f = 0; f < b->n_frames; ++f) hp[f] = host_pose_coef(poses + 3 * f);
    RCLC_CUDA(cudaMemcpyAsync(ctx->d_tmp2.p, hp, sizeof(PoseCoef) * b->n_frames, cudaMemcpyHostToDevice, ctx->stream));
    k_set_pose<<<b->n_frames, 256, 0, ctx->stream>>>(b->v, ctx->d_tmp2.as<PoseCoef>());
    ctx->launches += 1;
    RCLC_CUDA(cudaEventRecord(ctx->evk0, ctx->stream));
    RCLC_TRY(launch_sweep(b, false));
    RCLC_CUDA(cudaEventRecord(ctx->evk1, ctx->stream));
    return RCLC_OK;
}

int rclc_batch_read_eval(rclc_batch* b, double* residuals, double* jac, double* acc)
{
    RCLC_CHECK(b, RCLC_ERR_INVALID, "null batch");
    rclc_ctx* ctx = b->ctx;
    const size_t F = b->n_frames, nt = b->v.geom.nt;
    k_eval_out<<<b->n_frames, 128, 0, ctx->stream>>>(b->v);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    std::vector<double> h(F * nt * 4);
    RCLC_CUDA(cudaMemcpyAsync(h.data(), b->v.eval_out, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (acc)            // pose slot 0 of every frame
        RCLC_CUDA(cudaMemcpy2DAsync(acc, nt * 16 * 8, b->v.acc, (size_t)b->v.acc_stride * 8, nt * 16 * 8, F, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < F * nt; ++k) {
        if (residuals) residuals[k] = h[4 * k];
        if (jac) { jac[3 * k] = h[4 * k + 1]; jac[3 * k + 1] = h[4 * k + 2]; jac[3 * k + 2] = h[4 * k + 3]; }
    }
    return RCLC_OK;
}

#ifdef RCLC_SWEEP_TIMING
int rclc_lab_sweep_timing(long long* out, int32_t n_warps)       // lab build only
{
    RCLC_CUDA(cudaDeviceSynchronize());
    RCLC_CUDA(cudaMemcpyFromSymbol(out, g_sw_dbg, sizeof(long long) * 8 * std::min(n_warps, 8192)));
    return RCLC_OK;
}
#endif
#ifdef RCLC_ST_TIMING
int rclc_lab_st_timing(long long* out, int32_t n_warps)          // lab build only
{
    RCLC_CUDA(cudaDeviceSynchronize());
    RCLC_CUDA(cudaMemcpyFromSymbol(out, g_st_dbg, sizeof(long long) * kStDbgSlots * std::min(n_warps, 4096)));
    return RCLC_OK;
}
#endif

int rclc_batch_set_stream_sweep(rclc_batch* b, int32_t on)
{
    RCLC_CHECK(b, RCLC_ERR_INVALID, "null batch");
    RCLC_CHECK(!on || st_smem_bytes(b->v.geom.nt, b->v.geom.ncells) <= 72 * 1024, RCLC_ERR_UNSUPPORTED,
               "board too large for the streaming sweep's shared memory (three CTAs per SM)");
    b->stream_sweep = on != 0;
    for (cudaGraphExec_t ge : b->round_graphs) cudaGraphExecDestroy(ge);      // the graphs hold the launches
    b->round_graphs.clear();
    b->graphs_groups = 0;
    return RCLC_OK;
}

int rclc_batch_set_exact_sweep(rclc_batch* b, int32_t on)
{
    RCLC_CHECK(b, RCLC_ERR_INVALID, "null batch");
    b->v.force_exact = on ? 1 : 0;
    for (cudaGraphExec_t ge : b->round_graphs) cudaGraphExecDestroy(ge);      // the graphs hold the view by value
    b->round_graphs.clear();
    b->graphs_groups = 0;
    return RCLC_OK;
}

int rclc_batch_gridfit_solve(rclc_batch* b, const double* T_bl, double* T_bl_out, double* corners_out, rclc_gridfit_report* reports)
{
    RCLC_CHECK(b, RCLC_ERR_INVALID, "null batch");
    rclc_ctx* ctx = b->ctx;
    RCLC_CUDA(cudaSetDevice(ctx->device));
    const int F = b->n_frames;
    const GridGeom& g = b->v.geom;
    double* d_Tbl = nullptr;
    if (T_bl) {
        RCLC_TRY(ctx->d_tmp2.reserve((size_t)F * 16 * 8));
        d_Tbl = ctx->d_tmp2.as<double>();
        RCLC_CUDA(cudaMemcpyAsync(d_Tbl, T_bl, (size_t)F * 16 * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    // opt_grid_fitting.h:47-56 apply_noise_test is a debugging aid that perturbs the input; the caller
    // can apply it with rclc_transform_cloud. It is rejected here rather than silently ignored.
    RCLC_CHECK(!b->cfg.apply_noise_test, RCLC_ERR_UNSUPPORTED, "apply_noise_test: perturb the input with rclc_transform_cloud instead");
    b->binned_valid = false;
    RCLC_TRY(ctx->h_pin2.reserve(256));
    int* h_done = ctx->h_pin2.as<int>();           // [G] finished frames per group, as of each group's last poll
    // Frames are independent until the extrinsic solve, so the batch runs as G groups on G streams, each advancing on its own:
    // while the LM steps of one group (a single warp per frame, a few us of dependent fp64) hold the tail of its sweep kernel, the
    // other groups' sweeps keep the SMs busy, and a group whose frames need few rounds finishes early.
    const int G = solve_groups(b);
    RCLC_TRY(ensure_group_streams(ctx, G));
    b->uploads_pending = false;          // every group stream carries its own frames' uploads ahead of its kernels
    RCLC_CHECK(G <= 32, RCLC_ERR_INVALID, "at most 32 solve groups");
    RCLC_CUDA(cudaMemsetAsync(b->v.n_done, 0, 64 * sizeof(int), ctx->stream));
    RCLC_CUDA(cudaEventRecord(ctx->ev_sync, ctx->stream));              // uploads, T_bl, ... precede every group
    std::vector<GroupRun> groups(G);
    for (int gi = 0; gi < G; ++gi) {
        GroupRun& gr = groups[gi];
        gr.f0 = (int)((int64_t)F * gi / G);
        gr.nf = (int)((int64_t)F * (gi + 1) / G) - gr.f0;
        gr.st = gi == 0 ? ctx->stream : ctx->aux_streams[gi - 1];
        gr.v = slice_view(b->v, gr.f0, gr.nf, gi);
        gr.v.fuse_lm = 1;
        gr.round = 0;
        if (gi > 0) RCLC_CUDA(cudaStreamWaitEvent(gr.st, ctx->ev_sync, 0));
        RCLC_TRY(launch_init(b, gr, d_Tbl));
    }
    // Upper bound on rounds: every outer iteration needs at most max_lm_iters+2 evaluations.
    const int64_t max_rounds = (int64_t)(g.max_iter_time + 2) * (g.max_lm_iters + 3);
    // ---- first batch of rounds, launched directly: the one counting sort of the solve, then {k_move, k_sweep} ----
    for (int r = 0; r < kRoundsPerGraph; ++r)
        for (int gi = 0; gi < G; ++gi) {
            GroupRun& gr = groups[gi];
            if (r == 0) RCLC_TRY(launch_rebin(b, gr));
            else RCLC_TRY(launch_move(b, gr));
            RCLC_TRY(launch_sweep_view(b, gr.v, gr.nf, gr.st, F));      // sweep + fused LM step (last warp per frame)
        }
    // ---- every later batch is the same kRoundsPerGraph x {k_move, k_sweep} per group: captured once per batch object into a CUDA
    //      graph. From here on every group advances on its own: its graph is re-launched as soon as ITS previous batch has finished
    //      and still left frames unsolved. No group waits for another. ----
    static_assert(kRoundsPerGraph % 2 == 0, "a graph must leave the parity of the re-bin lists as it found it");
    if (b->graphs_groups != G) {
        for (cudaGraphExec_t ge : b->round_graphs) cudaGraphExecDestroy(ge);
        b->round_graphs.clear();
        b->graphs_groups = 0;
    }
    if (b->round_graphs.empty()) {
        for (int gi = 0; gi < G; ++gi) {
            GroupRun gr = groups[gi];                   // a copy: capturing must not advance the real round counters
            cudaGraph_t graph = nullptr;
            RCLC_CUDA(cudaStreamBeginCapture(gr.st, cudaStreamCaptureModeThreadLocal));
            int rc_cap = RCLC_OK;
            for (int r = 0; r < kRoundsPerGraph && rc_cap == RCLC_OK; ++r) {
                rc_cap = launch_move(b, gr);
                if (rc_cap == RCLC_OK) rc_cap = launch_sweep_view(b, gr.v, gr.nf, gr.st, F);
            }
            RCLC_CUDA(cudaStreamEndCapture(gr.st, &graph));
            RCLC_TRY(rc_cap);
            cudaGraphExec_t ge = nullptr;
            RCLC_CUDA(cudaGraphInstantiate(&ge, graph, 0));
            cudaGraphDestroy(graph);
            b->round_graphs.push_back(ge);
            ctx->launches -= 2 * kRoundsPerGraph;       // (counted when the graph is launched)
        }
        b->graphs_groups = G;
    }
    while ((int)ctx->ev_groups.size() < G) {
        cudaEvent_t e;
        RCLC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_groups.push_back(e);
    }
    auto poll = [&](int gi) -> int {
        RCLC_CUDA(cudaMemcpyAsync(h_done + gi, groups[gi].v.n_done, sizeof(int), cudaMemcpyDeviceToHost, groups[gi].st));
        RCLC_CUDA(cudaEventRecord(ctx->ev_groups[(size_t)gi], groups[gi].st));
        return RCLC_OK;
    };
    for (int gi = 0; gi < G; ++gi) RCLC_TRY(poll(gi));
    std::vector<char> finished((size_t)G, 0);
    std::vector<int64_t> g_rounds((size_t)G, kRoundsPerGraph);
    int n_finished = 0;
    while (n_finished < G) {
        bool progressed = false;
        for (int gi = 0; gi < G; ++gi) {
            if (finished[(size_t)gi]) continue;
            const cudaError_t q = cudaEventQuery(ctx->ev_groups[(size_t)gi]);
            if (q == cudaErrorNotReady) continue;
            RCLC_CUDA(q);
            progressed = true;
            if (h_done[gi] >= groups[gi].nf) { finished[(size_t)gi] = 1; ++n_finished; continue; }
            RCLC_CHECK(g_rounds[(size_t)gi] < max_rounds, RCLC_ERR_NUMERIC, "grid-fit state machine did not terminate");
            RCLC_CUDA(cudaGraphLaunch(b->round_graphs[(size_t)gi], groups[gi].st));
            groups[gi].round += kRoundsPerGraph;
            g_rounds[(size_t)gi] += kRoundsPerGraph;
            ctx->launches += 2 * kRoundsPerGraph;
            RCLC_TRY(poll(gi));
        }
        if (!progressed) std::this_thread::yield();
    }
    for (int gi = 1; gi < G; ++gi) {                                    // ctx->stream joins the other groups
        RCLC_CUDA(cudaEventRecord(ctx->ev_sync, groups[gi].st));
        RCLC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_sync, 0));
    }
    std::vector<LmState> h_lm(F);
    RCLC_CUDA(cudaMemcpyAsync(h_lm.data(), b->v.lm, sizeof(LmState) * F, cudaMemcpyDeviceToHost, ctx->stream));
    if (corners_out)
        RCLC_CUDA(cudaMemcpyAsync(corners_out, b->v.corners, (size_t)F * g.nc * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
#ifdef RCLC_LAB
    for (int f = 0; f < F; ++f) {
        const long long* d = h_lm[f].dbg;
        if (d[4] > 0)
            fprintf(stderr, "[rclc lab] frame %d: %lld LM steps; per step: load %.2f us, replay %.2f us, proposals %.2f us; first -> last step %.1f us (%.1f us per round)\n",
                    f, d[4], 1e-3 * d[0] / d[4], 1e-3 * d[1] / d[4], 1e-3 * d[2] / d[4], 1e-3 * (d[6] - d[5]),
                    d[4] > 1 ? 1e-3 * (d[6] - d[5]) / (d[4] - 1) : 0.0);
    }
#endif
    for (int f = 0; f < F; ++f) {
        if (reports) reports[f] = h_lm[f].rep;
        if (T_bl_out) for (int k = 0; k < 16; ++k) T_bl_out[(size_t)f * 16 + k] = h_lm[f].T_bl[k];
    }
    return RCLC_OK;
}

int rclc_gridfit_eval(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                      const double pose[3], double* residuals, double* jac, double* acc)
{
    RCLC_CHECK(ctx && cfg && pts && pose && residuals && n > 0, RCLC_ERR_INVALID, "null argument");
    rclc_batch* b = nullptr;
    RCLC_TRY(rclc_batch_create(ctx, cfg, 1, n, &b));
    int rc = rclc_batch_upload(b, 0, pts, stride, ioff, n);
    if (rc == RCLC_OK) rc = rclc_batch_bin(b);
    if (rc == RCLC_OK) rc = rclc_batch_sweep(b, pose);
    if (rc == RCLC_OK) rc = rclc_batch_read_eval(b, residuals, jac, acc);
    rclc_batch_destroy(b);
    return rc;
}

int rclc_gridfit_solve(rclc_ctx* ctx, const rclc_config* cfg, void* pts_io, int stride, int ioff, int64_t n,
                       double* T_bl_io, double* corners_out, rclc_gridfit_report* report)
{
    RCLC_CHECK(ctx && cfg && pts_io && corners_out && n > 0, RCLC_ERR_INVALID, "null argument");
    rclc_batch* b = nullptr;
    RCLC_TRY(rclc_batch_create(ctx, cfg, 1, n, &b));
    int rc = rclc_batch_upload(b, 0, pts_io, stride, ioff, n);
    rclc_gridfit_report rep;
    if (rc == RCLC_OK) rc = rclc_batch_gridfit_solve(b, T_bl_io, nullptr, corners_out, &rep);
    if (rc == RCLC_OK && report) *report = rep;
    rclc_batch_destroy(b);
    return rc;
}

int rclc_gridfit_multistart(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                            const double* poses, int32_t P, double* cost_out)
{
    RCLC_CHECK(ctx && cfg && pts && poses && cost_out && n > 0 && P > 0, RCLC_ERR_INVALID, "bad argument");
    rclc_batch* b = nullptr;
    RCLC_TRY(rclc_batch_create(ctx, cfg, 1, n, &b));
    int rc = rclc_batch_upload(b, 0, pts, stride, ioff, n);
    if (rc == RCLC_OK) rc = rclc_batch_bin(b);
    if (rc == RCLC_OK) rc = rclc_batch_multistart(b, 0, poses, P, cost_out);
    rclc_batch_destroy(b);
    return rc;
}

int rclc_batch_multistart(rclc_batch* b, int32_t frame, const double* poses, int32_t P, double* cost_out)
{
    RCLC_CHECK(b && poses && cost_out && P > 0 && frame >= 0 && frame < b->n_frames, RCLC_ERR_INVALID, "bad argument");
    RCLC_CHECK(b->binned_valid, RCLC_ERR_INVALID, "rclc_batch_bin must be called first");
    rclc_ctx* ctx = b->ctx;
    RCLC_CUDA(cudaSetDevice(ctx->device));
    const size_t nt = b->v.geom.nt;
    const size_t o_ctl = (sizeof(PoseCoef) * P + 255) & ~(size_t)255;
    const size_t o_acc = o_ctl + ((sizeof(FrameCtl) * P + 255) & ~(size_t)255);
    const size_t o_cost = o_acc + (size_t)P * nt * 16 * 8;
    RCLC_TRY(ctx->d_tmp.reserve(o_cost + (size_t)P * 8));
    RCLC_TRY(ctx->h_pin2.reserve(sizeof(PoseCoef) * P));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    PoseCoef* hp = ctx->h_pin2.as<PoseCoef>();
    for (int p = 0; p < P; ++p) hp[p] = host_pose_coef(poses + 3 * p);
    char* base = static_cast<char*>(ctx->d_tmp.p);
    PoseCoef* d_poses = reinterpret_cast<PoseCoef*>(base);
    FrameCtl* d_ctl = reinterpret_cast<FrameCtl*>(base + o_ctl);
    double* d_acc = reinterpret_cast<double*>(base + o_acc);
    double* d_cost = reinterpret_cast<double*>(base + o_cost);
    RCLC_CUDA(cudaMemcpyAsync(d_poses, hp, sizeof(PoseCoef) * P, cudaMemcpyHostToDevice, ctx->stream));
    RCLC_CUDA(cudaMemsetAsync(d_acc, 0, (size_t)P * nt * 16 * 8, ctx->stream));
    BatchView v = b->v;
    k_ms_setup<<<(P + 127) / 128, 128, 0, ctx->stream>>>(b->v, d_ctl, d_poses, P, frame);
    v.fixed_frame = frame;
    v.fuse_lm = 0;
    v.acc_stride = (int64_t)nt * 16;            // one pose slot per hypothesis
    // blockIdx.y is limited to 65535: sweep poses in slabs
    for (int p0 = 0; p0 < P; p0 += 32768) {
        const int np = std::min(32768, P - p0);
        BatchView vs = v;
        vs.ctl = d_ctl + p0;
        vs.acc = d_acc + (size_t)p0 * nt * 16;
        vs.n_frames = np;
        RCLC_TRY(launch_sweep_view(b, vs, np, ctx->stream, np));
        k_ms_cost<<<np, kLmThreads, 0, ctx->stream>>>(vs, d_cost + p0);
        ctx->launches += 1;
    }
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    RCLC_CUDA(cudaMemcpyAsync(cost_out, d_cost, (size_t)P * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    return RCLC_OK;
}

} // extern "C"
