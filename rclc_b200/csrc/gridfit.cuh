// gridfit.cuh — data structures of the grid-fit path (K4/K5): the batch view every kernel of gridfit.cu receives.
//
// Design: the reference evaluates 82 cost functors, each looping over its own kd-tree neighbour copy
// (board_fitting_cost.h:124-210, 3.3*N point visits per sweep, fp64). Here a frame's points are counting-sorted ONCE per solve
// into the cells of the half-square lattice on which all targets sit (and, inside a cell, into 4x4 sub-cells, so that 32
// consecutive points — a "row" — are spatial neighbours); later outer iterations move the sorted points in place. Every cell's
// segment is padded to a multiple of 32 points with far-away sentinels, and every row has a bounding box.
// A sweep is a GATHER: one warp owns one target of one frame, pulls in exactly the rows whose box touches the target's discs
// (as moved by the trial pose), keeps the target's six sums in registers and writes the target's 16 accumulators once — no
// atomics in the hot loop.
#pragma once
#include "common.cuh"

namespace rclc {

constexpr int kMaxTargets = 1024;
constexpr int kMaxCells = 65536;          // (cell, sub-cell) bins per frame: global histograms
constexpr int kMaxCoarse = 1024;         // lattice cells (kMaxCells bounds cells x sub-cells)
constexpr int kSweepThreads = 128;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kRowList = 256;            // live rows a warp collects before it streams them
constexpr int kGroupRows = 4;            // rows per cp.async group (one commit / wait per group)
constexpr int kRingGroups = 4;           // groups per warp ring (8 KB): one computing, up to three in flight (8 groups, or 8-row groups, measured no faster)
constexpr int kFlushRows = 448;          // rows a lane may accumulate before the count riding in the fp64 sum must be split off
constexpr int kBinThreads = 256;
constexpr int kSpec = 8;                 // most poses one round evaluates per frame: the LM candidate and the ones that follow it if every one is rejected

// Geometry derived from rclc_config on the host (gridfit.cu: build_geom).
struct GridGeom {
    int nt;                    // number of targets (col targets first, then row targets)
    int nc;                    // number of corners
    int ncx, ncy, ncells;      // lattice cells
    int sub, nfine;            // sub-cells per axis inside a cell (sort key only); ncells * sub * sub
    int i0, j0;                // site index of cell (0,0)'s lower corner
    double g, inv_g;           // lattice unit = side/2
    double rc2, rg2;           // cost / gradient radius squared (board_fitting_cost.h:58-59)
    float rn2f;                // (float)(neighbour_radius^2)  (FLANN radius)
    double huber_a;
    double function_tol;
    int max_lm_iters;
    int max_iter_time;
    int board_w, board_h;
    double side;
};

struct PoseCoef {              // R = Quaternion(cos(t/2),0,0,sin(t/2)).matrix(), t = (x,y)
    double r00, r01, r10, r11, tx, ty;
    double yaw_deg;
};

enum Phase { PH_EVAL_INIT = 1, PH_EVAL_CAND = 2, PH_DONE = 3 };

// Per-frame control block read by the streaming kernels.
struct alignas(16) FrameCtl {
    int sweep_active;          // sweep kernels process this frame
    int rebin;                 // count/scatter kernels process this frame (move + re-bin)
    int n_rows;                // 32-point rows of the current binning (cells padded to whole rows)
    int n_binned;              // points that landed in a lattice cell
    int use_exact;             // some intensity is outside {0} U [2^-3, 256): counts cannot ride in the fp64 sums -> all-fp64 body
    float drift;               // bound on how far any point has moved since the frame was sorted (in-place moves)
    float move[8];             // T_w1_w0 rows 0,1: m00 m01 m02 m03 / m10 m11 m12 m13
    int n_pose;                // poses the next sweep evaluates (1 .. kSpec)
    int pad_;
    PoseCoef pose[kSpec];      // pose[0]: the point the LM needs next; pose[1..]: its successors under "every step is rejected"
};
static_assert(sizeof(FrameCtl) % 16 == 0, "FrameCtl is staged with 16-byte loads");

// Ceres TrustRegionMinimizer state of the running inner solve. Kept apart from the rest of the frame's state because the
// LM step runs it twice: for real (lm_consume on the evaluations of a round) and on a scratch copy that is fed "rejected"
// to predict the next candidates (speculate_candidates).
struct LmCore {
    int phase;
    int iteration;
    int num_consecutive_invalid;
    int reuse_diagonal;
    int step_is_successful;
    int n_lm_iterations_this_solve;
    double x[3], cand[3], best[3];
    double x_cost, x_norm, radius, decrease_factor;
    double minimum_cost, initial_cost, min_iter_cost;
    double Au[6], bu[3];       // unscaled J'J (00,01,02,11,12,22) and J'r at x
    double scale[3], diag[3];
    double model_cost_change;
    double gmax;
};

// LmCore + opt_grid_fitting outer-loop state, one per frame.
struct alignas(16) LmState {
    LmCore core;
    int outer_iter;
    int status;
    int chain;                 // 0: no step of the running solve has been accepted yet; 1: after the first accepted step
    int n_rejected;            // rejections since the last accepted step (or since the start of the solve)
    double spec[kSpec][3];     // the poses of the round in flight: spec[0] is the LM's candidate, spec[k] what follows k rejections
    double spec_mcc[kSpec];    // model_cost_change of spec[k]
    double spec_radius[kSpec]; // the trust-region radius and decrease factor spec[k] was computed for: the replay adopts spec[k]
    double spec_factor[kSpec]; // only if, after k rejections, the minimiser's own are these, bit for bit
    float T_base_laser[16];
    double T_bl[16];
    rclc_gridfit_report rep;
#ifdef RCLC_LAB
    long long dbg[8];          // lab build only: ns spent in the stages of lm_frame_step, summed over its calls (scripts/lm_timing.py)
#endif
};
static_assert(sizeof(LmState) % 16 == 0, "LmState is staged with 16-byte loads");

struct BatchView {
    GridGeom geom;
    int n_frames;
    int64_t frame_stride;          // float4 elements between frames in raw/binned
    int parts;                     // warps that share one (frame, target): row groups are dealt round-robin
    int fixed_frame;               // >= 0: every blockIdx.y of k_sweep reads this frame's points (multi-start sweep)
    const int64_t* n_pts;          // [F]
    const float4* raw;             // [F][frame_stride]   the uploaded clouds (x,y,z,I); read once, by the initial sort
    float4* binned;                // [F][frame_stride]   current pc_iter, cell-sorted: {x, y, I, z}
    int* hist;                     // [F][nfine]
    int* cursor;                   // [F][nfine]
    int* cellstart;                // [F][ncells] first point of the cell's segment in `binned` (multiple of 32)
    int* cellcnt;                  // [F][ncells] points of the cell (the segment holds pad32 of it)
    float4* rowbox;                // [F][frame_stride / 32] {xmin, ymin, xmax, ymax} of every 32-point row
    unsigned* cellz2;              // [F][ncells] float bits of max z^2 over the cell's points (decides whether the neighbour mask can matter)
    double* acc;                   // [F][kSpec][nt][16] (acc_stride doubles per frame; a plain sweep uses pose slot 0)
    int64_t acc_stride;            // doubles between the accumulator tables of two frames
    int npose_max;                 // pose slots a sweep launch covers per frame: 1 (plain sweeps, multi-start) or kSpec (the solve)
    FrameCtl* ctl;                 // [F]
    LmState* lm;                   // [F]
    int* n_done;                   // [1]
    int* rebin_list;               // [2][F] frames that need a re-bin: the LM steps of round r fill list (r+1)&1
    int* rebin_cnt;                // [2]
    int list_parity;               // which list this round's binning kernels consume
    double* terms;                 // [F][kSpec][12][nt] per-target LM terms (cost, 6 J'J, 3 J'r, 2 flags) written by the sweep
    int* tgt_done;                 // [F][kSpec][nt] parts of a (pose, target) that have flushed their accumulators
    int* slot_done;                // [F][kSpec] targets of a pose slot whose terms are written (the last one sums them)
    double* evsum;                 // [F][kSpec][12] the summed terms of every pose slot: what Ceres' evaluation returns
    int* bin_done;                 // [F] CTAs of k_count that have finished their share of the histogram
    int* sweep_done;               // [F] pose slots of the current sweep that are summed (the last one runs the LM step)
    int flat_tasks;                // plain sweeps: the tasks of ALL frames packed into consecutive warps (gridDim.y == 1) — no two idle
                                   // warps per frame, and six full waves instead of 6.05 on a 128-frame batch
    int fuse_lm;                   // 1: k_sweep's last warp per frame advances the Ceres state machine
    int force_exact;               // all-fp64 sweep body for every frame (rclc_batch_set_exact_sweep: the A/B reference of the fast body)
    // geometry tables
    const double2* tgt_xy;         // [nt]
    const float2* tgt_xyf;         // [nt] (float)target  (FLANN query point)
    const uint8_t* tgt_is_row;     // [nt]
    double* eval_out;              // [F][nt][4] residual, j0, j1, j2 (rclc_batch_read_eval)
    double* corners;               // [F][nc][3]
};

} // namespace rclc

struct rclc_batch;
namespace rclc { int batch_join_uploads(rclc_batch* b); }      // ctx->stream waits for uploads that went to the other group streams

struct rclc_batch {
    rclc_ctx* ctx = nullptr;
    rclc_config cfg;
    rclc::BatchView v;
    int n_frames = 0;
    int64_t max_pts = 0;
    bool uploads_pending = false;     // some frames were uploaded on a group stream other than ctx->stream since the last join
    std::vector<int64_t> h_npts;
    float4* d_pristine = nullptr;     // [F][frame_stride] uploaded clouds (solve works on a copy)
    void* d_block = nullptr;          // one allocation for the small arrays
    int64_t* d_npts = nullptr;
    bool binned_valid = false;
    bool stream_sweep = false;        // rclc_batch_set_stream_sweep: plain sweeps and the solve use k_sweep_stream (CTA per range of rows, TMA ring) instead of k_sweep
    int round = 0;                    // parity of the re-bin list the next round consumes
    // one instantiated CUDA graph per solve group: kRoundsPerGraph rounds of {k_move, k_sweep} (the steady state of the solve)
    std::vector<cudaGraphExec_t> round_graphs;
    int graphs_groups = 0;
    // GMM scratch
    double* d_gmm = nullptr;
};
