// sweep_stream.cuh — k_sweep_stream: the grid-fit residual pass (K4) as a STREAM over the sorted cloud.
// Included by gridfit.cu (inside namespace rclc) after the gather sweep, whose helpers it shares.
//
// Replaces the same reference code as k_sweep (BOARD_FITTING_COST::Evaluate x 82, board_fitting_cost.h:74-262, and the
// kd-tree neighbour set, :264-280). What changes is who owns what:
//
//   k_sweep        one warp owns a TARGET and gathers the rows its disc touches — a point is pulled from L2 once per disc
//                  it lies in (1.7 x), after a long per-task chain of dependent look-ups (cells, row boxes) during which
//                  nothing streams;
//   k_sweep_stream one CTA owns a contiguous RANGE OF ROWS of one (frame, pose). A producer thread streams the range through a
//                  shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier, 8 KB of points + their 16 row boxes
//                  per stage) — every row is read from L2/DRAM exactly once, in address order, and the first copies leave
//                  before anything else of the CTA has been looked up. Four consumer warps classify each tile against the
//                  few targets whose discs can reach the lattice cell the rows belong to ("candidates": the two targets on
//                  the cell's corners, plus whatever the trial pose and the drift bring into reach). Warp (g, p) takes the
//                  candidates of rank g, g + 2 (nearest first) on the rows of parity p and keeps their six signed sums in
//                  registers; at the end of the cell's rows ("segment") the sums are folded over the warp and added to the
//                  target's 16 accumulators with fp64 atomics — the sums are exact (gridfit.cu, accumulate6), so the tables
//                  stay bit-identical to the gather sweep's and to the oracle's. Candidates beyond the four register sets,
//                  and frames whose intensities cannot carry the riding count (use_exact), go through an all-fp64 row routine
//                  that reduces and adds row by row.
//
// Per point and candidate the hot loop issues 22 instructions (k_sweep: ~35): the rotated offsets directly from the stored
// coordinates (w = R p - k, k = ref - R t: two FFMA chains; d^2 = |w|^2), the two boundary trackers as three-input minima
// (FMNMX3 with |.| modifiers), the class multipliers as 2 FSET (1.0f / 0.0f IS the high word of the double 2^-7) + 4 LOP3
// (a ^ (~w & signbit) in one instruction), six DFMAs. The accumulators therefore carry 2^-7 of the sums — exact, undone at
// the flush.
#pragma once

constexpr int kStConsWarps = 4;                            // consumer warps: 2 candidate groups x 2 row phases
constexpr int kStPhases = kStConsWarps / 2;                 // row phases
constexpr int kStConsThreads = kStConsWarps * 32;
constexpr int kStThreads = kStConsThreads + 32;            // + the producer warp
constexpr int kStTileRows = 16;
constexpr int kStStages = 6;
constexpr uint32_t kStTileBytes = kStTileRows * 512u;
constexpr uint32_t kStStageBytes = kStTileBytes + kStTileRows * 16u;        // points + row boxes
constexpr int kStMaxCand = 16;                             // sorted candidates kept per segment; more -> every target, row by row

// ---- mbarrier / TMA bulk copy ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// A waiting warp that re-issues try_wait back to back is not free: on B200 the polls of 580 producer threads alone kept the XU
// pipe — where the hot loop's F2F conversions go — 100 % busy (ncu, profiles/r02_stream_history.md). Waiters sleep between polls.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try(bar, parity)) __nanosleep(SLEEP_NS);
}
__device__ __forceinline__ void cons_bar_sync_all() { __syncthreads(); }
__device__ __forceinline__ void cons_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStConsThreads) : "memory"); }

#ifdef RCLC_ST_TIMING
// lab build only (scripts/st_timing.py): per consumer warp, cycles spent in the stages of one sweep
constexpr int kStDbgSlots = 8;          // 0 prologue, 1 tile waits, 2 next_segment, 3 flush, 4 rows, 5 total, 6 segments, 7 tiles
__device__ long long g_st_dbg[4096 * kStDbgSlots];
#define ST_T(var) const long long var = clock64()
#define ST_ADD(slot, t_from) dbg[slot] += clock64() - (t_from)
#else
#define ST_T(var)
#define ST_ADD(slot, t_from)
#endif

struct StTarget { float nkx, nky, cx, cy, tqx, tqy, disp, pad; double refx, refy; };      // 48 B per target and pose
static_assert(sizeof(StTarget) == 48, "one 48-byte record per target");

// What the producer warp tells the consumers about one segment (the rows of one lattice cell inside the CTA's range)
struct StSeg {
    int seg_end;                          // first row (of the frame) after the segment
    int n_cand;                           // candidates, nearest disc centre first (at most kStMaxCand kept)
    int all_targets;                      // more than kStMaxCand discs reach the cell: every target is a candidate
    int pad;
    unsigned short cand[kStMaxCand];
};
constexpr int kStSegRing = 8;

struct StShared {
    unsigned long long full[kStStages], empty[kStStages];
    unsigned long long seg_full[kStSegRing], seg_empty[kStSegRing];
    StSeg seg[kStSegRing];
    PoseCoef pc;
    float r00, r01, r10, r11, rg2, rc2, rn2, rgf;
    double rg2d, rc2d;
    unsigned z2max_bits, bc_bits;
    int last;
    // constants of the CTA that st_segment reads back (so that it needs three arguments)
    uint32_t ring;                        // shared address of the tile ring
    int r0, r1, nt, frame_exact;
    float E, E2, reach2, z2max;
    double* acc_slot;
    struct Warp { int cur, n_waited, seg_j, pad; } w[kStConsWarps];
#ifdef RCLC_ST_TIMING
    long long dbg[kStConsWarps][kStDbgSlots];
#endif
};

// fp32 geometry of one point against one candidate — the SAME intrinsics in the hot loop and in the correction.
// w = R p - k = x' - ref: the point is right / down of the target when wx / wy is positive.
__device__ __forceinline__ void st_geom(float r00, float r01, float r10, float r11, float nkx, float nky, float qx, float qy,
                                        float& wx, float& wy, float& d2)
{
    wx = __fmaf_rn(r00, qx, __fmaf_rn(r01, qy, nkx));
    wy = __fmaf_rn(r10, qx, __fmaf_rn(r11, qy, nky));
    d2 = __fmaf_rn(wx, wx, __fmul_rn(wy, wy));
}
__device__ __forceinline__ float fmin3_abs(float m, float a, float b)
{
    float r;
    asm("{\n\t.reg .f32 t, u;\n\tabs.f32 t, %2;\n\tabs.f32 u, %3;\n\tmin.f32 %0, %1, t, u;\n\t}" : "=f"(r) : "f"(m), "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint32_t fset_le(float a, float b)      // 0x3f800000 when a <= b, else 0
{
    uint32_t r;
    asm("set.le.f32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint32_t fset_lt(float a, float b)
{
    uint32_t r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
// h ^ (~w & 0x80000000): the multiplier's high word with the sign set when w is positive (right / down)
__device__ __forceinline__ uint32_t sign_mix(uint32_t h, float w)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD2;" : "=r"(r) : "r"(h), "r"(__float_as_uint(w)));
    return r;
}

// One point against one candidate: 22 instructions. a[] carries 2^-7 of the signed sums of accumulate6.
__device__ __forceinline__ void st_point(double (&a)[6], const float4& q, double v, float r00, float r01, float r10, float r11,
                                         float nkx, float nky, float rg2, float rc2, int zlo, float& mA, float& mB)
{
    float wx, wy, d2;
    st_geom(r00, r01, r10, r11, nkx, nky, q.x, q.y, wx, wy, d2);
    mA = fmin3_abs(mA, __fsub_rn(d2, rg2), __fsub_rn(d2, rc2));
    mB = fmin3_abs(mB, wx, wy);
    const uint32_t hin = fset_le(d2, rg2), hc = fset_lt(d2, rc2);
    a[0] = fma(__hiloint2double((int)hin, zlo), v, a[0]);
    a[1] = fma(__hiloint2double((int)sign_mix(hin, wy), zlo), v, a[1]);
    a[2] = fma(__hiloint2double((int)sign_mix(hin, wx), zlo), v, a[2]);
    a[3] = fma(__hiloint2double((int)hc, zlo), v, a[3]);
    a[4] = fma(__hiloint2double((int)sign_mix(hc, wy), zlo), v, a[4]);
    a[5] = fma(__hiloint2double((int)sign_mix(hc, wx), zlo), v, a[5]);
}

// Rare: the point is within the fp32 error bound of a decision boundary of this candidate. Its class is re-evaluated with the
// reference's own arithmetic and the signed sums are corrected by (exact - fast), in the accumulators' 2^-7 scale.
__device__ __noinline__ Corr6 st_correct(float4 q, const StShared* sh, const StTarget* st, float E, float E2)
{
    Corr6 c;
#pragma unroll
    for (int j = 0; j < 6; ++j) c.d[j] = 0.0;
    float wx, wy, d2;
    st_geom(sh->r00, sh->r01, sh->r10, sh->r11, st->nkx, st->nky, q.x, q.y, wx, wy, d2);
    const bool near = fabsf(__fsub_rn(d2, sh->rg2)) <= E2 || fabsf(__fsub_rn(d2, sh->rc2)) <= E2 || fabsf(wx) <= E || fabsf(wy) <= E;
    if (!near) return c;
    const double f_in = d2 <= sh->rg2 ? 1.0 : 0.0, f_c = d2 < sh->rc2 ? 1.0 : 0.0;
    const double f_sy = (__float_as_uint(wy) & 0x80000000u) ? 1.0 : -1.0, f_sx = (__float_as_uint(wx) & 0x80000000u) ? 1.0 : -1.0;
    const ExactClass e = classify_exact(q.x, q.y, sh->pc, st->refx, st->refy, sh->rg2d, sh->rc2d);
    const double e_in = e.in ? 1.0 : 0.0, e_c = (e.in && e.cost) ? 1.0 : 0.0;
    const double e_sy = e.down ? -1.0 : 1.0, e_sx = e.right ? -1.0 : 1.0;
    const double v = ((double)q.z + kCntUnit) * 0.0078125;
    c.d[0] = (e_in - f_in) * v;
    c.d[1] = (e_in * e_sy - f_in * f_sy) * v;
    c.d[2] = (e_in * e_sx - f_in * f_sx) * v;
    c.d[3] = (e_c - f_c) * v;
    c.d[4] = (e_c * e_sy - f_c * f_sy) * v;
    c.d[5] = (e_c * e_sx - f_c * f_sx) * v;
    return c;
}

// Twelve totals that every lane holds (six intensity sums, six counts: in, in&down, in&right, cost, cost&down, cost&right) ->
// the 16 accumulators of a target, one fp64 atomic per lane 0..15.
__device__ __forceinline__ void st_add_totals(double* acc_t, const double (&S)[6], const double (&N)[6], int lane)
{
    double o[16];
    store_totals(o, S[0], S[1], S[2], S[3], S[4], S[5], N[0], N[1], N[2], N[3], N[4], N[5]);
    double mine = 0.0;
#pragma unroll
    for (int l = 0; l < 16; ++l) if (lane == l) mine = o[l];
    if (lane < 16 && mine != 0.0) atomicAdd(acc_t + lane, mine);
}

// End of a segment (or the riding count is about to run out of room): the warp's signed sums of one register set -> the target.
// Every nested total is n * 2^17 + S with S a multiple of 2^-26 below 2^17: as a 52-bit integer its count and three 18-bit limbs
// of S are summed over the warp with REDUX (one instruction each, no shuffle trees), exactly.
__device__ __noinline__ void st_flush(double a0, double a1, double a2, double a3, double a4, double a5, double* acc_t)
{
    const int lane = threadIdx.x & 31;
    double n6[6];
    n6[0] = a0 * 128.0;
    n6[1] = (a0 - a1) * 64.0;
    n6[2] = (a0 - a2) * 64.0;
    n6[3] = a3 * 128.0;
    n6[4] = (a3 - a4) * 64.0;
    n6[5] = (a3 - a5) * 64.0;
    double S[6], N[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const long long Q = __double2ll_rn(n6[j] * 67108864.0);                // n * 2^43 + S * 2^26, exact
        const unsigned lo = (unsigned)Q, hi = (unsigned)((unsigned long long)Q >> 32);
        unsigned n = hi >> 11, l0 = lo & 0x3ffffu, l1 = __funnelshift_r(lo, hi, 18) & 0x3ffffu, l2 = (hi >> 4) & 0x7fu;
        n = __reduce_add_sync(0xffffffffu, n);
        l0 = __reduce_add_sync(0xffffffffu, l0);
        l1 = __reduce_add_sync(0xffffffffu, l1);
        l2 = __reduce_add_sync(0xffffffffu, l2);
        const unsigned long long s = (unsigned long long)l0 + ((unsigned long long)l1 << 18) + ((unsigned long long)l2 << 36);
        N[j] = (double)n;
        S[j] = (double)(long long)s * (1.0 / 67108864.0);
    }
    st_add_totals(acc_t, S, N, lane);
}

// One row against one candidate, the reference's arithmetic throughout (candidates beyond the register sets, segments in which
// the kd-tree neighbour test can matter, use_exact frames): reduced over the warp and added to the target at once.
__device__ __noinline__ void st_slow_row(float4 q, const StTarget* st, const StShared* sh, double* acc_t)
{
    const int lane = threadIdx.x & 31;
    bool fl[6] = {false, false, false, false, false, false};
    if (q.x != kSentinelXY && is_neighbour(st->tqx, st->tqy, sh->rn2, q.x, q.y, q.w)) {
        const ExactClass c = classify_exact(q.x, q.y, sh->pc, st->refx, st->refy, sh->rg2d, sh->rc2d);
        if (c.in) {
            fl[0] = true; fl[1] = c.down; fl[2] = c.right;
            if (c.cost) { fl[3] = true; fl[4] = c.down; fl[5] = c.right; }
        }
    }
    if (!__any_sync(0xffffffffu, fl[0])) return;
    const double I = (double)q.z;
    double S[6], N[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        S[j] = warp_sum(fl[j] ? I : 0.0);
        S[j] = __shfl_sync(0xffffffffu, S[j], 0);
        N[j] = (double)__popc(__ballot_sync(0xffffffffu, fl[j]));
    }
    st_add_totals(acc_t, S, N, lane);
}

__device__ __forceinline__ bool st_box_live(float4 b, float cx, float cy, float reach2)
{
    const float ex = fmaxf(fmaxf(b.x - cx, cx - b.z), 0.f), ey = fmaxf(fmaxf(b.y - cy, cy - b.w), 0.f);
    return ex * ex + ey * ey <= reach2;          // an empty box (+big, -big) is never live
}

// The cold paths of a tile, out of line so that the per-tile loop of k_sweep_stream stays within the L0 instruction cache (an
// earlier build with both inlined, two unrolled variants and eight-row blocks spent a third of its samples waiting for
// instruction fetches: profiles/r02_stream_history.md).
struct Corr12 { double d0[6], d1[6]; };
__device__ __noinline__ Corr12 st_fix_tile(uint32_t lbase, unsigned live0, unsigned live1, const StShared* sh, const StTarget* s0,
                                           const StTarget* s1, float E, float E2)
{
    Corr12 c;
#pragma unroll
    for (int j = 0; j < 6; ++j) { c.d0[j] = 0.0; c.d1[j] = 0.0; }
    unsigned rows = live0 | live1;
    while (rows) {
        const int r = __ffs(rows) - 1;
        rows &= rows - 1;
        const float4 qq = lds_f4(lbase + (uint32_t)r * 512u);
        if ((live0 >> r) & 1u) {
            const Corr6 d = st_correct(qq, sh, s0, E, E2);
#pragma unroll
            for (int j = 0; j < 6; ++j) c.d0[j] += d.d[j];
        }
        if ((live1 >> r) & 1u) {
            const Corr6 d = st_correct(qq, sh, s1, E, E2);
#pragma unroll
            for (int j = 0; j < 6; ++j) c.d1[j] += d.d[j];
        }
    }
    return c;
}

// candidates beyond the register sets (all of them when the warp has none) on this warp's rows of a tile: row by row
__device__ __noinline__ void st_slow_tile(uint32_t tile, float4 b, bool mine, const StSeg* sg, const StTarget* tgt, const StShared* sh,
                                          double* acc_slot, int nt, int grp, bool has_sets, float reach2)
{
    const int lane = threadIdx.x & 31;
    const int n_cand = sg->n_cand, all_targets = sg->all_targets;
    const int first_slow = has_sets ? 4 : 0;
    if (!(n_cand > first_slow || all_targets)) return;
    const int n_it = all_targets ? nt : n_cand;
    for (int k = all_targets ? 0 : first_slow + grp; k < n_it; k += all_targets ? 1 : 2) {
        int t = k;
        if (all_targets) {
            // the four nearest belong to the group of their rank (and sit in its register sets, if it has any); the rest by parity
            int rank = -1;
            for (int z = 0; z < min(n_cand, 4); ++z) if (sg->cand[z] == t) rank = z;
            if ((rank >= 0 ? (rank & 1) : (t & 1)) != grp) continue;
            if (rank >= 0 && has_sets) continue;
        } else t = sg->cand[k];
        const float2 cc2 = *reinterpret_cast<const float2*>(&tgt[t].cx);
        unsigned live = __ballot_sync(0xffffffffu, mine && st_box_live(b, cc2.x, cc2.y, reach2));
        while (live) {
            const int r = __ffs(live) - 1;
            live &= live - 1;
            st_slow_row(lds_f4(tile + (uint32_t)r * 512u + (uint32_t)lane * 16u), &tgt[t], sh, acc_slot + (size_t)t * 16);
        }
    }
}

// One segment for one consumer warp: takes the next descriptor from the producer warp, picks the warp's register sets (its
// candidates of rank g and g + 2, their constants from the target table) and walks the segment's rows tile by tile. Out of line
// and self-contained on purpose: its registers are its own (the six + six accumulators, the constants of the hot loop), and the
// loop over the tiles is small enough for the L0 instruction cache — one copy of the point update per set, two rows per
// iteration, every cold path a call.
__device__ __noinline__ void st_segment(StShared* sh, const StTarget* __restrict__ tgt)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp & 1, par = warp >> 1;
    StShared::Warp& ws = sh->w[warp];
    const uint32_t bar_full = smem_addr(&sh->full[0]), bar_empty = smem_addr(&sh->empty[0]);
    const uint32_t bar_sfull = smem_addr(&sh->seg_full[0]), bar_sempty = smem_addr(&sh->seg_empty[0]);
#ifdef RCLC_ST_TIMING
    long long* dbg = sh->dbg[warp];
    const long long t_seg = clock64();
#endif
    // ---- the descriptor ----
    int seg_j = ws.seg_j;
    if (seg_j >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(bar_sempty + 8u * (seg_j % kStSegRing)); }           // done with the old one
    ++seg_j;
    mbar_wait<20>(bar_sfull + 8u * (seg_j % kStSegRing), (uint32_t)(seg_j / kStSegRing) & 1u);
    const StSeg* sg = &sh->seg[seg_j % kStSegRing];
    const int seg_end = sg->seg_end, n_cand = sg->n_cand, all_targets = sg->all_targets;
    const int r0 = sh->r0, r1 = sh->r1;
    const float rgf = sh->rgf, rn2 = sh->rn2, z2max = sh->z2max, reach2 = sh->reach2;
    int t0 = -1, t1 = -1;                 // targets of the two register sets (-1: none)
    if (!sh->frame_exact) {
        if (grp < n_cand) t0 = sg->cand[grp];
        if (grp + 2 < n_cand) t1 = sg->cand[grp + 2];
    }
    float nkx0 = 0.f, nky0 = 0.f, cx0 = 0.f, cy0 = 0.f, nkx1 = 0.f, nky1 = 0.f, cx1 = 0.f, cy1 = 0.f;
    {
        int masked = 0;
        if (t0 >= 0) {
            const StTarget& s = tgt[t0];
            nkx0 = s.nkx; nky0 = s.nky; cx0 = s.cx; cy0 = s.cy;
            masked |= !((rgf + s.disp) * (rgf + s.disp) + z2max < rn2 * 0.9999f);
        }
        if (t1 >= 0) {
            const StTarget& s = tgt[t1];
            nkx1 = s.nkx; nky1 = s.nky; cx1 = s.cx; cy1 = s.cy;
            masked |= !((rgf + s.disp) * (rgf + s.disp) + z2max < rn2 * 0.9999f);
        }
        // a segment in which the kd-tree neighbour test can matter (trial pose centimetres away from the pose the neighbour sets
        // were built at) takes the row-by-row routine, which applies it like the reference does
        if (masked) { t0 = -1; t1 = -1; }
    }
    const bool has_sets = t0 >= 0, has1 = t1 >= 0;
    const bool has_slow = n_cand > (has_sets ? 4 : 0) || all_targets;
    double* acc_slot = sh->acc_slot;
    const float r00 = sh->r00, r01 = sh->r01, r10 = sh->r10, r11 = sh->r11, rg2 = sh->rg2, rc2 = sh->rc2;
    const float E = sh->E, E2 = sh->E2;
    const uint32_t ring = sh->ring;
    int zlo;
    asm volatile("mov.b32 %0, 0;" : "=r"(zlo));
    ST_ADD(2, t_seg);

    double a0[6] = {0, 0, 0, 0, 0, 0}, a1[6] = {0, 0, 0, 0, 0, 0};
    float mA = 3.0e38f, mB = 3.0e38f;
    int rows_acc = 0;                     // rows a lane has added since the last flush (the riding count has room for kFlushRows)
    int cur = ws.cur, n_waited = ws.n_waited;
    auto flush = [&]() {
        ST_T(tf);
        if (rows_acc > 0) {
            st_flush(a0[0], a0[1], a0[2], a0[3], a0[4], a0[5], acc_slot + (size_t)t0 * 16);
            if (has1) st_flush(a1[0], a1[1], a1[2], a1[3], a1[4], a1[5], acc_slot + (size_t)t1 * 16);
#pragma unroll
            for (int j = 0; j < 6; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
        }
        rows_acc = 0;
        ST_ADD(3, tf);
    };
#pragma unroll 1
    while (cur < seg_end) {
        const int i = (cur - r0) / kStTileRows;
        if (i >= n_waited) {                                                  // wait for the tile (once per warp)
            ST_T(tw);
            mbar_wait<20>(bar_full + 8u * (i % kStStages), (uint32_t)(i / kStStages) & 1u);
            n_waited = i + 1;
            ST_ADD(1, tw);
#ifdef RCLC_ST_TIMING
            dbg[7] += 1;
#endif
        }
        const uint32_t tile = ring + (uint32_t)(i % kStStages) * kStStageBytes;
        const int tile_row0 = r0 + i * kStTileRows, tile_rows = min(kStTileRows, r1 - tile_row0);
        const int lr = cur - tile_row0, lim = min(tile_rows, seg_end - tile_row0);
        const int first = lr + ((par - lr) & (kStPhases - 1));                // this warp's rows: tile-local index = par mod kStPhases
        const bool mine = lane >= first && lane < lim && ((lane & (kStPhases - 1)) == par);
        if (has_sets || has_slow) {
            // the tile's row boxes, one per lane: which of this warp's rows can touch the discs of its sets at all
            const float4 b = lane < tile_rows ? lds_f4(tile + kStTileBytes + (uint32_t)lane * 16u) : make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
            if (has_sets) {
                const unsigned live0 = __ballot_sync(0xffffffffu, mine && st_box_live(b, cx0, cy0, reach2));
                const unsigned live1 = has1 ? __ballot_sync(0xffffffffu, mine && st_box_live(b, cx1, cy1, reach2)) : 0u;
                const uint32_t lbase = tile + (uint32_t)lane * 16u;
                auto one = [&](const float4& q, int r) {
                    const double v = (double)q.z + kCntUnit;
                    if ((live0 >> r) & 1u) st_point(a0, q, v, r00, r01, r10, r11, nkx0, nky0, rg2, rc2, zlo, mA, mB);
                    if ((live1 >> r) & 1u) st_point(a1, q, v, r00, r01, r10, r11, nkx1, nky1, rg2, rc2, zlo, mA, mB);
                };
                unsigned todo = live0 | live1;                                  // (bits of this warp's rows only)
#pragma unroll 1
                while (todo) {
                    const int ra = __ffs(todo) - 1; todo &= todo - 1;
                    const int rb = todo ? __ffs(todo) - 1 : ra; todo &= todo - 1;
                    const float4 qa = lds_f4(lbase + (uint32_t)ra * 512u), qb = lds_f4(lbase + (uint32_t)rb * 512u);
                    one(qa, ra);
                    if (rb != ra) one(qb, rb);
                }
                if (__any_sync(0xffffffffu, mA <= E2 || mB <= E)) {            // cold: a few % of the tiles
                    if (mA <= E2 || mB <= E) {
                        const Corr12 cc = st_fix_tile(lbase, live0, live1, sh, &tgt[t0], &tgt[has1 ? t1 : t0], E, E2);
#pragma unroll
                        for (int j = 0; j < 6; ++j) { a0[j] += cc.d0[j]; a1[j] += cc.d1[j]; }
                    }
                    mA = 3.0e38f;
                    mB = 3.0e38f;
                    __syncwarp();
                }
                rows_acc += (lim - first + kStPhases - 1) / kStPhases;
            }
            if (has_slow) st_slow_tile(tile, b, mine, sg, tgt, sh, acc_slot, sh->nt, grp, has_sets, reach2);
        }
        cur = tile_row0 + lim;
        if (lim == tile_rows) {                                               // the warp has passed the tile's last row: give it back
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8u * (i % kStStages));
        }
        if (rows_acc + kStTileRows / kStPhases > kFlushRows) flush();
    }
    flush();
    if (lane == 0) { ws.cur = cur; ws.n_waited = n_waited; ws.seg_j = seg_j; }
    __syncwarp();
#ifdef RCLC_ST_TIMING
    if (lane == 0) { dbg[4] += clock64() - t_seg; dbg[6] += 1; }
#endif
}

constexpr size_t st_smem_bytes(int nt, int ncells)
{
    const size_t ring = (size_t)kStStages * kStStageBytes;
    const size_t tail = (size_t)12 * nt * 8 + sizeof(LmStage) + 16;
    return 128 + (ring > tail ? ring : tail) + ((sizeof(StShared) + 15) & ~(size_t)15) + (size_t)nt * sizeof(StTarget) + (size_t)ncells * 8;
}

// Three CTAs of five warps per SM: at most four warps per scheduler, i.e. 128 registers per thread (st_segment needs ~110;
// four CTAs / 96 registers spilled the accumulators inside the loop over the tiles).
__global__ void __launch_bounds__(kStThreads, 3) k_sweep_stream(const __grid_constant__ BatchView bv)
{
    const int f = blockIdx.y, slot = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef RCLC_ST_TIMING
    const long long t_begin = clock64();
#endif
    const GridGeom& g = bv.geom;
    const FrameCtl& c = bv.ctl[f];
    const PoseCoef pc = c.pose[min(slot, kSpec - 1)];                          // (requested together with the control words)
    const int active = c.sweep_active, n_pose = c.n_pose, n_rows = c.n_rows;
    if (!active || slot >= n_pose) return;                                     // the whole CTA
    const int frame_exact = c.use_exact;
    const float drift = c.drift;

    extern __shared__ __align__(128) unsigned char st_dyn[];
    unsigned char* sp = reinterpret_cast<unsigned char*>(((uintptr_t)st_dyn + 127) & ~(uintptr_t)127);
    const size_t ring_bytes = (size_t)kStStages * kStStageBytes, tail_bytes = (size_t)12 * g.nt * 8 + sizeof(LmStage) + 16;
    unsigned char* ring_gen = sp;                                               sp += ring_bytes > tail_bytes ? ring_bytes : tail_bytes;
    StShared* sh = reinterpret_cast<StShared*>(sp);                             sp += (sizeof(StShared) + 15) & ~(size_t)15;
    StTarget* tgt = reinterpret_cast<StTarget*>(sp);                            sp += (size_t)g.nt * sizeof(StTarget);
    int2* cells = reinterpret_cast<int2*>(sp);                                  // {first row, rows} of every lattice cell
    const uint32_t ring = smem_addr(ring_gen);
    const uint32_t bar_full = smem_addr(&sh->full[0]), bar_empty = smem_addr(&sh->empty[0]);
    const uint32_t bar_sfull = smem_addr(&sh->seg_full[0]), bar_sempty = smem_addr(&sh->seg_empty[0]);

    const int r0 = (int)((int64_t)n_rows * blockIdx.x / gridDim.x), r1 = (int)((int64_t)n_rows * (blockIdx.x + 1) / gridDim.x);
    const int n_tiles = (r1 - r0 + kStTileRows - 1) / kStTileRows;
    const float4* src = bv.binned + (size_t)f * bv.frame_stride;
    const float4* box = bv.rowbox + (size_t)f * (bv.frame_stride / 32);
    const double rgd = sqrt(g.rg2);
    const float rgf = (float)rgd, reach = rgf + 1e-5f, reach2 = reach * reach;

    // ---- everybody: the cell table and the per-target constants of this pose (the producer warp first arms the barriers and sends
    //      the first copies on their way) ----
    if (warp == kStConsWarps && lane == 0) {
        for (int s = 0; s < kStStages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, kStConsWarps); }
        for (int s = 0; s < kStSegRing; ++s) { mbar_init(bar_sfull + 8u * s, 1); mbar_init(bar_sempty + 8u * s, kStConsWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sh->z2max_bits = 0u;
        sh->bc_bits = 0u;
    }
    auto issue = [&](int i) {
        const int s = i % kStStages, rows = min(kStTileRows, r1 - (r0 + i * kStTileRows));
        const uint32_t dst = ring + (uint32_t)s * kStStageBytes;
        mbar_expect_tx(bar_full + 8u * s, (uint32_t)rows * 528u);
        bulk_g2s(dst, src + (size_t)(r0 + i * kStTileRows) * 32, (uint32_t)rows * 512u, bar_full + 8u * s);
        bulk_g2s(dst + kStTileBytes, box + r0 + i * kStTileRows, (uint32_t)rows * 16u, bar_full + 8u * s);
    };
    if (warp == kStConsWarps && lane == 0) for (int i = 0; i < min(n_tiles, kStStages); ++i) issue(i);
    __syncthreads();                                                            // #1: barriers and zeroed maxima are visible
    {
        unsigned z2 = 0u;
        for (int cc = tid; cc < g.ncells; cc += kStThreads) {
            const int cnt = bv.cellcnt[(size_t)f * g.ncells + cc], start = bv.cellstart[(size_t)f * g.ncells + cc];
            cells[cc] = make_int2(start >> 5, cnt > 0 ? pad32(cnt) >> 5 : 0);
            z2 = max(z2, bv.cellz2[(size_t)f * g.ncells + cc]);
        }
        float bcmax = 0.f;
        const float rot = sqrtf(fmaxf(0.f, 2.f - 2.f * (float)pc.r00)), shift = (float)(fabs(pc.tx) + fabs(pc.ty));
        for (int t = tid; t < g.nt; t += kStThreads) {
            const double2 ref = bv.tgt_xy[t];
            const float2 tq = bv.tgt_xyf[t];
            // p' = R (p + t); w = p' - ref = R p - k, k = ref - R t; the disc centre in the stored frame is c = R^T ref - t
            const double kx = ref.x - (pc.r00 * pc.tx + pc.r01 * pc.ty), ky = ref.y - (pc.r10 * pc.tx + pc.r11 * pc.ty);
            const double cxd = pc.r00 * ref.x + pc.r10 * ref.y - pc.tx, cyd = pc.r01 * ref.x + pc.r11 * ref.y - pc.ty;
            StTarget st;
            st.nkx = (float)(-kx); st.nky = (float)(-ky);
            st.cx = (float)cxd; st.cy = (float)cyd;
            st.tqx = tq.x; st.tqy = tq.y;
            // how far the pose moves a point of this target's discs (an upper bound): decides whether the kd-tree test can matter
            const float pmax = 1.0001f * (fabsf(st.cx) + fabsf(st.cy)) + 2.f * rgf;
            st.disp = rot * pmax + shift + 1e-5f;
            st.pad = 0.f;
            st.refx = ref.x; st.refy = ref.y;
            bcmax = fmaxf(bcmax, (float)(fabs(cxd) + fabs(cyd) + fabs(pc.tx) + fabs(pc.ty) + fabs(ref.x) + fabs(ref.y) + 0.2));
            tgt[t] = st;
        }
        z2 = __reduce_max_sync(0xffffffffu, z2);
        const unsigned bcb = __reduce_max_sync(0xffffffffu, __float_as_uint(bcmax));
        if (lane == 0) { atomicMax(&sh->z2max_bits, z2); atomicMax(&sh->bc_bits, bcb); }
        if (tid == 0) {
            sh->pc = pc;
            sh->r00 = (float)pc.r00; sh->r01 = (float)pc.r01; sh->r10 = (float)pc.r10; sh->r11 = (float)pc.r11;
            sh->rg2 = (float)g.rg2; sh->rc2 = (float)g.rc2; sh->rn2 = g.rn2f; sh->rgf = rgf;
            sh->rg2d = g.rg2; sh->rc2d = g.rc2;
            sh->ring = ring; sh->r0 = r0; sh->r1 = r1; sh->nt = g.nt; sh->frame_exact = frame_exact; sh->reach2 = reach2;
            sh->acc_slot = bv.acc + (size_t)f * bv.acc_stride + (size_t)slot * g.nt * 16;
        }
        if (tid < kStConsWarps) {
            sh->w[tid].cur = r0; sh->w[tid].n_waited = 0; sh->w[tid].seg_j = -1;
#ifdef RCLC_ST_TIMING
            for (int k = 0; k < kStDbgSlots; ++k) sh->dbg[tid][k] = 0;
#endif
        }
    }
    __syncthreads();                                                            // #2: tables complete
    if (tid == 0) {
        // error bounds of the fp32 geometry, for the largest coordinate magnitude any target of this pose sees (gridfit.cu, k.E / k.E2)
        const float E = (float)(6.0 * 5.9604644775390625e-8 * (double)__uint_as_float(sh->bc_bits) + 1e-9);
        sh->E = E;
        sh->E2 = 1.05f * (2.9f * rgf * E + 3e-7f * sh->rg2 + 2.f * E * E + 1e-9f);
        sh->z2max = __uint_as_float(sh->z2max_bits);
    }
    cons_bar_sync_all();                                                        // #3: (everybody) the bounds are there

    // ---- producer warp: keeps the ring of tiles full (lane 0) and, ahead of the consumers, works out the segments: which lattice
    //      cell the next rows belong to and which targets' discs can reach it, nearest first (all lanes) ----
    if (warp == kStConsWarps) {
        int next_tile = min(n_tiles, kStStages), seg_i = 0, cell = -1, row = r0;
        while (next_tile < n_tiles || row < r1) {
            bool progressed = false;
            if (row < r1) {
                const int k = seg_i % kStSegRing;
                int ok = 1;
                if (seg_i >= kStSegRing) {
                    if (lane == 0) ok = mbar_try(bar_sempty + 8u * k, (uint32_t)((seg_i / kStSegRing) - 1) & 1u);
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                }
                if (ok) {
                    do { ++cell; } while (cell < g.ncells && (cells[cell].y == 0 || cells[cell].x + cells[cell].y <= row));
                    const int seg_end = min(cells[cell].x + cells[cell].y, r1);          // (row < n_rows, so a cell is found)
                    const int ci = cell % g.ncx, cj = cell / g.ncx;
                    // the cell's rectangle, open-ended on the lattice border (points beyond the box are filed in the border cells with
                    // their own coordinates), widened by how far points may have drifted out of the cells they were sorted into
                    const float gx0 = (float)((g.i0 + ci) * g.g), gy0 = (float)((g.j0 + cj) * g.g), gs = (float)g.g;
                    const float slack = drift + 2e-5f + 4e-7f * (fabsf(gx0) + fabsf(gy0) + gs);
                    const float x0 = ci == 0 ? -3.0e38f : gx0 - slack, x1 = ci == g.ncx - 1 ? 3.0e38f : gx0 + gs + slack;
                    const float y0 = cj == 0 ? -3.0e38f : gy0 - slack, y1 = cj == g.ncy - 1 ? 3.0e38f : gy0 + gs + slack;
                    const float mx = gx0 + 0.5f * gs, my = gy0 + 0.5f * gs;
                    // key = distance of the disc centre to the cell centre (low 10 bits dropped) | target, + 1
                    auto key_of = [&](int t) -> unsigned {
                        if (t >= g.nt) return 0xffffffffu;
                        const float2 cc2 = *reinterpret_cast<const float2*>(&tgt[t].cx);
                        const float ex = fmaxf(fmaxf(x0 - cc2.x, cc2.x - x1), 0.f), ey = fmaxf(fmaxf(y0 - cc2.y, cc2.y - y1), 0.f);
                        if (!(ex * ex + ey * ey <= reach2)) return 0xffffffffu;
                        const float dx = cc2.x - mx, dy = cc2.y - my;
                        return ((__float_as_uint(dx * dx + dy * dy) & 0xfffffc00u) | (unsigned)t) + 1u;
                    };
                    StSeg& sg = sh->seg[k];
                    int n = 0;
                    if (g.nt <= 96) {                                      // three targets per lane: keys held in registers
                        unsigned k0 = key_of(lane), k1 = key_of(lane + 32), k2 = key_of(lane + 64);
                        for (; n <= kStMaxCand; ++n) {
                            const unsigned mine = min(k0, min(k1, k2));
                            const unsigned best = __reduce_min_sync(0xffffffffu, mine);
                            if (best == 0xffffffffu) break;
                            if (mine == best) {
                                if (n < kStMaxCand) sg.cand[n] = (unsigned short)((best - 1u) & 1023u);
                                if (k0 == best) k0 = 0xffffffffu; else if (k1 == best) k1 = 0xffffffffu; else k2 = 0xffffffffu;
                            }
                        }
                    } else {
                        unsigned last = 0u;
                        for (; n <= kStMaxCand; ++n) {
                            unsigned best = 0xffffffffu;
                            for (int t = lane; t < g.nt; t += 32) { const unsigned key = key_of(t); if (key > last && key < best) best = key; }
                            best = __reduce_min_sync(0xffffffffu, best);
                            if (best == 0xffffffffu) break;
                            if (n < kStMaxCand && lane == 0) sg.cand[n] = (unsigned short)((best - 1u) & 1023u);
                            last = best;
                        }
                    }
                    if (lane == 0) { sg.seg_end = seg_end; sg.n_cand = min(n, kStMaxCand); sg.all_targets = n > kStMaxCand; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_sfull + 8u * k);
                    row = seg_end;
                    ++seg_i;
                    progressed = true;
                }
            }
            if (next_tile < n_tiles) {
                int ok = 0;
                if (lane == 0) {
                    ok = mbar_try(bar_empty + 8u * (next_tile % kStStages), (uint32_t)((next_tile / kStStages) - 1) & 1u);
                    if (ok) issue(next_tile);
                }
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) { ++next_tile; progressed = true; }
            }
            if (!progressed) __nanosleep(200);
        }
        return;
    }

    // ---- consumers: segment after segment (st_segment) ----
    {
        StShared::Warp& ws = sh->w[warp];
        while (ws.cur < r1) st_segment(sh, tgt);
    }
    double* acc_slot = sh->acc_slot;
#ifdef RCLC_ST_TIMING
    if (lane == 0) {
        long long* dbg = sh->dbg[warp];
        dbg[5] = clock64() - t_begin;
        const int w = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kStConsWarps + warp;
        if (w < 4096) for (int k = 0; k < kStDbgSlots; ++k) g_st_dbg[w * kStDbgSlots + k] = dbg[k];
    }
#endif

    // ---- completion: the last CTA of the (frame, pose slot) turns the accumulators into the LM terms; the last slot of the frame
    //      runs the LM step ----
    if (!bv.fuse_lm) return;
    fence_acq_rel_gpu();
    cons_bar_sync();
    if (tid == 0) {
        fence_acq_rel_gpu();
        sh->last = atomicAdd(&bv.slot_done[(size_t)f * kSpec + slot], 1) == (int)gridDim.x - 1;
    }
    cons_bar_sync();
    if (!sh->last) return;
    fence_acq_rel_gpu();
    if (tid == 0) bv.slot_done[(size_t)f * kSpec + slot] = 0;
    double* s_terms = reinterpret_cast<double*>(ring_gen);                      // [12][nt] — the ring is idle now
    LmStage* stage = reinterpret_cast<LmStage*>(ring_gen + (((size_t)12 * g.nt * 8 + 15) & ~(size_t)15));
    {
        const double theta = (pc.yaw_deg * 3.14159265358979323846) / 180.0;
        const double cth = cos(theta), sth = sin(theta);
        for (int t = tid; t < g.nt; t += kStConsThreads) {
            double* acc_t = acc_slot + (size_t)t * 16;
            double a16[16];
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                const double2 v2 = __ldcg(reinterpret_cast<const double2*>(acc_t + q));
                a16[q] = v2.x; a16[q + 1] = v2.y;
            }
            double tt[12];
            target_terms(a16, bv.tgt_is_row[t] != 0, pc, cth, sth, g.huber_a, tt);
#pragma unroll
            for (int q = 0; q < 12; ++q) s_terms[(size_t)q * g.nt + t] = tt[q];
#pragma unroll
            for (int q = 0; q < 16; q += 2) *reinterpret_cast<double2*>(acc_t + q) = make_double2(0.0, 0.0);      // for the next sweep
        }
    }
    cons_bar_sync();
    if (warp != 0) return;
    {
        // slot_sums' order: lane-strided over the targets, then the butterfly
        const int nt = g.nt;
        double sums[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) sums[k] = 0.0;
        for (int tb = 0; tb < nt; tb += 96) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int t = tb + lane + 32 * u;
#pragma unroll
                for (int k = 0; k < 12; ++k) sums[k] += t < nt ? s_terms[(size_t)k * nt + t] : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) sums[k] = warp_sum(sums[k]);
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) if (lane == k) mine = sums[k];
        if (lane < 12) bv.evsum[((size_t)f * kSpec + slot) * 12 + lane] = mine;
    }
    fence_acq_rel_gpu();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&bv.sweep_done[f], 1) == n_pose - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    fence_acq_rel_gpu();
    if (lane == 0) bv.sweep_done[f] = 0;
    lm_frame_step(bv, f, stage);
}
