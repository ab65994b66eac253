// chain_stream.cuh -- K3, the streaming kernels for Processor runs without FIR and without resampler:
//     [gain* | copy] [biquad]? [gain*]
// i.e. BASELINE.json configs[0] (mock.Processor copy, reference mock/mock.go:147-154) and configs[1] (gain + biquad).
// Like K1 they replace the ProcessFunc walk of Processor.execute (reference pipe.go:425-451, the call at :438) for the
// run; unlike K1 these runs are HBM-bound (8 B per f32 sample, 16 B per f64 sample), so the kernels are built around
// bytes in flight, resident CTAs and the latency of the look-back:
//
//   stream_map_kernel     out = g * in over the flat buffer (copy / gain-only runs without meter): 16 B vectors,
//                         four independent loads per thread before the first store, streaming cache hints.
//   chain_stream_kernel   gain + biquad (+ fused meter sink).  A tile is 32 channels x 8 R frames (R = 32 for f32, 16 for
//                         f64: 32 KB in, 32 KB out); the R rows of each of the 8 warps go from HBM straight into the warp's
//                         slab of shared memory (cp.async, 16 B per lane, no register staging: 48 registers, 5 CTAs per SM),
//                         are read from there by both passes and leave as one coalesced 128 B / 256 B store per row, so
//                         one HBM read and one HBM write per sample is all the traffic there is (keeping the rows in
//                         registers instead, 3 CTAs per SM, measured slower).  The biquad (TDF-II, double, the same carried
//                         state [C][2] as K1/K2) is parallelised in time as in K1: zero-state end state of every R-row
//                         sub-chunk (2 independent DFMA per sample), tile aggregate, decoupled look-back across tiles, the
//                         true recursion from the resolved state (5 DFMA per sample).  What differs from K1, because the
//                         look-back latency is what bounds an HBM-bound tile kernel (measured: 60-80 % of a tile's time
//                         with one warp walking predecessors while the others wait at a barrier):
//                           * sub-chunk prefix states are computed by every warp in parallel from a table of A^(R j);
//                           * ALL 8 warps look back, each over one window of 32 predecessors [32 w, 32 w + 32): each
//                             folds its window (independent FMAs against a table of (A^T)^i) and the partial results are
//                             combined up to the first window that held an inclusive state.  A tile therefore never waits
//                             for an inclusive state to appear, only for aggregates, which depend on nothing -- with few
//                             channel groups (configs[1] has 2) more than a hundred tiles of one group are in flight;
//                             the last warp keeps sliding if 256 predecessors hold aggregates only;
//                           * one release per publication and no fence after the acquire (the warp barrier orders the
//                             lanes): MEMBARs were half of the look-back time;
//                           * full tiles take a path without per-row predicates, tables with static indices sit in the
//                             kernel parameters (constant-bank operands of the DFMAs): ~20 instead of ~78 instructions
//                             per sample.
#pragma once

#include "chain_tile.cuh"

namespace pb {

// Tuning history (profiles/r01_k3_summary.md; the measured-slower variants live outside the shipped kernel now): the tile waits
// between its two passes in shared memory (cp.async, 5 CTAs per SM: 65 % of the HBM peak at 1024 ch; in registers 57 %; two tiles
// with the next one's copies in flight 40 %, because a CTA that holds a ticket it is not yet working on publishes that tile's
// aggregate a whole tile late); every warp walks one look-back window (a two-level walk over blocks of 32 tiles read 1/32 of the
// payloads and still measured slower: the payload loads are not what a tile waits for).
#ifndef PB_ST_ROWS32
#define PB_ST_ROWS32 32   // rows per warp for f32
#endif
#ifndef PB_ST_MINB
#define PB_ST_MINB 5      // resident CTAs per SM the register budget is set for
#endif
// How the tiles of a launch learn their incoming biquad state:
//   kStOneSweep   decoupled look-back inside one launch (8 B of HBM traffic per f32 sample).  Right when there are many channel
//                 groups: the tiles in flight spread over them and the walk is short (1024 ch: 65 % of the HBM peak).
//   stream_aggregate_kernel + stream_scan_kernel + kStApply ("two sweeps")   with FEW channel groups (configs[1]: 64 ch = 2 groups) hundreds
//                 of tiles of one group are in flight, every tile folds up to 256 predecessors and the launch sits at 33 % of the
//                 peak waiting for them.  Reading the input twice costs 12 B per sample but nothing ever waits: sweep 1 stores
//                 every tile's aggregate, a one-CTA-per-group scan turns them into inclusive states, sweep 2 (in REVERSE tile
//                 order: the end of the batch is what sweep 1 left in L2) runs the recursion from the resolved states.
enum : int { kStOneSweep = 0, kStApply = 2 };
constexpr int kStTicketsPerCta = 1;  // tickets a CTA draws past the end of the batch
constexpr int kStThreads = 256;
constexpr int kStWarps = kStThreads / 32;        // 8 sub-chunks per tile, and 8 look-back windows
constexpr int kStWin = 32;                       // look-back window (one predecessor per lane)
template <typename T>
struct StShape {
    static constexpr int kRows = sizeof(T) == 4 ? PB_ST_ROWS32 : 16;   // rows a warp keeps in registers
    static constexpr int kTile = kStWarps * kRows;            // frames per tile: 256 (f32) / 128 (f64)
};
constexpr int kStMinTile = 128;

// double tables in global memory, copied to shared memory at kernel start (dynamic indices)
struct StTab {
    static constexpr int kPw = 0;                          // [9][4]   A^(R j), j = 0..8  (j = 8: the tile step A^T)
    static constexpr int kLb = kPw + 4 * (kStWarps + 1);   // [33][4]  (A^T)^i, i = 0..32
    static constexpr int kMw = kLb + 4 * (kStWin + 1);     // [33][4]  (A^T)^(32 i), i = 0..32
    static constexpr int kCount = kMw + 4 * (kStWin + 1);
};

template <typename T>
struct StreamParams {
    const T *in;
    T *out;
    int64_t n_frames;
    int C, n_tiles, n_groups;
    T g_load;             // all gains of a run without biquad, applied in T like K1; 1 for a biquad run, whose leading gains
                          // the host folds into wt and b0, b1, b2 (in double, like the oracle's y = g x)
    double g_bq;          // gains behind the biquad, applied to the double result before the single rounding to T
    int has_bq;
    double b0, b1, b2, a1, a2;
    double wt[32][2];     // A^k B, k < R: constant-bank operands (static indices after unrolling)
    const double *tab;    // StTab
    const double *bq_state;
    double *bq_state_next;
    double *lb_agg, *lb_inc;
    unsigned *lb_status;
    unsigned epoch;
    double *meter_peak, *meter_sumsq;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    int *err_flag;
    int vec_ok;           // 16-byte copies usable: C a multiple of the vector width and `in` 16-byte aligned
};

#ifdef __CUDACC__

__device__ __forceinline__ void mat2_fma(const double *__restrict__ m, double v0, double v1, double &a0, double &a1)
{
    a0 = fma(m[0], v0, fma(m[1], v1, a0));
    a1 = fma(m[2], v0, fma(m[3], v1, a1));
}

// f32 -> f64 for pass 2.  The rows are converted once per pass ON PURPOSE: written as a plain cast the compiler merges the two
// conversions of a row and keeps 32 doubles (64 registers) alive across the look-back, which spills half of the tile.
__device__ __forceinline__ double to_double_again(float v)
{
    double d;
    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ double to_double_again(double v) { return v; }

// One look-back window: lane i watches tile base - i of group status array st_g.  Returns the position of the first
// inclusive state (0..31), 32 when the window holds aggregates only, -1 on timeout.  Lanes before the stream start
// (base - i < 0) count as inclusive; they never come first because tile 0 publishes an inclusive state.
__device__ __forceinline__ int st_poll_window(const unsigned *st_g, int base, int lane, unsigned epoch, int *err_flag)
{
    const int j = base - lane;
    for (unsigned spins = 0;; spins++) {
        unsigned st = kLbInc;
        if (j >= 0) {
            st = ld_acquire_u32(st_g + j);
            st = ((st >> 2) == epoch) ? (st & 3u) : kLbNone;
        }
        const unsigned ready = __ballot_sync(0xffffffffu, st != kLbNone);
        const unsigned inc = __ballot_sync(0xffffffffu, st == kLbInc);
        if (inc) {
            const int first_inc = __ffs(inc) - 1;
            const unsigned need = (first_inc == 0) ? 0u : (0xffffffffu >> (32 - first_inc));
            if ((ready & need) == need) return first_inc;
        } else if (ready == 0xffffffffu) {
            return kStWin;
        }
        if (spins > (1u << 24)) {  // ~1 s: a predecessor never published
            if (lane == 0) atomicExch(err_flag, 1);
            return -1;
        }
        __nanosleep(32);
    }
}

// w += sum_{i < first_inc} (A^T)^i Z_{base-i}  (+ (A^T)^first_inc Inc_{base-first_inc} when first_inc < 32), this lane's channel
__device__ __forceinline__ void st_fold_window(const double *agg_g, const double *inc_g, const double *lb_s, int base, int first_inc,
                                               int lane, double &w0, double &w1)
{
#pragma unroll 1
    for (int i0 = 0; i0 < first_inc; i0 += 4) {  // payloads four at a time: one L2 round trip per batch is exposed
        double2 a[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = (i0 + u < first_inc) ? i0 + u : first_inc - 1;
            a[u] = __ldcg(reinterpret_cast<const double2 *>(agg_g + (size_t)(base - i) * 64 + lane * 2));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u < first_inc) mat2_fma(lb_s + 4 * (i0 + u), a[u].x, a[u].y, w0, w1);
    }
    if (first_inc < kStWin) {
        const double2 q = __ldcg(reinterpret_cast<const double2 *>(inc_g + (size_t)(base - first_inc) * 64 + lane * 2));
        mat2_fma(lb_s + 4 * first_inc, q.x, q.y, w0, w1);
    }
}

// 256 predecessors with aggregates only: the last warp keeps sliding window by window (tile 0 ends the walk).  Rare, and kept
// out of line so that its registers do not count against the kernel's.
__device__ __noinline__ double2 st_slide(const unsigned *st_g, const double *agg_g, const double *inc_g, const double *lb_s, int base,
                                         int lane, unsigned epoch, int *err_flag, double w0, double w1)
{
    double M[4] = {1.0, 0.0, 0.0, 1.0};  // (A^T)^(32 windows walked)
    const double *ML = lb_s + 4 * kStWin;
    for (;;) {
        base -= kStWin;
        const double n0 = M[0] * ML[0] + M[1] * ML[2], n1 = M[0] * ML[1] + M[1] * ML[3];
        const double n2 = M[2] * ML[0] + M[3] * ML[2], n3 = M[2] * ML[1] + M[3] * ML[3];
        M[0] = n0; M[1] = n1; M[2] = n2; M[3] = n3;
        const int first_inc = st_poll_window(st_g, base, lane, epoch, err_flag);
        __syncwarp();
        double v0 = 0.0, v1 = 0.0;
        if (first_inc >= 0) st_fold_window(agg_g, inc_g, lb_s, base, first_inc, lane, v0, v1);
        mat2_fma(M, v0, v1, w0, w1);
        if (first_inc < kStWin) return make_double2(w0, w1);
    }
}

// This warp's rows of `tile` go straight from HBM into its slab of shared memory (cp.async, no register staging): 16 B per
// lane, four (f64: two) rows per instruction; rows past the end and channels past C are zero-filled (src-size 0).
template <typename T>
__device__ __forceinline__ void st_issue_tile(const StreamParams<T> &p, int C, int tile, T *xs_w, int warp, int lane)
{
    constexpr int R = StShape<T>::kRows, kTile = StShape<T>::kTile;
    const int t = tile / p.n_groups, g = tile - t * p.n_groups;
    const int64_t f0 = (int64_t)t * kTile, ld = C;
    const int len = (int)((f0 + kTile < p.n_frames) ? kTile : p.n_frames - f0);
    const int r0 = warp * R;
    const int nrow = len - r0 < 0 ? 0 : (len - r0 > R ? R : len - r0);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(xs_w);
    if (p.vec_ok) {
        constexpr int EPC = 16 / (int)sizeof(T), CPR = kCg / EPC, RPI = 32 / CPR;
        const int rr = lane / CPR, ch = (lane % CPR) * EPC;
        const bool chv = g * kCg + ch < C;
        const T *sp0 = p.in + (f0 + r0 + rr) * ld + g * kCg + ch;
#pragma unroll
        for (int k = 0; k < R / RPI; k++) {
            const int row = k * RPI + rr;
            const bool ok = chv && row < nrow;
            const T *sp = ok ? sp0 + (int64_t)(k * RPI) * ld : p.in;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + (uint32_t)((row * kCg + ch) * sizeof(T))), "l"(sp),
                         "r"(ok ? 16 : 0)
                         : "memory");
        }
    } else {
        const int c = g * kCg + lane;
#pragma unroll
        for (int i = 0; i < R; i++) {
            const bool ok = c < C && i < nrow;
            const T *sp = ok ? p.in + (f0 + r0 + i) * ld + c : p.in;
            if (sizeof(T) == 4)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sbase + (uint32_t)((i * kCg + lane) * sizeof(T))), "l"(sp),
                             "r"(ok ? 4 : 0)
                             : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sbase + (uint32_t)((i * kCg + lane) * sizeof(T))), "l"(sp),
                             "r"(ok ? 8 : 0)
                             : "memory");
        }
    }
}

// CC: the channel count when it is one of the instantiated constants (then every row address is base + immediate: the
// 64-bit address arithmetic per row was a quarter of all instructions), 0 for any other count.  MODE: see kStOneSweep.
template <typename T, int CC, int MODE>
__global__ void __launch_bounds__(kStThreads, PB_ST_MINB) chain_stream_kernel(const __grid_constant__ StreamParams<T> p)
{
    constexpr int R = StShape<T>::kRows, kTile = StShape<T>::kTile;
    __shared__ double tab_s[StTab::kCount];
    __shared__ double zq_s[kStWarps * kCg * 2];    // zero-state end state of every sub-chunk; reused for the meter partials
    __shared__ double part_s[kStWarps * kCg * 2];  // look-back partial of every window
    __shared__ double zsum_s[kCg * 2];             // tile aggregate, parked by warp 0 across the look-back
    __shared__ int flag_s[kStWarps];               // window w held an inclusive state (the combination stops there)
    __shared__ int s_tile;
    __shared__ __align__(16) T xs[kStWarps * R * kCg];  // 32 KB: the tile, one slab of R rows per warp

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = CC ? CC : p.C;
    const int total_tiles = p.n_tiles * p.n_groups;
    const bool meter = p.meter_peak != nullptr;
    if (p.has_bq)
        for (int i = tid; i < StTab::kCount; i += kStThreads) tab_s[i] = p.tab[i];
    const double *pw_s = tab_s + StTab::kPw, *lb_s = tab_s + StTab::kLb, *mw_s = tab_s + StTab::kMw;

    for (int it = 0;; it++) {
        __syncthreads();  // previous tile done with shared memory (and the tables visible)
        // One sweep: time-major tickets -- a tile only waits on smaller tickets, which resident CTAs already hold.  Second of two
        // sweeps: nothing waits, so the tiles are dealt out statically (a ticket is an atomic round trip in front of every tile's
        // loads) and the batch is walked backwards (see kStApply).
        int tile;
        if (MODE == kStApply) {
            const int k = (int)blockIdx.x + it * (int)gridDim.x;
            if (k >= total_tiles) break;
            tile = total_tiles - 1 - k;
        } else {
            if (tid == 0) s_tile = (int)(atomicAdd(p.ticket, 1ULL) - p.ticket_base);
            __syncthreads();
            if (s_tile >= total_tiles) break;
            tile = s_tile;
        }
        T *xs_w = xs + warp * R * kCg;
        st_issue_tile<T>(p, C, tile, xs_w, warp, lane);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();  // a row is read by other lanes than the ones that copied it
        const int t = tile / p.n_groups, g = tile - t * p.n_groups;
        const bool first = (t == 0), last = (t == p.n_tiles - 1);
        const int64_t f0 = (int64_t)t * kTile;
        const int len = (int)((f0 + kTile < p.n_frames) ? kTile : p.n_frames - f0);
        const int c = g * kCg + lane;
        const bool cvalid = c < C;
        const int r0 = warp * R;
        const int nrow = len - r0 < 0 ? 0 : (len - r0 > R ? R : len - r0);
        const bool full = (nrow == R) && (g * kCg + kCg <= C);  // warp-uniform: no per-row predicates
        const int64_t ld = C;  // a compile-time constant when CC != 0

        T *dst = p.out + (f0 + r0) * ld + c;
        const T gl = p.g_load;  // the leading gains are applied when a row is read
#define PB_XV(i) (xs_w[(i) * kCg + lane] * gl)
#define PB_XR(i) xs_w[(i) * kCg + lane]  // biquad runs: the host folds the leading gains into wt and b0..b2 (one FMUL per read less)
        double m_peak = 0.0, m_sumsq = 0.0;

        if (p.has_bq) {
            // ---- pass 1: zero-state end state of the sub-chunk (the zero rows past the end of a ragged chunk leave its sum
            //      unused: only full tiles are chained, and the stream's final state comes out of pass 2)
            {
                // four accumulator pairs: a dependent DFMA costs ~100 cycles here, the chains must be short
                double z0 = 0.0, z1 = 0.0, y0 = 0.0, y1 = 0.0, u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int i = 0; i < R; i += 4) {
                    const double xa = (double)PB_XR(i), xb = (double)PB_XR(i + 1), xc = (double)PB_XR(i + 2), xd = (double)PB_XR(i + 3);
                    z0 = fma(p.wt[R - 1 - i][0], xa, z0);
                    z1 = fma(p.wt[R - 1 - i][1], xa, z1);
                    y0 = fma(p.wt[R - 2 - i][0], xb, y0);
                    y1 = fma(p.wt[R - 2 - i][1], xb, y1);
                    u0 = fma(p.wt[R - 3 - i][0], xc, u0);
                    u1 = fma(p.wt[R - 3 - i][1], xc, u1);
                    v0 = fma(p.wt[R - 4 - i][0], xd, v0);
                    v1 = fma(p.wt[R - 4 - i][1], xd, v1);
                }
                *reinterpret_cast<double2 *>(zq_s + (warp * kCg + lane) * 2) = make_double2((z0 + y0) + (u0 + v0), (z1 + y1) + (u1 + v1));
            }
            __syncthreads();
            const size_t slot = (size_t)g * p.n_tiles + t;
            if (MODE != kStApply && warp == 0) {
                // ---- warp 0: the tile aggregate sum_q A^(R (7-q)) z_q, published at once (it depends on nothing) and parked in
                //      shared memory until the inclusive state is due
                double Z0 = 0.0, Z1 = 0.0;
#pragma unroll
                for (int q = 0; q < kStWarps; q++) {
                    const double2 z = *reinterpret_cast<const double2 *>(zq_s + (q * kCg + lane) * 2);
                    mat2_fma(pw_s + 4 * (kStWarps - 1 - q), z.x, z.y, Z0, Z1);
                }
                *reinterpret_cast<double2 *>(zsum_s + lane * 2) = make_double2(Z0, Z1);
                if (!last && !first) {
                    // payload by every lane, then ONE release by lane 0: the warp barrier orders the lanes' stores before it
                    *reinterpret_cast<double2 *>(p.lb_agg + slot * 64 + lane * 2) = make_double2(Z0, Z1);
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
                }
            }
            double S0 = 0.0, S1 = 0.0;  // incoming state of the tile
            if (MODE == kStApply) {
                // the scan left the inclusive state after every full tile in lb_inc
                if (first) {
                    if (cvalid) {
                        S0 = p.bq_state[2 * c];
                        S1 = p.bq_state[2 * c + 1];
                    }
                } else {
                    const double2 q = __ldcg(reinterpret_cast<const double2 *>(p.lb_inc + (slot - 1) * 64 + lane * 2));
                    S0 = q.x;
                    S1 = q.y;
                }
            } else {
                // ---- look-back: warp w resolves window w of this group's predecessors (rotating the windows so that warp 0,
                //      which has just paid for the aggregate and its release, takes the farthest one measured no gain)
                {
                    const int win = warp;  // tiles t-1-32 win .. t-32-32 win
                    double w0 = 0.0, w1 = 0.0;
                    int terminal = 1;
                    if (first) {
                        if (win == 0 && cvalid) {
                            w0 = p.bq_state[2 * c];
                            w1 = p.bq_state[2 * c + 1];
                        }
                    } else {
                        const unsigned *st_g = p.lb_status + (size_t)g * p.n_tiles;
                        const double *agg_g = p.lb_agg + (size_t)g * p.n_tiles * 64, *inc_g = p.lb_inc + (size_t)g * p.n_tiles * 64;
                        const int base = t - 1 - kStWin * win;
                        if (base >= 0) {
                            const int first_inc = st_poll_window(st_g, base, lane, p.epoch, p.err_flag);
                            __syncwarp();  // the acquires of all lanes are ordered before every lane's payload loads
                            if (first_inc >= 0) st_fold_window(agg_g, inc_g, lb_s, base, first_inc, lane, w0, w1);
                            terminal = first_inc < kStWin;
                            if (win == kStWarps - 1 && !terminal) {
                                const double2 r = st_slide(st_g, agg_g, inc_g, lb_s, base, lane, p.epoch, p.err_flag, w0, w1);
                                w0 = r.x;
                                w1 = r.y;
                                terminal = 1;
                            }
                        }
                    }
                    *reinterpret_cast<double2 *>(part_s + (win * kCg + lane) * 2) = make_double2(w0, w1);
                    if (lane == 0) flag_s[win] = terminal;
                }
                __syncthreads();
                // ---- incoming state of the tile: the windows' partials up to the first one that held an inclusive state
#pragma unroll 1
                for (int w = 0; w < kStWarps; w++) {
                    const double2 v = *reinterpret_cast<const double2 *>(part_s + (w * kCg + lane) * 2);
                    mat2_fma(mw_s + 4 * w, v.x, v.y, S0, S1);
                    if (flag_s[w]) break;
                }
                if (warp == 0 && !last) {
                    // inclusive state after this (full) tile
                    const double2 Z = *reinterpret_cast<const double2 *>(zsum_s + lane * 2);
                    double I0 = Z.x, I1 = Z.y;
                    mat2_fma(pw_s + 4 * kStWarps, S0, S1, I0, I1);
                    *reinterpret_cast<double2 *>(p.lb_inc + slot * 64 + lane * 2) = make_double2(I0, I1);
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
                }
            }
            // ---- pass 2: the recursion itself from the true state at the first row of the sub-chunk,
            //      A^(R w) S + sum_{q<w} A^(R (w-1-q)) z_q
            double s1 = 0.0, s2 = 0.0;
            mat2_fma(pw_s + 4 * warp, S0, S1, s1, s2);
#pragma unroll 1
            for (int q = 0; q < warp; q++) {
                const double2 z = *reinterpret_cast<const double2 *>(zq_s + (q * kCg + lane) * 2);
                mat2_fma(pw_s + 4 * (warp - 1 - q), z.x, z.y, s1, s2);
            }
            if (full && !meter && !last) {
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const double xd = to_double_again(PB_XR(i));
                    const double v = fma(p.b0, xd, s1);
                    s1 = fma(-p.a1, v, fma(p.b1, xd, s2));
                    s2 = fma(-p.a2, v, p.b2 * xd);
                    __stcs(dst + i * ld, (T)(v * p.g_bq));
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const double xd = to_double_again(PB_XR(i));
                    const double v = fma(p.b0, xd, s1);
                    s1 = fma(-p.a1, v, fma(p.b1, xd, s2));
                    s2 = fma(-p.a2, v, p.b2 * xd);
                    const T o = (T)(v * p.g_bq);
                    if (i < nrow && cvalid) {  // rows past the end neither store nor reach the carried state
                        __stcs(dst + i * ld, o);
                        if (meter) {
                            const double a = fabs((double)o);
                            m_peak = a > m_peak ? a : m_peak;
                            m_sumsq = fma((double)o, (double)o, m_sumsq);
                        }
                        if (last && r0 + i + 1 == len) {
                            p.bq_state_next[2 * c] = s1;
                            p.bq_state_next[2 * c + 1] = s2;
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < R; i++)
                if (i < nrow && cvalid) {
                    const T o = PB_XV(i);
                    __stcs(dst + i * ld, o);
                    if (meter) {
                        const double a = fabs((double)o);
                        m_peak = a > m_peak ? a : m_peak;
                        m_sumsq = fma((double)o, (double)o, m_sumsq);
                    }
                }
        }
#undef PB_XV
#undef PB_XR
        if (meter) {
            __syncthreads();  // zq_s is reused: slower warps may still be reading the sub-chunk sums for their prefix
            zq_s[(warp * kCg + lane) * 2] = m_peak;
            zq_s[(warp * kCg + lane) * 2 + 1] = m_sumsq;
            __syncthreads();
            if (warp == 0 && cvalid) {
                double pk = 0.0, sq = 0.0;
#pragma unroll
                for (int q = 0; q < kStWarps; q++) {
                    const double a = zq_s[(q * kCg + lane) * 2];
                    pk = a > pk ? a : pk;
                    sq += zq_s[(q * kCg + lane) * 2 + 1];
                }
                atomic_max_nonneg(p.meter_peak + c, pk);
                atomicAdd(p.meter_sumsq + c, sq);
            }
        }
    }
}

// First of two sweeps: the aggregate (zero-state end state) of every full tile in front of the last one, and nothing else.  No
// tickets, no shared-memory staging: the R rows of a warp go from HBM straight into registers -- one 128 B line per row and warp,
// all R loads in flight before the first is used --, 2 DFMA per sample give the sub-chunk's end state, and the eight sub-chunk
// states of a tile are combined by one warp (a different one every tile, so that no warp is the slow one).  Tiles are dealt
// round-robin, tile = CTA + k * grid: the resident CTAs read one contiguous window of the batch.  (Measured alternatives, slower:
// one contiguous share of the batch per CTA, chaining the tile aggregates on the way -- 67 us against 57: hundreds of concurrent
// streams --; segments of 2..16 consecutive tiles per CTA -- 67-72 us.)
template <typename T, int CC>
__global__ void __launch_bounds__(kStThreads, 4) stream_aggregate_kernel(const __grid_constant__ StreamParams<T> p)
{
    constexpr int R = StShape<T>::kRows, kTile = StShape<T>::kTile;
    __shared__ double pw_s[4 * kStWarps];
    __shared__ __align__(16) double wt_s[32][2];   // A^k B from shared memory: as kernel parameters the compiler hoists all 64 out of the
                                                   // tile loop and spills them
    __shared__ __align__(16) double zq_s[2][kStWarps * kCg * 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = CC ? CC : p.C;
    const int64_t ld = C;
    if (tid < 4 * kStWarps) pw_s[tid] = p.tab[StTab::kPw + tid];
    if (tid >= 64 && tid < 128) wt_s[(tid - 64) >> 1][tid & 1] = p.wt[(tid - 64) >> 1][tid & 1];
    __syncthreads();
    const int total = (p.n_tiles - 1) * p.n_groups;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, it++) {
        const int t = tile / p.n_groups, g = tile - t * p.n_groups;
        const int c = g * kCg + lane;
        const bool cvalid = c < C;
        const T *src = p.in + ((int64_t)t * kTile + warp * R) * ld + (cvalid ? c : 0);
        T x[R];
#pragma unroll
        for (int i = 0; i < R; i++) x[i] = cvalid ? __ldg(src + i * ld) : T(0);
        double z0 = 0.0, z1 = 0.0, y0 = 0.0, y1 = 0.0, u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int i = 0; i < R; i += 4) {
            // (conversions pinned in program order: hoisted in front of the loop they keep 32 doubles alive and spill)
            const double xa = to_double_again(x[i]), xb = to_double_again(x[i + 1]), xc = to_double_again(x[i + 2]), xd = to_double_again(x[i + 3]);
            const double2 wa = *reinterpret_cast<const double2 *>(wt_s[R - 1 - i]), wb = *reinterpret_cast<const double2 *>(wt_s[R - 2 - i]);
            const double2 wc = *reinterpret_cast<const double2 *>(wt_s[R - 3 - i]), wd = *reinterpret_cast<const double2 *>(wt_s[R - 4 - i]);
            z0 = fma(wa.x, xa, z0);
            z1 = fma(wa.y, xa, z1);
            y0 = fma(wb.x, xb, y0);
            y1 = fma(wb.y, xb, y1);
            u0 = fma(wc.x, xc, u0);
            u1 = fma(wc.y, xc, u1);
            v0 = fma(wd.x, xd, v0);
            v1 = fma(wd.y, xd, v1);
        }
        double *zq = zq_s[it & 1];   // two buffers: the combining warp may still be reading the previous tile's
        *reinterpret_cast<double2 *>(zq + (warp * kCg + lane) * 2) = make_double2((z0 + y0) + (u0 + v0), (z1 + y1) + (u1 + v1));
        __syncthreads();
        if (warp == (it & (kStWarps - 1))) {
            double Z0 = 0.0, Z1 = 0.0;
#pragma unroll
            for (int q = 0; q < kStWarps; q++) {
                const double2 z = *reinterpret_cast<const double2 *>(zq + (q * kCg + lane) * 2);
                mat2_fma(pw_s + 4 * (kStWarps - 1 - q), z.x, z.y, Z0, Z1);
            }
            *reinterpret_cast<double2 *>(p.lb_agg + ((size_t)g * p.n_tiles + t) * 64 + lane * 2) = make_double2(Z0, Z1);
        }
    }
}

// Between the two sweeps: inclusive state after every full tile, S_t = (A^T) S_{t-1} + Z_t.  A dependent DFMA costs ~100 cycles
// on this part, so the recursion is cut three ways until no chain is long: the tiles of a channel group into kScanBlocks blocks
// (one CTA each), a block into 32 segments (one warp each, lane = channel).  First: segment aggregates by Horner, prefix over
// the block's segments from a zero state, block aggregate to global memory.  Then: the block's start state from the carried
// state and the aggregates of the blocks in front, the same prefix from it, and every warp walks its segment again and stores
// the inclusive states.  Payloads (L2) are fetched a batch ahead of the recursion.
constexpr int kScanWarps = 32, kScanBatch = 4, kScanBlocks = 16;
template <bool STORE>
__device__ __forceinline__ void scan_walk(const double *agg_g, double *inc_g, int t0, int t1, double m0, double m1, double m2, double m3,
                                          double &s0, double &s1)
{
    double2 A[kScanBatch], B[kScanBatch];
    auto load = [&](double2 (&v)[kScanBatch], int tb) {
#pragma unroll
        for (int u = 0; u < kScanBatch; u++) {
            const int t = tb + u < t1 ? tb + u : t1 - 1;
            v[u] = __ldcg(reinterpret_cast<const double2 *>(agg_g + (size_t)t * 64));
        }
    };
    auto run = [&](const double2 (&v)[kScanBatch], int tb) {
#pragma unroll
        for (int u = 0; u < kScanBatch; u++)
            if (tb + u < t1) {
                const double n0 = fma(m0, s0, fma(m1, s1, v[u].x)), n1 = fma(m2, s0, fma(m3, s1, v[u].y));
                s0 = n0;
                s1 = n1;
                if (STORE) *reinterpret_cast<double2 *>(inc_g + (size_t)(tb + u) * 64) = make_double2(s0, s1);
            }
    };
    if (t0 >= t1) return;
    load(A, t0);
    for (int tb = t0; tb < t1; tb += 2 * kScanBatch) {
        if (tb + kScanBatch < t1) load(B, tb + kScanBatch);
        run(A, tb);
        if (tb + 2 * kScanBatch < t1) load(A, tb + 2 * kScanBatch);
        run(B, tb + kScanBatch);
    }
}
__device__ __forceinline__ void mat2_pow(const double (&M)[4], int e, double (&P)[4])
{
    double B[4] = {M[0], M[1], M[2], M[3]};
    P[0] = 1.0; P[1] = 0.0; P[2] = 0.0; P[3] = 1.0;
    for (; e > 0; e >>= 1) {
        if (e & 1) {
            const double r0 = P[0] * B[0] + P[1] * B[2], r1 = P[0] * B[1] + P[1] * B[3], r2 = P[2] * B[0] + P[3] * B[2],
                         r3 = P[2] * B[1] + P[3] * B[3];
            P[0] = r0; P[1] = r1; P[2] = r2; P[3] = r3;
        }
        const double q0 = B[0] * B[0] + B[1] * B[2], q1 = B[0] * B[1] + B[1] * B[3], q2 = B[2] * B[0] + B[3] * B[2],
                     q3 = B[2] * B[1] + B[3] * B[3];
        B[0] = q0; B[1] = q1; B[2] = q2; B[3] = q3;
    }
}
// blk: [n_groups][kScanBlocks][32 lanes][2] block aggregates (scratch in global memory); flags: [n_groups][kScanBlocks] words, the
// epoch of the launch that wrote the block aggregate (release / acquire).  ONE launch of n_groups x kScanBlocks CTAs (at most 128
// with the channel groups two sweeps serve: always co-resident, so a CTA may wait for the CTAs in front of it): every CTA
// publishes its block aggregate as soon as it has it and then waits only for the blocks IN FRONT of it, whose aggregates depend
// on nothing.  A warp of this kernel is one dependent chain at ~20 cycles per instruction, so what counts is the number of
// instructions on the longest chain: the 32 segment maps of a block are combined by a Kogge-Stone scan over the warps (five
// rounds) instead of a 32-step walk by one warp, and the matrix powers come from the host.
struct ScanParams {
    const double *agg;
    double *inc, *blk;
    unsigned *flags;
    const double *bq_state, *tab;
    int *err_flag;
    unsigned epoch;
    int C, n_tiles, n_full, span, per;   // tiles per block / per segment (warp)
    double Mper[4];                      // (tile step)^per: over a whole segment
    double Mspan[4];                     // (tile step)^span: over a whole block
};
__global__ void __launch_bounds__(kScanWarps * 32) stream_scan_kernel(const __grid_constant__ ScanParams p)
{
    __shared__ double mat_s[2][kScanWarps][4];
    __shared__ __align__(16) double vec_s[2][kScanWarps][kCg][2];
    __shared__ __align__(16) double blk_start_s[kCg][2];
    const int g = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = g * kCg + lane;
    const double *AT = p.tab + StTab::kPw + 4 * kStWarps;   // the tile step A^T
    const double M[4] = {AT[0], AT[1], AT[2], AT[3]};
    const int b0 = b * p.span < p.n_full ? b * p.span : p.n_full, b1 = b0 + p.span < p.n_full ? b0 + p.span : p.n_full;
    const int t0 = b0 + w * p.per < b1 ? b0 + w * p.per : b1, t1 = t0 + p.per < b1 ? t0 + p.per : b1;
    const double *agg_g = p.agg + (size_t)g * p.n_tiles * 64 + lane * 2;
    double *inc_g = p.inc + (size_t)g * p.n_tiles * 64 + lane * 2;
    double *blk_g = p.blk + ((size_t)g * kScanBlocks) * 64 + lane * 2;
    unsigned *flag_g = p.flags + (size_t)g * kScanBlocks;
    // the segment's affine map from a zero state: v <- (tile step)^n v + (what its n tiles add); n = per except at the block's end
    double v0 = 0.0, v1 = 0.0;
    scan_walk<false>(agg_g, inc_g, t0, t1, M[0], M[1], M[2], M[3], v0, v1);
    double W0 = p.Mper[0], W1 = p.Mper[1], W2 = p.Mper[2], W3 = p.Mper[3];
    const int n = t1 - t0;
    if (n != p.per) {   // the block's last segment (shorter) and the empty ones behind it
        double R[4];
        mat2_pow(M, n, R);
        W0 = R[0]; W1 = R[1]; W2 = R[2]; W3 = R[3];
    }
    int cur = 0;
    if (lane == 0) {
        mat_s[0][w][0] = W0; mat_s[0][w][1] = W1; mat_s[0][w][2] = W2; mat_s[0][w][3] = W3;
    }
    *reinterpret_cast<double2 *>(vec_s[0][w][lane]) = make_double2(v0, v1);
    __syncthreads();
    // inclusive scan of the maps over the warps: (W, v)_w <- (W, v)_w o (W, v)_(w - d)
#pragma unroll 1
    for (int d = 1; d < kScanWarps; d <<= 1) {
        if (w >= d) {
            const double *Wp = mat_s[cur][w - d];
            const double p0 = Wp[0], p1 = Wp[1], p2 = Wp[2], p3 = Wp[3];
            const double2 vp = *reinterpret_cast<const double2 *>(vec_s[cur][w - d][lane]);
            v0 = fma(W0, vp.x, fma(W1, vp.y, v0));
            v1 = fma(W2, vp.x, fma(W3, vp.y, v1));
            const double r0 = W0 * p0 + W1 * p2, r1 = W0 * p1 + W1 * p3, r2 = W2 * p0 + W3 * p2, r3 = W2 * p1 + W3 * p3;
            W0 = r0; W1 = r1; W2 = r2; W3 = r3;
        }
        cur ^= 1;
        if (lane == 0) {
            mat_s[cur][w][0] = W0; mat_s[cur][w][1] = W1; mat_s[cur][w][2] = W2; mat_s[cur][w][3] = W3;
        }
        *reinterpret_cast<double2 *>(vec_s[cur][w][lane]) = make_double2(v0, v1);
        __syncthreads();
    }
    if (w == kScanWarps - 1) {
        if (b + 1 < kScanBlocks) {   // the block from a zero state: published for the blocks behind (nobody is behind the last one)
            *reinterpret_cast<double2 *>(blk_g + (size_t)b * 64) = make_double2(v0, v1);
            __syncwarp();
            if (lane == 0) st_release_u32(flag_g + b, p.epoch);
        }
        // start state of the block: the carried state through the blocks in front (every block in front of this one is full)
        double s0 = c < p.C ? p.bq_state[2 * c] : 0.0, s1 = c < p.C ? p.bq_state[2 * c + 1] : 0.0;
        if (b > 0) {
            for (unsigned spins = 0;; spins++) {
                const unsigned f = lane < b ? ld_acquire_u32(flag_g + lane) : p.epoch;
                if (__all_sync(0xffffffffu, f == p.epoch)) break;
                if (spins > (1u << 22)) {
                    if (lane == 0) atomicExch(p.err_flag, 1);
                    break;
                }
                __nanosleep(32);
            }
            __syncwarp();
#pragma unroll 1
            for (int k0 = 0; k0 < b; k0 += 8) {   // eight payloads per L2 round trip
                double2 z[8];
#pragma unroll
                for (int u = 0; u < 8; u++) z[u] = __ldcg(reinterpret_cast<const double2 *>(blk_g + (size_t)(k0 + u < b ? k0 + u : 0) * 64));
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (k0 + u < b) {
                        const double n0 = fma(p.Mspan[0], s0, fma(p.Mspan[1], s1, z[u].x)), n1 = fma(p.Mspan[2], s0, fma(p.Mspan[3], s1, z[u].y));
                        s0 = n0;
                        s1 = n1;
                    }
            }
        }
        *reinterpret_cast<double2 *>(blk_start_s[lane]) = make_double2(s0, s1);
    }
    __syncthreads();
    // true state at the start of this warp's segment: the map of the segments in front of it applied to the block's start state
    const double2 bs = *reinterpret_cast<const double2 *>(blk_start_s[lane]);
    double s0 = bs.x, s1 = bs.y;
    if (w > 0) {
        const double *Wp = mat_s[cur][w - 1];
        const double2 vp = *reinterpret_cast<const double2 *>(vec_s[cur][w - 1][lane]);
        s0 = fma(Wp[0], bs.x, fma(Wp[1], bs.y, vp.x));
        s1 = fma(Wp[2], bs.x, fma(Wp[3], bs.y, vp.y));
    }
    scan_walk<true>(agg_g, inc_g, t0, t1, M[0], M[1], M[2], M[3], s0, s1);
}

// out = g * in over a flat buffer of n values (g == 1: a copy, bit-exact).  V is the 16-byte vector of T.
template <typename T, typename V>
__global__ void __launch_bounds__(256) stream_map_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t n, T g)
{
    constexpr int kPer = (int)(sizeof(V) / sizeof(T));
    constexpr int kUnroll = 4;
    const int64_t nv = n / kPer;
    const V *vin = reinterpret_cast<const V *>(in);
    V *vout = reinterpret_cast<V *>(out);
    const bool scale = (g != T(1));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * kUnroll;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x * kUnroll + threadIdx.x; i < nv; i += stride) {
        V v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
            if (i + (int64_t)u * blockDim.x < nv) v[u] = __ldcs(vin + i + (int64_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
            if (i + (int64_t)u * blockDim.x < nv) {
                if (scale) {
                    T *e = reinterpret_cast<T *>(&v[u]);
#pragma unroll
                    for (int k = 0; k < kPer; k++) e[k] *= g;
                }
                __stcs(vout + i + (int64_t)u * blockDim.x, v[u]);
            }
    }
    // tail (n not a multiple of the vector width)
    const int64_t i = nv * kPer + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scale ? in[i] * g : in[i];
}

#endif  // __CUDACC__

}  // namespace pb
