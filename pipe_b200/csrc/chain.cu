// chain.cu -- the pb_chain object: planning a run of Processors into fused
// segments, carried state, launches, and the host-marshalling paths.
//
// Reference boundary: a pb_chain is what a ProcessorAllocatorFunc
// (reference line.go:30) would allocate for a contiguous run of GPU Processors,
// and pb_chain_process* is the body of the resulting ProcessFunc
// (pipe.go:64, invoked from Processor.execute at pipe.go:438).
#include <cmath>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <new>
#include <map>
#include <vector>

#include "chain_stream.cuh"
#include "chain_tc.cuh"
#include "chain_tile.cuh"

namespace pb {

struct StageCopy {
    pb_stage_desc d{};
    std::vector<double> taps;
};

struct Segment {
    int fir_stage = -1, bq_stage = -1, rs_stage = -1;
    std::vector<std::pair<int, int>> gain_stages;  // (stage index, gain point 0..3)
    // shape
    int Hf = 0, Hr = 0, P = 0, up = 1, down = 1;
    int L = 0, FB = 0, tp_len = 0, tp_off = 0, wt_len = 0, apow_len = 0;
    size_t smem = 0;
    int grid_max = 0;
    // coefficient tables (device)
    void *d_taps = nullptr, *d_wt = nullptr, *d_apow = nullptr, *d_coef = nullptr;
    // carried state, ping-ponged
    void *d_xhist[2] = {nullptr, nullptr}, *d_yhist[2] = {nullptr, nullptr}, *d_state[2] = {nullptr, nullptr};
    int pp = 0;
    int64_t acc = 0;
    // look-back workspace
    void *d_agg = nullptr, *d_inc = nullptr;
    unsigned *d_status = nullptr;
    int lb_tiles = 0;
    // parameters (host, double)
    double g[4] = {1, 1, 1, 1};
    double b[3] = {1, 0, 0}, a[2] = {0, 0};
    // K2 (tcgen05 path): eligibility, fp16 tables, fixed-point shifts, A^160
    bool tc_ok = false;
    void *d_tc_tables = nullptr;
    void *d_tc_rc = nullptr;
    int tc_sh = 0;
    double tc_AL[4] = {1, 0, 0, 1};
    int tc_sh2 = 0;                                  // fixed-point shift of the folded biquad+resampler matrix P
    double tc_A16[4], tc_ALf[4];                     // A^16, A^145
    float tc_Wz[16][2];
    float tc_Mb[2][kTcBlocks][4];
    double tc_Wb[4], tc_Wbi[4];
    bool tc_level_ok = true;                         // the biquad keeps the broadband level (see build_tc_tables)
    double tc_fir_l1 = 0;                            // sum |h|: bounds the FIR output on the channel's grid (fp16 range of the f pieces)
    // per-channel block exponent of K2 (chain_tc.cuh): two ping-pong copies of {sigma[C], 1/sigma[C]} and of the measured peaks
    float *d_tc_scale = nullptr;                     // [2 copies][sigma, 1 / sigma][2 scale classes][C]
    unsigned *d_tc_peak = nullptr;                   // [2 copies][2 scale classes][C]
    int tc_k = 0;                                    // copy the next call's pass A reads
    std::vector<float> tc_rc;                        // [147 + 32][8] output correction per block state (see chain_tc.cuh)
    // K2 tiles start at the first frame of a call, whatever the resampler's integer phase is there (160 input frames give 147
    // outputs from ANY phase): the resampler-dependent tables (P blocks, rc, the grid shift of P) exist once per phase a chain
    // has met, built on first use -- 4096-frame buffers visit 5 of the 160 phases
    struct TcPhase {
        void *d_tables = nullptr;                    // Toeplitz pieces (a copy) + P blocks of this phase, TcTables::kBytes
        void *d_rc = nullptr;
        int sh2 = 0;
        bool ok = false;                             // the slice schedule covers every output at this phase
    };
    std::map<int, TcPhase> tc_phase;
    // K3 (streaming kernels): runs without FIR and without resampler
    bool st_ok = false;
    int st_grid = 0, st_agg_grid = 0;
    double st_tile_step[4] = {1, 0, 0, 1};          // K3: the biquad state step over one tile (the table entry the kernels call A^T)
    void *d_st_tab = nullptr;                        // StTab
    void *d_scan_blk = nullptr;                      // K3 two sweeps: block aggregates of the scan [groups][kScanBlocks][32][2] doubles
    double st_wt[32][2] = {};                        // A^k B, passed in the kernel parameters
};

struct TmapEntry {  // cached TMA descriptor of a [rows][C] f32 buffer (K2)
    const void *base = nullptr;
    int64_t rows = 0;
    CUtensorMap map;
};

struct Slot {  // one in-flight batch of the pipelined host path
    void *h_in = nullptr, *h_out = nullptr;  // pinned staging
    void *d_in = nullptr, *d_out = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
    std::vector<int64_t> out_counts;
    void *user_out = nullptr;  // pageable destination to fill at collect (nullptr when D2H went direct)
    int64_t out_frames = 0;
    bool busy = false;
};

}  // namespace pb

using namespace pb;

struct pb_chain {
    int device = 0, dtype = PB_F32, C = 0, buffer_frames = 0, max_batch = 1;
    unsigned flags = 0;
    double sample_rate = 0, out_sample_rate = 0;
    size_t elem = 4;
    int64_t max_frames = 0;
    int num_sms = 148;
    std::vector<StageCopy> stages;
    std::vector<Segment> segs;
    void *d_mid[2] = {nullptr, nullptr};
    unsigned long long *d_ticket = nullptr;  // [0] ticket counter, [1] low word = kernel error flag
    unsigned long long ticket_base = 0;
    unsigned epoch = 0;
    double *d_meter = nullptr;  // [2][C]: peak, sumsq
    double *d_meter_scratch = nullptr;  // [2][C]: what pass A of a K2 call metered, until pass B accepts or drops it
    int64_t meter_frames = 0;
    cudaStream_t st_compute = nullptr, st_h2d = nullptr, st_d2h = nullptr;
    Slot slots[2];
    int slot_head = 0, slot_tail = 0, slots_busy = 0;
    int last_path = 0;
    int64_t launches = 0;
    std::vector<TmapEntry> tmaps;
    static constexpr int kMaxPieces = 8;   // pb_chain_process cuts one large host buffer into pieces (process_pieces)
    cudaEvent_t ev_piece[3][kMaxPieces] = {};   // [h2d | done | d2h][piece]
};

namespace pb {

// ------------------------------------------------------------------ helpers --

template <typename T>
static cudaError_t upload(void *dst, const std::vector<double> &v)
{
    std::vector<T> tmp(v.size());
    for (size_t i = 0; i < v.size(); i++) tmp[i] = (T)v[i];
    return cudaMemcpy(dst, tmp.data(), sizeof(T) * tmp.size(), cudaMemcpyHostToDevice);
}

static cudaError_t upload_any(int dtype, void *dst, const std::vector<double> &v)
{
    return dtype == PB_F32 ? upload<float>(dst, v) : upload<double>(dst, v);
}

static int32_t validate_stage(const pb_stage_desc &s, int idx)
{
    switch (s.kind) {
    case PB_STAGE_COPY:
    case PB_STAGE_GAIN:
        if (!std::isfinite(s.gain) && s.kind == PB_STAGE_GAIN) return fail(PB_ERR_INVALID, "stage %d: gain is not finite", idx);
        return PB_OK;
    case PB_STAGE_BIQUAD:
        for (int i = 0; i < 3; i++)
            if (!std::isfinite(s.b[i])) return fail(PB_ERR_INVALID, "stage %d: biquad b[%d] is not finite", idx, i);
        for (int i = 0; i < 2; i++)
            if (!std::isfinite(s.a[i])) return fail(PB_ERR_INVALID, "stage %d: biquad a[%d] is not finite", idx, i);
        return PB_OK;
    case PB_STAGE_FIR:
        if (s.n_taps < 1 || !s.taps) return fail(PB_ERR_INVALID, "stage %d: FIR needs n_taps >= 1 and taps", idx);
        return PB_OK;
    case PB_STAGE_RESAMPLE:
        if (s.up < 1 || s.down < 1 || !s.taps || s.n_taps < s.up || s.n_taps % s.up != 0)
            return fail(PB_ERR_INVALID, "stage %d: resample needs up,down >= 1 and n_taps a multiple of up", idx);
        if (s.up > s.down)
            return fail(PB_ERR_UNSUPPORTED,
                        "stage %d: resample up > down would emit more frames than bufferSize (pipe.go:437-443)", idx);
        return PB_OK;
    default:
        return fail(PB_ERR_INVALID, "stage %d: unknown kind %d", idx, s.kind);
    }
}

// Cut the stage list into fused segments [gain*][FIR]?[gain*][biquad]?[gain*][resample]?[gain*].
static void plan_segments(pb_chain *c)
{
    c->segs.clear();
    Segment cur;
    int rank_used = 0;  // 0 none, 1 FIR, 2 biquad, 3 resample
    for (int i = 0; i < (int)c->stages.size(); i++) {
        const pb_stage_desc &s = c->stages[i].d;
        if (s.kind == PB_STAGE_COPY) continue;
        if (s.kind == PB_STAGE_GAIN) {
            cur.gain_stages.push_back({i, rank_used});
            continue;
        }
        const int rank = s.kind == PB_STAGE_FIR ? 1 : s.kind == PB_STAGE_BIQUAD ? 2 : 3;
        if (rank <= rank_used) {
            c->segs.push_back(cur);
            cur = Segment();
            rank_used = 0;
        }
        if (rank == 1) cur.fir_stage = i;
        if (rank == 2) cur.bq_stage = i;
        if (rank == 3) cur.rs_stage = i;
        rank_used = rank;
    }
    c->segs.push_back(cur);
}

static void free_segment(Segment &s)
{
    void *ptrs[] = {s.d_taps, s.d_wt, s.d_apow, s.d_coef, s.d_xhist[0], s.d_xhist[1], s.d_yhist[0], s.d_yhist[1],
                    s.d_state[0], s.d_state[1], s.d_agg, s.d_inc, (void *)s.d_status, s.d_tc_tables, s.d_tc_rc, s.d_st_tab, s.d_scan_blk,
                    (void *)s.d_tc_scale, (void *)s.d_tc_peak};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto &kv : s.tc_phase) {
        if (kv.second.d_tables) cudaFree(kv.second.d_tables);
        if (kv.second.d_rc) cudaFree(kv.second.d_rc);
    }
    s.tc_phase.clear();
}

template <typename T, int FB>
static int32_t configure_kernel(pb_chain *c, Segment &s)
{
    auto kern = chain_tile_kernel<T, FB>;
    // the attribute is per function, not per chain: always raise it to the architectural maximum
    cudaFuncAttributes fa{};
    PB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int max_dyn = 232448 - (int)fa.sharedSizeBytes;  // 227 KB per CTA, static part included
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    if ((int)s.smem > max_dyn) return fail(PB_ERR_UNSUPPORTED, "tile needs %zu B of shared memory (max %d)", s.smem, max_dyn);
    int per_sm = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, s.smem));
    if (per_sm < 1) return fail(PB_ERR_UNSUPPORTED, "fused tile kernel does not fit on an SM (%zu B shared)", s.smem);
    s.grid_max = per_sm * c->num_sms;
    return PB_OK;
}


// K3: dynamic shared memory attribute of every instantiation (per function, per device) and the resident CTAs per SM
template <typename T>
static cudaError_t configure_stream_kernels(int *per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, chain_stream_kernel<T, 0, kStOneSweep>, kStThreads, 0);
}
template <typename T>
static cudaError_t configure_aggregate_kernel(int *per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, stream_aggregate_kernel<T, 0>, kStThreads, 0);
}

// one launch of chain_stream_kernel in the given mode; the channel counts of BASELINE.json's configs get row addresses with
// immediate offsets
template <typename T, int MODE>
static void launch_stream_mode(const StreamParams<T> &p, int grid, int C, cudaStream_t stream)
{
    if (C == 1024) chain_stream_kernel<T, 1024, MODE><<<grid, kStThreads, 0, stream>>>(p);
    else if (C == 256) chain_stream_kernel<T, 256, MODE><<<grid, kStThreads, 0, stream>>>(p);
    else if (C == 64) chain_stream_kernel<T, 64, MODE><<<grid, kStThreads, 0, stream>>>(p);
    else chain_stream_kernel<T, 0, MODE><<<grid, kStThreads, 0, stream>>>(p);
}

// ---- K2 host side --------------------------------------------------------------

constexpr double kTcMaxPoleRadius = 0.994;   // biquads with poles beyond this stay on K1 (see tc_level_ok)

// fixed-point split: scale 2^shift so that |v| <= 2047 and sum|v| <= 8191 (with |x*2^10| <= 2048 every
// partial sum of the exact accumulator then stays below 2^24)
static int fixed_shift(const double *v, int n)
{
    double mx = 0, sum = 0;
    for (int i = 0; i < n; i++) {
        mx = std::max(mx, std::fabs(v[i]));
        sum += std::fabs(v[i]);
    }
    if (mx == 0) return 0;
    const double lim = std::min(2047.0 / mx, 8191.0 / sum);
    int sh = (int)std::floor(std::log2(lim));
    return std::max(-24, std::min(24, sh));
}

static void split_fixed3(double v, __half &p0, __half &p1, __half &p2)
{
    const double r = std::nearbyint(v);
    p0 = __float2half_rn((float)r);
    const double r1 = v - (double)__half2float(p0);
    p1 = __float2half_rn((float)r1);
    p2 = __float2half_rn((float)(r1 - (double)__half2float(p1)));
}

// ph == nullptr: everything that does not depend on the resampler phase (Toeplitz pieces of the FIR, the biquad's block-state
// matrices, the level check); ph != nullptr: the tables of the tiles that start at integer phase acc0 of the resampler.
static int32_t build_tc_tables(pb_chain *c, Segment &s, int acc0 = 0, Segment::TcPhase *ph = nullptr)
{
    const auto &fir = c->stages[s.fir_stage];
    std::vector<double> h(kTcMaxTaps, 0.0);
    for (size_t k = 0; k < fir.taps.size(); k++) h[k] = fir.taps[k];
    auto gtap = [&](int t) { return (t >= 0 && t < kTcMaxTaps) ? h[(size_t)t] : 0.0; };
    // biquad state response W[k] = A^k B and A^160
    const double A[4] = {-s.a[0], 1.0, -s.a[1], 0.0};
    const double B[2] = {s.b[1] - s.a[0] * s.b[0], s.b[2] - s.a[1] * s.b[0]};
    double M[4] = {1, 0, 0, 1};
    for (int k = 0; k <= kTcFrames; k++) {
        if (k == kTcFrames)
            for (int i = 0; i < 4; i++) s.tc_AL[i] = M[i];
        const double N[4] = {M[0] * A[0] + M[1] * A[2], M[0] * A[1] + M[1] * A[3], M[2] * A[0] + M[3] * A[2],
                             M[2] * A[1] + M[3] * A[3]};
        for (int i = 0; i < 4; i++) M[i] = N[i];
    }
    s.tc_sh = fixed_shift(h.data(), kTcMaxTaps);
    s.tc_fir_l1 = 0;
    for (double v : h) s.tc_fir_l1 += std::fabs(v);
    const double sh = std::ldexp(1.0, s.tc_sh);
    std::vector<__half> tab((size_t)TcTables::kHalfs);
    __half *T0 = tab.data(), *T1 = T0 + TcTables::kT, *T2 = T1 + TcTables::kT, *T3 = T2 + TcTables::kT;
    for (int e = -21; e <= 53; e++)
        for (int n_i = 0; n_i < 8; n_i++)
            for (int k_i = 0; k_i < 8; k_i++) {
                const size_t idx = (size_t)(e + 21) * 64 + n_i * 8 + k_i;
                split_fixed3(gtap(8 * e + n_i - k_i + 1) * sh, T0[idx], T1[idx], T2[idx]);
                T3[idx] = __float2half_rn((float)(gtap(8 * e + n_i - k_i + 1) * sh));
            }
    if (!ph) PB_CUDA(cudaMemcpy(s.d_tc_tables, tab.data(), tab.size() * sizeof(__half), cudaMemcpyHostToDevice));
    // K2's fixed-point grids make its error relative to the level INSIDE the chain (FIR output at full scale), because the
    // biquad is folded into the resampler matrix P.  A biquad that removes most of a broadband signal (a 120 Hz low-pass on
    // white noise: -29 dB) would leave that error standing against a much smaller output, so such chains stay on K1:
    // the L2 norm of the biquad's impulse response (its gain for white input) must be at least -6 dB.
    if (!ph) {
        double s1 = 0, s2 = 0, e2 = 0;
        for (int n = 0; n < 1 << 14; n++) {
            const double x = n == 0 ? 1.0 : 0.0, y = s.b[0] * x + s1;
            s1 = s.b[1] * x - s.a[0] * y + s2;
            s2 = s.b[2] * x - s.a[1] * y;
            e2 += y * y;
        }
        s.tc_level_ok = std::isfinite(e2) && std::sqrt(e2) >= 0.5;
        // K2 hands the biquad state from block to block as floats (6e-8 of the state).  After a large level drop that rounding stands
        // against the quiet signal for as long as the filter remembers the loud one, so filters with a long memory -- poles beyond
        // |z| = 0.994, i.e. more than ~1150 samples for 60 dB: a 40 Hz high-pass at 48 kHz -- stay on K1, whose state is double
        // (tools/k2_soak.py: 2e-6 .. 7e-6 of the channel's peak behind a 60 dB drop with 40 Hz filters, nothing with the others).
        {
            const double a1 = s.a[0], a2 = s.a[1], disc = a1 * a1 - 4.0 * a2;
            const double radius = disc < 0.0 ? std::sqrt(std::fabs(a2)) : 0.5 * (std::fabs(a1) + std::sqrt(disc));
            if (!(radius <= kTcMaxPoleRadius)) s.tc_level_ok = false;
        }
    }
    // MMA2: P = blockdiag(G16) * R.  R[row][m] = coef[branch(m)][i_m - row] is the polyphase matrix of one tile (row = frame + 15;
    // output m is triggered by tile-relative frame i_m = ceil((m+1)*160/147) - 1 with branch 146 - ((i_m+1)*147 % 160)); G16 is the
    // biquad's zero-state response inside a block of 16 rows: y[r'] = sum_{r <= r', same block} g[r'-r] f[r], g[0] = b0,
    // g[k] = (A^(k-1) B)[0].  So P[row][m] = sum_{r' = row .. end of row's block} R[r'][m] g[r'-row].
    // Slice s (outputs [32s, 32s+32)) reads row blocks [2s, 2s+4) (the last slice 3): 19 blocks [32 outputs][16 rows], K-major 8x8
    // core matrices, two fp16 pieces on the fixed grid 2^sh2, the two pieces of a block adjacent ([p0 | p1] is one N = 64 operand).
    // Entry 19 is block (0, 0) for tile 0: rows 0..14 are the carried y history (plain R, no biquad), row 15 starts the recursion.
    const auto &rs = c->stages[s.rs_stage];
    auto rtap = [&](int row, int m) -> double {
        if (m >= kTcOut || row < 0 || row >= kTcN) return 0.0;
        // the integer accumulator starts the tile at acc0: after tile frame i it has seen acc0 + 147 (i + 1), output m leaves at
        // the first i where that reaches 160 (m + 1), with what is left over as the accumulator (resampler)
        const int im = (kTcFrames * (m + 1) - acc0 + kTcUp - 1) / kTcUp - 1;
        const int k = im + kTcHr - row;
        if (k < 0 || k >= kTcP) return 0.0;
        const int br = kTcUp - 1 - (acc0 + (im + 1) * kTcUp - kTcFrames * (m + 1));
        return rs.taps[(size_t)br + (size_t)k * kTcUp];
    };
    auto trigger = [&](int m) { return (kTcFrames * (m + 1) - acc0 + kTcUp - 1) / kTcUp - 1; };
    double g[16], Apow[kTcN + 1][4];
    {
        double Mk[4] = {1, 0, 0, 1};
        for (int k = 0; k <= kTcN; k++) {
            for (int i = 0; i < 4; i++) Apow[k][i] = Mk[i];
            const double N[4] = {Mk[0] * A[0] + Mk[1] * A[2], Mk[0] * A[1] + Mk[1] * A[3], Mk[2] * A[0] + Mk[3] * A[2],
                                 Mk[2] * A[1] + Mk[3] * A[3]};
            for (int i = 0; i < 4; i++) Mk[i] = N[i];
        }
        g[0] = s.b[0];
        for (int k = 1; k < 16; k++) g[k] = Apow[k - 1][0] * B[0] + Apow[k - 1][1] * B[1];
    }
    // first0: the tile-0 variant of block 0
    auto ptap = [&](int row, int m, bool first0) -> double {
        if (first0 && row < kTcHr) return rtap(row, m);
        const int end = (row / 16) * 16 + 15;
        double v = 0;
        for (int rp = row; rp <= end && rp < kTcN; rp++) v += rtap(rp, m) * g[rp - row];
        return v;
    };
    {
        double mx = 0, smax = 0;
        for (int f0 = 0; f0 < 2; f0++)
            for (int m = 0; m < kTcOut; m++) {
                double sum = 0;  // the grid must keep sum|p0| of ONE output below 8191
                for (int row = 0; row < kTcN; row++) {
                    const double v = std::fabs(ptap(row, m, f0 && row < 16));
                    sum += v;
                    mx = std::max(mx, v);
                }
                smax = std::max(smax, sum);
            }
        int sh2 = 0;
        if (mx > 0) sh2 = (int)std::floor(std::log2(std::min(2047.0 / mx, 8191.0 / smax)));
        sh2 = std::max(-24, std::min(24, sh2));
        if (ph) ph->sh2 = sh2;
        else s.tc_sh2 = sh2;
    }
    const double sh2 = std::ldexp(1.0, ph ? ph->sh2 : s.tc_sh2);
    std::vector<__half> b2((size_t)TcTables::kB2);
    auto fill_pair = [&](int pair, int sl, int chunk, bool first0) {
        for (int n = 0; n < kRsN; n++)
            for (int kk = 0; kk < 16; kk++) {
                const double v = ptap(16 * chunk + kk, kRsN * sl + n, first0) * sh2;
                const size_t idx = (size_t)(n / 8) * 128 + (size_t)(kk / 8) * 64 + (size_t)(n % 8) * 8 + (size_t)(kk % 8);
                __half r0 = __float2half_rn((float)std::nearbyint(v));
                __half r1 = __float2half_rn((float)(v - (double)__half2float(r0)));
                b2[(size_t)pair * 1024 + idx] = r0;
                b2[(size_t)pair * 1024 + 512 + idx] = r1;
            }
    };
    int pair = 0;
    for (int sl = 0; sl < kRsSlices; sl++) {
        const int nch = (sl == kRsSlices - 1) ? 3 : 4;
        for (int k = 0; k < nch; k++, pair++) fill_pair(pair, sl, 2 * sl + k, false);
    }
    fill_pair(kRsPairs, 0, 0, true);
    // every tap of every output must be inside the blocks its slice reads
    for (int m = 0; m < kTcOut; m++) {
        const int im = trigger(m), sl = m / kRsN;
        const int lo = im, hi = im + kTcHr, nch = (sl == kRsSlices - 1) ? 3 : 4;
        if (im < 0 || im >= kTcFrames || lo < 32 * sl || hi >= 32 * sl + 16 * nch)
            return fail(PB_ERR_UNSUPPORTED, "resampler slice schedule does not cover output %d at phase %d", m, acc0);
    }
    if (ph) {
        PB_CUDA(cudaMalloc(&ph->d_tables, (size_t)TcTables::kBytes));
        PB_CUDA(cudaMalloc(&ph->d_rc, (size_t)kTcRcRows * 8 * sizeof(float)));
        PB_CUDA(cudaMemcpy(ph->d_tables, s.d_tc_tables, (size_t)TcTables::kHalfs * 2, cudaMemcpyDeviceToDevice));
        PB_CUDA(cudaMemcpy((char *)ph->d_tables + (size_t)TcTables::kHalfs * 2, b2.data(), b2.size() * sizeof(__half),
                           cudaMemcpyHostToDevice));
    }
    // Block states.  The free response of a block to the state s at its start is y_sr[i] = (A^i s)[0], i = 0..15: a 16 x 2
    // matrix H.  With H = U S V^T the states travel as w = W s, W = S V^T ("balanced" coordinates): y_sr = U w with
    // orthonormal U, so neither the tables below nor the float arithmetic on w see the cancellation that the TDF-II basis
    // has for poles near the unit circle.  The response of output m to w at the start of block b is
    //   rc[m][2k..2k+1] = g_bq g_out sum_{rows i of block b} R[16b+i][m] U[i][.],   b = 2 (m/32) + k, k = 0..3;
    // tile 0, block 0: the state enters at row 15 (only U[0][.] applies, to row 15).
    {
        double H[16][2], G[3] = {0, 0, 0};  // G = H^T H
        for (int i = 0; i < 16; i++) {
            H[i][0] = Apow[i][0];
            H[i][1] = Apow[i][1];
            G[0] += H[i][0] * H[i][0];
            G[1] += H[i][0] * H[i][1];
            G[2] += H[i][1] * H[i][1];
        }
        // eigen-decomposition of the symmetric 2x2 G: V (columns), S^2
        const double th = 0.5 * std::atan2(2.0 * G[1], G[0] - G[2]);
        const double cs = std::cos(th), sn = std::sin(th);
        const double V[2][2] = {{cs, -sn}, {sn, cs}};
        double S[2];
        for (int q = 0; q < 2; q++) {
            double n2 = 0;
            for (int i = 0; i < 16; i++) {
                const double v = H[i][0] * V[0][q] + H[i][1] * V[1][q];
                n2 += v * v;
            }
            S[q] = std::sqrt(n2);
            if (!(S[q] > 1e-300)) S[q] = 1e-300;
        }
        double U[16][2], W[4], Wi[4];
        for (int i = 0; i < 16; i++)
            for (int q = 0; q < 2; q++) U[i][q] = (H[i][0] * V[0][q] + H[i][1] * V[1][q]) / S[q];
        for (int q = 0; q < 2; q++) {
            W[2 * q] = S[q] * V[0][q];       // w_q = S_q (V^T s)_q
            W[2 * q + 1] = S[q] * V[1][q];
            Wi[q] = V[0][q] / S[q];          // s = V S^-1 w
            Wi[2 + q] = V[1][q] / S[q];
        }
        if (!ph)
            for (int i = 0; i < 4; i++) {
                s.tc_Wb[i] = W[i];
                s.tc_Wbi[i] = Wi[i];
            }
        const double gg = s.g[2] * s.g[3];
        s.tc_rc.assign((size_t)kTcRcRows * 8, 0.f);
        for (int m = 0; m < kTcOut; m++)
            for (int b = 0; b < kTcBlocks; b++) {
                double v0 = 0, v1 = 0;
                for (int i = 0; i < 16; i++) {
                    const double r = rtap(16 * b + i, m);
                    v0 += r * U[i][0];
                    v1 += r * U[i][1];
                }
                const int k = b - 2 * (m / kRsN);
                if (k >= 0 && k < 4) {
                    s.tc_rc[(size_t)m * 8 + 2 * k] = (float)(v0 * gg);
                    s.tc_rc[(size_t)m * 8 + 2 * k + 1] = (float)(v1 * gg);
                } else if (v0 != 0.0 || v1 != 0.0) {
                    return fail(PB_ERR_UNSUPPORTED, "resampler output %d reads rows outside the blocks of its slice", m);
                }
            }
        for (int m = 0; m < kRsN; m++) {
            for (int q = 0; q < 8; q++) s.tc_rc[(size_t)(kTcRcFirst + m) * 8 + q] = s.tc_rc[(size_t)m * 8 + q];
            s.tc_rc[(size_t)(kTcRcFirst + m) * 8 + 0] = (float)(rtap(kTcHr, m) * U[0][0] * gg);  // block 0: only row 15, y_sr = U[0] w
            s.tc_rc[(size_t)(kTcRcFirst + m) * 8 + 1] = (float)(rtap(kTcHr, m) * U[0][1] * gg);
        }
        if (ph) {
            PB_CUDA(cudaMemcpy(ph->d_rc, s.tc_rc.data(), s.tc_rc.size() * sizeof(float), cudaMemcpyHostToDevice));
            ph->ok = true;
            return PB_OK;
        }
        auto sim = [&](const double *M, double *out) {  // W M W^-1
            double t[4] = {W[0] * M[0] + W[1] * M[2], W[0] * M[1] + W[1] * M[3], W[2] * M[0] + W[3] * M[2], W[2] * M[1] + W[3] * M[3]};
            out[0] = t[0] * Wi[0] + t[1] * Wi[2];
            out[1] = t[0] * Wi[1] + t[1] * Wi[3];
            out[2] = t[2] * Wi[0] + t[3] * Wi[2];
            out[3] = t[2] * Wi[1] + t[3] * Wi[3];
        };
        for (int b = 0; b < kTcBlocks; b++) {
            double m0[4], m1[4];
            sim(Apow[16 * b], m0);
            if (b == 0) { m1[0] = 1; m1[1] = 0; m1[2] = 0; m1[3] = 1; }
            else sim(Apow[16 * b - kTcHr], m1);
            for (int i = 0; i < 4; i++) {
                s.tc_Mb[0][b][i] = (float)m0[i];
                s.tc_Mb[1][b][i] = (float)m1[i];
            }
        }
        sim(Apow[16], s.tc_A16);  // block step in balanced coordinates
        for (int i = 0; i < 4; i++) s.tc_ALf[i] = Apow[kTcFrames - kTcHr][i];
        for (int i = 0; i < 16; i++) {  // W A^(15-i) B
            const double z0 = Apow[15 - i][0] * B[0] + Apow[15 - i][1] * B[1], z1 = Apow[15 - i][2] * B[0] + Apow[15 - i][3] * B[1];
            s.tc_Wz[i][0] = (float)(W[0] * z0 + W[1] * z1);
            s.tc_Wz[i][1] = (float)(W[2] * z0 + W[3] * z1);
        }
    }
    return PB_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D f32 tensor map over a [rows][C] frame-major buffer, box 128 channels x 16 frames (one K chunk of a K2 tile), no swizzle:
// the converter threads read it one channel per lane, 128 B per warp and row
static int32_t make_frame_map(CUtensorMap *map, const void *base, int C, int64_t rows)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(PB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTcCh, 16};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PB_OK;
}

// Encoding a tensor map costs a driver call; a chain sees the same few (pointer, rows) pairs call after call (its two history
// buffers, the caller's or the slots' input buffers), so the last few are kept.
static int32_t cached_frame_map(pb_chain *c, CUtensorMap *out, const void *base, int64_t rows)
{
    for (auto &e : c->tmaps)
        if (e.base == base && e.rows == rows) {
            *out = e.map;
            return PB_OK;
        }
    TmapEntry e;
    e.base = base;
    e.rows = rows;
    int32_t r = make_frame_map(&e.map, base, c->C, rows);
    if (r != PB_OK) return r;
    if (c->tmaps.size() >= 16) c->tmaps.erase(c->tmaps.begin());
    c->tmaps.push_back(e);
    *out = e.map;
    return PB_OK;
}

static bool tc_shape_ok(const pb_chain *c, const Segment &s)
{
    if (c->dtype != PB_F32 || (c->flags & PB_CHAIN_NO_TENSOR) || c->C % kTcCh != 0) return false;
    if (c->C / kTcCh > c->num_sms) return false;  // a CTA never owns two tiles of the last time step (chain_tc.cuh, carried state)
    if (s.fir_stage < 0 || s.bq_stage < 0 || s.rs_stage < 0) return false;
    if (s.Hf + 1 > kTcMaxTaps || s.Hf < 1) return false;
    return s.up == kTcUp && s.down == kTcFrames && s.P == kTcP;
}

// per-launch switches of the K2 function attributes (per device, set once)
static int32_t tc_configure_once(int device)
{
    static std::once_flag once[64];
    static cudaError_t err[64];
    if (device < 0 || device >= 64) return fail(PB_ERR_INVALID, "device ordinal %d", device);
    std::call_once(once[device], [&] {
        err[device] = cudaFuncSetAttribute(chain_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
        if (err[device] == cudaSuccess)
            err[device] = cudaFuncSetAttribute(chain_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
        if (err[device] == cudaSuccess)
            err[device] = cudaFuncSetAttribute(chain_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
        if (err[device] == cudaSuccess)
            err[device] = cudaFuncSetAttribute(chain_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
        if (err[device] == cudaSuccess)
            err[device] = cudaFuncSetAttribute(chain_tc_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    });
    PB_CUDA(err[device]);
    return PB_OK;
}

static int32_t tc_reset_scales(pb_chain *c, Segment &s, cudaStream_t stream)
{
    // before the first call nothing is known about the levels: assume full scale (|g x| = 1); pass B corrects it
    // layout: [2 ping-pong copies][sigma | 1 / sigma][2 scale classes][C]
    const size_t C2 = 2 * (size_t)c->C;
    std::vector<float> init(4 * C2);
    for (int k = 0; k < 2; k++)
        for (size_t i = 0; i < C2; i++) {
            init[(size_t)(2 * k) * C2 + i] = kSigTarget;
            init[(size_t)(2 * k + 1) * C2 + i] = 1.0f / kSigTarget;
        }
    PB_CUDA(cudaMemcpyAsync(s.d_tc_scale, init.data(), init.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemsetAsync(s.d_tc_peak, 0, sizeof(unsigned) * 2 * C2, stream));
    PB_CUDA(cudaStreamSynchronize(stream));  // `init` is pageable host memory
    s.tc_k = 0;
    return PB_OK;
}

// The tables of the tiles that start at resampler phase acc0 (built on first use; *out == nullptr with PB_OK when the slice schedule
// does not cover that phase -- the caller then falls back to tiles that start at phase 0).
static int32_t tc_get_phase(pb_chain *c, Segment &s, int acc0, const Segment::TcPhase **out)
{
    auto it = s.tc_phase.find(acc0);
    if (it == s.tc_phase.end()) {
        Segment::TcPhase ph;
        const int32_t r = build_tc_tables(c, s, acc0, &ph);
        if (r != PB_OK && r != PB_ERR_UNSUPPORTED) return r;
        if (r != PB_OK) {
            if (ph.d_tables) cudaFree(ph.d_tables);
            if (ph.d_rc) cudaFree(ph.d_rc);
            ph = Segment::TcPhase();
        }
        it = s.tc_phase.emplace(acc0, ph).first;
    }
    *out = it->second.ok ? &it->second : nullptr;
    return PB_OK;
}

// One K2 call: `n` (>= 160) input frames of segment `s`, starting at the resampler phase s.acc, as two launches of the same kernel: pass A with the scales the
// previous call ended with, pass B which verifies them and returns at once unless a channel left its grid window (chain_tc.cuh).
static int32_t launch_segment_tc(pb_chain *c, Segment &s, const Segment::TcPhase &ph, const void *in, int64_t n, void *out,
                                 bool is_last_segment, cudaStream_t stream)
{
    int32_t r = tc_configure_once(c->device);
    if (r != PB_OK) return r;
    TcParams p{};
    r = cached_frame_map(c, &p.tm_in, in, n);
    if (r != PB_OK) return r;
    r = cached_frame_map(c, &p.tm_hist, s.d_xhist[s.pp], s.Hf);
    if (r != PB_OK) return r;
    p.out = (float *)out;
    p.tables = (const __half *)ph.d_tables;
    p.yhist = (const float *)s.d_yhist[s.pp];
    p.yhist_next = (float *)s.d_yhist[s.pp ^ 1];
    p.xhist_next = (float *)s.d_xhist[s.pp ^ 1];
    p.bq_state = (const double *)s.d_state[s.pp];
    p.bq_state_next = (double *)s.d_state[s.pp ^ 1];
    p.lb_agg = (double *)s.d_agg;
    p.lb_inc = (double *)s.d_inc;
    p.lb_status = s.d_status;
    const bool meter = is_last_segment && (c->flags & PB_CHAIN_METER);
    p.meter_main = meter ? c->d_meter : nullptr;
    p.meter_scratch = meter ? c->d_meter_scratch : nullptr;
    p.scale = s.d_tc_scale + (size_t)s.tc_k * 4 * c->C;
    p.scale_next = s.d_tc_scale + (size_t)(s.tc_k ^ 1) * 4 * c->C;
    p.peak = s.d_tc_peak + (size_t)s.tc_k * 2 * c->C;
    p.peak_next = s.d_tc_peak + (size_t)(s.tc_k ^ 1) * 2 * c->C;
    p.err_flag = reinterpret_cast<int *>(c->d_ticket + 1);
    p.C = c->C;
    // the last tile may be partial: TMA zero-fills the rows behind the call's last frame, the kernel stores only the outputs the
    // real frames trigger and takes the carried state at the last real row
    p.n_tiles = (int)((n + kTcFrames - 1) / kTcFrames);
    p.last_frames = (int)(n - (int64_t)(p.n_tiles - 1) * kTcFrames);
    p.last_outputs = (int)((s.acc + (int64_t)p.last_frames * kTcUp) / kTcFrames);
    p.n_cg = c->C / kTcCh;
    p.hist_rows = s.Hf;
    p.g_load = (float)s.g[0];
    p.fscale = (float)(s.g[1] / std::ldexp(1.0, s.tc_sh));
    p.inv_gbq = (float)(1.0 / s.g[2]);
    p.descale_rs = (float)(s.g[2] * s.g[3] / std::ldexp(1.0, ph.sh2));
    p.b0 = s.b[0]; p.b1 = s.b[1]; p.b2 = s.b[2]; p.a1 = s.a[0]; p.a2 = s.a[1];
    p.g_bq = s.g[2];
    for (int i = 0; i < 4; i++) {
        p.AL[i] = s.tc_AL[i];
        p.AL_first[i] = s.tc_ALf[i];
        p.A16[i] = s.tc_A16[i];
    }
    {   // two block steps (balanced coordinates, like A16)
        const double *a = s.tc_A16;
        p.A32[0] = a[0] * a[0] + a[1] * a[2];
        p.A32[1] = a[0] * a[1] + a[1] * a[3];
        p.A32[2] = a[2] * a[0] + a[3] * a[2];
        p.A32[3] = a[2] * a[1] + a[3] * a[3];
    }
    for (int i = 0; i < 16; i++) {  // applied to the raw accumulator sum: fold fscale in
        p.Wz[i][0] = (float)((double)s.tc_Wz[i][0] * (double)p.fscale);
        p.Wz[i][1] = (float)((double)s.tc_Wz[i][1] * (double)p.fscale);
    }
    memcpy(p.Mb, s.tc_Mb, sizeof(p.Mb));
    for (int i = 0; i < 4; i++) {
        p.Wb[i] = s.tc_Wb[i];
        p.Wbi[i] = s.tc_Wbi[i];
    }
    p.rc = (const float *)ph.d_rc;
    if (p.n_tiles > s.lb_tiles) return fail(PB_ERR_CAPACITY, "batch of %lld frames exceeds the chain's max_batch", (long long)n);
    const int total = p.n_tiles * p.n_cg;
    const int grid = std::min(total, c->num_sms);
    // development aid: PB_TC_PROF=1 prints per-role cycle counters (averaged over CTAs) for every K2 launch
    static const int prof_mode = getenv("PB_TC_PROF") ? atoi(getenv("PB_TC_PROF")) : 0;   // 1: counters, 2: timeline of CTA 0
    const bool prof_on = prof_mode != 0;
    static const int dbg = getenv("PB_TC_DBG") ? atoi(getenv("PB_TC_DBG")) : 0;
    p.dbg = dbg;
    long long *d_prof = nullptr;
    if (prof_on) {
        PB_CUDA(cudaMalloc((void **)&d_prof, sizeof(long long) * tc::kProfCount * grid));
        PB_CUDA(cudaMemset(d_prof, 0, sizeof(long long) * tc::kProfCount * grid));
        p.prof = d_prof;
        PB_CUDA(cudaMalloc((void **)&p.trace, sizeof(long long) * tc::kTraceTiles * tc::kTraceRoles * 32));
        PB_CUDA(cudaMemset(p.trace, 0, sizeof(long long) * tc::kTraceTiles * tc::kTraceRoles * 32));
    }
    for (int pass = 0; pass < 2; pass++) {
        p.pass = pass;
        // the two passes publish their look-back states under different epochs; pass A meters into the scratch copy
        c->epoch = (c->epoch % 0x3ffffffeu) + 1u;
        p.epoch = c->epoch;
        double *mtr = !meter ? nullptr : pass == 0 ? c->d_meter_scratch : c->d_meter;
        p.meter_peak = mtr;
        p.meter_sumsq = mtr ? mtr + c->C : nullptr;
        // Cooperative launch: the static tile schedule makes a CTA's look-back wait on tiles of OTHER CTAs of the grid, so the
        // whole grid (<= one CTA per SM) has to be resident at once.  A cooperative grid is only started when all of it fits
        // next to whatever else holds the device's SMs (another chain's kernel on its own stream, another process under MPS):
        // a late CTA can then no longer leave its successors spinning into the look-back bound.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = tc::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute coop[1];
        coop[0].id = cudaLaunchAttributeCooperative;
        coop[0].val.cooperative = 1;
        static const bool no_coop = getenv("PB_TC_NO_COOP") != nullptr;   // development: plain launches (timing comparison)
        cfg.attrs = coop;
        cfg.numAttrs = no_coop ? 0 : 1;
        if (prof_mode == 1) PB_CUDA(cudaLaunchKernelEx(&cfg, chain_tc_kernel<1>, p));
        else if (prof_mode == 2) PB_CUDA(cudaLaunchKernelEx(&cfg, chain_tc_kernel<2>, p));
        else if (p.dbg != 0) PB_CUDA(cudaLaunchKernelEx(&cfg, (chain_tc_kernel<0, false, true>), p));   // development switches (whole tiles only)
        else if (p.last_frames < kTcFrames) PB_CUDA(cudaLaunchKernelEx(&cfg, (chain_tc_kernel<0, true>), p));
        else PB_CUDA(cudaLaunchKernelEx(&cfg, chain_tc_kernel<0>, p));
        c->launches++;
        if (prof_on && pass == 0) {
            PB_CUDA(cudaStreamSynchronize(stream));
            std::vector<long long> h((size_t)tc::kProfCount * grid);
            PB_CUDA(cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            static const char *names[] = {"prod_wait_empty", "mma_wait_tmem", "mma_wait_cvt", "mma_issue", "cvt_wait_raw",
                                          "cvt_wait_a1", "cvt_work", "drain_wait_blk", "drain_work", "drain_lookback", "total",
                                          "out_wait", "out_main", "mma2_wait_a2", "out_tmem_ld", "out_math", "ns (MMA1 warp)", "-"};
            const double tiles_per_cta = (double)total / grid;
            fprintf(stderr, "[PB_TC_PROF] grid %d, %.1f tiles/CTA; cycles per tile (mean over CTAs):\n", grid, tiles_per_cta);
            for (int k = 0; k < tc::kProfCount; k++) {
                double sum = 0;
                for (int b = 0; b < grid; b++) sum += (double)h[(size_t)b * tc::kProfCount + k];
                fprintf(stderr, "  %-16s %10.0f\n", names[k], sum / grid / tiles_per_cta);
            }
            PB_CUDA(cudaMemset(d_prof, 0, sizeof(long long) * tc::kProfCount * grid));
            // event timeline of CTA 0 (clock64, relative to the first event)
            std::vector<long long> tr((size_t)tc::kTraceTiles * tc::kTraceRoles * 32);
            PB_CUDA(cudaMemcpy(tr.data(), p.trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            long long t0 = 0;
            const long long kRawTag = 1ll << 62;   // entries above this are raw values, not clocks
            for (long long v : tr)
                if (v && v < kRawTag && (!t0 || v < t0)) t0 = v;
            static const char *roles[] = {"mma1", "cvt0", "cvt1", "drainA", "drainB", "mma2", "out"};
            for (int ti = 0; ti < tc::kTraceTiles; ti++)
                for (int r = 0; r < tc::kTraceRoles; r++) {
                    fprintf(stderr, "[PB_TC_TRACE] tile %d %-6s", ti, roles[r]);
                    for (int k = 0; k < 32; k++) {
                        const long long v = tr[((size_t)ti * tc::kTraceRoles + r) * 32 + k];
                        if (v) fprintf(stderr, " %d:%lld", k, v >= kRawTag ? v - kRawTag : v - t0);
                    }
                    fprintf(stderr, "\n");
                }
        }
    }
    if (prof_on) {
        PB_CUDA(cudaStreamSynchronize(stream));
        cudaFree(d_prof);
        cudaFree(p.trace);
    }
    s.tc_k ^= 1;
    s.pp ^= 1;
    return PB_OK;
}

// (Re)compute the parameter-dependent tables of one segment and upload them.
static int32_t refresh_segment_params(pb_chain *c, Segment &s)
{
    for (int i = 0; i < 4; i++) s.g[i] = 1.0;
    for (auto &gs : s.gain_stages) s.g[gs.second] *= c->stages[gs.first].d.gain;
    if (s.fir_stage >= 0) {
        const auto &st = c->stages[s.fir_stage];
        std::vector<double> tp((size_t)s.tp_len, 0.0);
        for (int k = 0; k < (int)st.taps.size(); k++) tp[(size_t)k + s.tp_off] = st.taps[k];
        PB_CUDA(upload_any(c->dtype, s.d_taps, tp));
    }
    if (s.bq_stage >= 0) {
        const pb_stage_desc &d = c->stages[s.bq_stage].d;
        for (int i = 0; i < 3; i++) s.b[i] = d.b[i];
        for (int i = 0; i < 2; i++) s.a[i] = d.a[i];
        // TDF-II as a state-space system: s' = A s + B x, y = s[0] + b0 x
        const double A[4] = {-s.a[0], 1.0, -s.a[1], 0.0};
        const double B[2] = {s.b[1] - s.a[0] * s.b[0], s.b[2] - s.a[1] * s.b[0]};
        std::vector<double> apow((size_t)s.apow_len * 4), wt((size_t)s.wt_len * 2);
        double M[4] = {1, 0, 0, 1};
        for (int k = 0; k < s.apow_len; k++) {
            for (int i = 0; i < 4; i++) apow[(size_t)k * 4 + i] = M[i];
            if (k < s.wt_len) {
                wt[(size_t)k * 2] = M[0] * B[0] + M[1] * B[1];
                wt[(size_t)k * 2 + 1] = M[2] * B[0] + M[3] * B[1];
            }
            const double N[4] = {M[0] * A[0] + M[1] * A[2], M[0] * A[1] + M[1] * A[3],
                                 M[2] * A[0] + M[3] * A[2], M[2] * A[1] + M[3] * A[3]};
            for (int i = 0; i < 4; i++) M[i] = N[i];
        }
        PB_CUDA(upload<double>(s.d_apow, apow));
        PB_CUDA(upload<double>(s.d_wt, wt));
        if (s.st_ok) {
            // K3 tables (chain_stream.cuh), R rows per warp: A^k B for k < R (kernel parameters), A^(R j) for j <= 16,
            // (A^T)^i and (A^T)^(32 i) for i <= 32, with the tile step T = 8 R
            const int R = c->dtype == PB_F32 ? StShape<float>::kRows : StShape<double>::kRows;
            std::vector<double> tab((size_t)StTab::kCount);
            auto mul = [](const double *X, const double *Y, double *Z) {
                const double r[4] = {X[0] * Y[0] + X[1] * Y[2], X[0] * Y[1] + X[1] * Y[3], X[2] * Y[0] + X[3] * Y[2], X[2] * Y[1] + X[3] * Y[3]};
                for (int i = 0; i < 4; i++) Z[i] = r[i];
            };
            double P[4] = {1, 0, 0, 1};
            for (int k = 0; k <= kStWarps * R; k++) {
                if (k < R) {
                    s.st_wt[k][0] = P[0] * B[0] + P[1] * B[1];
                    s.st_wt[k][1] = P[2] * B[0] + P[3] * B[1];
                }
                if (k % R == 0)
                    for (int i = 0; i < 4; i++) tab[StTab::kPw + 4 * (k / R) + i] = P[i];
                mul(P, A, P);
            }
            const double *AT = &tab[StTab::kPw + 4 * kStWarps];  // A^T
            double Q[4] = {1, 0, 0, 1};
            for (int i = 0; i <= kStWin; i++) {
                for (int e = 0; e < 4; e++) tab[StTab::kLb + 4 * i + e] = Q[e];
                mul(Q, AT, Q);
            }
            const double *A32 = &tab[StTab::kLb + 4 * kStWin];   // (A^T)^32
            double W[4] = {1, 0, 0, 1};
            for (int w = 0; w <= kStWin; w++) {
                for (int e = 0; e < 4; e++) tab[StTab::kMw + 4 * w + e] = W[e];
                mul(W, A32, W);
            }
            for (int e = 0; e < 4; e++) s.st_tile_step[e] = AT[e];
            PB_CUDA(upload<double>(s.d_st_tab, tab));
        }
    }
    if (s.rs_stage >= 0) {
        const auto &st = c->stages[s.rs_stage];
        std::vector<double> coef((size_t)s.up * s.P);
        for (int br = 0; br < s.up; br++)
            for (int k = 0; k < s.P; k++) coef[(size_t)br * s.P + k] = st.taps[(size_t)br + (size_t)k * s.up];
        PB_CUDA(upload_any(c->dtype, s.d_coef, coef));
    }
    if (s.tc_ok) {
        // the per-phase tables hold the biquad, the resampler taps and the gains behind them: rebuilt on next use
        for (auto &kv : s.tc_phase) {
            if (kv.second.d_tables) cudaFree(kv.second.d_tables);
            if (kv.second.d_rc) cudaFree(kv.second.d_rc);
        }
        s.tc_phase.clear();
        return build_tc_tables(c, s);
    }
    return PB_OK;
}

static int32_t build_segment(pb_chain *c, Segment &s)
{
    const bool f32 = c->dtype == PB_F32;
    const size_t el = c->elem;
    s.FB = f32 ? 16 : 8;
    if (s.fir_stage >= 0) s.Hf = c->stages[s.fir_stage].d.n_taps - 1;
    if (s.rs_stage >= 0) {
        const pb_stage_desc &d = c->stages[s.rs_stage].d;
        s.up = d.up;
        s.down = d.down;
        s.P = d.n_taps / d.up;
        s.Hr = s.P - 1;
    }
    const bool has_fir = s.fir_stage >= 0;
    s.st_ok = !has_fir && s.rs_stage < 0 && !(c->flags & PB_CHAIN_NO_STREAM);
    // tile length
    auto smem_for = [&](int L) {
        const int tp_len = has_fir ? (s.Hf + 1) + 2 * s.FB + s.FB : 0;
        const int wt_len = (L + s.Hr) / kNW + 2;
        return f32 ? TileSmem<float>(L, s.Hf, s.Hr, has_fir, tp_len, wt_len, s.FB).total
                   : TileSmem<double>(L, s.Hf, s.Hr, has_fir, tp_len, wt_len, s.FB).total;
    };
    if (has_fir) {
        int rows = kNW * s.FB * 2;
        while (rows <= s.Hr * 2) rows += kNW * s.FB;
        if (smem_for(rows - s.Hr) > 110 * 1024 && rows - kNW * s.FB > s.Hr * 2) rows -= kNW * s.FB;
        s.L = rows - s.Hr;
    } else {
        s.L = f32 ? 256 : 128;
        while (s.L <= s.Hr * 2) s.L *= 2;
    }
    s.tp_off = 2 * s.FB;
    s.tp_len = has_fir ? (s.Hf + 1) + s.tp_off + s.FB : 0;
    s.wt_len = (s.L + s.Hr) / kNW + 2;
    s.apow_len = s.L + s.Hr + 1;
    s.smem = smem_for(s.L);
    if (s.smem > 226 * 1024)
        return fail(PB_ERR_UNSUPPORTED, "FIR with %d taps needs %zu B of shared memory per tile (max 232448)", s.Hf + 1, s.smem);
    if (f32) {
        int32_t r = configure_kernel<float, 16>(c, s);
        if (r != PB_OK) return r;
    } else {
        int32_t r = configure_kernel<double, 8>(c, s);
        if (r != PB_OK) return r;
    }
    if (s.st_ok) {
        int per_sm = 0;
        if (f32) PB_CUDA((configure_stream_kernels<float>(&per_sm)));
        else PB_CUDA((configure_stream_kernels<double>(&per_sm)));
        if (per_sm < 1) return fail(PB_ERR_UNSUPPORTED, "streaming kernel does not fit on an SM");
        s.st_grid = per_sm * c->num_sms;
        int agg_per_sm = 0;
        if (f32) PB_CUDA((configure_aggregate_kernel<float>(&agg_per_sm)));
        else PB_CUDA((configure_aggregate_kernel<double>(&agg_per_sm)));
        s.st_agg_grid = std::max(1, agg_per_sm) * c->num_sms;
    }
    // tables
    if (has_fir) PB_CUDA(cudaMalloc(&s.d_taps, el * (size_t)s.tp_len));
    if (s.bq_stage >= 0) {
        PB_CUDA(cudaMalloc(&s.d_apow, sizeof(double) * (size_t)s.apow_len * 4));
        PB_CUDA(cudaMalloc(&s.d_wt, sizeof(double) * (size_t)s.wt_len * 2));
        s.lb_tiles = (int)ceil_div64(c->max_frames, std::min(s.L, s.st_ok ? kStMinTile : kTcFrames)) + 1;
        if (s.st_ok) PB_CUDA(cudaMalloc(&s.d_st_tab, sizeof(double) * (size_t)StTab::kCount));
        if (s.st_ok) {   // block aggregates of the scan, and behind them one flag word per block (the epoch that wrote it)
            const size_t blk_bytes = sizeof(double) * 64 * (size_t)kScanBlocks * ((size_t)(c->C + kCg - 1) / kCg);
            const size_t flag_bytes = sizeof(unsigned) * (size_t)kScanBlocks * ((size_t)(c->C + kCg - 1) / kCg);
            PB_CUDA(cudaMalloc(&s.d_scan_blk, blk_bytes + flag_bytes));
            PB_CUDA(cudaMemset(s.d_scan_blk, 0, blk_bytes + flag_bytes));
        }
        const size_t groups = (size_t)(c->C + kCg - 1) / kCg;
        PB_CUDA(cudaMalloc(&s.d_agg, sizeof(double) * groups * s.lb_tiles * 64));
        PB_CUDA(cudaMalloc(&s.d_inc, sizeof(double) * groups * s.lb_tiles * 64));
        PB_CUDA(cudaMalloc((void **)&s.d_status, sizeof(unsigned) * groups * s.lb_tiles));
        PB_CUDA(cudaMemset(s.d_status, 0, sizeof(unsigned) * groups * s.lb_tiles));
    }
    if (s.rs_stage >= 0) PB_CUDA(cudaMalloc(&s.d_coef, el * (size_t)s.up * s.P));
    for (int i = 0; i < 2; i++) {
        if (s.Hf > 0) {
            PB_CUDA(cudaMalloc(&s.d_xhist[i], el * (size_t)s.Hf * c->C));
            PB_CUDA(cudaMemset(s.d_xhist[i], 0, el * (size_t)s.Hf * c->C));
        }
        if (s.Hr > 0) {
            PB_CUDA(cudaMalloc(&s.d_yhist[i], el * (size_t)s.Hr * c->C));
            PB_CUDA(cudaMemset(s.d_yhist[i], 0, el * (size_t)s.Hr * c->C));
        }
        if (s.bq_stage >= 0) {
            PB_CUDA(cudaMalloc(&s.d_state[i], sizeof(double) * (size_t)c->C * 2));
            PB_CUDA(cudaMemset(s.d_state[i], 0, sizeof(double) * (size_t)c->C * 2));
        }
    }
    s.tc_ok = tc_shape_ok(c, s);
    if (s.tc_ok) PB_CUDA(cudaMalloc(&s.d_tc_tables, (size_t)TcTables::kBytes));
    if (s.tc_ok) PB_CUDA(cudaMalloc(&s.d_tc_rc, (size_t)kTcRcRows * 8 * sizeof(float)));
    if (s.tc_ok) {
        PB_CUDA(cudaMalloc((void **)&s.d_tc_scale, sizeof(float) * 8 * (size_t)c->C));   // [2 copies][sigma, 1 / sigma][2 classes][C]
        PB_CUDA(cudaMalloc((void **)&s.d_tc_peak, sizeof(unsigned) * 4 * (size_t)c->C));  // [2 copies][2 classes][C]
        int32_t r = tc_reset_scales(c, s, c->st_compute);
        if (r != PB_OK) return r;
    }
    return refresh_segment_params(c, s);
}

static int32_t reset_segment(pb_chain *c, Segment &s)
{
    const size_t el = c->elem;
    for (int i = 0; i < 2; i++) {
        if (s.d_xhist[i]) PB_CUDA(cudaMemsetAsync(s.d_xhist[i], 0, el * (size_t)s.Hf * c->C, c->st_compute));
        if (s.d_yhist[i]) PB_CUDA(cudaMemsetAsync(s.d_yhist[i], 0, el * (size_t)s.Hr * c->C, c->st_compute));
        if (s.d_state[i]) PB_CUDA(cudaMemsetAsync(s.d_state[i], 0, sizeof(double) * (size_t)c->C * 2, c->st_compute));
    }
    s.acc = 0;
    if (s.tc_ok) return tc_reset_scales(c, s, c->st_compute);
    return PB_OK;
}

// One fused launch: `n` input frames of segment `s` from `in` to `out`.
template <typename T, int FB>
static int32_t launch_segment(pb_chain *c, Segment &s, const void *in, int64_t n, void *out, bool is_last_segment,
                              cudaStream_t stream)
{
    TileParams<T> p{};
    p.in = (const T *)in;
    p.out = (T *)out;
    p.n_frames = n;
    p.C = c->C;
    p.L = s.L;
    p.n_tiles = (int)ceil_div64(n, s.L);
    p.n_groups = (c->C + kCg - 1) / kCg;
    const bool has_fir = s.fir_stage >= 0, has_bq = s.bq_stage >= 0, has_rs = s.rs_stage >= 0;
    // fold gains whose stage is absent into the previous present point (see TileParams)
    double g_load = s.g[0], g_fir = s.g[1], g_bq = s.g[2], g_out = s.g[3];
    if (!has_fir) { g_load *= g_fir; g_fir = 1.0; }
    if (!has_bq) { (has_fir ? g_fir : g_load) *= g_bq; g_bq = 1.0; }
    if (!has_rs) { (has_bq ? g_bq : has_fir ? g_fir : g_load) *= g_out; g_out = 1.0; }
    p.g_load = (T)g_load; p.g_fir = (T)g_fir; p.g_bq = (T)g_bq; p.g_out = (T)g_out;
    p.Hf = s.Hf;
    p.has_fir = has_fir;
    p.taps_padded = (const T *)s.d_taps;
    p.tp_len = s.tp_len;
    p.tp_off = s.tp_off;
    p.xhist = (const T *)s.d_xhist[s.pp];
    p.xhist_next = (T *)s.d_xhist[s.pp ^ 1];
    p.has_bq = has_bq;
    p.b0 = s.b[0]; p.b1 = s.b[1]; p.b2 = s.b[2]; p.a1 = s.a[0]; p.a2 = s.a[1];
    p.bq_wt = (const double *)s.d_wt;
    p.bq_apow = (const double *)s.d_apow;
    p.wt_len = s.wt_len;
    p.bq_state = (const double *)s.d_state[s.pp];
    p.bq_state_next = (double *)s.d_state[s.pp ^ 1];
    p.lb_agg = (double *)s.d_agg;
    p.lb_inc = (double *)s.d_inc;
    p.lb_status = s.d_status;
    c->epoch = (c->epoch % 0x3ffffffeu) + 1u;
    p.epoch = c->epoch;
    p.has_rs = has_rs;
    p.rs_up = s.up; p.rs_down = s.down; p.rs_P = s.P; p.Hr = s.Hr;
    p.rs_acc0 = s.acc;
    p.rs_coef = (const T *)s.d_coef;
    p.yhist = (const T *)s.d_yhist[s.pp];
    p.yhist_next = (T *)s.d_yhist[s.pp ^ 1];
    const bool meter = is_last_segment && (c->flags & PB_CHAIN_METER);
    p.meter_peak = meter ? c->d_meter : nullptr;
    p.meter_sumsq = meter ? c->d_meter + c->C : nullptr;
    p.ticket = c->d_ticket;
    p.ticket_base = c->ticket_base;
    p.err_flag = reinterpret_cast<int *>(c->d_ticket + 1);
    p.vec_ok = (sizeof(T) == 4 && c->C % 4 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0) ? 1 : 0;
    if (has_bq && p.n_tiles > s.lb_tiles) return fail(PB_ERR_CAPACITY, "batch of %lld frames exceeds the chain's max_batch", (long long)n);
    const int64_t total = (int64_t)p.n_tiles * p.n_groups;
    const int grid = (int)std::min<int64_t>(total, s.grid_max);
    chain_tile_kernel<T, FB><<<grid, kTileThreads, s.smem, stream>>>(p);
    PB_CUDA(cudaGetLastError());
    c->ticket_base += (unsigned long long)total + (unsigned long long)grid;
    c->launches++;
    s.pp ^= 1;
    return PB_OK;
}

// K3: one launch of the streaming kernels (chain_stream.cuh) for a run without FIR and without resampler.
template <typename T, typename V>
static int32_t launch_segment_stream(pb_chain *c, Segment &s, const void *in, int64_t n, void *out, bool is_last_segment,
                                     cudaStream_t stream)
{
    const bool has_bq = s.bq_stage >= 0;
    const bool meter = is_last_segment && (c->flags & PB_CHAIN_METER);
    // same gain folding as K1: everything in front of the biquad is applied at load in T, everything behind it to the double result
    const double g_front = s.g[0] * s.g[1], g_back = s.g[2] * s.g[3];
    if (!has_bq && !meter && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0) {
        const int64_t nvals = n * c->C;
        const int64_t per_cta = 256 * 4 * (int64_t)(16 / sizeof(T));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(nvals, per_cta), (int64_t)c->num_sms * 8));
        stream_map_kernel<T, V><<<grid, 256, 0, stream>>>((const T *)in, (T *)out, nvals, (T)(g_front * g_back));
        PB_CUDA(cudaGetLastError());
        c->launches++;
        return PB_OK;
    }
    StreamParams<T> p{};
    p.in = (const T *)in;
    p.out = (T *)out;
    p.n_frames = n;
    p.C = c->C;
    p.n_tiles = (int)ceil_div64(n, StShape<T>::kTile);
    p.n_groups = (c->C + kCg - 1) / kCg;
    p.has_bq = has_bq;
    // biquad runs: the leading gains go into the input-side coefficients (b0, b1, b2 and A^k B), the trailing ones scale the
    // double result; the carried state keeps K1's meaning (it is driven by g x)
    p.g_load = has_bq ? (T)1 : (T)(g_front * g_back);
    p.g_bq = has_bq ? g_back : 1.0;
    p.b0 = s.b[0] * g_front; p.b1 = s.b[1] * g_front; p.b2 = s.b[2] * g_front; p.a1 = s.a[0]; p.a2 = s.a[1];
    for (int k = 0; k < 32; k++) {
        p.wt[k][0] = s.st_wt[k][0] * g_front;
        p.wt[k][1] = s.st_wt[k][1] * g_front;
    }
    p.vec_ok = (c->C % (int)(16 / sizeof(T)) == 0 && ((uintptr_t)in % 16) == 0) ? 1 : 0;
    p.tab = (const double *)s.d_st_tab;
    p.bq_state = (const double *)s.d_state[s.pp];
    p.bq_state_next = (double *)s.d_state[s.pp ^ 1];
    p.lb_agg = (double *)s.d_agg;
    p.lb_inc = (double *)s.d_inc;
    p.lb_status = s.d_status;
    c->epoch = (c->epoch % 0x3ffffffeu) + 1u;
    p.epoch = c->epoch;
    p.meter_peak = meter ? c->d_meter : nullptr;
    p.meter_sumsq = meter ? c->d_meter + c->C : nullptr;
    p.ticket = c->d_ticket;
    p.ticket_base = c->ticket_base;
    p.err_flag = reinterpret_cast<int *>(c->d_ticket + 1);
    if (has_bq && p.n_tiles > s.lb_tiles) return fail(PB_ERR_CAPACITY, "batch of %lld frames exceeds the chain's max_batch", (long long)n);
    const int64_t total = (int64_t)p.n_tiles * p.n_groups;
    // PB_ST_CTAS_PER_SM=<n>: fewer resident CTAs (fewer tiles in flight, a shorter look-back), for measurements
    static const int ctas_per_sm = getenv("PB_ST_CTAS_PER_SM") ? atoi(getenv("PB_ST_CTAS_PER_SM")) : 0;
    const int grid = (int)std::min<int64_t>(total, ctas_per_sm > 0 ? std::min(s.st_grid, ctas_per_sm * c->num_sms) : s.st_grid);
    // With few channel groups hundreds of tiles of one group are in flight and the look-back of a single sweep is what the launch
    // waits for: two sweeps with a scan in between instead (chain_stream.cuh, kStOneSweep).  PB_ST_TWO_SWEEPS=<groups> moves the
    // threshold (0: never), for measurements.
    // (at most 8 groups: the scan's n_groups x kScanBlocks CTAs wait for each other and must be co-resident)
    static const int two_sweep_groups = std::min(8, getenv("PB_ST_TWO_SWEEPS") ? atoi(getenv("PB_ST_TWO_SWEEPS")) : 4);
    if (has_bq && p.n_groups <= two_sweep_groups && p.n_tiles >= 256) {
        // neither sweep draws tickets (static tile schedules), so the chain's ticket counter does not move
        const int agg_grid = (int)std::min<int64_t>((int64_t)(p.n_tiles - 1) * p.n_groups, (int64_t)s.st_agg_grid);
        if (c->C == 64) stream_aggregate_kernel<T, 64><<<agg_grid, kStThreads, 0, stream>>>(p);
        else if (c->C == 256) stream_aggregate_kernel<T, 256><<<agg_grid, kStThreads, 0, stream>>>(p);
        else stream_aggregate_kernel<T, 0><<<agg_grid, kStThreads, 0, stream>>>(p);
        PB_CUDA(cudaGetLastError());
        // block aggregates: the first kScanBlocks slots of the (otherwise unused) inclusive-state array of the LAST tile row would
        // alias live data, so they have their own scratch behind the look-back arrays
        const dim3 sgrid((unsigned)p.n_groups, (unsigned)kScanBlocks);
        ScanParams sp{};
        sp.agg = p.lb_agg;
        sp.inc = p.lb_inc;
        sp.blk = (double *)s.d_scan_blk;
        sp.flags = reinterpret_cast<unsigned *>(sp.blk + 64 * (size_t)kScanBlocks * (size_t)p.n_groups);
        sp.bq_state = p.bq_state;
        sp.tab = p.tab;
        sp.err_flag = p.err_flag;
        sp.epoch = p.epoch;
        sp.C = c->C;
        sp.n_tiles = p.n_tiles;
        sp.n_full = p.n_tiles - 1;
        sp.span = (sp.n_full + kScanBlocks - 1) / kScanBlocks;
        sp.per = (sp.span + kScanWarps - 1) / kScanWarps;
        {   // powers of the tile step over a segment and over a block (2x2, double, on the host)
            auto mpow = [](const double *M, int e, double *P) {
                double B[4] = {M[0], M[1], M[2], M[3]}, R[4] = {1.0, 0.0, 0.0, 1.0};
                for (; e > 0; e >>= 1) {
                    if (e & 1) {
                        const double r[4] = {R[0] * B[0] + R[1] * B[2], R[0] * B[1] + R[1] * B[3], R[2] * B[0] + R[3] * B[2], R[2] * B[1] + R[3] * B[3]};
                        for (int i = 0; i < 4; i++) R[i] = r[i];
                    }
                    const double q[4] = {B[0] * B[0] + B[1] * B[2], B[0] * B[1] + B[1] * B[3], B[2] * B[0] + B[3] * B[2], B[2] * B[1] + B[3] * B[3]};
                    for (int i = 0; i < 4; i++) B[i] = q[i];
                }
                for (int i = 0; i < 4; i++) P[i] = R[i];
            };
            mpow(s.st_tile_step, sp.per, sp.Mper);
            mpow(s.st_tile_step, sp.span, sp.Mspan);
        }
        {   // cooperative for the same reason as K2: a block waits for the blocks in front of it, the grid must be resident at once
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = sgrid;
            cfg.blockDim = dim3(kScanWarps * 32);
            cfg.stream = stream;
            cudaLaunchAttribute coop[1];
            coop[0].id = cudaLaunchAttributeCooperative;
            coop[0].val.cooperative = 1;
            cfg.attrs = coop;
            cfg.numAttrs = 1;
            PB_CUDA(cudaLaunchKernelEx(&cfg, stream_scan_kernel, sp));
        }
        launch_stream_mode<T, kStApply>(p, grid, c->C, stream);
        PB_CUDA(cudaGetLastError());
        c->launches += 3;
    } else {
        launch_stream_mode<T, kStOneSweep>(p, grid, c->C, stream);
        PB_CUDA(cudaGetLastError());
        c->ticket_base += (unsigned long long)total + (unsigned long long)grid * kStTicketsPerCta;
        c->launches++;
    }
    if (has_bq) s.pp ^= 1;
    return PB_OK;
}

// per-buffer output frame counts through every resampling segment (integer bookkeeping)
static void count_outputs(pb_chain *c, const int64_t *buf_frames, int n_buffers, int64_t *buf_out, int64_t *total_in,
                          int64_t *total_out, bool commit)
{
    std::vector<int64_t> acc(c->segs.size());
    for (size_t i = 0; i < c->segs.size(); i++) acc[i] = c->segs[i].acc;
    int64_t tin = 0, tout = 0;
    for (int bi = 0; bi < n_buffers; bi++) {
        int64_t n = buf_frames[bi];
        tin += n;
        for (size_t i = 0; i < c->segs.size(); i++) {
            const Segment &s = c->segs[i];
            if (s.rs_stage < 0) continue;
            const int64_t tot = acc[i] + n * s.up;
            n = tot / s.down;
            acc[i] = tot % s.down;
        }
        if (buf_out) buf_out[bi] = n;
        tout += n;
    }
    if (commit)
        for (size_t i = 0; i < c->segs.size(); i++) c->segs[i].acc = acc[i];
    *total_in = tin;
    *total_out = tout;
}

static int32_t run_batch_device(pb_chain *c, const void *in_dev, const int64_t *buf_frames, int n_buffers, void *out_dev,
                                int64_t out_capacity_frames, int64_t *buf_out_frames, cudaStream_t stream)
{
    if (!buf_frames || n_buffers < 1) return fail(PB_ERR_INVALID, "process: need at least one buffer");
    if (n_buffers > c->max_batch) return fail(PB_ERR_CAPACITY, "process: %d buffers > max_batch %d", n_buffers, c->max_batch);
    for (int i = 0; i < n_buffers; i++) {
        if (buf_frames[i] < 0 || buf_frames[i] > c->buffer_frames)
            return fail(PB_ERR_INVALID, "process: buffer %d has %lld frames (bufferSize %d)", i, (long long)buf_frames[i], c->buffer_frames);
        // only the final buffer of a stream may be short (Source.execute, pipe.go:404-406)
        if (i + 1 < n_buffers && buf_frames[i] != c->buffer_frames)
            return fail(PB_ERR_INVALID, "process: only the last buffer of a batch may be short");
    }
    int64_t tin = 0, tout = 0;
    count_outputs(c, buf_frames, n_buffers, nullptr, &tin, &tout, false);
    if (tout > out_capacity_frames) return fail(PB_ERR_CAPACITY, "process: output needs %lld frames, capacity %lld", (long long)tout, (long long)out_capacity_frames);
    if (tin > 0 && (!in_dev || !out_dev)) return fail(PB_ERR_INVALID, "process: NULL buffer");
    int path = 0;
    if (tin > 0) {
        const void *src = in_dev;
        int64_t n = tin;
        for (size_t i = 0; i < c->segs.size(); i++) {
            Segment &s = c->segs[i];
            const bool lastseg = (i + 1 == c->segs.size());
            void *dst = lastseg ? out_dev : c->d_mid[i & 1];
            auto generic = [&](const void *from, int64_t frames, void *to) -> int32_t {
                return s.st_ok ? (c->dtype == PB_F32 ? launch_segment_stream<float, float4>(c, s, from, frames, to, lastseg, stream)
                                                     : launch_segment_stream<double, double2>(c, s, from, frames, to, lastseg, stream))
                       : c->dtype == PB_F32 ? launch_segment<float, 16>(c, s, from, frames, to, lastseg, stream)
                                            : launch_segment<double, 8>(c, s, from, frames, to, lastseg, stream);
            };
            // K2 works on 160-frame tiles that start where the resampler phase is 0, i.e. at stream positions that are multiples
            // of 160.  A call that does not start or end there (a 4096-frame buffer never does both) is cut into a head up to the
            // next such position, the aligned middle and a tail; head and tail go through K1, which shares every piece of
            // carried state with K2 (launches on one stream are ordered like separate calls).
            // (the f pieces of MMA2 are fp16: the FIR output of a channel at its grid peak, |g_fir| sum|h| * 2047, must stay inside)
            const bool tc_eligible = s.tc_ok && s.g[2] != 0.0 && s.g[0] != 0.0 && s.tc_level_ok && std::fabs(s.g[1]) * s.tc_fir_l1 <= 28.0 &&
                                     ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0;
            // The tiles start at the first frame of the call, with the tables of the resampler phase found there (160 frames give
            // 147 outputs from any phase); what is left behind the last whole tile (< 160 frames) goes through K1.  Only if the
            // slice schedule did not cover a phase would the call be cut at the next phase-0 position instead (K1 head).
            int64_t head = 0;
            const Segment::TcPhase *ph = nullptr;
            int32_t r = PB_OK;
            if (tc_eligible && n >= kTcFrames) {
                r = tc_get_phase(c, s, (int)s.acc, &ph);
                if (r != PB_OK) return r;
                if (!ph) {
                    while (head < s.down && (s.acc + head * s.up) % s.down != 0) head++;
                    if (n - head >= kTcFrames) {
                        r = tc_get_phase(c, s, 0, &ph);
                        if (r != PB_OK) return r;
                    }
                }
            }
            // PB_TC_TAIL_K1=1 (development): whole tiles only, the rest of the call on K1 as before
            static const bool tail_k1 = getenv("PB_TC_TAIL_K1") && atoi(getenv("PB_TC_TAIL_K1")) != 0;
            const int64_t mid = !ph ? 0 : tail_k1 ? ((n - head) / kTcFrames) * kTcFrames : n - head;
            if (mid > 0) {
                const int64_t acc0 = s.acc, pieces[3] = {head, mid, n - head - mid};
                const char *from = (const char *)src;
                char *to = (char *)dst;
                for (int k = 0; k < 3 && r == PB_OK; k++) {
                    const int64_t len = pieces[k];
                    if (len == 0) continue;
                    r = (k == 1) ? launch_segment_tc(c, s, *ph, from, len, to, lastseg, stream) : generic(from, len, to);
                    const int64_t tot = s.acc + len * s.up;  // the phase each piece starts from, restored below
                    from += (size_t)len * c->C * c->elem;
                    to += (size_t)(tot / s.down) * c->C * c->elem;
                    s.acc = tot % s.down;
                }
                s.acc = acc0;  // committed for the whole call by count_outputs
            } else {
                r = generic(src, n, dst);
            }
            const bool use_tc = mid > 0;
            if (r != PB_OK) return r;
            path = std::max(path, use_tc ? 2 : s.st_ok ? 3 : 1);
            if (s.rs_stage >= 0) n = (s.acc + n * s.up) / s.down;  // s.acc is committed below, after every launch used it
            src = dst;
            if (n == 0 && !lastseg) {
                // nothing left to feed downstream this call; later segments keep their state
                break;
            }
        }
        c->last_path = path;
    }
    count_outputs(c, buf_frames, n_buffers, buf_out_frames, &tin, &tout, true);
    c->meter_frames += tout;
    return PB_OK;
}

}  // namespace pb

// =============================================================================
//                                   C-ABI
// =============================================================================

extern "C" int32_t pb_abi_version(void) { return PB_ABI_VERSION; }

extern "C" const char *pb_last_error(void) { return tls_error_buf(); }

extern "C" int32_t pb_chain_destroy(pb_chain *c)
{
    if (!c) return PB_OK;
    DeviceGuard dg(c->device);
    cudaDeviceSynchronize();
    for (auto &s : c->segs) free_segment(s);
    for (int i = 0; i < 2; i++) {
        if (c->d_mid[i]) cudaFree(c->d_mid[i]);
        Slot &sl = c->slots[i];
        if (sl.h_in) cudaFreeHost(sl.h_in);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
    }
    for (auto &row : c->ev_piece)
        for (auto &ev : row)
            if (ev) cudaEventDestroy(ev);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->d_meter) cudaFree(c->d_meter);
    if (c->d_meter_scratch) cudaFree(c->d_meter_scratch);
    if (c->st_compute) cudaStreamDestroy(c->st_compute);
    if (c->st_h2d) cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
    delete c;
    return PB_OK;
}

extern "C" int32_t pb_chain_create(const pb_chain_desc *desc, pb_chain **out)
{
    if (!desc || !out) return fail(PB_ERR_INVALID, "pb_chain_create: NULL argument");
    *out = nullptr;
    if (desc->abi_version != PB_ABI_VERSION) return fail(PB_ERR_INVALID, "pb_chain_create: abi_version %d != %d", desc->abi_version, PB_ABI_VERSION);
    if (desc->dtype != PB_F32 && desc->dtype != PB_F64) return fail(PB_ERR_INVALID, "pb_chain_create: bad dtype");
    if (desc->channels < 1 || desc->buffer_frames < 1 || desc->max_batch < 1 || desc->n_stages < 0 ||
        (desc->n_stages > 0 && !desc->stages))
        return fail(PB_ERR_INVALID, "pb_chain_create: bad channels/buffer_frames/max_batch/stages");
    for (int i = 0; i < desc->n_stages; i++) {
        int32_t r = validate_stage(desc->stages[i], i);
        if (r != PB_OK) return r;
    }
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev < 1)
            return fail(PB_ERR_NO_DEVICE, "no CUDA device (%s); pipe_b200 has no CPU fallback", cudaGetErrorString(e));
    }
    if (desc->device < 0 || desc->device >= ndev) return fail(PB_ERR_INVALID, "pb_chain_create: device %d of %d", desc->device, ndev);
    DeviceGuard dg(desc->device);
    PB_CUDA(dg.err);
    cudaDeviceProp prop{};
    PB_CUDA(cudaGetDeviceProperties(&prop, desc->device));
    if (prop.major < 10)
        return fail(PB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", desc->device, prop.major, prop.minor);

    pb_chain *c = new (std::nothrow) pb_chain();
    if (!c) return fail(PB_ERR_NOMEM, "out of host memory");
    c->device = desc->device;
    c->dtype = desc->dtype;
    c->elem = desc->dtype == PB_F32 ? 4 : 8;
    c->C = desc->channels;
    c->buffer_frames = desc->buffer_frames;
    c->max_batch = desc->max_batch;
    c->flags = (unsigned)desc->flags;
    c->sample_rate = desc->sample_rate;
    c->max_frames = (int64_t)desc->buffer_frames * desc->max_batch;
    c->num_sms = prop.multiProcessorCount;
    c->stages.resize((size_t)desc->n_stages);
    for (int i = 0; i < desc->n_stages; i++) {
        c->stages[i].d = desc->stages[i];
        if (desc->stages[i].taps && desc->stages[i].n_taps > 0)
            c->stages[i].taps.assign(desc->stages[i].taps, desc->stages[i].taps + desc->stages[i].n_taps);
        c->stages[i].d.taps = nullptr;  // never retain a caller pointer
    }
    auto bail = [&](int32_t r) {
        char keep[512];
        snprintf(keep, sizeof keep, "%s", tls_error_buf());
        pb_chain_destroy(c);
        snprintf(tls_error_buf(), 512, "%s", keep);
        return r;
    };
#define PB_TRY(expr)                         \
    do {                                     \
        int32_t _r = (expr);                 \
        if (_r != PB_OK) return bail(_r);    \
    } while (0)
#define PB_TRY_CUDA(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return bail(fail(_e == cudaErrorMemoryAllocation ? PB_ERR_NOMEM : PB_ERR_CUDA, "%s failed: %s", #expr, \
                             cudaGetErrorString(_e)));                                             \
    } while (0)

    PB_TRY_CUDA(cudaStreamCreateWithFlags(&c->st_compute, cudaStreamNonBlocking));
    PB_TRY_CUDA(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
    PB_TRY_CUDA(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
    PB_TRY_CUDA(cudaMalloc((void **)&c->d_ticket, 2 * sizeof(unsigned long long)));
    PB_TRY_CUDA(cudaMemset(c->d_ticket, 0, 2 * sizeof(unsigned long long)));
    if (c->flags & PB_CHAIN_METER) {
        PB_TRY_CUDA(cudaMalloc((void **)&c->d_meter, sizeof(double) * 2 * (size_t)c->C));
        PB_TRY_CUDA(cudaMemset(c->d_meter, 0, sizeof(double) * 2 * (size_t)c->C));
        PB_TRY_CUDA(cudaMalloc((void **)&c->d_meter_scratch, sizeof(double) * 2 * (size_t)c->C));
        PB_TRY_CUDA(cudaMemset(c->d_meter_scratch, 0, sizeof(double) * 2 * (size_t)c->C));
    }
    plan_segments(c);
    double rate = c->sample_rate;
    for (auto &s : c->segs) {
        PB_TRY(build_segment(c, s));
        if (s.rs_stage >= 0) rate = rate * s.up / s.down;
    }
    c->out_sample_rate = rate;
    if (c->segs.size() > 1)
        for (int i = 0; i < 2; i++) PB_TRY_CUDA(cudaMalloc(&c->d_mid[i], c->elem * (size_t)c->max_frames * c->C));
    PB_TRY_CUDA(cudaDeviceSynchronize());
#undef PB_TRY
#undef PB_TRY_CUDA
    *out = c;
    return PB_OK;
}

extern "C" int32_t pb_chain_reset(pb_chain *c)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_reset: NULL chain");
    if (c->slots_busy) return fail(PB_ERR_STATE, "pb_chain_reset: %d submitted batches not collected", c->slots_busy);
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaDeviceSynchronize());
    for (auto &s : c->segs) {
        int32_t r = reset_segment(c, s);
        if (r != PB_OK) return r;
    }
    if (c->d_meter) PB_CUDA(cudaMemsetAsync(c->d_meter, 0, sizeof(double) * 2 * (size_t)c->C, c->st_compute));
    if (c->d_meter_scratch) PB_CUDA(cudaMemsetAsync(c->d_meter_scratch, 0, sizeof(double) * 2 * (size_t)c->C, c->st_compute));
    // a look-back timeout (the only kernel-side error) is not sticky: a reset chain starts clean
    PB_CUDA(cudaMemsetAsync(c->d_ticket + 1, 0, sizeof(unsigned long long), c->st_compute));
    c->meter_frames = 0;
    PB_CUDA(cudaStreamSynchronize(c->st_compute));
    return PB_OK;
}

extern "C" int32_t pb_chain_out_properties(const pb_chain *c, int32_t *channels, double *sample_rate)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_out_properties: NULL chain");
    if (channels) *channels = c->C;
    if (sample_rate) *sample_rate = c->out_sample_rate;
    return PB_OK;
}

extern "C" int32_t pb_chain_peek_out_frames(const pb_chain *c, int64_t in_frames, int64_t *out_frames)
{
    if (!c || !out_frames || in_frames < 0) return fail(PB_ERR_INVALID, "pb_chain_peek_out_frames: bad argument");
    int64_t n = in_frames;
    for (const auto &s : c->segs)
        if (s.rs_stage >= 0) n = (s.acc + n * s.up) / s.down;
    *out_frames = n;
    return PB_OK;
}

extern "C" int32_t pb_chain_process_batch_device(pb_chain *c, const void *in_dev, const int64_t *buf_frames,
                                                 int32_t n_buffers, void *out_dev, int64_t out_capacity_frames,
                                                 int64_t *buf_out_frames, void *stream)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_process_batch_device: NULL chain");
    if (c->slots_busy) return fail(PB_ERR_STATE, "process while %d submitted batches are in flight", c->slots_busy);
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    return run_batch_device(c, in_dev, buf_frames, n_buffers, out_dev, out_capacity_frames, buf_out_frames,
                            (cudaStream_t)stream);
}

// The only kernel-side failure left is a look-back wait that ran into its bound (a tile's predecessors did not publish
// for seconds).  K1 and the one-sweep K3 claim tiles in ticket order and K2 / the K3 scan are launched cooperatively, so
// the CTAs a tile waits for are always resident: the bound is a guard against a fault, not a scheduling hazard.  It is
// reported once and cleared, so the chain can be reset and used again; what the affected call wrote is undefined.
static int32_t take_kernel_error(pb_chain *c)
{
    int flag = 0;
    PB_CUDA(cudaMemcpy(&flag, c->d_ticket + 1, sizeof(int), cudaMemcpyDeviceToHost));
    if (!flag) return PB_OK;
    PB_CUDA(cudaMemset(c->d_ticket + 1, 0, sizeof(unsigned long long)));
    return fail(PB_ERR_CUDA, "fused kernel: look-back wait timed out (kernel error flag %d); the carried state is undefined, call pb_chain_reset", flag);
}

extern "C" int32_t pb_chain_sync(pb_chain *c, void *stream)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_sync: NULL chain");
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    PB_CUDA(cudaStreamSynchronize(c->st_compute));
    return take_kernel_error(c);
}

static int32_t ensure_slot(pb_chain *c, Slot &sl)
{
    const size_t bytes = c->elem * (size_t)c->max_frames * c->C;
    if (!sl.d_in) PB_CUDA(cudaMalloc(&sl.d_in, bytes));
    if (!sl.d_out) PB_CUDA(cudaMalloc(&sl.d_out, bytes));
    if (!sl.ev_h2d) PB_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
    if (!sl.ev_done) PB_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    if (!sl.ev_d2h) PB_CUDA(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
    return PB_OK;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int32_t pb_chain_pipeline_depth(const pb_chain *c) { return c ? 2 : 0; }

extern "C" int32_t pb_chain_submit(pb_chain *c, const void *in_host, const int64_t *buf_frames, int32_t n_buffers,
                                   void *out_host, int64_t out_capacity_frames)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_submit: NULL chain");
    if (c->slots_busy >= 2) return fail(PB_ERR_STATE, "pb_chain_submit: pipeline full, collect first");
    if (!buf_frames || n_buffers < 1) return fail(PB_ERR_INVALID, "pb_chain_submit: need at least one buffer");
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    Slot &sl = c->slots[c->slot_head];
    int32_t r = ensure_slot(c, sl);
    if (r != PB_OK) return r;
    // validate the whole call before anything is enqueued: a rejected call must leave the carried state untouched
    if (n_buffers > c->max_batch) return fail(PB_ERR_CAPACITY, "pb_chain_submit: %d buffers > max_batch %d", n_buffers, c->max_batch);
    for (int i = 0; i < n_buffers; i++) {
        if (buf_frames[i] < 0 || buf_frames[i] > c->buffer_frames)
            return fail(PB_ERR_INVALID, "pb_chain_submit: buffer %d has %lld frames (bufferSize %d)", i, (long long)buf_frames[i], c->buffer_frames);
        if (i + 1 < n_buffers && buf_frames[i] != c->buffer_frames)
            return fail(PB_ERR_INVALID, "pb_chain_submit: only the last buffer of a batch may be short");
    }
    int64_t tin = 0, tout_need = 0;
    count_outputs(c, buf_frames, n_buffers, nullptr, &tin, &tout_need, false);
    if (tin > c->max_frames) return fail(PB_ERR_CAPACITY, "pb_chain_submit: %lld frames > buffer_frames*max_batch", (long long)tin);
    if (tout_need > out_capacity_frames)
        return fail(PB_ERR_CAPACITY, "pb_chain_submit: output needs %lld frames, capacity %lld", (long long)tout_need, (long long)out_capacity_frames);
    const size_t in_bytes = c->elem * (size_t)tin * c->C;
    if (tin > 0 && (!in_host || !out_host)) return fail(PB_ERR_INVALID, "pb_chain_submit: NULL buffer");
    // H2D: pinned caller memory goes direct, pageable memory (the Go heap) is
    // staged first so that no caller pointer outlives this call.
    if (in_bytes) {
        const void *src = in_host;
        if (!is_pinned(in_host)) {
            if (!sl.h_in) PB_CUDA(cudaMallocHost(&sl.h_in, c->elem * (size_t)c->max_frames * c->C));
            memcpy(sl.h_in, in_host, in_bytes);
            src = sl.h_in;
        }
        PB_CUDA(cudaMemcpyAsync(sl.d_in, src, in_bytes, cudaMemcpyHostToDevice, c->st_h2d));
    }
    PB_CUDA(cudaEventRecord(sl.ev_h2d, c->st_h2d));
    PB_CUDA(cudaStreamWaitEvent(c->st_compute, sl.ev_h2d, 0));
    // the slot's previous D2H must have drained before d_out is overwritten
    PB_CUDA(cudaStreamWaitEvent(c->st_compute, sl.ev_d2h, 0));
    sl.out_counts.assign((size_t)n_buffers, 0);
    r = run_batch_device(c, sl.d_in, buf_frames, n_buffers, sl.d_out, c->max_frames, sl.out_counts.data(), c->st_compute);
    if (r != PB_OK) return r;
    int64_t tout = 0;
    for (auto v : sl.out_counts) tout += v;
    PB_CUDA(cudaEventRecord(sl.ev_done, c->st_compute));
    PB_CUDA(cudaStreamWaitEvent(c->st_d2h, sl.ev_done, 0));
    sl.out_frames = tout;
    sl.user_out = nullptr;
    const size_t out_bytes = c->elem * (size_t)tout * c->C;
    if (out_bytes) {
        void *dst = out_host;
        if (!is_pinned(out_host)) {
            if (!sl.h_out) PB_CUDA(cudaMallocHost(&sl.h_out, c->elem * (size_t)c->max_frames * c->C));
            dst = sl.h_out;
            sl.user_out = out_host;
        }
        PB_CUDA(cudaMemcpyAsync(dst, sl.d_out, out_bytes, cudaMemcpyDeviceToHost, c->st_d2h));
    }
    PB_CUDA(cudaEventRecord(sl.ev_d2h, c->st_d2h));
    sl.busy = true;
    c->slot_head ^= 1;
    c->slots_busy++;
    return PB_OK;
}

extern "C" int32_t pb_chain_collect(pb_chain *c, int64_t *buf_out_frames, int32_t n_buffers)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_collect: NULL chain");
    if (c->slots_busy < 1) return fail(PB_ERR_STATE, "pb_chain_collect: nothing submitted");
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    Slot &sl = c->slots[c->slot_tail];
    if (n_buffers != (int)sl.out_counts.size()) return fail(PB_ERR_INVALID, "pb_chain_collect: batch had %zu buffers", sl.out_counts.size());
    PB_CUDA(cudaEventSynchronize(sl.ev_d2h));
    const int32_t kerr = take_kernel_error(c);
    if (kerr == PB_OK) {
        if (sl.user_out && sl.out_frames) memcpy(sl.user_out, sl.h_out, c->elem * (size_t)sl.out_frames * c->C);
        if (buf_out_frames)
            for (int i = 0; i < n_buffers; i++) buf_out_frames[i] = sl.out_counts[(size_t)i];
    }
    // the slot is released whether or not the batch succeeded: a failed batch must not wedge the pipeline
    sl.busy = false;
    sl.user_out = nullptr;
    c->slot_tail ^= 1;
    c->slots_busy--;
    return kerr;
}

// ONE large host buffer (the ProcessFunc call of pipe.go:438: synchronous, nothing else in flight) cut into pieces that follow
// each other through H2D -> kernels -> D2H on the chain's three streams: the copy of piece i+1 in, the kernels of piece i and
// the copy of piece i-1 out overlap (PCIe is full duplex), and for pageable callers so do the staging memcpys on the host.  To the chain a piece is a call
// of its own (the carried state crosses calls by construction; pieces are multiples of K2's 160-frame tile, so the resampler
// phase at the start of every piece is the phase at the start of the buffer).  Measured per 4096 x 1024 buffer from pinned memory
// (tools/process_time.py): K2 f32 0.651 -> 0.541 ms, K3 f32 0.653 -> 0.522 ms, K1 f64 1.83 -> 1.00 ms; eight pieces are slower than four.
static int32_t process_pieces(pb_chain *c, const char *in_host, int64_t in_frames, char *out_host, int64_t out_capacity_frames,
                              int64_t *out_frames, int64_t piece)
{
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    Slot &sl = c->slots[c->slot_head];
    int32_t r = ensure_slot(c, sl);
    if (r != PB_OK) return r;
    // validate the whole call before anything is enqueued: a rejected call must leave the carried state untouched
    int64_t tin = 0, tout_need = 0;
    count_outputs(c, &in_frames, 1, nullptr, &tin, &tout_need, false);
    if (tout_need > out_capacity_frames)
        return fail(PB_ERR_CAPACITY, "pb_chain_process: output needs %lld frames, capacity %lld", (long long)tout_need, (long long)out_capacity_frames);
    if (!in_host || !out_host) return fail(PB_ERR_INVALID, "pb_chain_process: NULL buffer");
    const bool pin_in = is_pinned(in_host), pin_out = is_pinned(out_host);
    const size_t slot_bytes = c->elem * (size_t)c->max_frames * c->C, frame_bytes = c->elem * (size_t)c->C;
    if (!pin_in && !sl.h_in) PB_CUDA(cudaMallocHost(&sl.h_in, slot_bytes));
    if (!pin_out && !sl.h_out) PB_CUDA(cudaMallocHost(&sl.h_out, slot_bytes));
    const int n_pieces = (int)((in_frames + piece - 1) / piece);
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < n_pieces; i++)
            if (!c->ev_piece[k][i]) PB_CUDA(cudaEventCreateWithFlags(&c->ev_piece[k][i], cudaEventDisableTiming));
    int64_t out_done[pb_chain::kMaxPieces + 1] = {0};
    int32_t rc = PB_OK;
    int enq = 0;
    for (int i = 0; i < n_pieces && rc == PB_OK; i++) {
        const int64_t f0 = (int64_t)i * piece, n = std::min(piece, in_frames - f0);
        const size_t off = (size_t)f0 * frame_bytes, nbytes = (size_t)n * frame_bytes;
        const char *src = in_host + off;
        if (!pin_in) {
            memcpy((char *)sl.h_in + off, src, nbytes);
            src = (const char *)sl.h_in + off;
        }
        auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == PB_OK) rc = fail(PB_ERR_CUDA, "pb_chain_process: %s", cudaGetErrorString(e)); };
        cu(cudaMemcpyAsync((char *)sl.d_in + off, src, nbytes, cudaMemcpyHostToDevice, c->st_h2d));
        cu(cudaEventRecord(c->ev_piece[0][i], c->st_h2d));
        cu(cudaStreamWaitEvent(c->st_compute, c->ev_piece[0][i], 0));
        if (rc != PB_OK) break;
        int64_t cnt = 0;
        const size_t ooff = (size_t)out_done[i] * frame_bytes;
        rc = run_batch_device(c, (const char *)sl.d_in + off, &n, 1, (char *)sl.d_out + ooff, c->max_frames - out_done[i], &cnt, c->st_compute);
        if (rc != PB_OK) break;
        out_done[i + 1] = out_done[i] + cnt;
        cu(cudaEventRecord(c->ev_piece[1][i], c->st_compute));
        cu(cudaStreamWaitEvent(c->st_d2h, c->ev_piece[1][i], 0));
        if (cnt) cu(cudaMemcpyAsync((pin_out ? out_host : (char *)sl.h_out) + ooff, (char *)sl.d_out + ooff, (size_t)cnt * frame_bytes, cudaMemcpyDeviceToHost, c->st_d2h));
        cu(cudaEventRecord(c->ev_piece[2][i], c->st_d2h));
        enq = i + 1;
    }
    // drain piece by piece (a pageable destination is filled while the later pieces are still on their way)
    for (int i = 0; i < enq; i++) {
        const cudaError_t e = cudaEventSynchronize(c->ev_piece[2][i]);
        if (e != cudaSuccess && rc == PB_OK) rc = fail(PB_ERR_CUDA, "pb_chain_process: %s", cudaGetErrorString(e));
        if (rc == PB_OK && !pin_out && out_done[i + 1] > out_done[i])
            memcpy(out_host + (size_t)out_done[i] * frame_bytes, (char *)sl.h_out + (size_t)out_done[i] * frame_bytes,
                   (size_t)(out_done[i + 1] - out_done[i]) * frame_bytes);
    }
    if (rc != PB_OK) {
        cudaStreamSynchronize(c->st_h2d);
        cudaStreamSynchronize(c->st_compute);
        cudaStreamSynchronize(c->st_d2h);
        return rc;
    }
    PB_CUDA(cudaEventRecord(sl.ev_d2h, c->st_d2h));   // the slot protocol of submit / collect: this slot's last D2H
    r = take_kernel_error(c);
    if (r != PB_OK) return r;
    if (out_frames) *out_frames = out_done[enq];
    return PB_OK;
}

extern "C" int32_t pb_chain_process(pb_chain *c, const void *in_host, int64_t in_frames, void *out_host,
                                    int64_t out_capacity_frames, int64_t *out_frames)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_process: NULL chain");
    if (c->slots_busy) return fail(PB_ERR_STATE, "pb_chain_process while submitted batches are in flight");
    // A single call may carry up to max_batch buffers' worth of frames.
    if (in_frames < 0 || in_frames > c->max_frames) return fail(PB_ERR_INVALID, "pb_chain_process: %lld frames", (long long)in_frames);
    {
        // one buffer of >= 8 MiB whose rows keep every piece 16-byte aligned: four pieces (multiples of 160 frames) in a row
        static const int want = getenv("PB_PROCESS_PIECES") ? atoi(getenv("PB_PROCESS_PIECES")) : 4;   // 1: off (development)
        const size_t frame_bytes = c->elem * (size_t)c->C;
        if (want > 1 && want <= pb_chain::kMaxPieces && in_frames <= c->buffer_frames && frame_bytes % 16 == 0 &&
            (size_t)in_frames * frame_bytes >= ((size_t)8 << 20)) {
            const int64_t piece = ((in_frames + want - 1) / want + kTcFrames - 1) / kTcFrames * kTcFrames;
            if (piece < in_frames)
                return process_pieces(c, (const char *)in_host, in_frames, (char *)out_host, out_capacity_frames, out_frames, piece);
        }
    }
    std::vector<int64_t> bf;
    int64_t left = in_frames;
    do {
        const int64_t n = left < c->buffer_frames ? left : c->buffer_frames;
        bf.push_back(n);
        left -= n;
    } while (left > 0);
    int32_t r = pb_chain_submit(c, in_host, bf.data(), (int32_t)bf.size(), out_host, out_capacity_frames);
    if (r != PB_OK) return r;
    std::vector<int64_t> oc(bf.size());
    r = pb_chain_collect(c, oc.data(), (int32_t)oc.size());
    if (r != PB_OK) return r;
    int64_t tot = 0;
    for (auto v : oc) tot += v;
    if (out_frames) *out_frames = tot;
    return PB_OK;
}

extern "C" int32_t pb_chain_set_stage(pb_chain *c, int32_t idx, const pb_stage_desc *st)
{
    if (!c || !st) return fail(PB_ERR_INVALID, "pb_chain_set_stage: NULL argument");
    if (idx < 0 || idx >= (int)c->stages.size()) return fail(PB_ERR_INVALID, "pb_chain_set_stage: stage %d of %zu", idx, c->stages.size());
    if (c->slots_busy) return fail(PB_ERR_STATE, "pb_chain_set_stage while submitted batches are in flight");
    StageCopy &cur = c->stages[(size_t)idx];
    if (st->kind != cur.d.kind) return fail(PB_ERR_INVALID, "pb_chain_set_stage: kind may not change");
    int32_t r = validate_stage(*st, idx);
    if (r != PB_OK) return r;
    if ((st->kind == PB_STAGE_FIR || st->kind == PB_STAGE_RESAMPLE) &&
        (st->n_taps != cur.d.n_taps || st->up != cur.d.up || st->down != cur.d.down))
        return fail(PB_ERR_INVALID, "pb_chain_set_stage: n_taps/up/down may not change");
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    // mutations land between buffers (pipe.go:433): drain what is in flight first
    PB_CUDA(cudaDeviceSynchronize());
    cur.d = *st;
    if (st->taps && st->n_taps > 0) cur.taps.assign(st->taps, st->taps + st->n_taps);
    cur.d.taps = nullptr;
    for (auto &s : c->segs) {
        bool mine = s.fir_stage == idx || s.bq_stage == idx || s.rs_stage == idx;
        for (auto &gs : s.gain_stages) mine = mine || gs.first == idx;
        if (mine) return refresh_segment_params(c, s);
    }
    return PB_OK;  // a COPY stage: nothing to update
}

// Copy the carried state of the stateful stage `old_stage` of plan `olds` into whichever segment of the new plan now holds that
// stage (`new_stage`).  State is per STAGE in meaning -- a FIR's past input, a biquad's TDF-II state, a resampler's past input
// and phase -- and every kernel family stores it in the same form, so it survives any regrouping of the
// stages into segments.
static int32_t carry_stage_state(pb_chain *c, const std::vector<Segment> &olds, int old_stage, int new_stage)
{
    const Segment *so = nullptr;
    Segment *sn = nullptr;
    for (const auto &s : olds)
        if (s.fir_stage == old_stage || s.bq_stage == old_stage || s.rs_stage == old_stage) so = &s;
    for (auto &s : c->segs)
        if (s.fir_stage == new_stage || s.bq_stage == new_stage || s.rs_stage == new_stage) sn = &s;
    if (!so || !sn) return PB_OK;  // a stage without carried state (copy, gain)
    const size_t el = c->elem;
    if (so->fir_stage == old_stage && so->Hf > 0)
        PB_CUDA(cudaMemcpy(sn->d_xhist[sn->pp], so->d_xhist[so->pp], el * (size_t)so->Hf * c->C, cudaMemcpyDeviceToDevice));
    if (so->bq_stage == old_stage)
        PB_CUDA(cudaMemcpy(sn->d_state[sn->pp], so->d_state[so->pp], sizeof(double) * (size_t)c->C * 2, cudaMemcpyDeviceToDevice));
    if (so->rs_stage == old_stage) {
        if (so->Hr > 0)
            PB_CUDA(cudaMemcpy(sn->d_yhist[sn->pp], so->d_yhist[so->pp], el * (size_t)so->Hr * c->C, cudaMemcpyDeviceToDevice));
        sn->acc = so->acc;
    }
    return PB_OK;
}

// InsertProcessor on a fused run (pipe.go:297-333, run.go:134-169): the run is re-planned -- the stage list is cut into fused
// segments again, tables and workspaces are rebuilt -- and every stage that was already there keeps its carried state; the new
// stage starts from zero state, like a freshly allocated Processor.  Takes effect at the next process call (the reference
// applies the edit as a mutation between buffers, pipe.go:302).
extern "C" int32_t pb_chain_insert_stage(pb_chain *c, int32_t pos, const pb_stage_desc *st)
{
    if (!c || !st) return fail(PB_ERR_INVALID, "pb_chain_insert_stage: NULL argument");
    if (pos < 0 || pos > (int)c->stages.size()) return fail(PB_ERR_INVALID, "pb_chain_insert_stage: position %d of %zu", pos, c->stages.size());
    if (c->slots_busy) return fail(PB_ERR_STATE, "pb_chain_insert_stage while submitted batches are in flight");
    int32_t r = validate_stage(*st, pos);
    if (r != PB_OK) return r;
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaDeviceSynchronize());   // edits land between buffers: drain what is in flight first
    StageCopy ns;
    ns.d = *st;
    if (st->taps && st->n_taps > 0) ns.taps.assign(st->taps, st->taps + st->n_taps);
    ns.d.taps = nullptr;
    std::vector<StageCopy> old_stages = c->stages;
    std::vector<Segment> olds = std::move(c->segs);
    void *old_mid[2] = {c->d_mid[0], c->d_mid[1]};
    const double old_rate = c->out_sample_rate;
    c->stages.insert(c->stages.begin() + pos, ns);
    c->d_mid[0] = c->d_mid[1] = nullptr;
    c->tmaps.clear();
    plan_segments(c);
    auto rollback = [&](int32_t code) {
        char keep[512];
        snprintf(keep, sizeof keep, "%s", tls_error_buf());
        for (auto &s : c->segs) free_segment(s);
        for (int i = 0; i < 2; i++)
            if (c->d_mid[i]) cudaFree(c->d_mid[i]);
        c->segs = std::move(olds);
        c->stages = old_stages;
        c->d_mid[0] = old_mid[0];
        c->d_mid[1] = old_mid[1];
        c->out_sample_rate = old_rate;
        snprintf(tls_error_buf(), 512, "%s", keep);
        return code;
    };
    double rate = c->sample_rate;
    for (auto &s : c->segs) {
        r = build_segment(c, s);
        if (r != PB_OK) return rollback(r);
        if (s.rs_stage >= 0) rate = rate * s.up / s.down;
    }
    if (c->segs.size() > 1)
        for (int i = 0; i < 2; i++)
            if (cudaMalloc(&c->d_mid[i], c->elem * (size_t)c->max_frames * c->C) != cudaSuccess)
                return rollback(fail(PB_ERR_NOMEM, "pb_chain_insert_stage: intermediate buffers"));
    for (int i = 0; i < (int)old_stages.size(); i++) {
        r = carry_stage_state(c, olds, i, i < pos ? i : i + 1);
        if (r != PB_OK) return rollback(r);
    }
    PB_CUDA(cudaDeviceSynchronize());
    c->out_sample_rate = rate;
    for (auto &s : olds) free_segment(s);
    for (int i = 0; i < 2; i++)
        if (old_mid[i]) cudaFree(old_mid[i]);
    return PB_OK;
}

extern "C" int32_t pb_chain_meter_read(pb_chain *c, double *peak, double *sumsq, int64_t *frames)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_meter_read: NULL chain");
    if (!c->d_meter) return fail(PB_ERR_STATE, "pb_chain_meter_read: chain was created without PB_CHAIN_METER");
    DeviceGuard dg(c->device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaDeviceSynchronize());
    if (peak) PB_CUDA(cudaMemcpy(peak, c->d_meter, sizeof(double) * (size_t)c->C, cudaMemcpyDeviceToHost));
    if (sumsq) PB_CUDA(cudaMemcpy(sumsq, c->d_meter + c->C, sizeof(double) * (size_t)c->C, cudaMemcpyDeviceToHost));
    if (frames) *frames = c->meter_frames;
    return PB_OK;
}

extern "C" int32_t pb_chain_last_path(const pb_chain *c, int32_t *path, int64_t *kernel_launches)
{
    if (!c) return fail(PB_ERR_INVALID, "pb_chain_last_path: NULL chain");
    if (path) *path = c->last_path;
    if (kernel_launches) *kernel_launches = c->launches;
    return PB_OK;
}

// ---- memory helpers ------------------------------------------------------------

extern "C" int32_t pb_device_count(int32_t *count)
{
    if (!count) return fail(PB_ERR_INVALID, "pb_device_count: NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(PB_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return PB_OK;
}

extern "C" int32_t pb_device_alloc(int32_t device, int64_t bytes, void **ptr)
{
    if (!ptr || bytes < 0) return fail(PB_ERR_INVALID, "pb_device_alloc: bad argument");
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaMalloc(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return PB_OK;
}

extern "C" int32_t pb_device_free(int32_t device, void *ptr)
{
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaFree(ptr));
    return PB_OK;
}

extern "C" int32_t pb_host_alloc_pinned(int64_t bytes, void **ptr)
{
    if (!ptr || bytes < 0) return fail(PB_ERR_INVALID, "pb_host_alloc_pinned: bad argument");
    PB_CUDA(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return PB_OK;
}

extern "C" int32_t pb_host_free_pinned(void *ptr)
{
    PB_CUDA(cudaFreeHost(ptr));
    return PB_OK;
}

extern "C" int32_t pb_memcpy_h2d(int32_t device, void *dst, const void *src, int64_t bytes)
{
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
    return PB_OK;
}

extern "C" int32_t pb_memcpy_d2h(int32_t device, void *dst, const void *src, int64_t bytes)
{
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return PB_OK;
}

extern "C" int32_t pb_device_synchronize(int32_t device)
{
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaDeviceSynchronize());
    return PB_OK;
}

extern "C" int32_t pb_ipc_export(int32_t device, void *ptr, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    if (!ptr || !handle) return fail(PB_ERR_INVALID, "pb_ipc_export: NULL");
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return PB_OK;
}

// cudaIpcGetMemHandle names the ALLOCATION a pointer lies in, and cudaIpcOpenMemHandle returns that allocation's base: a buffer
// carved out of a larger block (any caching allocator) sits at this offset from it.
extern "C" int32_t pb_ipc_offset(int32_t device, void *ptr, int64_t *offset)
{
    if (!ptr || !offset) return fail(PB_ERR_INVALID, "pb_ipc_offset: NULL");
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*PFN_range)(CUdeviceptr *, size_t *, CUdeviceptr);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(PB_ERR_CUDA, "cuMemGetAddressRange is unavailable");
    const CUresult r = ((PFN_range)fn)(&base, &size, (CUdeviceptr)(uintptr_t)ptr);
    if (r != CUDA_SUCCESS) return fail(PB_ERR_CUDA, "cuMemGetAddressRange failed (%d)", (int)r);
    *offset = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
    return PB_OK;
}

extern "C" int32_t pb_ipc_open(int32_t device, const uint8_t handle[64], void **ptr)
{
    if (!ptr || !handle) return fail(PB_ERR_INVALID, "pb_ipc_open: NULL");
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    PB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PB_OK;
}

extern "C" int32_t pb_ipc_close(int32_t device, void *ptr)
{
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    PB_CUDA(cudaIpcCloseMemHandle(ptr));
    return PB_OK;
}
