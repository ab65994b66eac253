// chain_tc.cuh -- K2: the headline Processor run [gain, FIR<=257, biquad, resample 147/160 x16] on
// tcgen05 / TMEM / TMA.  Replaces the ProcessFunc walk of Processor.execute (reference pipe.go:438)
// for that run when the call is aligned to 160-frame tiles; everything else goes through K1
// (chain_tile.cuh), with which it shares every piece of carried state.
//
// FIR as a tensor-core contraction:  D[ch x frames] = X^T[ch x window] * Toeplitz[window x frames]
//   M = 128 channels (TMEM lane = channel), N = 176 columns (15 left-context frames recomputed for
//   the resampler + 160 frames + 1 pad), K = 432 input frames in 27 chunks of 16.
//
// Precision (measured with tools/tc_probe.cu): the tensor core adds into its fp32 accumulator with
// truncation, which costs ~1e-6 over 27..81 steps.  So the operands are split on FIXED grids:
//   x*2^11 = x0 + x1,  h*2^sh = h0 + h1 + h2,  x0 and h0 integer-valued fp16 (|.| <= 2048), the others
//   fp16 remainders (h needs the third piece: the remainder of a tap has ABSOLUTE precision 2^-13 of the
//   grid, which summed over 257 taps was 5.7e-7 of the peak).  x0*h0 goes to accumulator E: every product
//   and every partial sum is an integer below 2^24, so E is EXACT.  x0*h1 + x0*h2 + x1*h0 + x1*h1 go to
//   accumulator X, 2^-9 of E in magnitude, whose truncation is negligible.  Modelled FIR error 6e-8.
//
// B operand: the Toeplitz matrix is never materialised.  A K-major 8x8 core matrix depends only on
// (column block - row block); with the two K-blocks of an instruction stored swapped in A, the
// descriptor strides LBO = SBO = 128 B make 75 core matrices (9.6 KB per piece) serve all 27 x 22
// positions.
//
// Roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9 two converter groups
// on alternate chunks (f32 tile -> x0/x1 in the MN-major UMMA layout), warps 10-13 biquad (drain half
// of TMEM to a shared staging tile, look-back, double-precision recursion per channel, y written back in
// place), warps 14-17 resampler (other half of the drain, then the statically unrolled 147/160 polyphase
// streaming 16-row blocks behind the biquad warps, coalesced stores).  One warp per scheduler per role
// exposed every latency (measured: 71 k cycles per tile, 50 k of them in a single-warp epilogue), hence
// the split.  Tiles follow a static time-major schedule over a persistent grid.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include <utility>

#include "chain_tile.cuh"

namespace pb {

constexpr int kTcCh = 128;          // channels per tile (UMMA M)
constexpr int kTcFrames = 160;      // input frames per tile == resampler period (down)
constexpr int kTcUp = 147, kTcP = 16, kTcHr = kTcP - 1;
constexpr int kTcOut = 147;         // output frames per tile
constexpr int kTcN = 176;           // accumulator columns: 15 + 160 + 1 pad
constexpr int kTcChunks = 27;       // K window = 432 frames = [f0-272, f0+160)
constexpr int kTcWin = kTcChunks * 16;
constexpr int kTcLead = 272;        // frames of the window before the tile start
constexpr int kTcMaxTaps = 257;
constexpr int kTcCores = 75;        // Toeplitz core matrices per piece
constexpr int kTcThreads = 576;     // 18 warps: TMA, MMA, 2x4 converters, 4 biquad, 4 resampler
constexpr int kRawStages = 8, kCvtStages = 4;
constexpr int kTcSplit = 80;        // rows [0,80) drained / Z-summed by the biquad warps, [80,176) by the resampler warps
constexpr int kTcBlocks = 11;       // 16-row hand-off blocks between the biquad and resampler warps

struct TcTables {  // tables in global memory, copied to shared at kernel start
    static constexpr int kT = kTcCores * 64;           // T0 T1 T2: 75*64 halfs each
    static constexpr int kHalfs = 3 * kT;
    static constexpr int kSeq = kTcN * 16;              // then the resampler coefficients, [row][slot] floats (+1 spare row)
    static constexpr int kBytes = kHalfs * 2 + kSeq * 4;
};

struct TcParams {
    CUtensorMap tm_in;    // [n_frames][C] f32, box 32 ch x 16 frames, SWIZZLE_128B
    CUtensorMap tm_hist;  // xhist [256][C] f32 (already gain-scaled), same box
    float *out;
    const __half *tables;
    const float *yhist;
    float *yhist_next;
    float *xhist_next;
    const double *bq_state;
    double *bq_state_next;
    double *lb_agg, *lb_inc;
    unsigned *lb_status;
    double *meter_peak, *meter_sumsq;
    int *err_flag;
    long long *prof;     // optional per-CTA cycle counters (PB_TC_PROF=1), nullptr otherwise
    int dbg;             // development switches (PB_TC_DBG): bit0 skip the MMAs, bit1 skip the TMA loads
    int C, n_tiles, n_cg;
    int hist_rows;       // rows of xhist == FIR taps - 1 (<= 256)
    unsigned epoch;
    float scale_in;      // g_load * 2^11 (applied to frames of this call)
    float scale_hist;    // 2^11          (history frames are already gain-scaled)
    float inv_scale_in;  // 2^-11: turns scaled input back into xhist_next values
    float descale_fir;   // g_fir / (2^11 * 2^sh)
    float g_out;
    double b0, b1, b2, a1, a2, g_bq;
    double AL[4];        // A^160
    float Wf[kTcFrames][2];  // W[k] = A^k B: zero-state end state Z = sum_r W[159-r] * fir[r]
    unsigned rs_emit[6];  // bit r: row r completes an output (the oldest in-flight one)
};

#ifdef __CUDACC__

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"  // suspend-time hint: park the warp in
        "@p bra DONE;\n"                                               // hardware instead of spinning (a spinning
        "bra WAIT_LOOP;\n"                                             // warp steals issue slots from its scheduler)
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // sm_100 descriptor version; SWIZZLE_NONE, base offset 0
    return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc(int n)
{
    // c=F32 (1<<4), a=b=F16 (0), a MN-major (1<<15), b K-major, N>>3 at bit 17, M>>4 at bit 24
    return (1u << 4) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ long long clk() { return clock64(); }
enum { kProfProdWait = 0, kProfMmaWaitTmem, kProfMmaWaitCvt, kProfMmaIssue, kProfCvtWaitRaw, kProfCvtWaitCvt, kProfCvtWork,
       kProfEpiWaitTmem, kProfEpiDrain, kProfEpiZ, kProfEpiLookback, kProfEpiMain, kProfTotal, kProfRsWait, kProfRsMain, kProfCount = 16 };

// shared memory map (bytes)
constexpr int kRawStageBytes = 16 * kTcCh * 4;                  // 8 KB: 4 sub-tiles of 16 rows x 128 B
constexpr int kCvtStageBytes = 2 * 16 * kTcCh * 2;              // 8 KB: x0 then x1
constexpr int kOffRaw = 0;
constexpr int kOffCvt = kOffRaw + kRawStages * kRawStageBytes;  // 32 KB
constexpr int kOffTab = kOffCvt + kCvtStages * kCvtStageBytes;  // 64 KB
constexpr int kTabBytes = TcTables::kBytes;                     // 28800 + 11200
constexpr int kOffStage = ((kOffTab + kTabBytes + 127) / 128) * 128;
constexpr int kStageBytes = kTcN * kTcCh * 4;                   // 90112: FIR output tile [176][128] f32
constexpr int kOffZpart = kOffStage + kStageBytes;              // [4][32] double2: resampler warps' half of Z
constexpr int kOffBar = kOffZpart + 4 * 32 * 16;
constexpr int kNumBars = 2 * kRawStages + 2 * kCvtStages + 2 + 4 + 4 * kTcBlocks + 4;
constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemSlot + 16;
static_assert(kSmemBytes <= 227 * 1024, "K2 shared memory budget");

}  // namespace tc

struct BqCoef {
    double b0, b1, b2, na1, na2, gbq;
};

// ROWS rows of the TDF-II recursion for one channel, in place in the staging column.
// All loads and conversions are issued up front; the loop-carried path is two DP operations per row:
//   v = b0 x + s1;  s1' = (b1 x + s2) - a1 v;  s2' = b2 x - a2 v
template <int ROWS>
__device__ __forceinline__ void bq_block(float *__restrict__ col, const BqCoef &k, double &s1, double &s2)
{
    float xf[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) xf[i] = col[i * kTcCh];
    double xd[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) xd[i] = (double)xf[i];
#pragma unroll
    for (int i = 0; i < ROWS; i++) {
        const double t = fma(k.b1, xd[i], s2);
        const double p2 = k.b2 * xd[i];
        const double v = fma(k.b0, xd[i], s1);
        s1 = fma(k.na1, v, t);
        s2 = fma(k.na2, v, p2);
        col[i * kTcCh] = (float)(v * k.gbq);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) chain_tc_kernel(const __grid_constant__ TcParams p)
{
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *raw = smem + kOffRaw;
    unsigned char *cvt = smem + kOffCvt;
    __half *tab = reinterpret_cast<__half *>(smem + kOffTab);
    const float *rs_seq = reinterpret_cast<const float *>(smem + kOffTab + TcTables::kHalfs * 2);
    float *stage = reinterpret_cast<float *>(smem + kOffStage);
    double2 *zpart = reinterpret_cast<double2 *>(smem + kOffZpart);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *raw_full = bars, *raw_empty = bars + kRawStages;
    uint64_t *cvt_full = bars + 2 * kRawStages, *cvt_empty = cvt_full + kCvtStages;
    uint64_t *tmem_full = cvt_empty + kCvtStages, *tmem_empty = tmem_full + 1;
    uint64_t *zb_ready = tmem_empty + 1;            // [4]      resampler warp e -> biquad warp e: Z half is in zpart
    uint64_t *yblk = zb_ready + 4;                  // [4][11]  biquad warp e -> resampler warp e: y block k is in the staging tile
    uint64_t *wg2_done = yblk + 4 * kTcBlocks;      // [4]      resampler warp e finished the tile
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffTmemSlot);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_tiles = p.n_tiles * p.n_cg;

    // ---- one-time setup ------------------------------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int i = 0; i < kRawStages; i++) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], 4);
        }
        for (int i = 0; i < kCvtStages; i++) {
            mbar_init(&cvt_full[i], 4);
            mbar_init(&cvt_empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);
        for (int i = 0; i < 4; i++) {
            mbar_init(&zb_ready[i], 1);
            mbar_init(&wg2_done[i], 1);
        }
        for (int i = 0; i < 4 * kTcBlocks; i++) mbar_init(&yblk[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tables);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = tid; i < kTabBytes / 16; i += kTcThreads) dst[i] = src[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColE = 0, kColX = 192;  // TMEM columns: E [0,176), X [192,368)

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int s = 0, ph = 0;
            long long pw = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
                const int f0 = t * kTcFrames, ch0 = cg * kTcCh;
                for (int q = 0; q < kTcChunks; q++) {
                    const long long c0 = clk();
                    mbar_wait(&raw_empty[s], ph ^ 1);
                    pw += clk() - c0;
                    mbar_expect_tx(&raw_full[s], kRawStageBytes);
                    const int fr = f0 - kTcLead + 16 * q;  // first frame of the chunk, call-relative
                    // chunks never straddle frame 0 (kTcLead and tile starts are multiples of 16)
                    const CUtensorMap *map = (fr < 0) ? &p.tm_hist : &p.tm_in;
                    const int row = (fr < 0) ? fr + p.hist_rows : fr;  // history map holds frames [-hist_rows, 0); rows < 0 are zero-filled
                    unsigned char *dst = raw + s * kRawStageBytes;
#pragma unroll
                    for (int g = 0; g < 4; g++) tma_load_2d(dst + g * 2048, map, ch0 + 32 * g, row, &raw_full[s]);
                    if (++s == kRawStages) { s = 0; ph ^= 1; }
                }
            }
            if (p.prof) p.prof[blockIdx.x * kProfCount + kProfProdWait] = pw;
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t t0 = smem_u32(tab), t1 = t0 + TcTables::kT * 2, t2 = t1 + TcTables::kT * 2;
            constexpr uint32_t idesc_main = make_idesc(kTcN);
            int s = 0, ph = 0, tph = 0;
            long long w_t = 0, w_c = 0, w_i = 0;
            const long long kstart = clk();
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                long long c0 = clk();
                mbar_wait(tmem_empty, tph ^ 1);  // both halves of the previous tile have been drained
                w_t += clk() - c0;
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (int q = 0; q < kTcChunks; q++) {
                    c0 = clk();
                    mbar_wait(&cvt_full[s], ph);
                    const long long c1 = clk();
                    w_c += c1 - c0;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_base = smem_u32(cvt + s * kCvtStageBytes);
                    const uint64_t a0 = make_desc(a_base, 2048, 128), a1 = make_desc(a_base + 4096, 2048, 128);
                    const uint32_t toff = (52 - 2 * q) * 128;
                    const uint64_t b0 = make_desc(t0 + toff, 128, 128), b1 = make_desc(t1 + toff, 128, 128);
                    const uint64_t b2 = make_desc(t2 + toff, 128, 128);
                    const uint32_t acc = q > 0;
                    if (!(p.dbg & 1)) {
                        umma(tmem_base + kColE, a0, b0, idesc_main, acc);   // exact: integers < 2^24
                        umma(tmem_base + kColX, a0, b1, idesc_main, acc);
                        umma(tmem_base + kColX, a0, b2, idesc_main, 1);
                        umma(tmem_base + kColX, a1, b0, idesc_main, 1);
                        umma(tmem_base + kColX, a1, b1, idesc_main, 1);
                    }
                    umma_commit(&cvt_empty[s]);  // frees the A stage when these MMAs have read it
                    w_i += clk() - c1;
                    if (++s == kCvtStages) { s = 0; ph ^= 1; }
                }
                umma_commit(tmem_full);
                tph ^= 1;
            }
            if (p.prof) {
                long long *pr = p.prof + blockIdx.x * kProfCount;
                pr[kProfMmaWaitTmem] = w_t;
                pr[kProfMmaWaitCvt] = w_c;
                pr[kProfMmaIssue] = w_i;
                pr[kProfTotal] = clk() - kstart;
            }
        }
    } else if (warp < 10) {
        // ================================ converters ==================================
        // two groups of 4 warps take alternate chunks; in a group warp cw owns channel sub-tile cw
        // (32 channels); lane -> (fr_i = lane % 8, mbq = lane / 8)
        const int grp = (warp - 2) >> 2, cw = (warp - 2) & 3;
        const int fr_i = lane & 7, mbq = lane >> 3, mb = cw * 4 + mbq;
        float vmax = 0.f;
        long long w_r = 0, w_c = 0, w_w = 0;
        const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int n_chunks = my_tiles * kTcChunks;
        for (int g = grp; g < n_chunks; g += 2) {  // g: this CTA's running chunk number
            const int it = g / kTcChunks, q = g - it * kTcChunks;
            const int tile = blockIdx.x + it * gridDim.x;
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const int f0 = t * kTcFrames;
            const bool last = (t == p.n_tiles - 1);
            const int rs = g % kRawStages, rph = (g / kRawStages) & 1;
            const int cs = g % kCvtStages, cph = (g / kCvtStages) & 1;
            const long long c0 = clk();
            mbar_wait(&raw_full[rs], rph);
            const long long c1 = clk();
            mbar_wait(&cvt_empty[cs], cph ^ 1);
            const long long c2 = clk();
            w_r += c1 - c0;
            w_c += c2 - c1;
            const bool hist = (f0 - kTcLead + 16 * q) < 0;
            const float sc = hist ? p.scale_hist : p.scale_in;
            const unsigned char *src = raw + rs * kRawStageBytes + cw * 2048;
            unsigned char *dst = cvt + cs * kCvtStageBytes;
#pragma unroll
            for (int kb = 0; kb < 2; kb++) {
                const int row = 8 * kb + fr_i;
                // SWIZZLE_128B: 16 B chunk c of a row lives at chunk position c ^ (row % 8)
                const float4 va = *reinterpret_cast<const float4 *>(src + row * 128 + (((2 * mbq) ^ fr_i) << 4));
                const float4 vb = *reinterpret_cast<const float4 *>(src + row * 128 + (((2 * mbq + 1) ^ fr_i) << 4));
                float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                __half2 hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float a = v[2 * i] * sc, b = v[2 * i + 1] * sc;
                    // round to nearest integer on the FMA pipe (exact for |a| < 2^22; larger values trip the range check)
                    const float ra = (a + 12582912.f) - 12582912.f, rb = (b + 12582912.f) - 12582912.f;
                    hi[i] = __floats2half2_rn(ra, rb);
                    lo[i] = __floats2half2_rn(a - ra, b - rb);
                    vmax = fmaxf(vmax, fmaxf(fabsf(a), fabsf(b)));
                    v[2 * i] = a;
                    v[2 * i + 1] = b;
                }
                const int off = (1 - kb) * 2048 + mb * 128 + fr_i * 16;  // K-blocks swapped (Toeplitz trick)
                *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<uint4 *>(hi);
                *reinterpret_cast<uint4 *>(dst + 4096 + off) = *reinterpret_cast<uint4 *>(lo);
                const int hrow = 16 * q + row - 176 - (256 - p.hist_rows);
                if (last && hrow >= 0) {
                    // carried FIR input history: frames [n-hist_rows, n) in gain-scaled units (K1's convention)
                    float *hp = p.xhist_next + (size_t)hrow * p.C + cg * kTcCh + mb * 8;
                    const float is = p.inv_scale_in;
                    *reinterpret_cast<float4 *>(hp) = make_float4(v[0] * is, v[1] * is, v[2] * is, v[3] * is);
                    *reinterpret_cast<float4 *>(hp + 4) = make_float4(v[4] * is, v[5] * is, v[6] * is, v[7] * is);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&cvt_full[cs]);
                mbar_arrive(&raw_empty[rs]);
            }
            w_w += clk() - c2;
        }
        if (vmax > 60000.f) atomicExch(p.err_flag, 2);
        if (p.prof && warp == 2 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfCvtWaitRaw] = w_r;
            pr[kProfCvtWaitCvt] = w_c;
            pr[kProfCvtWork] = w_w;
        }
    } else if (warp < 14) {
        // ================================ biquad warps ================================
        // warp e owns TMEM lanes [32e, 32e+32) == channels cg*128 + 32e + lane: an independent chain.
        const int e = warp & 3;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        float *st = stage + e * 32 + lane;  // stage[row][128]: this thread's column
        long long e_w = 0, e_d = 0, e_z = 0, e_l = 0, e_m = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const bool first = (t == 0), last = (t == p.n_tiles - 1);
            const int c = cg * kTcCh + e * 32 + lane;
            const int grp = cg * 4 + e;  // 32-channel look-back group, same indexing as K1
            const uint32_t par = it & 1;
            const long long k0 = clk();
            mbar_wait(tmem_full, par);
            if (it > 0) mbar_wait(&wg2_done[e], par ^ 1);  // the resampler warp is done with the previous tile's rows
            const long long k1 = clk();
            e_w += k1 - k0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            // ---- drain rows [0,80): FIR = (E + X) * descale into the staging tile; the look-back aggregate
            //      Z = sum_r W[159-r] fir[r] is accumulated on the way (16-term float partial sums folded in double)
            double Z0 = 0.0, Z1 = 0.0;
            const bool chained = !first && !last;
#pragma unroll
            for (int c0 = 0; c0 < kTcSplit; c0 += 16) {
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + c0, re);
                tmem_ld16(tmem_base + lane_base + kColX + c0, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float v = (__uint_as_float(re[i]) + __uint_as_float(rx[i])) * p.descale_fir;
                    st[(c0 + i) * kTcCh] = v;
                    p0 = fmaf(p.Wf[kTcFrames - 1 - (c0 + i)][0], v, p0);
                    p1 = fmaf(p.Wf[kTcFrames - 1 - (c0 + i)][1], v, p1);
                }
                Z0 += (double)p0;
                Z1 += (double)p1;
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            const long long k2 = clk();
            e_d += k2 - k1;
            mbar_wait(&zb_ready[e], par);  // also orders this warp after the other half of the drain
            if (chained) {
                const double2 zb = zpart[e * 32 + lane];
                Z0 += zb.x;
                Z1 += zb.y;
            }
            const long long k3 = clk();
            e_z += k3 - k2;

            // ---- biquad state chain (same protocol and arrays as K1, 32-channel groups)
            const size_t slot = (size_t)grp * p.n_tiles + t;
            if (chained) {
                p.lb_agg[slot * 64 + lane * 2] = Z0;
                p.lb_agg[slot * 64 + lane * 2 + 1] = Z1;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
            }
            double s1, s2;
            if (first) {
                s1 = p.bq_state[2 * c];
                s2 = p.bq_state[2 * c + 1];
            } else {
                const int base = t - 1, j = base - lane;
                int first_inc = 0;
                for (unsigned spins = 0;; spins++) {
                    unsigned stt = kLbInc;
                    if (j >= 0) {
                        stt = ld_acquire_u32(p.lb_status + (size_t)grp * p.n_tiles + j);
                        stt = ((stt >> 2) == p.epoch) ? (stt & 3u) : kLbNone;
                    }
                    const unsigned ready = __ballot_sync(0xffffffffu, stt != kLbNone);
                    const unsigned inc = __ballot_sync(0xffffffffu, stt == kLbInc);
                    if (inc) {
                        first_inc = __ffs(inc) - 1;
                        const unsigned need = (first_inc == 0) ? 0u : (0xffffffffu >> (32 - first_inc));
                        if ((ready & need) == need) break;
                    }
                    if (spins > (1u << 24)) {
                        if (lane == 0) atomicExch(p.err_flag, 1);
                        first_inc = -1;
                        break;
                    }
                    __nanosleep(20);
                }
                __syncwarp();
                const size_t s_inc = (size_t)grp * p.n_tiles + (first_inc < 0 ? 0 : base - first_inc);
                s1 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2);
                s2 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2 + 1);
                // Horner over the aggregates between that inclusive state and this tile; payloads are fetched
                // sixteen at a time so that only one L2 round trip per batch is exposed
                for (int i0 = first_inc - 1; i0 >= 0; i0 -= 16) {
                    double a0[16], a1[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int i = i0 - u;
                        const size_t sa = (size_t)grp * p.n_tiles + (base - (i >= 0 ? i : 0));
                        a0[u] = ld_cg(p.lb_agg + sa * 64 + lane * 2);
                        a1[u] = ld_cg(p.lb_agg + sa * 64 + lane * 2 + 1);
                    }
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        if (i0 - u >= 0) {
                            mat2_apply(p.AL, s1, s2);
                            s1 += a0[u];
                            s2 += a1[u];
                        }
                }
            }
            if (chained) {
                double I0 = s1, I1 = s2;
                mat2_apply(p.AL, I0, I1);
                I0 += Z0;
                I1 += Z1;
                p.lb_inc[slot * 64 + lane * 2] = I0;
                p.lb_inc[slot * 64 + lane * 2 + 1] = I1;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
            }
            const long long k4 = clk();
            e_l += k4 - k3;

            // ---- recursion in double, y written back over the FIR value; a block of 16 rows at a time is
            //      handed to the resampler warp
            const BqCoef kc = {p.b0, p.b1, p.b2, -p.a1, -p.a2, p.g_bq};
            int k_start = 0;
            if (first) {
                // tile 0: the 15 left-context rows come from the carried history, the recursion starts at row 15
                for (int r = 0; r < kTcHr; r++) st[r * kTcCh] = p.yhist[(size_t)r * p.C + c];
                bq_block<1>(st + kTcHr * kTcCh, kc, s1, s2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&yblk[e * kTcBlocks]);
                k_start = 1;
            }
#pragma unroll 1
            for (int k = k_start; k < kTcBlocks - 1; k++) {
                bq_block<16>(st + 16 * k * kTcCh, kc, s1, s2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&yblk[e * kTcBlocks + k]);
            }
            if (first && !last) {
                // tile 0 publishes its inclusive state (after row 159, frame 144) from the recursion itself
                p.lb_inc[slot * 64 + lane * 2] = s1;
                p.lb_inc[slot * 64 + lane * 2 + 1] = s2;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
            }
            bq_block<15>(st + 160 * kTcCh, kc, s1, s2);  // rows 160..174; column 175 is padding
            __syncwarp();
            if (lane == 0) mbar_arrive(&yblk[e * kTcBlocks + kTcBlocks - 1]);
            if (last)
                for (int j = 0; j < kTcHr; j++) p.yhist_next[(size_t)j * p.C + c] = st[(kTcFrames + j) * kTcCh];
            if (last) {
                p.bq_state_next[2 * c] = s1;
                p.bq_state_next[2 * c + 1] = s2;
            }
            e_m += clk() - k4;
        }
        if (p.prof && warp == 10 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfEpiWaitTmem] = e_w;
            pr[kProfEpiDrain] = e_d;
            pr[kProfEpiZ] = e_z;
            pr[kProfEpiLookback] = e_l;
            pr[kProfEpiMain] = e_m;
        }
    } else {
        // ================================ resampler warps =============================
        const int e = warp & 3;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        float *st = stage + e * 32 + lane;
        long long r_w = 0, r_m = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const bool first = (t == 0), last = (t == p.n_tiles - 1);
            const int c = cg * kTcCh + e * 32 + lane;
            const uint32_t par = it & 1;
            mbar_wait(tmem_full, par);
            asm volatile("tcgen05.fence::after_thread_sync;");
            // ---- drain rows [80,176) with this warp's half of Z (rows [80,160)) accumulated on the way
            double Z0 = 0.0, Z1 = 0.0;
#pragma unroll
            for (int c0 = kTcSplit; c0 < kTcN; c0 += 16) {
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + c0, re);
                tmem_ld16(tmem_base + lane_base + kColX + c0, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float v = (__uint_as_float(re[i]) + __uint_as_float(rx[i])) * p.descale_fir;
                    st[(c0 + i) * kTcCh] = v;
                    if (c0 + i < kTcFrames) {
                        p0 = fmaf(p.Wf[kTcFrames - 1 - (c0 + i)][0], v, p0);
                        p1 = fmaf(p.Wf[kTcFrames - 1 - (c0 + i)][1], v, p1);
                    }
                }
                Z0 += (double)p0;
                Z1 += (double)p1;
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            zpart[e * 32 + lane] = make_double2(Z0, Z1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&zb_ready[e]);

            // ---- 147/160 polyphase, input-driven, as a rolled loop over the y rows the biquad warp releases.
            //      acc[0..15] are the in-flight outputs ordered by completion; every row adds its tap to each of
            //      them, and on "emit" rows the oldest one is finished and the others move up one slot.  The host
            //      lays the coefficients out per row in exactly that slot order (rs_seq), so the loop body is the
            //      same 16 FFMAs for every row: compact code instead of 60 KB of unrolled, I-cache-missing SASS.
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = 0.f;
            float *outp = p.out + (size_t)t * kTcOut * p.C + c;
            float m_peak = 0.f;
            double m_sumsq = 0.0;
            const bool meter = p.meter_peak != nullptr;
            const float g_out = p.g_out;
            const long long k5 = clk();
#pragma unroll 1
            for (int blk = 0; blk < kTcBlocks; blk++) {
                const long long k6 = clk();
                mbar_wait(&yblk[e * kTcBlocks + blk], par);
                r_w += clk() - k6;
                const int nrows = (blk == kTcBlocks - 1) ? 15 : 16;
                const unsigned emask = (p.rs_emit[blk >> 1] >> ((blk & 1) * 16)) & ((1u << nrows) - 1u);
                const float *col = st + 16 * blk * kTcCh;
                const float4 *cq = reinterpret_cast<const float4 *>(rs_seq + 16 * blk * 16);
                // Branch-free rows (a branch per row would pin every row's loads behind its reconvergence point):
                // t[j] = cf[j]*y + acc[j]; on an emit row t[0] is the finished output and the others move up a slot.
                // Rows past nrows have zero coefficients in the table and no emit bit, so they are no-ops.
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float y = col[(i < nrows ? i : nrows - 1) * kTcCh];
                    const float4 c0 = cq[4 * i], c1 = cq[4 * i + 1], c2 = cq[4 * i + 2], c3 = cq[4 * i + 3];
                    const float cf[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w,
                                          c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
                    const bool emit = (emask >> i) & 1u;
                    float tt[17];
#pragma unroll
                    for (int j = 0; j < 16; j++) tt[j] = fmaf(cf[j], y, acc[j]);
                    tt[16] = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[j] = emit ? tt[j + 1] : tt[j];
                    if (emit) {
                        const float o = tt[0] * g_out;
                        *outp = o;
                        if (meter) {
                            m_peak = fmaxf(m_peak, fabsf(o));
                            m_sumsq += (double)o * (double)o;
                        }
                    }
                    outp += emit ? p.C : 0;
                }
            }
            r_m += clk() - k5;
            if (meter) {
                atomic_max_nonneg(p.meter_peak + c, (double)m_peak);
                atomicAdd(p.meter_sumsq + c, m_sumsq);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&wg2_done[e]);
        }
        if (p.prof && warp == 14 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfRsWait] = r_w;
            pr[kProfRsMain] = r_m;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

#endif  // __CUDACC__

}  // namespace pb
