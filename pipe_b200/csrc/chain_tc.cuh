// chain_tc.cuh -- K2: the headline Processor run [gain, FIR<=257, biquad, resample 147/160 x16] on
// tcgen05 / TMEM / TMA.  Replaces the ProcessFunc walk of Processor.execute (reference pipe.go:438)
// for that run when the call is aligned to 160-frame tiles; everything else goes through K1
// (chain_tile.cuh), with which it shares every piece of carried state.
//
// Two contractions per tile of 128 channels x 160 frames, both with channels on M (TMEM lane = channel):
//
//  MMA1 (FIR)        D1[ch x 176] = X^T[ch x 432] * Toeplitz[432 x 176]
//      N = 176 columns (15 left-context frames recomputed for the resampler + 160 frames + 1 pad),
//      K = 432 input frames in 27 chunks of 16.
//  MMA2 (resampler)  D2[ch x 147] = Y^T[ch x 176] * R[176 x 147]
//      R is the 147/160 polyphase matrix of the tile (every output has 16 taps, so R is a narrow band):
//      five slices of 32 outputs, each touching 4 (the last: 3) chunks of 16 rows -- 19 (slice, chunk) blocks.
//
// Precision (measured with tools/tc_probe.cu): the tensor core adds into its fp32 accumulator with
// truncation, which costs ~1e-6 over 27..81 steps.  So the operands are split on FIXED grids:
//   x*2^11 = x0 + x1,  h*2^sh = h0 + h1 + h2,  x0 and h0 integer-valued fp16 (|.| <= 2048), the others
//   fp16 remainders (h needs the third piece: the remainder of a tap has ABSOLUTE precision 2^-13 of the
//   grid, which summed over 257 taps was 5.7e-7 of the peak).  x0*h0 goes to accumulator E: every product
//   and every partial sum is an integer below 2^24, so E is EXACT.  x0*h1 + x0*h2 + x1*h0 + x1*h1 go to
//   accumulator X, 2^-9 of E in magnitude, whose truncation is negligible.  Modelled FIR error 6e-8.
//   The resampler uses the same scheme with two pieces each (y*2^11 = y0 + y1, r*2^sh2 = r0 + r1; only 16
//   taps per output): E2 = y0*r0 exact, X2 = y0*r1 + y1*r0 + y1*r1.  Modelled error 2.2e-7 of the peak, the
//   same as the f32 FFMA chain it replaces.
//
// B operand of MMA1: the Toeplitz matrix is never materialised.  A K-major 8x8 core matrix depends only on
// (column block - row block); with the two K-blocks of an instruction stored swapped in A, the
// descriptor strides LBO = SBO = 128 B make 75 core matrices (9.6 KB per piece) serve all 27 x 22
// positions.  B operand of MMA2: the 19 blocks [32 outputs x 16 rows] of R, 1 KB per piece.
//
// Biquad between the two contractions: channel-per-lane, double precision.  The recursion is sequential in
// time, so one thread runs FOUR zero-state chains (rows 0-47, 48-95, 96-135, 136-175) interleaved for
// instruction-level parallelism and never waits for the state coming from the previous tile: y is linear in
// that state, so its contribution is a rank-2 correction per chain, carried through the resampler on the host
// (rc = R * (A^k)_row0) and added to the outputs when D2 is drained.  The state itself comes from the same
// decoupled look-back as K1 (aggregate Z published right after the drain, inclusive state after the look-back),
// which now overlaps MMA2 instead of stalling the recursion.
//
// Shared-memory staging: D1 is drained to an f32 tile so that TMEM is free for the next tile's MMA1 at once.
// The biquad threads overwrite that tile IN PLACE with the fp16 pieces of y in the MN-major UMMA layout (one
// 16-row chunk = 8 KB of f32 = 2 pieces x 4 KB of fp16; each warp's f32 values live inside the footprint of its
// own 32 channels' pieces, a half-chunk of 8 rows is read completely before it is overwritten).  Core-matrix
// stride 144 B instead of 128 B keeps the 2-byte scatter stores conflict-free.
//
// Roles (608 threads): warp 0 TMA producer, warp 1 MMA1 issuer, warps 2-5 / 6-9 two converter groups on
// alternate chunks (f32 tile -> x0/x1 in the MN-major UMMA layout), warps 10-13 biquad (drain rows [0,80),
// Z, recursion, pieces, look-back, chain states), warps 14-17 output (drain rows [80,176), then per slice:
// D2 -> registers, correction, coalesced stores, meter), warp 18 MMA2 issuer.  Tiles follow a static
// time-major schedule over a persistent grid.
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
constexpr int kTcThreads = 736;     // 23 warps: TMA, MMA1, 2x4 converters, 2x4 biquad, 4 output, MMA2
constexpr int kRawStages = 4, kCvtStages = 3;
constexpr int kTcDrainB = 64;       // rows [0,64) drained / Z-summed by biquad role A, [64,128) by role B, [128,176) by the output warps
constexpr int kTcRowChunks = kTcN / 16;  // 11 chunks of 16 rows of y (K of MMA2)
constexpr int kRsSlices = 5;        // slices of 32 outputs; slice s reads row chunks [2s, 2s+4) (the last one 3)
constexpr int kRsN = 32;
constexpr int kRsPairs = 19;        // (slice, chunk) blocks of R
__host__ __device__ constexpr int rs_chain_of_slice(int s) { return s < 2 ? 0 : (s == 2 ? 1 : 2); }  // the two biquad chains (j, j+1) whose rows a slice reads
constexpr int kBqChains = 4;        // zero-state recursion chains per tile, in half-chunks of 8 rows:
constexpr int kBqHc0 = 0, kBqHc1 = 6, kBqHc2 = 12, kBqHc3 = 17;   //   first half-chunk of each chain
constexpr int kBqLen0 = 6, kBqLen1 = 6, kBqLen2 = 5, kBqLen3 = 5; //   half-chunks per chain (rows 0,48,96,136)

struct TcTables {  // tables in global memory, copied to shared at kernel start
    static constexpr int kT = kTcCores * 64;           // T0 T1 T2: 75*64 halfs each
    static constexpr int kHalfs = 3 * kT;
    static constexpr int kB2 = kRsPairs * 2 * 512;      // then R blocks [pair][piece][nb 4][kb 2][8][8] halfs
    static constexpr int kBytes = (kHalfs + kB2) * 2;
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
    int dbg;             // development switches (PB_TC_DBG): bit0 skip MMA1, bit1 skip MMA2, bit2 skip the conversion, bit4 skip the TMA loads
    int C, n_tiles, n_cg;
    int hist_rows;       // rows of xhist == FIR taps - 1 (<= 256)
    unsigned epoch;
    float scale_in;      // g_load * 2^11 (applied to frames of this call)
    float scale_hist;    // 2^11          (history frames are already gain-scaled)
    float inv_scale_in;  // 2^-11: turns scaled input back into xhist_next values
    float descale_fir;   // g_fir / (2^11 * 2^sh)
    float descale_rs;    // g_out / (2^11 * 2^sh2)
    float yh_scale;      // 2^11: carried y history rows -> the fixed-point grid of the pieces
    double b0, b1, b2, a1, a2;
    double g_bq;         // gain after the biquad (y = v * g_bq)
    double ysc;          // g_bq * 2^11 * 2^13 (y -> fixed point with 13 fractional bits below the 2^11 grid)
    double AL[4];        // A^160 (look-back step)
    double AP48[4], AP40[4], AP24[4], AP39[4];  // chain-to-chain / snapshot transitions
    float Wf[kTcFrames][2];   // W[k] = A^k B: zero-state end state Z = sum_r W[159-r] * fir[r]
    float rc[kTcOut][4];      // output correction: out[m] += rc[m][0..1] . q(j) + rc[m][2..3] . q(j+1), j = kRsChainOfSlice[m/32]
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__device__ __forceinline__ long long clk() { return clock64(); }
enum { kProfProdWait = 0, kProfMmaWaitTmem, kProfMmaWaitCvt, kProfMmaIssue, kProfCvtWaitRaw, kProfCvtWaitCvt, kProfCvtWork,
       kProfBqWaitTmem, kProfBqDrain, kProfBqZ, kProfBqMain, kProfBqLookback, kProfTotal, kProfOutWait, kProfOutMain,
       kProfMma2Wait, kProfBWaitTmem, kProfBDrain, kProfBZ, kProfBMain, kProfBLookback, kProfCount = 24 };

// shared memory map (bytes)
constexpr int kRawStageBytes = 16 * kTcCh * 4;                  // 8 KB: 4 sub-tiles of 16 rows x 128 B
constexpr int kCvtStageBytes = 2 * 16 * kTcCh * 2;              // 8 KB: x0 then x1
constexpr int kOffRaw = 0;
constexpr int kOffCvt = kOffRaw + kRawStages * kRawStageBytes;
constexpr int kOffTab = kOffCvt + kCvtStages * kCvtStageBytes;
constexpr int kTabBytes = TcTables::kBytes;                     // 28800 + 38912
// staging tile: 11 chunks of 16 rows; a chunk is [piece 2][kb 2][mb 16] core matrices of 8 rows x 16 B at a
// 144 B stride (see the header comment); the f32 FIR values of warp e live in the 576 B ranges of its own 4 mb
constexpr int kMbStride = 144, kKbStride = 16 * kMbStride, kPieceBytes = 2 * kKbStride, kChunkBytes = 2 * kPieceBytes;
constexpr int kOffStage = ((kOffTab + kTabBytes + 127) / 128) * 128;
constexpr int kStageBytes = kTcRowChunks * kChunkBytes;         // 101376
constexpr int kOffSstate = kOffStage + kStageBytes;             // [4][8][32] float chain states; aliases zpart [4][32] double2
constexpr int kOffBar = kOffSstate + 4 * 8 * 32 * 4;
constexpr int kNumBars = 2 * kRawStages + 2 * kCvtStages + 2 + 4 + 1 + 1 + 2 + 2 + 4 + 4;
constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemSlot + 16;
static_assert(kBqLen0 == kBqLen2 + 1 && kBqLen1 == kBqLen3 + 1 && kBqLen2 == kBqLen3, "chain schedule");
static_assert(kSmemBytes <= 227 * 1024, "K2 shared memory budget");
static_assert(kOffStage % 16 == 0 && kChunkBytes % 16 == 0, "UMMA descriptor alignment");

// TMEM columns: D1 = E [0,176) + X [192,368); D2 double-buffered slices E2/X2 of 32 columns from 368
constexpr uint32_t kColE = 0, kColX = 192, kColD2 = 368;

// byte offset of the f32 staging value of row r16 (0..15) of a chunk, relative to the thread's base
__host__ __device__ constexpr int f32_off(int r16) { return (r16 >> 3) * kKbStride + ((r16 >> 2) & 1) * kPieceBytes + (r16 & 3) * 128; }

}  // namespace tc

struct BqCoef {
    double b0, b1, b2, na1, na2, ysc;
};

#ifdef __CUDACC__
// One half-chunk (8 rows) of NCH independent TDF-II recursions for one channel, straight-line so that the chains
// interleave:  v = b0 x + s1;  s1' = (b1 x + s2) - a1 v;  s2' = b2 x - a2 v;  y*2^11 = v * ysc is split into its fp16
// pieces, which overwrite the f32 FIR values of the half-chunk in place (all lanes read them first).
// g[j]: half-chunk of chain j.  HIST (tile 0 only, chain 0): rows < 15 are the carried y history, not recursion output.
template <int NCH, bool HIST, bool CAP>
__device__ __forceinline__ void bq_step(unsigned char *stf, unsigned char *stp, const int (&g)[NCH], const BqCoef &kc,
                                        double *cs1, double *cs2, double &e3_1, double &e3_2, float &vmax,
                                        const float *yh, int C, float yh_scale)
{
    using namespace tc;
    bool oflow = false;
    float xf[NCH][8];
    unsigned char *dst[NCH];
#pragma unroll
    for (int j = 0; j < NCH; j++) {
        const int off = (g[j] >> 1) * kChunkBytes + (g[j] & 1) * kKbStride;
        const unsigned char *b = stf + off;
        dst[j] = stp + off;
#pragma unroll
        for (int rr = 0; rr < 8; rr++) xf[j][rr] = *reinterpret_cast<const float *>(b + (rr >> 2) * kPieceBytes + (rr & 3) * 128);
    }
    __syncwarp();  // every lane has read the half-chunks that the piece stores below overwrite
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            // f32 -> f64 and f64 -> fixed point by bit manipulation: the F2F conversions that involve a 64-bit type
            // cost ~40 cycles per warp on this part (measured: they, not the DFMAs, bounded the recursion)
            const unsigned xu = __float_as_uint(xf[j][rr]);
            const unsigned xa = xu & 0x7fffffffu;
            const int xhi = (int)((xu & 0x80000000u) | (xa < 0x00800000u ? 0u : (xa >> 3) + 0x38000000u));  // zero / denormal -> 0
            const double x = __hiloint2double(xhi, (int)(xu << 29));
            const double tt = fma(kc.b1, x, cs2[j]);
            const double p2 = kc.b2 * x;
            const double v = fma(kc.b0, x, cs1[j]);
            double n1 = fma(kc.na1, v, tt);
            double n2 = fma(kc.na2, v, p2);
            // K = rint(y * 2^11 * 2^13) from the low word of v * ysc13 + 1.5 * 2^52;  y * 2^11 = i + k / 2^13
            const double rk = fma(v, kc.ysc, 6755399441055744.0);
            int K = __double2loint(rk);
            oflow |= (unsigned)(__double2hiint(rk) + 1 - 0x43380000) > 1u;  // |y * 2^24| >= 2^31: K has wrapped
            if (HIST && j == 0) {
                const int row = 8 * g[0] + rr;
                const bool use = row < kTcHr;
                const float yv = use ? yh[(size_t)row * C] : 0.f;
                K = use ? __float2int_rn(yv * yh_scale * 8192.f) : K;
                n1 = use ? cs1[0] : n1;
                n2 = use ? cs2[0] : n2;
            }
            cs1[j] = n1;
            cs2[j] = n2;
            const int ii = (K + 4096) >> 13, kk = K - (ii << 13);
            const float ra = __int_as_float(0x4B400000 + ii) - 12582912.f;   // exact int -> float for |.| < 2^22
            const float fk = __int_as_float(0x4B400000 + kk) - 12582912.f;
            const __half h0 = __float2half_rn(ra);            // above 2048 the fp16 grid is coarser than 1:
            const __half h1 = __float2half_rn(fmaf(fk, 1.f / 8192.f, ra - __half2float(h0)));  // the remainder is taken from what h0 really holds
            vmax = fmaxf(vmax, fabsf(ra));
            *reinterpret_cast<__half *>(dst[j] + rr * 16) = h0;
            *reinterpret_cast<__half *>(dst[j] + rr * 16 + kPieceBytes) = h1;
        }
        if (CAP && rr == 6) {  // last chain of the call after row 8g+6: on chain 3's last half-chunk that is row 174
            e3_1 = cs1[NCH - 1];
            e3_2 = cs2[NCH - 1];
        }
    }
    if (oflow) vmax = 1e30f;
}
#endif

__global__ void __launch_bounds__(kTcThreads, 1) chain_tc_kernel(const __grid_constant__ TcParams p)
{
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *raw = smem + kOffRaw;
    unsigned char *cvt = smem + kOffCvt;
    __half *tab = reinterpret_cast<__half *>(smem + kOffTab);
    unsigned char *stage = smem + kOffStage;
    float *sstate = reinterpret_cast<float *>(smem + kOffSstate);
    double2 *zpart = reinterpret_cast<double2 *>(smem + kOffSstate);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *raw_full = bars, *raw_empty = bars + kRawStages;
    uint64_t *cvt_full = bars + 2 * kRawStages, *cvt_empty = cvt_full + kCvtStages;
    uint64_t *tmem_full = cvt_empty + kCvtStages, *tmem_empty = tmem_full + 1;
    uint64_t *zb_ready = tmem_empty + 1;            // [4]  biquad warp B and output warp e -> biquad warps e: rows staged, Z parts in zpart
    uint64_t *y_ready = zb_ready + 4;               //      biquad warps -> MMA2: the y pieces of the tile are in the staging tile
    uint64_t *stage_free = y_ready + 1;             //      MMA2 (commit) -> drain: the staging tile has been consumed
    uint64_t *d2_full = stage_free + 1;             // [2]  MMA2 (commit) -> output warps
    uint64_t *d2_empty = d2_full + 2;               // [2]  output warps -> MMA2
    uint64_t *state_ready = d2_empty + 2;           // [4]  biquad warps e (A and B) -> output warp e: chain states in sstate
    uint64_t *q2_ready = state_ready + 4;           // [4]  biquad warp A -> B: state at the start of chain 2 in the q2 slot
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
        mbar_init(tmem_empty, 12);
        for (int i = 0; i < 4; i++) {
            mbar_init(&zb_ready[i], 2);
            mbar_init(&state_ready[i], 2);
            mbar_init(&q2_ready[i], 1);
        }
        mbar_init(y_ready, 8);
        mbar_init(stage_free, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 4);
        }
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
                    if (p.dbg & 16) {  // development: no loads
                        mbar_arrive(&raw_full[s]);
                        if (++s == kRawStages) { s = 0; ph ^= 1; }
                        continue;
                    }
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
        // ================================ MMA1 issuer =================================
        if (lane == 0) {
            const uint32_t t0 = smem_u32(tab), t1 = t0 + TcTables::kT * 2, t2 = t1 + TcTables::kT * 2;
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
                    // Only the band of the Toeplitz matrix is multiplied: chunk q (frames f0-272+16q ..+15) reaches output
                    // columns [16q-257, 16q+14], i.e. 8-column blocks [2q-33, 2q+1] clipped to [0, 21] and widened to an even
                    // count (N % 16 == 0).  Chunk 0 runs at full width with accumulate = 0: it zeroes the accumulators.
                    int nb0 = 0, nbl = kTcN / 8;
                    if (q > 0) {
                        int lo = 2 * q - 33 > 0 ? 2 * q - 33 : 0, hi = 2 * q + 1 < kTcN / 8 - 1 ? 2 * q + 1 : kTcN / 8 - 1;
                        if ((hi - lo + 1) & 1) lo--;  // odd only when lo > 0
                        nb0 = lo;
                        nbl = hi - lo + 1;
                    }
                    const uint32_t toff = (52 - 2 * q + nb0) * 128;
                    const uint64_t b0 = make_desc(t0 + toff, 128, 128), b1 = make_desc(t1 + toff, 128, 128);
                    const uint64_t b2 = make_desc(t2 + toff, 128, 128);
                    const uint32_t acc = q > 0, idesc = make_idesc(8 * nbl);
                    const uint32_t dE = tmem_base + kColE + 8 * nb0, dX = tmem_base + kColX + 8 * nb0;
                    if (!(p.dbg & 1)) {
                        umma(dE, a0, b0, idesc, acc);   // exact: integers < 2^24
                        umma(dX, a0, b1, idesc, acc);
                        umma(dX, a0, b2, idesc, 1);
                        umma(dX, a1, b0, idesc, 1);
                        umma(dX, a1, b1, idesc, 1);
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
            for (int kb = 0; kb < ((p.dbg & 4) ? 0 : 2); kb++) {
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
    } else if (warp < 18) {
        // ================================ biquad warps ================================
        // Two warps per TMEM lane quadrant e (channels cg*128 + 32e + lane): role A (warps 10-13) drains rows [0,64)
        // and runs chains 0,1 (rows 0-95) and the look-back; role B (warps 14-17) drains rows [64,128) and runs chains
        // 2,3 (rows 96-175), then continues A's state chain (q2 arrives through a 16 B slot of the staging tile).
        const int e = warp & 3;
        const bool roleB = warp >= 14;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        unsigned char *stf = stage + e * 4 * kMbStride + lane * 4;                        // f32 view: + chunk + f32_off(r16)
        unsigned char *stp = stage + (e * 4 + (lane >> 3)) * kMbStride + (lane & 7) * 2;  // piece view: + chunk + piece + kb + fr*16
        // never touched by the f32 view, the pieces or the MMA: the 16 B pad of core-matrix column 4e+3 of block `lane`
        double2 *q2slot = reinterpret_cast<double2 *>(stage + (lane >> 2) * kChunkBytes + ((lane >> 1) & 1) * kPieceBytes +
                                                      (lane & 1) * kKbStride + (e * 4 + 3) * kMbStride + 128);
        const BqCoef kc = {p.b0, p.b1, p.b2, -p.a1, -p.a2, p.ysc};
        float vmax = 0.f;
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
            if (it > 0) mbar_wait(stage_free, par ^ 1);  // MMA2 has consumed the previous tile's pieces
            const long long k1 = clk();
            e_w += k1 - k0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            // ---- drain 64 rows: FIR = (E + X) * descale into the staging tile; the look-back aggregate
            //      Z = sum_r W[159-r] fir[r] is accumulated on the way (16-term float partial sums folded in double)
            double Z0 = 0.0, Z1 = 0.0;
            const bool chained = !first && !last;
            const int cbase = roleB ? kTcDrainB : 0;
#pragma unroll
            for (int cc = 0; cc < kTcDrainB; cc += 16) {
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + cbase + cc, re);
                tmem_ld16(tmem_base + lane_base + kColX + cbase + cc, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
                unsigned char *dstc = stf + ((cbase + cc) >> 4) * kChunkBytes;
                const float(*wf)[2] = p.Wf + (kTcFrames - 1 - (cbase + cc));
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float v = (__uint_as_float(re[i]) + __uint_as_float(rx[i])) * p.descale_fir;
                    *reinterpret_cast<float *>(dstc + f32_off(i)) = v;
                    p0 = fmaf(wf[-i][0], v, p0);
                    p1 = fmaf(wf[-i][1], v, p1);
                }
                Z0 += (double)p0;
                Z1 += (double)p1;
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            if (roleB) {
                zpart[e * 32 + lane] = make_double2(Z0, Z1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&zb_ready[e]);
            }
            const long long k2 = clk();
            e_d += k2 - k1;
            mbar_wait(&zb_ready[e], par);  // all rows of this quadrant are staged; both other Z parts are in zpart
            const size_t slot = (size_t)grp * p.n_tiles + t;
            if (!roleB && chained) {
                const double2 zb = zpart[e * 32 + lane], zo = zpart[128 + e * 32 + lane];
                Z0 += zb.x + zo.x;
                Z1 += zb.y + zo.y;
                p.lb_agg[slot * 64 + lane * 2] = Z0;
                p.lb_agg[slot * 64 + lane * 2 + 1] = Z1;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
            }
            const long long k3 = clk();
            e_z += k3 - k2;

            // ---- two interleaved zero-state recursions in double over half-chunks of 8 rows; y*2^11 is split into
            //      its fp16 pieces over the FIR values it was computed from
            double cs1[2] = {0.0, 0.0}, cs2[2] = {0.0, 0.0};
            double s159_1 = 0.0, s159_2 = 0.0, e3_1 = 0.0, e3_2 = 0.0;
            const float *yh = p.yhist + c;
            if (!roleB) {
                if (first) {
                    cs1[0] = p.bq_state[2 * c];
                    cs2[0] = p.bq_state[2 * c + 1];
                }
#pragma unroll 1
                for (int i = 0; i < kBqLen0; i++) {
                    const int g2[2] = {kBqHc0 + i, kBqHc1 + i};
                    if (first && i < 2) bq_step<2, true, false>(stf, stp, g2, kc, cs1, cs2, e3_1, e3_2, vmax, yh, p.C, p.yh_scale);
                    else bq_step<2, false, false>(stf, stp, g2, kc, cs1, cs2, e3_1, e3_2, vmax, yh, p.C, p.yh_scale);
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < kBqLen2; i++) {
                    const int g2[2] = {kBqHc2 + i, kBqHc3 + i};
                    bq_step<2, false, true>(stf, stp, g2, kc, cs1, cs2, e3_1, e3_2, vmax, yh, p.C, p.yh_scale);
                    if (i == 2) {  // chain 3 after row 159: the state the next tile's row 0 starts from
                        s159_1 = cs1[1];
                        s159_2 = cs2[1];
                    }
                }
                if (last) {
                    // carried y history rows 160..174, zero-state part (the state response is added below), read back
                    // from the pieces just written: y * 2^11 = y0 + y1
                    for (int r = 0; r < kTcHr; r++) {
                        const unsigned char *d = stp + 10 * kChunkBytes + (r >> 3) * kKbStride + (r & 7) * 16;
                        const float v = __half2float(*reinterpret_cast<const __half *>(d)) + __half2float(*reinterpret_cast<const __half *>(d + kPieceBytes));
                        p.yhist_next[(size_t)r * p.C + c] = v / p.yh_scale;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
            __syncwarp();
            if (lane == 0) mbar_arrive(y_ready);
            const long long k4 = clk();
            e_m += k4 - k3;

            if (!roleB) {
                // ---- incoming state: decoupled look-back (same protocol and arrays as K1, 32-channel groups)
                double s1 = 0.0, s2 = 0.0;
                if (!first) {
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
                    // eight at a time so that only one L2 round trip per batch is exposed
                    for (int i0 = first_inc - 1; i0 >= 0; i0 -= 8) {
                        double a0[8], a1[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int i = i0 - u;
                            const size_t sa = (size_t)grp * p.n_tiles + (base - (i >= 0 ? i : 0));
                            a0[u] = ld_cg(p.lb_agg + sa * 64 + lane * 2);
                            a1[u] = ld_cg(p.lb_agg + sa * 64 + lane * 2 + 1);
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (i0 - u >= 0) {
                                mat2_apply(p.AL, s1, s2);
                                s1 += a0[u];
                                s2 += a1[u];
                            }
                    }
                }
                // ---- true state at the start of chains 0..2 (chain 0 of tile 0 carried it itself: its entry is zero)
                double u1 = s1, u2 = s2;
                mat2_apply(p.AP48, u1, u2);
                const double q11 = u1 + cs1[0], q12 = u2 + cs2[0];
                u1 = q11;
                u2 = q12;
                mat2_apply(p.AP48, u1, u2);
                *q2slot = make_double2(u1 + cs1[1], u2 + cs2[1]);
                sstate[(e * 8 + 0) * 32 + lane] = (float)s1;
                sstate[(e * 8 + 1) * 32 + lane] = (float)s2;
                sstate[(e * 8 + 2) * 32 + lane] = (float)q11;
                sstate[(e * 8 + 3) * 32 + lane] = (float)q12;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&q2_ready[e]);
                    mbar_arrive(&state_ready[e]);
                }
            } else {
                mbar_wait(&q2_ready[e], par);
                const double2 q2 = *q2slot;
                double q31 = q2.x, q32 = q2.y;
                mat2_apply(p.AP40, q31, q32);
                q31 += cs1[0];
                q32 += cs2[0];
                if (!last) {
                    // inclusive state after row 159 (== before the next tile's row 0)
                    double I0 = q31, I1 = q32;
                    mat2_apply(p.AP24, I0, I1);
                    I0 += s159_1;
                    I1 += s159_2;
                    p.lb_inc[slot * 64 + lane * 2] = I0;
                    p.lb_inc[slot * 64 + lane * 2 + 1] = I1;
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
                }
                sstate[(e * 8 + 4) * 32 + lane] = (float)q2.x;
                sstate[(e * 8 + 5) * 32 + lane] = (float)q2.y;
                sstate[(e * 8 + 6) * 32 + lane] = (float)q31;
                sstate[(e * 8 + 7) * 32 + lane] = (float)q32;
                __syncwarp();
                if (lane == 0) mbar_arrive(&state_ready[e]);
                if (last) {
                    // carried state: after row 174; carried y history: rows 160..174 with the state response added
                    double E0 = q31, E1 = q32;
                    mat2_apply(p.AP39, E0, E1);
                    p.bq_state_next[2 * c] = E0 + e3_1;
                    p.bq_state_next[2 * c + 1] = E1 + e3_2;
                    double u1 = q31, u2 = q32;
                    mat2_apply(p.AP24, u1, u2);
                    const double A[4] = {-p.a1, 1.0, -p.a2, 0.0};
                    for (int r = 0; r < kTcHr; r++) {
                        float *hp = p.yhist_next + (size_t)r * p.C + c;
                        *hp = (float)((double)*hp + u1 * p.g_bq);
                        mat2_apply(A, u1, u2);
                    }
                }
            }
            e_l += clk() - k4;
        }
        if (vmax > 60000.f) atomicExch(p.err_flag, 2);
        if (p.prof && (warp == 10 || warp == 14) && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount + (roleB ? kProfBWaitTmem - kProfBqWaitTmem : 0);
            pr[kProfBqWaitTmem] = e_w;
            pr[kProfBqDrain] = e_d;
            pr[kProfBqZ] = e_z;
            pr[kProfBqMain] = e_m;
            pr[kProfBqLookback] = e_l;
        }
    } else if (warp < 22) {
        // ================================ output warps ================================
        const int e = warp & 3;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        unsigned char *stf = stage + e * 4 * kMbStride + lane * 4;
        long long r_w = 0, r_m = 0;
        int it = 0;
        unsigned nsl = 0;  // running slice number: D2 buffer nsl & 1, phase (nsl >> 1) & 1
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const int c = cg * kTcCh + e * 32 + lane;
            const uint32_t par = it & 1;
            mbar_wait(tmem_full, par);
            if (it > 0) mbar_wait(stage_free, par ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            // ---- drain rows [128,176) with this warp's part of Z (rows [128,160)) accumulated on the way
            double Z0 = 0.0, Z1 = 0.0;
#pragma unroll
            for (int c0 = 2 * kTcDrainB; c0 < kTcN; c0 += 16) {
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + c0, re);
                tmem_ld16(tmem_base + lane_base + kColX + c0, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float v = (__uint_as_float(re[i]) + __uint_as_float(rx[i])) * p.descale_fir;
                    *reinterpret_cast<float *>(stf + (c0 >> 4) * kChunkBytes + f32_off(i)) = v;
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
            zpart[128 + e * 32 + lane] = make_double2(Z0, Z1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&zb_ready[e]);

            // ---- per slice: D2 -> registers, add the state response of the chains, store
            float *outp = p.out + (size_t)t * kTcOut * p.C + c;
            float sv[8];
            float m_peak = 0.f;
            double m_sumsq = 0.0;
            const bool meter = p.meter_peak != nullptr;
            const long long k5 = clk();
#pragma unroll 1
            for (int s = 0; s < kRsSlices; s++, nsl++) {
                const uint32_t b = nsl & 1u;
                const long long k6 = clk();
                mbar_wait(&d2_full[b], (nsl >> 1) & 1u);
                r_w += clk() - k6;
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (s == 0) {
                    mbar_wait(&state_ready[e], par);
#pragma unroll
                    for (int q = 0; q < 8; q++) sv[q] = sstate[(e * 8 + q) * 32 + lane];
                }
                // the rows of this slice belong to chains (j, j+1): four state components matter
                const float q0 = s < 2 ? sv[0] : (s == 2 ? sv[2] : sv[4]), q1 = s < 2 ? sv[1] : (s == 2 ? sv[3] : sv[5]);
                const float q2 = s < 2 ? sv[2] : (s == 2 ? sv[4] : sv[6]), q3 = s < 2 ? sv[3] : (s == 2 ? sv[5] : sv[7]);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t re[16], rx[16];
                    tmem_ld16(tmem_base + lane_base + kColD2 + 64 * b + 16 * h, re);
                    tmem_ld16(tmem_base + lane_base + kColD2 + 64 * b + 32 + 16 * h, rx);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d2_empty[b]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int m = kRsN * s + 16 * h + i;
                        if (m < kTcOut) {
                            const float corr = fmaf(p.rc[m][0], q0, fmaf(p.rc[m][1], q1, fmaf(p.rc[m][2], q2, p.rc[m][3] * q3)));
                            const float o = fmaf(__uint_as_float(re[i]) + __uint_as_float(rx[i]), p.descale_rs, corr);
                            outp[(size_t)m * p.C] = o;
                            if (meter) {
                                m_peak = fmaxf(m_peak, fabsf(o));
                                m_sumsq += (double)o * (double)o;
                            }
                        }
                    }
                }
            }
            r_m += clk() - k5;
            if (meter) {
                atomic_max_nonneg(p.meter_peak + c, (double)m_peak);
                atomicAdd(p.meter_sumsq + c, m_sumsq);
            }
        }
        if (p.prof && warp == 18 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfOutWait] = r_w;
            pr[kProfOutMain] = r_m;
        }
    } else {
        // ================================ MMA2 issuer =================================
        if (lane == 0) {
            const uint32_t st0 = smem_u32(stage);
            const uint32_t b2 = smem_u32(tab) + TcTables::kHalfs * 2;
            constexpr uint32_t idesc_rs = make_idesc(kRsN);
            unsigned nsl = 0;
            int it = 0;
            long long w_y = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const long long c0 = clk();
                mbar_wait(y_ready, it & 1);
                w_y += clk() - c0;
                asm volatile("tcgen05.fence::after_thread_sync;");
                int pair = 0;
                for (int s = 0; s < kRsSlices; s++, nsl++) {
                    const uint32_t b = nsl & 1u;
                    mbar_wait(&d2_empty[b], ((nsl >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t dE = tmem_base + kColD2 + 64 * b, dX = dE + 32;
                    const int nch = (s == kRsSlices - 1) ? 3 : 4;
                    for (int k = 0; k < nch; k++, pair++) {
                        const uint32_t a_base = st0 + (2 * s + k) * kChunkBytes;
                        const uint64_t a0 = make_desc(a_base, kKbStride, kMbStride), a1 = make_desc(a_base + kPieceBytes, kKbStride, kMbStride);
                        const uint64_t r0 = make_desc(b2 + pair * 2048, 128, 256), r1 = make_desc(b2 + pair * 2048 + 1024, 128, 256);
                        const uint32_t acc = k > 0;
                        if (!(p.dbg & 2)) {
                            umma(dE, a0, r0, idesc_rs, acc);   // exact: integers < 2^24
                            umma(dX, a0, r1, idesc_rs, acc);
                            umma(dX, a1, r0, idesc_rs, 1);
                            umma(dX, a1, r1, idesc_rs, 1);
                        }
                    }
                    umma_commit(&d2_full[b]);
                }
                umma_commit(stage_free);
            }
            if (p.prof) p.prof[blockIdx.x * kProfCount + kProfMma2Wait] = w_y;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

#endif  // __CUDACC__

}  // namespace pb
