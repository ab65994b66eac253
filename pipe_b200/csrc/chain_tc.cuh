// chain_tc.cuh -- K2: the headline Processor run [gain, FIR<=257, biquad, resample 147/160 x16] on
// tcgen05 / TMEM / TMA.  Replaces the ProcessFunc walk of Processor.execute (reference pipe.go:438)
// for that run when the call is aligned to 160-frame tiles; everything else goes through K1
// (chain_tile.cuh), with which it shares every piece of carried state.
//
// Two contractions per tile of 128 channels x 160 frames, both with channels on M (TMEM lane = channel):
//
//  MMA1 (FIR)                 D1[ch x 176] = X^T[ch x 432] * Toeplitz[432 x 176]
//      N = 176 columns (15 left-context frames recomputed for the resampler + 160 frames + 1 pad),
//      K = 432 input frames in 27 chunks of 16; only the band of the Toeplitz matrix is multiplied.
//  MMA2 (biquad + resampler)  D2[ch x 147] = F^T[ch x 176] * P[176 x 147]
//      F is the FIR output.  The biquad is linear, so inside one block of 16 rows its zero-state response is a
//      lower-triangular Toeplitz matrix G16; P = blockdiag(G16) * R folds it into the 147/160 polyphase matrix R
//      of the tile and keeps R's block structure: five slices of 32 outputs, each touching 4 (the last: 3) chunks
//      of 16 rows -- 19 (slice, chunk) blocks.  What the blocks do not see is the biquad state at each block
//      start: y is linear in it, so its contribution is a rank-2 correction per block, carried through the
//      resampler on the host (rc = R * (A^k)_row0) and added when D2 is drained.
//
// Why the biquad is NOT a per-row recursion here (it was): the recursion has to run in double (an f32 TDF-II with
// poles near z = 1 measured 1e-5), and on this part a running tcgen05.mma stream throttles DFMA about 7x (measured:
// 3.0 -> 21.9 cycles per warp instruction, tools/mma_probe.cu), conversions between f32 and f64 cost ~40 cycles each,
// so 176 rows x 6 DFMA per channel bounded the kernel at ~20 k cycles per tile.  Folded into P, the only double
// arithmetic left is the block-state recursion s' = A^16 s + Z_b: 4 DFMA per 16 rows.
//
// Precision (measured with tools/tc_probe.cu): the tensor core adds into its fp32 accumulator with
// truncation, which costs ~1e-6 over 27..81 steps.  So the operands are split on FIXED grids:
//   x*2^11 = x0 + x1,  h*2^sh = h0 + h1 + h2,  x0 and h0 integer-valued fp16 (|.| <= 2048), the others
//   fp16 remainders (h needs the third piece: the remainder of a tap has ABSOLUTE precision 2^-13 of the
//   grid, which summed over 257 taps was 5.7e-7 of the peak).  x0*h0 goes to accumulator E: every product
//   and every partial sum is an integer below 2^24, so E is EXACT.  x0*h1 + x0*h2 + x1*h go to
//   accumulator X, 2^-9 of E in magnitude, whose truncation is negligible (x1 is multiplied by the tap rounded once:
//   four passes).  Modelled FIR error 1.3e-7.
//   MMA2 uses the same scheme with two pieces each (f*2^11 = f0 + f1, P*2^sh2 = p0 + p1; at most 32 rows per
//   output): E2 = f0*p0 exact, X2 = f0*p1 + f1*p0 + f1*p1.  Modelled error 2.2e-7 of the peak.  The grids are FIXED, so
//   the noise floor is 2^-24 of full scale (|g x| = 1), not of each channel's own level.
//
// B operand of MMA1: the Toeplitz matrix is never materialised.  A K-major 8x8 core matrix depends only on
// (column block - row block); with the two K-blocks of an instruction stored swapped in A, the
// descriptor strides LBO = SBO = 128 B make 75 core matrices (9.6 KB per piece) serve all 27 x 22
// positions.  B operand of MMA2: the 19 blocks [32 outputs x 16 rows] of P, 1 KB per piece, the two pieces
// adjacent so that one N = 64 instruction multiplies f0 by [p0 | p1] into [E2 | X2].
//
// Roles (864 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA1 issuer, warps 2-5 / 6-9 two converter
// groups on alternate chunks (f32 tile -> x0/x1 in the MN-major UMMA layout), warps 10-13 / 14-17 two FIR drain groups on
// the even / odd 16-column blocks of D1 (one warp per TMEM lane quadrant in each: D1 -> f pieces in the MN-major UMMA
// layout for MMA2, block by block as MMA1's last chunks complete them; the first group then runs the block-state recursion
// and the look-back), warps 18-21 / 22-25 two output groups on the two halves of every slice (D2 -> registers, free the
// buffer, block-state correction, coalesced stores, meter), warp 26 MMA2 issuer.  Tiles follow a static time-major schedule.
// The drain warps hand the 11 block states to the output warps through 24 spare TMEM columns.  Inside a role group only the
// first warp polls mbarriers, the others wait on a named barrier; the issuing warps run converged and elect one lane per
// tcgen05 instruction.  MMA1 writes the first touch of every 16 columns with accumulate = 0 and waits for the previous
// tile's drain per column block, so it overlaps the tail of that drain.
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
constexpr int kTcThreads = 864;     // 27 warps: TMA, MMA1, 2x4 converters, 2x4 drain, 2x4 output, MMA2
constexpr int kRawStages = 5, kCvtStages = 3;
constexpr int kTcRing = 8;          // chunks of the MMA2 A operand kept in shared memory (block b lives in slot b % 8)
constexpr int kTcBlocks = kTcN / 16;     // 11 blocks of 16 rows: biquad blocks == K chunks of MMA2
constexpr int kTcFirstDone = 17;    // column block b of D1 is complete after chunk b + 17 (the last two after chunk 26)
constexpr int kRsSlices = 5;        // slices of 32 outputs; slice s reads row blocks [2s, 2s+4) (the last one 3)
constexpr int kRsN = 32;
constexpr int kRsPairs = 19;        // (slice, block) blocks of P; entry 19 is (0, 0) for tile 0 (rows 0..14 are y history)

struct TcTables {  // tables in global memory, copied to shared at kernel start
    static constexpr int kT = kTcCores * 64;           // T0 T1 T2 (three-piece split) and T3 (the tap rounded once): 75*64 halfs each
    static constexpr int kHalfs = 4 * kT;
    static constexpr int kB2 = (kRsPairs + 1) * 2 * 512;  // then P blocks [pair][piece][nb 4][kb 2][8][8] halfs
    static constexpr int kBytes = (kHalfs + kB2) * 2;
};

struct TcParams {
    CUtensorMap tm_in;    // [n_frames][C] f32, box 32 ch x 16 frames, SWIZZLE_128B
    CUtensorMap tm_hist;  // xhist [256][C] f32 (already gain-scaled), same box
    float *out;
    const __half *tables;
    const float *rc;      // [kTcRcRows][8]
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
    float fscale;        // (E + X) -> FIR output on the 2^11 grid: g_fir * 2^11 / (2^11 * 2^sh)
    float inv_fgrid;     // 2^-11
    float yh_scale;      // 2^11 / g_bq: carried y history (tile 0, rows 0..14) -> the same grid, biquad gain undone
    float descale_rs;    // g_bq * g_out / (2^11 * 2^sh2)
    double b0, b1, b2, a1, a2, g_bq;
    double A16[4];       // block step
    double AL[4];        // A^160: look-back step, incoming state -> state after row 159
    double AL_first[4];  // A^145: the same for tile 0, whose state enters at row 15
    float Wz[16][2];     // zero-state end state of a block: Z = sum_i Wz[i] * (E+X)_i, Wz[i] = A^(15-i) B * fscale / 2^11
    double Wb[4], Wbi[4];          // balanced state coordinates w = Wb s, s = Wbi w; A16 and Wz are given in them
    float Mb[2][kTcBlocks][4];     // in those coordinates: [0]: W A^(16 b) W^-1, incoming state -> state at the start of block b;
                                   // [1]: tile 0: identity for b = 0 (the state enters at row 15), W A^(16 b - 15) W^-1 after
};

// out[m] += sum_k rc[m][2k..2k+1] . s(block 2*(m/32) + k), k = 0..3; rows 147.. : tile 0, slice 0.  In global memory, copied at
// kernel start into the 16 B pads of the staging tile's core-matrix slots (shared memory is full; kernel parameters would be
// read with indexed constant loads, which bounded the output warps).
constexpr int kTcRcFirst = 152;      // row of the tile-0 variant of slice 0 (a multiple of 8: the pad address of a row then splits into slice base + constant)
constexpr int kTcRcRows = kTcRcFirst + kRsN;

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
// Named barrier over the 4 warps of a role group.  Only the group's first warp sleeps on an mbarrier; the others wait here,
// which costs no issue slots (every mbarrier arrival wakes the warps parked on the CTA's mbarriers).
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
// Packed FP32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100): the kernel is bound by instruction issue, and these halve the FP32
// arithmetic instructions of the converter, the drain and the output warps.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// One lane of a converged warp.  The tcgen05 / TMA instructions take their operands from uniform registers: issued under
// `if (lane == 0)` (divergent code) every one of them is wrapped in an ELECT / BRA.U.ANY waterfall with R2UR moves, which cost
// the single issuing thread ~100 cycles per MMA (measured).  With the whole warp running the control flow and only the issue
// elected, the descriptors stay in uniform registers.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &r0, uint32_t &r1)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
// f32 -> f64 and back by bit manipulation: the F2F conversions that involve a 64-bit type cost ~40 cycles per warp here
__device__ __forceinline__ double f2d_bits(float f)
{
    const unsigned u = __float_as_uint(f), a = u & 0x7fffffffu;
    const int hi = (int)((u & 0x80000000u) | (a < 0x00800000u ? 0u : (a >> 3) + 0x38000000u));  // zero / denormal -> 0
    return __hiloint2double(hi, (int)(u << 29));
}
__device__ __forceinline__ float d2f_bits(double d)  // round to nearest (ties up); |d| in the normal f32 range or tiny -> 0
{
    const unsigned hi = (unsigned)__double2hiint(d), lo = (unsigned)__double2loint(d);
    const unsigned a = hi & 0x7fffffffu;
    const unsigned m = ((a - 0x38000000u) << 3) | (lo >> 29);
    const unsigned r = m + ((lo >> 28) & 1u);
    return __uint_as_float((hi & 0x80000000u) | (a < 0x38100000u ? 0u : r));
}

__device__ __forceinline__ long long clk() { return clock64(); }
enum { kProfProdWait = 0, kProfMmaWaitTmem, kProfMmaWaitCvt, kProfMmaIssue, kProfCvtWaitRaw, kProfCvtWaitCvt, kProfCvtWork,
       kProfEpWaitBlk, kProfEpWork, kProfEpLookback, kProfTotal, kProfOutWait, kProfOutMain, kProfMma2Wait, kProfChunk0 = 16, kProfCount = 48 };

// shared memory map (bytes)
constexpr int kRawStageBytes = 16 * kTcCh * 4;                  // 8 KB: 4 sub-tiles of 16 rows x 128 B
constexpr int kCvtStageBytes = 2 * 16 * kTcCh * 2;              // 8 KB: x0 then x1
constexpr int kOffRaw = 0;
constexpr int kOffCvt = kOffRaw + kRawStages * kRawStageBytes;
constexpr int kOffTab = kOffCvt + kCvtStages * kCvtStageBytes;
constexpr int kTabBytes = TcTables::kBytes;                     // 38400 + 40960
// A operand of MMA2: 11 chunks of 16 rows; a chunk is [piece 2][kb 2][mb 16] core matrices of 8 rows x 16 B at a
// 144 B stride (instead of 128 B: the 2-byte scatter stores of a warp then hit 16 different words)
constexpr int kMbStride = 144, kKbStride = 16 * kMbStride, kPieceBytes = 2 * kKbStride, kChunkBytes = 2 * kPieceBytes;
constexpr int kOffStage = ((kOffTab + kTabBytes + 127) / 128) * 128;
constexpr int kStageBytes = kTcRing * kChunkBytes;              // 73728
constexpr int kOffZx = kOffStage + kStageBytes;                 // [11][2][128] float: zero-state end state Z_b of every block
constexpr int kOffBar = kOffZx + kTcBlocks * 2 * kTcCh * 4;
constexpr int kNumBlkBars = kTcChunks - kTcFirstDone;           // 10: one per chunk 17..26
constexpr int kNumBars = 2 * kRawStages + 2 * kCvtStages + kNumBlkBars + kTcBlocks + kTcBlocks + 1 + 2 + 2 + 4 + 4 + 4 + 2;
constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemSlot + 16;
static_assert(kSmemBytes <= 227 * 1024, "K2 shared memory budget");
static_assert(kOffStage % 16 == 0 && kChunkBytes % 16 == 0, "UMMA descriptor alignment");

// byte offset (from the staging tile) of half h (4 floats) of row m of the rc table: pad 2m+h of the 704 pads
__host__ __device__ constexpr int rc_pad_off(int m, int h) { return (m >> 3) * kKbStride + (2 * (m & 7) + h) * kMbStride + 128; }
static_assert(rc_pad_off(kTcRcRows - 1, 1) + 16 <= kStageBytes, "rc table fits in the pads");
static_assert(kTcBlocks - kTcRing == 3, "blocks 8, 9 reuse the slots of blocks 0, 1 (read by slice 0 only), block 10 that of block 2 (slices 0, 1)");

// TMEM columns: D1 = E [0,176) + X [176,352); D2 double-buffered slices [E2 | X2] of 32 + 32 columns from 352;
// mailbox of the 11 block states (2 columns each, 24 with padding) from 480
constexpr uint32_t kColE = 0, kColX = 176, kColD2 = 352, kColMbox = 480;

}  // namespace tc

// One block of 16 FIR columns of one channel: f*2^11 = (E + X) * fscale -> pieces f0 (integer grid) + f1 in the A operand
// of MMA2, and the 16-term sums of the block's zero-state end state.  FIRST0: block 0 of tile 0, whose rows 0..14 are the
// carried y history (they do not drive the biquad) and whose row 15 is frame 0.
template <bool FIRST0>
__device__ __forceinline__ void ep_block(const uint32_t (&re)[16], const uint32_t (&rx)[16], unsigned char *dst, const TcParams &p,
                                         const float *yh, float &p0, float &p1, float &vmax)
{
    using namespace tc;
    f32x2 z0 = pk2(0.f, 0.f), z1 = z0;  // even / odd rows of the two block sums
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        f32x2 t2 = add2(pk2(__uint_as_float(re[i]), __uint_as_float(re[i + 1])), pk2(__uint_as_float(rx[i]), __uint_as_float(rx[i + 1])));
        f32x2 sc2 = pk2(p.fscale, p.fscale);
        if (FIRST0 && i < kTcHr) {
            // rows 0..14 of tile 0 are the carried y history (row 15, in the last pair, is frame 0)
            float ta, tb;
            upk2(t2, ta, tb);
            t2 = pk2(yh[(size_t)i * p.C], i + 1 < kTcHr ? yh[(size_t)(i + 1) * p.C] : tb);
            sc2 = pk2(p.yh_scale, i + 1 < kTcHr ? p.yh_scale : p.fscale);
        }
        float ra, rb;  // t * sc to the nearest integer (|.| < 2^22)
        upk2(add2(fma2(t2, sc2, pk2(12582912.f, 12582912.f)), pk2(-12582912.f, -12582912.f)), ra, rb);
        const __half2 h0 = __floats2half2_rn(ra, rb);  // above 2048 the fp16 grid is coarser than 1:
        const float2 f0 = __half22float2(h0);           // the remainder is taken from what h0 really holds
        float la, lb;
        upk2(fma2(t2, sc2, pk2(-f0.x, -f0.y)), la, lb);
        const __half2 h1 = __floats2half2_rn(la, lb);
        vmax = fmaxf(vmax, fmaxf(fabsf(ra), fabsf(rb)));
        // Wz is pre-multiplied by fscale; the history rows of tile 0 do not drive the biquad
        const float w0a = (FIRST0 && i < kTcHr) ? 0.f : p.Wz[i][0], w1a = (FIRST0 && i < kTcHr) ? 0.f : p.Wz[i][1];
        const float w0b = (FIRST0 && i + 1 < kTcHr) ? 0.f : p.Wz[i + 1][0], w1b = (FIRST0 && i + 1 < kTcHr) ? 0.f : p.Wz[i + 1][1];
        z0 = fma2(pk2(w0a, w0b), t2, z0);
        z1 = fma2(pk2(w1a, w1b), t2, z1);
        unsigned char *d = dst + (i >> 3) * kKbStride + (i & 7) * 16;
        *reinterpret_cast<__half *>(d) = __low2half(h0);
        *reinterpret_cast<__half *>(d + 16) = __high2half(h0);
        *reinterpret_cast<__half *>(d + kPieceBytes) = __low2half(h1);
        *reinterpret_cast<__half *>(d + 16 + kPieceBytes) = __high2half(h1);
    }
    float a, b;
    upk2(z0, a, b);
    p0 += a + b;
    upk2(z1, a, b);
    p1 += a + b;
}

// PROF: per-role cycle counters (PB_TC_PROF=1); a compile-time switch, the counters cost the single-warp issue loops dearly
template <bool PROF>
__global__ void __launch_bounds__(kTcThreads, 1) chain_tc_kernel(const __grid_constant__ TcParams p)
{
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *raw = smem + kOffRaw;
    unsigned char *cvt = smem + kOffCvt;
    __half *tab = reinterpret_cast<__half *>(smem + kOffTab);
    unsigned char *stage = smem + kOffStage;
    float *zx = reinterpret_cast<float *>(smem + kOffZx);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *raw_full = bars, *raw_empty = bars + kRawStages;
    uint64_t *cvt_full = bars + 2 * kRawStages, *cvt_empty = cvt_full + kCvtStages;
    uint64_t *blk_full = cvt_empty + kCvtStages;    // [10] MMA1 (commit after chunk 17+i) -> drain: column block i of D1 is final
    uint64_t *d1_free = blk_full + kNumBlkBars;     // [11] drain warps -> MMA1: column block b of D1 has been read (the next tile may overwrite it)
    uint64_t *a2_ready = d1_free + kTcBlocks;            // [11] drain warps -> MMA2: the f pieces of block b are in the staging tile
    uint64_t *stage_free = a2_ready + kTcBlocks;    //      MMA2 (commit) -> drain: the staging tile has been consumed
    uint64_t *d2_full = stage_free + 1;             // [2]  MMA2 (commit) -> output warps
    uint64_t *d2_empty = d2_full + 2;               // [2]  output warps -> MMA2
    uint64_t *mbox_ready = d2_empty + 2;            // [4]  drain warps A (all four arrive on [0]) -> output warps: block states in the TMEM mailbox
    uint64_t *mbox_free = mbox_ready + 4;           // [4]  output warp e -> drain warps e
    uint64_t *zx_ready = mbox_free + 4;             // [4]  drain warp B -> drain warp A: the Z of the odd blocks are in D1's dead columns
    uint64_t *slice_done = zx_ready + 4;            // [2]  MMA2 (commit) -> drain: slices 0 / 1 have read ring slots 0,1 / 2
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
            mbar_init(&raw_empty[i], 1);
        }
        for (int i = 0; i < kCvtStages; i++) {
            mbar_init(&cvt_full[i], 1);
            mbar_init(&cvt_empty[i], 1);
        }
        for (int i = 0; i < kNumBlkBars; i++) mbar_init(&blk_full[i], 1);
        for (int i = 0; i < kTcBlocks; i++) mbar_init(&d1_free[i], 1);
        for (int i = 0; i < kTcBlocks; i++) mbar_init(&a2_ready[i], 1);
        mbar_init(stage_free, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 1);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(&mbox_ready[i], 4);  // only [0] is used: all four quadrants arrive, one output warp waits
            mbar_init(&mbox_free[i], 2);
            mbar_init(&zx_ready[i], 1);
            if (i < 2) mbar_init(&slice_done[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tables);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = tid; i < kTabBytes / 16; i += kTcThreads) dst[i] = src[i];
    }
    for (int i = tid; i < kTcRcRows * 2; i += kTcThreads)
        *reinterpret_cast<float4 *>(stage + rc_pad_off(i >> 1, i & 1)) = reinterpret_cast<const float4 *>(p.rc)[i];
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
                    const long long c0 = (PROF ? clk() : 0ll);
                    mbar_wait(&raw_empty[s], ph ^ 1);
                    pw += (PROF ? clk() : 0ll) - c0;
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
            if (PROF && p.prof) p.prof[blockIdx.x * kProfCount + kProfProdWait] = pw;
        }
    } else if (warp == 1) {
        // ================================ MMA1 issuer =================================
        // Only the band of the Toeplitz matrix is multiplied: chunk q (frames f0-272+16q ..+15) reaches output columns
        // [16q-257, 16q+14], i.e. 8-column blocks [2q-33, 2q+1] clipped to [0, 21] and widened to an even count (N % 16 == 0).
        // Chunk 0 runs at full width with accumulate = 0: it zeroes the accumulators.  The warp runs the loop converged and elects
        // one lane per issue (x: B descriptor offset >> 4, y: instruction descriptor, z: first accumulator column).
        {
            const uint32_t t0 = smem_u32(tab);
            const uint64_t bd0 = make_desc(t0, 128, 128), bd1 = make_desc(t0 + TcTables::kT * 2, 128, 128);
            const uint64_t bd2 = make_desc(t0 + TcTables::kT * 4, 128, 128), bd3 = make_desc(t0 + TcTables::kT * 6, 128, 128);
            const uint64_t ad0 = make_desc(smem_u32(cvt), 2048, 128);  // stage s: + s * kCvtStageBytes >> 4; x1: + 4096 >> 4
            int s = 0, ph = 0, tph = 0;
            long long w_t = 0, w_c = 0, w_i = 0;
            const long long kstart = (PROF ? clk() : 0ll);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                static_assert(kTcChunks % kCvtStages == 0, "the chunk loop is unrolled by the number of A stages");
#pragma unroll 3
                for (int q = 0; q < kTcChunks; q++) {
                    // Window of chunk q in 8-column blocks: [lo, hi], widened to an even count.  For q <= 10 its last two blocks
                    // (2q, 2q+1) are touched for the first time in this tile: they are written with accumulate = 0 by a
                    // separate N = 16 instruction, so D1 never needs zeroing and the previous tile's drain only has to have
                    // released column block q (16 columns) before chunk q -- MMA1 of the next tile overlaps the drain's tail.
                    int lo = 2 * q - 33 > 0 ? 2 * q - 33 : 0, hi = 2 * q + 1 < kTcN / 8 - 1 ? 2 * q + 1 : kTcN / 8 - 1;
                    if ((hi - lo + 1) & 1) lo--;  // odd only when lo > 0
                    const bool has_new = q < kTcBlocks;           // blocks 2q, 2q+1 are new
                    const int n_old = has_new ? 2 * q : hi - lo + 1;  // 8-column blocks that already hold partial sums
                    const uint32_t toff = (uint32_t)((52 - 2 * q + lo) * 128) >> 4;
                    const uint64_t a0 = ad0 + (uint64_t)(s * (kCvtStageBytes >> 4)), a1 = a0 + (4096 >> 4);
                    const uint64_t b0 = bd0 + toff, b1 = bd1 + toff, b2 = bd2 + toff, b3 = bd3 + toff;
                    const uint32_t dE = tmem_base + kColE + 8 * lo, dX = tmem_base + kColX + 8 * lo;
                    long long c0 = (PROF ? clk() : 0ll);
                    if (has_new && tph) mbar_wait(&d1_free[q], (tph - 1) & 1);  // tph counts this CTA's tiles
                    w_t += (PROF ? clk() : 0ll) - c0;
                    c0 = (PROF ? clk() : 0ll);
                    mbar_wait(&cvt_full[s], ph);
                    const long long c1 = (PROF ? clk() : 0ll);
                    w_c += c1 - c0;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
                        if (!(p.dbg & 1)) {
                            if (n_old > 0) {
                                const uint32_t idn = make_idesc(8 * n_old);
                                umma(dE, a0, b0, idn, 1);   // x0 h0: exact, integers < 2^24
                                umma(dX, a0, b1, idn, 1);   // x0 (h1 + h2): the tap's remainder in two pieces
                            }
                            if (has_new) {
                                const uint32_t idn = make_idesc(16), off = (uint32_t)(n_old * 128) >> 4;
                                umma(dE + 8 * n_old, a0, b0 + off, idn, 0);
                                umma(dX + 8 * n_old, a0, b1 + off, idn, 0);
                            }
                            const uint32_t idw = make_idesc(8 * (hi - lo + 1));
                            umma(dX, a0, b2, idw, 1);
                            umma(dX, a1, b3, idw, 1);       // x1 h: |x1| <= 1/2, the tap rounded once (2^-12 relative) is enough
                        }
                        umma_commit(&cvt_empty[s]);  // frees the A stage when these MMAs have read it
                        if (q >= kTcFirstDone) umma_commit(&blk_full[q - kTcFirstDone]);  // column block q-17 (after 26: 9 and 10) is final
                    }
                    __syncwarp();
                    const long long c2 = (PROF ? clk() : 0ll);
                    w_i += c2 - c1;
                    if (++s == kCvtStages) { s = 0; ph ^= 1; }
                }
                tph++;
            }
            if (PROF && p.prof && lane == 0) {
                long long *pr = p.prof + blockIdx.x * kProfCount;
                pr[kProfMmaWaitTmem] = w_t;
                pr[kProfMmaWaitCvt] = w_c;
                pr[kProfMmaIssue] = w_i;
                pr[kProfTotal] = (PROF ? clk() : 0ll) - kstart;
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
            const long long c0 = (PROF ? clk() : 0ll);
            if (cw == 0) mbar_wait(&raw_full[rs], rph);
            const long long c1 = (PROF ? clk() : 0ll);
            if (cw == 0) mbar_wait(&cvt_empty[cs], cph ^ 1);
            group_sync(1 + grp);
            const long long c2 = (PROF ? clk() : 0ll);
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
                    // x * scale rounded to the nearest integer on the FMA pipe (|.| < 2^22; larger values trip the range check),
                    // and the remainder with a single rounding; two values per instruction
                    const f32x2 x2 = pk2(v[2 * i], v[2 * i + 1]), s2 = pk2(sc, sc);
                    const f32x2 r2 = add2(fma2(x2, s2, pk2(12582912.f, 12582912.f)), pk2(-12582912.f, -12582912.f));
                    float ra, rb, la, lb;
                    upk2(r2, ra, rb);
                    upk2(fma2(x2, s2, pk2(-ra, -rb)), la, lb);
                    hi[i] = __floats2half2_rn(ra, rb);
                    lo[i] = __floats2half2_rn(la, lb);
                    vmax = fmaxf(vmax, fmaxf(fabsf(ra), fabsf(rb)));
                }
                const int off = (1 - kb) * 2048 + mb * 128 + fr_i * 16;  // K-blocks swapped (Toeplitz trick)
                *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<uint4 *>(hi);
                *reinterpret_cast<uint4 *>(dst + 4096 + off) = *reinterpret_cast<uint4 *>(lo);
                const int hrow = 16 * q + row - 176 - (256 - p.hist_rows);
                if (last && hrow >= 0) {
                    // carried FIR input history: frames [n-hist_rows, n) in gain-scaled units (K1's convention)
                    float *hp = p.xhist_next + (size_t)hrow * p.C + cg * kTcCh + mb * 8;
                    const float is = sc * p.inv_scale_in;
                    *reinterpret_cast<float4 *>(hp) = make_float4(v[0] * is, v[1] * is, v[2] * is, v[3] * is);
                    *reinterpret_cast<float4 *>(hp + 4) = make_float4(v[4] * is, v[5] * is, v[6] * is, v[7] * is);
                }
            }
            if (!(p.dbg & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
            group_sync(1 + grp);
            if (cw == 0 && lane == 0) {
                mbar_arrive(&cvt_full[cs]);
                mbar_arrive(&raw_empty[rs]);
            }
            w_w += (PROF ? clk() : 0ll) - c2;
        }
        // x0 must stay on the integer grid the fp16 piece represents exactly (|g x| <= 1): beyond it the split would be silently
        // inexact, so the chain reports an error instead
        if (vmax > 2048.5f) atomicExch(p.err_flag, 2);
        if (PROF && p.prof && warp == 2 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfCvtWaitRaw] = w_r;
            pr[kProfCvtWaitCvt] = w_c;
            pr[kProfCvtWork] = w_w;
        }
    } else if (warp < 18) {
        // ================================ FIR drain warps =============================
        // Two warps per TMEM lane quadrant e (channels cg*128 + 32e + lane): role A (warps 10-13) takes the even
        // blocks of 16 columns, role B (warps 14-17) the odd ones.  Per block: f*2^11 = (E + X) * fscale -> fp16 pieces
        // f0 (integer grid) + f1 into the A operand of MMA2, and the zero-state end state of the block Z_b (16-term float
        // sums) into the TMEM mailbox.  Role A then runs the block-state recursion in double and the look-back.
        const int e = warp & 3;
        const bool roleB = warp >= 14;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        unsigned char *stp = stage + (e * 4 + (lane >> 3)) * kMbStride + (lane & 7) * 2;  // + chunk + piece + kb + fr*16
        float vmax = 0.f;
        long long e_w = 0, e_m = 0, e_l = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const bool first = (t == 0), last = (t == p.n_tiles - 1);
            const int c = cg * kTcCh + e * 32 + lane;
            const int grp = cg * 4 + e;  // 32-channel look-back group, same indexing as K1
            const uint32_t par = it & 1;
            const size_t slot = (size_t)grp * p.n_tiles + t;
            const bool chained = !first && !last;
            const float *yh = p.yhist + c;
            const bool gl = (warp & 3) == 2;  // first warp of the role group (warps 10 / 14)
            const int gid = roleB ? 4 : 3;
            if (gl && it > 0) mbar_wait(stage_free, par ^ 1);  // MMA2 has consumed the previous tile's pieces
#pragma unroll 1
            for (int b = roleB ? 1 : 0; b < kTcBlocks; b += 2) {
                const long long k0 = (PROF ? clk() : 0ll);
                if (gl) {
                    mbar_wait(&blk_full[b < kNumBlkBars ? b : kNumBlkBars - 1], par);
                    if (b >= kTcRing) mbar_wait(&slice_done[b == kTcBlocks - 1 ? 1 : 0], par);  // the slot's previous block has been multiplied
                }
                group_sync(gid);
                const long long k1 = (PROF ? clk() : 0ll);
                e_w += k1 - k0;
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + 16 * b, re);
                tmem_ld16(tmem_base + lane_base + kColX + 16 * b, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
                unsigned char *dst = stp + (b % kTcRing) * kChunkBytes;
                if (first && b == 0) ep_block<true>(re, rx, dst, p, yh, p0, p1, vmax);
                else ep_block<false>(re, rx, dst, p, yh, p0, p1, vmax);
                if (!(p.dbg & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
                asm volatile("tcgen05.fence::before_thread_sync;");
                group_sync(gid);
                if (gl && lane == 0) {
                    mbar_arrive(&a2_ready[b]);
                    // all four quadrants have read the block's columns.  The last block is released only after role A has
                    // read the block sums below: that keeps the next tile's role B (which overwrites them) behind MMA1 chunk 10
                    if (b != kTcBlocks - 1) mbar_arrive(&d1_free[b]);
                }
                zx[(2 * b + 0) * kTcCh + e * 32 + lane] = p0;
                zx[(2 * b + 1) * kTcCh + e * 32 + lane] = p1;
                e_m += (PROF ? clk() : 0ll) - k1;
            }
            const long long k4 = (PROF ? clk() : 0ll);
            if (roleB) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&zx_ready[e]);
                e_l += (PROF ? clk() : 0ll) - k4;
                continue;
            }
            // ---- role A: block-state recursion in double, s(b+1) = A^16 s(b) + Z_b from a zero state; the mailbox entries
            //      Z_b are replaced by the state at the START of block b (float) for the output warp
            mbar_wait(&zx_ready[e], par);
            double s10_1 = 0.0, s10_2 = 0.0;  // zero-state state after row 159
            float szs[kTcBlocks][2];          // zero-state state at the start of each block
            {
                uint32_t z[kTcBlocks][2];
#pragma unroll
                for (int b = 0; b < kTcBlocks; b++) {
                    z[b][0] = __float_as_uint(zx[(2 * b + 0) * kTcCh + e * 32 + lane]);
                    z[b][1] = __float_as_uint(zx[(2 * b + 1) * kTcCh + e * 32 + lane]);
                }
                group_sync(gid);
                if (gl && lane == 0) mbar_arrive(&d1_free[kTcBlocks - 1]);
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int b = 0; b < kTcBlocks; b++) {
                    // the whole block recursion runs in balanced coordinates w = W s (W = S V^T of the block's free-response
                    // matrix): in the TDF-II basis a filter with poles near z = 1 has free responses that cancel to 1e-2 of
                    // their terms, which neither the float sums Z_b nor the float states can afford (measured 6e-6 on a 200 Hz
                    // high-pass)
                    szs[b][0] = d2f_bits(s1);
                    szs[b][1] = d2f_bits(s2);
                    const double n1 = fma(p.A16[0], s1, fma(p.A16[1], s2, f2d_bits(__uint_as_float(z[b][0]))));
                    const double n2 = fma(p.A16[2], s1, fma(p.A16[3], s2, f2d_bits(__uint_as_float(z[b][1]))));
                    s1 = n1;
                    s2 = n2;
                    if (b == kTcBlocks - 2) {  // back to the TDF-II basis, which the look-back arrays and K1 use
                        s10_1 = fma(p.Wbi[0], s1, p.Wbi[1] * s2);
                        s10_2 = fma(p.Wbi[2], s1, p.Wbi[3] * s2);
                    }
                }
            }
            if (chained) {
                // the aggregate of the tile for the look-back of the tiles behind it
                p.lb_agg[slot * 64 + lane * 2] = s10_1;
                p.lb_agg[slot * 64 + lane * 2 + 1] = s10_2;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
            }
            // ---- incoming state: decoupled look-back (same protocol and arrays as K1, 32-channel groups)
            double q1 = 0.0, q2 = 0.0;
            if (first) {
                q1 = p.bq_state[2 * c];
                q2 = p.bq_state[2 * c + 1];
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
                q1 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2);
                q2 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2 + 1);
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
                            mat2_apply(p.AL, q1, q2);
                            q1 += a0[u];
                            q2 += a1[u];
                        }
                }
            }
            // ---- true state at the start of every block = zero-state part + response to the incoming state -> output warps
            {
                const float qf0 = d2f_bits(fma(p.Wb[0], q1, p.Wb[1] * q2)), qf1 = d2f_bits(fma(p.Wb[2], q1, p.Wb[3] * q2));
                const int fi = first ? 1 : 0;
                if (it > 0) mbar_wait(&mbox_free[e], par ^ 1);  // the output warps have read the previous tile's block states
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int b = 0; b < kTcBlocks; b++) {
                    const float *M = p.Mb[fi][b];
                    const float v0 = fmaf(M[0], qf0, fmaf(M[1], qf1, szs[b][0])), v1 = fmaf(M[2], qf0, fmaf(M[3], qf1, szs[b][1]));
                    tmem_st2(tmem_base + lane_base + kColMbox + 2 * b, __float_as_uint(v0), __float_as_uint(v1));
                }
                tmem_st2(tmem_base + lane_base + kColMbox + 2 * kTcBlocks, 0u, 0u);  // the last slice reads a fourth, absent block
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(&mbox_ready[0]);
            }
            // true state after row 159
            double I0 = q1, I1 = q2;
            mat2_apply(first ? p.AL_first : p.AL, I0, I1);
            I0 += s10_1;
            I1 += s10_2;
            if (!last) {
                p.lb_inc[slot * 64 + lane * 2] = I0;
                p.lb_inc[slot * 64 + lane * 2 + 1] = I1;
                __syncwarp();
                if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
            } else {
                // carried biquad state (after row 174) and carried y history (rows 160..174): the only place where the
                // recursion runs row by row, from the pieces of f just written (f * 2^11 = f0 + f1)
                const double nb1 = p.b1, nb2 = p.b2, na1 = -p.a1, na2 = -p.a2;
                for (int r = 0; r < kTcHr; r++) {
                    const unsigned char *d = stp + ((kTcBlocks - 1) % kTcRing) * kChunkBytes + (r >> 3) * kKbStride + (r & 7) * 16;
                    const double x = (double)((__half2float(*reinterpret_cast<const __half *>(d)) +
                                               __half2float(*reinterpret_cast<const __half *>(d + kPieceBytes))) * p.inv_fgrid);
                    const double v = fma(p.b0, x, I0);
                    const double tt = fma(nb1, x, I1);
                    I0 = fma(na1, v, tt);
                    I1 = fma(na2, v, nb2 * x);
                    p.yhist_next[(size_t)r * p.C + c] = (float)(v * p.g_bq);
                }
                p.bq_state_next[2 * c] = I0;
                p.bq_state_next[2 * c + 1] = I1;
            }
            e_l += (PROF ? clk() : 0ll) - k4;
        }
        if (vmax > 60000.f) atomicExch(p.err_flag, 2);
        if (PROF && p.prof && warp == 10 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfEpWaitBlk] = e_w;
            pr[kProfEpWork] = e_m;
            pr[kProfEpLookback] = e_l;
        }
    } else if (warp < 26) {
        // ================================ output warps ================================
        // two warps per TMEM lane quadrant: warps 18-21 take outputs 0..15 of every slice, warps 22-25 outputs 16..31
        const int e = warp & 3;
        const int hsel = warp >= 22 ? 1 : 0;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        long long r_w = 0, r_m = 0, o_sync = 0, o_mbox = 0, o_ld = 0, o_out = 0;
        int it = 0;
        unsigned nsl = 0;  // running slice number: D2 buffer nsl & 1, phase (nsl >> 1) & 1
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const bool first = (t == 0);
            const int c = cg * kTcCh + e * 32 + lane;
            const uint32_t par = it & 1;
            float *outp = p.out + (size_t)t * kTcOut * p.C + c;
            float m_peak = 0.f;
            double m_sumsq = 0.0;
            const bool meter = p.meter_peak != nullptr;
            const long long k5 = (PROF ? clk() : 0ll);
#pragma unroll 1
            for (int s = 0; s < kRsSlices; s++, nsl++) {
                const uint32_t b = nsl & 1u;
                const long long k6 = (PROF ? clk() : 0ll);
                if (warp == 18) {
                    mbar_wait(&d2_full[b], (nsl >> 1) & 1u);
                    if (s == 0) mbar_wait(&mbox_ready[0], par);  // the block states of all four quadrants are in the mailbox
                }
                const long long g0 = (PROF ? clk() : 0ll);
                asm volatile("bar.sync 5, 256;" ::: "memory");
                o_sync += (PROF ? clk() : 0ll) - g0;
                const long long k7 = (PROF ? clk() : 0ll);
                r_w += k7 - k6;
                asm volatile("tcgen05.fence::after_thread_sync;");
                // true state at the start of the four blocks this slice reads, and this warp's 16 accumulator columns
                uint32_t zs[8], e16[16], x16[16];
                tmem_ld8(tmem_base + lane_base + kColMbox + 4 * s, zs);
                tmem_ld16(tmem_base + lane_base + kColD2 + 64 * b + 16 * hsel, e16);
                tmem_ld16(tmem_base + lane_base + kColD2 + 64 * b + 32 + 16 * hsel, x16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                asm volatile("bar.sync 5, 256;" ::: "memory");
                if (warp == 18 && lane == 0) mbar_arrive(&d2_empty[b]);
                if (s == kRsSlices - 1 && lane == 0) mbar_arrive(&mbox_free[e]);
                const long long k9 = (PROF ? clk() : 0ll);
                o_ld += k9 - k7;
                const f32x2 sb01 = pk2(__uint_as_float(zs[0]), __uint_as_float(zs[1])), sb23 = pk2(__uint_as_float(zs[2]), __uint_as_float(zs[3]));
                const f32x2 sb45 = pk2(__uint_as_float(zs[4]), __uint_as_float(zs[5])), sb67 = pk2(__uint_as_float(zs[6]), __uint_as_float(zs[7]));
                const int nout = ((s == kRsSlices - 1) ? kTcOut - kRsN * (kRsSlices - 1) : kRsN) - 16 * hsel;
                const unsigned char *rcg = stage + ((((first && s == 0) ? kTcRcFirst : kRsN * s) >> 3) + 2 * hsel) * kKbStride + 128;
                float *op = outp + (size_t)(kRsN * s + 16 * hsel) * p.C;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const unsigned char *rci = rcg + (i >> 3) * kKbStride + 2 * (i & 7) * kMbStride;
                    const float4 ca = *reinterpret_cast<const float4 *>(rci);
                    const float4 cb = *reinterpret_cast<const float4 *>(rci + kMbStride);
                    f32x2 acc = mul2(pk2(ca.x, ca.y), sb01);
                    acc = fma2(pk2(ca.z, ca.w), sb23, acc);
                    acc = fma2(pk2(cb.x, cb.y), sb45, acc);
                    acc = fma2(pk2(cb.z, cb.w), sb67, acc);
                    float c0, c1;
                    upk2(acc, c0, c1);
                    const float o = fmaf(__uint_as_float(e16[i]) + __uint_as_float(x16[i]), p.descale_rs, c0 + c1);
                    if (i < nout) {
                        op[(size_t)i * p.C] = o;
                        if (meter) {
                            m_peak = fmaxf(m_peak, fabsf(o));
                            m_sumsq += (double)o * (double)o;
                        }
                    }
                }
                o_out += (PROF ? clk() : 0ll) - k9;
            }
            r_m += (PROF ? clk() : 0ll) - k5;
            if (meter) {
                atomic_max_nonneg(p.meter_peak + c, (double)m_peak);
                atomicAdd(p.meter_sumsq + c, m_sumsq);
            }
        }
        if (PROF && p.prof && warp == 18 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfOutWait] = r_w;
            pr[kProfOutMain] = r_m;
            pr[kProfChunk0 + 0] = o_sync;
            pr[kProfChunk0 + 1] = o_mbox;
            pr[kProfChunk0 + 2] = o_ld;
            pr[kProfChunk0 + 3] = o_out;
        }
    } else {
        // ================================ MMA2 issuer =================================
        {
            const uint32_t st0 = smem_u32(stage);
            const uint32_t b2 = smem_u32(tab) + TcTables::kHalfs * 2;
            constexpr uint32_t idesc32 = make_idesc(kRsN), idesc64 = make_idesc(2 * kRsN);
            unsigned nsl = 0;
            int it = 0;
            long long w_y = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const bool first = (tile / p.n_cg) == 0;
                int pair = 0;
#pragma unroll 1
                for (int s = 0; s < kRsSlices; s++, nsl++) {
                    const uint32_t b = nsl & 1u;
                    const int nch = (s == kRsSlices - 1) ? 3 : 4;
                    const long long c0 = (PROF ? clk() : 0ll);
                    mbar_wait(&a2_ready[2 * s + nch - 1], it & 1);  // each drain role stages its blocks (even / odd) in order
                    mbar_wait(&a2_ready[2 * s + nch - 2], it & 1);
                    w_y += (PROF ? clk() : 0ll) - c0;
                    mbar_wait(&d2_empty[b], ((nsl >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t dE = tmem_base + kColD2 + 64 * b, dX = dE + kRsN;
#pragma unroll 1
                    for (int k = 0; k < nch; k++, pair++) {
                        const uint32_t a_base = st0 + ((2 * s + k) % kTcRing) * kChunkBytes;
                        const uint64_t a0 = make_desc(a_base, kKbStride, kMbStride), a1 = make_desc(a_base + kPieceBytes, kKbStride, kMbStride);
                        const uint32_t tb = b2 + ((first && pair == 0) ? kRsPairs : pair) * 2048;
                        const uint64_t r0 = make_desc(tb, 128, 256), r1 = make_desc(tb + 1024, 128, 256);
                        if (!(p.dbg & 2) && elect_one()) {
                            umma(dE, a0, r0, idesc64, k > 0);  // f0 * [p0 | p1] -> [E2 | X2]; E2 exact: integers < 2^24
                            umma(dX, a1, r0, idesc32, 1);
                            umma(dX, a1, r1, idesc32, 1);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(&d2_full[b]);
                        if (s < 2) umma_commit(&slice_done[s]);
                        if (s == kRsSlices - 1) umma_commit(stage_free);
                    }
                    __syncwarp();
                }
            }
            if (PROF && p.prof && lane == 0) p.prof[blockIdx.x * kProfCount + kProfMma2Wait] = w_y;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

#endif  // __CUDACC__

}  // namespace pb
