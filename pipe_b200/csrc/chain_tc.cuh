// chain_tc.cuh -- K2: the headline Processor run [gain, FIR<=257, biquad, resample 147/160 x16] on
// tcgen05 / TMEM / TMA.  Replaces the ProcessFunc walk of Processor.execute (reference pipe.go:438)
// for that run when the call is aligned to 160-frame tiles; everything else goes through K1
// (chain_tile.cuh), with which it shares every piece of carried state.
//
// Two contractions per tile of 128 channels x 160 frames, both with channels on M (TMEM lane = channel) and BOTH WITH
// THE A OPERAND IN TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): the tensor core then reads only the small B tables
// from shared memory.  With A in shared memory an M = 128 instruction fetches 4 KB of A on top of N x 32 B of B, which at
// N <= 176 is 100 of the 128 B/cycle the SM's shared memory delivers -- the tensor core has priority, so the converter,
// drain and TMA traffic of the same SM starved (profiles/r02_probe_ts_mma.txt: next to a saturated SS stream the other
// warps got 35 B/cycle, next to a TS stream 73; small-N instructions drop from 54-60 to 21-42 cycles).
//
//  MMA1 (FIR)                 D1[ch x 176] = X^T[ch x 432] * Toeplitz[432 x 176]
//      N = 176 columns (15 left-context frames recomputed for the resampler + 160 frames + 1 pad),
//      K = 432 input frames in 27 chunks of 16; only the band of the Toeplitz matrix is multiplied.
//      A: the converter warps (thread = channel) split 16 frames of their channel into fp16 pieces and write them
//      with tcgen05.st into a 4-stage ring of 16 TMEM columns (x0: 8 columns, x1: 8 columns, two frames per column).
//  MMA2 (biquad + resampler)  D2[ch x 147] = F^T[ch x 176] * P[176 x 147]
//      F is the FIR output.  The biquad is linear, so inside one block of 16 rows its zero-state response is a
//      lower-triangular Toeplitz matrix G16; P = blockdiag(G16) * R folds it into the 147/160 polyphase matrix R
//      of the tile and keeps R's block structure: five slices of 32 outputs, each touching 4 (the last: 3) chunks
//      of 16 rows -- 19 (slice, chunk) blocks.  What the blocks do not see is the biquad state at each block
//      start: y is linear in it, so its contribution is a rank-2 correction per block, carried through the
//      resampler on the host (rc = R * (A^k)_row0) and added when D2 is drained.
//      A: the drain warps write the pieces of F IN PLACE over the 16 accumulator columns of the block they just read
//      (f0: 8 columns, f1: 8 columns); MMA1 of the next tile may touch a block again once the MMA2 slices that read
//      it have completed (tcgen05.commit -> slice_done).
//
// Why the biquad is NOT a per-row recursion here (it was): the recursion has to run in double (an f32 TDF-II with
// poles near z = 1 measured 1e-5), and on this part a running tcgen05.mma stream throttles DFMA about 7x (measured:
// 3.0 -> 21.9 cycles per warp instruction, tools/mma_probe.cu), conversions between f32 and f64 cost ~40 cycles each,
// so 176 rows x 6 DFMA per channel bounded the kernel at ~20 k cycles per tile.  Folded into P, the only double
// arithmetic left is the block-state recursion s' = A^16 s + Z_b: 4 DFMA per 16 rows.
//
// Precision (measured with tools/tc_probe.cu): the tensor core adds into its fp32 accumulator with
// truncation, which costs ~1e-6 over 27..81 steps.  So the operands are split on integer grids:
//   x*sigma_c = x0 + x1,  h*2^sh = h0 + h1 + h2,  x0 and h0 integer-valued fp16 (|.| <= 2048), the others
//   fp16 remainders (h needs the third piece: the remainder of a tap has ABSOLUTE precision 2^-13 of the
//   grid, which summed over 257 taps was 5.7e-7 of the peak).  x0*h0 goes to accumulator E: every product
//   and every partial sum is an integer below 2^24, so E is EXACT.  x0*h1 + x0*h2 + x1*h go to
//   accumulator X, 2^-9 of E in magnitude, whose truncation is negligible (x1 is multiplied by the tap rounded once:
//   four passes).
//   MMA2 uses the same scheme with two pieces each (f*sigma_c = f0 + f1, P*2^sh2 = p0 + p1; at most 32 rows per
//   output): E2 = f0*p0 exact, X2 = f0*p1 + f1*p0 + f1*p1.
//
// PER-CHANNEL BLOCK EXPONENT.  sigma_c is a scale per channel (a float, not a power of two) chosen so that the channel's
// own peak |g x| over the call lands at ~1900 of the 2047 grid units: the noise floor of the split then follows each
// channel's level (a channel at -60 dBFS is served with the same relative accuracy as one at full scale, and |g x| may
// be anything -- nothing raises an error).  The peak of a call is only known after it has been read, so the scale is
// SPECULATED and VERIFIED: pass A runs with the scale the previous call ended with and measures every channel's peak on
// the way (one FMNMX per sample in the converter, one atomicMax per channel and tile); pass B is the same kernel launched
// right behind it: its prologue checks every channel's peak against the window [kSigLo, kSigHi] grid units, writes the
// exact scale for the next call, and returns at once when every channel was inside (the steady state: ~2 us).  Otherwise
// it redoes the call with the exact scales; carried state is ping-ponged and the look-back arrays are epoch-tagged, so
// the second pass simply overwrites the first.  The fused meter accumulates pass A into a scratch copy that pass B's
// prologue folds into the chain's meter (or drops when it redoes the call).
//
// B operand of MMA1: the Toeplitz matrix is never materialised.  A K-major 8x8 core matrix depends only on
// (column block - row block); with the two K-blocks of an instruction stored swapped in A, the
// descriptor strides LBO = SBO = 128 B make 75 core matrices (9.6 KB per piece) serve all 27 x 22
// positions.  B operand of MMA2: the 19 blocks [32 outputs x 16 rows] of P, 1 KB per piece, the two pieces
// adjacent so that one N = 64 instruction multiplies f0 by [p0 | p1] into [E2 | X2].
//
// Roles (864 threads, one persistent CTA per SM): warp 0 TMA producer (8-stage ring of 8 KB chunks, one box of
// 128 channels x 16 frames each), warp 1 MMA1 issuer, warps 2-5 / 6-9 two converter groups on alternate chunks (one warp
// per TMEM lane quadrant), warps 10-13 / 14-17 two FIR drain groups on the even / odd 16-column blocks of D1 (D1 -> f
// pieces, block by block as MMA1's last chunks complete them, each group chaining the zero-state end states of its blocks on
// the way; once block 9 is drained the second group combines the two chains into the tile's aggregate, publishes it and does
// the look-back while the first group runs the block-state recursion), warps 18-21 / 22-25 two output groups on the two halves of every slice (D2 -> registers, free the
// buffer, block-state correction, coalesced stores, meter), warp 26 MMA2 issuer.  Tiles follow a static time-major
// schedule: the look-back spins on tiles owned by other CTAs of the same grid, so the grid (<= one CTA per SM) must be
// co-resident -- the host launches the kernel cooperatively (cudaLaunchAttributeCooperative), which starts a grid only when all of it fits.  The drain warps hand the 11
// block states to the output warps through 24 spare TMEM columns.  Inside a role group only the first warp polls
// mbarriers, the others wait on a named barrier; the issuing warps run converged and elect one lane per tcgen05
// instruction.  MMA1 writes the first touch of every 16 columns with accumulate = 0 and waits per column block for the
// previous tile's MMA2 slice, so it overlaps the tail of that tile.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>
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
constexpr int kRawStages = 8;       // TMA ring (shared memory): ~2 k cycles of L2 latency at 8 KB per ~450 cycles, plus the 4 chunks the converters hold
constexpr int kA1Stages = 8;        // converted-chunk ring (tensor memory)
constexpr int kTcBlocks = kTcN / 16;     // 11 blocks of 16 rows: biquad blocks == K chunks of MMA2
constexpr int kTcFirstDone = 17;    // column block b of D1 is complete after chunk b + 17 (the last two after chunk 26)
constexpr int kRsSlices = 5;        // slices of 32 outputs; slice s reads row blocks [2s, 2s+4) (the last one 3)
constexpr int kRsN = 32;
constexpr int kRsPairs = 19;        // (slice, block) blocks of P; entry 19 is (0, 0) for tile 0 (rows 0..14 are y history)

// Two scale classes per channel: tiles 0 and 1 of a call, whose windows reach into the carried history (class 0), and all the others
// (class 1).  With one scale per call a channel that was loud in the previous call and is quiet in this one would be served on the
// loud grid to the end of the call, and the carried state it leaves behind (biquad state, y history) would carry the loud grid's
// absolute error into the next call (found by tools/k2_soak.py: 1e-4 of the quiet channel's peak in the first outputs of that call).
constexpr int kTcHistTiles = 2;
__host__ __device__ constexpr int tc_scale_class(int t) { return t < kTcHistTiles ? 0 : 1; }
// per-channel scale window, in grid units of the channel's peak |g x| (the x0 piece must stay below 2048)
constexpr float kSigTarget = 1900.f, kSigLo = 1400.f, kSigHi = 2047.f, kSigCap = 7.9e28f /* 2^96 */;

struct TcTables {  // tables in global memory, copied to shared at kernel start
    static constexpr int kT = kTcCores * 64;           // T0 T1 T2 (three-piece split) and T3 (the tap rounded once): 75*64 halfs each
    static constexpr int kHalfs = 4 * kT;
    static constexpr int kB2 = (kRsPairs + 1) * 2 * 512;  // then P blocks [pair][piece][nb 4][kb 2][8][8] halfs
    static constexpr int kBytes = (kHalfs + kB2) * 2;
};

struct TcParams {
    CUtensorMap tm_in;    // [n_frames][C] f32, box 128 ch x 16 frames, no swizzle
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
    double *meter_peak, *meter_sumsq;       // where THIS pass accumulates (pass A: the scratch copy, pass B: the chain's meter); or nullptr
    double *meter_main, *meter_scratch;     // [2][C] each: pass B's prologue folds (or drops) what pass A accumulated
    // per-channel block exponent (see the header comment): scale[0] = sigma, scale[1] = 1 / sigma, [2 classes][C] each (tc_scale_class)
    const float *scale;       // used by pass A
    float *scale_next;        // written by pass B's prologue from the peaks pass A measured; used by pass B and by the next call
    unsigned *peak;           // [2 classes][C] float bits of max |g x| over the windows of the class's tiles (pass A: atomicMax)
    unsigned *peak_next;      // [2 classes][C] zeroed by pass B's prologue for the next call
    int pass;                 // 0: pass A (speculated scales, measures the peaks); 1: pass B (verifies; redoes the call if needed)
    int *err_flag;
    long long *prof;     // optional per-CTA cycle counters (PB_TC_PROF=1), nullptr otherwise
    long long *trace;    // optional event timeline of CTA 0, tiles kTraceIt0.. (PB_TC_PROF=1): [kTraceTiles][kTraceRoles][32] clock64 values
    int dbg;             // development switches (PB_TC_DBG): bit0 skip MMA1, bit1 skip MMA2, bit2 skip the conversion, bit4 skip the TMA loads, bit6 no look-back, bit7 no output arithmetic
    int C, n_tiles, n_cg;
    int hist_rows;       // rows of xhist == FIR taps - 1 (<= 256)
    int last_frames;     // input frames the call's last tile really holds (1..160; the rows behind them are zero-filled by TMA)
    int last_outputs;    // outputs those frames trigger at the call's resampler phase (<= 147): the rest of that tile is not stored
    unsigned epoch;
    float g_load;        // gain applied to the frames of this call (history frames are already gain-scaled)
    float fscale;        // (E + X) -> FIR output on the channel's grid: g_fir / 2^sh
    float inv_gbq;       // 1 / g_bq: carried y history (tile 0, rows 0..14) -> biquad gain undone
    float descale_rs;    // g_bq * g_out / 2^sh2 (times 1 / sigma_c per channel)
    double b0, b1, b2, a1, a2, g_bq;
    double A16[4];       // block step
    double A32[4];       // two block steps: the drain groups each chain the blocks of one parity (the early tile aggregate)
    double AL[4];        // A^160: look-back step, incoming state -> state after row 159
    double AL_first[4];  // A^145: the same for tile 0, whose state enters at row 15
    float Wz[16][2];     // zero-state end state of a block on the channel's grid: Z = sum_i Wz[i] * (E+X)_i, Wz[i] = A^(15-i) B * fscale
    double Wb[4], Wbi[4];          // balanced state coordinates w = Wb s, s = Wbi w; A16 and Wz are given in them
    float Mb[2][kTcBlocks][4];     // in those coordinates: [0]: W A^(16 b) W^-1, incoming state -> state at the start of block b;
                                   // [1]: tile 0: identity for b = 0 (the state enters at row 15), W A^(16 b - 15) W^-1 after
};

// out[m] += sum_k rc[m][2k..2k+1] . s(block 2*(m/32) + k), k = 0..3; rows kTcRcFirst.. : tile 0, slice 0
constexpr int kTcRcFirst = 152;
constexpr int kTcRcRows = kTcRcFirst + kRsN;

#ifdef __CUDACC__

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Wait for a phase of an mbarrier with the hardware-suspended try_wait (SASS TRYWAIT + NANOSLEEP.SYNCS): ~90 cycles when the
// phase is already complete, ~60 from the arrival otherwise.  (Measured alternative, profiles/r02_k2v4_summary.md: test_wait +
// nanosleep back-off removes the wake-up instructions -- a parked warp is woken by EVERY mbarrier arrival of the CTA, 17 % of
// all executed instructions -- but costs 149 cycles per successful test and the back-off latency, and the kernel is bound by
// the latency of its hand-offs, not by issue slots: 0.343 -> 0.383 ms.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
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
// plain bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion on an mbarrier; size and addresses multiples of 16
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
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
    // c=F32 (1<<4), a=b=F16 (0), a K-major (the A operand lives in tensor memory), b K-major, N>>3 at bit 17, M>>4 at bit 24
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: A is 128 lanes x 8 columns, column c of a lane = K elements 2c (low half) and 2c+1
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(a[0]),
                 "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]),
                 "r"(b[5]), "r"(b[6]), "r"(b[7])
                 : "memory");
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
constexpr int kTraceIt0 = 8, kTraceTiles = 4, kTraceRoles = 7;   // roles: 0 MMA1 (per chunk), 1/2 converter groups (per chunk), 3/4 drain A/B, 5 MMA2 (per slice), 6 output
#define PB_TRACE(role, it_, idx)                                                                                          \
    do {                                                                                                                  \
        if (PROF == 2 && p.trace && blockIdx.x == 0 && (it_) >= kTraceIt0 && (it_) < kTraceIt0 + kTraceTiles && lane == 0)      \
            p.trace[(((it_) - kTraceIt0) * kTraceRoles + (role)) * 32 + (idx)] = clock64();                               \
    } while (0)
enum { kProfProdWait = 0, kProfMmaWaitTmem, kProfMmaWaitCvt, kProfMmaIssue, kProfCvtWaitRaw, kProfCvtWaitCvt, kProfCvtWork,
       kProfEpWaitBlk, kProfEpWork, kProfEpLookback, kProfTotal, kProfOutWait, kProfOutMain, kProfMma2Wait, kProfOutLd, kProfOutMath,
       kProfNs, kProfCount = 18 };

// shared memory map (bytes)
constexpr int kRawStageBytes = 16 * kTcCh * 4;                  // 8 KB: 16 frames x 128 channels f32, frame-major
constexpr int kOffRaw = 0;
constexpr int kOffTab = kOffRaw + kRawStages * kRawStageBytes;
constexpr int kTabBytes = TcTables::kBytes;                     // 38400 + 40960
constexpr int kOffRc = ((kOffTab + kTabBytes + 127) / 128) * 128;
constexpr int kRcBytes = kTcRcRows * 8 * 4;
constexpr int kOffZx = kOffRc + kRcBytes;                       // [11][2][128] float: zero-state end state Z_b of every block
constexpr int kZxFloats = kTcBlocks * 2 * kTcCh;
constexpr int kOffPark = kOffZx + kZxFloats * 4;                // [128][128] float: the resampler sums E2 + X2 of slices 0..3, parked until the block states arrive
constexpr int kParkRows = (kRsSlices - 1) * kRsN;               // (the last slice waits in registers: nothing queues behind it)
constexpr int kOffBar = kOffPark + kParkRows * kTcCh * 4;
constexpr int kNumBlkBars = kTcChunks - kTcFirstDone;           // 10: one per chunk 17..26
constexpr int kNumBars = 2 * kRawStages + 2 * kA1Stages + kNumBlkBars + kTcBlocks + kRsSlices + 2 + 1 + 4 + 4 + 4 + 4 + 4 + 1;
constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr int kOffSa = ((kOffTmemSlot + 16 + 15) / 16) * 16;                      // [2][128][2] double: drain group A's chained state of the even blocks (per tile parity)
constexpr int kSmemBytes = kOffSa + 2 * kTcCh * 2 * 8;
static_assert(kSmemBytes <= 227 * 1024, "K2 shared memory budget");
static_assert(kOffTab % 128 == 0 && kOffBar % 8 == 0 && kTabBytes % 16 == 0 && kRcBytes % 16 == 0, "alignment");

// TMEM columns: D1 = E [0,176) + X [176,352) -- block b's 16 E columns are overwritten in place by the pieces of F (f0: 8
// columns, f1: 8 columns), the A operand of MMA2; ring of 8 converted chunks (A operand of MMA1: x0 8 columns, x1 8 columns) from
// 352; mailbox of the 11 block states (2 columns each, 24 with padding) from 480, the tile's zero-state end state (2 doubles) from 504
constexpr uint32_t kColE = 0, kColX = 176, kColA1 = 352, kColMbox = 480, kColS10 = 504;
static_assert(kColA1 + 16 * kA1Stages == kColMbox && kColMbox + 2 * kTcBlocks + 2 <= kColS10 && kColS10 + 4 <= 512, "TMEM map");
// D2 of slice s = [E2 | X2], 32 + 32 columns, lives in the X columns of the blocks the slice reads: the drain has read them, so
// they are dead until the next tile's MMA1 touches them again (which waits for the output warps to have read the slice)
__host__ __device__ constexpr uint32_t d2_col(int s) { return kColX + (s < kRsSlices - 1 ? 32 * s : kTcN - 64); }

// What the MMA1 issuer needs for chunk q, worked out at compile time (constant bank, uniform loads): the issuing warp is a single
// chain of dependent instructions at ~7 cycles each, and deriving the band window, the instruction descriptors and the table
// offsets of a chunk in that chain cost ~90 instructions per chunk -- the chunk pace (650 cycles) that the accumulator hand-over
// between consecutive tiles multiplies by twenty.
struct TcChunk {
    uint32_t toff;   // descriptor offset (16 B units) of the Toeplitz window of the chunk
    uint32_t dcol;   // first accumulator column of the window
    uint32_t idn;    // instruction descriptor over the columns that already hold partial sums (0: none)
    uint32_t idw;    // ... over the whole window
    uint32_t ncol;   // first-touch blocks: column offset inside the window (q < 11) ...
    uint32_t noff;   // ... and their descriptor offset
    uint32_t has_new;
    int32_t blk;     // >= 0: the commit of this chunk completes column block `blk` of D1 (blk_full barrier index)
};
__host__ __device__ constexpr uint32_t make_idesc_c(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__host__ __device__ constexpr TcChunk make_chunk(int q)
{
    int lo = 2 * q - 33 > 0 ? 2 * q - 33 : 0;
    const int hi = 2 * q + 1 < kTcN / 8 - 1 ? 2 * q + 1 : kTcN / 8 - 1;
    if ((hi - lo + 1) & 1) lo--;  // odd only when lo > 0
    const bool has_new = q < kTcBlocks;
    const int n_old = has_new ? 2 * q : hi - lo + 1;
    return TcChunk{(uint32_t)((52 - 2 * q + lo) * 128) >> 4,
                   (uint32_t)(8 * lo),
                   n_old > 0 ? make_idesc_c(8 * n_old) : 0u,
                   make_idesc_c(8 * (hi - lo + 1)),
                   (uint32_t)(8 * n_old),
                   has_new ? (uint32_t)(n_old * 128) >> 4 : 0u,
                   has_new ? 1u : 0u,
                   q >= kTcFirstDone ? q - kTcFirstDone : -1};
}
struct TcChunkTab {
    TcChunk c[kTcChunks];
    constexpr TcChunkTab() : c{}
    {
        for (int q = 0; q < kTcChunks; q++) c[q] = make_chunk(q);
    }
};
static __constant__ TcChunkTab c_tc_chunks = TcChunkTab();

}  // namespace tc

// One block of 16 FIR columns of one channel: f on the channel's grid = (E + X) * fscale -> pieces f0 (integer grid) + f1, packed
// two rows per register for the A operand of MMA2, and the 16-term sums of the block's zero-state end state.  FIRST0: block 0 of
// tile 0, whose rows 0..14 are the carried y history (they do not drive the biquad) and whose row 15 is frame 0.
template <bool FIRST0>
__device__ __forceinline__ void ep_block(const uint32_t (&re)[16], const uint32_t (&rx)[16], uint32_t (&q0)[8], uint32_t (&q1)[8],
                                         const TcParams &p, const float *yh, float yh_scale, float &p0, float &p1)
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
            sc2 = pk2(yh_scale, i + 1 < kTcHr ? yh_scale : p.fscale);
        }
        float ra, rb;  // t * sc to the nearest integer (|.| < 2^22)
        upk2(add2(fma2(t2, sc2, pk2(12582912.f, 12582912.f)), pk2(-12582912.f, -12582912.f)), ra, rb);
        const __half2 h0 = __floats2half2_rn(ra, rb);  // above 2048 the fp16 grid is coarser than 1:
        const float2 f0 = __half22float2(h0);           // the remainder is taken from what h0 really holds
        float la, lb;
        upk2(fma2(t2, sc2, pk2(-f0.x, -f0.y)), la, lb);
        const __half2 h1 = __floats2half2_rn(la, lb);
        // Wz is pre-multiplied by fscale; the history rows of tile 0 do not drive the biquad
        const float w0a = (FIRST0 && i < kTcHr) ? 0.f : p.Wz[i][0], w1a = (FIRST0 && i < kTcHr) ? 0.f : p.Wz[i][1];
        const float w0b = (FIRST0 && i + 1 < kTcHr) ? 0.f : p.Wz[i + 1][0], w1b = (FIRST0 && i + 1 < kTcHr) ? 0.f : p.Wz[i + 1][1];
        z0 = fma2(pk2(w0a, w0b), t2, z0);
        z1 = fma2(pk2(w1a, w1b), t2, z1);
        q0[i >> 1] = h2_bits(h0);   // column i/2 of the block: rows i (low half) and i+1
        q1[i >> 1] = h2_bits(h1);
    }
    float a, b;
    upk2(z0, a, b);
    p0 += a + b;
    upk2(z1, a, b);
    p1 += a + b;
}

// Tail of a call whose last tile is partial (see the state code of the drain warps): out of line, so that its registers (two
// blocks of f pieces, the double recursion) do not count against the kernel's hot loops.
__device__ __noinline__ void tc_partial_tail(uint32_t taddr_a, uint32_t taddr_b, uint32_t taddr_state, double wi0, double wi1, double wi2,
                                             double wi3, double b0, double b1, double b2, double a1, double a2, double g_bq, float isig,
                                             int row0, int last_frames, float *yhist_next, double *bq_state_next, int C)
{
    using namespace tc;
    uint32_t fa[16], fb[16], ws[2];
    tmem_ld16(taddr_a, fa);
    tmem_ld16(taddr_b, fb);
    // the block's true state, as this warp has just written it to the mailbox (balanced coordinates) -> TDF-II basis
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(ws[0]), "=r"(ws[1]) : "r"(taddr_state));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const double w0d = (double)__uint_as_float(ws[0]), w1d = (double)__uint_as_float(ws[1]);
    double S0 = fma(wi0, w0d, wi1 * w1d), S1 = fma(wi2, w0d, wi3 * w1d);
    const double na1 = -a1, na2 = -a2;
    const int nrows = last_frames + kTcHr - row0;   // rows row0 .. last_frames + 14
#pragma unroll
    for (int k = 0; k < 32; k++) {
        if (k < nrows) {
            const uint32_t w0 = k < 16 ? fa[k >> 1] : fb[(k - 16) >> 1], w1 = k < 16 ? fa[8 + (k >> 1)] : fb[8 + ((k - 16) >> 1)];
            const __half2 h0 = *reinterpret_cast<const __half2 *>(&w0), h1 = *reinterpret_cast<const __half2 *>(&w1);
            const float fg = (k & 1) ? __high2float(h0) + __high2float(h1) : __low2float(h0) + __low2float(h1);
            const double x = (double)(fg * isig);
            const double v = fma(b0, x, S0);
            const double tt = fma(b1, x, S1);
            S0 = fma(na1, v, tt);
            S1 = fma(na2, v, b2 * x);
            const int hr = row0 + k - last_frames;   // 0..14: position in the carried history
            if (hr >= 0) yhist_next[(size_t)hr * C] = (float)(v * g_bq);
        }
    }
    bq_state_next[0] = S0;
    bq_state_next[1] = S1;
}

// PROF: 1 = per-role cycle counters (PB_TC_PROF=1; a compile-time switch: the counters cost the issue loops dearly, ~6 k cycles
// per tile), 2 = event timeline of CTA 0 only (PB_TC_PROF=2; nearly free)
// PARTIAL: the call's last tile may hold fewer than 160 frames (a separate instantiation: the few extra instructions in the output
// and drain warps cost a whole-tile batch 6 %, measured -- this kernel is that sensitive to its register allocation).
// DBG: the development switches of TcParams::dbg are compiled in (a separate instantiation, launched only when PB_TC_DBG is set: tested
// at run time they cost the converter ~15 instructions per chunk).
template <int PROF, bool PARTIAL = false, bool DBG = false>
__global__ void __launch_bounds__(kTcThreads, 1) chain_tc_kernel(const __grid_constant__ TcParams p)
{
    const int dbg = DBG ? p.dbg : 0;
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *raw = smem + kOffRaw;
    __half *tab = reinterpret_cast<__half *>(smem + kOffTab);
    const float *rcs = reinterpret_cast<const float *>(smem + kOffRc);
    float *zx = reinterpret_cast<float *>(smem + kOffZx);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *raw_full = bars, *raw_empty = bars + kRawStages;
    uint64_t *a1_full = bars + 2 * kRawStages, *a1_empty = a1_full + kA1Stages;  // converters -> MMA1 (tensor-memory ring) and back (commit)
    uint64_t *blk_full = a1_empty + kA1Stages;      // [10] MMA1 (commit after chunk 17+i) -> drain: column block i of D1 is final
    uint64_t *a2_ready = blk_full + kNumBlkBars;    // [11] drain warps -> MMA2: the f pieces of block b are in tensor memory
    uint64_t *slice_read = a2_ready + kTcBlocks;    // [5]  output warps -> MMA1: slice s has been multiplied (its f pieces are dead) AND read out of its D2 columns
    uint64_t *d2_full = slice_read + kRsSlices;     //      MMA2 (commit) -> output warps
    uint64_t *d2_empty = d2_full + 1;               //      output warps -> MMA2
    uint64_t *mbox_ready = d2_empty + 1;            //      drain warps A (all four arrive) -> output warps: block states in the TMEM mailbox
    uint64_t *mbox_free = mbox_ready + 1;           // [4]  output warps of quadrant e -> drain warp A of quadrant e
    uint64_t *zx_ready = mbox_free + 4;             // [4]  drain warp B -> drain warp A: the Z of the odd blocks are in shared memory
    uint64_t *szs_ready = zx_ready + 4;             // [4]  drain warp A -> drain warp B: zero-state block states and s10 are in the mailbox
    uint64_t *zx_free = szs_ready + 4;              // [4]  drain warp A -> drain warp B: the previous tile's Z have been read
    uint64_t *sa_ready = zx_free + 4;               // [4]  drain warp A -> drain warp B: the chained state of the even blocks 0..8 is in shared memory
    uint64_t *tab_ready = sa_ready + 4;             //      bulk copies of the B tables and the correction table -> both MMA issuers, output warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffTmemSlot);
    double *sa_s = reinterpret_cast<double *>(smem + kOffSa);

    // Role numbers (`warp` below) are not the hardware warp ids: a scheduler prefers its highest warp id, so the ids follow
    // the pipeline's priorities -- the three single-warp roles whose latency every tile waits for (TMA producer, the two MMA
    // issuers) take hardware warps 24-26, the converters (which feed MMA1, the longest stage) 16-23, the drain 8-15 and the
    // output warps, which have a whole tile of slack, 0-7.  (With the converters on the lowest ids one chunk of ~100
    // instructions took ~1100 cycles: every runnable drain / output warp went first.)  hw_q is the TMEM lane quadrant a warp
    // may touch.
    const int tid = threadIdx.x, hw_warp = tid >> 5, lane = tid & 31;
    const int warp = hw_warp < 8 ? hw_warp + 18 : hw_warp < 16 ? hw_warp + 2 : hw_warp < 24 ? hw_warp - 14 : hw_warp == 24 ? 0 : hw_warp == 25 ? 1 : 26;
    const int hw_q = hw_warp & 3;
    const int total_tiles = p.n_tiles * p.n_cg;

    // ---- pass B: verify the scales pass A speculated -----------------------------------
    // Every CTA scans all channels (the decision must be the same everywhere): peak in grid units outside [kSigLo, kSigHi]
    // -> the call is redone with the exact scales; in any case the exact scale goes to scale_next for the next call.
    const float *scale = p.scale;
    if (p.pass == 1) {
        int bad = 0;
        for (int c = tid; c < 2 * p.C; c += kTcThreads) {   // both scale classes of every channel
            const float pk = __uint_as_float(p.peak[c]), su = p.scale[c];
            const float units = pk * su;
            bad |= (units > kSigHi) || (pk > 0.f && units < kSigLo);
        }
        const int rerun = __syncthreads_or(bad);
        // the scales for the next call: written by CTA 0 alone in the steady state (148 CTAs storing the same 24 KB were most of
        // this launch's time); when the call is redone every CTA writes them (the same values) and reads its own copy back
        if (rerun || blockIdx.x == 0) {
            for (int c = tid; c < 2 * p.C; c += kTcThreads) {
                const float pk = __uint_as_float(p.peak[c]), su = p.scale[c];
                const float ns = pk > 0.f ? fminf(kSigTarget / pk, kSigCap) : su;
                p.scale_next[c] = ns;
                p.scale_next[2 * p.C + c] = 1.0f / ns;
                p.peak_next[c] = 0u;
            }
            if (rerun) __syncthreads();
        }
        if (blockIdx.x == 0 && p.meter_scratch) {
            // what pass A metered: fold it into the chain's meter, or drop it when the call is redone
            for (int c = tid; c < p.C; c += kTcThreads) {
                if (!rerun) {
                    p.meter_main[c] = fmax(p.meter_main[c], p.meter_scratch[c]);
                    p.meter_main[p.C + c] += p.meter_scratch[p.C + c];
                }
                p.meter_scratch[c] = 0.0;
                p.meter_scratch[p.C + c] = 0.0;
            }
        }
        if (!rerun) return;
        scale = p.scale_next;
    }
    const float *iscale = scale + 2 * p.C;

    // ---- one-time setup ------------------------------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int i = 0; i < kRawStages; i++) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], 1);   // the converter group's leader releases the stage once the group has read it (one arrival
                                           // instead of four: every mbarrier arrival of the CTA wakes every parked warp)
        }
        for (int i = 0; i < kA1Stages; i++) {
            mbar_init(&a1_full[i], 1);
            mbar_init(&a1_empty[i], 1);
        }
        for (int i = 0; i < kNumBlkBars; i++) mbar_init(&blk_full[i], 1);
        for (int i = 0; i < kTcBlocks; i++) mbar_init(&a2_ready[i], 1);
        for (int i = 0; i < kRsSlices; i++) mbar_init(&slice_read[i], 1);
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, 1);
        mbar_init(mbox_ready, 4);
        for (int i = 0; i < 4; i++) {
            mbar_init(&mbox_free[i], 2);
            mbar_init(&zx_ready[i], 1);
            mbar_init(&szs_ready[i], 1);
            mbar_init(&zx_free[i], 1);
            mbar_init(&sa_ready[i], 1);
        }
        mbar_init(tab_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
        // the tables arrive by bulk copy behind the CTA's back: the producer and the converters start on the first chunks at
        // once, only the two MMA issuers and the output warps wait for tab_ready (was: a copy loop of all threads in front of
        // the first __syncthreads)
        mbar_expect_tx(tab_ready, kTabBytes + kRcBytes);
        bulk_load(tab, p.tables, kTabBytes, tab_ready);
        bulk_load(smem + kOffRc, p.rc, kRcBytes, tab_ready);
    }
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
                    const long long c0 = (PROF == 1 ? clk() : 0ll);
                    mbar_wait(&raw_empty[s], ph ^ 1);
                    pw += (PROF == 1 ? clk() : 0ll) - c0;
                    if (dbg & 16) {  // development: no loads
                        mbar_arrive(&raw_full[s]);
                        if (++s == kRawStages) { s = 0; ph ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&raw_full[s], kRawStageBytes);
                    const int fr = f0 - kTcLead + 16 * q;  // first frame of the chunk, call-relative
                    // chunks never straddle frame 0 (kTcLead and tile starts are multiples of 16)
                    const CUtensorMap *map = (fr < 0) ? &p.tm_hist : &p.tm_in;
                    const int row = (fr < 0) ? fr + p.hist_rows : fr;  // history map holds frames [-hist_rows, 0); rows < 0 are zero-filled
                    tma_load_2d(raw + s * kRawStageBytes, map, ch0, row, &raw_full[s]);
                    if (++s == kRawStages) { s = 0; ph ^= 1; }
                }
            }
            if (PROF == 1 && p.prof) p.prof[blockIdx.x * kProfCount + kProfProdWait] = pw;
        }
    } else if (warp == 1) {
        // ================================ MMA1 issuer =================================
        // Only the band of the Toeplitz matrix is multiplied: chunk q (frames f0-272+16q ..+15) reaches output columns
        // [16q-257, 16q+14], i.e. 8-column blocks [2q-33, 2q+1] clipped to [0, 21] and widened to an even count (N % 16 == 0).
        // The warp runs the loop converged and elects one lane per issue.
        {
            mbar_wait(tab_ready, 0);
            const uint32_t t0 = smem_u32(tab);
            const uint64_t bd0 = make_desc(t0, 128, 128), bd1 = make_desc(t0 + TcTables::kT * 2, 128, 128);
            const uint64_t bd2 = make_desc(t0 + TcTables::kT * 4, 128, 128), bd3 = make_desc(t0 + TcTables::kT * 6, 128, 128);
            long long w_t = 0, w_c = 0, w_i = 0;
            const long long kstart = (PROF == 1 ? clk() : 0ll);
            unsigned long long ns0 = 0;
            if (PROF == 1) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
            const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int n_chunks = my_tiles * kTcChunks;
            // the MMAs of chunk q of a tile, A operand in ring stage st; called by the elected lane only
            auto issue_chunk = [&](int q, int st) {
                // Window of chunk q in 8-column blocks: [lo, hi], widened to an even count.  For q <= 10 its last two blocks
                // (2q, 2q+1) are touched for the first time in this tile: they are written with accumulate = 0 by a
                // separate N = 16 instruction, so D1 never needs zeroing and the previous tile only has to have
                // released column block q (16 columns) before chunk q -- MMA1 of the next tile overlaps that tile's tail.
                const TcChunk k = c_tc_chunks.c[q];
                const uint32_t a0 = tmem_base + kColA1 + 16 * st, a1 = a0 + 8;   // x0, x1 pieces of the chunk
                const uint64_t b0 = bd0 + k.toff, b1 = bd1 + k.toff, b2 = bd2 + k.toff, b3 = bd3 + k.toff;
                const uint32_t dE = tmem_base + kColE + k.dcol, dX = dE + (kColX - kColE);
                if (!(dbg & 1)) {
                    if (k.idn) {
                        umma_ts(dE, a0, b0, k.idn, 1);   // x0 h0: exact, integers < 2^24
                        umma_ts(dX, a0, b1, k.idn, 1);   // x0 (h1 + h2): the tap's remainder in two pieces
                    }
                    if (k.has_new) {
                        constexpr uint32_t id16 = make_idesc_c(16);
                        umma_ts(dE + k.ncol, a0, b0 + k.noff, id16, 0);
                        umma_ts(dX + k.ncol, a0, b1 + k.noff, id16, 0);
                    }
                    umma_ts(dX, a0, b2, k.idw, 1);
                    umma_ts(dX, a1, b3, k.idw, 1);       // x1 h: |x1| <= 1/2, the tap rounded once (2^-12 relative) is enough
                }
                umma_commit(&a1_empty[st]);  // frees the A stage when these MMAs have read it
                if (k.blk >= 0) umma_commit(&blk_full[k.blk]);  // column block q-17 (after 26: 9 and 10) is final
            };
            // The issuing warp is a serial chain of long-latency instructions (mbarrier test ~90 cycles, fence, elect, ~45 per
            // UTCHMMA issue, ~60 per commit): it takes TWO chunks -- the pair one converter group hands over together -- per
            // wait / elect round trip, otherwise the tensor pipe idles behind it (measured: 870 cycles per chunk for ~400 of MMA).
            int ita = 0, qa = 0;   // tile and chunk of ga, kept incrementally
            for (int ga = 0; ga < n_chunks; ga += 2) {
                const int gb = ga + 1 < n_chunks ? ga + 1 : ga;
                const int qb = gb == ga ? qa : (qa + 1 < kTcChunks ? qa + 1 : 0), itb = (gb != ga && qa + 1 == kTcChunks) ? ita + 1 : ita;
                long long c0 = (PROF == 1 ? clk() : 0ll);
                // First touch of column block q (q < 11): its E columns hold the previous tile's f pieces until that tile's MMA2
                // slice has read them, its X columns the resampler sums of slice q/2 (blocks 7..10: slice 4) until the output
                // warps have read them.  Slices are read in order, so the later chunk's slice covers the pair.
                if (qb < kTcBlocks && itb > 0) mbar_wait(&slice_read[qb < 7 ? (qb >> 1) : kRsSlices - 1], (itb - 1) & 1);
                else if (qa < kTcBlocks && ita > 0) mbar_wait(&slice_read[qa < 7 ? (qa >> 1) : kRsSlices - 1], (ita - 1) & 1);
                w_t += (PROF == 1 ? clk() : 0ll) - c0;
                c0 = (PROF == 1 ? clk() : 0ll);
                mbar_wait(&a1_full[gb % kA1Stages], (gb / kA1Stages) & 1);   // the group's leader arrives on ga's barrier first
                const long long c1 = (PROF == 1 ? clk() : 0ll);
                w_c += c1 - c0;
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (elect_one()) {
                    issue_chunk(qa, ga % kA1Stages);
                    if (gb != ga) issue_chunk(qb, gb % kA1Stages);
                }
                __syncwarp();
                const long long c2 = (PROF == 1 ? clk() : 0ll);
                w_i += c2 - c1;
                PB_TRACE(0, ita, qa);
                if (gb != ga) PB_TRACE(0, itb, qb);
                qa += 2;
                if (qa >= kTcChunks) { qa -= kTcChunks; ita++; }
            }
            if (PROF == 1 && p.prof && lane == 0) {
                long long *pr = p.prof + blockIdx.x * kProfCount;
                pr[kProfMmaWaitTmem] = w_t;
                pr[kProfMmaWaitCvt] = w_c;
                pr[kProfMmaIssue] = w_i;
                pr[kProfTotal] = (PROF == 1 ? clk() : 0ll) - kstart;
                unsigned long long ns1 = 0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
                pr[kProfNs] = (long long)(ns1 - ns0);
            }
        }
    } else if (warp < 10) {
        // ================================ converters ==================================
        // Two groups of 4 warps take alternate chunks; in a group the warp of TMEM lane quadrant e owns channels
        // [32 e, 32 e + 32) of the tile, one channel per thread: 16 frames -> x * sigma_c = x0 (integer grid) + x1 -> two
        // tcgen05.st of 8 columns.  The K-blocks are stored swapped (columns 0-3: frames 8-15, columns 4-7: frames 0-7): that
        // is what lets the Toeplitz B operand be addressed with LBO = SBO.
        const int grp = (warp - 2) >> 2, e = hw_q;
        const bool gl = ((warp - 2) & 3) == 0;  // first warp of the group (warps 2 / 6): the one that sleeps on mbarriers
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        const bool track = p.pass == 0;
        long long w_r = 0, w_c = 0, w_w = 0;
        const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int n_chunks = my_tiles * kTcChunks;
        int cur_it = -1, c = 0, f0 = 0, sc_idx = 0;
        bool last = false, early = false;
        float sig = 0.f, amax = 0.f;
        // one chunk (tile `it` of this CTA, chunk q of the tile, ring stage st): shared memory -> registers -> pieces -> tensor
        // memory (no synchronisation in here)
        auto convert = [&](int it, int q, int st) {
            if (it != cur_it) {   // once per tile
                if (track && cur_it >= 0) atomicMax(p.peak + sc_idx, __float_as_uint(amax));
                cur_it = it;
                const int tile = blockIdx.x + it * gridDim.x;
                const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
                f0 = t * kTcFrames;
                last = (t == p.n_tiles - 1);
                c = cg * kTcCh + e * 32 + lane;
                sc_idx = tc_scale_class(t) * p.C + c;
                // the first two tiles of class 1 look back at frames that are "new" in tiles of class 0: they measure their whole window
                early = (t >= kTcHistTiles && t < 2 * kTcHistTiles);
                sig = scale[sc_idx];
                amax = 0.f;
            }
            const bool hist = (f0 - kTcLead + 16 * q) < 0;
            const float gsc = hist ? 1.f : p.g_load;   // history frames are already gain-scaled
            const float sc = gsc * sig;
            const float *src = reinterpret_cast<const float *>(raw + st * kRawStageBytes) + e * 32 + lane;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = (dbg & 4) ? 0.f : src[i * kTcCh];
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                // x * scale rounded to the nearest integer on the FMA pipe (|.| < 2^22), and the remainder with a single
                // rounding; two values per instruction
                const f32x2 x2 = pk2(v[i], v[i + 1]), s2 = pk2(sc, sc);
                const f32x2 r2 = add2(fma2(x2, s2, pk2(12582912.f, 12582912.f)), pk2(-12582912.f, -12582912.f));
                float ra, rb, la, lb;
                upk2(r2, ra, rb);
                upk2(fma2(x2, s2, pk2(-ra, -rb)), la, lb);
                const int col = (i < 8) ? 4 + (i >> 1) : ((i - 8) >> 1);  // K-blocks swapped (Toeplitz trick)
                hi[col] = h2_bits(__floats2half2_rn(ra, rb));
                lo[col] = h2_bits(__floats2half2_rn(la, lb));
            }
            // peak of |g x| (true units) for the scale check: every frame of the call is "new" (chunks 17..26) in exactly one tile;
            // the history frames are seen by the first tiles only
            if (track && (q >= kTcFirstDone || hist || early)) {
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i += 2) m = fmaxf(m, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
                amax = fmaxf(amax, m * fabsf(gsc));
            }
            if (!(dbg & 32)) tmem_st16(tmem_base + lane_base + kColA1 + 16 * st, hi, lo);   // x0: 8 columns, x1: 8 columns
            if (last) {
                // carried FIR input history: frames [n-hist_rows, n) in gain-scaled units (K1's convention)
                // (the window ends kTcFrames - last_frames rows behind the call's last frame when the last tile is partial)
                const int hrow0 = 16 * q - (kTcLead + (PARTIAL ? p.last_frames : kTcFrames) - p.hist_rows);
                if (hrow0 + 15 >= 0 && hrow0 < p.hist_rows) {
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (hrow0 + i >= 0 && hrow0 + i < p.hist_rows) p.xhist_next[(size_t)(hrow0 + i) * p.C + c] = v[i] * gsc;
                }
            }
        };
        static_assert(kRawStages == 8 && kA1Stages == 8, "the converter uses one stage index (g & 7) for both rings");
        // The loop body is a chain of latencies (mbarrier wake-up, named barrier, LDS, tcgen05.st + wait, named barrier,
        // arrive: ~1 k cycles around ~70 instructions of arithmetic), so a group takes TWO consecutive chunks per round trip
        // (the pair may straddle two tiles: 27 is odd).
        int ita = 0, qa = 2 * grp;   // tile and chunk of ga, kept incrementally (ga advances by 4)
        for (int ga = 2 * grp; ga < n_chunks; ga += 4) {
            const int gb = ga + 1;
            const bool two = gb < n_chunks;
            const int qb = qa + 1 < kTcChunks ? qa + 1 : 0, itb = qa + 1 < kTcChunks ? ita : ita + 1;
            const int sa = ga & 7, sb = gb & 7;
            const uint32_t pha = (ga >> 3) & 1, phb = (gb >> 3) & 1;
            const long long c0 = (PROF == 1 ? clk() : 0ll);
            if (gl) {
                mbar_wait(&raw_full[sa], pha);
                if (two) mbar_wait(&raw_full[sb], phb);
            }
            const long long c1 = (PROF == 1 ? clk() : 0ll);
            if (gl) {
                mbar_wait(&a1_empty[sa], pha ^ 1);
                if (two) mbar_wait(&a1_empty[sb], phb ^ 1);
            }
            group_sync(1 + grp);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const long long c2 = (PROF == 1 ? clk() : 0ll);
            w_r += c1 - c0;
            w_c += c2 - c1;
            convert(ita, qa, sa);
            if (two) convert(itb, qb, sb);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
            group_sync(1 + grp);
            if (gl && lane == 0) {
                mbar_arrive(&a1_full[sa]);
                if (two) mbar_arrive(&a1_full[sb]);
                mbar_arrive(&raw_empty[sa]);   // (the group has read both chunks out of shared memory: the barrier above)
                if (two) mbar_arrive(&raw_empty[sb]);
            }
            if (gl) PB_TRACE(1 + grp, ita, qa);
            w_w += (PROF == 1 ? clk() : 0ll) - c2;
            qa += 4;
            if (qa >= kTcChunks) { qa -= kTcChunks; ita++; }
        }
        if (track && cur_it >= 0) atomicMax(p.peak + sc_idx, __float_as_uint(amax));
        if (PROF == 1 && p.prof && warp == 2 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfCvtWaitRaw] = w_r;
            pr[kProfCvtWaitCvt] = w_c;
            pr[kProfCvtWork] = w_w;
        }
    } else if (warp < 18) {
        // ================================ FIR drain warps =============================
        // Two warps per TMEM lane quadrant e (channels cg*128 + 32e + lane): role A (warps 10-13) takes the even
        // blocks of 16 columns, role B (warps 14-17) the odd ones.  Per block: f on the channel's grid = (E + X) * fscale ->
        // fp16 pieces f0 (integer grid) + f1 written back over the block's E columns (the A operand of MMA2), and the zero-state
        // end state of the block Z_b (16-term float sums) into shared memory.  Role A then runs the block-state recursion in
        // double and the look-back.  Everything between the accumulators and the block states stays on the channel's grid (the
        // chain is linear); 1 / sigma_c is applied where states leave for the look-back arrays and the mailbox.
        const int e = hw_q;
        const bool roleB = warp >= 14;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
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
            const float sig = scale[tc_scale_class(t) * p.C + c], isig = iscale[tc_scale_class(t) * p.C + c];
            const bool gl = (warp & 3) == 2;  // first warp of the role group (warps 10 / 14)
            const int gid = roleB ? 4 : 3;
            float *zxt = zx;
            // Each group chains the zero-state end states of ITS blocks on the way (two block steps per link): the state after row
            // 159 from a zero state -- the tile's aggregate, which the look-back of the tiles behind waits for -- is then
            // A^16 (even blocks 0..8) + (odd blocks 1..9), known the moment block 9 has been drained instead of a full 11-step
            // recursion (~2.5 k cycles) later.
            double ps1 = 0.0, ps2 = 0.0;
            if (roleB && it > 0) mbar_wait(&zx_free[e], par ^ 1);  // role A has read the previous tile's block sums
#pragma unroll 1
            for (int b = roleB ? 1 : 0; b < kTcBlocks; b += 2) {
                const long long k0 = (PROF == 1 ? clk() : 0ll);
                if (gl) mbar_wait(&blk_full[b < kNumBlkBars ? b : kNumBlkBars - 1], par);
                group_sync(gid);
                const long long k1 = (PROF == 1 ? clk() : 0ll);
                e_w += k1 - k0;
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t re[16], rx[16];
                tmem_ld16(tmem_base + lane_base + kColE + 16 * b, re);
                tmem_ld16(tmem_base + lane_base + kColX + 16 * b, rx);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float p0 = 0.f, p1 = 0.f;
                uint32_t q0[8], q1[8];
                if (first && b == 0) ep_block<true>(re, rx, q0, q1, p, yh, sig * p.inv_gbq, p0, p1);
                else ep_block<false>(re, rx, q0, q1, p, yh, 0.f, p0, p1);
                tmem_st8(tmem_base + lane_base + kColE + 16 * b, q0);
                tmem_st8(tmem_base + lane_base + kColE + 16 * b + 8, q1);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                group_sync(gid);
                if (gl && lane == 0) mbar_arrive(&a2_ready[b]);
                zxt[(2 * b + 0) * kTcCh + e * 32 + lane] = p0;
                zxt[(2 * b + 1) * kTcCh + e * 32 + lane] = p1;
                if (b < kTcBlocks - 1) {   // blocks 0..9 make the state after row 159
                    const double n1 = fma(p.A32[0], ps1, fma(p.A32[1], ps2, f2d_bits(p0)));
                    const double n2 = fma(p.A32[2], ps1, fma(p.A32[3], ps2, f2d_bits(p1)));
                    ps1 = n1;
                    ps2 = n2;
                }
                if (!roleB && b == kTcBlocks - 3) {   // after block 8: hand the even chain to group B
                    *reinterpret_cast<double2 *>(sa_s + ((size_t)(it & 1) * kTcCh + e * 32 + lane) * 2) = make_double2(ps1, ps2);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sa_ready[e]);
                }
                if (gl) PB_TRACE(roleB ? 4 : 3, it, b);
                e_m += (PROF == 1 ? clk() : 0ll) - k1;
            }
            const long long k4 = (PROF == 1 ? clk() : 0ll);
            // The tail of the tile is split between the two roles so that neither is a long serial chain:
            //   role A: block-state recursion from a zero state, the tile's aggregate for the look-back of the tiles behind it,
            //           zero-state block states -> mailbox (tensor memory);
            //   role B: (meanwhile) the decoupled look-back for the tile's incoming state -- it depends on the predecessors
            //           only --, then mailbox += response to that state, the tile's inclusive state, the carried state.
            if (!roleB) {
                // ---- role A: s(b+1) = A^16 s(b) + Z_b in double, on the channel's grid
                mbar_wait(&zx_ready[e], par);
                float szs[kTcBlocks][2];          // zero-state state at the start of each block (true units)
                {
                    double s1 = 0.0, s2 = 0.0;
#pragma unroll
                    for (int b = 0; b < kTcBlocks; b++) {
                        // the whole block recursion runs in balanced coordinates w = W s (W = S V^T of the block's free-response
                        // matrix): in the TDF-II basis a filter with poles near z = 1 has free responses that cancel to 1e-2 of
                        // their terms, which neither the float sums Z_b nor the float states can afford (measured 6e-6 on a 200 Hz
                        // high-pass)
                        szs[b][0] = d2f_bits(s1) * isig;
                        szs[b][1] = d2f_bits(s2) * isig;
                        const double n1 = fma(p.A16[0], s1, fma(p.A16[1], s2, f2d_bits(zxt[(2 * b + 0) * kTcCh + e * 32 + lane])));
                        const double n2 = fma(p.A16[2], s1, fma(p.A16[3], s2, f2d_bits(zxt[(2 * b + 1) * kTcCh + e * 32 + lane])));
                        s1 = n1;
                        s2 = n2;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&zx_free[e]);
                if (it > 0) mbar_wait(&mbox_free[e], par ^ 1);  // the output warps have read the previous tile's block states
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int b = 0; b < kTcBlocks; b++)
                    tmem_st2(tmem_base + lane_base + kColMbox + 2 * b, __float_as_uint(szs[b][0]), __float_as_uint(szs[b][1]));
                tmem_st2(tmem_base + lane_base + kColMbox + 2 * kTcBlocks, 0u, 0u);  // the last slice reads a fourth, absent block
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(&szs_ready[e]);
                if (gl) PB_TRACE(3, it, 12);
                e_l += (PROF == 1 ? clk() : 0ll) - k4;
                continue;
            }
            // ---- role B
            __syncwarp();
            if (lane == 0) mbar_arrive(&zx_ready[e]);
            // the tile's aggregate: state after row 159 from a zero state = A^16 (even chain, group A) + (odd chain), back in the
            // TDF-II basis and in true units, which the look-back arrays and K1 use
            double s10_1, s10_2;
            {
                mbar_wait(&sa_ready[e], par);
                const double2 sa = *reinterpret_cast<const double2 *>(sa_s + ((size_t)(it & 1) * kTcCh + e * 32 + lane) * 2);
                const double t1 = fma(p.A16[0], sa.x, fma(p.A16[1], sa.y, ps1)), t2 = fma(p.A16[2], sa.x, fma(p.A16[3], sa.y, ps2));
                const double isig_d = f2d_bits(isig);
                s10_1 = fma(p.Wbi[0], t1, p.Wbi[1] * t2) * isig_d;
                s10_2 = fma(p.Wbi[2], t1, p.Wbi[3] * t2) * isig_d;
                if (chained) {
                    p.lb_agg[slot * 64 + lane * 2] = s10_1;
                    p.lb_agg[slot * 64 + lane * 2 + 1] = s10_2;
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
                }
            }
            // incoming state: decoupled look-back (same protocol and arrays as K1, 32-channel groups)
            double q1 = 0.0, q2 = 0.0;
            if (first) {
                q1 = p.bq_state[2 * c];
                q2 = p.bq_state[2 * c + 1];
            } else if (dbg & 64) {   // development: no look-back (wrong results)
            } else {
                const int base = t - 1, j = base - lane;
                int first_inc = 0;
                unsigned lb_ns = 32;   // poll with back-off: every poll is an L2 round trip plus an L1 invalidate (ld.acquire)
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
                    if (spins > (1u << 22)) {
                        if (lane == 0) atomicExch(p.err_flag, 1);
                        first_inc = -1;
                        break;
                    }
                    __nanosleep(lb_ns);
                    lb_ns = lb_ns < 128 ? lb_ns * 2 : 128;
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
            if (gl) PB_TRACE(4, it, 12);   // look-back done
            // ---- true state at the start of every block = zero-state part (role A, in the mailbox) + response to the incoming state
            const int end_blk = (PARTIAL && last && p.last_frames < kTcFrames) ? p.last_frames >> 4 : -1;   // block of tile row `last_frames`
            {
                const float qf0 = d2f_bits(fma(p.Wb[0], q1, p.Wb[1] * q2)), qf1 = d2f_bits(fma(p.Wb[2], q1, p.Wb[3] * q2));
                const int fi = first ? 1 : 0;
                mbar_wait(&szs_ready[e], par);
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t u0[8], u1[8], u2[8];
                tmem_ld8(tmem_base + lane_base + kColMbox, u0);
                tmem_ld8(tmem_base + lane_base + kColMbox + 8, u1);
                tmem_ld8(tmem_base + lane_base + kColMbox + 16, u2);   // blocks 8..10, the zero pad
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int b = 0; b < kTcBlocks; b++) {
                    const float *M = p.Mb[fi][b];
                    const uint32_t z0 = b < 4 ? u0[2 * b] : b < 8 ? u1[2 * b - 8] : u2[2 * b - 16];
                    const uint32_t z1 = b < 4 ? u0[2 * b + 1] : b < 8 ? u1[2 * b - 7] : u2[2 * b - 15];
                    const float v0 = fmaf(M[0], qf0, fmaf(M[1], qf1, __uint_as_float(z0))), v1 = fmaf(M[2], qf0, fmaf(M[3], qf1, __uint_as_float(z1)));
                    tmem_st2(tmem_base + lane_base + kColMbox + 2 * b, __float_as_uint(v0), __float_as_uint(v1));
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(mbox_ready);
                if (gl) PB_TRACE(4, it, 13);
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
            } else if (PARTIAL && end_blk >= 0) {
                // Partial last tile: the call ends after tile row 14 + last_frames (row = frame + 15).  Carried biquad state (after
                // that row) and carried y history (its last 15 rows) by running the recursion row by row from the true state at
                // the start of the block that holds row `last_frames` -- a float in balanced coordinates, like every block state
                // (6e-8 of the state; the full-tile case below starts from the double inclusive state) -- over the pieces of f in
                // that block's and the next block's E columns: at most 30 rows.
                tc_partial_tail(tmem_base + lane_base + kColE + 16 * end_blk,
                                tmem_base + lane_base + kColE + 16 * (end_blk + 1 < kTcBlocks ? end_blk + 1 : end_blk),
                                tmem_base + lane_base + kColMbox + 2 * end_blk, p.Wbi[0], p.Wbi[1], p.Wbi[2], p.Wbi[3],
                                p.b0, p.b1, p.b2, p.a1, p.a2, p.g_bq, isig, 16 * end_blk, p.last_frames, p.yhist_next + c,
                                p.bq_state_next + 2 * c, p.C);
            } else {
                // carried biquad state (after row 174) and carried y history (rows 160..174): the only place where the
                // recursion runs row by row, from the pieces of f that role A wrote over block 10's E columns
                uint32_t fp[16];
                tmem_ld16(tmem_base + lane_base + kColE + 16 * (kTcBlocks - 1), fp);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const double nb1 = p.b1, nb2 = p.b2, na1 = -p.a1, na2 = -p.a2;
#pragma unroll
                for (int r = 0; r < kTcHr; r++) {
                    const __half2 h0 = *reinterpret_cast<const __half2 *>(&fp[r >> 1]), h1 = *reinterpret_cast<const __half2 *>(&fp[8 + (r >> 1)]);
                    const float fg = (r & 1) ? __high2float(h0) + __high2float(h1) : __low2float(h0) + __low2float(h1);
                    const double x = (double)(fg * isig);
                    const double v = fma(p.b0, x, I0);
                    const double tt = fma(nb1, x, I1);
                    I0 = fma(na1, v, tt);
                    I1 = fma(na2, v, nb2 * x);
                    p.yhist_next[(size_t)r * p.C + c] = (float)(v * p.g_bq);
                }
                p.bq_state_next[2 * c] = I0;
                p.bq_state_next[2 * c + 1] = I1;
            }
            e_l += (PROF == 1 ? clk() : 0ll) - k4;
        }
        if (PROF == 1 && p.prof && warp == 10 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfEpWaitBlk] = e_w;
            pr[kProfEpWork] = e_m;
            pr[kProfEpLookback] = e_l;
        }
    } else if (warp < 26) {
        // ================================ output warps ================================
        // two warps per TMEM lane quadrant: warps 18-21 take outputs 0..15 of every slice, warps 22-25 outputs 16..31.
        // Phase 1, per slice as MMA2 completes it: D2 -> registers, free the (single) D2 buffer at once, park E2 + X2 in shared
        // memory -- nothing here waits for the look-back, so MMA2's slices flow and release the accumulator blocks the next
        // tile's MMA1 is waiting for.  Phase 2, once the drain warps have put the block states into the mailbox: parked sum ->
        // descale + block-state correction -> coalesced stores, meter.  Every thread parks and re-reads its own values only
        // (row-major [output][channel]: conflict-free), so the park buffer needs no synchronisation of its own.
        const int e = hw_q;
        const int hsel = warp >= 22 ? 1 : 0;
        const uint32_t lane_base = (uint32_t)(e * 32) << 16;
        float *park = reinterpret_cast<float *>(smem + kOffPark) + 16 * hsel * kTcCh + e * 32 + lane;
        long long r_w = 0, r_m = 0, o_ld = 0, o_out = 0;
        int it = 0;
        unsigned nsl = 0;  // running slice number: phase of the D2 barriers
        mbar_wait(tab_ready, 0);   // the correction table (rcs)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int t = tile / p.n_cg, cg = tile - t * p.n_cg;
            const bool first = (t == 0), part_tile = PARTIAL && (t == p.n_tiles - 1) && p.last_frames < kTcFrames;
            const int c = cg * kTcCh + e * 32 + lane;
            const uint32_t par = it & 1;
            float *outp = p.out + (size_t)t * kTcOut * p.C + c;
            const float dsc = p.descale_rs * iscale[tc_scale_class(t) * p.C + c];
            float m_peak = 0.f;
            double m_sumsq = 0.0;
            const bool meter = p.meter_peak != nullptr;
            const long long k5 = (PROF == 1 ? clk() : 0ll);
            float hold[16];   // the last slice's sums: they wait in registers (nothing queues behind the last slice)
#pragma unroll 1
            for (int s = 0; s < kRsSlices; s++, nsl++) {
                const long long k6 = (PROF == 1 ? clk() : 0ll);
                if (warp == 18) mbar_wait(d2_full, nsl & 1u);
                asm volatile("bar.sync 5, 256;" ::: "memory");
                const long long k7 = (PROF == 1 ? clk() : 0ll);
                r_w += k7 - k6;
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t e16[16], x16[16];
                tmem_ld16(tmem_base + lane_base + d2_col(s) + 16 * hsel, e16);
                tmem_ld16(tmem_base + lane_base + d2_col(s) + kRsN + 16 * hsel, x16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                asm volatile("bar.sync 5, 256;" ::: "memory");
                if (warp == 18 && lane == 0) {
                    mbar_arrive(d2_empty);
                    mbar_arrive(&slice_read[s]);
                }
                if (s < kRsSlices - 1) {
                    float *pp = park + (size_t)(kRsN * s) * kTcCh;
#pragma unroll
                    for (int i = 0; i < 16; i++) pp[i * kTcCh] = __uint_as_float(e16[i]) + __uint_as_float(x16[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) hold[i] = __uint_as_float(e16[i]) + __uint_as_float(x16[i]);
                }
                if (warp == 18) PB_TRACE(6, it, s);
                o_ld += (PROF == 1 ? clk() : 0ll) - k7;
            }
            const long long k8 = (PROF == 1 ? clk() : 0ll);
            if (warp == 18) mbar_wait(mbox_ready, par);  // the block states of all four quadrants are in the mailbox
            asm volatile("bar.sync 5, 256;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;");
            const long long k9 = (PROF == 1 ? clk() : 0ll);
            r_w += k9 - k8;
            if (warp == 18) PB_TRACE(6, it, 8);
            // 16 outputs of slice s from their sums: descale, block-state correction, store, meter
            auto emit = [&](int s, auto sum_of, auto partial) {
                // true state at the start of the four blocks this slice reads
                uint32_t zs[8];
                tmem_ld8(tmem_base + lane_base + kColMbox + 4 * s, zs);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (s == kRsSlices - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&mbox_free[e]);
                }
                if (dbg & 128) return;   // development: no output arithmetic, no stores (wrong results)
                const f32x2 sb01 = pk2(__uint_as_float(zs[0]), __uint_as_float(zs[1])), sb23 = pk2(__uint_as_float(zs[2]), __uint_as_float(zs[3]));
                const f32x2 sb45 = pk2(__uint_as_float(zs[4]), __uint_as_float(zs[5])), sb67 = pk2(__uint_as_float(zs[6]), __uint_as_float(zs[7]));
                // (a partial last tile stores only the outputs its frames have triggered)
                // (PARTIAL: the call's last tile when it is not full -- it stores only the outputs its frames have triggered; a
                // separate instantiation, so that the loop of every other tile keeps its two possible trip counts)
                const int nout = (decltype(partial)::value ? p.last_outputs - kRsN * s : (s == kRsSlices - 1) ? kTcOut - kRsN * (kRsSlices - 1) : kRsN) - 16 * hsel;
                const float *rcg = rcs + (size_t)(((first && s == 0) ? kTcRcFirst : kRsN * s) + 16 * hsel) * 8;
                float *op = outp + (size_t)(kRsN * s + 16 * hsel) * p.C;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float4 ca = *reinterpret_cast<const float4 *>(rcg + i * 8);
                    const float4 cb = *reinterpret_cast<const float4 *>(rcg + i * 8 + 4);
                    f32x2 acc = mul2(pk2(ca.x, ca.y), sb01);
                    acc = fma2(pk2(ca.z, ca.w), sb23, acc);
                    acc = fma2(pk2(cb.x, cb.y), sb45, acc);
                    acc = fma2(pk2(cb.z, cb.w), sb67, acc);
                    float c0, c1;
                    upk2(acc, c0, c1);
                    const float o = fmaf(sum_of(i), dsc, c0 + c1);
                    if (i < nout) {
                        op[(size_t)i * p.C] = o;
                        if (meter) {
                            m_peak = fmaxf(m_peak, fabsf(o));
                            m_sumsq += (double)o * (double)o;
                        }
                    }
                }
            };
#pragma unroll 1
            for (int s = 0; s < kRsSlices - 1; s++) {
                const float *pp = park + (size_t)(kRsN * s) * kTcCh;
                if (PARTIAL && part_tile) emit(s, [&](int i) { return pp[i * kTcCh]; }, std::integral_constant<bool, PARTIAL>{});
                else emit(s, [&](int i) { return pp[i * kTcCh]; }, std::false_type{});
            }
            if (PARTIAL && part_tile) emit(kRsSlices - 1, [&](int i) { return hold[i]; }, std::integral_constant<bool, PARTIAL>{});
            else emit(kRsSlices - 1, [&](int i) { return hold[i]; }, std::false_type{});
            o_out += (PROF == 1 ? clk() : 0ll) - k9;
            if (warp == 18) PB_TRACE(6, it, 9);
            r_m += (PROF == 1 ? clk() : 0ll) - k5;
            if (meter) {
                atomic_max_nonneg(p.meter_peak + c, (double)m_peak);
                atomicAdd(p.meter_sumsq + c, m_sumsq);
            }
        }
        if (PROF == 1 && p.prof && warp == 18 && lane == 0) {
            long long *pr = p.prof + blockIdx.x * kProfCount;
            pr[kProfOutWait] = r_w;
            pr[kProfOutMain] = r_m;
            pr[kProfOutLd] = o_ld;
            pr[kProfOutMath] = o_out;
        }
    } else {
        // ================================ MMA2 issuer =================================
        {
            mbar_wait(tab_ready, 0);
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
                    const int nch = (s == kRsSlices - 1) ? 3 : 4;
                    const long long c0 = (PROF == 1 ? clk() : 0ll);
                    mbar_wait(&a2_ready[2 * s + nch - 1], it & 1);  // each drain role stages its blocks (even / odd) in order
                    mbar_wait(&a2_ready[2 * s + nch - 2], it & 1);
                    w_y += (PROF == 1 ? clk() : 0ll) - c0;
                    mbar_wait(d2_empty, (nsl & 1u) ^ 1u);           // the output warps have read the previous slice (its D2 columns overlap this one's)
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t dE = tmem_base + d2_col(s), dX = dE + kRsN;
#pragma unroll 1
                    for (int k = 0; k < nch; k++, pair++) {
                        const uint32_t f0 = tmem_base + kColE + 16 * (2 * s + k), f1 = f0 + 8;   // the pieces of block 2s+k
                        const uint32_t tb = b2 + ((first && pair == 0) ? kRsPairs : pair) * 2048;
                        const uint64_t r0 = make_desc(tb, 128, 256), r1 = make_desc(tb + 1024, 128, 256);
                        if (!(dbg & 2) && elect_one()) {
                            umma_ts(dE, f0, r0, idesc64, k > 0);  // f0 * [p0 | p1] -> [E2 | X2]; E2 exact: integers < 2^24
                            umma_ts(dX, f1, r0, idesc32, 1);
                            umma_ts(dX, f1, r1, idesc32, 1);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(d2_full);
                    }
                    __syncwarp();
                    PB_TRACE(5, it, s);
                }
            }
            if (PROF == 1 && p.prof && lane == 0) p.prof[blockIdx.x * kProfCount + kProfMma2Wait] = w_y;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

#endif  // __CUDACC__

}  // namespace pb
