// chain_tile.cuh -- K1, the generic fused Processor-run kernel.
//
// One launch walks one contiguous run of Processors
//     [gain*] [FIR]? [biquad]? [resample]?      (each optional, gains folded)
// over a whole batch of buffers.  It replaces the per-buffer ProcessFunc walk of
// Processor.execute (reference pipe.go:425-451, the call at :438) for that run.
//
// Work decomposition: the batch is cut into tiles of
// 32 channels x L frames.  Persistent CTAs take tiles from a ticket counter in
// time-major order, so a tile only ever waits on tiles with smaller tickets,
// which are already resident: forward progress without a grid barrier.
//
//   * FIR:      tile + (T-1)-frame halo of the input, re-read through L2.
//   * biquad:   time is split inside the tile (8 sub-chunks) AND across tiles.
//               Each tile publishes its zero-state end state Z (aggregate), then
//               finds its incoming state with a decoupled look-back over its
//               predecessors' aggregates:  S_in = sum_i (A^L)^i Z_{t-1-i} + ...
//               (a first-order linear recurrence on the 2-vector TDF-II state).
//   * resample: polyphase with the integer phase accumulator; the P-1 frames of
//               left context are recomputed by the tile (tiles t > 0) or come
//               from the carried y-history (tile 0).
//
// Carried state between calls (the reference guarantees in-order one-at-a-time
// delivery per stage: fitting.go:58, run.go:175): FIR input history (T-1
// frames), biquad state at the call end, resampler input history (P-1 frames)
// and phase accumulator.  All are ping-ponged so a launch never reads what it
// writes.
#pragma once

#include "common.cuh"

namespace pb {

constexpr int kCg = 32;           // channels per tile == warp width (128 B rows in f32)
constexpr int kTileThreads = 256;
constexpr int kNW = kTileThreads / 32;

enum : unsigned { kLbNone = 0u, kLbAgg = 1u, kLbInc = 2u };

template <typename T>
struct TileParams {
    const T *in;          // [n_frames][C]
    T *out;               // [out_frames][C]
    int64_t n_frames;
    int C, L, n_tiles, n_groups;
    // gains at the four points of the run: at load, after FIR, after biquad, after resample
    // (the host folds a gain into the previous point when the stage it follows is absent)
    T g_load, g_fir, g_bq, g_out;
    // FIR
    int Hf;               // taps-1, 0 when absent
    int has_fir;
    const T *taps_padded; // [tp_len]: taps[k] at index k + tp_off, zeros elsewhere
    int tp_len, tp_off;
    const T *xhist;       // [Hf][C] gain-scaled input frames before this call
    T *xhist_next;
    // biquad
    int has_bq;
    // The recursion runs in double whatever T is: an fp32 TDF-II with poles near z = 1 has a
    // rounding-noise gain far above the 1e-6 parity bar, and 7 DFMA/sample is cheap next to the FIR.
    double b0, b1, b2, a1, a2;
    const double *bq_wt;     // [wt_len][2]  Wt[k] = A^k B
    const double *bq_apow;   // [L+Hr+1][4]  A^k row-major
    int wt_len;
    const double *bq_state;  // [C][2] state at the first frame of this call
    double *bq_state_next;   // [C][2] state after the last frame of this call
    double *lb_agg;          // [n_groups][n_tiles][32][2]
    double *lb_inc;
    unsigned *lb_status;  // [n_groups][n_tiles]
    unsigned epoch;
    // resample
    int has_rs, rs_up, rs_down, rs_P, Hr;
    int64_t rs_acc0;
    const T *rs_coef;     // [up][P]: coef[br][k] = proto[br + k*up]
    const T *yhist;       // [Hr][C] resampler input frames before this call
    T *yhist_next;
    // fused meter sink
    double *meter_peak;   // [C] or nullptr
    double *meter_sumsq;
    // scheduling
    unsigned long long *ticket;
    unsigned long long ticket_base;
    int *err_flag;        // set to 1 if a look-back wait exceeds its bound (never expected; avoids a hung GPU)
    int vec_ok;           // float4 path usable (f32, C%4==0, 16B-aligned pointers)
};

__device__ __forceinline__ void mat2_apply(const double *__restrict__ m, double &v0, double &v1)
{
    const double r0 = m[0] * v0 + m[1] * v1;
    const double r1 = m[2] * v0 + m[3] * v1;
    v0 = r0;
    v1 = r1;
}

// host+device: byte layout of the dynamic shared memory
template <typename T>
struct TileSmem {
    int rowsA_alloc, rowsB_alloc;
    size_t off_tp, off_wt, off_zq, off_sin, off_red, off_bufA, off_bufB, total;
    __host__ __device__ TileSmem(int L, int Hf, int Hr, int has_fir, int tp_len, int wt_len, int FB)
    {
        rowsB_alloc = L + Hr + FB;               // partial FIR blocks write nothing past rowsB but index math stays in range
        rowsA_alloc = L + Hr + Hf + 2 * FB;      // FIR chunking reads up to 2*FB-1 zeroed rows past the tile
        size_t o = 0;
        off_wt = o;   o += sizeof(double) * (size_t)((2 * wt_len + 1) & ~1);
        off_zq = o;   o += sizeof(double) * kNW * kCg * 2;
        off_sin = o;  o += sizeof(double) * kNW * kCg * 2;
        off_red = o;  o += sizeof(double) * kNW * kCg * 2;
        off_tp = o;   o += sizeof(T) * (size_t)((tp_len + 3) & ~3);
        o = (o + 15) & ~(size_t)15;
        off_bufA = o; o += sizeof(T) * (size_t)rowsA_alloc * kCg;
        off_bufB = o; o += has_fir ? sizeof(T) * (size_t)rowsB_alloc * kCg : 0;
        total = o;
    }
};

template <typename T, int FB>
__global__ void __launch_bounds__(kTileThreads) chain_tile_kernel(const TileParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_tile;

    const TileSmem<T> lay(p.L, p.Hf, p.Hr, p.has_fir, p.tp_len, p.wt_len, FB);
    T *tp_s = reinterpret_cast<T *>(smem_raw + lay.off_tp);
    double *wt_s = reinterpret_cast<double *>(smem_raw + lay.off_wt);
    double *zq_s = reinterpret_cast<double *>(smem_raw + lay.off_zq);
    double *sin_s = reinterpret_cast<double *>(smem_raw + lay.off_sin);
    double *red_s = reinterpret_cast<double *>(smem_raw + lay.off_red);
    T *bufA = reinterpret_cast<T *>(smem_raw + lay.off_bufA);
    T *bufB = p.has_fir ? reinterpret_cast<T *>(smem_raw + lay.off_bufB) : bufA;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, L = p.L, Hf = p.Hf, Hr = p.Hr;
    const int total_tiles = p.n_tiles * p.n_groups;

    // per-launch constants into shared memory
    for (int i = tid; i < p.tp_len; i += kTileThreads) tp_s[i] = p.has_fir ? p.taps_padded[i] : T(0);
    for (int i = tid; i < 2 * p.wt_len; i += kTileThreads) wt_s[i] = p.has_bq ? p.bq_wt[i] : 0.0;

    for (;;) {
        __syncthreads();  // previous tile fully done with shared memory (and constants visible)
        if (tid == 0) s_tile = (int)(atomicAdd(p.ticket, 1ULL) - p.ticket_base);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= total_tiles) break;
        const int t = tile / p.n_groups, g = tile - t * p.n_groups;
        const bool first = (t == 0), last = (t == p.n_tiles - 1);
        const int64_t f0 = (int64_t)t * L;
        const int64_t f1 = (f0 + L < p.n_frames) ? f0 + L : p.n_frames;
        const int len = (int)(f1 - f0);
        const int rowsB = len + Hr;           // y rows: row j <-> frame f0 - Hr + j
        const int rowsA = rowsB + Hf;         // x rows: row r <-> frame f0 - Hr - Hf + r
        const int c = g * kCg + lane;
        const bool cvalid = c < C;

        // ------------------------------------------------ load x tile (+halo) ---
        if (p.vec_ok) {
            // float4: 8 threads cover one 32-channel row, 32 rows per pass
            const int c4 = (tid & 7) * 4, cg4 = g * kCg + c4;
            for (int r = tid >> 3; r < rowsA; r += kTileThreads / 8) {
                const int64_t fr = f0 - Hr - Hf + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cg4 < C) {
                    if (!p.has_fir && first && r < Hr) {
                        v = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.yhist) + (int64_t)r * C + cg4);
                    } else if (fr >= 0) {
                        v = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.in) + fr * C + cg4));
                        const float gsc = (float)p.g_load;
                        v.x *= gsc; v.y *= gsc; v.z *= gsc; v.w *= gsc;
                    } else if (fr >= -Hf) {
                        v = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.xhist) + (Hf + fr) * C + cg4);
                    }
                }
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(bufA) + r * kCg + c4) = v;
            }
        } else {
            for (int r = warp; r < rowsA; r += kNW) {
                const int64_t fr = f0 - Hr - Hf + r;
                T v = T(0);
                if (cvalid) {
                    if (!p.has_fir && first && r < Hr) v = p.yhist[(int64_t)r * C + c];
                    else if (fr >= 0) v = __ldg(p.in + fr * C + c) * p.g_load;
                    else if (fr >= -Hf) v = p.xhist[(Hf + fr) * C + c];
                }
                bufA[r * kCg + lane] = v;
            }
        }
        if (p.has_fir) {
            // zero the rows the chunked FIR may touch past the tile
            for (int i = tid; i < 2 * FB * kCg; i += kTileThreads) bufA[rowsA * kCg + i] = T(0);
            if (first)
                for (int j = warp; j < Hr; j += kNW) bufB[j * kCg + lane] = cvalid ? p.yhist[(int64_t)j * C + c] : T(0);
        }
        __syncthreads();

        // ------------------------------------------- carried FIR input history ---
        if (Hf > 0 && cvalid) {
            const int64_t hs = p.n_frames - Hf;  // first frame kept
            const int64_t a = (hs > f0) ? hs : f0;
            for (int64_t fr = a + warp; fr < f1; fr += kNW)
                p.xhist_next[(fr - hs) * C + c] = bufA[(int)(fr - f0 + Hr + Hf) * kCg + lane];
            if (first && hs < 0)
                for (int64_t j = warp; j < -hs; j += kNW) p.xhist_next[j * C + c] = p.xhist[(p.n_frames + j) * C + c];
        }

        // ---------------------------------------------------------------- FIR ---
        if (p.has_fir) {
            const int fir_r0 = first ? Hr : 0;
            for (int j0 = fir_r0 + warp * FB; j0 < rowsB; j0 += kNW * FB) {
                // Each chunk of FB x-rows is summed in T, then folded into a double accumulator:
                // a 257-term fp32 running sum alone costs ~2e-6 of the channel peak.
                double acc[FB];
#pragma unroll
                for (int i = 0; i < FB; i++) acc[i] = 0.0;
                // y[j0+jj] = sum_k taps[k] * x[row j0+jj+Hf-k];  walk x rows r = j0 + rr
                for (int r0 = 0; r0 < Hf + FB; r0 += FB) {
                    T part[FB];
#pragma unroll
                    for (int i = 0; i < FB; i++) part[i] = T(0);
                    // tap for (rr = r0+i, jj): k = jj + Hf - r0 - i ; padded index k + tp_off
                    const int wbase = Hf - r0 + p.tp_off - (FB - 1);  // index of d = 0, d = jj - i + FB-1
                    T tw[2 * FB - 1];
#pragma unroll
                    for (int d = 0; d < 2 * FB - 1; d++) tw[d] = tp_s[wbase + d];
#pragma unroll
                    for (int i = 0; i < FB; i++) {
                        const T x = bufA[(j0 + r0 + i) * kCg + lane];
#pragma unroll
                        for (int jj = 0; jj < FB; jj++) part[jj] += tw[jj - i + FB - 1] * x;
                    }
#pragma unroll
                    for (int jj = 0; jj < FB; jj++) acc[jj] += (double)part[jj];
                }
#pragma unroll
                for (int jj = 0; jj < FB; jj++)
                    if (j0 + jj < rowsB) bufB[(j0 + jj) * kCg + lane] = (T)(acc[jj] * (double)p.g_fir);
            }
            __syncthreads();
        }
        T *y = bufB;

        // ------------------------------------------------------------- biquad ---
        if (p.has_bq) {
            const int ra = first ? Hr : 0;
            const int rb = last ? rowsB : len;   // chain region [ra, rb); rows [rb, rowsB) are continued but not chained
            const int nR = rb - ra;
            const int cbq = ra + (int)(((int64_t)warp * nR) / kNW);
            const int cbq1 = ra + (int)(((int64_t)(warp + 1) * nR) / kNW);
            // pass 1: zero-state end state of this warp's sub-chunk, z = sum_n A^(end-1-n) B y[n]
            {
                double z0 = 0.0, z1 = 0.0;
                for (int n = cbq; n < cbq1; n++) {
                    const double x = (double)y[n * kCg + lane];
                    const int k = cbq1 - 1 - n;
                    z0 += wt_s[2 * k] * x;
                    z1 += wt_s[2 * k + 1] * x;
                }
                zq_s[(warp * kCg + lane) * 2] = z0;
                zq_s[(warp * kCg + lane) * 2 + 1] = z1;
            }
            __syncthreads();
            if (warp == 0) {
                // tile aggregate (zero-state end state over the whole chain region)
                double Z0 = 0.0, Z1 = 0.0;
                for (int q = 0; q < kNW; q++) {
                    const int lq = (int)(((int64_t)(q + 1) * nR) / kNW) - (int)(((int64_t)q * nR) / kNW);
                    mat2_apply(p.bq_apow + 4 * lq, Z0, Z1);
                    Z0 += zq_s[(q * kCg + lane) * 2];
                    Z1 += zq_s[(q * kCg + lane) * 2 + 1];
                }
                const size_t slot = (size_t)g * p.n_tiles + t;
                if (!last && !first) {
                    p.lb_agg[slot * 64 + lane * 2] = Z0;
                    p.lb_agg[slot * 64 + lane * 2 + 1] = Z1;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbAgg);
                }
                // incoming state
                double S0, S1;
                if (first) {
                    S0 = cvalid ? p.bq_state[2 * c] : 0.0;
                    S1 = cvalid ? p.bq_state[2 * c + 1] : 0.0;
                } else {
                    // decoupled look-back: lane i watches tile t-1-i
                    const int base = t - 1;
                    const int j = base - lane;
                    int first_inc = 0;
                    for (unsigned spins = 0;; spins++) {
                        unsigned st = kLbInc;  // lanes before the stream start count as resolved
                        if (j >= 0) {
                            st = ld_acquire_u32(p.lb_status + (size_t)g * p.n_tiles + j);
                            st = ((st >> 2) == p.epoch) ? (st & 3u) : kLbNone;
                        }
                        const unsigned ready = __ballot_sync(0xffffffffu, st != kLbNone);
                        const unsigned inc = __ballot_sync(0xffffffffu, st == kLbInc);
                        if (inc) {
                            first_inc = __ffs(inc) - 1;
                            const unsigned need = (first_inc == 0) ? 0u : (0xffffffffu >> (32 - first_inc));
                            if ((ready & need) == need) break;
                        }
                        if (spins > (1u << 24)) {  // ~1 s: a predecessor never published
                            if (lane == 0) atomicExch(p.err_flag, 1);
                            first_inc = -1;
                            break;
                        }
                        __nanosleep(40);
                    }
                    __threadfence();
                    __syncwarp();
                    const size_t s_inc = (size_t)g * p.n_tiles + (first_inc < 0 ? 0 : base - first_inc);
                    S0 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2);
                    S1 = ld_cg(p.lb_inc + s_inc * 64 + lane * 2 + 1);
                    const double *ML = p.bq_apow + 4 * L;  // every chained predecessor covers exactly L frames
                    for (int i = first_inc - 1; i >= 0; i--) {
                        const size_t sa = (size_t)g * p.n_tiles + (base - i);
                        const double a0 = ld_cg(p.lb_agg + sa * 64 + lane * 2);
                        const double a1 = ld_cg(p.lb_agg + sa * 64 + lane * 2 + 1);
                        mat2_apply(ML, S0, S1);
                        S0 += a0;
                        S1 += a1;
                    }
                }
                // inclusive state after the chain region
                double I0 = S0, I1 = S1;
                mat2_apply(p.bq_apow + 4 * nR, I0, I1);
                I0 += Z0;
                I1 += Z1;
                if (!last) {
                    p.lb_inc[slot * 64 + lane * 2] = I0;
                    p.lb_inc[slot * 64 + lane * 2 + 1] = I1;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_u32(p.lb_status + slot, (p.epoch << 2) | kLbInc);
                } else if (cvalid) {
                    p.bq_state_next[2 * c] = I0;
                    p.bq_state_next[2 * c + 1] = I1;
                }
                // incoming state of every sub-chunk
                for (int q = 0; q < kNW; q++) {
                    sin_s[(q * kCg + lane) * 2] = S0;
                    sin_s[(q * kCg + lane) * 2 + 1] = S1;
                    const int lq = (int)(((int64_t)(q + 1) * nR) / kNW) - (int)(((int64_t)q * nR) / kNW);
                    mat2_apply(p.bq_apow + 4 * lq, S0, S1);
                    S0 += zq_s[(q * kCg + lane) * 2];
                    S1 += zq_s[(q * kCg + lane) * 2 + 1];
                }
            }
            __syncthreads();
            // pass 2: the actual recurrence from the true incoming state, in place
            {
                double s1 = sin_s[(warp * kCg + lane) * 2], s2 = sin_s[(warp * kCg + lane) * 2 + 1];
                const int end = (warp == kNW - 1) ? rowsB : cbq1;
                const double gbq = (double)p.g_bq;
                for (int n = cbq; n < end; n++) {
                    const double x = (double)y[n * kCg + lane];
                    const double v = p.b0 * x + s1;
                    s1 = p.b1 * x - p.a1 * v + s2;
                    s2 = p.b2 * x - p.a2 * v;
                    y[n * kCg + lane] = (T)(v * gbq);
                }
            }
            __syncthreads();
        }

        // --------------------------------------- carried resampler input history ---
        if (Hr > 0 && last && cvalid)
            for (int j = warp; j < Hr; j += kNW) p.yhist_next[(int64_t)j * C + c] = y[(len + j) * kCg + lane];

        // ------------------------------------------------- resample / store out ---
        double m_peak = 0.0, m_sumsq = 0.0;
        const bool meter = p.meter_peak != nullptr;
        if (p.has_rs) {
            const int64_t up = p.rs_up, down = p.rs_down;
            const int64_t m_lo = (p.rs_acc0 + f0 * up) / down, m_hi = (p.rs_acc0 + f1 * up) / down;
            const int P = p.rs_P;
            for (int64_t m = m_lo + warp; m < m_hi; m += kNW) {
                const int64_t num = (m + 1) * down - p.rs_acc0;
                const int64_t im = (num + up - 1) / up - 1;                  // frame that triggers output m
                const int br = (int)(up - 1 - (p.rs_acc0 + (im + 1) * up - (m + 1) * down));
                const int row = (int)(im - f0) + Hr;
                const T *__restrict__ cf = p.rs_coef + (size_t)br * P;
                double dacc = 0.0;
                for (int k0 = 0; k0 < P; k0 += 8) {
                    T part = T(0);
                    const int k1 = (k0 + 8 < P) ? k0 + 8 : P;
                    for (int k = k0; k < k1; k++) part += __ldg(cf + k) * y[(row - k) * kCg + lane];
                    dacc += (double)part;
                }
                const T acc = (T)(dacc * (double)p.g_out);
                if (cvalid) {
                    p.out[m * C + c] = acc;
                    if (meter) {
                        const double a = fabs((double)acc);
                        m_peak = a > m_peak ? a : m_peak;
                        m_sumsq += (double)acc * (double)acc;
                    }
                }
            }
        } else if (p.vec_ok && !meter) {
            const int c4 = (tid & 7) * 4, cg4 = g * kCg + c4;
            if (cg4 < C)
                for (int r = tid >> 3; r < len; r += kTileThreads / 8)
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + (f0 + r) * C + cg4) =
                        *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(y) + r * kCg + c4);
        } else if (cvalid) {
            for (int r = warp; r < len; r += kNW) {
                const T v = y[r * kCg + lane];
                p.out[(f0 + r) * C + c] = v;
                if (meter) {
                    const double a = fabs((double)v);
                    m_peak = a > m_peak ? a : m_peak;
                    m_sumsq += (double)v * (double)v;
                }
            }
        }
        if (meter) {
            red_s[(warp * kCg + lane) * 2] = m_peak;
            red_s[(warp * kCg + lane) * 2 + 1] = m_sumsq;
            __syncthreads();
            if (warp == 0 && cvalid) {
                double pk = 0.0, sq = 0.0;
                for (int q = 0; q < kNW; q++) {
                    const double a = red_s[(q * kCg + lane) * 2];
                    pk = a > pk ? a : pk;
                    sq += red_s[(q * kCg + lane) * 2 + 1];
                }
                atomic_max_nonneg(p.meter_peak + c, pk);
                atomicAdd(p.meter_sumsq + c, sq);
            }
        }
    }
}

}  // namespace pb
