// ts_probe.cu -- development probe (not part of the product library): tcgen05.mma with the A operand in TENSOR MEMORY.
//   1. layout check: A[m][k] written with tcgen05.st (lane = row m, 8 columns = 16 fp16 K-values), B = identity:
//      D[m][n] must come back as A[m][n] -> tells which half of which column is K index k;
//   2. numerics check with a dense integer B;
//   3. cycles per MMA for SS (A in shared memory) and TS (A in tensor memory), alone and next to warps that
//      saturate the shared-memory crossbar (the converter / drain traffic of the chain kernel);
//   4. tcgen05.ld / tcgen05.st throughput with 4 and 8 warps.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#define DEVI __device__ __forceinline__
DEVI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
DEVI void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DEVI uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// a_mn: A is MN-major (SS mode of the chain kernel); TS mode needs K-major A
DEVI uint32_t make_idesc(int n, int a_mn) { return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
DEVI void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
DEVI void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
DEVI void commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
DEVI void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
DEVI void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

constexpr int kBBytes = 32768;   // B operand region
constexpr int kABytes = 16384;   // A operand region (SS mode), 3 stages of 4 KB + slack
constexpr int kHamBytes = 65536; // region the hammer warps stream through
constexpr int kOffA = kBBytes, kOffHam = kBBytes + kABytes, kOffBar = kOffHam + kHamBytes;
constexpr int kSmem = kOffBar + 128;

// B element (n, k) of a K-major no-swizzle operand with LBO = 128 (K-adjacent core matrices), SBO = 256
DEVI int b_off(int n, int k) { return (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }

// ---------------------------------------------------------------- 1 + 2: layout and numerics -----------------
__global__ void __launch_bounds__(128) check_kernel(int mode, float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + kOffBar + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kBBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    __syncthreads();
    // B[k][n]: mode 0 identity (n < 16), mode 1 dense integers ((3 n + 5 k) % 7) - 3
    for (int i = tid; i < 64 * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        const float v = mode == 0 ? (n == k ? 1.f : 0.f) : (float)(((3 * n + 5 * k) % 7) - 3);
        *reinterpret_cast<__half *>(smem + b_off(n, k)) = __float2half(v);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = *slot;
    // A[m][k] = m + 100 k   (exact in fp16 for m < 128, k < 16: max 1627)... keep below 2048
    {
        uint32_t r[8];
        for (int c = 0; c < 8; c++) {
            const __half lo = __float2half((float)(tid + 100 * (2 * c))), hi = __float2half((float)(tid + 100 * (2 * c + 1)));
            r[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
        }
        tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 256, r);   // A at columns [256, 264)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        umma_ts(tb, tb + 256, make_desc(smem_u32(smem), 128, 256), make_idesc(64, 0), 0);
        commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) out[tid * 64 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

// ---------------------------------------------------------------- 3: cycles per MMA -----------------------------
struct Cfg { int n, ts, hammer, per_commit; };

DEVI bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

template <bool TS>
__global__ void __launch_bounds__(32 * 10) rate_kernel(Cfg c, int reps, long long *cyc, float *sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + kOffBar + 64);
    volatile int *stop = reinterpret_cast<volatile int *>(smem + kOffBar + 96);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (kSmem - 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // halves 1.0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = *slot;
    if (warp == 0) {
        // the whole warp runs the loop converged and elects one lane per issue (operands stay in uniform registers)
        const uint64_t a = make_desc(smem_u32(smem) + kOffA, 2048, 128);   // MN-major A as in the chain kernel
        const uint64_t b = make_desc(smem_u32(smem), 128, 128);            // Toeplitz-style overlapping core matrices
        const uint32_t id = make_idesc(c.n, TS ? 0 : 1);
        const long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
            if (c.per_commit == 4) {
                // the chain kernel's MMA1 loop: 4 chunks of (E, X, X, X) + a commit nobody waits for, then one waited commit
#pragma unroll 1
                for (int q = 0; q < 4; q++) {
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
                        const uint32_t a0 = tb + 400 + 16 * (q & 3);
                        umma_ts(tb, a0, b, id, 1);
                        umma_ts(tb + 176, a0, b + 600, id, 1);
                        umma_ts(tb + 176, a0, b + 1200, id, 1);
                        umma_ts(tb + 176, a0 + 8, b + 1800, id, 1);
                        commit(bar + 1);
                    }
                    __syncwarp();
                }
                if (elect_one()) commit(bar);
            } else if (c.per_commit == 1) {
                // 16 MMAs accumulating into the SAME columns (dependent chain)
                if (elect_one()) {
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        if (TS) umma_ts(tb, tb + 400 + 16 * (u % 3), b + (uint64_t)(u * 8), id, 1);
                        else umma_ss(tb, a + (uint64_t)((u % 3) * (4096 >> 4)), b + (uint64_t)(u * 8), id, 1);
                    }
                    commit(bar);
                }
            } else if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    if (TS) umma_ts(tb + (u & 1) * 176, tb + 400 + 16 * (u % 3), b + (uint64_t)(u * 8), id, 1);
                    else umma_ss(tb + (u & 1) * 176, a + (uint64_t)((u % 3) * (4096 >> 4)), b + (uint64_t)(u * 8), id, 1);
                }
                commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, r & 1);
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) cyc[0] = t1 - t0;
        if (lane == 0) *stop = 1;
    } else if (warp <= c.hammer) {
        // shared-memory traffic like the converter's: LDS.128 + STS.128 streams over a 64 KB region, conflict-free
        float4 acc = make_float4(0, 0, 0, 0);
        long long n = 0;
        const long long t0 = clock64();
        unsigned char *base = smem + kOffHam;
        while (!*stop) {
#pragma unroll 8
            for (int u = 0; u < 32; u++) {
                const int off = ((warp * 4096 + u * 512 + lane * 16) & (kHamBytes - 1));
                const float4 v = *reinterpret_cast<const float4 *>(base + off);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                *reinterpret_cast<float4 *>(base + ((off + 32768) & (kHamBytes - 1))) = acc;
            }
            n += 64;  // 64 warp-wide 512 B accesses
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) { cyc[2 * warp] = t1 - t0; cyc[2 * warp + 1] = n; }
        sink[blockIdx.x * blockDim.x + tid] = acc.x + acc.y + acc.z + acc.w;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

// ---------------------------------------------------------------- 4: tcgen05.ld / st throughput -----------------
__global__ void __launch_bounds__(256) tmem_rate_kernel(int nwarps, int store, int reps, long long *cyc, float *sink)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    float acc = 0.f;
    uint32_t r[16];
    for (int i = 0; i < 16; i++) r[i] = tid + i;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < reps; it++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (store) {
                    uint32_t q[8] = {r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]};
                    tmem_st8(tb + 8 * u, q);
                    tmem_st8(tb + 128 + 8 * u, q);
                } else {
                    tmem_ld16(tb + 16 * u, r);
                }
            }
            if (store) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[it & 15]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) cyc[0] = t1 - t0;
    sink[blockIdx.x * 256 + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

int main()
{
    long long *c, h[32];
    float *sink, *out;
    cudaMalloc(&c, sizeof h);
    cudaMalloc(&sink, 4 * 148 * 320);
    cudaMalloc(&out, 4 * 128 * 64);
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    static float ho[128 * 64];
    for (int mode = 0; mode < 2; mode++) {
        cudaMemset(out, 0, sizeof ho);
        check_kernel<<<1, 128, kSmem>>>(mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
        if (mode == 0) {
            printf("TS layout check (%s): D[m][n] for m = 5, n = 0..15 (A[m][k] = m + 100 k, B = identity):\n ", cudaGetErrorString(e));
            for (int n = 0; n < 16; n++) printf(" %g", ho[5 * 64 + n]);
            int bad = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 16; n++) bad += ho[m * 64 + n] != (float)(m + 100 * n);
            printf("\n  mismatches vs 'column c holds K = 2c (low half), 2c+1 (high half)': %d of 2048\n", bad);
        } else {
            int bad = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 64; n++) {
                    float ref = 0;
                    for (int k = 0; k < 16; k++) ref += (float)(m + 100 * k) * (float)(((3 * n + 5 * k) % 7) - 3);
                    bad += ho[m * 64 + n] != ref;
                }
            printf("TS numerics check (%s): mismatches %d of 8192\n", cudaGetErrorString(e), bad);
        }
    }
    const int reps = 200;
    Cfg cfgs[] = {
        {176, 1, 0, 1}, {176, 0, 0, 1}, {96, 1, 0, 1}, {32, 1, 0, 1}, {176, 1, 0, 4}, {96, 1, 0, 4}, {32, 1, 0, 4}, {176, 1, 8, 4}, {32, 1, 8, 4},
        {176, 0, 0, 16}, {176, 1, 0, 16}, {128, 0, 0, 16}, {128, 1, 0, 16}, {64, 0, 0, 16}, {64, 1, 0, 16}, {32, 1, 0, 16}, {16, 1, 0, 16},
        {176, 0, 4, 16}, {176, 1, 4, 16}, {176, 0, 8, 16}, {176, 1, 8, 16}, {128, 0, 8, 16}, {128, 1, 8, 16}, {64, 0, 8, 16}, {64, 1, 8, 16},
    };
    for (auto &cf : cfgs) {
        cudaMemset(c, 0, sizeof h);
        if (cf.ts) rate_kernel<true><<<148, 320, kSmem>>>(cf, reps, c, sink);
        else rate_kernel<false><<<148, 320, kSmem>>>(cf, reps, c, sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
        printf("N %3d %s hammer warps %d %s: %.1f cycles/MMA (floor %.0f)", cf.n, cf.ts ? "TS" : "SS", cf.hammer,
               cf.per_commit == 1 ? "same D (dependent chain) " : cf.per_commit == 4 ? "chain-kernel pattern E,X,X,X per chunk " : "", (double)h[0] / reps / 16,
               cf.n / 2.0);
        if (cf.hammer) {
            double bpc = 0;
            for (int w = 1; w <= cf.hammer; w++) bpc += h[2 * w] ? 512.0 * (double)h[2 * w + 1] / (double)h[2 * w] : 0.0;
            printf("  | hammer smem traffic %.1f B/cycle", bpc);
        }
        printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    for (int store = 0; store < 2; store++)
        for (int nw : {4, 8}) {
            cudaMemset(c, 0, sizeof h);
            tmem_rate_kernel<<<148, 256>>>(nw, store, 100, c, sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * 100 * 16 * 16 * 128;  // per warp and iteration: 16 x (16 columns x 32 lanes x 4 B)
            printf("tcgen05.%s %d warps: %.1f B/cycle/SM  %s\n", store ? "st" : "ld", nw, bytes / (double)h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
