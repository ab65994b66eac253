// mma_probe.cu -- development probe: tcgen05.mma (M=128, kind::f16, SS) cost versus N, accumulator column offset and
// operand layout, and whether a running UMMA stream slows DFMA on the same SM.
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
DEVI uint32_t make_idesc(int n) { return (1u << 4) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
DEVI void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
struct Cfg { int n, col, a_kb, a_mb, b_lbo, b_sbo, alt, dp, batch4; };

__global__ void __launch_bounds__(160) k(Cfg c, int reps, long long *cyc, double *sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 65536 + 32);
    volatile int *stop = reinterpret_cast<volatile int *>(smem + 65536 + 48);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // halves 1.0
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
    if (tid == 0) {
        const uint64_t a = make_desc(smem_u32(smem), c.a_kb, c.a_mb);
        const uint64_t b = make_desc(smem_u32(smem) + 16384, c.b_lbo, c.b_sbo);
        const uint32_t id = make_idesc(c.n);
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
            if (c.batch4) {
                // the kernel's pattern: per chunk 4 MMAs (E, X, X, X) and a commit nobody waits for; 4 chunks per rep
#pragma unroll 1
                for (int u = 0; u < 4; u++) {
                    const uint64_t as = a + (uint64_t)((u % 3) * (8192 >> 4));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    umma(tb + c.col, as, b, id, 1);
                    umma(tb + c.col + 176, as, b + 600, id, 1);
                    umma(tb + c.col + 176, as, b + 1200, id, 1);
                    umma(tb + c.col + 176, as + 256, b + 1800, id, 1);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
                }
                if ((r & 7) == 7) {
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
                    mbar_wait(bar, (r >> 3) & 1);
                }
            } else if (c.n > 0) {
#pragma unroll
                for (int u = 0; u < 16; u++) umma(tb + c.col + ((c.alt && (u & 1)) ? 192 : 0), a, b, id, 1);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
                mbar_wait(bar, r & 1);
            } else {
                __nanosleep(1000);
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) cyc[0] = t1 - t0;
        *stop = 1;
    } else if (warp >= 1 && c.dp) {
        double acc[8];
        for (int i = 0; i < 8; i++) acc[i] = tid * 1e-3 + i;
        long long t0 = clock64();
        long long n = 0;
        while (!*stop) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], 0.999, 0.5);
            n += 64;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && tid == 32) { cyc[1] = t1 - t0; cyc[2] = n; }
        double s = 0;
        for (int i = 0; i < 8; i++) s += acc[i];
        sink[blockIdx.x * 160 + tid] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
int main()
{
    long long *c, h[3];
    double *sink;
    cudaMalloc(&c, 24); cudaMalloc(&sink, 8 * 148 * 160);
    const int smem = 65536 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 200;
    Cfg cfgs[] = {
        {176, 0, 2048, 128, 128, 128, 0, 0}, {176, 0, 2048, 128, 128, 128, 1, 0}, {128, 0, 2048, 128, 128, 128, 0, 0}, {96, 0, 2048, 128, 128, 128, 0, 0},
        {64, 0, 2048, 128, 128, 128, 0, 0},  {32, 0, 2048, 128, 128, 128, 0, 0},  {16, 0, 2048, 128, 128, 128, 0, 0},
        {64, 8, 2048, 128, 128, 128, 0, 0},  {64, 16, 2048, 128, 128, 128, 0, 0}, {64, 24, 2048, 128, 128, 128, 0, 0}, {64, 32, 2048, 128, 128, 128, 0, 0},
        {160, 8, 2048, 128, 128, 128, 1, 0}, {160, 16, 2048, 128, 128, 128, 1, 0},
        {32, 368, 2304, 144, 128, 256, 0, 0}, {64, 368, 2304, 144, 128, 256, 0, 0},
        {176, 0, 2048, 128, 128, 352, 0, 0},  // B with non-overlapping core matrices (dense K-major tile)
        {256, 0, 2048, 128, 128, 256, 0, 0},
        {0, 0, 2048, 128, 128, 128, 0, 1},    // DFMA warps alone
        {176, 0, 2048, 128, 128, 128, 1, 1},  // DFMA warps next to a saturated UMMA stream
        {32, 368, 2304, 144, 128, 256, 0, 1},
        {176, 0, 2048, 128, 128, 128, 0, 0, 1}, {96, 8, 2048, 128, 128, 128, 0, 0, 1}, {32, 16, 2048, 128, 128, 128, 0, 0, 1}, {16, 0, 2048, 128, 128, 128, 0, 0, 1},
    };
    for (auto &cf : cfgs) {
        h[0] = h[1] = h[2] = 0;
        cudaMemset(c, 0, 24);
        k<<<148, 160, smem>>>(cf, reps, c, sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
        printf("N %3d col %3d A(kb %4d mb %3d) B(lbo %3d sbo %3d) alt %d : %.1f cycles/MMA (floor %.0f)", cf.n, cf.col, cf.a_kb, cf.a_mb, cf.b_lbo,
               cf.b_sbo, cf.alt, cf.n ? (double)h[0] / reps / 16 : 0.0, cf.n / 2.0);
        if (cf.batch4) printf(" [batches of 4 + commit]");
        if (cf.dp) printf("  | DFMA: %.2f cycles/op/warp (4 warps)", h[2] ? (double)h[1] / h[2] : 0.0);
        printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
