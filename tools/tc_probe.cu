// tc_probe.cu -- development probe for the tcgen05 FIR formulation (not part of the product library).
// One CTA computes one tile D[128 ch x 192 cols] = X^T[128 x 432] * B[432 x 192] on tcgen05 with
//   * A = X^T, MN-major fp16 hi/lo pieces written by the threads,
//   * B = Toeplitz matrix addressed through overlapping core matrices (LBO = SBO = 128 B),
//   * 16 extra dense columns (the "V" rows),
// and dumps the accumulator so numpy can check descriptors, layout and split-precision numerics.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#define DEVI __device__ __forceinline__

DEVI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVI void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DEVI void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

DEVI uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

DEVI uint32_t make_idesc(int n)
{
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 0u << 7;                    // a_format = F16
    d |= 0u << 10;                   // b_format = F16
    d |= 1u << 15;                   // a_major = MN
    d |= 0u << 16;                   // b_major = K
    d |= (uint32_t)(n >> 3) << 17;   // n_dim
    d |= (uint32_t)(128 >> 4) << 24; // m_dim
    return d;
}

DEVI void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
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

constexpr int kRows = 432, kCh = 128, kCols = 192, kChunks = 27;

extern "C" __global__ void __launch_bounds__(128) tc_probe_kernel(const float *__restrict__ x,  // [432][128]
                                                                  const __half *__restrict__ tbl,  // hi[75*64] lo[75*64] vhi[27*256] vlo[27*256]
                                                                  float *__restrict__ out,         // [128][192]
                                                                  float xscale, int split_acc, int fixed_hi)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __half *A_hi = reinterpret_cast<__half *>(smem);             // 4096 B
    __half *A_lo = reinterpret_cast<__half *>(smem + 4096);      // 4096 B
    __half *T_hi = reinterpret_cast<__half *>(smem + 8192);      // 9600 B
    __half *T_lo = reinterpret_cast<__half *>(smem + 8192 + 9600);
    __half *V_hi = reinterpret_cast<__half *>(smem + 8192 + 19200);           // 27*512
    __half *V_lo = reinterpret_cast<__half *>(smem + 8192 + 19200 + 13824);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192 + 19200 + 27648);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 8192 + 19200 + 27648 + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 75 * 64; i += 128) {
        T_hi[i] = tbl[i];
        T_lo[i] = tbl[75 * 64 + i];
    }
    for (int i = tid; i < 27 * 256; i += 128) {
        V_hi[i] = tbl[2 * 75 * 64 + i];
        V_lo[i] = tbl[2 * 75 * 64 + 27 * 256 + i];
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t idesc_main = make_idesc(176), idesc_v = make_idesc(16);
    // converter mapping: lane -> (fr_i = lane % 8, mbq = lane / 8); warp -> channel group of 32
    const int fr_i = lane & 7, mbq = lane >> 3, mb = warp * 4 + mbq;

    for (int q = 0; q < kChunks; q++) {
        for (int kb = 0; kb < 2; kb++) {
            const int row = 16 * q + 8 * kb + fr_i;
            const float *src = x + (size_t)row * kCh + mb * 8;
            __half2 hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float a = src[2 * i] * xscale, b = src[2 * i + 1] * xscale;
                if (fixed_hi == 1) {  // hi piece on a fixed integer grid: products and partial sums stay exact in fp32
                    const float ra = rintf(a), rb = rintf(b);
                    hi[i] = __floats2half2_rn(ra, rb);
                    lo[i] = __floats2half2_rn(a - ra, b - rb);
                } else {
                    const __half2 h = __floats2half2_rn(a, b);
                    const float2 hf = __half22float2(h);
                    hi[i] = h;
                    lo[i] = __floats2half2_rn(a - hf.x, b - hf.y);
                }
            }
            const int kbpos = 1 - kb;  // the two K-blocks are stored swapped (Toeplitz overlap trick)
            const int off = kbpos * 2048 + mb * 128 + fr_i * 16;
            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(A_hi) + off) = *reinterpret_cast<uint4 *>(hi);
            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(A_lo) + off) = *reinterpret_cast<uint4 *>(lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint64_t a_hi = make_desc(smem_u32(A_hi), 2048, 128), a_lo = make_desc(smem_u32(A_lo), 2048, 128);
            const uint64_t b_hi = make_desc(smem_u32(T_hi) + (52 - 2 * q) * 128, 128, 128);
            const uint64_t b_lo = make_desc(smem_u32(T_lo) + (52 - 2 * q) * 128, 128, 128);
            const uint64_t v_hi = make_desc(smem_u32(V_hi) + q * 512, 256, 128);
            const uint64_t v_lo = make_desc(smem_u32(V_lo) + q * 512, 256, 128);
            if (split_acc) {
                umma(tmem_base, a_hi, b_hi, idesc_main, q > 0);
                umma(tmem_base + 192, a_lo, b_hi, idesc_main, q > 0);
                umma(tmem_base + 192, a_hi, b_lo, idesc_main, 1);
                umma(tmem_base + 176, a_hi, v_hi, idesc_v, q > 0);
                umma(tmem_base + 368, a_lo, v_hi, idesc_v, q > 0);
                umma(tmem_base + 368, a_hi, v_lo, idesc_v, 1);
            } else {
                umma(tmem_base, a_hi, b_hi, idesc_main, q > 0);
                umma(tmem_base, a_lo, b_hi, idesc_main, 1);
                umma(tmem_base, a_hi, b_lo, idesc_main, 1);
                umma(tmem_base + 176, a_hi, v_hi, idesc_v, q > 0);
                umma(tmem_base + 176, a_lo, v_hi, idesc_v, 1);
                umma(tmem_base + 176, a_hi, v_lo, idesc_v, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(bar, q & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < 2 * kCols; c0 += 16) {
        if (!split_acc && c0 >= kCols) break;
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) out[(size_t)(warp * 32 + lane) * (2 * kCols) + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

extern "C" int tc_probe_run(const float *x_host, const void *tbl_host, float *out_host, float xscale, int split_acc, int fixed_hi)
{
    float *dx, *dout;
    void *dtbl;
    const size_t tbl_bytes = (2 * 75 * 64 + 2 * 27 * 256) * 2;
    if (cudaMalloc(&dx, kRows * kCh * 4) || cudaMalloc(&dout, kCh * kCols * 8) || cudaMalloc(&dtbl, tbl_bytes)) return -1;
    cudaMemcpy(dx, x_host, kRows * kCh * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dtbl, tbl_host, tbl_bytes, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, kCh * kCols * 8);
    const int smem = 8192 + 19200 + 27648 + 64;
    cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_probe_kernel<<<1, 128, smem>>>(dx, (const __half *)dtbl, dout, xscale, split_acc, fixed_hi);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("tc_probe: %s\n", cudaGetErrorString(e));
        return -2;
    }
    cudaMemcpy(out_host, dout, kCh * kCols * 8, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dout);
    cudaFree(dtbl);
    return 0;
}
