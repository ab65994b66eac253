// aux_kernels.cu -- Source- and Sink-side kernels around the Processor run:
// synthetic Source fill, meter Sink (peak / sum of squares, warp-shuffle
// reductions) and the fan-in mixer Sink (sum over Lines, peer pointers allowed).
#include "common.cuh"

namespace pb {

// ---- synthetic Source --------------------------------------------------------
// mock.Source analogue (reference mock/mock.go:86-105) with the BASELINE.md
// formula instead of a constant so that parity sees real signal content.
template <typename T>
__global__ void source_fill_kernel(T *__restrict__ out, int64_t first_index, int64_t n, uint64_t seed, uint64_t line)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t z = splitmix64(seed ^ (line << 48) ^ (uint64_t)(first_index + i));
        out[i] = (T)((double)(z >> 40) * (1.0 / 8388608.0) - 1.0);
    }
}

// ---- meter Sink --------------------------------------------------------------
// One warp owns 32 consecutive channels (coalesced 128 B rows); blocks stride
// over frames; per-channel partials are combined across the block's warps in
// shared memory, then one atomic per channel per block.
template <typename T>
__global__ void __launch_bounds__(256) meter_kernel(const T *__restrict__ in, int64_t frames, int C,
                                                    double *__restrict__ peak, double *__restrict__ sumsq,
                                                    int frame_blocks)
{
    __shared__ double s_pk[8][32], s_sq[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x / frame_blocks, fb = blockIdx.x % frame_blocks;
    const int c = g * 32 + lane;
    double pk = 0.0, sq = 0.0;
    if (c < C) {
        for (int64_t f = (int64_t)fb * 8 + warp; f < frames; f += (int64_t)frame_blocks * 8) {
            const double v = (double)__ldg(in + f * C + c);
            const double a = fabs(v);
            pk = a > pk ? a : pk;
            sq += v * v;
        }
    }
    s_pk[warp][lane] = pk;
    s_sq[warp][lane] = sq;
    __syncthreads();
    if (warp == 0 && c < C) {
        for (int w = 1; w < 8; w++) {
            pk = s_pk[w][lane] > pk ? s_pk[w][lane] : pk;
            sq += s_sq[w][lane];
        }
        atomic_max_nonneg(peak + c, pk);
        atomicAdd(sumsq + c, sq);
    }
}

// ---- fan-in mixer Sink ---------------------------------------------------------
// out[i] = sum_l in[l][i].  Inputs may live on peer GPUs (pointers opened with
// pb_ipc_open): the loads then cross NVLink inside this kernel, so the transfer
// and the sum are one pass with no staging copy.  Summation order is l = 0..n-1,
// the order the oracle uses.
constexpr int kMaxMixInputs = 16;
struct MixPtrs {
    const void *p[kMaxMixInputs];
};

template <typename T, typename V, int VEC>
__global__ void __launch_bounds__(256) mix_sum_kernel(const __grid_constant__ MixPtrs ins, int n_inputs, int64_t n_vec, V *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        V acc = reinterpret_cast<const V *>(ins.p[0])[i];
        T *a = reinterpret_cast<T *>(&acc);
        for (int l = 1; l < n_inputs; l++) {
            const V v = reinterpret_cast<const V *>(ins.p[l])[i];
            const T *b = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int k = 0; k < VEC; k++) a[k] += b[k];
        }
        out[i] = acc;
    }
}

static int num_sms(int device)
{
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n;
}

}  // namespace pb

using namespace pb;

extern "C" int32_t pb_source_fill_device(int32_t device, int32_t dtype, void *out_dev, int64_t first_index,
                                         int64_t n_values, uint64_t seed, uint64_t line, void *stream)
{
    if (!out_dev || n_values < 0 || (dtype != PB_F32 && dtype != PB_F64)) return fail(PB_ERR_INVALID, "pb_source_fill_device: bad argument");
    if (n_values == 0) return PB_OK;
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    const int blocks = (int)std::min<int64_t>(ceil_div64(n_values, 256), (int64_t)num_sms(device) * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == PB_F32)
        source_fill_kernel<float><<<blocks, 256, 0, s>>>((float *)out_dev, first_index, n_values, seed, line);
    else
        source_fill_kernel<double><<<blocks, 256, 0, s>>>((double *)out_dev, first_index, n_values, seed, line);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int32_t pb_meter_device(int32_t device, int32_t dtype, const void *in_dev, int64_t frames, int32_t channels,
                                   double *peak_dev, double *sumsq_dev, void *stream)
{
    if (!in_dev || !peak_dev || !sumsq_dev || frames < 0 || channels < 1 || (dtype != PB_F32 && dtype != PB_F64))
        return fail(PB_ERR_INVALID, "pb_meter_device: bad argument");
    if (frames == 0) return PB_OK;
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    const int groups = (channels + 31) / 32;
    int fblocks = (int)std::min<int64_t>(ceil_div64(frames, 64), std::max(1, num_sms(device) * 8 / groups));
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == PB_F32)
        meter_kernel<float><<<groups * fblocks, 256, 0, s>>>((const float *)in_dev, frames, channels, peak_dev, sumsq_dev, fblocks);
    else
        meter_kernel<double><<<groups * fblocks, 256, 0, s>>>((const double *)in_dev, frames, channels, peak_dev, sumsq_dev, fblocks);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int32_t pb_mix_sum_device(int32_t device, int32_t dtype, const void *const *inputs_dev, int32_t n_inputs,
                                     int64_t n_values, void *out_dev, void *stream)
{
    if (!inputs_dev || !out_dev || n_inputs < 1 || n_inputs > kMaxMixInputs || n_values < 0 ||
        (dtype != PB_F32 && dtype != PB_F64))
        return fail(PB_ERR_INVALID, "pb_mix_sum_device: bad argument (1..%d inputs)", kMaxMixInputs);
    if (n_values == 0) return PB_OK;
    DeviceGuard dg(device);
    PB_CUDA(dg.err);
    MixPtrs ptrs{};
    bool aligned = ((uintptr_t)out_dev % 16) == 0;
    for (int i = 0; i < n_inputs; i++) {
        if (!inputs_dev[i]) return fail(PB_ERR_INVALID, "pb_mix_sum_device: input %d is NULL", i);
        ptrs.p[i] = inputs_dev[i];
        aligned = aligned && ((uintptr_t)inputs_dev[i] % 16) == 0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int sms = num_sms(device);
    if (dtype == PB_F32) {
        if (aligned && n_values % 4 == 0) {
            const int64_t nv = n_values / 4;
            const int blocks = (int)std::min<int64_t>(ceil_div64(nv, 256), (int64_t)sms * 8);
            mix_sum_kernel<float, float4, 4><<<blocks, 256, 0, s>>>(ptrs, n_inputs, nv, (float4 *)out_dev);
        } else {
            const int blocks = (int)std::min<int64_t>(ceil_div64(n_values, 256), (int64_t)sms * 8);
            mix_sum_kernel<float, float, 1><<<blocks, 256, 0, s>>>(ptrs, n_inputs, n_values, (float *)out_dev);
        }
    } else {
        if (aligned && n_values % 2 == 0) {
            const int64_t nv = n_values / 2;
            const int blocks = (int)std::min<int64_t>(ceil_div64(nv, 256), (int64_t)sms * 8);
            mix_sum_kernel<double, double2, 2><<<blocks, 256, 0, s>>>(ptrs, n_inputs, nv, (double2 *)out_dev);
        } else {
            const int blocks = (int)std::min<int64_t>(ceil_div64(n_values, 256), (int64_t)sms * 8);
            mix_sum_kernel<double, double, 1><<<blocks, 256, 0, s>>>(ptrs, n_inputs, n_values, (double *)out_dev);
        }
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}
