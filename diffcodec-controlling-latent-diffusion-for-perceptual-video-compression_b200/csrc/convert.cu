// convert.cu -- element-type conversion for the host-buffer path (sm_100a): decoded video is 8-bit and flows / metrics
// travel well as half precision, so the PCIe upload of dcb's host entry point shrinks from 24 to 9 bytes per pixel and
// the download from 12 to 6 (bf16 result); the splat itself still runs in fp32 on the device.
//
//   dst[i] = (T_dst)(scale * (float)src[i])        src: U8 | F16 | BF16 | F32      dst: F32 | BF16 | F16
//
// No counterpart in the reference: its dataset code does this on the host (`.astype(np.float32) / 255.0`,
// controlnet/dataset.py) before the upload.
#include "dcb_common.cuh"

#include <cuda_fp16.h>

namespace dcb {

template <class S> __device__ __forceinline__ float to_f32(S v);
template <> __device__ __forceinline__ float to_f32<unsigned char>(unsigned char v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <class D> __device__ __forceinline__ D from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// 8 consecutive elements per thread per step (one 8..32-byte load, one 16..32-byte store), grid-stride
template <class S, class D>
__global__ void __launch_bounds__(256) k_convert(const S* __restrict__ src, D* __restrict__ dst, long long n, float scale, int vec) {
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        struct alignas(sizeof(S) * 8) SV { S v[8]; };
        struct alignas(sizeof(D) * 8) DV { D v[8]; };
        const long long n8 = n / 8;
        for (; i < n8; i += stride) {
            const SV in = reinterpret_cast<const SV*>(src)[i];
            DV out;
#pragma unroll
            for (int k = 0; k < 8; ++k) out.v[k] = from_f32<D>(mul_rn(to_f32<S>(in.v[k]), scale));
            reinterpret_cast<DV*>(dst)[i] = out;
        }
        i = n8 * 8 + ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    }
    for (; i < n; i += stride) dst[i] = from_f32<D>(mul_rn(to_f32<S>(src[i]), scale));
}

template <class S, class D> static int launch_convert(const void* src, void* dst, long long n, float scale, cudaStream_t st) {
    const bool vec = ((uintptr_t)src % (sizeof(S) * 8) == 0) && ((uintptr_t)dst % (sizeof(D) * 8) == 0);
    long long blocks = (n / 8 + 255) / 256;
    const long long cap = (long long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    DCB_CHECK_CUDA(launch_pdl(k_convert<S, D>, dim3((unsigned)blocks), dim3(256), 0, st, (const S*)src, (D*)dst, n, scale, vec ? 1 : 0));
    count_launch();
    return DCB_OK;
}

template <class S> static int convert_dst(const void* src, void* dst, int dst_dtype, long long n, float scale, cudaStream_t st) {
    switch (dst_dtype) {
        case DCB_F32: return launch_convert<S, float>(src, dst, n, scale, st);
        case DCB_BF16: return launch_convert<S, __nv_bfloat16>(src, dst, n, scale, st);
        case DCB_F16: return launch_convert<S, __half>(src, dst, n, scale, st);
    }
    return set_error(DCB_E_DTYPE, "dcb_convert: destination must be F32, BF16 or F16, got %d", dst_dtype);
}

int convert_impl(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, float scale, cudaStream_t st) {
    if (n == 0) return DCB_OK;
    switch (src_dtype) {
        case DCB_U8: return convert_dst<unsigned char>(src, dst, dst_dtype, n, scale, st);
        case DCB_F16: return convert_dst<__half>(src, dst, dst_dtype, n, scale, st);
        case DCB_BF16: return convert_dst<__nv_bfloat16>(src, dst, dst_dtype, n, scale, st);
        case DCB_F32: return convert_dst<float>(src, dst, dst_dtype, n, scale, st);
    }
    return set_error(DCB_E_DTYPE, "dcb_convert: source must be U8, F16, BF16 or F32, got %d", src_dtype);
}

}  // namespace dcb
