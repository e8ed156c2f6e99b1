// Shared host/device helpers of libpips_b200.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pips_b200.h"

namespace pips {

// error reporting (thread-local message behind pips_last_error())
int fail(const char* msg);
int fail_cuda(const char* where, cudaError_t e);

int sm_count();

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda,
// so it also loads on a machine without a driver -- the CPU-side ABI tests rely on that).
bool encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, void* base, const cuuint64_t* gdim,
                  const cuuint64_t* gstride_bytes, const cuuint32_t* box, const cuuint32_t* estride,
                  CUtensorMapSwizzle swizzle);

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

// v ~= hi + lo
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float gelu_exact(float v) {
    // nn.GELU() default (approximate='none'): 0.5 x (1 + erf(x / sqrt(2)))
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pips
