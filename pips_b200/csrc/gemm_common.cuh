// Shared pieces of the tcgen05 dense-layer kernels (gemm_tc.cu: one CTA per tile; gemm_tc2.cu: CTA pairs).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pips {

constexpr int BK = 64;                       // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;

struct GemmArgs {
    int M, N, K;                 // valid rows / cols, K multiple of 64
    const float* bias;           // [N]
    int epilogue;                // PIPS_EPI_*
    float* out_f32;              // BIAS / BIAS_RESID target (row stride ldo)
    int ldo;
    __nv_bfloat16* out_hi;       // BIAS_GELU target (row stride ldh); out_lo may be null
    __nv_bfloat16* out_lo;
    int ldh;
};


// Epilogue of one 32-column chunk of one output row: bias, then GELU + (hi, lo) split, or fp32 store with
// optional residual.  `col` and the chunk predicate are warp-uniform; `row_ok` is per thread.
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& args, const uint32_t (&v)[32], int row, bool row_ok, int col) {
    if (col >= args.N) return;                       // warp-uniform
    const bool full_chunk = col + 32 <= args.N;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 b;
        if (full_chunk) {
            b = __ldg(reinterpret_cast<const float4*>(args.bias + col + j));
        } else {
            b.x = col + j + 0 < args.N ? __ldg(args.bias + col + j + 0) : 0.f;
            b.y = col + j + 1 < args.N ? __ldg(args.bias + col + j + 1) : 0.f;
            b.z = col + j + 2 < args.N ? __ldg(args.bias + col + j + 2) : 0.f;
            b.w = col + j + 3 < args.N ? __ldg(args.bias + col + j + 3) : 0.f;
        }
        f[j + 0] = __uint_as_float(v[j + 0]) + b.x;
        f[j + 1] = __uint_as_float(v[j + 1]) + b.y;
        f[j + 2] = __uint_as_float(v[j + 2]) + b.z;
        f[j + 3] = __uint_as_float(v[j + 3]) + b.w;
    }
    if (row_ok && args.epilogue == PIPS_EPI_BIAS_GELU) {
        __nv_bfloat16* ph = args.out_hi + static_cast<size_t>(row) * args.ldh + col;
        __nv_bfloat16* pl = args.out_lo ? args.out_lo + static_cast<size_t>(row) * args.ldh + col : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float g0 = gelu_fast(f[j + 2 * e]);
                const float g1 = gelu_fast(f[j + 2 * e + 1]);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(g0), h1 = __float2bfloat16_rn(g1);
                hw[e] = pack_bf16(h0, h1);
                lw[e] = pack_bf16(__float2bfloat16_rn(g0 - __bfloat162float(h0)),
                                  __float2bfloat16_rn(g1 - __bfloat162float(h1)));
            }
            if (full_chunk) {
                *reinterpret_cast<uint4*>(ph + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                if (pl) *reinterpret_cast<uint4*>(pl + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            } else {
                for (int e = 0; e < 8; ++e) {
                    if (col + j + e < args.N) {
                        const uint32_t hh = hw[e >> 1], ll = lw[e >> 1];
                        reinterpret_cast<uint16_t*>(ph)[j + e] = (e & 1) ? (hh >> 16) : (hh & 0xffff);
                        if (pl) reinterpret_cast<uint16_t*>(pl)[j + e] = (e & 1) ? (ll >> 16) : (ll & 0xffff);
                    }
                }
            }
        }
    } else if (row_ok) {
        float* po = args.out_f32 + static_cast<size_t>(row) * args.ldo + col;
        const bool resid = args.epilogue == PIPS_EPI_BIAS_RESID;
        if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                if (resid) {
                    const float4 r = *reinterpret_cast<const float4*>(po + j);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(po + j) = o;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (col + j < args.N) po[j] = resid ? po[j] + f[j] : f[j];
        }
    }
}

}  // namespace pips
