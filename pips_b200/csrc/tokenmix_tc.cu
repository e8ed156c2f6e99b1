// Token mixing with the two small contractions on the tensor cores.
//
//   x (8 frames x 512 channels per track)
//   y = LN1(x);  h[c, j] = gelu(sum_s y[s, c] W1[j, s] + b1[j]);  z[c, s'] = sum_j h[c, j] W2[s', j] + b2[s'];
//   x += z;  out = LN2(x)                                                    (nets/pips.py:117, :93-109, :100)
//
// K = 8 and K = 32 are far too small for a conventional tensor-core GEMM, but with the CHANNELS as the MMA's M
// dimension each track is four M = 128 tiles and the per-channel MLP becomes
//   GEMM1  H (128 x 32) = A1 (128 x 16) . B^T      A1 row c = [y_hi[0..8) | y_lo[0..8)]   (bf16, K-major, no swizzle)
//   GEMM2  Z (128 x 16) = G  (128 x 64) . B^T      G  row c = [g_hi[0..32) | g_lo[0..32)]
// where the bf16x3 split is folded into K: B = [w_hi | w_hi] gives hi*hi + lo*hi in one MMA and B' = [w_lo | 0]
// adds hi*lo.  8 tcgen05.mma per tile replace 512 FMAs per channel; LayerNorm, GELU, the hi/lo splits and all
// global traffic stay exactly as in the CUDA-core kernel (thread t owns channels 4t..4t+3 == lane t of tile
// 0..3, so loads and stores are 16-byte vectors).  One CTA = 128 threads = the 128 TMEM lanes; two CTAs per SM
// (256 TMEM columns each) overlap each other's MMA round trips.
#include "mixer_common.cuh"
#include "ptx.cuh"

namespace pips {

constexpr int TT_THREADS = 128;
constexpr uint32_t TT_A1_TILE = 128 * 32;            // 4 KB: 128 rows x K=16 bf16
constexpr uint32_t TT_G_TILE = 128 * 128;            // 16 KB: 128 rows x K=64 bf16
constexpr uint32_t TT_OFF_A1 = 0;
constexpr uint32_t TT_OFF_G = TT_OFF_A1 + 4 * TT_A1_TILE;          // 16 KB
constexpr uint32_t TT_OFF_B1A = TT_OFF_G + 4 * TT_G_TILE;          // [w1_hi | w1_hi]  32 rows x 32 B
constexpr uint32_t TT_OFF_B1B = TT_OFF_B1A + 1024;                 // [w1_lo | 0]
constexpr uint32_t TT_OFF_B2A = TT_OFF_B1B + 1024;                 // [w2_hi | w2_hi]  16 rows x 128 B
constexpr uint32_t TT_OFF_B2B = TT_OFF_B2A + 2048;                 // [w2_lo | 0]
constexpr uint32_t TT_OFF_MISC = TT_OFF_B2B + 2048;                // barriers, tmem slot, reductions, biases
constexpr uint32_t TT_SMEM = TT_OFF_MISC + 512 + 1024;

// K-major operand without swizzle: 8-row x 16-byte core matrices; element (r, k) lives at
// (r / 8) * sbo + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;       // between core matrices along K
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;       // between 8-row groups
    d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ uint32_t core_off(int r, int kcore, uint32_t sbo) {
    return static_cast<uint32_t>(r >> 3) * sbo + static_cast<uint32_t>(kcore) * 128 + static_cast<uint32_t>(r & 7) * 16;
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

__global__ void __launch_bounds__(TT_THREADS, 2)
tokenmix_tc_kernel(float* __restrict__ x, int seqs, const float* __restrict__ ln1_w, const float* __restrict__ ln1_b,
                   const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                   const float* __restrict__ b2, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
                   __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TT_OFF_MISC);
    const uint32_t bar1 = smem_u32(bars), bar2 = bar1 + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float (*red)[8] = reinterpret_cast<float (*)[8]>(smem + TT_OFF_MISC + 32);
    float* s_b1 = reinterpret_cast<float*>(smem + TT_OFF_MISC + 32 + 128);
    float* s_b2 = s_b1 + 32;

    const int t = threadIdx.x, warp = t >> 5;

    // ---- one-time setup: weight operands, barriers, TMEM
    for (int i = t; i < 32 * 16; i += TT_THREADS) {                 // GEMM1 B operands: row j, k in [0,16)
        const int j = i >> 4, k = i & 15;
        __nv_bfloat16 h, l;
        split_bf16(w1[j * 8 + (k & 7)], h, l);
        const uint32_t off = core_off(j, k >> 3, 256) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + TT_OFF_B1A + off) = h;                                   // [hi | hi]
        *reinterpret_cast<__nv_bfloat16*>(smem + TT_OFF_B1B + off) = k < 8 ? l : __float2bfloat16_rn(0.f); // [lo | 0]
    }
    for (int i = t; i < 16 * 64; i += TT_THREADS) {                 // GEMM2 B operands: row s' (8 real of 16), k in [0,64)
        const int r = i >> 6, k = i & 63;
        __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
        if (r < 8) split_bf16(w2[r * 32 + (k & 31)], h, l);
        const uint32_t off = core_off(r, k >> 3, 1024) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + TT_OFF_B2A + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(smem + TT_OFF_B2B + off) = k < 32 ? l : __float2bfloat16_rn(0.f);
    }
    if (t < 32) s_b1[t] = b1[t];
    if (t < 8) s_b2[t] = b2[t];
    if (t == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), 256);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 32);
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, 16);

    const float4 g1 = *reinterpret_cast<const float4*>(ln1_w + t * 4), c1 = *reinterpret_cast<const float4*>(ln1_b + t * 4);
    const float4 g2 = *reinterpret_cast<const float4*>(ln2_w + t * 4), c2 = *reinterpret_cast<const float4*>(ln2_b + t * 4);

    uint32_t parity = 0;
    for (int seq = blockIdx.x; seq < seqs; seq += gridDim.x, parity ^= 1) {
        const size_t base = static_cast<size_t>(seq) * 8 * 512 + t * 4;
        float xv[8][4], yv[8][4];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(x + base + s * 512);
            xv[s][0] = v.x; xv[s][1] = v.y; xv[s][2] = v.z; xv[s][3] = v.w;
        }
        layernorm8(xv, yv, g1, c1, red);

        // ---- A1: row t of tile i = [y_hi over the 8 frames | y_lo over the 8 frames] of channel 4t+i
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = yv[2 * e][i], b = yv[2 * e + 1][i];
                hw[e] = cvt_bf16x2(a, b);
                lw[e] = cvt_bf16x2(a - __uint_as_float(hw[e] << 16), b - __uint_as_float(hw[e] & 0xffff0000u));
            }
            uint8_t* row = smem + TT_OFF_A1 + i * TT_A1_TILE;
            *reinterpret_cast<uint4*>(row + core_off(t, 0, 256)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(row + core_off(t, 1, 256)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const uint64_t ba = umma_desc_nosw(sbase + TT_OFF_B1A, 128, 256), bb = umma_desc_nosw(sbase + TT_OFF_B1B, 128, 256);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t a = umma_desc_nosw(sbase + TT_OFF_A1 + i * TT_A1_TILE, 128, 256);
                umma_f16(tmem_base + i * 32, a, ba, idesc1, 0);
                umma_f16(tmem_base + i * 32, a, bb, idesc1, 1);
            }
            umma_commit(bar1);
        }
        mbar_wait(bar1, parity);
        tc_fence_after();

        // ---- GELU on H, G row t of tile i = [g_hi over j | g_lo over j]
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + i * 32, v);
            tmem_ld_wait();
            uint8_t* row = smem + TT_OFF_G + i * TT_G_TILE;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 8 * p + 2 * e;
                    const float2 g = gelu_fast2(make_float2(__uint_as_float(v[j]) + s_b1[j], __uint_as_float(v[j + 1]) + s_b1[j + 1]));
                    hw[e] = cvt_bf16x2(g.x, g.y);
                    const float2 lo = fma2(make_float2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u)),
                                           bcast2(-1.0f), g);
                    lw[e] = cvt_bf16x2(lo.x, lo.y);
                }
                *reinterpret_cast<uint4*>(row + core_off(t, p, 1024)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(row + core_off(t, p + 4, 1024)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t d = tmem_base + 128 + i * 16;
                const uint32_t ga = sbase + TT_OFF_G + i * TT_G_TILE;
#pragma unroll
                for (int k = 0; k < 4; ++k)                      // [g_hi | g_lo] . [w2_hi | w2_hi]
                    umma_f16(d, umma_desc_nosw(ga + k * 256, 128, 1024), umma_desc_nosw(sbase + TT_OFF_B2A + k * 256, 128, 1024), idesc2, k != 0);
#pragma unroll
                for (int k = 0; k < 2; ++k)                      // g_hi . w2_lo
                    umma_f16(d, umma_desc_nosw(ga + k * 256, 128, 1024), umma_desc_nosw(sbase + TT_OFF_B2B + k * 256, 128, 1024), idesc2, 1);
            }
            umma_commit(bar2);
        }
        mbar_wait(bar2, parity);
        tc_fence_after();

        // ---- residual, LN2, outputs
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v[8];
            tmem_ld_32x8(lane_addr + 128 + i * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < 8; ++s) xv[s][i] += __uint_as_float(v[s]) + s_b2[s];
        }
#pragma unroll
        for (int s = 0; s < 8; ++s)
            *reinterpret_cast<float4*>(x + base + s * 512) = make_float4(xv[s][0], xv[s][1], xv[s][2], xv[s][3]);
        layernorm8(xv, yv, g2, c2, red);
#pragma unroll
        for (int s = 0; s < 8; ++s) store_row4(yv[s], base + s * 512, y_hi, y_lo, nullptr);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace pips

using namespace pips;

// called by pips_tokenmix (mixer_simt.cu) for the bf16 / bf16x3 precisions
int tokenmix_tc_launch(float* x, int seqs, const float* ln1_w, const float* ln1_b, const float* w1, const float* b1, const float* w2,
                       const float* b2, const float* ln2_w, const float* ln2_b, void* y_hi, void* y_lo, cudaStream_t st) {
    static bool attr[kMaxDevices] = {};
    {
        cudaError_t e = ensure_dyn_smem(tokenmix_tc_kernel, attr, static_cast<int>(TT_SMEM));
        if (e != cudaSuccess) return fail_cuda("pips_tokenmix (tc): smem attribute", e);
    }
    const int cap = 2 * sm_count();
    const int grid = seqs < cap ? seqs : cap;
    tokenmix_tc_kernel<<<grid, TT_THREADS, TT_SMEM, st>>>(x, seqs, ln1_w, ln1_b, w1, b1, w2, b2, ln2_w, ln2_b,
                                                          static_cast<__nv_bfloat16*>(y_hi), static_cast<__nv_bfloat16*>(y_lo));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_tokenmix (tc): launch", e);
}

// explicit entry point of the tensor-core variant (same contract as pips_tokenmix with bf16 outputs)
extern "C" int pips_tokenmix_tc(float* x, int seqs, const float* ln1_w, const float* ln1_b, const float* w1, const float* b1,
                                const float* w2, const float* b2, const float* ln2_w, const float* ln2_b, void* y_hi, void* y_lo,
                                void* stream) {
    if (!x || !ln1_w || !ln1_b || !w1 || !b1 || !w2 || !b2 || !ln2_w || !ln2_b || !y_hi) return fail("pips_tokenmix_tc: null pointer");
    if (seqs <= 0) return fail("pips_tokenmix_tc: no sequences");
    return tokenmix_tc_launch(x, seqs, ln1_w, ln1_b, w1, b1, w2, b2, ln2_w, ln2_b, y_hi, y_lo, static_cast<cudaStream_t>(stream));
}
