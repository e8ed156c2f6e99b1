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
// global traffic stay on the CUDA cores (thread t owns channels 4t..4t+3 == lane t of tile 0..3, so loads and
// stores are 16-byte vectors).  One CTA = 128 threads = the 128 TMEM lanes.
//
// Round 2 (this version): the round-1 kernel ran one track at a time per CTA with every MMA round trip exposed and
// only 2 CTAs per SM (87 KB of smem, 256 TMEM columns): 0.147 ms per launch at 4096 tracks, slower than the CUDA-core
// kernel (0.117 ms).  Now
//   * the G operand is double-buffered per TILE (2 x 16 KB instead of 4 x 16 KB) and Z_i overwrites the TMEM columns
//     of H_i (consumed by then): 55 KB of smem and 128 TMEM columns per CTA -> 4 CTAs (16 warps) per SM;
//   * the second contraction is software-pipelined over the four tiles: GELU(H_{i+1}) runs on the CUDA cores while
//     the tensor cores execute GEMM2 of tile i (one mbarrier per G buffer);
//   * the 8-row LayerNorm statistics use a transposing butterfly (9 shuffles per reduction instead of 40).
#include "mixer_common.cuh"
#include "ptx.cuh"

namespace pips {

constexpr int TT_THREADS = 128;
constexpr int TT_CTAS_PER_SM = 4;
constexpr uint32_t TT_TMEM_COLS = 128;               // H_i: columns [32 i, 32 i + 32); Z_i overwrites [32 i, 32 i + 16)
constexpr uint32_t TT_A1_TILE = 128 * 32;            // 4 KB: 128 rows x K=16 bf16
constexpr uint32_t TT_G_TILE = 128 * 128;            // 16 KB: 128 rows x K=64 bf16
constexpr uint32_t TT_OFF_A1 = 0;
constexpr uint32_t TT_OFF_G = TT_OFF_A1 + 4 * TT_A1_TILE;          // two G buffers (tiles i and i+1 in flight)
constexpr uint32_t TT_OFF_B1A = TT_OFF_G + 2 * TT_G_TILE;          // [w1_hi | w1_hi]  32 rows x 32 B
constexpr uint32_t TT_OFF_B1B = TT_OFF_B1A + 1024;                 // [w1_lo | 0]
constexpr uint32_t TT_OFF_B2A = TT_OFF_B1B + 1024;                 // [w2_hi | w2_hi]  16 rows x 128 B
constexpr uint32_t TT_OFF_B2B = TT_OFF_B2A + 2048;                 // [w2_lo | 0]
constexpr uint32_t TT_OFF_MISC = TT_OFF_B2B + 2048;                // barriers, tmem slot, reductions, biases
constexpr uint32_t TT_SMEM = TT_OFF_MISC + 512 + 128;              // + alignment slack; 4 CTAs: 4 x (55.9 + 1) KB <= 228 KB

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

// ---- 8-row statistics with a transposing butterfly: every lane contributes 8 partial sums (one per frame); three
// exchange steps halve the number of live values while doubling the lanes summed, two more finish the warp:
// 9 shuffles instead of 8 x 5.  Lane l ends with the warp total of row  4*bit4(l) + 2*bit3(l) + bit2(l).
__device__ __forceinline__ float warp_rowsum8(const float (&v)[8]) {
    const int lane = threadIdx.x & 31;
    float a[4], b[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? a[i] : a[i + 2], keep = up ? a[i + 2] : a[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    const bool up = lane & 4;
    float c = (up ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, up ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

// CTA-wide sums of 8 per-thread values (128 threads): v[s] <- total over the CTA.  `red` is [4 warps][8 rows].
__device__ __forceinline__ void cta_rowsum8(float (&v)[8], float (*red)[8]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float c = warp_rowsum8(v);
    __syncthreads();                                        // previous readers of `red` are done
    if ((lane & 3) == 0) red[warp][(lane >> 2) & 7] = c;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 r0 = *reinterpret_cast<const float4*>(&red[0][4 * h]), r1 = *reinterpret_cast<const float4*>(&red[1][4 * h]);
        const float4 r2 = *reinterpret_cast<const float4*>(&red[2][4 * h]), r3 = *reinterpret_cast<const float4*>(&red[3][4 * h]);
        v[4 * h + 0] = (r0.x + r1.x) + (r2.x + r3.x);
        v[4 * h + 1] = (r0.y + r1.y) + (r2.y + r3.y);
        v[4 * h + 2] = (r0.z + r1.z) + (r2.z + r3.z);
        v[4 * h + 3] = (r0.w + r1.w) + (r2.w + r3.w);
    }
}

// two-pass LayerNorm over the 512 channels of the 8 rows held as x[s][0..3] (biased variance, eps 1e-5); the
// per-element arithmetic runs on packed fp32 pairs (FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ void layernorm8_fast(const float (&x)[8][4], float (&y)[8][4], const float4 g, const float4 b, float (*red)[8]) {
    float m[8], q[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const float2 t = add2(make_float2(x[s][0], x[s][1]), make_float2(x[s][2], x[s][3]));
        m[s] = t.x + t.y;
    }
    cta_rowsum8(m, red);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        m[s] *= (1.0f / 512.0f);
        const float2 nm = bcast2(-m[s]);
        const float2 d0 = add2(make_float2(x[s][0], x[s][1]), nm), d1 = add2(make_float2(x[s][2], x[s][3]), nm);
        const float2 t = fma2(d1, d1, mul2(d0, d0));
        q[s] = t.x + t.y;
    }
    cta_rowsum8(q, red);
    const float2 g0 = make_float2(g.x, g.y), g1 = make_float2(g.z, g.w), b0 = make_float2(b.x, b.y), b1 = make_float2(b.z, b.w);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const float r = rsqrtf(q[s] * (1.0f / 512.0f) + 1e-5f);
        const float2 nm = bcast2(-m[s]), r2 = bcast2(r);
        const float2 o0 = fma2(mul2(add2(make_float2(x[s][0], x[s][1]), nm), r2), g0, b0);
        const float2 o1 = fma2(mul2(add2(make_float2(x[s][2], x[s][3]), nm), r2), g1, b1);
        y[s][0] = o0.x; y[s][1] = o0.y; y[s][2] = o1.x; y[s][3] = o1.y;
    }
}

// elect.sync needs the whole warp converged: called at a point where all 128 threads arrive, evaluated in warp 0 only
__device__ __forceinline__ bool elect_one_in_warp0(int warp) {
    bool e = false;
    if (warp == 0) e = elect_one();
    return e;
}

__global__ void __launch_bounds__(TT_THREADS, TT_CTAS_PER_SM)
tokenmix_tc_kernel(float* __restrict__ x, int seqs, const float* __restrict__ ln1_w, const float* __restrict__ ln1_b,
                   const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                   const float* __restrict__ b2, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
                   __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TT_OFF_MISC);
    const uint32_t bar_h = smem_u32(bars);                  // GEMM1 of all four tiles done
    const uint32_t bar_g0 = bar_h + 8;                      // GEMM2 that read G buffer b done (bar_g0 + 8 b)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
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
        mbar_init(bar_h, 1);
        mbar_init(bar_g0, 1);
        mbar_init(bar_g0 + 8, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), TT_TMEM_COLS);
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
    const uint64_t d_b1a = umma_desc_nosw(sbase + TT_OFF_B1A, 128, 256), d_b1b = umma_desc_nosw(sbase + TT_OFF_B1B, 128, 256);
    const bool elected = warp == 0 && elect_one_in_warp0(warp);   // the lane of warp 0 that issues the MMAs

    pdl_wait();                                             // x is written by the previous kernel (weights above are constants)
    // x of the track after the current one is fetched into registers while the current track's GELU phase runs; the
    // current x is dropped after LayerNorm 1 and parked in shared memory for the residual, so that both fit 128 registers
    float4 nx[8];
    if (static_cast<int>(blockIdx.x) < seqs) {
        const float* src = x + static_cast<size_t>(blockIdx.x) * 8 * 512 + t * 4;
#pragma unroll
        for (int s = 0; s < 8; ++s) nx[s] = *reinterpret_cast<const float4*>(src + s * 512);
    }
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(ln1_w + t * 4)), c1 = __ldg(reinterpret_cast<const float4*>(ln1_b + t * 4));

    uint32_t parity = 0;                                    // of bar_h: one completion per track
    for (int seq = blockIdx.x; seq < seqs; seq += gridDim.x, parity ^= 1) {
        const size_t base = static_cast<size_t>(seq) * 8 * 512 + t * 4;
        {
            float xv[8][4], yv[8][4];
#pragma unroll
            for (int s = 0; s < 8; ++s) { xv[s][0] = nx[s].x; xv[s][1] = nx[s].y; xv[s][2] = nx[s].z; xv[s][3] = nx[s].w; }
            layernorm8_fast(xv, yv, g1, c1, red);

            // ---- A1: row t of tile i = [y_hi over the 8 frames | y_lo over the 8 frames] of channel 4t+i
            // (the tensor cores finished reading the previous track's A1 long ago: its bar_h wait was passed)
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
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {                                     // warp-uniform: descriptors in uniform registers, one lane issues
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t a = umma_desc_nosw(sbase + TT_OFF_A1 + i * TT_A1_TILE, 128, 256);
                if (elected) {
                    umma_f16(tmem_base + i * 32, a, d_b1a, idesc1, 0);
                    umma_f16(tmem_base + i * 32, a, d_b1b, idesc1, 1);
                }
            }
            if (elected) umma_commit(bar_h);
        }
        if (seq + static_cast<int>(gridDim.x) < seqs) {      // next track's x: in flight during the GELU phase
            const float* src = x + static_cast<size_t>(seq + gridDim.x) * 8 * 512 + t * 4;
#pragma unroll
            for (int s = 0; s < 8; ++s) nx[s] = *reinterpret_cast<const float4*>(src + s * 512);
        }
        mbar_wait(bar_h, parity);
        tc_fence_after();
        // GEMM1 is done with the A1 tiles: their 16 KB now take a copy of this track's x (cp.async, no registers held)
        // for the residual at the end -- re-reading it from L2 there exposed ~500 cycles of latency per track
        {
            const uint32_t dst = sbase + TT_OFF_A1 + t * 16;
            const float* src = x + base;
#pragma unroll
            for (int sr = 0; sr < 8; ++sr)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + sr * 2048), "l"(src + sr * 512) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        // ---- per tile: GELU(H_i) -> G buffer (i & 1) -> GEMM2_i (Z_i lands on H_i's first 16 columns) while the
        // CUDA cores already work on H_{i+1}.  Each G-buffer barrier completes twice per track (parities 0, 1).
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + i * 32, v);
            if (i >= 2) mbar_wait(bar_g0 + 8 * (i & 1), 0);  // GEMM2_{i-2} has finished reading this G buffer
            tmem_ld_wait();
            uint8_t* row = smem + TT_OFF_G + (i & 1) * TT_G_TILE;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 8 * p + 2 * e;
                    const float2 bj = *reinterpret_cast<const float2*>(s_b1 + j);
                    const float2 g = gelu_fast2_abs(add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), bj));
                    hw[e] = cvt_bf16x2(g.x, g.y);
                    const float2 lo = fma2(make_float2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u)),
                                           bcast2(-1.0f), g);
                    lw[e] = cvt_bf16x2(lo.x, lo.y);
                }
                *reinterpret_cast<uint4*>(row + core_off(t, p, 1024)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(row + core_off(t, p + 4, 1024)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            fence_proxy_async_smem();
            tc_fence_before();                               // H_i fully read by this thread before Z_i may overwrite it
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                const uint32_t d = tmem_base + i * 32;
                const uint32_t ga = sbase + TT_OFF_G + (i & 1) * TT_G_TILE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {                    // [g_hi | g_lo] . [w2_hi | w2_hi]
                    const uint64_t da = umma_desc_nosw(ga + k * 256, 128, 1024), db = umma_desc_nosw(sbase + TT_OFF_B2A + k * 256, 128, 1024);
                    if (elected) umma_f16(d, da, db, idesc2, k != 0);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {                    // g_hi . w2_lo
                    const uint64_t da = umma_desc_nosw(ga + k * 256, 128, 1024), db = umma_desc_nosw(sbase + TT_OFF_B2B + k * 256, 128, 1024);
                    if (elected) umma_f16(d, da, db, idesc2, 1);
                }
                if (elected) umma_commit(bar_g0 + 8 * (i & 1));
            }
        }
        // ---- residual, LN2, outputs
        float xv[8][4];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int sr = 0; sr < 8; ++sr) {                     // this thread's own copies: no CTA-wide barrier needed
            const float4 v = *reinterpret_cast<const float4*>(smem + TT_OFF_A1 + sr * 2048 + t * 16);
            xv[sr][0] = v.x; xv[sr][1] = v.y; xv[sr][2] = v.z; xv[sr][3] = v.w;
        }
        mbar_wait(bar_g0, 1);                                // GEMM2 of tiles 2 and 3 (and therefore of all) done
        mbar_wait(bar_g0 + 8, 1);
        tc_fence_after();
        {
            uint32_t z[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) tmem_ld_32x8(lane_addr + i * 32, z[i]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int s = 0; s < 8; ++s) xv[s][i] += __uint_as_float(z[i][s]) + s_b2[s];
        }
        tc_fence_before();                                   // Z read before the next track's GEMM1 overwrites the columns
#pragma unroll
        for (int s = 0; s < 8; ++s)
            *reinterpret_cast<float4*>(x + base + s * 512) = make_float4(xv[s][0], xv[s][1], xv[s][2], xv[s][3]);
        {
            float yv[8][4];
            const float4 g2 = __ldg(reinterpret_cast<const float4*>(ln2_w + t * 4)), c2 = __ldg(reinterpret_cast<const float4*>(ln2_b + t * 4));
            layernorm8_fast(xv, yv, g2, c2, red);            // its __syncthreads also order the Z reads before the next GEMM1
#pragma unroll
            for (int s = 0; s < 8; ++s) store_row4(yv[s], base + s * 512, y_hi, y_lo, nullptr);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TT_TMEM_COLS);
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
    const int cap = TT_CTAS_PER_SM * sm_count();
    const int grid = seqs < cap ? seqs : cap;
    cudaError_t e = launch_pdl(tokenmix_tc_kernel, dim3(grid), dim3(TT_THREADS), TT_SMEM, st, x, seqs, ln1_w, ln1_b, w1, b1, w2, b2, ln2_w,
                               ln2_b, static_cast<__nv_bfloat16*>(y_hi), static_cast<__nv_bfloat16*>(y_lo));
    if (e == cudaSuccess) e = cudaGetLastError();
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
