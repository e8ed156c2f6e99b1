// Mixer dense layers on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )
//
// A (activations) and W (nn.Linear weight, (out,in) == N x K, K-major) are bf16.  In the
// "bf16x3" precision mode both operands are carried as a (hi, lo) bf16 pair with
// v ~= hi + lo, and the product is evaluated as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo with fp32
// accumulation in TMEM (relative error ~2^-16 per product instead of 2^-8), which is what the
// 1e-3 px parity target of the refinement loop needs (reference: fp32 nn.Linear,
// nets/pips.py:104-107,:115,:122).
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles; BN = 256 for large problems, 128 / 64 when
// 256-wide tiles would leave most SMs without a tile -- few particles: the demo clip, chained windows):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled K-chunks of A and W -> smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (M=128,N=256,K=16) per chunk/term
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       : tcgen05.ld the fp32 tile (one row per thread, two warps per lane quadrant),
//                                bias / GELU(erf) / residual / hi-lo split, vectorised stores
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "gemm_common.cuh"

namespace pips {

constexpr int BM = 128;
constexpr int GEMM_THREADS = 384;                 // 4 control warps + 8 epilogue warps
constexpr uint32_t A_TILE_BYTES = BM * BK * 2;     // 16 KB

template <int TERMS, int BN>
struct GemmCfg {
    static constexpr uint32_t kWTileBytes = BN * BK * 2;                  // 32 / 16 / 8 KB
    static constexpr int kOperandCopies = TERMS == 3 ? 2 : 1;             // hi (+ lo)
    static constexpr uint32_t kStageBytes = kOperandCopies * (A_TILE_BYTES + kWTileBytes);
    static constexpr uint32_t kFixedBytes = EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kFit = static_cast<int>((227u * 1024u - kFixedBytes) / kStageBytes);
    static constexpr int kStages = kFit > 8 ? 8 : kFit;                   // x3: 2 / 3 / 4 for BN = 256 / 128 / 64
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int TERMS, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const GemmArgs args) {
    using Cfg = GemmCfg<TERMS, BN>;
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t W_TILE_BYTES = Cfg::kWTileBytes;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle atoms need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* epi_stage = smem + kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPI_STAGE_BYTES);
    // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * kStages;
    const uint32_t tfull0 = empty0 + 8 * kStages;
    const uint32_t tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();

    const int tiles_m = (args.M + BM - 1) / BM;
    const int tiles_n = (args.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = args.K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_w_hi);
        if (TERMS == 3) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_w_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, GEMM_THREADS - 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                             // operands and the residual stream come from the previous kernel

    // Register budget per warpgroup: the control warps (0-3) need few, the epilogue warpgroups (4-7, 8-11) many.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * BM;
                const int n0 = (t % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
                    const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
                    const int k0 = kb * BK;
                    tma_load_2d(base, &map_a_hi, fb, k0, m0);
                    tma_load_2d(base + A_TILE_BYTES, &map_w_hi, fb, k0, n0);
                    if (TERMS == 3) {
                        tma_load_2d(base + A_TILE_BYTES + W_TILE_BYTES, &map_a_lo, fb, k0, m0);
                        tma_load_2d(base + 2 * A_TILE_BYTES + W_TILE_BYTES, &map_w_lo, fb, k0, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // whole warp, warp-uniform control flow (descriptors stay in uniform registers); one elected lane issues
        {
            const bool elected = elect_one();
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * as, aphase ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t a_hi = umma_desc_sw128(base);
                    const uint64_t w_hi = umma_desc_sw128(base + A_TILE_BYTES);
                    const uint64_t a_lo = umma_desc_sw128(base + A_TILE_BYTES + W_TILE_BYTES);
                    const uint64_t w_lo = umma_desc_sw128(base + 2 * A_TILE_BYTES + W_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);   // 32 B per K=16 step
                        if (elected) umma_f16(d_tmem, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                        if (TERMS == 3) {
                            if (elected) umma_f16(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
                            if (elected) umma_f16(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
                        }
                    }
                    if (elected) umma_commit(empty0 + 8 * stage);          // smem slot reusable once these MMAs retire
                    if (elected && kb == num_kb - 1) umma_commit(tfull0 + 8 * as);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ------------------------------------------------------------ epilogue (256 threads)
        // Two warps per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31); each takes half
        // of the tile's 256 columns.  Thread = one output row; tcgen05.ld of the next 32-column chunk is
        // in flight while the current one goes through bias / GELU / split / store.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint8_t* my_stage = epi_stage + (warp - 4) * 4096;
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int m0 = (t / tiles_n) * BM;
            const int n0 = (t % tiles_n) * BN + half * (BN / 2);
            const int row0 = m0 + q * 32;
            mbar_wait(tfull0 + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * BN + half * (BN / 2) + (static_cast<uint32_t>(q * 32) << 16);

            uint32_t va[32], vb[32];
            float bias[32];
            tmem_ld_32x32(taddr, va);
            load_bias_chunk(args, n0, bias);
            constexpr int NCH = BN / 64;                         // 32-column chunks per half tile: 4 / 2 / 1
#pragma unroll 1
            for (int c = 0; c < NCH; c += 2) {
                const int c0 = c * 32;
                tmem_ld_wait();
                if (c + 1 < NCH) tmem_ld_32x32(taddr + c0 + 32, vb);
                epilogue_chunk(args, va, bias, row0, n0 + c0, c + 1 < NCH ? n0 + c0 + 32 : -1, my_stage);
                __syncwarp();                                    // tcgen05.ld / wait are warp-collective
                if (c + 1 < NCH) {
                    tmem_ld_wait();
                    if (c + 2 < NCH) tmem_ld_32x32(taddr + c0 + 64, va);
                    epilogue_chunk(args, vb, bias, row0, n0 + c0 + 32, c + 2 < NCH ? n0 + c0 + 64 : -1, my_stage);
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * as);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------- host side

static bool make_operand_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld_elems, int box_rows) {
    // global tensor (innermost first): {K, rows}, row stride ld_elems*2 bytes; box {64, box_rows}, 128B swizzle
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_SWIZZLE_128B);
}

int gemm_tc_pair_dispatch(const void* a_hi, const void* a_lo, int lda, int a_rows, const void* w_hi, const void* w_lo, int ldw,
                          int w_rows, const GemmArgs& args, cudaStream_t st);       // gemm_tc2.cu

// 0: automatic, 1: pair kernel, 64/128/256: single-CTA kernel with that tile width
static int forced_tile() {
    const char* e = getenv("PIPS_B200_GEMM_TILE");
    if (!e || !e[0]) return 0;
    if (e[0] == 'p') return 1;
    const int v = atoi(e);
    return (v == 64 || v == 128 || v == 256) ? v : 0;
}

int launch_single(int bn, bool x3, const void* a_hi, const void* a_lo, int lda, int a_rows, const void* w_hi, const void* w_lo,
                  int ldw, int w_rows, const GemmArgs& args, cudaStream_t st);

static bool use_pair_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PIPS_B200_GEMM_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace pips

using namespace pips;

extern "C" int pips_gemm_tc(const void* a_hi, const void* a_lo, int lda, int a_rows,
                            const void* w_hi, const void* w_lo, int ldw, int w_rows,
                            int M, int N, int K, const float* bias, int epilogue,
                            float* out_f32, int ldo, void* out_hi, void* out_lo, int ldh,
                            void* stream) {
    if (K % BK != 0 || K <= 0) return fail("pips_gemm_tc: K must be a positive multiple of 64");
    if (M <= 0 || N <= 0) return fail("pips_gemm_tc: empty problem");
    if (a_rows < M || w_rows < N) return fail("pips_gemm_tc: operand allocations smaller than the problem");
    if ((lda % 8) || (ldw % 8)) return fail("pips_gemm_tc: leading dimensions must be multiples of 8 elements (16 B)");
    if (!bias) return fail("pips_gemm_tc: bias is required");
    const bool x3 = a_lo != nullptr && w_lo != nullptr;
    if ((a_lo != nullptr) != (w_lo != nullptr)) return fail("pips_gemm_tc: need both or neither lo operands");
    if (epilogue == PIPS_EPI_BIAS_GELU) {
        if (!out_hi || (ldh % 8)) return fail("pips_gemm_tc: GELU epilogue needs out_hi with ldh % 8 == 0");
    } else if (epilogue == PIPS_EPI_BIAS || epilogue == PIPS_EPI_BIAS_RESID) {
        if (!out_f32 || (ldo % 4)) return fail("pips_gemm_tc: fp32 epilogue needs out_f32 with ldo % 4 == 0");
    } else {
        return fail("pips_gemm_tc: unknown epilogue");
    }
    GemmArgs args;
    args.M = M; args.N = N; args.K = K; args.bias = bias; args.epilogue = epilogue;
    args.out_f32 = out_f32; args.ldo = ldo;
    args.out_hi = static_cast<__nv_bfloat16*>(out_hi); args.out_lo = static_cast<__nv_bfloat16*>(out_lo); args.ldh = ldh;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Tile shape.  Every variant accumulates each output element over K in the same order (same MMA K-steps, same
    // hi/lo term order), so the choice changes the schedule, not the bits (tests/test_kernels_gpu.py).
    //   pair  256x256 per CTA pair (cta_group::2): least operand traffic; needs enough pair tiles to fill the GPU
    //   128 x {256,128,64} single CTA: when 256-wide tiles would leave most SMs idle (few particles)
    const int force = forced_tile();                  // PIPS_B200_GEMM_TILE = pair | 256 | 128 | 64 (tests, profiling)
    const bool pair_ok = use_pair_kernel() && M > 128 && (a_rows % 256) == 0 && (w_rows % 256) == 0;
    const int pair_tiles = ((M + 255) / 256) * ((N + 255) / 256);
    const int sms = sm_count();
    int bn = 256;
    bool pair = pair_ok && pair_tiles * 2 >= (sms * 3) / 4;          // at least 3/4 of the SMs get a tile
    if (!pair) {
        const int tm = (M + BM - 1) / BM;
        while (bn > 64 && tm * ((N + bn - 1) / bn) < (sms * 3) / 4) bn >>= 1;
    }
    if (force == 1) pair = pair_ok;
    else if (force > 1) { pair = false; bn = force; }
    if ((w_rows % bn) != 0) { bn = 256; pair = pair && pair_ok; }   // TMA boxes must stay inside the weight allocation
    if (pair)
        return gemm_tc_pair_dispatch(a_hi, x3 ? a_lo : nullptr, lda, a_rows, w_hi, x3 ? w_lo : nullptr, ldw, w_rows, args, st);
    return launch_single(bn, x3, a_hi, a_lo, lda, a_rows, w_hi, w_lo, ldw, w_rows, args, st);
}

namespace pips {

template <int TERMS, int BN>
static int launch_variant(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mw_hi, const CUtensorMap& mw_lo,
                          const GemmArgs& args, cudaStream_t st) {
    using Cfg = GemmCfg<TERMS, BN>;
    static bool attr[kMaxDevices] = {};
    {
        cudaError_t e = ensure_dyn_smem(gemm_tc_kernel<TERMS, BN>, attr, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail_cuda("pips_gemm_tc: smem attribute", e);
    }
    const int tiles = ((args.M + BM - 1) / BM) * ((args.N + BN - 1) / BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    cudaError_t e = launch_pdl(gemm_tc_kernel<TERMS, BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ma_hi, ma_lo, mw_hi, mw_lo, args);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_gemm_tc: launch", e);
}

int launch_single(int bn, bool x3, const void* a_hi, const void* a_lo, int lda, int a_rows, const void* w_hi, const void* w_lo,
                  int ldw, int w_rows, const GemmArgs& args, cudaStream_t st) {
    const int K = args.K;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    if (!make_operand_map(&ma_hi, a_hi, a_rows, K, lda, BM)) return fail("pips_gemm_tc: tensor map (A hi) failed");
    if (!make_operand_map(&mw_hi, w_hi, w_rows, K, ldw, bn)) return fail("pips_gemm_tc: tensor map (W hi) failed");
    ma_lo = ma_hi;
    mw_lo = mw_hi;
    if (x3) {
        if (!make_operand_map(&ma_lo, a_lo, a_rows, K, lda, BM)) return fail("pips_gemm_tc: tensor map (A lo) failed");
        if (!make_operand_map(&mw_lo, w_lo, w_rows, K, ldw, bn)) return fail("pips_gemm_tc: tensor map (W lo) failed");
    }
    int rc;
    if (x3) rc = bn == 256 ? launch_variant<3, 256>(ma_hi, ma_lo, mw_hi, mw_lo, args, st)
               : bn == 128 ? launch_variant<3, 128>(ma_hi, ma_lo, mw_hi, mw_lo, args, st)
                           : launch_variant<3, 64>(ma_hi, ma_lo, mw_hi, mw_lo, args, st);
    else rc = bn == 256 ? launch_variant<1, 256>(ma_hi, ma_lo, mw_hi, mw_lo, args, st)
            : bn == 128 ? launch_variant<1, 128>(ma_hi, ma_lo, mw_hi, mw_lo, args, st)
                        : launch_variant<1, 64>(ma_hi, ma_lo, mw_hi, mw_lo, args, st);
    if (rc) return rc;
    cudaError_t e;
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda("pips_gemm_tc: launch", e);
    return 0;
}

}  // namespace pips
