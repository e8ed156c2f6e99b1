// Mixer dense layers on CTA pairs: tcgen05.mma.cta_group::2, one 256x256 output tile per pair.
//
// Same contract and epilogues as gemm_tc.cu, but two CTAs of a cluster (the two SMs of a TPC) execute
// every MMA together: each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256
// output features), the tensor cores of both SMs read both halves, and each CTA ends up with its 128
// rows x 256 columns of the accumulator in its own TMEM.  Per SM and K-chunk this moves 64 KB (bf16x3)
// instead of the 96 KB of the single-CTA kernel for the same MMA time -- the single-CTA kernel needs
// ~62 B/clk/SM from L2, right at the SM's L2 ingest limit, which is what capped it (profiles/, round 1) --
// and the smaller stage allows 3 pipeline stages instead of 2.
//
//   warp 0      TMA producer (both CTAs): A rows m0 + rank*128, W rows n0 + rank*128; all transaction bytes
//               are signalled on the LEADER's "full" barrier (cta_group::2 form of cp.async.bulk.tensor)
//   warp 1      MMA issuer (leader CTA only): M=256,N=256,K=16; tcgen05.commit multicast to both CTAs
//   warp 2      TMEM allocator (cta_group::2 allocation, issued in both CTAs)
//   warps 4-11  epilogue (both CTAs, own 128 rows); "accumulator drained" arrives on the leader's barrier
//
// Tail tiles.  Tiles are dealt round-robin to the resident pairs; a last wave that covers only part of the pairs
// (FC2 at M = 32768: 256 tiles on 74 pairs = 3 waves + 34 tiles) leaves the rest idle for a whole tile time.  When
// the leftover tiles number at most half the pairs they are split into two 256x128 column halves each, one per
// pair (68 half tiles on 74 pairs), which ends the kernel ~0.4 tile times earlier.  A half tile issues the same
// K-steps in the same term order into the same accumulator rows, so the output bits do not change.
#include <stdlib.h>

#include "gemm_common.cuh"

namespace pips {

constexpr int P_BM = 256;                    // pair tile
constexpr int P_BN = 256;
constexpr int P_THREADS = 384;
constexpr uint32_t P_A_BYTES = 128 * BK * 2;         // this CTA's 128 rows of A
constexpr uint32_t P_W_BYTES = 128 * BK * 2;         // this CTA's half of the W tile

template <int TERMS>
struct PairCfg {
    static constexpr int kCopies = TERMS == 3 ? 2 : 1;
    static constexpr uint32_t kStageBytes = kCopies * (P_A_BYTES + P_W_BYTES);      // per CTA: 64 KB / 32 KB
    static constexpr int kStages = TERMS == 3 ? 3 : 6;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + EPI_STAGE_BYTES + 1024 + 256;
};

template <int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ CUtensorMap map_wn_hi, const __grid_constant__ CUtensorMap map_wn_lo,
                const GemmArgs args) {
    using Cfg = PairCfg<TERMS>;
    constexpr int kStages = Cfg::kStages;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* epi_stage = smem + kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPI_STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars);                 // used in the leader only
    const uint32_t empty0 = full0 + 8 * kStages;           // per CTA
    const uint32_t tfull0 = empty0 + 8 * kStages;          // per CTA
    const uint32_t tempty0 = tfull0 + 16;                  // used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    const int tiles_m = (args.M + P_BM - 1) / P_BM;
    const int tiles_n = (args.N + P_BN - 1) / P_BN;
    const int num_tiles = args.pair_full_tiles;             // 256x256 tiles dealt round-robin
    const bool has_narrow = pair < args.pair_narrow_tiles;  // then (at most) one 256x128 half tile for this pair
    const int narrow_tile = args.pair_full_tiles + (pair >> 1), narrow_half = pair & 1;
    const int num_kb = args.K / BK;
    (void)tiles_m;

    cluster_sync_all();                                     // both CTAs resident before the pair-wide TMEM allocation
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_w_hi);
        if (has_narrow) tma_prefetch_desc(&map_wn_hi);
        if (TERMS == 3) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_w_lo);
            if (has_narrow) tma_prefetch_desc(&map_wn_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 2);                    // one arrival per producer CTA (+ their bytes)
            mbar_init(empty0 + 8 * s, 1);                   // multicast commit from the leader's MMA thread
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 2 * (P_THREADS - 128));   // epilogue threads of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();                                     // barriers + TMEM of both CTAs ready before any remote use
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                             // operands and the residual stream come from the previous kernel

    // Register budget per warpgroup: the control warps (0-3) need few, the epilogue warpgroups (4-7, 8-11) many.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                const int m0 = (t / tiles_n) * P_BM + static_cast<int>(rank) * 128;
                const int n0 = (t % tiles_n) * P_BN + static_cast<int>(rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = leader_addr(full0 + 8 * stage);
                    if (leader) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * Cfg::kStageBytes);
                    else mbar_arrive_cluster(full0 + 8 * stage, 0);
                    const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
                    const int k0 = kb * BK;
                    tma_load_2d_pair(base, &map_a_hi, fb, k0, m0);
                    tma_load_2d_pair(base + P_A_BYTES, &map_w_hi, fb, k0, n0);
                    if (TERMS == 3) {
                        tma_load_2d_pair(base + P_A_BYTES + P_W_BYTES, &map_a_lo, fb, k0, m0);
                        tma_load_2d_pair(base + 2 * P_A_BYTES + P_W_BYTES, &map_w_lo, fb, k0, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (has_narrow) {                               // 256x128 half tile: this CTA's 128 rows of A, 64 rows of W
                constexpr uint32_t kNarrowBytes = Cfg::kCopies * (P_A_BYTES + P_W_BYTES / 2);
                const int m0 = (narrow_tile / tiles_n) * P_BM + static_cast<int>(rank) * 128;
                const int n0 = (narrow_tile % tiles_n) * P_BN + narrow_half * 128 + static_cast<int>(rank) * 64;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = leader_addr(full0 + 8 * stage);
                    if (leader) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * kNarrowBytes);
                    else mbar_arrive_cluster(full0 + 8 * stage, 0);
                    const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
                    const int k0 = kb * BK;
                    tma_load_2d_pair(base, &map_a_hi, fb, k0, m0);
                    tma_load_2d_pair(base + P_A_BYTES, &map_wn_hi, fb, k0, n0);
                    if (TERMS == 3) {
                        tma_load_2d_pair(base + P_A_BYTES + P_W_BYTES, &map_a_lo, fb, k0, m0);
                        tma_load_2d_pair(base + 2 * P_A_BYTES + P_W_BYTES, &map_wn_lo, fb, k0, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader only)
        // whole warp, warp-uniform control flow (descriptors stay in uniform registers); one elected lane issues
        if (leader) {
            const bool elected = elect_one();
            constexpr uint32_t idesc_wide = umma_idesc_bf16(P_BM, P_BN);
            constexpr uint32_t idesc_narrow = umma_idesc_bf16(P_BM, P_BN / 2);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int t = pair; t < num_tiles || (has_narrow && t < num_tiles + num_pairs); t += num_pairs, ++it) {
                const bool narrow = t >= num_tiles;          // the one item after this pair's 256x256 tiles
                const uint32_t idesc = narrow ? idesc_narrow : idesc_wide;
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * P_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t a_hi = umma_desc_sw128(base);
                    const uint64_t w_hi = umma_desc_sw128(base + P_A_BYTES);
                    const uint64_t a_lo = umma_desc_sw128(base + P_A_BYTES + P_W_BYTES);
                    const uint64_t w_lo = umma_desc_sw128(base + 2 * P_A_BYTES + P_W_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);
                        if (elected) umma_f16_pair(d_tmem, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                        if (TERMS == 3) {
                            if (elected) umma_f16_pair(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
                            if (elected) umma_f16_pair(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
                        }
                    }
                    if (elected) umma_commit_pair(empty0 + 8 * stage);
                    if (elected && kb == num_kb - 1) umma_commit_pair(tfull0 + 8 * as);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint8_t* my_stage = epi_stage + (warp - 4) * 4096;
        int it = 0;
        for (int t = pair; t < num_tiles || (has_narrow && t < num_tiles + num_pairs); t += num_pairs, ++it) {
            const bool narrow = t >= num_tiles;
            const int tile = narrow ? narrow_tile : t;
            const int hc = narrow ? P_BN / 4 : P_BN / 2;    // columns per epilogue half: 64 / 128
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int m0 = (tile / tiles_n) * P_BM + static_cast<int>(rank) * 128;
            const int n0 = (tile % tiles_n) * P_BN + (narrow ? narrow_half * 128 : 0) + half * hc;
            const int row0 = m0 + q * 32;
            mbar_wait(tfull0 + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * P_BN + half * hc + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t va[32], vb[32];
            float bias[32];
            tmem_ld_32x32(taddr, va);
            load_bias_chunk(args, n0, bias);
#pragma unroll 1
            for (int c0 = 0; c0 < hc; c0 += 64) {
                tmem_ld_wait();
                tmem_ld_32x32(taddr + c0 + 32, vb);
                epilogue_chunk(args, va, bias, row0, n0 + c0, n0 + c0 + 32, my_stage);
                __syncwarp();                                    // tcgen05.ld / wait are warp-collective
                tmem_ld_wait();
                if (c0 + 64 < hc) tmem_ld_32x32(taddr + c0 + 64, va);
                epilogue_chunk(args, vb, bias, row0, n0 + c0 + 32, c0 + 64 < hc ? n0 + c0 + 64 : -1, my_stage);
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive_cluster(tempty0 + 8 * as, 0);       // the leader's MMA thread owns the accumulator hand-off
        }
    }

    tc_fence_before();
    cluster_sync_all();                                     // nobody leaves while the peer may still signal us
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

static bool make_pair_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld_elems, int box_rows = 128) {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int TERMS>
static int launch_pair(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mw_hi, const CUtensorMap& mw_lo,
                       const CUtensorMap& mwn_hi, const CUtensorMap& mwn_lo, GemmArgs args, cudaStream_t st) {
    using Cfg = PairCfg<TERMS>;
    static bool attr[kMaxDevices] = {};
    {
        cudaError_t e = ensure_dyn_smem(gemm_tc2_kernel<TERMS>, attr, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail_cuda("pips_gemm_tc (pair): smem attribute", e);
    }
    const int tiles = ((args.M + P_BM - 1) / P_BM) * ((args.N + P_BN - 1) / P_BN);
    const int max_pairs = sm_count() / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    const int tail = tiles % pairs;                         // 256x256 tiles of a last, partial wave
    const char* e_tail = getenv("PIPS_B200_GEMM_TAIL");     // "0": keep the leftover tiles whole (A/B timing)
    if (tail > 0 && 2 * tail <= pairs && !(e_tail && e_tail[0] == '0')) {
        args.pair_full_tiles = tiles - tail;
        args.pair_narrow_tiles = 2 * tail;
    } else {
        args.pair_full_tiles = tiles;
        args.pair_narrow_tiles = 0;
    }
    cudaError_t e = launch_pdl(gemm_tc2_kernel<TERMS>, dim3(2 * pairs), dim3(P_THREADS), Cfg::kSmemBytes, st, ma_hi, ma_lo, mw_hi, mw_lo,
                               mwn_hi, mwn_lo, args);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_gemm_tc (pair): launch", e);
}

// called by pips_gemm_tc (gemm_tc.cu) after argument validation
int gemm_tc_pair_dispatch(const void* a_hi, const void* a_lo, int lda, int a_rows, const void* w_hi, const void* w_lo, int ldw,
                          int w_rows, const GemmArgs& args, cudaStream_t st) {
    const bool x3 = a_lo != nullptr;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo, mwn_hi, mwn_lo;          // mwn_*: 64-row boxes for the 256x128 tail tiles
    if (!make_pair_map(&ma_hi, a_hi, a_rows, args.K, lda)) return fail("pips_gemm_tc (pair): tensor map (A hi) failed");
    if (!make_pair_map(&mw_hi, w_hi, w_rows, args.K, ldw)) return fail("pips_gemm_tc (pair): tensor map (W hi) failed");
    if (!make_pair_map(&mwn_hi, w_hi, w_rows, args.K, ldw, 64)) return fail("pips_gemm_tc (pair): tensor map (W hi, 64 rows) failed");
    ma_lo = ma_hi;
    mw_lo = mw_hi;
    mwn_lo = mwn_hi;
    if (x3) {
        if (!make_pair_map(&ma_lo, a_lo, a_rows, args.K, lda)) return fail("pips_gemm_tc (pair): tensor map (A lo) failed");
        if (!make_pair_map(&mw_lo, w_lo, w_rows, args.K, ldw)) return fail("pips_gemm_tc (pair): tensor map (W lo) failed");
        if (!make_pair_map(&mwn_lo, w_lo, w_rows, args.K, ldw, 64)) return fail("pips_gemm_tc (pair): tensor map (W lo, 64 rows) failed");
    }
    return x3 ? launch_pair<3>(ma_hi, ma_lo, mw_hi, mw_lo, mwn_hi, mwn_lo, args, st)
              : launch_pair<1>(ma_hi, ma_lo, mw_hi, mw_lo, mwn_hi, mwn_lo, args, st);
}

}  // namespace pips
