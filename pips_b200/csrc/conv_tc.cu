// 2-D convolution of the feature encoder as an implicit GEMM on CTA pairs (tcgen05.mma.cta_group::2).
//
//   out[n, oy, ox, co] = sum_{r,s,ci} in[n, oy*st + r - pad, ox*st + s - pad, ci] * w[co, r, s, ci]   (+ bias[co])
//
// for the 3x3 / 1x1, stride 1 / 2 convolutions of BasicEncoder's residual stages and head
// (nets/pips.py:135-136, :170, :221-223).  Activations are channels-last bf16 (hi, lo) pairs (v ~= hi + lo)
// written by pips_inorm_apply / pips_resize_pair, weights are packed (Cout_pad, taps*Cp) bf16 (hi, lo), and
// every product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM -- the same bf16x3
// scheme as the mixer GEMMs (gemm_tc2.cu), whose pipeline this kernel reuses:
//   GEMM-M  = output pixels, one CTA = TH rows x TW columns = 128 pixels (a pair = two adjacent tiles)
//   GEMM-N  = output channels (64 / 128 / 256; W rows split between the two CTAs of the pair)
//   GEMM-K  = taps x input channels; one K-chunk = 64 channels of one filter tap
// The A tile of a K-chunk is ONE 4-D TMA box {64 ch, TW px, TH rows, 1 image} of the input at the tap's
// offset (element strides = conv stride; out-of-bounds = zero padding), so no im2col buffer exists anywhere.
#include "gemm_common.cuh"

namespace pips {

constexpr int C_THREADS = 384;
constexpr uint32_t C_A_BYTES = 128 * BK * 2;          // 128 pixels x 64 channels bf16
constexpr int C_MAX_STAGES = 8;
constexpr uint32_t C_RING_BYTES = 192 * 1024;         // smem ring: stage = A_hi | A_lo | W_hi | W_lo, 2*(16 KB + BN/2*128 B)
constexpr uint32_t C_COMB_BYTES = 8 * 4 * 64 * 4;      // statistics exchange: 8 epilogue warps x up to 4 chunks x (32 sums + 32 squares)
constexpr uint32_t C_SMEM_BYTES = C_RING_BYTES + C_COMB_BYTES + 1024 + 256;

struct ConvArgs {
    int N, Ho, Wo, Cout;        // output (NHWC fp32, row stride Cout)
    int Cp;                     // padded input channels per tap (multiple of 64)
    int R, S, sy, sx, py, px;   // filter taps, conv stride (rows, cols), zero padding (rows, cols)
    int TW, TH;                 // pixel tile: TW * TH == 128
    int tiles_x, tiles_y;       // tiles per image
    int BN;                     // GEMM-N tile == padded Cout: 64, 128 or 256
    int stages;                 // ring depth: 3 (BN = 256) .. 4 (BN = 64)
    const float* bias;          // [Cout] or null
    float* out;
    // Work items: the tiles of ONE image in groups of G consecutive pair tiles (pt_img pair tiles per image, groups_img
    // groups).  A group is the unit of the InstanceNorm partial statistics (nets/pips.py:154-157): with `partial` set, the
    // epilogue accumulates (sum, sum of squares) per output channel over the group's valid pixels and every CTA writes one
    // row partial[(img * groups_img + g) * 2 + rank][2][Cout] -- a layout fixed by the image geometry alone (never by the
    // batch), reduced in fp64 by pips_inorm_finalize.
    int G, groups_img, pt_img;
    float* partial;             // optional
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const ConvArgs a) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    float* comb = reinterpret_cast<float*>(smem + C_RING_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_RING_BYTES + C_COMB_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * C_MAX_STAGES;
    const uint32_t tfull0 = empty0 + 8 * C_MAX_STAGES;
    const uint32_t tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C_MAX_STAGES + 4);
    const int C_STAGES = a.stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    const int tiles_img = a.tiles_x * a.tiles_y;
    const int num_items = a.N * a.groups_img;
    const int chunks = a.Cp / BK;
    const int num_kb = a.R * a.S * chunks;
    const uint32_t w_bytes = static_cast<uint32_t>(a.BN / 2) * BK * 2;       // this CTA's half of the weight tile
    const uint32_t stage_tx = 2 * (C_A_BYTES + w_bytes);                      // per CTA: hi + lo
    const uint32_t C_STAGE_BYTES = stage_tx;                                  // multiple of 1 KB for every BN
    const uint32_t off_w_hi = 2 * C_A_BYTES, off_w_lo = 2 * C_A_BYTES + w_bytes;

    cluster_sync_all();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 2);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 2 * (C_THREADS - 128));
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                         // the activation operand comes from the previous kernel

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 0) {
            // ------------------------------------------------------------ TMA producer (both CTAs)
            if (lane == 0) {
                uint32_t stage = 0, phase = 0;
                for (int item = pair; item < num_items; item += num_pairs) {
                  const int img = item / a.groups_img, p0 = (item - img * a.groups_img) * a.G, p1 = min(a.pt_img, p0 + a.G);
                  for (int ptl = p0; ptl < p1; ++ptl) {
                    const int rem = 2 * ptl + static_cast<int>(rank);       // tile inside the image; == tiles_img: a tile past the
                                                                            // last one (odd tile count), computed but never stored
                    const int oy0 = (rem / a.tiles_x) * a.TH, ox0 = (rem % a.tiles_x) * a.TW;
                    const int n0 = static_cast<int>(rank) * (a.BN / 2);
                    int kb = 0;
                    for (int r = 0; r < a.R; ++r) {
                        for (int s = 0; s < a.S; ++s) {
                            const int ix = ox0 * a.sx + s - a.px, iy = oy0 * a.sy + r - a.py;
                            for (int c = 0; c < chunks; ++c, ++kb) {
                                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                                const uint32_t fb = leader_addr(full0 + 8 * stage);
                                if (leader) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * stage_tx);
                                else mbar_arrive_cluster(full0 + 8 * stage, 0);
                                const uint32_t base = smem_u32(smem + stage * C_STAGE_BYTES);
                                tma_load_4d_pair(base, &map_a_hi, fb, c * BK, ix, iy, img);
                                tma_load_4d_pair(base + C_A_BYTES, &map_a_lo, fb, c * BK, ix, iy, img);
                                tma_load_2d_pair(base + off_w_hi, &map_w_hi, fb, kb * BK, n0);
                                tma_load_2d_pair(base + off_w_lo, &map_w_lo, fb, kb * BK, n0);
                                if (++stage == C_STAGES) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                  }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ------------------------------------------------------------ MMA issuer (leader only)
            // whole warp, warp-uniform control flow (descriptors stay in uniform registers); one elected lane issues
            if (leader) {
                const bool elected = elect_one();
                const uint32_t idesc = umma_idesc_bf16(256, a.BN);
                uint32_t stage = 0, phase = 0;
                int it = 0;
                for (int item = pair; item < num_items; item += num_pairs) {
                  const int g0 = (item % a.groups_img) * a.G, ntl = min(a.pt_img, g0 + a.G) - g0;
                  for (int jt = 0; jt < ntl; ++jt, ++it) {
                    const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                    mbar_wait(tempty0 + 8 * as, aphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * 256;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t base = smem_u32(smem + stage * C_STAGE_BYTES);
                        const uint64_t a_hi = umma_desc_sw128(base);
                        const uint64_t a_lo = umma_desc_sw128(base + C_A_BYTES);
                        const uint64_t w_hi = umma_desc_sw128(base + off_w_hi);
                        const uint64_t w_lo = umma_desc_sw128(base + off_w_lo);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);
                            if (elected) umma_f16_pair(d_tmem, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                            if (elected) umma_f16_pair(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
                            if (elected) umma_f16_pair(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
                        }
                        if (elected) umma_commit_pair(empty0 + 8 * stage);
                        if (elected && kb == num_kb - 1) umma_commit_pair(tfull0 + 8 * as);
                        if (++stage == C_STAGES) { stage = 0; phase ^= 1; }
                    }
                  }
                }
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 pixels)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int half_cols = a.BN / 2;                       // 32, 64 or 128 columns per epilogue half
        int it = 0;
        for (int item = pair; item < num_items; item += num_pairs) {
            const int img = item / a.groups_img, grp = item - img * a.groups_img;
            const int p0 = grp * a.G, p1 = min(a.pt_img, p0 + a.G);
            float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};     // lane j: channel chunk*32 + j of this warp's half
            for (int ptl = p0; ptl < p1; ++ptl, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                const int rem = 2 * ptl + static_cast<int>(rank);
                const int i = q * 32 + lane;                      // pixel index inside the tile
                const int oy = (rem / a.tiles_x) * a.TH + i / a.TW, ox = (rem % a.tiles_x) * a.TW + i % a.TW;
                const bool ok = rem < tiles_img && oy < a.Ho && ox < a.Wo;
                float* orow = a.out + ((static_cast<size_t>(img) * a.Ho + oy) * a.Wo + ox) * a.Cout;
                mbar_wait(tfull0 + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + as * 256 + half * half_cols + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int c0 = ci * 32;
                    if (c0 < half_cols) {                         // warp-uniform
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c0, v);
                        tmem_ld_wait();
                        const int col = half * half_cols + c0;
                        if (ok && col < a.Cout) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col + j < a.Cout) {           // Cout is a multiple of 4
                                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                           __uint_as_float(v[j + 3]));
                                    if (a.bias) {
                                        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col + j));
                                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                                        v[j] = __float_as_uint(o.x); v[j + 1] = __float_as_uint(o.y);
                                        v[j + 2] = __float_as_uint(o.z); v[j + 3] = __float_as_uint(o.w);
                                    }
                                    *reinterpret_cast<float4*>(orow + col + j) = o;
                                }
                            }
                        }
                        if (a.partial) {
                            // transpose-reduce over the warp's 32 pixels (pixels outside the image contribute 0):
                            // lane j ends with this tile's totals of channel col + j
                            float s[32], sq[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float f = ok ? __uint_as_float(v[j]) : 0.f;
                                s[j] = f; sq[j] = f * f;
                            }
#pragma unroll
                            for (int w = 16; w >= 1; w >>= 1) {
                                const bool up = lane & w;
#pragma unroll
                                for (int j = 0; j < w; ++j) {
                                    const float send_s = up ? s[j] : s[j + w], keep_s = up ? s[j + w] : s[j];
                                    const float send_q = up ? sq[j] : sq[j + w], keep_q = up ? sq[j + w] : sq[j];
                                    s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
                                    sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
                                }
                            }
                            ssum[ci] += s[0];
                            ssq[ci] += sq[0];
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                mbar_arrive_cluster(tempty0 + 8 * as, 0);
            }
            if (a.partial) {
                float* mine = comb + (warp - 4) * 256;            // [chunk][sum 32 | squares 32]
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) { mine[ci * 64 + lane] = ssum[ci]; mine[ci * 64 + 32 + lane] = ssq[ci]; }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (q == 0) {                                      // warps 4 and 8: the four pixel quadrants in a fixed order
                    const float* c4 = comb + (half * 4) * 256;
                    float* p = a.partial + ((static_cast<size_t>(img) * a.groups_img + grp) * 2 + rank) * 2 * a.Cout;
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int col = half * half_cols + ci * 32 + lane;
                        if (ci * 32 < half_cols && col < a.Cout) {
                            const float* e = c4 + ci * 64;
                            p[col] = ((e[lane] + e[256 + lane]) + e[512 + lane]) + e[768 + lane];
                            p[a.Cout + col] = ((e[32 + lane] + e[256 + 32 + lane]) + e[512 + 32 + lane]) + e[768 + 32 + lane];
                        }
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");    // comb is rewritten by the next item
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace pips

using namespace pips;

// x_hi/x_lo: (N, H, W, Cp) bf16; w_hi/w_lo: (BN, R*S*Cp) bf16 with BN = Cout rounded up to 64/128/256 (zero rows);
// out: (N, Ho, Wo, Cout) fp32.
// tile shape and work-item grouping from the output geometry alone (the partial-statistics layout depends on it)
static void conv_geometry(int Ho, int Wo, int sx, ConvArgs& a) {
    // widest power-of-two tile row that does not exceed the output width (<= 128), the rest in rows
    int tw = 128;
    while (tw > 8 && tw > Wo) tw >>= 1;
    if (sx == 2 && (tw - 1) * 2 + 1 > 256) tw = 64;
    a.TW = tw; a.TH = 128 / tw;
    a.tiles_x = (Wo + a.TW - 1) / a.TW; a.tiles_y = (Ho + a.TH - 1) / a.TH;
    a.pt_img = (a.tiles_x * a.tiles_y + 1) / 2;               // pair tiles per image
    a.G = (a.pt_img + 31) / 32;                               // at most 32 groups (64 partial rows) per image
    a.groups_img = (a.pt_img + a.G - 1) / a.G;
}

static int conv_tc_impl(const void* x_hi, const void* x_lo, int N, int H, int W, int Cp, const void* w_hi, const void* w_lo,
                        int Cout, int R, int S, int sy, int sx, int py, int px, const float* bias, float* out, float* partial,
                        void* stream) {
    if (!x_hi || !x_lo || !w_hi || !w_lo || !out) return fail("pips_conv_tc: null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || Cp <= 0 || (Cp % BK)) return fail("pips_conv_tc: Cp must be a positive multiple of 64");
    if (Cout <= 0 || Cout > 256 || (Cout % 4)) return fail("pips_conv_tc: Cout must be a multiple of 4, at most 256");
    if (R <= 0 || S <= 0 || R > 7 || S > 7 || (sy != 1 && sy != 2) || (sx != 1 && sx != 2) || py < 0 || px < 0)
        return fail("pips_conv_tc: unsupported filter geometry");
    const int Ho = (H + 2 * py - R) / sy + 1, Wo = (W + 2 * px - S) / sx + 1;
    if (Ho <= 0 || Wo <= 0) return fail("pips_conv_tc: empty output");
    ConvArgs a;
    a.N = N; a.Ho = Ho; a.Wo = Wo; a.Cout = Cout; a.Cp = Cp; a.R = R; a.S = S; a.sy = sy; a.sx = sx; a.py = py; a.px = px;
    a.BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    {
        const uint32_t stage = 2 * (C_A_BYTES + static_cast<uint32_t>(a.BN / 2) * BK * 2);
        int st = static_cast<int>(C_RING_BYTES / stage);
        a.stages = st > C_MAX_STAGES ? C_MAX_STAGES : st;      // BN 256: 3 x 64 KB, 128: 4 x 48 KB, 64: 4 x 40 KB
    }
    conv_geometry(Ho, Wo, sx, a);
    a.bias = bias; a.out = out; a.partial = partial;

    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    {
        cuuint64_t gdim[4] = {static_cast<cuuint64_t>(Cp), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
        cuuint64_t gstr[3] = {static_cast<cuuint64_t>(Cp) * 2, static_cast<cuuint64_t>(W) * Cp * 2, static_cast<cuuint64_t>(H) * W * Cp * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>((a.TW - 1) * sx + 1),
                             static_cast<cuuint32_t>((a.TH - 1) * sy + 1), 1};
        cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(sx), static_cast<cuuint32_t>(sy), 1};
        if (!encode_tiled(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_hi), gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !encode_tiled(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_lo), gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_128B))
            return fail("pips_conv_tc: activation tensor map failed");
    }
    {
        const int K = R * S * Cp;
        cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(a.BN)};
        cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(a.BN / 2)};
        cuuint32_t estr[2] = {1, 1};
        if (!encode_tiled(&mw_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_hi), gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !encode_tiled(&mw_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_lo), gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_128B))
            return fail("pips_conv_tc: weight tensor map failed");
    }
    static bool attr[kMaxDevices] = {};
    {
        cudaError_t e = ensure_dyn_smem(conv_tc_kernel, attr, C_SMEM_BYTES);
        if (e != cudaSuccess) return fail_cuda("pips_conv_tc: smem attribute", e);
    }
    const int items = N * a.groups_img;
    const int max_pairs = sm_count() / 2;
    const int pairs = items < max_pairs ? items : max_pairs;
    cudaError_t e = launch_pdl(conv_tc_kernel, dim3(2 * pairs), dim3(C_THREADS), C_SMEM_BYTES, static_cast<cudaStream_t>(stream), ma_hi, ma_lo,
                               mw_hi, mw_lo, a);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_conv_tc: launch", e);
}

extern "C" int pips_conv_tc(const void* x_hi, const void* x_lo, int N, int H, int W, int Cp, const void* w_hi, const void* w_lo,
                            int Cout, int R, int S, int stride, int pad, const float* bias, float* out, void* stream) {
    return conv_tc_impl(x_hi, x_lo, N, H, W, Cp, w_hi, w_lo, Cout, R, S, stride, stride, pad, pad, bias, out, nullptr, stream);
}

// Anisotropic form (separate row / column stride and padding): the 7x7 stem runs as a 7x1 convolution over the
// column-unfolded image written by pips_stem_pack.
extern "C" int pips_conv_tc_aniso(const void* x_hi, const void* x_lo, int N, int H, int W, int Cp, const void* w_hi, const void* w_lo,
                                  int Cout, int R, int S, int stride_y, int stride_x, int pad_y, int pad_x, const float* bias,
                                  float* out, void* stream) {
    return conv_tc_impl(x_hi, x_lo, N, H, W, Cp, w_hi, w_lo, Cout, R, S, stride_y, stride_x, pad_y, pad_x, bias, out, nullptr, stream);
}

// Number of partial-statistics rows per image pips_conv_tc_stats writes for this geometry (0: invalid geometry).
extern "C" int pips_conv_tc_chunks(int H, int W, int R, int S, int stride_y, int stride_x, int pad_y, int pad_x) {
    if (H <= 0 || W <= 0 || R <= 0 || S <= 0 || stride_y <= 0 || stride_x <= 0) return 0;
    const int Ho = (H + 2 * pad_y - R) / stride_y + 1, Wo = (W + 2 * pad_x - S) / stride_x + 1;
    if (Ho <= 0 || Wo <= 0) return 0;
    ConvArgs a;
    conv_geometry(Ho, Wo, stride_x, a);
    return a.groups_img * 2;
}

// pips_conv_tc_aniso that also accumulates the InstanceNorm partial statistics of its output in the epilogue:
// partial (N, pips_conv_tc_chunks(...), 2, Cout) fp32 = per chunk (sum, sum of squares) per channel, for pips_inorm_finalize.
extern "C" int pips_conv_tc_stats(const void* x_hi, const void* x_lo, int N, int H, int W, int Cp, const void* w_hi, const void* w_lo,
                                  int Cout, int R, int S, int stride_y, int stride_x, int pad_y, int pad_x, const float* bias,
                                  float* out, float* partial, void* stream) {
    if (!partial) return fail("pips_conv_tc_stats: null partial buffer");
    return conv_tc_impl(x_hi, x_lo, N, H, W, Cp, w_hi, w_lo, Cout, R, S, stride_y, stride_x, pad_y, pad_x, bias, out, partial, stream);
}
