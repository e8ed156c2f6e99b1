// Sparse score map: the dense heat-map of nets/pips.py:504-511 for a handful of
// chosen particles, straight from the resident pyramid -- the (B,S,N,H_l,W_l) all-pairs volumes that the
// reference materialises for every particle (CorrBlock.corr, nets/pips.py:384-398) never exist.
//
//   fcp[b,s,j] = sum_l  interpolate( <ffeats[b,s,sel[j]], fmaps_l[b,s,:,y,x]> / sqrt(C), (H8,W8),
//                                    mode='bilinear', align_corners=True )
//
// Two launches: (1) heat_corr -- one warp per pyramid pixel dots its 128 channels with every selected query
// of that frame (queries staged in shared memory); (2) heat_upsample -- one thread per output pixel gathers
// the 4 taps of each level and adds the levels in the reference's order (0, 1, 2, 3).
#include "common.cuh"

namespace pips {
namespace {

constexpr int C = PIPS_C;
constexpr int L = PIPS_LEVELS;
constexpr int MAX_SEL = 32;          // queries per launch of heat_corr (16 KB of shared memory)

struct HeatLevels {
    const void* lvl[L];
    int H[L], W[L];
    int off[L + 1];                  // pixel offset of each level inside one (frame, query) scratch row
};

template <bool BF16>
__global__ void __launch_bounds__(256) heat_corr_kernel(HeatLevels lv, const float* __restrict__ ffeats,
                                                        const int* __restrict__ sel, int n_sel, int j0, int nj, int S,
                                                        int N, float* __restrict__ scratch) {
    __shared__ float4 q[MAX_SEL][C / 4];
    const int frame = blockIdx.y;                        // b * S + s
    const int b = frame / S, s = frame - b * S;
    for (int i = threadIdx.x; i < nj * (C / 4); i += blockDim.x) {
        const int j = i / (C / 4), c4 = i - j * (C / 4);
        const int n = sel[j0 + j];
        q[j][c4] = reinterpret_cast<const float4*>(ffeats + (static_cast<size_t>(b * N + n) * S + s) * C)[c4];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (blockDim.x >> 5) * gridDim.x;
    const int P = lv.off[L];
    const float scale = 0.08838834764831845f;            // 1 / sqrt(128), nets/pips.py:396
    for (int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < P; p += warps) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < L; ++k) l += (p >= lv.off[k]);
        const size_t pix = static_cast<size_t>(frame) * lv.H[l] * lv.W[l] + (p - lv.off[l]);
        float4 f;
        if (BF16) {
            const uint2 raw = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(lv.lvl[l]) + pix * C)[lane];
            f.x = __uint_as_float(raw.x << 16);
            f.y = __uint_as_float(raw.x & 0xffff0000u);
            f.z = __uint_as_float(raw.y << 16);
            f.w = __uint_as_float(raw.y & 0xffff0000u);
        } else {
            f = reinterpret_cast<const float4*>(static_cast<const float*>(lv.lvl[l]) + pix * C)[lane];
        }
        for (int j = 0; j < nj; ++j) {
            const float4 w = q[j][lane];
            float acc = fmaf(f.x, w.x, fmaf(f.y, w.y, fmaf(f.z, w.z, f.w * w.w)));
            acc = warp_sum(acc);
            if (lane == 0) scratch[(static_cast<size_t>(frame) * n_sel + j0 + j) * P + p] = acc * scale;
        }
    }
}

// ATen's upsample_bilinear2d with align_corners=True: src = dst * (in-1)/(out-1), i0 = (int)src, lambda = src - i0,
// second tap i0 + (i0 < in-1).
__global__ void __launch_bounds__(256) heat_upsample_kernel(HeatLevels lv, const float* __restrict__ scratch,
                                                            const int* __restrict__ slot, int n_sel, int H8, int W8,
                                                            size_t out_frame_stride, float* __restrict__ out, size_t total) {
    const int P = lv.off[L];
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W8);
        size_t r = i / W8;
        const int y = static_cast<int>(r % H8);
        r /= H8;
        const int j = static_cast<int>(r % n_sel);
        const size_t frame = r / n_sel;
        const float* row = scratch + (frame * n_sel + j) * P;
        float acc = 0.0f;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int h = lv.H[l], w = lv.W[l];
            const float sy = H8 > 1 ? static_cast<float>(h - 1) / static_cast<float>(H8 - 1) : 0.0f;
            const float sx = W8 > 1 ? static_cast<float>(w - 1) / static_cast<float>(W8 - 1) : 0.0f;
            const float fy = sy * y, fx = sx * x;
            const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
            const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
            const float ly = fy - y0, lx = fx - x0;
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            const float* m = row + lv.off[l];
            const float v00 = m[y0 * w + x0], v01 = m[y0 * w + x0 + xp];
            const float v10 = m[(y0 + yp) * w + x0], v11 = m[(y0 + yp) * w + x0 + xp];
            acc += hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
        }
        out[frame * out_frame_stride + (static_cast<size_t>(slot ? slot[j] : j) * H8 + y) * W8 + x] = acc;
    }
}

}  // namespace
}  // namespace pips

extern "C" size_t pips_heatmap_scratch_floats(int frames, int n_sel, int H8, int W8) {
    size_t P = 0;
    int h = H8, w = W8;
    for (int l = 0; l < PIPS_LEVELS; ++l) {
        P += static_cast<size_t>(h) * w;
        h /= 2;
        w /= 2;
    }
    return static_cast<size_t>(frames) * n_sel * P;
}

extern "C" int pips_heatmap(const void* const* lvl, int feat_dtype, int B, int S, int N, int H8, int W8, const float* ffeats,
                            const int* sel, const int* slot, int n_sel, float* scratch, float* out, size_t out_frame_stride,
                            void* stream) {
    using namespace pips;
    if (!lvl || !ffeats || !sel || !scratch || !out) return fail("pips_heatmap: null pointer");
    if (B <= 0 || S <= 0 || N <= 0 || n_sel <= 0) return fail("pips_heatmap: empty problem");
    if ((H8 >> (PIPS_LEVELS - 1)) < 1 || (W8 >> (PIPS_LEVELS - 1)) < 1) return fail("pips_heatmap: feature map too small for 4 levels");
    if (feat_dtype != PIPS_FEAT_F32 && feat_dtype != PIPS_FEAT_BF16) return fail("pips_heatmap: bad feat_dtype");
    if (!slot && out_frame_stride < static_cast<size_t>(n_sel) * H8 * W8) return fail("pips_heatmap: out_frame_stride too small");
    HeatLevels lv;
    int h = H8, w = W8, off = 0;
    for (int l = 0; l < PIPS_LEVELS; ++l) {
        if (!lvl[l]) return fail("pips_heatmap: null pyramid level");
        lv.lvl[l] = lvl[l];
        lv.H[l] = h;
        lv.W[l] = w;
        lv.off[l] = off;
        off += h * w;
        h /= 2;
        w /= 2;
    }
    lv.off[PIPS_LEVELS] = off;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int frames = B * S;
    const int bx = (off + 8 * 16 - 1) / (8 * 16);           // 8 warps per block, ~16 pixels per warp
    for (int j0 = 0; j0 < n_sel; j0 += MAX_SEL) {
        const int nj = n_sel - j0 < MAX_SEL ? n_sel - j0 : MAX_SEL;
        dim3 grid(bx, frames);
        if (feat_dtype == PIPS_FEAT_BF16)
            heat_corr_kernel<true><<<grid, 256, 0, st>>>(lv, ffeats, sel, n_sel, j0, nj, S, N, scratch);
        else
            heat_corr_kernel<false><<<grid, 256, 0, st>>>(lv, ffeats, sel, n_sel, j0, nj, S, N, scratch);
    }
    const size_t total = static_cast<size_t>(frames) * n_sel * H8 * W8;
    const size_t blocks = (total + 255) / 256;
    const int cap = sm_count() * 16;
    heat_upsample_kernel<<<static_cast<unsigned>(blocks < static_cast<size_t>(cap) ? blocks : cap), 256, 0, st>>>(
        lv, scratch, slot, n_sel, H8, W8, out_frame_stride, out, total);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_heatmap", e);
}
