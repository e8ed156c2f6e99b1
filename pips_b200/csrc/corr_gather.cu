// Fused local correlation + motion embedding + mixer-input packing (the "corr_gather" kernel).
//
// Replaces, per refinement iteration, CorrBlock.corr (all-pairs matmul, nets/pips.py:384-398),
// CorrBlock.sample (4x grid_sample, :355-382 + :313-328), the permute/reshape/cat glue (:517-522,
// :307-308) and get_3d_embedding (utils/misc.py:44-69).  The all-pairs volume is never formed: by
// linearity, bilinearly sampling the dot-product map at the 49 taps of a level equals blending the
// dot products with the 8x8 pixel footprint those taps share (they have one common fractional
// offset).  Zero padding outside the map (grid_sample's default) is exactly the TMA out-of-bounds fill.
//
// Work unit = one (b, s, n): for each of the 4 pyramid levels two TMA boxes {128 ch, 8 px, 4 rows}
// are staged into shared memory (channels-last pyramid => every pixel is one contiguous 512 B / 256 B
// line), each lane owns 4 channels, the 32 partial dot products of a box are reduced with a 31-shuffle
// butterfly that leaves pixel i's total in lane i, and the 7x7 blend, the sin/cos embedding and the
// feature copy are assembled in a per-warp row buffer that is written out as one contiguous mixer row.
// Every warp runs its own 3-deep TMA ring (it is both producer and consumer, so only "full" mbarriers
// are needed); 4 warps per CTA, persistent grid.
//
// Algorithmic bytes per unit: 4 levels x 64 px x 128 ch x e_f  +  128 x 4 (query)
// +  output row.
#include "common.cuh"
#include "ptx.cuh"

namespace pips {

constexpr int CG_WARPS = 4;
constexpr int CG_STAGES = 3;
constexpr int CG_ROWS_PER_BOX = 4;
constexpr int CG_ROWBUF = PIPS_KITCHEN_PAD;   // 576 floats

template <typename T>
struct CgCfg {
    static constexpr uint32_t kBoxBytes = CG_ROWS_PER_BOX * 8 * 128 * sizeof(T);
    static constexpr uint32_t kWarpBytes = CG_STAGES * kBoxBytes + CG_ROWBUF * 4 + 64 * 4 + 128 /*barriers; keeps 128 B alignment*/;
    static constexpr uint32_t kSmemBytes = CG_WARPS * kWarpBytes + 128;
};

struct CgArgs {
    int B, S, N;
    int H[PIPS_LEVELS], W[PIPS_LEVELS];
    const float* coords;     // (B,S,N,2)
    const float* ffeats;     // (B*N,S,128)
    const float* times;      // [S]
    const int* frame_base;   // optional [B*N]: window start of each track inside its clip of T frames (chained tracking)
    int T;                   // frames per batch element in the pyramid (== S when frame_base is null)
    __nv_bfloat16* x_hi;
    __nv_bfloat16* x_lo;
    float* x_f32;
    int ldx;
};

__device__ __forceinline__ float4 ld_chan4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_chan4(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    float4 r;
    r.x = __uint_as_float(u.x << 16);
    r.y = __uint_as_float(u.x & 0xffff0000u);
    r.z = __uint_as_float(u.y << 16);
    r.w = __uint_as_float(u.y & 0xffff0000u);
    return r;
}

struct UnitInfo {
    float cx, cy;      // coords of (b,s,n), level-0 pixels
    int frame;         // b*S + s
    int valid;
};

__device__ __forceinline__ UnitInfo load_unit(const CgArgs& a, long long u, long long total) {
    UnitInfo ui;
    ui.valid = u < total;
    ui.cx = 0.f; ui.cy = 0.f; ui.frame = 0;
    if (ui.valid) {
        const int s = static_cast<int>(u % a.S);
        const long long seq = u / a.S;
        const int b = static_cast<int>(seq / a.N), n = static_cast<int>(seq % a.N);
        const float2 c = *reinterpret_cast<const float2*>(a.coords + ((static_cast<size_t>(b) * a.S + s) * a.N + n) * 2);
        ui.cx = c.x; ui.cy = c.y;
        // chained windows replicate the clip's last frame past its end (chain_demo.py:50-52)
        ui.frame = a.frame_base ? b * a.T + min(a.frame_base[seq] + s, a.T - 1) : b * a.S + s;
    }
    return ui;
}

// TMA box origin of (level, half) for a unit; coordinates are clamped in float so that diverged or
// non-finite tracks cannot overflow the int conversion (anything beyond +-8 px of the map is all-zero anyway).
__device__ __forceinline__ void box_origin(const CgArgs& a, const UnitInfo& ui, int level, int half, int& x0, int& y0) {
    const float sc = 1.0f / static_cast<float>(1 << level);        // coords / 2**i  (nets/pips.py:373), exact
    const float cxl = fminf(fmaxf(ui.cx * sc, -8.0f), static_cast<float>(a.W[level]) + 8.0f);
    const float cyl = fminf(fmaxf(ui.cy * sc, -8.0f), static_cast<float>(a.H[level]) + 8.0f);
    x0 = static_cast<int>(floorf(cxl)) - PIPS_RADIUS;
    y0 = static_cast<int>(floorf(cyl)) - PIPS_RADIUS + half * CG_ROWS_PER_BOX;
}

template <typename T>
__global__ void __launch_bounds__(CG_WARPS * 32)
corr_gather_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                   const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const CgArgs a) {
    using Cfg = CgCfg<T>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wbase = smem + warp * Cfg::kWarpBytes;
    T* stage_buf = reinterpret_cast<T*>(wbase);
    float* rowbuf = reinterpret_cast<float*>(wbase + CG_STAGES * Cfg::kBoxBytes);
    float* dots = rowbuf + CG_ROWBUF;
    const uint32_t bar0 = smem_u32(dots + 64);
    const uint32_t stage0 = smem_u32(stage_buf);

    const long long total = static_cast<long long>(a.B) * a.S * a.N;
    const long long nwarps = static_cast<long long>(gridDim.x) * CG_WARPS;
    long long u = static_cast<long long>(blockIdx.x) * CG_WARPS + warp;
    if (u >= total) return;                         // whole warp exits together

    if (lane == 0) {
        for (int s = 0; s < CG_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_barrier_init();
    }
    for (int i = PIPS_KITCHEN + lane; i < CG_ROWBUF; i += 32) rowbuf[i] = 0.f;     // zero K padding, written once
    __syncwarp();

    const CUtensorMap* maps[PIPS_LEVELS] = {&map0, &map1, &map2, &map3};
    auto issue = [&](const UnitInfo& ui, int j, int slot) {
        // j in [0,8): level = j>>1, half = j&1
        if (lane == 0) {
            int x0, y0;
            box_origin(a, ui, j >> 1, j & 1, x0, y0);
            const uint32_t bar = bar0 + 8 * slot;
            mbar_arrive_expect_tx(bar, Cfg::kBoxBytes);
            tma_load_4d(stage0 + slot * Cfg::kBoxBytes, maps[j >> 1], bar, 0, x0, y0, ui.frame);
        }
    };

    UnitInfo cur = load_unit(a, u, total);
    UnitInfo nxt = load_unit(a, u + nwarps, total);
    float4 q = *reinterpret_cast<const float4*>(a.ffeats + static_cast<size_t>(u) * 128 + lane * 4);
    issue(cur, 0, 0);
    issue(cur, 1, 1);
    int slot = 0;
    uint32_t parity = 0;

    for (; u < total; u += nwarps) {
        // prefetch the next unit's query while this one is processed
        float4 nq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nxt.valid) nq = *reinterpret_cast<const float4*>(a.ffeats + static_cast<size_t>(u + nwarps) * 128 + lane * 4);
        const UnitInfo nxt2 = load_unit(a, u + 2 * nwarps, total);

#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            // keep the ring full: item j+2 of this unit, or item j-6 of the next one
            {
                int slot2 = slot + 2;
                if (slot2 >= CG_STAGES) slot2 -= CG_STAGES;
                if (j < 6) issue(cur, j + 2, slot2);
                else if (nxt.valid) issue(nxt, j - 6, slot2);
            }
            mbar_wait(bar0 + 8 * slot, parity);
            const T* sb = stage_buf + static_cast<size_t>(slot) * (Cfg::kBoxBytes / sizeof(T)) + lane * 4;
            float acc[32];
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 v = ld_chan4(sb + p * 128);
                acc[p] = (q.x * v.x + q.y * v.y) + (q.z * v.z + q.w * v.w);
            }
            // butterfly: after the 5 steps lane i holds the full dot product of pixel i of this box
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bool up = lane & 16;
                const float send = up ? acc[i] : acc[i + 16], keep = up ? acc[i + 16] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool up = lane & 8;
                const float send = up ? acc[i] : acc[i + 8], keep = up ? acc[i + 8] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool up = lane & 4;
                const float send = up ? acc[i] : acc[i + 4], keep = up ? acc[i + 4] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool up = lane & 2;
                const float send = up ? acc[i] : acc[i + 2], keep = up ? acc[i + 2] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            {
                const bool up = lane & 1;
                const float send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
                acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            // corrs / sqrt(C)  (nets/pips.py:397)
            dots[(j & 1) * 32 + lane] = __fdiv_rn(acc[0], 11.313708498984761f);
            __syncwarp();                           // stage slot fully consumed, dots visible
            if (j & 1) {
                const int level = j >> 1;
                const float sc = 1.0f / static_cast<float>(1 << level);
                const float cxl = fminf(fmaxf(cur.cx * sc, -8.0f), static_cast<float>(a.W[level]) + 8.0f);
                const float cyl = fminf(fmaxf(cur.cy * sc, -8.0f), static_cast<float>(a.H[level]) + 8.0f);
                const float fx = cxl - floorf(cxl), fy = cyl - floorf(cyl);
                const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
#pragma unroll
                for (int k = lane; k < 49; k += 32) {
                    const int ax = k / 7, by = k - ax * 7;        // channel k = a*7+b samples x = cx+(a-3), y = cy+(b-3)
                    const float* d = dots + by * 8 + ax;
                    rowbuf[PIPS_C + level * 49 + k] = ((w00 * d[0] + w01 * d[1]) + w10 * d[8]) + w11 * d[9];
                }
                __syncwarp();                       // dots are rewritten by the next level
            }
            if (++slot == CG_STAGES) { slot = 0; parity ^= 1; }
        }

        // feature copy + motion embedding (utils/misc.py:44-69; flows nets/pips.py:518-520)
        {
            const int s = static_cast<int>(u % a.S);
            const long long seq = u / a.S;
            const int b = static_cast<int>(seq / a.N), n = static_cast<int>(seq % a.N);
            const float2 c0 = *reinterpret_cast<const float2*>(a.coords + ((static_cast<size_t>(b) * a.S + 0) * a.N + n) * 2);
            const float flow_x = cur.cx - c0.x, flow_y = cur.cy - c0.y, t = a.times[s];
            *reinterpret_cast<float4*>(rowbuf + lane * 4) = q;
            const float div = static_cast<float>(lane) * 31.25f;          // arange(0,64,2) * (1000/64)
            float sn, cs;
            sincosf(__fmul_rn(flow_x, div), &sn, &cs);
            rowbuf[324 + 2 * lane] = sn; rowbuf[325 + 2 * lane] = cs;
            sincosf(__fmul_rn(flow_y, div), &sn, &cs);
            rowbuf[388 + 2 * lane] = sn; rowbuf[389 + 2 * lane] = cs;
            sincosf(__fmul_rn(t, div), &sn, &cs);
            rowbuf[452 + 2 * lane] = sn; rowbuf[453 + 2 * lane] = cs;
            if (lane == 0) { rowbuf[516] = flow_x; rowbuf[517] = flow_y; rowbuf[518] = t; }
        }
        __syncwarp();
        {
            const size_t ro = static_cast<size_t>(u) * a.ldx;
            for (int g = lane; g < CG_ROWBUF / 4; g += 32) {
                const float4 v = *reinterpret_cast<const float4*>(rowbuf + g * 4);
                if (a.x_f32) *reinterpret_cast<float4*>(a.x_f32 + ro + g * 4) = v;
                if (a.x_hi) {
                    __nv_bfloat16 h[4], l[4];
                    split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]);
                    split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
                    *reinterpret_cast<uint2*>(a.x_hi + ro + g * 4) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
                    if (a.x_lo) *reinterpret_cast<uint2*>(a.x_lo + ro + g * 4) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
                }
            }
        }
        __syncwarp();
        cur = nxt;
        nxt = nxt2;
        q = nq;
    }
}

template <typename T>
static int launch_corr_gather(const void* const* lvl, CUtensorMapDataType dt, const CgArgs& args, int frames, cudaStream_t st) {
    CUtensorMap maps[PIPS_LEVELS];
    for (int l = 0; l < PIPS_LEVELS; ++l) {
        const cuuint64_t es = sizeof(T);
        cuuint64_t gdim[4] = {128, static_cast<cuuint64_t>(args.W[l]), static_cast<cuuint64_t>(args.H[l]), static_cast<cuuint64_t>(frames)};
        cuuint64_t gstr[3] = {128 * es, static_cast<cuuint64_t>(args.W[l]) * 128 * es,
                              static_cast<cuuint64_t>(args.H[l]) * args.W[l] * 128 * es};
        cuuint32_t box[4] = {128, 8, CG_ROWS_PER_BOX, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (!encode_tiled(&maps[l], dt, 4, const_cast<void*>(lvl[l]), gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE))
            return fail("pips_corr_gather: cuTensorMapEncodeTiled failed");
    }
    using Cfg = CgCfg<T>;
    static bool attr[kMaxDevices] = {};
    {
        cudaError_t e = ensure_dyn_smem(corr_gather_kernel<T>, attr, static_cast<int>(Cfg::kSmemBytes));
        if (e != cudaSuccess) return fail_cuda("pips_corr_gather: smem attribute", e);
    }
    const long long units = static_cast<long long>(args.B) * args.S * args.N;
    const int ctas_per_sm = sizeof(T) == 2 ? 2 : 1;
    long long grid = (units + CG_WARPS - 1) / CG_WARPS;
    const long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;
    if (grid > cap) grid = cap;
    corr_gather_kernel<T><<<static_cast<unsigned>(grid), CG_WARPS * 32, Cfg::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], args);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_corr_gather: launch", e);
}

}  // namespace pips

using namespace pips;

extern "C" int pips_corr_gather(const void* const* lvl, int feat_dtype, int B, int S, int N, int H, int W, const float* coords,
                                const float* ffeats, const float* times, const int* frame_base, int frames_per_batch,
                                void* x_hi, void* x_lo, float* x_f32, int ldx, void* stream) {
    if (!lvl || !coords || !ffeats || !times) return fail("pips_corr_gather: null pointer");
    if (S != PIPS_S) return fail("pips_corr_gather: S must be 8");
    if (B <= 0 || N <= 0) return fail("pips_corr_gather: empty problem");
    if (H < 8 || W < 8) return fail("pips_corr_gather: level-0 map must be at least 8x8");
    if (!x_hi && !x_f32) return fail("pips_corr_gather: no output buffer");
    if (x_lo && !x_hi) return fail("pips_corr_gather: x_lo without x_hi");
    if (ldx < PIPS_KITCHEN_PAD || (ldx % 8)) return fail("pips_corr_gather: ldx must be >= 576 and a multiple of 8");
    CgArgs a;
    a.B = B; a.S = S; a.N = N;
    int h = H, w = W;
    for (int l = 0; l < PIPS_LEVELS; ++l) {
        if (!lvl[l]) return fail("pips_corr_gather: null level pointer");
        a.H[l] = h; a.W[l] = w; h /= 2; w /= 2;
    }
    if (frame_base && frames_per_batch <= 0) return fail("pips_corr_gather: frame_base needs frames_per_batch > 0");
    a.coords = coords; a.ffeats = ffeats; a.times = times;
    a.frame_base = frame_base; a.T = frame_base ? frames_per_batch : S;
    a.x_hi = static_cast<__nv_bfloat16*>(x_hi); a.x_lo = static_cast<__nv_bfloat16*>(x_lo); a.x_f32 = x_f32; a.ldx = ldx;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (feat_dtype == PIPS_FEAT_F32) return launch_corr_gather<float>(lvl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a, B * a.T, st);
    if (feat_dtype == PIPS_FEAT_BF16) return launch_corr_gather<__nv_bfloat16>(lvl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a, B * a.T, st);
    return fail("pips_corr_gather: unknown feat_dtype");
}
