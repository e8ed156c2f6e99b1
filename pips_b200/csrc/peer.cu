// Node-local exchange of results without a collective library.
//
// Particle-sharded ranks (one process per GPU) each own a "slab" of device memory that every other rank maps
// through CUDA IPC.  Producers store their slice of a result directly into all slabs over NVLink (the coordinate
// predictions from inside the update kernel, mixer_simt.cu; visibility logits and features with peer_scatter
// below), and a flag barrier replaces the collective's implicit synchronisation.  Nothing here allocates or
// synchronises on the hot path: alloc / open / close / free are set-up calls.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pips {
namespace {

__global__ void __launch_bounds__(256) peer_scatter_kernel(const float* __restrict__ src, int rows, int cols,
                                                           pips_peer_out dst, int cols_total, int col_offset) {
    const size_t total = static_cast<size_t>(rows) * cols;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / cols, c = i - r * cols;
        const float v = src[i];
        const size_t o = r * cols_total + col_offset + c;
        for (int p = 0; p < dst.n_peers; ++p) dst.out[p][o] = v;
    }
}

// 16-byte form (cols, cols_total, col_offset multiples of 4, pointers 16-byte aligned): a warp stores 512 contiguous
// bytes per peer and instruction -- NVLink carries writes in packets of up to 256 B of payload behind one header, so
// whole lines matter far more here than for local stores.
__global__ void __launch_bounds__(256) peer_scatter4_kernel(const float4* __restrict__ src, int rows, int cols4,
                                                            pips_peer_out dst, int cols_total4, int col_offset4) {
    const size_t total = static_cast<size_t>(rows) * cols4;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / cols4, c = i - r * cols4;
        const float4 v = src[i];
        const size_t o = r * cols_total4 + col_offset4 + c;
        for (int p = 0; p < dst.n_peers; ++p) reinterpret_cast<float4*>(dst.out[p])[o] = v;
    }
}

struct FlagPtrs {
    int* f[PIPS_MAX_PEERS];
};

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// one thread per rank: publish my arrival in rank p's flag array, then wait for rank p's arrival in mine
__global__ void __launch_bounds__(32) peer_barrier_kernel(FlagPtrs flags, int rank, int n_peers, int epoch,
                                                          unsigned long long timeout_ns) {
    const int p = threadIdx.x;
    if (p >= n_peers) return;
    __threadfence_system();                                   // everything this stream wrote before is visible first
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flags.f[p] + rank), "r"(epoch) : "memory");
    const int* mine = flags.f[rank] + p;
    const unsigned long long t0 = now_ns();
    int v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (now_ns() - t0 > timeout_ns) {
            printf("pips_peer_barrier: rank %d waited for rank %d (epoch %d, saw %d)\n", rank, p, epoch, v);
            __trap();
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace
}  // namespace pips

extern "C" int pips_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    using namespace pips;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ptr || !handle64 || bytes == 0) return fail("pips_peer_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail_cuda("pips_peer_alloc: cudaMalloc", e);
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail_cuda("pips_peer_alloc: cudaIpcGetMemHandle", e);
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return 0;
}

extern "C" int pips_peer_open(const unsigned char* handle64, void** ptr) {
    using namespace pips;
    if (!ptr || !handle64) return fail("pips_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail_cuda("pips_peer_open: cudaIpcOpenMemHandle", e);
    *ptr = p;
    return 0;
}

extern "C" int pips_peer_close(void* ptr) {
    using namespace pips;
    if (!ptr) return fail("pips_peer_close: null pointer");
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? 0 : fail_cuda("pips_peer_close", e);
}

extern "C" int pips_peer_free(void* ptr) {
    using namespace pips;
    if (!ptr) return fail("pips_peer_free: null pointer");
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? 0 : fail_cuda("pips_peer_free", e);
}

extern "C" int pips_peer_scatter(const float* src, int rows, int cols, float* const* dst, int n_peers, int cols_total,
                                 int col_offset, void* stream) {
    using namespace pips;
    if (!src || !dst) return fail("pips_peer_scatter: null pointer");
    if (rows <= 0 || cols <= 0) return fail("pips_peer_scatter: empty block");
    if (n_peers <= 0 || n_peers > PIPS_MAX_PEERS) return fail("pips_peer_scatter: n_peers must be 1..PIPS_MAX_PEERS");
    if (col_offset < 0 || col_offset + cols > cols_total) return fail("pips_peer_scatter: block outside the destination row");
    pips_peer_out d;
    d.n_peers = n_peers;
    d.n_offset = 0;
    d.n_total = 0;
    for (int r = 0; r < n_peers; ++r) {
        if (!dst[r]) return fail("pips_peer_scatter: null destination");
        d.out[r] = dst[r];
    }
    bool vec = (cols % 4 == 0) && (cols_total % 4 == 0) && (col_offset % 4 == 0) && !(reinterpret_cast<uintptr_t>(src) & 15);
    for (int r = 0; r < n_peers; ++r) vec = vec && !(reinterpret_cast<uintptr_t>(dst[r]) & 15);
    const size_t total = static_cast<size_t>(rows) * cols / (vec ? 4 : 1);
    const size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(sm_count()) * 8;
    const unsigned grid = static_cast<unsigned>(blocks < cap ? blocks : cap);
    if (vec)
        peer_scatter4_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), rows, cols / 4, d,
                                                                                 cols_total / 4, col_offset / 4);
    else
        peer_scatter_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, rows, cols, d, cols_total, col_offset);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_peer_scatter", e);
}

extern "C" int pips_peer_barrier(int* const* flags, int rank, int n_peers, int epoch, int timeout_ms, void* stream) {
    using namespace pips;
    if (!flags) return fail("pips_peer_barrier: null pointer");
    if (n_peers <= 0 || n_peers > PIPS_MAX_PEERS || rank < 0 || rank >= n_peers) return fail("pips_peer_barrier: bad rank / n_peers");
    if (timeout_ms <= 0) return fail("pips_peer_barrier: timeout must be positive");
    FlagPtrs f;
    for (int r = 0; r < n_peers; ++r) {
        if (!flags[r]) return fail("pips_peer_barrier: null flag array");
        f.f[r] = flags[r];
    }
    peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, rank, n_peers, epoch,
                                                                          static_cast<unsigned long long>(timeout_ms) * 1000000ull);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pips_peer_barrier", e);
}
