// scan.cuh — exclusive scan of 32-bit values on the device (three kernels: tile sums, scan of the tile sums by one
// block, apply). Used by the level-synchronous builders (kd.cu, sah.cu) for their segmented compactions.
#pragma once
#include "rtds_internal.cuh"

namespace rtds_scan {

constexpr int SC_BLOCK = 256, SC_ITEMS = 8, SC_TILE = SC_BLOCK * SC_ITEMS;

static __global__ void __launch_bounds__(SC_BLOCK) scan_tile_sums(const int* __restrict__ in, int n, int* __restrict__ sums)
{
    __shared__ int ws[SC_BLOCK / 32];
    long long base = (long long)blockIdx.x * SC_TILE;
    int v = 0;
    for (int i = threadIdx.x; i < SC_TILE; i += SC_BLOCK) { long long p = base + i; v += p < n ? in[p] : 0; }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < SC_BLOCK / 32; ++w) t += ws[w]; sums[blockIdx.x] = t; }
}
static __global__ void __launch_bounds__(1024) scan_sums(int* sums, int tiles, int* total)
{
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < tiles ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < tiles) sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
static __global__ void __launch_bounds__(SC_BLOCK) scan_apply(const int* __restrict__ in, int n, const int* __restrict__ sums, int* __restrict__ out)
{
    __shared__ int ws[SC_BLOCK / 32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = sums[blockIdx.x];
    __syncthreads();
    long long base = (long long)blockIdx.x * SC_TILE;
    for (int it = 0; it < SC_ITEMS; ++it) {
        long long p = base + it * SC_BLOCK + threadIdx.x;
        int v = p < n ? in[p] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        int woff = 0, chunk = 0;
        for (int w = 0; w < SC_BLOCK / 32; ++w) { int c = ws[w]; woff += (w < warp) ? c : 0; chunk += c; }
        const int start = running;
        if (p < n) out[p] = start + woff + x - v;
        __syncthreads();
        if (threadIdx.x == 0) running = start + chunk;
        __syncthreads();
    }
}

struct Scanner {
    rtds_ctx* ctx; int* sums; int cap_tiles; int* d_total;
    int run(const int* in, int* out, int n, int* launches)
    {
        if (n <= 0) { cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream); return RTDS_OK; }
        int tiles = (n + SC_TILE - 1) / SC_TILE;
        if (tiles > cap_tiles) { rtds_set_error("scan: workspace too small"); return RTDS_ERR_CAPACITY; }
        scan_tile_sums<<<tiles, SC_BLOCK, 0, ctx->stream>>>(in, n, sums);
        scan_sums<<<1, 1024, 0, ctx->stream>>>(sums, tiles, d_total);
        scan_apply<<<tiles, SC_BLOCK, 0, ctx->stream>>>(in, n, sums, out);
        if (launches) *launches += 3;
        return RTDS_OK;
    }
};

}  // namespace rtds_scan
