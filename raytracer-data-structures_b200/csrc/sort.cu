// sort.cu — K3: onesweep least-significant-digit radix sort of (key, value) pairs.
//
// Replaces radixSort (accelerators.h:348-369: a decimal, linked-list radix sort of a by-value copy whose
// result is thrown away).  Contract: stable ascending sort of the Morton keys with the primitive index as
// payload.  One upfront histogram kernel reads the keys once and produces every digit pass's global bin
// counts; each digit pass is then ONE kernel ("onesweep"): tiles are claimed with an atomic ticket, ranked
// inside the tile with warp match/ballot, and chained to their predecessors with a decoupled look-back
// over packed {flag,count} words, so every key is read once and written once per pass (8 B + 8 B for
// 32-bit keys with a 32-bit payload).
#include "rtds_internal.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

constexpr uint32_t FLAG_AGG = 1u << 30;     // tile's own count is published
constexpr uint32_t FLAG_PREFIX = 2u << 30;  // inclusive prefix over tiles 0..t is published
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS) histogram_kernel(const K* __restrict__ keys, int n, int passes,
                                                                 uint32_t* __restrict__ ghist)
{
    extern __shared__ uint32_t sh[];  // [passes][RADIX]
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        K k = keys[i];
#pragma unroll 1
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * RADIX + (int)((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// exclusive scan of each pass's 256 bins: one block per pass
__global__ void __launch_bounds__(RADIX) scan_bins_kernel(uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t s[RADIX];
    uint32_t* h = ghist + blockIdx.x * RADIX;
    uint32_t v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < RADIX; off <<= 1) {
        uint32_t t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS)
onesweep_pass_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out, int n, int shift, const uint32_t* __restrict__ gbase /*[RADIX]*/,
                     uint32_t* __restrict__ ticket, volatile uint32_t* __restrict__ status /*[tiles][RADIX]*/)
{
    __shared__ uint32_t warp_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t digit_base[RADIX];
    __shared__ int s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const int tile = s_tile;
    const long long warp_base = (long long)tile * SORT_TILE + (long long)warp * (32 * SORT_ITEMS);

    K key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        long long idx = warp_base + j * 32 + lane;
        bool ok = idx < n;
        key[j] = ok ? keys_in[idx] : (K)~(K)0;
        val[j] = ok ? vals_in[idx] : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int d = (int)((key[j] >> shift) & (RADIX - 1));
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t base = warp_hist[warp][d];
        __syncwarp();
        rank[j] = base + __popc(peers & lt_mask);
        if (lane == 31 - __clz(peers)) warp_hist[warp][d] = base + __popc(peers);
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: prefix over warps, publish, look back
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
        const uint32_t count = run;
        uint32_t excl = 0;
        if (tile == 0) {
            status[(size_t)tile * RADIX + d] = FLAG_PREFIX | count;
        } else {
            status[(size_t)tile * RADIX + d] = FLAG_AGG | count;
            int t = tile - 1;
            while (true) {
                uint32_t s = status[(size_t)t * RADIX + d];
                uint32_t f = s & FLAG_MASK;
                if (f == 0) continue;  // predecessor not published yet (it holds an earlier ticket, so it runs)
                excl += s & VALUE_MASK;
                if (f == FLAG_PREFIX) break;
                --t;
            }
            status[(size_t)tile * RADIX + d] = FLAG_PREFIX | (excl + count);
        }
        digit_base[d] = gbase[d] + excl;
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        long long idx = warp_base + j * 32 + lane;
        if (idx < n) {
            int d = (int)((key[j] >> shift) & (RADIX - 1));
            uint32_t pos = digit_base[d] + warp_hist[warp][d] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

template <typename K>
int onesweep_sort(rtds_ctx* ctx, K* d_keys, uint32_t* d_vals, K* d_keys_tmp, uint32_t* d_vals_tmp, int n, int key_bits,
                  int* launches)
{
    if (n <= 1) return RTDS_OK;
    if (n >= (1 << 30)) { rtds_set_error("onesweep: n must be < 2^30"); return RTDS_ERR_INVALID; }
    const int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    const int tiles = (n + SORT_TILE - 1) / SORT_TILE;
    // workspace: ghist[passes][RADIX] | ticket[passes] (padded to RADIX) | status[passes][tiles][RADIX]
    size_t words = (size_t)passes * RADIX + RADIX + (size_t)passes * tiles * RADIX;
    size_t bytes = words * sizeof(uint32_t);
    if (ctx->sort_ws_bytes < bytes) {
        if (ctx->d_sort_ws) cudaFree(ctx->d_sort_ws);
        ctx->d_sort_ws = nullptr; ctx->sort_ws_bytes = 0;
        RTDS_CUDA(cudaMalloc(&ctx->d_sort_ws, bytes));
        ctx->sort_ws_bytes = bytes;
    }
    uint32_t* ghist = (uint32_t*)ctx->d_sort_ws;
    uint32_t* ticket = ghist + (size_t)passes * RADIX;
    uint32_t* status = ticket + RADIX;
    RTDS_TRY(rtds_zero_async(ctx->d_sort_ws, bytes, ctx->stream, launches));

    int hist_blocks = min(tiles, ctx->sm_count * 8);
    histogram_kernel<K><<<hist_blocks, SORT_THREADS, passes * RADIX * sizeof(uint32_t), ctx->stream>>>(d_keys, n, passes, ghist);
    scan_bins_kernel<<<passes, RADIX, 0, ctx->stream>>>(ghist);
    if (launches) *launches += 2;

    K* kin = d_keys; uint32_t* vin = d_vals; K* kout = d_keys_tmp; uint32_t* vout = d_vals_tmp;
    for (int p = 0; p < passes; ++p) {
        onesweep_pass_kernel<K><<<tiles, SORT_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, p * RADIX_BITS,
                                                                          ghist + (size_t)p * RADIX, ticket + p,
                                                                          status + (size_t)p * tiles * RADIX);
        if (launches) *launches += 1;
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    if (kin != d_keys) {  // odd number of passes: bring the result home
        RTDS_CUDA(cudaMemcpyAsync(d_keys, kin, sizeof(K) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        RTDS_CUDA(cudaMemcpyAsync(d_vals, vin, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    RTDS_CUDA(cudaGetLastError());
    return RTDS_OK;
}

}  // namespace

int rtds_onesweep_sort_u32(rtds_ctx* ctx, uint32_t* d_keys, uint32_t* d_vals, uint32_t* d_keys_tmp, uint32_t* d_vals_tmp,
                           int n, int key_bits, int* launches)
{
    return onesweep_sort<uint32_t>(ctx, d_keys, d_vals, d_keys_tmp, d_vals_tmp, n, key_bits, launches);
}

int rtds_onesweep_sort_u64(rtds_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_tmp, uint32_t* d_vals_tmp,
                           int n, int key_bits, int* launches)
{
    return onesweep_sort<uint64_t>(ctx, d_keys, d_vals, d_keys_tmp, d_vals_tmp, n, key_bits, launches);
}
