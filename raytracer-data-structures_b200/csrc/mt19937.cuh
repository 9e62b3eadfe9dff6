// mt19937.cuh — K11: the reference's jitter stream on the device, fused with primary-ray generation.
//   random_double()   main.cpp:503-508    std::mt19937 (seed 5489) + uniform_real_distribution<double>:
//                                         generate_canonical<double,53> = (lo + hi*2^32) / 2^64
//   render()          main.cpp:554-557    xx / yy from the two jitter doubles (double arithmetic, narrowed to float)
//   Vec3::normalize() geometry.h:125-134  factor = (float)(1.0 / sqrt((double)n))
// Included by render.cu only (inside its anonymous namespace's translation unit).
#pragma once
#include "rtds_internal.cuh"

namespace {

// ===================================================================================================
// K11: MT19937
// ===================================================================================================
constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_SNAP_EVERY = 8;  // regenerations between stored state snapshots
constexpr int MT_THREADS = 256;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v)
{
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// One regeneration A -> B (624 words) by a block of >= 227 threads, ONE phase and one barrier: x[k+624] =
// x[k+397] ^ twist(x[k], x[k+1]) makes B[t+227] depend on B[t] and B[t+454] on B[t+227], so thread t carries its own
// chain of three words in registers; B[623] needs B[396] (thread 169's second word) and B[0], which that thread recomputes.
// (Three 227-wide phases with a barrier each cost the jitter kernel 24 barriers per block and left a third of the warps idle.)
__device__ __forceinline__ void mt_regen(const uint32_t* __restrict__ A, uint32_t* __restrict__ B)
{
    const int t = threadIdx.x;
    if (t < 227) {
        const uint32_t b0 = A[t + MT_M] ^ mt_twist(A[t], A[t + 1]);
        B[t] = b0;
        const uint32_t b1 = b0 ^ mt_twist(A[t + 227], A[t + 228]);
        B[t + 227] = b1;
        if (t < 169) B[t + 454] = b1 ^ mt_twist(A[t + 454], A[t + 455]);
        else if (t == 169) B[623] = b1 ^ mt_twist(A[623], A[MT_M] ^ mt_twist(A[0], A[1]));
    }
    __syncthreads();
}

// Sequential walk of the generator by ONE block: stores the state after every MT_SNAP_EVERY regenerations.
// snap[k] = state after k*MT_SNAP_EVERY regenerations; k in [k0, k1). snap[0] is the seeded state.
__global__ void __launch_bounds__(MT_THREADS) mt_snapshot_kernel(uint32_t* __restrict__ snap, int k0, int k1, uint32_t seed)
{
    __shared__ uint32_t S[2][MT_N];
    const int t = threadIdx.x;
    int cur = 0;
    if (k0 == 0) {
        if (t == 0) {
            uint32_t x = seed;
            S[0][0] = x;
            for (int i = 1; i < MT_N; ++i) { x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i; S[0][i] = x; }
        }
        __syncthreads();
        for (int i = t; i < MT_N; i += MT_THREADS) snap[i] = S[0][i];
        k0 = 1;
    } else {
        for (int i = t; i < MT_N; i += MT_THREADS) S[0][i] = snap[(size_t)(k0 - 1) * MT_N + i];
        __syncthreads();
    }
    for (int k = k0; k < k1; ++k) {
        for (int r = 0; r < MT_SNAP_EVERY; ++r) { mt_regen(S[cur], S[cur ^ 1]); cur ^= 1; }
        for (int i = t; i < MT_N; i += MT_THREADS) snap[(size_t)k * MT_N + i] = S[cur][i];
    }
}

// Which part of the stream a rank needs: the samples of the scanline tiles it owns (all of it when world == 1).
struct JitterOwner {
    unsigned long long first_word;   // stream word of sample 0 of pixel 0
    unsigned long long row_words;    // 4 * width * spp
    int tile_rows, rank, world, height;
};

// Block b regenerates MT_SNAP_EVERY times from snapshot (s0 + b) and writes the tempered words.
// out[0] is stream word (s0 * MT_SNAP_EVERY * 624). Chunks that hold no sample of an owned row are skipped.
__global__ void __launch_bounds__(MT_THREADS) mt_expand_kernel(const uint32_t* __restrict__ snap, int s0,
                                                               uint32_t* __restrict__ out, size_t n_words, const JitterOwner own)
{
    __shared__ uint32_t S[2][MT_N];
    const int t = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * MT_SNAP_EVERY * MT_N;
    if (own.world > 1) {
        const unsigned long long wlo = (unsigned long long)(s0 + blockIdx.x) * MT_SNAP_EVERY * MT_N;
        const unsigned long long whi = wlo + MT_SNAP_EVERY * MT_N - 1;
        long long ylo = wlo > own.first_word ? (long long)((wlo - own.first_word) / own.row_words) : 0;
        long long yhi = whi > own.first_word ? (long long)((whi - own.first_word) / own.row_words) : 0;
        if (yhi >= own.height) yhi = own.height - 1;
        bool mine = false;
        for (long long tl = ylo / own.tile_rows; tl <= yhi / own.tile_rows; ++tl) mine |= (tl % own.world) == own.rank;
        if (!mine) return;
    }
    const uint32_t* src = snap + (size_t)(s0 + blockIdx.x) * MT_N;
    for (int i = t; i < MT_N; i += MT_THREADS) S[0][i] = src[i];
    __syncthreads();
    int cur = 0;
    for (int r = 0; r < MT_SNAP_EVERY; ++r) {
        mt_regen(S[cur], S[cur ^ 1]);
        cur ^= 1;
        for (int i = t; i < MT_N; i += MT_THREADS) {
            size_t w = base + (size_t)r * MT_N + i;
            if (w < n_words) out[w] = mt_temper(S[cur][i]);
        }
    }
}

// exact uint32 -> double without the conversion unit: 2^52 + u is representable, the subtraction is exact
__device__ __forceinline__ double u32_to_double(uint32_t u) { return __hiloint2double(0x43300000, (int)u) - 4503599627370496.0; }

// generate_canonical<double,53>(mt19937): two draws, (lo + hi*2^32)/2^64, clamped below 1.
__device__ __forceinline__ double canonical53(uint32_t lo, uint32_t hi)
{
    // hi * 2^32 exactly: 2^84 + hi * 2^32 is representable (ulp 2^32), the subtraction is exact; the sum rounds once
    // (to nearest even), exactly like the reference's long-double sum narrowed to double
    const double hi32 = __hiloint2double(0x45300000, (int)hi) - 19342813113834066795298816.0;
    const double sum = u32_to_double(lo) + hi32;
    double r = sum * 5.42101086242752217e-20;    // / 2^64, exact
    if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;  // nextafter(1,0)
    return r;
}

__global__ void jitter_doubles_kernel(const uint32_t* __restrict__ words, size_t first_double_rel, int n, double* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        size_t w = (first_double_rel + (size_t)i) * 2;
        out[i] = canonical53(words[w], words[w + 1]);
    }
}

// geometry.h:125-134: n = x*x+y*y+z*z (float); factor = (float)(1 / sqrt((double)n))
__device__ __forceinline__ void normalize3(float& x, float& y, float& z)
{
    float n = x * x + y * y + z * z;
    if (n > 0) {
        float factor = (float)(1.0 / sqrt((double)n));
        x *= factor; y *= factor; z *= factor;
    }
}

// Primary ray direction of sample (px, py) from its four jitter words: main.cpp:554-557 (double -> float exactly as there).
struct RayGen { float angle, aspect, inv_w, inv_h; };
__device__ __forceinline__ void primary_dir(const uint4 jw, int px, int py, const RayGen& G, float& dx, float& dy, float& dz)
{
    const double r1 = canonical53(jw.x, jw.y), r2 = canonical53(jw.z, jw.w);
    dx = (float)((2 * ((u32_to_double((unsigned)px) + r1) * (double)G.inv_w) - 1) * (double)G.angle * (double)G.aspect);
    dy = (float)((1 - 2 * ((u32_to_double((unsigned)py) + r2) * (double)G.inv_h)) * (double)G.angle);
    dz = -1;
    normalize3(dx, dy, dz);
}

// exact unsigned division by a launch-time constant (Granlund-Montgomery round-up form): magic == 0 -> power of two
struct FastDiv { uint32_t magic, shift; };
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv d)
{
    if (d.magic == 0) return n >> d.shift;
    const uint32_t q = __umulhi(n, d.magic);
    return (((n - q) >> 1) + q) >> d.shift;
}

// K11 + ray generation fused: block b regenerates MT_SNAP_EVERY times from snapshot (s0 + b), each regeneration
// written straight into the next 624-word slice of one shared array (the state after regeneration r IS slice r), then
// all threads turn the chunk's 1,248 samples (4 words each) into primary directions — the double-precision part of
// main.cpp:554-557 and Vec3::normalize — and store 3 floats per sample: 12 bytes instead of the 16 bytes of raw
// words, and the render kernel starts from ready directions. dirs[3 * g] is frame sample g = pixel * spp + k.
__global__ void __launch_bounds__(MT_THREADS) mt_expand_dirs_kernel(const uint32_t* __restrict__ snap, int s0, float* __restrict__ dirs,
                                                                    unsigned long long first_sample, unsigned n_samples, int width,
                                                                    int spp, const FastDiv div_spp, const FastDiv div_width,
                                                                    const RayGen G, const JitterOwner own, int chunks_per_tile)
{
    __shared__ __align__(16) uint32_t words[(MT_SNAP_EVERY + 1) * MT_N];      // slice 0 = the snapshot
    const int t = threadIdx.x;
    {
        const unsigned vb = blockIdx.x;
        // world == 1: block b = chunk s0 + b. world > 1: the grid covers only the rank's OWN scanline tiles (chunks_per_tile
        // blocks per owned tile). A chunk that straddles a tile boundary is regenerated by a block of each tile it touches, and every
        // block stores only the samples of ITS tile [lo, hi) - no direction is written twice, also when a tile is shorter than a chunk.
        unsigned long long chunk = (unsigned long long)s0 + vb;
        long long lo = 0, hi = (long long)n_samples;      // frame samples this block may write
        if (own.world > 1) {
            const unsigned k = vb / (unsigned)chunks_per_tile, c = vb - k * (unsigned)chunks_per_tile;
            const unsigned long long tile = (unsigned long long)own.rank + (unsigned long long)k * own.world;
            const unsigned long long row0 = tile * own.tile_rows;
            if (row0 >= (unsigned long long)own.height) return;
            const unsigned long long row1 = min(row0 + own.tile_rows, (unsigned long long)own.height);
            const unsigned long long w0 = own.first_word + row0 * own.row_words, w1 = own.first_word + row1 * own.row_words - 1;
            chunk = w0 / (MT_SNAP_EVERY * MT_N) + c;
            if (chunk > w1 / (MT_SNAP_EVERY * MT_N)) return;
            lo = (long long)(row0 * (own.row_words / 4));
            hi = (long long)(row1 * (own.row_words / 4));
        }
        const unsigned long long wlo = chunk * MT_SNAP_EVERY * MT_N;
        const uint32_t* src = snap + (size_t)chunk * MT_N;
        for (int i = t; i < MT_N; i += MT_THREADS) words[i] = src[i];
        __syncthreads();
        for (int r = 0; r < MT_SNAP_EVERY; ++r) mt_regen(words + r * MT_N, words + (r + 1) * MT_N);
        // all threads, no barriers: temper, canonical doubles, direction, normalise, store
        const unsigned long long as0 = wlo / 4;           // absolute stream sample of the chunk's first four words
        const uint4* w4 = reinterpret_cast<const uint4*>(words + MT_N);
        constexpr int CHUNK_SAMPLES = MT_SNAP_EVERY * MT_N / 4;
        // frame sample of the chunk's first stream sample; chunks that lie wholly inside the frame (all but the first and the
        // last) need no 64-bit range check per sample, and when aa_samples divides the thread stride the pixel coordinates
        // advance by additions instead of two divisions per sample
        const long long rel = (long long)as0 - (long long)first_sample;
        const bool inside = rel >= lo && rel + CHUNK_SAMPLES <= hi;
        const bool stepping = inside && (MT_THREADS % spp) == 0;
        const unsigned step_px = stepping ? (unsigned)(MT_THREADS / spp) : 0u;
        unsigned px = 0, py = 0;
        if (stepping) {
            const unsigned pix = fast_div((unsigned)rel + (unsigned)t, div_spp);
            py = fast_div(pix, div_width); px = pix - py * (unsigned)width;
        }
        for (int i = t; i < CHUNK_SAMPLES; i += MT_THREADS) {
            const long long gl = rel + i;
            if (inside || (gl >= lo && gl < hi)) {
                const unsigned g = (unsigned)gl;
                if (!stepping) {
                    const unsigned pix = fast_div(g, div_spp);
                    py = fast_div(pix, div_width); px = pix - py * (unsigned)width;
                }
                const uint4 w = w4[i];
                const uint4 jw = make_uint4(mt_temper(w.x), mt_temper(w.y), mt_temper(w.z), mt_temper(w.w));
                float dx, dy, dz;
                primary_dir(jw, (int)px, (int)py, G, dx, dy, dz);
                float* o = dirs + 3 * (size_t)g;
                o[0] = dx; o[1] = dy; o[2] = dz;
                if (stepping) { px += step_px; while (px >= (unsigned)width) { px -= (unsigned)width; ++py; } }
            }
        }
    }
}

}  // namespace
