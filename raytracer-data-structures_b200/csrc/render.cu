// render.cu — K10: the render kernels (shading + accumulate + 8-bit quantise around the traversals of traverse.cuh,
// primary directions from mt19937.cuh), the parity-probe trace kernel, the multi-GPU frame flags, and their host side.
//
// Reference semantics reproduced here (all citations relative to /root/reference/project/raytracer/):
//   render()            main.cpp:541-566   per-pixel sample loop, accumulation order
//   castRay()           main.cpp:291-500   Phong shading, material branches, depth limit, shadow query
//   write_into_file()   main.cpp:516-528   (unsigned char)(min(1, c/aa) * 255)
// The translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true so that every float
// operation rounds exactly like the reference's SSE2 code.
#include "rtds_internal.cuh"
#include <math.h>
#include <algorithm>
#include <functional>
#include <stdlib.h>
#include <string.h>

#include "mt19937.cuh"
#include "traverse.cuh"

namespace {

// ===================================================================================================
// shading: castRay's DIFFUSE_AND_GLOSSY branch (main.cpp:394-497)
// ===================================================================================================
struct ShadeParams {
    int       n_lights;
    RtdsLight lights[RTDS_MAX_LIGHTS];
    float     bg[3];
    float     bias;
    int       max_depth;
    int       shadows;
};

// powf(x, 25) for x in [0, ~1]: exact-product chain in double, one rounding to float. glibc's powf is
// correctly rounded except in astronomically rare cases, and so is this.
__device__ __forceinline__ float pow25f(float xf)
{
    double x = (double)xf;
    double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
    return (float)(x16 * x8 * x);
}

__device__ __forceinline__ void shade_diffuse(const ShadeParams& P, float dx, float dy, float dz, float hx, float hy, float hz,
                                              float nx, float ny, float nz, float sr, float sg, float sb, float& r, float& g, float& b)
{
    // (hx,hy,hz) = rayorig + raydir * tnear; (nx,ny,nz) = hitPoint - centre (un-normalised): normalise, flip towards the ray
    normalize3(nx, ny, nz);
    if (dx * nx + dy * ny + dz * nz > 0) { nx = -nx; ny = -ny; nz = -nz; }
    float hr = 0, hg = 0, hb = 0;
    for (int i = 0; i < P.n_lights; ++i) {
        const RtdsLight& L = P.lights[i];
        float lx = L.c[0] - hx, ly = L.c[1] - hy, lz = L.c[2] - hz;
        normalize3(lx, ly, lz);
        float LdotN = fmaxf(0.f, lx * nx + ly * ny + lz * nz);
        // lightAmt = (1 - inShadow) * emissionColor * LdotN      (inShadow = 0: trace_more is a stub)
        float ar = (L.le[0] * 1.0f) * LdotN, ag = (L.le[1] * 1.0f) * LdotN, ab = (L.le[2] * 1.0f) * LdotN;
        // reflect(-lightDir, N) = I - 2*dot(I,N)*N
        float ix = -lx, iy = -ly, iz = -lz;
        float s2 = 2 * (ix * nx + iy * ny + iz * nz);
        float rx = ix - nx * s2, ry = iy - ny * s2, rz = iz - nz * s2;
        float sp = pow25f(fmaxf(0.f, -(rx * dx + ry * dy + rz * dz)));
        float spr = L.le[0] * sp, spg = L.le[1] * sp, spb = L.le[2] * sp;
        // hitColor += lightAmt * (diffuse * 0.8) / 2 + specular * 0.5; diffuse = (0.815,0.235,0.031)
        hr += (ar * (0.815f * 0.8f)) / 2.0f + spr * 0.5f;
        hg += (ag * (0.235f * 0.8f)) / 2.0f + spg * 0.5f;
        hb += (ab * (0.031f * 0.8f)) / 2.0f + spb * 0.5f;
        hr += sr; hg += sg; hb += sb;
    }
    r = hr; g = hg; b = hb;
}

// ===================================================================================================
// K10: the render kernel. One thread per pixel of this rank's rows, samples looped in order so the
// per-pixel float accumulation matches main.cpp:553-560.
// ===================================================================================================
struct RenderArgs {
    int   width, height, spp;
    int   rank, world, tile_rows, local_rows;
    int   lrow0;                 // first local row of this launch (the frame may be rendered in row bands)
    float angle, aspect, inv_w, inv_h;
    const uint32_t* jitter;      // word 0 = stream word jitter_base (strip kernel / tests only)
    const float*    dirs;        // primary directions, 3 floats per frame sample (pixel * spp + k), from mt_expand_dirs_kernel
    const uint32_t* mt_snap;     // MT19937 state snapshots (strip kernel regenerates its words itself)
    unsigned long long first_sample;   // stream index of sample 0 of pixel 0 (jitter_offset)
    int             chunk_first; // absolute snapshot chunk of block 0 of the strip kernel
    uint64_t        jitter_rel;  // (4*first_sample - jitter_base): word offset of sample 0 of pixel 0
    BvhView bvh;
    KdView  kd;
    const float4* sph;           // objId-indexed {c, r}
    const float4* tri;           // objId-indexed v0,v1,v2 (triangle scenes)
    int           prim_type;
    const float4* mat;           // objId-indexed {rgb, material}
    int           n;             // primitive count for NONE
    ShadeParams   shade;
    uint8_t* out_rgb;            // local rows x width x 3, or the whole frame (height x width x 3) when out_global_rows
    int      out_global_rows;    // 1: out_rgb is a full frame indexed by the GLOBAL row (possibly another GPU's memory)
    int      block_order;        // 0 quadrant-major, 1 row-major, 2 quadrant-major from the centre rows outwards (default)
    int      out_vec8;           // out_rgb is 8-byte aligned and width % 8 == 0: warps store whole 8-byte words
    int*     out_hit;            // optional, local rows x width
    float*   out_accum;          // optional, local rows x width x 3
    unsigned long long* counters;
    const int*     heavy_list;   // lpt_split: tiles rendered by render_heavy_kernel this frame (-1: unused entry), or nullptr
    const unsigned char* skip;   // ... and the per-tile flag that makes the packet kernel leave them alone
    float*         wave_t;       // wavefront form (shadows on): per local sample (lrow * width + px) * spp + k, the primary hit's tnear ...
    int*           wave_leaf;    // ... and leaf (-1: sky), written by wave_primary_kernel, read by wave_shade_kernel
    const int* order;            // optional: launch position -> block id (the previous frame's heaviest blocks first), see block_order_kernel
    unsigned*  cost;             // optional: per block id, the largest per-thread node-visit count of this frame (input of the next frame's order)
};

// Heaviest blocks first. A handful of 16 x 8-pixel blocks (silhouettes, rays that graze many leaf boxes) run 15-25 times the median
// block - measured with tools/block_timeline.py: median 7.5 us, p99 105 us, max 187 us on the bench frame - and the kernel cannot end
// before "start of the slowest block + its duration". In launch order those blocks start late (the lower image half 55 us into
// a 240 us per-rank kernel at 8 GPUs; 800 us into the 950 us single-GPU kernel). So every frame records each block's cost (the
// largest node-visit count among its threads), and the NEXT frame of the same geometry launches the few hundred heaviest blocks
// first, everything else in the usual quadrant-major order (which keeps one direction octant's loop copy per SM: +20 % per block
// when that is given up). Pure scheduling: pixels do not depend on the order blocks run in.
__device__ __forceinline__ int logical_block(const RenderArgs& A) { return A.order ? __ldg(A.order + blockIdx.x) : (int)blockIdx.x; }
__device__ __forceinline__ void report_block_cost(const RenderArgs& A, int lb, unsigned cost)      // all 32 lanes
{
    if (A.cost) {
        const unsigned m = __reduce_max_sync(0xffffffffu, cost);
        if ((threadIdx.x & 31) == 0 && m) atomicMax(A.cost + lb, m);
    }
}

// cost[0..n) of the finished frame -> order[0..n) for the next one: the heaviest blocks first, heaviest of all at the front, the
// rest behind them in ascending id (the usual launch order). "Heaviest" = cost in the top 32nds of [0, max] from the top down
// for as long as at most `cap` blocks (half a wave of resident blocks) qualify, and never below max/4: enough to start every
// straggler in the first wave, few enough to leave the octant-coherent launch order of everything else alone (with 2 % of a
// 64,800-block frame moved to the front the frame got 2 % SLOWER: an SM full of heavy blocks of all four octants). Clears cost.
// One block of 1024 threads per band.
constexpr int LPT_MAX_HEAVY = 1024;
constexpr int LPT_SPLIT_MAX = 256;       // most tiles handed to render_heavy_kernel per frame
constexpr int LPT_TRIAL_FRAMES = 6;      // lpt = 1: timed frames of a geometry spent comparing the schedules (2 or 3 modes, round robin)
struct LptBands { int n_bands; int off[RTDS_MAX_BANDS + 1]; };      // a frame rendered as row bands: one launch (and one order) per band
__global__ void __launch_bounds__(1024) block_order_kernel(unsigned* __restrict__ cost, int* __restrict__ order, const LptBands bands, int cap,
                                                           int* __restrict__ heavy_list, unsigned char* __restrict__ skip, int n_split, int min_bin)
{
    cost += bands.off[blockIdx.x]; order += bands.off[blockIdx.x]; skip += bands.off[blockIdx.x];
    const int n = bands.off[blockIdx.x + 1] - bands.off[blockIdx.x];
    __shared__ unsigned s_u[32];
    __shared__ int s_i[32], s_j[32];
    __shared__ unsigned s_hist[32];
    __shared__ unsigned s_M;
    __shared__ int s_tbin, s_heavy;
    __shared__ unsigned s_hc[LPT_MAX_HEAVY];
    __shared__ int s_hid[LPT_MAX_HEAVY];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    cap = min(cap, LPT_MAX_HEAVY);
    unsigned mx = 0;
    for (int i = t; i < n; i += 1024) mx = max(mx, cost[i]);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) s_u[warp] = mx;
    if (t < 32) s_hist[t] = 0u;
    __syncthreads();
    if (warp == 0) { unsigned v = s_u[lane]; v = __reduce_max_sync(0xffffffffu, v); if (lane == 0) s_M = v; }
    __syncthreads();
    const unsigned M = s_M;
    auto bin_of = [M](unsigned c) { return M ? (int)min(31ull, (unsigned long long)c * 32ull / M) : 0; };
    for (int i = t; i < n; i += 1024) { const unsigned c = cost[i]; if (c) atomicAdd(&s_hist[bin_of(c)], 1u); }
    __syncthreads();
    if (t == 0) {
        int tb = 32, total = 0;
        for (int b2 = 31; b2 >= min_bin && M; --b2) {    // default min_bin 8: cost >= max / 4
            if (total + (int)s_hist[b2] > cap) break;
            total += (int)s_hist[b2];
            tb = b2;
        }
        s_tbin = tb;
    }
    __syncthreads();
    const int tbin = s_tbin;
    // stable split: thread t owns ids [t*per, (t+1)*per)
    const int per = (n + 1023) / 1024, i0 = min(n, t * per), i1 = min(n, i0 + per);
    int mine = 0;
    for (int i = i0; i < i1; ++i) { const unsigned c = cost[i]; mine += (c && bin_of(c) >= tbin); }
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_i[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_i[lane], sc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += u; }
        s_j[lane] = sc - v;                       // exclusive offset of each warp
        if (lane == 31) s_heavy = sc;
    }
    __syncthreads();
    int before = s_j[warp] + incl - mine;         // heavy ids below i0
    const int n_heavy = s_heavy;                  // <= cap <= LPT_MAX_HEAVY
    for (int i = i0; i < i1; ++i) {
        const unsigned c = cost[i];
        const bool h = c && bin_of(c) >= tbin;
        if (h) { s_hc[before] = c; s_hid[before] = i; } else order[n_heavy + i - before] = i;
        before += h;
        cost[i] = 0u;
        skip[i] = 0;
    }
    // the heavy ones by descending cost (ties: ascending id): bitonic sort of the padded list in shared memory
    for (int i = n_heavy + t; i < LPT_MAX_HEAVY; i += 1024) { s_hc[i] = 0u; s_hid[i] = 0x7fffffff; }
    __syncthreads();
    for (int k = 2; k <= LPT_MAX_HEAVY; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (t < LPT_MAX_HEAVY) {
                const int x = t ^ j;
                if (x > t) {
                    const unsigned ca = s_hc[t], cb = s_hc[x];
                    const int ia = s_hid[t], ib = s_hid[x];
                    const bool a_first = ca > cb || (ca == cb && ia < ib);       // a belongs before b in the final order
                    const bool up = (t & k) == 0;
                    if (up ? !a_first : a_first) { s_hc[t] = cb; s_hc[x] = ca; s_hid[t] = ib; s_hid[x] = ia; }
                }
            }
            __syncthreads();
        }
    if (t < n_heavy) order[t] = s_hid[t];
    // lpt_split (single-band frames): the heaviest of them go to render_heavy_kernel next frame
    if (heavy_list && blockIdx.x == 0) {
        const int ns = min(n_split, n_heavy);
        if (t < LPT_SPLIT_MAX) heavy_list[t] = t < ns ? s_hid[t] : -1;
        if (t < ns) skip[s_hid[t]] = 1;
    }
}

// Frame-buffer write of a warp's 8 x 4 pixels. The warp's RGB8 values are staged in shared memory and leave as 12
// aligned 8-byte stores (4 rows x 24 bytes) instead of 96 single-byte stores: when the frame lives in ANOTHER GPU's
// memory (out_global_rows: every rank stores its tiles straight into rank 0's frame over NVLink, there is no gather
// afterwards) the transfer is made of full 8-byte writes. Warp-level only (__syncwarp): a block barrier here made
// finished warps wait for the slowest one (10 % of the kernel's stall samples). Falls back to byte stores for ragged
// widths / edge warps / unaligned buffers. All 32 lanes of the warp must call.
__device__ __forceinline__ int global_row_of(const RenderArgs& A, int lrow)
{
    const int tile = lrow / A.tile_rows, within = lrow - tile * A.tile_rows;
    return (tile * A.world + A.rank) * A.tile_rows + within;
}
__device__ __forceinline__ void store_warp_rgb(const RenderArgs& A, int px, int lrow, bool active, unsigned char r8, unsigned char g8,
                                               unsigned char b8)
{
    __shared__ __align__(16) unsigned char tile[4][4][24];       // [warp][row][8 pixels x 3]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px0 = px - (lane & 7), row0 = lrow - (lane >> 3);   // the warp's top-left pixel
    const bool fast = A.out_vec8 && px0 + 8 <= A.width && row0 + 4 <= A.local_rows;     // warp-uniform
    if (fast) {
        unsigned char* t = &tile[warp][lane >> 3][3 * (lane & 7)];
        t[0] = r8; t[1] = g8; t[2] = b8;
        __syncwarp();
        if (lane < 12) {
            const int row = lane / 3, seg = lane - 3 * row;
            const int orow = A.out_global_rows ? global_row_of(A, row0 + row) : row0 + row;
            const uint2 v = *reinterpret_cast<const uint2*>(&tile[warp][row][8 * seg]);
            *reinterpret_cast<uint2*>(A.out_rgb + ((size_t)orow * A.width + px0) * 3 + 8 * seg) = v;
        }
    } else if (active) {
        const size_t o = (size_t)(A.out_global_rows ? global_row_of(A, lrow) : lrow) * A.width + px;
        A.out_rgb[3 * o] = r8; A.out_rgb[3 * o + 1] = g8; A.out_rgb[3 * o + 2] = b8;
    }
}

// Block coordinates from a linear block id, image QUADRANT by quadrant (top-left, top-right, bottom-left,
// bottom-right), row-major inside a quadrant. Primary rays of one quadrant share the direction octant, so the blocks
// resident on an SM at any time run the same octant copy of the traversal loop (instruction-cache footprint).
__device__ __forceinline__ void quadrant_block(int b, int nbx, int nby, int& bx, int& by, int order)
{
    if (order == 1) { bx = b % nbx; by = b / nbx; return; }       // plain row-major (experiment)
    if (order == 3) {
        // block rows from the image centre outwards, alternating below / above, full width: the rows that usually hold the
        // model - and with it the few blocks that run 10-20x longer than the median - all start in the first wave. For the short
        // per-rank kernels of a multi-GPU frame the kernel's length is "start of the slowest block + its duration" (measured:
        // tools/block_timeline.py), and quadrant after quadrant starts the lower half 55 us late.
        const int r = b / nbx, hy = nby >> 1;
        bx = b - r * nbx;
        int k = (r & 1) ? hy - 1 - (r >> 1) : hy + (r >> 1);      // hy, hy-1, hy+1, hy-2, ...
        if (k < 0) k = r;                                         // the shorter side is used up: the rest in order
        else if (k >= nby) k = nby - 1 - r;
        by = k;
        return;
    }
    const int hx = nbx >> 1, hy = nby >> 1, wx = nbx - hx;
    const int n0 = hx * hy, n1 = wx * hy, n2 = hx * (nby - hy);
    // order 2: inside the two upper quadrants the block rows run from the image centre UP, so every quadrant starts at
    // the centre rows (experiment: heavy rows first)
    if (b < n0) { bx = b % hx; by = b / hx; if (order == 2) by = hy - 1 - by; return; }
    b -= n0;
    if (b < n1) { bx = hx + b % wx; by = b / wx; if (order == 2) by = hy - 1 - by; return; }
    b -= n1;
    if (b < n2) { bx = b % hx; by = hy + b / hx; return; }
    b -= n2;
    bx = hx + b % wx; by = hy + b / wx;
}

// K10: one thread per pixel (warp = 8 x 4 pixels, block = 16 x 8), samples looped in order so the per-pixel float
// accumulation matches main.cpp:553-560.
template <int MODE /*0 = BVH exact, 1 = BVH ordered, 2 = NONE, 3 = KDTREE (any-hit, unshaded: main.cpp:362-372),
                     4 = KDTREE closest hit, shaded (extension: rtds_render_params.kd_closest)*/>
// resident blocks of the one-ray-per-thread kernels, measured (config 4 / bunny 1080p, ms): 8 (64 registers) 1.97 / 0.179, 10 1.99 / 0.190,
// 12 2.09 / 0.193, 16 2.27 / 0.217 - unlike the per-sample shadow kernel, spilling for occupancy loses here
#ifndef RTDS_RK_MINB
#define RTDS_RK_MINB 8
#endif
__global__ void __launch_bounds__(128, RTDS_RK_MINB) render_kernel(const __grid_constant__ RenderArgs A)
{
    __shared__ float4 sh_sph[MODE == 2 ? NONE_CHUNK : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    const int lblock = logical_block(A);
    quadrant_block(lblock, (A.width + 15) / 16, (A.local_rows - A.lrow0 + 7) / 8, bx, by, A.block_order);
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int lrow = A.lrow0 + by * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool active = px < A.width && lrow < A.local_rows;
    const int py = active ? global_row_of(A, lrow) : 0;
    Counters cnt = {0, 0, 0, 0};
    float acc_r = 0, acc_g = 0, acc_b = 0;
    int last_hit = -1;
    const size_t pix = (size_t)py * A.width + px;
    for (int k = 0; k < A.spp; ++k) {
        float dx = 0, dy = 0, dz = -1;
        if (active) {
            const float* dp = A.dirs + 3 * (pix * A.spp + k);      // main.cpp:554-557, computed by mt_expand_dirs_kernel
            dx = __ldg(dp); dy = __ldg(dp + 1); dz = __ldg(dp + 2);
            cnt.rays++;
        }
        float tnear = INFINITY;
        int best_key = 0, best_leaf = -1, hit_obj = -1;
        if (MODE == 2) {
            brute_force_block(A.prim_type, A.sph, A.tri, A.n, active, 0.f, 0.f, 0.f, dx, dy, dz, tnear, hit_obj, cnt, sh_sph);
        } else if (MODE == 3) {
            if (active) hit_obj = kd_any_hit(A.kd, 0.f, 0.f, 0.f, dx, dy, dz, cnt) ? 1 : -1;
        } else if (MODE == 4) {
            if (active) kd_closest_hit(A.kd, 0.f, 0.f, 0.f, dx, dy, dz, tnear, hit_obj, cnt);
        } else if (active) {
            if (MODE == 0) traverse_bvh_exact(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
            else if (A.bvh.wide) traverse_fast<true, false, true, true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
            else traverse_fast<true, false, true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
            if (best_leaf >= 0) hit_obj = __ldg(A.bvh.prim_order + best_leaf);
        }
        if (active) {
            float r, g, b;
            if (hit_obj < 0) { r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2]; }
            else if (MODE == 3) { r = 0.f; g = 0.f; b = 0.f; }       // main.cpp:369: a KD hit is black
            else {
                float4 m = __ldg(A.mat + hit_obj);
                const float hx = 0.f + dx * tnear, hy = 0.f + dy * tnear, hz = 0.f + dz * tnear;     // main.cpp:396
                float nx, ny, nz;
                if (MODE == 2 || MODE == 4) raw_normal(A.prim_type, A.sph, A.tri, (size_t)hit_obj, hx, hy, hz, nx, ny, nz);
                else raw_normal(A.bvh.prim_type, A.bvh.leaf_sph, A.bvh.leaf_tri, (size_t)best_leaf, hx, hy, hz, nx, ny, nz);
                shade_diffuse(A.shade, dx, dy, dz, hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b);
            }
            acc_r += r; acc_g += g; acc_b += b;
            last_hit = hit_obj;
        }
    }
    const float fs = (float)(unsigned)A.spp;
    store_warp_rgb(A, px, lrow, active, (unsigned char)(fminf(1.0f, acc_r / fs) * 255),
                   (unsigned char)(fminf(1.0f, acc_g / fs) * 255), (unsigned char)(fminf(1.0f, acc_b / fs) * 255));
    if (active) {
        const size_t o = (size_t)lrow * A.width + px;
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
    }
    report_block_cost(A, lblock, cnt.node_visits + cnt.prim_tests);
    // counters: warp reduce, one atomic per warp per counter
    unsigned v[4] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

// the shadow query of main.cpp:468-473 for the packet kernel (one ray, any-hit), out of line; same answers as occluded<1>
struct ShadowHit { int occluded; unsigned node_tests, prim_tests, node_visits; };
__device__ __forceinline__ ShadowHit shadow_query(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz, float dist2)
{
    Counters c = {0, 0, 0, 0};
    float ts = INFINITY;
    int bk = 0, bl = -1;
    const float len2 = dx * dx + dy * dy + dz * dz;
    int occ;
    if (fabsf(len2 - 1.0f) < 1e-3f) {
        traverse_fast<false, true>(B, ox, oy, oz, dx, dy, dz, ts, bk, bl, c, dist2);
        occ = bl >= 0;
    } else {
        const ColdHit h = traverse_exact_cold(&B, ox, oy, oz, dx, dy, dz, ts, bk, bl);
        c.node_tests = h.node_tests; c.prim_tests = h.prim_tests; c.node_visits = h.node_visits;
        occ = h.leaf >= 0 && h.tnear * h.tnear < dist2;
    }
    return ShadowHit{occ, c.node_tests, c.prim_tests, c.node_visits};
}
static __device__ __noinline__ ShadowHit shadow_query_cold(const BvhView* B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                          float dist2)
{
    return shadow_query(*B, ox, oy, oz, dx, dy, dz, dist2);
}

// castRay's DIFFUSE_AND_GLOSSY branch with the shadow query evaluated (main.cpp:447-494), as in render_full_kernel
// COLD: the any-hit traversal out of line (callers that hold a packet's state in registers) or inline (one thread = one sample)
template <bool COLD = true>
__device__ __forceinline__ void shade_diffuse_shadowed(const RenderArgs& A, float dx, float dy, float dz, float hx, float hy, float hz,
                                                       float nx, float ny, float nz, float sr, float sg, float sb, float& r, float& g,
                                                       float& b, Counters& cnt, unsigned& shadow_rays)
{
    normalize3(nx, ny, nz);
    if (dx * nx + dy * ny + dz * nz > 0) { nx = -nx; ny = -ny; nz = -nz; }
    const float bias = A.shade.bias;
    float hr = 0, hg = 0, hb = 0;
    for (int i = 0; i < A.shade.n_lights; ++i) {
        const RtdsLight& L = A.shade.lights[i];
        float lx = L.c[0] - hx, ly = L.c[1] - hy, lz = L.c[2] - hz;
        const float dist2 = lx * lx + ly * ly + lz * lz;
        normalize3(lx, ly, lz);
        const float LdotN = fmaxf(0.f, lx * nx + ly * ny + lz * nz);
        const bool front = dx * nx + dy * ny + dz * nz < 0;
        const float sx = front ? hx + nx * bias : hx - nx * bias;
        const float sy = front ? hy + ny * bias : hy - ny * bias;
        const float sz = front ? hz + nz * bias : hz - nz * bias;
        cnt.rays++;
        shadow_rays++;
        // the answer only scales the diffuse term by 0 or 1 (main.cpp:471-472), and that term is (Le * lit) * LdotN: for a light
        // behind the surface (LdotN == 0) both answers give the same float, so the ray is counted but not traced
        float lit = 1.0f;
        if (LdotN > 0.f) {
            const ShadowHit sh = COLD ? shadow_query_cold(&A.bvh, sx, sy, sz, lx, ly, lz, dist2) : shadow_query(A.bvh, sx, sy, sz, lx, ly, lz, dist2);
            cnt.node_tests += sh.node_tests; cnt.prim_tests += sh.prim_tests; cnt.node_visits += sh.node_visits;
            lit = sh.occluded ? 0.0f : 1.0f;
        }
        const float ar = (L.le[0] * lit) * LdotN, ag = (L.le[1] * lit) * LdotN, ab = (L.le[2] * lit) * LdotN;
        const float ix = -lx, iy = -ly, iz = -lz;
        const float s2 = 2 * (ix * nx + iy * ny + iz * nz);
        const float qx = ix - nx * s2, qy = iy - ny * s2, qz = iz - nz * s2;
        const float sp = pow25f(fmaxf(0.f, -(qx * dx + qy * dy + qz * dz)));
        hr += (ar * (0.815f * 0.8f)) / 2.0f + (L.le[0] * sp) * 0.5f;
        hg += (ag * (0.235f * 0.8f)) / 2.0f + (L.le[1] * sp) * 0.5f;
        hb += (ab * (0.031f * 0.8f)) / 2.0f + (L.le[2] * sp) * 0.5f;
        hr += sr; hg += sg; hb += sb;
    }
    r = hr; g = hg; b = hb;
}

// K10, packet form: one thread per pixel, the pixel's samples traced four at a time by traverse_packet. Ordered
// (exact = 0) BVH / LBVH traversal of sphere scenes with aa_samples % 4 == 0; everything else uses render_kernel.
#ifndef RTDS_PK_THREADS
#define RTDS_PK_THREADS 128   // block of the packet kernels: a 16 x (RTDS_PK_THREADS / 16) pixel tile, one 8 x 4 sub-tile per warp (128 or 64).
                              // Measured (bench frame, same registers): 128 x 6 blocks 0.947 ms, 64 x 12 0.971, 64 x 10 0.994
#endif
constexpr int PK_TILE_H = RTDS_PK_THREADS / 16;
#ifndef RTDS_PK_MINB
#define RTDS_PK_MINB 6   // measured on B200: 4 blocks (112 regs) 1.40 ms, 5 (96) 1.245, 6 (80, 188 B spilled) 1.223, 8 (64) 1.238
#endif
// closest hits of four primary rays (origin 0, dz < 0): one packet when they share a direction octant and the ordered
// traversal's preconditions hold, four single-ray traversals (out of line) otherwise
template <bool HULL>
__device__ __forceinline__ void trace_packet4(const RenderArgs& A, const float (&dx)[PK], const float (&dy)[PK], const float (&dz)[PK],
                                              float margin, float (&tnear)[PK], int (&best_leaf)[PK], Counters& cnt)
{
    float ix[PK], iy[PK], iz[PK];
    int best_key[PK];
    bool ok = A.bvh.root_ref >= 0;
    int oct0 = 0;
#pragma unroll
    for (int j = 0; j < PK; ++j) {
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix[j]) : "f"(dx[j]));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy[j]) : "f"(dy[j]));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz[j]) : "f"(dz[j]));
        const float amin = fminf(fminf(fabsf(ix[j]), fabsf(iy[j])), fabsf(iz[j]));
        const float amax = fmaxf(fmaxf(fabsf(ix[j]), fabsf(iy[j])), fabsf(iz[j]));
        const int oct = (dx[j] < 0 ? 1 : 0) | (dy[j] < 0 ? 2 : 0) | 4;
        if (j == 0) oct0 = oct;
        ok = ok && amin > 1e-30f && amax < 1e30f && oct == oct0;
        tnear[j] = INFINITY; best_key[j] = 0; best_leaf[j] = -1;
    }
    if (ok) {
        switch (oct0) {
            case 4: traverse_packet<4, HULL>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
            case 5: traverse_packet<5, HULL>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
            case 6: traverse_packet<6, HULL>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
            default: traverse_packet<7, HULL>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
        }
    } else {
#pragma unroll
        for (int j = 0; j < PK; ++j) {
            const ColdHit h = trace_primary_cold(&A.bvh, dx[j], dy[j], dz[j]);
            tnear[j] = h.tnear; best_leaf[j] = h.leaf;
            cnt.node_tests += h.node_tests; cnt.prim_tests += h.prim_tests; cnt.node_visits += h.node_visits;
        }
    }
}

// colour of one primary ray of a packet (sphere leaves), main.cpp:394-497
template <bool SHADOWS>
__device__ __forceinline__ int shade_packet_ray(const RenderArgs& A, float dx, float dy, float dz, float tnear, int best_leaf, float& r,
                                                float& g, float& b, Counters& cnt, unsigned& shadow_rays)
{
    int hit_obj = -1;
    if (best_leaf >= 0) hit_obj = __ldg(A.bvh.prim_order + best_leaf);
    if (hit_obj < 0) { r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2]; }
    else {
        const float4 m = __ldg(A.mat + hit_obj);
        const float hx = 0.f + dx * tnear, hy = 0.f + dy * tnear, hz = 0.f + dz * tnear;     // main.cpp:396
        float nx, ny, nz;
        raw_normal(0, A.bvh.leaf_sph, nullptr, (size_t)best_leaf, hx, hy, hz, nx, ny, nz);
        if (SHADOWS) shade_diffuse_shadowed(A, dx, dy, dz, hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b, cnt, shadow_rays);
        else shade_diffuse(A.shade, dx, dy, dz, hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b);
    }
    return hit_obj;
}

// the same, out of line (RTDS_SHADE_OOL experiment): one copy of the shading code per kernel instead of four
struct ShadedRay { float r, g, b; int hit; };
static __device__ __noinline__ ShadedRay shade_packet_ray_ool(const RenderArgs* A, float dx, float dy, float dz, float tnear, int best_leaf)
{
    ShadedRay o;
    Counters c = {0, 0, 0, 0};
    unsigned sr = 0;
    o.hit = shade_packet_ray<false>(*A, dx, dy, dz, tnear, best_leaf, o.r, o.g, o.b, c, sr);
    return o;
}

// castRay's non-diffuse material branches for one ray of a packet (defined with render_full_kernel's pieces below)
struct MatResult { float r, g, b; unsigned node_tests, prim_tests, node_visits, rays, shadow_rays, secondary_rays; };
static __device__ __noinline__ MatResult cast_material_cold(const RenderArgs* A, float dx, float dy, float dz, float tnear, int hit_obj, int hit_leaf);

#ifdef RTDS_BLOCK_TIMING     // profiling variant (tools/block_timeline.py): start / end / SM of every block of the packet kernel
__device__ unsigned long long g_block_times[3 * 70000];
#endif
template <bool SHADOWS /*evaluate the shadow query (extension; the reference's trace_more is a stub)*/,
          bool HULL /*interior boxes tested once per packet against the hull of the four reciprocal directions*/,
          bool MATERIALS = false /*the scene holds REFLECTION_AND_REFRACTION / REFLECTION primitives: rays that hit one continue
                                   through castRay's material branches (single rays, out of line); diffuse hits are shaded as always*/>
__global__ void __launch_bounds__(RTDS_PK_THREADS, RTDS_PK_MINB) render_packet_kernel(const __grid_constant__ RenderArgs A)
{
#ifdef RTDS_BLOCK_TIMING
    unsigned long long bt0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(bt0));
#endif
    unsigned shadow_rays = 0, secondary_rays = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    const int lblock = logical_block(A);
    if (A.skip && __ldg(A.skip + lblock)) return;      // lpt_split: this tile is render_heavy_kernel's (block-uniform)
    quadrant_block(lblock, (A.width + 15) / 16, (A.local_rows - A.lrow0 + PK_TILE_H - 1) / PK_TILE_H, bx, by, A.block_order);
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int lrow = A.lrow0 + by * PK_TILE_H + (warp >> 1) * 4 + (lane >> 3);
    const bool active = px < A.width && lrow < A.local_rows;
    Counters cnt = {0, 0, 0, 0};
    unsigned char r8 = 0, g8 = 0, b8 = 0;
    if (active) {
        const int py = global_row_of(A, lrow);
        float acc_r = 0, acc_g = 0, acc_b = 0;
        int last_hit = -1;
        const size_t pix = (size_t)py * A.width + px;
        const float margin = prune_margin(A.bvh.root_box, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < A.spp; k0 += PK) {
            float dx[PK], dy[PK], dz[PK], tnear[PK];
            int best_leaf[PK];
            // the packet's 4 directions are 48 contiguous, 16-byte aligned bytes (main.cpp:554-557, from mt_expand_dirs_kernel)
            const float4* dp = reinterpret_cast<const float4*>(A.dirs + 3 * (pix * A.spp + k0));
            // read once: streaming loads (evict-first) leave L1/L2 to the tree; measured 0.948 -> 0.942 ms on the bench frame
            const float4 d0 = __ldcs(dp), d1 = __ldcs(dp + 1), d2 = __ldcs(dp + 2);
            dx[0] = d0.x; dy[0] = d0.y; dz[0] = d0.z; dx[1] = d0.w; dy[1] = d1.x; dz[1] = d1.y;
            dx[2] = d1.z; dy[2] = d1.w; dz[2] = d2.x; dx[3] = d2.y; dy[3] = d2.z; dz[3] = d2.w;
            cnt.rays += PK;
            trace_packet4<HULL>(A, dx, dy, dz, margin, tnear, best_leaf, cnt);
#pragma unroll
            for (int j = 0; j < PK; ++j) {
                float r, g, b;
                if (MATERIALS && best_leaf[j] >= 0) {
                    const int obj = __ldg(A.bvh.prim_order + best_leaf[j]);
                    if (__ldg(&A.mat[obj].w) != 0.0f) {
                        const MatResult mr = cast_material_cold(&A, dx[j], dy[j], dz[j], tnear[j], obj, best_leaf[j]);
                        cnt.node_tests += mr.node_tests; cnt.prim_tests += mr.prim_tests; cnt.node_visits += mr.node_visits; cnt.rays += mr.rays;
                        shadow_rays += mr.shadow_rays; secondary_rays += mr.secondary_rays;
                        acc_r += mr.r; acc_g += mr.g; acc_b += mr.b;     // sample order, main.cpp:553-560
                        last_hit = obj;
                        continue;
                    }
                }
#ifndef RTDS_SHADE_INLINE
                if (!SHADOWS) { const ShadedRay sh = shade_packet_ray_ool(&A, dx[j], dy[j], dz[j], tnear[j], best_leaf[j]); r = sh.r; g = sh.g; b = sh.b; last_hit = sh.hit; }
                else
#endif
                last_hit = shade_packet_ray<SHADOWS>(A, dx[j], dy[j], dz[j], tnear[j], best_leaf[j], r, g, b, cnt, shadow_rays);
                acc_r += r; acc_g += g; acc_b += b;     // sample order, main.cpp:553-560
            }
        }
        const float fs = (float)(unsigned)A.spp;
        r8 = (unsigned char)(fminf(1.0f, acc_r / fs) * 255);
        g8 = (unsigned char)(fminf(1.0f, acc_g / fs) * 255);
        b8 = (unsigned char)(fminf(1.0f, acc_b / fs) * 255);
        const size_t o = (size_t)lrow * A.width + px;
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
    }
    store_warp_rgb(A, px, lrow, active, r8, g8, b8);
#ifdef RTDS_BLOCK_TIMING
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < 70000) {
        unsigned long long bt1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(bt1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_block_times[3 * blockIdx.x] = bt0; g_block_times[3 * blockIdx.x + 1] = bt1; g_block_times[3 * blockIdx.x + 2] = ((unsigned long long)smid << 32) | cnt.node_visits;
    }
#endif
    report_block_cost(A, lblock, cnt.node_visits + cnt.prim_tests);
    unsigned v[6] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays, shadow_rays, secondary_rays};
#pragma unroll
    for (int c = 0; c < (MATERIALS ? 6 : (SHADOWS ? 5 : 4)); ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

// ===================================================================================================
// K10 + K11 fused ("strip" kernel): one block per MT19937 snapshot chunk (8 regenerations = 4,992 words = 1,248
// samples). The block regenerates its chunk of the reference's jitter stream into SHARED memory (plus one more
// regeneration for a pixel whose samples straddle the chunk end) and renders the pixels whose first sample lies in
// the chunk — consecutive pixels in scanline order, i.e. the order the reference consumes the stream in. The 531 MB
// of jitter words a 4K x 4 spp frame needs are never written to or read from HBM. Warps hold 32 consecutive pixels
// of a row; with sub-pixel primitives the traversal is as coherent as with 8x4 tiles (measured).
// ===================================================================================================
constexpr int STRIP_THREADS = 128;
constexpr int STRIP_REGENS = MT_SNAP_EVERY + 1;
constexpr int STRIP_SAMPLES = MT_SNAP_EVERY * MT_N / 4;   // 1,248 samples per chunk

// regeneration A -> B by a block of 128 threads (two passes per 227-wide phase)
__device__ __forceinline__ void mt_regen_128(const uint32_t* __restrict__ A, uint32_t* __restrict__ B)
{
    const int t = threadIdx.x;
    for (int i = t; i < 227; i += STRIP_THREADS) B[i] = A[i + MT_M] ^ mt_twist(A[i], A[i + 1]);
    __syncthreads();
    for (int i = t; i < 227; i += STRIP_THREADS) B[i + 227] = B[i] ^ mt_twist(A[i + 227], A[i + 228]);
    __syncthreads();
    for (int i = t; i < 169; i += STRIP_THREADS) B[i + 454] = B[i + 227] ^ mt_twist(A[i + 454], A[i + 455]);
    if (t == 0) B[623] = B[396] ^ mt_twist(A[623], B[0]);
    __syncthreads();
}

template <int MODE /*0 = BVH exact, 1 = BVH ordered, 3 = KDTREE*/>
__global__ void __launch_bounds__(STRIP_THREADS) render_strip_kernel(const __grid_constant__ RenderArgs A)
{
    __shared__ uint32_t S[2][MT_N];
    __shared__ __align__(16) uint32_t words[STRIP_REGENS * MT_N];
    const int lane = threadIdx.x & 31;
    const unsigned long long chunk = (unsigned long long)A.chunk_first + blockIdx.x;
    const unsigned long long s_lo = chunk * STRIP_SAMPLES, s_hi = s_lo + STRIP_SAMPLES;     // absolute sample range
    const unsigned long long n_pix = (unsigned long long)A.width * A.height;
    // pixels whose first sample (first_sample + p*spp) lies in [s_lo, s_hi)
    unsigned long long p_lo = s_lo > A.first_sample ? (s_lo - A.first_sample + A.spp - 1) / A.spp : 0;
    unsigned long long p_hi = s_hi > A.first_sample ? (s_hi - A.first_sample + A.spp - 1) / A.spp : 0;
    if (p_hi > n_pix) p_hi = n_pix;
    if (p_lo >= p_hi) return;
    if (A.world > 1) {   // skip chunks that hold no row of this rank
        const int y0 = (int)(p_lo / A.width), y1 = (int)((p_hi - 1) / A.width);
        bool mine = false;
        for (int t = y0 / A.tile_rows; t <= y1 / A.tile_rows; ++t) mine |= (t % A.world) == A.rank;
        if (!mine) return;
    }
    // regenerate the chunk (+1 regeneration) from its snapshot
    {
        const uint32_t* src = A.mt_snap + chunk * MT_N;
        for (int i = threadIdx.x; i < MT_N; i += STRIP_THREADS) S[0][i] = __ldg(src + i);
        __syncthreads();
        int cur = 0;
        for (int r = 0; r < STRIP_REGENS; ++r) {
            mt_regen_128(S[cur], S[cur ^ 1]);
            cur ^= 1;
            for (int i = threadIdx.x; i < MT_N; i += STRIP_THREADS) words[r * MT_N + i] = mt_temper(S[cur][i]);
        }
        __syncthreads();
    }
    Counters cnt = {0, 0, 0, 0};
    for (unsigned long long p = p_lo + threadIdx.x; p < p_hi; p += STRIP_THREADS) {
        const int py = (int)(p / A.width), px = (int)(p - (unsigned long long)py * A.width);
        const int tile = py / A.tile_rows;
        if (A.world > 1 && (tile % A.world) != A.rank) continue;
        const int lrow = (tile / A.world) * A.tile_rows + (py - tile * A.tile_rows);
        const unsigned woff = (unsigned)((A.first_sample + p * A.spp - s_lo) * 4);   // word offset of the pixel's first sample
        float acc_r = 0, acc_g = 0, acc_b = 0;
        int last_hit = -1;
        for (int k = 0; k < A.spp; ++k) {
            const uint4 jw = *reinterpret_cast<const uint4*>(&words[woff + 4 * k]);
            float dx, dy, dz;
            primary_dir(jw, px, py, RayGen{A.angle, A.aspect, A.inv_w, A.inv_h}, dx, dy, dz);
            cnt.rays++;
            float tnear = INFINITY;
            int best_key = 0, best_leaf = -1, hit_obj = -1;
            if (MODE == 3) hit_obj = kd_any_hit(A.kd, 0.f, 0.f, 0.f, dx, dy, dz, cnt) ? 1 : -1;
            else {
                if (MODE == 0) traverse_bvh_exact(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
                else traverse_fast<true, false, true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
                if (best_leaf >= 0) hit_obj = __ldg(A.bvh.prim_order + best_leaf);
            }
            float r, g, b;
            if (hit_obj < 0) { r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2]; }
            else if (MODE == 3) { r = 0.f; g = 0.f; b = 0.f; }
            else {
                float4 m = __ldg(A.mat + hit_obj);
                const float hx = 0.f + dx * tnear, hy = 0.f + dy * tnear, hz = 0.f + dz * tnear;
                float nx, ny, nz;
                raw_normal(A.bvh.prim_type, A.bvh.leaf_sph, A.bvh.leaf_tri, (size_t)best_leaf, hx, hy, hz, nx, ny, nz);
                shade_diffuse(A.shade, dx, dy, dz, hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b);
            }
            acc_r += r; acc_g += g; acc_b += b;
            last_hit = hit_obj;
        }
        const size_t o = (size_t)lrow * A.width + px;
        const float fs = (float)(unsigned)A.spp;
        A.out_rgb[3 * o]     = (unsigned char)(fminf(1.0f, acc_r / fs) * 255);
        A.out_rgb[3 * o + 1] = (unsigned char)(fminf(1.0f, acc_g / fs) * 255);
        A.out_rgb[3 * o + 2] = (unsigned char)(fminf(1.0f, acc_b / fs) * 255);
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
    }
    unsigned v[4] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

// ===================================================================================================
// K10b: castRay in full (main.cpp:291-500) — material branches, recursion to depth 2 unrolled into a loop, and the
// shadow query whose contract main.cpp:468-473 states (trace_more itself is a stub: shadows are an extension).
// Used when the scene holds non-diffuse materials or shadows are requested; the all-diffuse, shadow-free path
// (everything the reference's own loader can produce) stays on render_kernel above.
// ===================================================================================================
__device__ __forceinline__ float clampf3(float lo, float hi, float v) { return fmaxf(lo, fminf(hi, v)); }   // main.cpp:75-79

__device__ __forceinline__ float fresnel_kr(float ix, float iy, float iz, float nx, float ny, float nz, float ior)   // main.cpp:85-103
{
    float cosi = clampf3(-1.f, 1.f, ix * nx + iy * ny + iz * nz);
    float etai = 1, etat = ior;
    if (cosi > 0) { float t = etai; etai = etat; etat = t; }
    float sint = etai / etat * sqrtf(fmaxf(0.f, 1 - cosi * cosi));
    if (sint >= 1) return 1.f;
    float cost = sqrtf(fmaxf(0.f, 1 - sint * sint));
    cosi = fabsf(cosi);
    float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
    float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
    return (Rs * Rs + Rp * Rp) / 2;
}

__device__ __forceinline__ void refract_dir(float ix, float iy, float iz, float nx, float ny, float nz, float ior, float& rx, float& ry,
                                            float& rz)   // main.cpp:223-233
{
    float cosi = clampf3(-1.f, 1.f, ix * nx + iy * ny + iz * nz);
    float etai = 1, etat = ior;
    if (cosi < 0) cosi = -cosi;
    else { float t = etai; etai = etat; etat = t; nx = -nx; ny = -ny; nz = -nz; }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) { rx = ry = rz = 0.f; return; }
    float f = eta * cosi - sqrtf(k);
    rx = ix * eta + nx * f; ry = iy * eta + ny * f; rz = iz * eta + nz * f;
}

template <int MODE>
__device__ __forceinline__ void closest_hit(const RenderArgs& A, float ox, float oy, float oz, float dx, float dy, float dz, float& tnear,
                                            int& hit_obj, int& hit_leaf, Counters& cnt)
{
    tnear = INFINITY; hit_obj = -1; hit_leaf = -1;
    if (MODE == 2) {   // main.cpp:376-386, straight from global memory (all lanes read the same address)
        for (int i = 0; i < A.n; ++i) {
            float t0, t1;
            if (obj_test(A.prim_type, A.sph, A.tri, i, ox, oy, oz, dx, dy, dz, t0, t1)) {
                if (t0 < 0) t0 = t1;
                if (t0 < tnear) { tnear = t0; hit_obj = i; }
            }
        }
        cnt.prim_tests += A.n;
    } else {
        int best_key = 0, best_leaf = -1;
        float len2 = dx * dx + dy * dy + dz * dz;
        if (MODE == 0 || !(fabsf(len2 - 1.0f) < 1e-3f)) traverse_bvh_exact(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
        else traverse_fast<false>(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
        if (best_leaf >= 0) { hit_obj = __ldg(A.bvh.prim_order + best_leaf); hit_leaf = best_leaf; }
    }
}

// the shadow query of main.cpp:468-473: is there a hit with tNearShadow^2 < lightDistance2?
template <int MODE>
__device__ __forceinline__ bool occluded(const RenderArgs& A, float ox, float oy, float oz, float dx, float dy, float dz, float dist2, Counters& cnt)
{
    if (MODE == 1) {
        float len2 = dx * dx + dy * dy + dz * dz;
        if (fabsf(len2 - 1.0f) < 1e-3f) {
            float ts = INFINITY; int bk = 0, bl = -1;
            traverse_fast<false, true>(A.bvh, ox, oy, oz, dx, dy, dz, ts, bk, bl, cnt, dist2);
            return bl >= 0;
        }
    }
    float ts; int sh, shl;
    closest_hit<MODE>(A, ox, oy, oz, dx, dy, dz, ts, sh, shl, cnt);
    return sh >= 0 && ts * ts < dist2;
}

// castRay (main.cpp:291-500) for one primary ray, recursion to depth 2 unrolled into a loop. HAVE_FIRST: the primary ray's closest
// hit (tnear0 / hit_obj0 / hit_leaf0) was already found by the caller (the packet kernel) and is not traced again.
template <int MODE /*0 = BVH exact, 1 = BVH ordered, 2 = NONE*/, bool HAVE_FIRST>
__device__ __forceinline__ void cast_ray_full(const RenderArgs& A, float dx, float dy, float dz, float tnear0, int hit_obj0, int hit_leaf0,
                                              float& r, float& g, float& b, int& first_hit, Counters& cnt, unsigned& shadow_rays,
                                              unsigned& secondary_rays)
{
    const float bias = A.shade.bias;
    float ox = 0, oy = 0, oz = 0;
    r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2];
    float mult[2] = {1.f, 1.f};   // (1 - kr) of the REFLECTION_AND_REFRACTION hits on the way, outermost first
    int nmult = 0;
    first_hit = -1;
    for (int depth = 1;; ++depth) {
        if (depth > A.shade.max_depth) break;                        // main.cpp:311-313: sky
        float tnear;
        int hit_obj, hit_leaf;
        if (HAVE_FIRST && depth == 1) { tnear = tnear0; hit_obj = hit_obj0; hit_leaf = hit_leaf0; }
        else {
            closest_hit<MODE>(A, ox, oy, oz, dx, dy, dz, tnear, hit_obj, hit_leaf, cnt);
            cnt.rays++;
        }
        if (depth == 1) first_hit = hit_obj;
        if (hit_obj < 0) break;                                      // sky
        const float4 m = __ldg(A.mat + hit_obj);
        const int material = (int)m.w;
        float hx = ox + dx * tnear, hy = oy + dy * tnear, hz = oz + dz * tnear;
        float nx, ny, nz;
        if (MODE == 2) raw_normal(A.prim_type, A.sph, A.tri, (size_t)hit_obj, hx, hy, hz, nx, ny, nz);
        else raw_normal(A.bvh.prim_type, A.bvh.leaf_sph, A.bvh.leaf_tri, (size_t)hit_leaf, hx, hy, hz, nx, ny, nz);
        normalize3(nx, ny, nz);
        if (dx * nx + dy * ny + dz * nz > 0) { nx = -nx; ny = -ny; nz = -nz; }
        if (material == RTDS_REFLECTION_AND_REFRACTION) {            // main.cpp:418-434
            float rx, ry, rz;
            refract_dir(dx, dy, dz, nx, ny, nz, 3.f, rx, ry, rz);
            normalize3(rx, ry, rz);
            const bool neg = rx * nx + ry * ny + rz * nz < 0;
            const float kr = fresnel_kr(dx, dy, dz, nx, ny, nz, 2.f);
            if (nmult < 2) mult[nmult++] = 1 - kr;
            ox = neg ? hx - nx * bias : hx + nx * bias;
            oy = neg ? hy - ny * bias : hy + ny * bias;
            oz = neg ? hz - nz * bias : hz + nz * bias;
            dx = rx; dy = ry; dz = rz;
            if (depth + 1 <= A.shade.max_depth) secondary_rays++;   // a deeper ray returns the sky untraced
            // (the reflection ray of main.cpp:428 is traced by the reference but its colour is discarded)
            continue;
        }
        if (material == RTDS_REFLECTION) {                           // main.cpp:435-446
            const float kr = fresnel_kr(dx, dy, dz, nx, ny, nz, 2.f);
            r = g = b = 1 - kr;
            break;
        }
        float hr = 0, hg = 0, hb = 0;                                // main.cpp:447-494
        for (int i = 0; i < A.shade.n_lights; ++i) {
            const RtdsLight& L = A.shade.lights[i];
            float lx = L.c[0] - hx, ly = L.c[1] - hy, lz = L.c[2] - hz;
            const float dist2 = lx * lx + ly * ly + lz * lz;
            normalize3(lx, ly, lz);
            const float LdotN = fmaxf(0.f, lx * nx + ly * ny + lz * nz);
            float lit = 1.0f;
            if (A.shade.shadows) {
                const bool front = dx * nx + dy * ny + dz * nz < 0;
                float sx = front ? hx + nx * bias : hx - nx * bias;
                float sy = front ? hy + ny * bias : hy - ny * bias;
                float sz = front ? hz + nz * bias : hz - nz * bias;
                cnt.rays++;
                shadow_rays++;
                // fast form: a light behind the surface (LdotN == 0) gives the same float for either answer - counted, not traced
                // (see shade_diffuse_shadowed); the exact and brute-force forms keep the reference's intersection-test counts
                if ((MODE != 1 || LdotN > 0.f) && occluded<MODE>(A, sx, sy, sz, lx, ly, lz, dist2, cnt)) lit = 0.0f;   // main.cpp:471-472
            }
            const float ar = (L.le[0] * lit) * LdotN, ag = (L.le[1] * lit) * LdotN, ab = (L.le[2] * lit) * LdotN;
            const float ix = -lx, iy = -ly, iz = -lz;
            const float s2 = 2 * (ix * nx + iy * ny + iz * nz);
            const float qx = ix - nx * s2, qy = iy - ny * s2, qz = iz - nz * s2;
            const float sp = pow25f(fmaxf(0.f, -(qx * dx + qy * dy + qz * dz)));
            hr += (ar * (0.815f * 0.8f)) / 2.0f + (L.le[0] * sp) * 0.5f;
            hg += (ag * (0.235f * 0.8f)) / 2.0f + (L.le[1] * sp) * 0.5f;
            hb += (ab * (0.031f * 0.8f)) / 2.0f + (L.le[2] * sp) * 0.5f;
            hr += m.x; hg += m.y; hb += m.z;
        }
        r = hr; g = hg; b = hb;
        break;
    }
    for (int i = nmult - 1; i >= 0; --i) { r = r * mult[i]; g = g * mult[i]; b = b * mult[i]; }   // main.cpp:432, innermost first
}

// a packet ray that hit a REFLECTION_AND_REFRACTION / REFLECTION primitive: castRay's material branches from the known primary
// hit on (out of line: one copy per kernel, off the packet loop's hot path)
static __device__ __noinline__ MatResult cast_material_cold(const RenderArgs* A, float dx, float dy, float dz, float tnear, int hit_obj, int hit_leaf)
{
    MatResult o;
    Counters c = {0, 0, 0, 0};
    unsigned sh = 0, sec = 0;
    int first;
    cast_ray_full<1, true>(*A, dx, dy, dz, tnear, hit_obj, hit_leaf, o.r, o.g, o.b, first, c, sh, sec);
    o.node_tests = c.node_tests; o.prim_tests = c.prim_tests; o.node_visits = c.node_visits; o.rays = c.rays; o.shadow_rays = sh; o.secondary_rays = sec;
    return o;
}

template <int MODE /*0 = BVH exact, 1 = BVH ordered, 2 = NONE*/>
__global__ void __launch_bounds__(128) render_full_kernel(const __grid_constant__ RenderArgs A)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    const int lblock = logical_block(A);
    quadrant_block(lblock, (A.width + 15) / 16, (A.local_rows - A.lrow0 + 7) / 8, bx, by, A.block_order);
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int lrow = A.lrow0 + by * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool active = px < A.width && lrow < A.local_rows;
    Counters cnt = {0, 0, 0, 0};
    unsigned shadow_rays = 0, secondary_rays = 0;
    unsigned char r8 = 0, g8 = 0, b8 = 0;
    if (active) {
        const int py = global_row_of(A, lrow);
        float acc_r = 0, acc_g = 0, acc_b = 0;
        int last_hit = -1;
        const size_t pix = (size_t)py * A.width + px;
        for (int k = 0; k < A.spp; ++k) {
            const float* dp = A.dirs + 3 * (pix * A.spp + k);      // main.cpp:554-557, computed by mt_expand_dirs_kernel
            const float dx = __ldg(dp), dy = __ldg(dp + 1), dz = __ldg(dp + 2);
            float r, g, b;
            cast_ray_full<MODE, false>(A, dx, dy, dz, 0.f, -1, -1, r, g, b, last_hit, cnt, shadow_rays, secondary_rays);
            acc_r += r; acc_g += g; acc_b += b;
        }
        size_t o = (size_t)lrow * A.width + px;
        float fs = (float)(unsigned)A.spp;
        r8 = (unsigned char)(fminf(1.0f, acc_r / fs) * 255);
        g8 = (unsigned char)(fminf(1.0f, acc_g / fs) * 255);
        b8 = (unsigned char)(fminf(1.0f, acc_b / fs) * 255);
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
    }
    store_warp_rgb(A, px, lrow, active, r8, g8, b8);
    report_block_cost(A, lblock, cnt.node_visits + cnt.prim_tests);
    unsigned v[6] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays, shadow_rays, secondary_rays};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

// ===================================================================================================
// K10d: the heaviest tiles of the packet kernel, one RAY per thread (lpt_split). What bounds a rank's kernel in the strong-scaling
// run is a handful of 16 x 8 tiles whose packets graze hundreds of leaf boxes: in the packet kernel one thread
// walks its pixel's four samples' leaf tests one after the other. The tiles that were heaviest in the previous frame are therefore
// rendered here instead - four blocks per tile, two pixel rows each, a thread per (pixel, sample), four neighbouring lanes per
// pixel - on the highest-priority stream, while render_packet_kernel skips them (A.skip). Same rays, same single-ray traversal the
// packet traversal is tested against, same shading function, sums in sample order through shuffles: same bytes.
// ===================================================================================================
__global__ void __launch_bounds__(128) render_heavy_kernel(const __grid_constant__ RenderArgs A)
{
    static_assert(PK_TILE_H == 8, "render_heavy_kernel covers a 16 x 8 tile with 4 blocks of 2 rows");
    const int tile = __ldg(A.heavy_list + (blockIdx.x >> 2));
    if (tile < 0) return;
    const int lane = threadIdx.x & 31, q = blockIdx.x & 3;
    int bx, by;
    quadrant_block(tile, (A.width + 15) / 16, (A.local_rows - A.lrow0 + PK_TILE_H - 1) / PK_TILE_H, bx, by, A.block_order);
    const int p = threadIdx.x >> 2, k = threadIdx.x & 3;
    const int px = bx * 16 + (p & 15), lrow = A.lrow0 + by * PK_TILE_H + 2 * q + (p >> 4);
    const bool active = px < A.width && lrow < A.local_rows;
    Counters cnt = {0, 0, 0, 0};
    float acc_r = 0, acc_g = 0, acc_b = 0;
    int last_hit = -1;
    unsigned est_visits = 0, est_prims = 0;
    const size_t pix = active ? (size_t)global_row_of(A, lrow) * A.width + px : 0;
    const int base = lane & ~3;
    for (int k0 = 0; k0 < A.spp; k0 += PK) {
        float r = 0.f, g = 0.f, b = 0.f;
        int hit = -1;
        if (active) {
            const float* dp = A.dirs + 3 * (pix * A.spp + k0 + k);
            const float dx = __ldcs(dp), dy = __ldcs(dp + 1), dz = __ldcs(dp + 2);
            float tnear = INFINITY;
            int key = 0, leaf = -1;
            Counters c1 = {0, 0, 0, 0};
            traverse_fast<true, false, true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, key, leaf, c1);
            cnt.node_tests += c1.node_tests; cnt.prim_tests += c1.prim_tests; cnt.node_visits += c1.node_visits; cnt.rays++;
            est_visits = max(est_visits, c1.node_visits); est_prims += c1.prim_tests;
            const ShadedRay sh = shade_packet_ray_ool(&A, dx, dy, dz, tnear, leaf);
            r = sh.r; g = sh.g; b = sh.b; hit = sh.hit;
        }
#pragma unroll
        for (int j = 0; j < PK; ++j) {                                          // sample order, main.cpp:553-560
            acc_r += __shfl_sync(0xffffffffu, r, base + j); acc_g += __shfl_sync(0xffffffffu, g, base + j); acc_b += __shfl_sync(0xffffffffu, b, base + j);
        }
        last_hit = __shfl_sync(0xffffffffu, hit, base + PK - 1);
    }
    if (active && k == 0) {
        const float fs = (float)(unsigned)A.spp;
        const unsigned char r8 = (unsigned char)(fminf(1.0f, acc_r / fs) * 255);
        const unsigned char g8 = (unsigned char)(fminf(1.0f, acc_g / fs) * 255);
        const unsigned char b8 = (unsigned char)(fminf(1.0f, acc_b / fs) * 255);
        const size_t o = (size_t)lrow * A.width + px;
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
        const size_t oo = (size_t)(A.out_global_rows ? global_row_of(A, lrow) : lrow) * A.width + px;
        A.out_rgb[3 * oo] = r8; A.out_rgb[3 * oo + 1] = g8; A.out_rgb[3 * oo + 2] = b8;
    }
    // the tile's cost for the next frame's order, in the packet kernel's currency (a packet's node visits ~ its longest ray's,
    // its primitive tests = the four rays' together): the largest such pixel
    unsigned pv = est_visits, pp = est_prims;
    pv = max(pv, __shfl_xor_sync(0xffffffffu, pv, 1)); pv = max(pv, __shfl_xor_sync(0xffffffffu, pv, 2));
    pp += __shfl_xor_sync(0xffffffffu, pp, 1); pp += __shfl_xor_sync(0xffffffffu, pp, 2);
    report_block_cost(A, tile, pv + pp);
    unsigned v[4] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned x = __reduce_add_sync(0xffffffffu, v[c]);
        if (lane == 0 && x) atomicAdd(&A.counters[c], (unsigned long long)x);
    }
}

// ===================================================================================================
// K10c: castRay with shadow rays as a WAVEFRONT (frames with shadow rays, aa_samples % 4 == 0 and <= 128, sphere leaves). The
// single-kernel form (render_full_kernel, one thread = one pixel = 16 samples x (1 primary + n_lights shadow rays) one after the
// other) runs with 12.6 of 32 lanes active on config 5 (ncu, profiles/ncu_r02u_kernels.md): the lanes of a warp are in different
// phases of different rays. Here the frame is two kernels over the same sample set, each with warps full of like work:
//   wave_primary_kernel  one thread per pixel, its samples four at a time as packets (traverse_packet): tnear + leaf per sample
//   wave_shade_kernel    one thread per SAMPLE: the shadow rays of the hit (a warp = 32 consecutive samples - two pixels' worth -
//                        towards one light at a time: one direction octant, origins a pixel apart), castRay's shading, material
//                        hits through cast_material_cold; then the block adds each pixel's samples in order (main.cpp:553-560)
// Same arithmetic per ray as the single-kernel form: identical hit ids, float sums, bytes and ray counts (tests run both).
// ===================================================================================================
__global__ void __launch_bounds__(RTDS_PK_THREADS, RTDS_PK_MINB) wave_primary_kernel(const __grid_constant__ RenderArgs A)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    const int lblock = logical_block(A);
    quadrant_block(lblock, (A.width + 15) / 16, (A.local_rows - A.lrow0 + PK_TILE_H - 1) / PK_TILE_H, bx, by, A.block_order);
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int lrow = A.lrow0 + by * PK_TILE_H + (warp >> 1) * 4 + (lane >> 3);
    Counters cnt = {0, 0, 0, 0};
    if (px < A.width && lrow < A.local_rows) {
        const size_t pix = (size_t)global_row_of(A, lrow) * A.width + px, lpix = (size_t)lrow * A.width + px;
        const float margin = prune_margin(A.bvh.root_box, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < A.spp; k0 += PK) {
            float dx[PK], dy[PK], dz[PK], tnear[PK];
            int best_leaf[PK];
            const float4* dp = reinterpret_cast<const float4*>(A.dirs + 3 * (pix * A.spp + k0));
            const float4 d0 = __ldcs(dp), d1 = __ldcs(dp + 1), d2 = __ldcs(dp + 2);
            dx[0] = d0.x; dy[0] = d0.y; dz[0] = d0.z; dx[1] = d0.w; dy[1] = d1.x; dz[1] = d1.y;
            dx[2] = d1.z; dy[2] = d1.w; dz[2] = d2.x; dx[3] = d2.y; dy[3] = d2.z; dz[3] = d2.w;
            cnt.rays += PK;
            trace_packet4<true>(A, dx, dy, dz, margin, tnear, best_leaf, cnt);
            *reinterpret_cast<float4*>(A.wave_t + lpix * A.spp + k0) = make_float4(tnear[0], tnear[1], tnear[2], tnear[3]);
            *reinterpret_cast<int4*>(A.wave_leaf + lpix * A.spp + k0) = make_int4(best_leaf[0], best_leaf[1], best_leaf[2], best_leaf[3]);
        }
    }
    report_block_cost(A, lblock, cnt.node_visits + cnt.prim_tests);
    unsigned v[4] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

// one thread per local sample of the launch's rows [lrow0, local_rows): a block = a small tile of pixels (wave_tile); a warp holds 32
// consecutive samples (two pixels at 16 spp) - like rays, like phases. Each thread finishes its sample's castRay from the stored
// hit (shadow rays light after light, shading; material hits through cast_material_cold), then every pixel's samples are added in
// sample order (main.cpp:553-560): warp shuffles, or shared memory when aa_samples is not 4, 8, 16 or 32.
// block size x resident blocks, measured on config 5's stand-in (whole frame, ms): 256x4 50.5, 256x6 48.0, 128x8 48.8, 128x10 46.5,
// 128x12 48.3, 128x16 64.3, 64x16 48.6, 64x20 46.7, 32x32 49.4 - small blocks (a block leaves when its slowest ray is done) and 40
// warps per SM (48 registers, some spilled) win
#ifndef WAVE_THREADS
#define WAVE_THREADS 128
#endif
__device__ __forceinline__ int wave_slot(int i) { return i + (i >> 5); }      // padded: a pixel's leader reads stride-spp without bank conflicts
// the block's pixels: a tw x th tile, tw * th = the largest power of two <= WAVE_THREADS / spp (4 x 2 at 16 spp, 8 x 4 at 4 spp). A
// square tile, not a run of one scanline: the block's rays then share the nodes they visit through L1 (the lesson of the strip kernel)
__host__ __device__ __forceinline__ void wave_tile(int spp, int& tw, int& th)
{
    int lg = 0;
    while ((2 << lg) <= WAVE_THREADS / spp) ++lg;
    tw = 1 << ((lg + 1) / 2);
    th = (1 << lg) / tw;
}
// WARP_SUM (aa_samples = 4, 8, 16 or 32: a pixel's samples sit in one warp): the per-pixel sums go through shuffles, no block barrier -
// a barrier makes every warp wait for the block's slowest shadow ray (ncu, first version: 31 % of the stall cycles).
#ifndef RTDS_WAVE_MINB
#define RTDS_WAVE_MINB 10
#endif
template <bool WARP_SUM>
__global__ void __launch_bounds__(WAVE_THREADS, RTDS_WAVE_MINB) wave_shade_kernel(const __grid_constant__ RenderArgs A)
{
    __shared__ float col[WARP_SUM ? 1 : 3][WARP_SUM ? 1 : WAVE_THREADS + WAVE_THREADS / 32];
    __shared__ int last_obj[WARP_SUM ? 1 : WAVE_THREADS];
    const int lane = threadIdx.x & 31;
    int tw, th;
    wave_tile(A.spp, tw, th);
    const int ppb = tw * th, tiles_x = (A.width + tw - 1) / tw;
    const int tile_y = (int)blockIdx.x / tiles_x, tile_x = (int)blockIdx.x - tile_y * tiles_x;
    const int lp = WARP_SUM ? (int)threadIdx.x >> (__ffs(A.spp) - 1) : (int)threadIdx.x / A.spp, k = (int)threadIdx.x - lp * A.spp;
    const int px = tile_x * tw + (lp & (tw - 1)), lrow = A.lrow0 + tile_y * th + (lp >> (__ffs(tw) - 1));
    const bool active = lp < ppb && px < A.width && lrow < A.local_rows;
    Counters cnt = {0, 0, 0, 0};
    unsigned shadow_rays = 0, secondary_rays = 0;
    float r = 0.f, g = 0.f, b = 0.f;
    int obj = -1;
    if (active) {
        const size_t gl = ((size_t)lrow * A.width + px) * A.spp + k;
        const int leaf = __ldcs(A.wave_leaf + gl);
        if (leaf >= 0) obj = __ldg(A.bvh.prim_order + leaf);
        if (obj < 0) { r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2]; }
        else {
            const float* dp = A.dirs + 3 * (((size_t)global_row_of(A, lrow) * A.width + px) * A.spp + k);
            const float dx = __ldcs(dp), dy = __ldcs(dp + 1), dz = __ldcs(dp + 2);
            const float tnear = __ldcs(A.wave_t + gl);
            const float4 m = __ldg(A.mat + obj);
            if (m.w != 0.0f) {
                const MatResult mr = cast_material_cold(&A, dx, dy, dz, tnear, obj, leaf);
                cnt.node_tests += mr.node_tests; cnt.prim_tests += mr.prim_tests; cnt.node_visits += mr.node_visits; cnt.rays += mr.rays;
                shadow_rays += mr.shadow_rays; secondary_rays += mr.secondary_rays;
                r = mr.r; g = mr.g; b = mr.b;
            } else {
                const float hx = 0.f + dx * tnear, hy = 0.f + dy * tnear, hz = 0.f + dz * tnear;     // main.cpp:396
                float nx, ny, nz;
                raw_normal(0, A.bvh.leaf_sph, nullptr, (size_t)leaf, hx, hy, hz, nx, ny, nz);
                shade_diffuse_shadowed<false>(A, dx, dy, dz, hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b, cnt, shadow_rays);
            }
        }
    }
    float acc_r = 0, acc_g = 0, acc_b = 0;
    int last_hit = -1;
    bool leader;
    if (WARP_SUM) {
        __syncwarp();
        const int base = lane & ~(A.spp - 1);
        for (int j = 0; j < A.spp; ++j) {                                       // sample order, main.cpp:553-560
            acc_r += __shfl_sync(0xffffffffu, r, base + j); acc_g += __shfl_sync(0xffffffffu, g, base + j); acc_b += __shfl_sync(0xffffffffu, b, base + j);
        }
        last_hit = __shfl_sync(0xffffffffu, obj, base + A.spp - 1);
        leader = active && k == 0;
    } else {
        const int sl = wave_slot(threadIdx.x);
        col[0][sl] = r; col[1][sl] = g; col[2][sl] = b;
        last_obj[threadIdx.x] = obj;
        __syncthreads();
        // thread t < ppb now speaks for the tile's pixel t
        const int qx = tile_x * tw + ((int)threadIdx.x & (tw - 1)), qrow = A.lrow0 + tile_y * th + (int)threadIdx.x / tw;
        leader = (int)threadIdx.x < ppb && qx < A.width && qrow < A.local_rows;
        if (leader) {
            const int s0 = (int)threadIdx.x * A.spp;
            for (int j = 0; j < A.spp; ++j) {
                const int sj = wave_slot(s0 + j);
                acc_r += col[0][sj]; acc_g += col[1][sj]; acc_b += col[2][sj];
            }
            last_hit = last_obj[s0 + A.spp - 1];
        }
    }
    if (leader) {
        const int opx = WARP_SUM ? px : tile_x * tw + ((int)threadIdx.x & (tw - 1));
        const int orow = WARP_SUM ? lrow : A.lrow0 + tile_y * th + (int)threadIdx.x / tw;
        const float fs = (float)(unsigned)A.spp;
        const unsigned char r8 = (unsigned char)(fminf(1.0f, acc_r / fs) * 255);
        const unsigned char g8 = (unsigned char)(fminf(1.0f, acc_g / fs) * 255);
        const unsigned char b8 = (unsigned char)(fminf(1.0f, acc_b / fs) * 255);
        const size_t o = (size_t)orow * A.width + opx;
        if (A.out_hit) A.out_hit[o] = last_hit;
        if (A.out_accum) { A.out_accum[3 * o] = acc_r; A.out_accum[3 * o + 1] = acc_g; A.out_accum[3 * o + 2] = acc_b; }
        const size_t oo = (size_t)(A.out_global_rows ? global_row_of(A, orow) : orow) * A.width + opx;
        A.out_rgb[3 * oo] = r8; A.out_rgb[3 * oo + 1] = g8; A.out_rgb[3 * oo + 2] = b8;
    }
    // per-warp totals fit 32 bits here (one sample per thread): one REDUX per counter, nothing for a warp of sky samples
    unsigned v[6] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays, shadow_rays, secondary_rays};
    if (__any_sync(0xffffffffu, (v[0] | v[3] | v[4] | v[5]) != 0)) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const unsigned x = __reduce_add_sync(0xffffffffu, v[c]);
            if (lane == 0 && x) atomicAdd(&A.counters[c], (unsigned long long)x);
        }
    }
}

// ===================================================================================================
// parity probe: arbitrary rays
// ===================================================================================================
struct TraceArgs {
    const float* o; const float* d; int nrays;
    BvhView bvh;
    KdView kd;
    const float4* sph; const float4* tri; int prim_type; int n;
    int* hit; float* t;
    unsigned long long* counters;
    int exact;
};

template <int MODE /*0 BVH, 2 NONE, 3 KDTREE any-hit, 4 KDTREE closest hit*/>
__global__ void __launch_bounds__(128) trace_kernel(const __grid_constant__ TraceArgs A)
{
    __shared__ float4 sh_sph[MODE == 2 ? NONE_CHUNK : 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = i < A.nrays;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = -1;
    if (active) {
        ox = A.o[3 * i]; oy = A.o[3 * i + 1]; oz = A.o[3 * i + 2];
        dx = A.d[3 * i]; dy = A.d[3 * i + 1]; dz = A.d[3 * i + 2];
    }
    Counters cnt = {0, 0, 0, 0};
    float tnear = INFINITY;
    int best_key = 0, best_leaf = -1, hit_obj = -1;
    if (MODE == 2) {
        brute_force_block(A.prim_type, A.sph, A.tri, A.n, active, ox, oy, oz, dx, dy, dz, tnear, hit_obj, cnt, sh_sph);
    } else if (MODE == 3) {
        if (active) { hit_obj = kd_any_hit(A.kd, ox, oy, oz, dx, dy, dz, cnt) ? 1 : -1; tnear = 0.f; }
    } else if (MODE == 4) {
        if (active) kd_closest_hit(A.kd, ox, oy, oz, dx, dy, dz, tnear, hit_obj, cnt);
    } else if (active) {
        // the ordered traversal's pruning bound assumes a unit direction (as every ray castRay makes has)
        float len2 = dx * dx + dy * dy + dz * dz;
        bool unit = fabsf(len2 - 1.0f) < 1e-3f;
        if (A.exact || !unit) traverse_bvh_exact(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
        else traverse_fast<false>(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
        if (best_leaf >= 0) hit_obj = __ldg(A.bvh.prim_order + best_leaf);
    }
    if (active) { A.hit[i] = hit_obj; A.t[i] = tnear; cnt.rays = 1; }
    unsigned v[4] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
    }
}

BvhView make_view(const DeviceBvh& b)
{
    BvhView v;
    v.nodes = b.nodes; v.leaf_sph = b.leaf_sph; v.prim_order = b.prim_order; v.leaf_parent = b.leaf_parent; v.leaf_tri = b.leaf_tri; v.prim_type = b.prim_type;
    v.root_ref = b.root_ref; v.tie_by_objid = b.tie_by_objid; v.leaf_box_prim = b.leaf_box_prim;
    for (int i = 0; i < 6; ++i) v.root_box[i] = b.root_box[i];
    v.wide = (b.wide_valid && b.root_ref >= 0) ? b.wide : nullptr;
    return v;
}

KdView make_kd_view(const rtds_ctx* ctx)
{
    KdView v;
    v.nodes = ctx->kd.nodes; v.prim_idx = ctx->kd.prim_idx; v.sph = ctx->d_sph; v.tri = ctx->d_tris; v.prim_type = ctx->prim_type;
    for (int i = 0; i < 6; ++i) v.bounds[i] = ctx->kd.bounds[i];
    return v;
}

int check_bvh(rtds_ctx* ctx, int acc)
{
    if (!ctx->bvh.valid || ctx->bvh_acc != acc) {
        rtds_set_error("acc_type %d requested but the last BVH/LBVH build was acc_type %d (valid=%d): call rtds_build first",
                       acc, ctx->bvh_acc, (int)ctx->bvh.valid);
        return RTDS_ERR_NOT_BUILT;
    }
    if (ctx->bvh.max_depth >= STACK_MAX) {
        rtds_set_error("tree depth %d exceeds the traversal stack (%d)", ctx->bvh.max_depth, STACK_MAX);
        return RTDS_ERR_UNSUPPORTED;
    }
    return RTDS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// MT19937 state snapshots for chunks [0, need_snaps): computed once per context by one sequential block, extended on demand
static int ensure_snapshots(rtds_ctx* ctx, int need_snaps, int* launches)
{
    cudaStream_t s = ctx->stream;
    if (ctx->n_snap < need_snaps) {
        // grow geometrically; keep existing snapshots
        int cap = need_snaps + need_snaps / 4 + 16;
        uint32_t* nsnap = nullptr;
        RTDS_CUDA(cudaMalloc(&nsnap, sizeof(uint32_t) * MT_N * (size_t)cap));
        if (ctx->n_snap > 0)
            RTDS_CUDA(cudaMemcpyAsync(nsnap, ctx->d_mt_snap, sizeof(uint32_t) * MT_N * (size_t)ctx->n_snap, cudaMemcpyDeviceToDevice, s));
        mt_snapshot_kernel<<<1, MT_THREADS, 0, s>>>(nsnap, ctx->n_snap, cap, 5489u);
        if (launches) *launches += 1;
        RTDS_CUDA(cudaGetLastError());
        RTDS_CUDA(cudaStreamSynchronize(s));
        if (ctx->d_mt_snap) cudaFree(ctx->d_mt_snap);
        ctx->d_mt_snap = nsnap;
        ctx->n_snap = cap;
    }
    return RTDS_OK;
}

static inline unsigned __float_as_uint_host(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static int jitter_prepare(rtds_ctx* ctx, uint64_t first_word, size_t n_words, int* launches, const JitterOwner& own)
{
    if (n_words == 0) return RTDS_OK;
    const uint64_t words_per_snap = (uint64_t)MT_SNAP_EVERY * MT_N;
    const uint64_t s0 = first_word / words_per_snap;
    const uint64_t s1 = (first_word + n_words - 1) / words_per_snap;  // last snapshot chunk needed (inclusive)
    if (s1 + 2 >= (1ull << 31)) { rtds_set_error("jitter stream position too large"); return RTDS_ERR_INVALID; }
    RTDS_TRY(ensure_snapshots(ctx, (int)(s1 + 1), launches));
    cudaStream_t s = ctx->stream;
    const int blocks = (int)(s1 - s0 + 1);
    const size_t out_words = (size_t)blocks * words_per_snap;
    if (ctx->jitter_cap_words < out_words) {
        if (ctx->d_jitter) cudaFree(ctx->d_jitter);
        ctx->d_jitter = nullptr; ctx->jitter_cap_words = 0;
        RTDS_CUDA(cudaMalloc(&ctx->d_jitter, sizeof(uint32_t) * out_words));
        ctx->jitter_cap_words = out_words;
    }
    mt_expand_kernel<<<blocks, MT_THREADS, 0, s>>>(ctx->d_mt_snap, (int)s0, ctx->d_jitter, out_words, own);
    if (launches) *launches += 1;
    RTDS_CUDA(cudaGetLastError());
    ctx->jitter_first_word = s0 * words_per_snap;
    ctx->jitter_n_words = out_words;
    return RTDS_OK;
}

// Primary directions of the frame's samples (pixel * spp + k), generated from the jitter stream starting at sample
// `first_sample`; only the chunks that hold rows of `own` are filled. dirs_plan() sizes the buffers and fills the launch
// (grid + arguments of mt_expand_dirs_kernel), dirs_prepare() also issues it on `stream`.
struct DirsLaunch {
    unsigned grid = 0;
    const uint32_t* snap; int s0; float* dirs; unsigned long long first_sample; unsigned n_samples; int width, spp;
    FastDiv div_spp, div_width; RayGen G; JitterOwner own; int chunks_per_tile;
    void* args[12];
    void bind() { void* a[12] = {&snap, &s0, &dirs, &first_sample, &n_samples, &width, &spp, &div_spp, &div_width, &G, &own, &chunks_per_tile};
                  for (int i = 0; i < 12; ++i) args[i] = a[i]; }
};

static int dirs_plan(rtds_ctx* ctx, uint64_t first_sample, int W, int H, int spp, const RayGen& G, const JitterOwner& own, int* launches,
                     DirsLaunch& L)
{
    float*& buf = ctx->d_dirs;
    size_t& cap = ctx->dirs_cap_floats;
    uint64_t* key = ctx->dirs_key;
    bool& valid = ctx->dirs_valid;
    const uint64_t n_samples = (uint64_t)W * H * spp;
    if (n_samples >= (1ull << 32)) { rtds_set_error("render: width*height*aa_samples must be below 2^32"); return RTDS_ERR_INVALID; }
    const uint64_t words_per_snap = (uint64_t)MT_SNAP_EVERY * MT_N;
    const uint64_t first_word = 4 * first_sample, n_words = 4 * n_samples;
    const uint64_t s0 = first_word / words_per_snap, s1 = (first_word + n_words - 1) / words_per_snap;
    if (s1 + 2 >= (1ull << 31)) { rtds_set_error("jitter stream position too large"); return RTDS_ERR_INVALID; }
    RTDS_TRY(ensure_snapshots(ctx, (int)(s1 + 1), launches));
    if (cap < 3 * n_samples) {
        if (buf) cudaFree(buf);
        buf = nullptr; cap = 0;
        RTDS_CUDA(cudaMalloc(&buf, sizeof(float) * 3 * n_samples));
        cap = 3 * n_samples;
    }
    auto make_div = [](uint32_t d) {
        FastDiv f{0u, 0u};
        uint32_t fl = 0;
        while ((2ull << fl) <= d) ++fl;                         // floor(log2 d)
        if ((d & (d - 1)) == 0) { f.shift = fl; return f; }
        const uint64_t num = 1ull << (32 + fl);
        uint64_t pm = num / d;
        const uint64_t rem = num % d;
        pm += pm;
        if (rem + rem >= d) pm += 1;
        f.magic = (uint32_t)(1 + pm);
        f.shift = fl;
        return f;
    };
    unsigned grid = (unsigned)(s1 - s0 + 1);
    int chunks_per_tile = 0;
    if (own.world > 1) {
        // only the chunks that hold rows of this rank's tiles: (#owned tiles) x (max chunks a tile can span)
        const uint64_t tile_words = (uint64_t)own.tile_rows * own.row_words;
        chunks_per_tile = (int)((tile_words + words_per_snap - 1) / words_per_snap + 1);
        const int n_tiles = (H + own.tile_rows - 1) / own.tile_rows;
        const int owned = own.rank < n_tiles ? (n_tiles - own.rank + own.world - 1) / own.world : 0;
        grid = (unsigned)owned * (unsigned)chunks_per_tile;
    }
    L.grid = grid; L.snap = ctx->d_mt_snap; L.s0 = (int)s0; L.dirs = buf; L.first_sample = first_sample; L.n_samples = (unsigned)n_samples;
    L.width = W; L.spp = spp; L.div_spp = make_div((uint32_t)spp); L.div_width = make_div((uint32_t)W); L.G = G; L.own = own;
    L.chunks_per_tile = chunks_per_tile;
    L.bind();
    key[0] = first_sample; key[1] = (uint64_t)W; key[2] = (uint64_t)H; key[3] = (uint64_t)spp;
    key[4] = ((uint64_t)__float_as_uint_host(G.angle) << 32) | __float_as_uint_host(G.aspect);
    key[5] = ((uint64_t)(unsigned)own.rank << 40) | ((uint64_t)(unsigned)own.world << 20) | (uint64_t)(unsigned)own.tile_rows;
    valid = true;
    return RTDS_OK;
}

static int dirs_prepare(rtds_ctx* ctx, uint64_t first_sample, int W, int H, int spp, const RayGen& G, const JitterOwner& own, int* launches,
                        cudaStream_t stream)
{
    DirsLaunch L;
    RTDS_TRY(dirs_plan(ctx, first_sample, W, H, spp, G, own, launches, L));
    if (L.grid > 0) RTDS_CUDA(cudaLaunchKernel((const void*)mt_expand_dirs_kernel, dim3(L.grid), dim3(MT_THREADS), L.args, 0, stream));
    if (launches) *launches += 1;
    RTDS_CUDA(cudaGetLastError());
    return RTDS_OK;
}

int rtds_jitter_prepare(rtds_ctx* ctx, uint64_t first_word, size_t n_words, int* launches)
{
    JitterOwner all{0, 1, 1, 0, 1, 1};
    return jitter_prepare(ctx, first_word, n_words, launches, all);
}

int rtds_jitter_stream_impl(rtds_ctx* ctx, uint64_t first, int n, double* out)
{
    if (n <= 0) return RTDS_OK;
    RTDS_TRY(rtds_jitter_prepare(ctx, first * 2, (size_t)n * 2, nullptr));
    RTDS_TRY(rtds_ensure_scratch(ctx, sizeof(double) * (size_t)n));
    double* d_out = (double*)ctx->d_scratch;
    jitter_doubles_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_jitter, (size_t)(first - ctx->jitter_first_word / 2), n, d_out);
    RTDS_CUDA(cudaGetLastError());
    RTDS_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
    return RTDS_OK;
}

static RayGen make_raygen(const rtds_render_params* p)
{
    // main.cpp:544-546
    const float fov = p->fov > 0 ? p->fov : 30.0f;
    RayGen G;
    G.inv_w = 1 / float(p->width); G.inv_h = 1 / float(p->height);
    G.aspect = p->width / float(p->height);
    G.angle = (float)tan(3.141592653589793 * 0.5 * fov / 180.);
    return G;
}

int rtds_ensure_band_streams(rtds_ctx* ctx)
{
    if (ctx->band_streams[0]) return RTDS_OK;
    int least = 0, greatest = 0;
    RTDS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));     // numerically: greatest <= least
    for (int k = 0; k < RTDS_MAX_BANDS; ++k)
        RTDS_CUDA(cudaStreamCreateWithPriority(&ctx->band_streams[k], cudaStreamNonBlocking, std::min(least, greatest + k)));
    RTDS_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    // timing only when RTDS_TRACE_FRAME reads them (measured: timing-enabled band events cost the frame ~0.1 ms)
    const bool tr = ctx->opt.trace_frame > 1;
    for (int k = 0; k < RTDS_MAX_BANDS; ++k) RTDS_CUDA(cudaEventCreateWithFlags(&ctx->ev_bands[k], tr ? cudaEventDefault : cudaEventDisableTiming));
    return RTDS_OK;
}

int rtds_prefetch_dirs(rtds_ctx* ctx, const rtds_render_params* p)
{
    const int W = p->width, H = p->height, spp = p->aa_samples;
    if (W <= 0 || H <= 0 || spp <= 0) return RTDS_OK;        // the render call reports the error
    const int world = p->world > 0 ? p->world : 1, rank = p->rank, tile_rows = p->tile_rows > 0 ? p->tile_rows : 8;
    if (rank < 0 || rank >= world) return RTDS_OK;
    if (ctx->opt.strip == 1) return RTDS_OK;
    if (ctx->dirs_pending) RTDS_CUDA(cudaStreamSynchronize(ctx->jit_stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));            // a render still reading d_dirs
    const RayGen G = make_raygen(p);
    JitterOwner own{4ull * p->jitter_offset, 4ull * (unsigned long long)W * spp, tile_rows, rank, world, H};
    int launches = 0;
    RTDS_TRACE_RECORD(ctx, 1, ctx->jit_stream);
    RTDS_TRY(dirs_prepare(ctx, p->jitter_offset, W, H, spp, G, own, &launches, ctx->jit_stream));
    RTDS_TRACE_RECORD(ctx, 2, ctx->jit_stream);
    RTDS_CUDA(cudaEventRecord(ctx->ev_dirs, ctx->jit_stream));
    ctx->dirs_pending = true;
    ctx->dirs_pending_launches = launches;
    return RTDS_OK;
}

// completion flags of the multi-GPU shared frame (one 128-byte line per rank behind the frame)
__global__ void frame_signal_kernel(volatile uint32_t* flag, uint32_t seq)
{
    __threadfence_system();        // this rank's tile stores (previous kernel on the stream) before the flag
    *flag = seq;
}
__global__ void frame_wait_kernel(const volatile uint32_t* flags, int world, uint32_t seq, long long timeout_cycles, int* status)
{
    const int r = threadIdx.x;
    if (r < world) {
        const long long t0 = clock64();
        while ((int)(flags[32 * r] - seq) < 0) {     // a rank may already have signalled a later frame
            if (clock64() - t0 > timeout_cycles) { atomicExch(status, 1 + r); break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// Streams the tree (interior nodes + leaf spheres) into L2 with evict-last priority while the direction kernel runs: the
// render kernel's first wave then finds the upper levels in L2 instead of every block missing on them at once, and the
// 12 B/ray direction stream that follows does not push the tree out again. Only issued when the tree fits L2.
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const char* __restrict__ a, size_t na, const char* __restrict__ b, size_t nb)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128, stride = (size_t)gridDim.x * blockDim.x * 128;
    for (size_t o = i0; o < na; o += stride) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a + o));
    for (size_t o = i0; o < nb; o += stride) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(b + o));
}
constexpr size_t L2_PREFETCH_MAX_BYTES = 100u << 20;
constexpr unsigned L2_PREFETCH_BLOCKS = 148 * 2;

struct PrefetchArgs { const char* a; size_t na; const char* b; size_t nb; };
static bool prefetch_plan(const rtds_ctx* ctx, bool bvh_path, PrefetchArgs& P)
{
    if (!ctx->opt.l2_prefetch || !bvh_path || ctx->bvh.prim_type != 0 || ctx->bvh.n_internal <= 0) return false;
    P.a = (const char*)ctx->bvh.nodes; P.na = sizeof(Node64) * (size_t)ctx->bvh.n_internal;
    P.b = (const char*)ctx->bvh.leaf_sph; P.nb = sizeof(float4) * (size_t)ctx->bvh.n_prims;
    return P.na + P.nb <= L2_PREFETCH_MAX_BYTES;
}

static bool same_bytes(std::vector<unsigned char>& last, const void* now, size_t n)
{
    if (last.size() == n && memcmp(last.data(), now, n) == 0) return true;
    last.assign((const unsigned char*)now, (const unsigned char*)now + n);
    return false;
}

// frame_graph option: the whole frame as one graph launch on ctx->stream (see FrameGraph). The caller synchronises.
static unsigned render_threads(const void* fn);
static int render_frame_graph(rtds_ctx* ctx, const RenderArgs& A, const void* fn, unsigned lin, const DirsLaunch* DL, const PrefetchArgs* PF,
                              bool shared, uint32_t seq, int* launches)
{
    FrameGraph& g = ctx->fg;
    SharedFrame& f = ctx->shared;
    const bool has_dirs = DL && DL->grid > 0, has_pf = PF != nullptr, has_signal = shared, has_wait = shared && f.owner;
    // kernel parameter blocks
    void* a_render[] = {(void*)&A};
    PrefetchArgs pf = PF ? *PF : PrefetchArgs{nullptr, 0, nullptr, 0};
    void* a_pf[] = {&pf.a, &pf.na, &pf.b, &pf.nb};
    volatile uint32_t* sig_flag = shared ? f.flags + 32 * f.rank : nullptr;
    void* a_signal[] = {&sig_flag, &seq};
    const volatile uint32_t* wait_flags = f.flags;
    int wait_world = f.world;
    long long wait_timeout = 20000000000ll;      // ~10 s
    int* wait_status = (int*)(ctx->d_counters + 7);
    void* a_wait[] = {&wait_flags, &wait_world, &seq, &wait_timeout, &wait_status};
    auto kparams = [](const void* func, unsigned grid, unsigned block, void** args) {
        cudaKernelNodeParams kp = {};
        kp.func = (void*)func; kp.gridDim = dim3(grid); kp.blockDim = dim3(block); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        return kp;
    };
    const bool same_shape = g.exec && g.fn_render == fn && g.has_dirs == has_dirs && g.has_pf == has_pf && g.has_signal == has_signal &&
                            g.has_wait == has_wait;
    if (!same_shape) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (g.graph) { cudaGraphDestroy(g.graph); g.graph = nullptr; }
        RTDS_CUDA(cudaGraphCreate(&g.graph, 0));
        cudaGraphNode_t n_ev0, n_zero, n_ev2, n_ev3, n_ev1, n_copy;
        RTDS_CUDA(cudaGraphAddEventRecordNode(&n_ev0, g.graph, nullptr, 0, ctx->ev0));
        cudaMemsetParams mp = {};
        mp.dst = ctx->d_counters; mp.value = 0; mp.elementSize = 4; mp.width = 16; mp.height = 1; mp.pitch = 0;
        RTDS_CUDA(cudaGraphAddMemsetNode(&n_zero, g.graph, &n_ev0, 1, &mp));
        std::vector<cudaGraphNode_t> pre{n_zero};
        if (has_dirs) {
            cudaKernelNodeParams kp = kparams((const void*)mt_expand_dirs_kernel, DL->grid, MT_THREADS, const_cast<void**>(DL->args));
            RTDS_CUDA(cudaGraphAddKernelNode(&g.k_dirs, g.graph, &n_ev0, 1, &kp));
            pre.push_back(g.k_dirs);
        }
        if (has_pf) {
            cudaKernelNodeParams kp = kparams((const void*)l2_prefetch_kernel, L2_PREFETCH_BLOCKS, 256, a_pf);
            RTDS_CUDA(cudaGraphAddKernelNode(&g.k_pf, g.graph, &n_ev0, 1, &kp));
            pre.push_back(g.k_pf);
        }
        RTDS_CUDA(cudaGraphAddEventRecordNode(&n_ev2, g.graph, pre.data(), pre.size(), ctx->ev2));
        {
            cudaKernelNodeParams kp = kparams(fn, lin, render_threads(fn), a_render);
            RTDS_CUDA(cudaGraphAddKernelNode(&g.k_render, g.graph, &n_ev2, 1, &kp));
        }
        RTDS_CUDA(cudaGraphAddEventRecordNode(&n_ev3, g.graph, &g.k_render, 1, ctx->ev3));
        cudaGraphNode_t last = n_ev3;
        if (has_signal) {
            cudaKernelNodeParams kp = kparams((const void*)frame_signal_kernel, 1, 1, a_signal);
            RTDS_CUDA(cudaGraphAddKernelNode(&g.k_signal, g.graph, &last, 1, &kp));
            last = g.k_signal;
        }
        if (has_wait) {
            cudaKernelNodeParams kp = kparams((const void*)frame_wait_kernel, 1, 32, a_wait);
            RTDS_CUDA(cudaGraphAddKernelNode(&g.k_wait, g.graph, &last, 1, &kp));
            last = g.k_wait;
        }
        RTDS_CUDA(cudaGraphAddEventRecordNode(&n_ev1, g.graph, &last, 1, ctx->ev1));
        RTDS_CUDA(cudaGraphAddMemcpyNode1D(&n_copy, g.graph, &n_ev1, 1, ctx->h_counters, ctx->d_counters, sizeof(unsigned long long) * 8,
                                           cudaMemcpyDeviceToHost));
        RTDS_CUDA(cudaGraphInstantiate(&g.exec, g.graph, 0));
        g.fn_render = fn; g.has_dirs = has_dirs; g.has_pf = has_pf; g.has_signal = has_signal; g.has_wait = has_wait;
        g.grid_dirs = has_dirs ? DL->grid : 0; g.grid_render = lin;
        g.last_render.assign((const unsigned char*)&A, (const unsigned char*)&A + sizeof A);
        if (has_dirs) g.last_dirs.assign((const unsigned char*)DL, (const unsigned char*)DL + offsetof(DirsLaunch, args));
        if (has_pf) g.last_pf.assign((const unsigned char*)&pf, (const unsigned char*)&pf + sizeof pf);
        ++g.rebuilds;
    } else {
        if (!same_bytes(g.last_render, &A, sizeof A) || g.grid_render != lin) {
            cudaKernelNodeParams kp = kparams(fn, lin, render_threads(fn), a_render);
            RTDS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.k_render, &kp));
            g.grid_render = lin;
        }
        if (has_dirs && (!same_bytes(g.last_dirs, DL, offsetof(DirsLaunch, args)) || g.grid_dirs != DL->grid)) {
            cudaKernelNodeParams kp = kparams((const void*)mt_expand_dirs_kernel, DL->grid, MT_THREADS, const_cast<void**>(DL->args));
            RTDS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.k_dirs, &kp));
            g.grid_dirs = DL->grid;
        }
        if (has_pf && !same_bytes(g.last_pf, &pf, sizeof pf)) {
            cudaKernelNodeParams kp = kparams((const void*)l2_prefetch_kernel, L2_PREFETCH_BLOCKS, 256, a_pf);
            RTDS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.k_pf, &kp));
        }
        if (has_signal) {      // the frame sequence number changes every frame
            cudaKernelNodeParams kp = kparams((const void*)frame_signal_kernel, 1, 1, a_signal);
            RTDS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.k_signal, &kp));
        }
        if (has_wait) {
            cudaKernelNodeParams kp = kparams((const void*)frame_wait_kernel, 1, 32, a_wait);
            RTDS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.k_wait, &kp));
        }
    }
    RTDS_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
    ++g.launches;
    *launches += 1 + (has_dirs ? 1 : 0) + (has_pf ? 1 : 0) + (has_signal ? 1 : 0) + (has_wait ? 1 : 0);
    return RTDS_OK;
}

static void lpt_frame_timed(rtds_ctx* ctx, float ms_kernel, bool was_lpt_frame, int mode_run);

// wavefront = 1: the two forms of a frame with shadow rays give the same bytes; which is faster depends on the scene and the frame
// (config 5 stand-in, 16 samples x 3 lights: wavefront 53.7 ms, single kernel 104.6; bunny x 30 at 4 samples, 1 light: 2.67 vs 1.96).
// The first WAVE_TRIAL_FRAMES timed frames of a frame geometry alternate (single kernel, wavefront, ...), the best time of each
// decides. Frames nobody times (no statistics requested) and the time before the decision use the rule samples x lights >= 16.
constexpr int WAVE_TRIAL_FRAMES = 4;
static bool wave_choose(rtds_ctx* ctx, const RenderArgs& A, int n_lights, bool timed)
{
    ctx->wave_trial = false;
    if (ctx->opt.wavefront >= 2) return true;
    const uint64_t key[4] = {((uint64_t)(unsigned)A.width << 32) | (unsigned)A.height, ((uint64_t)(unsigned)A.spp << 32) | (unsigned)n_lights,
                             ((uint64_t)(unsigned)A.rank << 40) | ((uint64_t)(unsigned)A.world << 20) | (uint64_t)(unsigned)A.tile_rows,
                             (uint64_t)A.n};
    if (memcmp(key, ctx->wave_key, sizeof key) != 0) {
        memcpy(ctx->wave_key, key, sizeof key);
        ctx->wave_phase = 0; ctx->wave_ms[0] = ctx->wave_ms[1] = 0.f;
        ctx->wave_use = A.spp * n_lights >= 16;
    }
    if (timed && ctx->wave_phase < WAVE_TRIAL_FRAMES) { ctx->wave_trial = true; return (ctx->wave_phase & 1) != 0; }
    return ctx->wave_use;
}
static void wave_frame_timed(rtds_ctx* ctx, float ms_kernel)
{
    if (!ctx->wave_trial || !(ms_kernel > 0.f)) return;
    ctx->wave_trial = false;
    float& best = ctx->wave_ms[ctx->wave_this_frame ? 1 : 0];
    best = best > 0.f ? std::min(best, ms_kernel) : ms_kernel;
    if (++ctx->wave_phase == WAVE_TRIAL_FRAMES) ctx->wave_use = ctx->wave_ms[1] < ctx->wave_ms[0];
}

// render statistics from the counters the frame left in pinned memory (the stream is already synchronised)
static int render_stats_out(rtds_ctx* ctx, rtds_render_stats* st, int launches, int rows, bool wait_copies, bool)
{
    const unsigned long long* c = ctx->h_counters;
    memset(st, 0, sizeof *st);
    st->node_tests = c[0]; st->prim_tests = c[1]; st->node_visits = c[2];
    st->rays = c[3]; st->shadow_rays = c[4]; st->secondary_rays = c[5];
    st->primary_rays = c[3] - c[4] - c[5];
    RTDS_CUDA(cudaEventElapsedTime(&st->ms_kernel, ctx->ev2, ctx->ev3));
    RTDS_CUDA(cudaEventElapsedTime(&st->ms_total, ctx->ev0, ctx->ev1));
    st->kernel_launches = launches;
    st->rows = rows;
    st->reserved[0] = (int)(unsigned)c[7];      // shared frame: 1 + the rank the owner timed out on (0 = complete)
    wave_frame_timed(ctx, st->ms_kernel);
    lpt_frame_timed(ctx, st->ms_kernel, ctx->lpt_active, ctx->lpt_last_mode);
    ctx->lpt_active = false;
    if (wait_copies) RTDS_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return RTDS_OK;
}

// grid of a render kernel over `rows` local rows: one block per 16 x 8 pixels, quadrant-major linear order
static bool is_packet_fn(const void* fn)
{
    return fn == (const void*)render_packet_kernel<false, false> || fn == (const void*)render_packet_kernel<false, true> ||
           fn == (const void*)render_packet_kernel<true, false> || fn == (const void*)render_packet_kernel<true, true> ||
           fn == (const void*)render_packet_kernel<false, true, true> || fn == (const void*)wave_primary_kernel;
}
static unsigned render_threads(const void* fn) { return is_packet_fn(fn) ? RTDS_PK_THREADS : 128; }
static unsigned render_grid(rtds_ctx*, const void* fn, int W, int rows)
{
    const int th = is_packet_fn(fn) ? PK_TILE_H : 8;
    return (unsigned)((W + 15) / 16) * (unsigned)((rows + th - 1) / th);
}

// Which render kernel serves this frame (all take one RenderArgs; grid = quadrant-major linear grid of 16 x 8-pixel blocks).
static const void* select_render_kernel(const rtds_ctx* ctx, const RenderArgs& A, const rtds_render_params* p, bool full, bool kdt, bool brute,
                                        bool kd_closest)
{
    // four samples of a pixel per thread as one packet (see traverse_packet); the packet option turns it off
    // (shadows without reflective / refractive materials stay on the packet kernel; the shadow rays are single)
    // Only the samples of ONE pixel are packed. Packets of neighbouring pixels were built twice and measured slower both
    // times, also with the hull test (round 2, B200): 2 x 2 pixels x 1 sample 0.205 vs 0.087 ms (bunny 640x480), 0.25 vs
    // 0.20 ms (1080p), 4.7 vs 2.3 ms (7 M spheres); 2 pixels x 4 samples 1.07 vs 0.97 ms on the bench frame - node visits per
    // ray drop 2-3.4x, but a thread then carries 4-8 rays' leaf work in sequence and the grid has 2-4x fewer threads
    // to hide latency with (negative results). Removed again.
    bool packet = !kdt && !brute && !p->exact && A.spp % PK == 0 && A.bvh.leaf_box_prim && A.shade.max_depth >= 1;
    packet = packet && ctx->opt.packet != 0;
    // scenes with REFLECTION_AND_REFRACTION / REFLECTION primitives, no shadow rays: the primary rays still go as packets, a ray
    // that hits such a primitive continues through castRay's material branches on its own. WITH shadow rays and materials the
    // single-kernel form is the one-ray-at-a-time castRay kernel (config 5: packets inside ONE kernel 115.9 vs 109.7 ms - the shadow
    // rays are 2/3 of the rays and stay single either way); what such frames normally run is the wavefront form (K10c, chosen by
    // the caller before this function is asked: 47.6 ms)
    if (packet && ctx->has_materials && !p->shadows) return (const void*)render_packet_kernel<false, true, true>;
    if (ctx->has_materials) packet = false;
    // interior boxes tested once per packet against the hull of the four reciprocal directions (default; hull option 0:
    // once per ray). Bench frame: same frame, node visits 3.17 -> 3.18 per ray, kernel 1.16 -> 0.99 ms.
    const bool hull = ctx->opt.hull != 0;
    if (packet && full) return hull ? (const void*)render_packet_kernel<true, true> : (const void*)render_packet_kernel<true, false>;
    if (full) {
        if (brute) return (const void*)render_full_kernel<2>;
        return p->exact ? (const void*)render_full_kernel<0> : (const void*)render_full_kernel<1>;
    }
    if (packet) return hull ? (const void*)render_packet_kernel<false, true> : (const void*)render_packet_kernel<false, false>;
    if (kd_closest) return (const void*)render_kernel<4>;
    if (kdt) return (const void*)render_kernel<3>;
    if (brute) return (const void*)render_kernel<2>;
    return p->exact ? (const void*)render_kernel<0> : (const void*)render_kernel<1>;
}

// lpt option, before a frame's launches: size the cost / order arrays for the frame's bands (one launch each) and decide whether
// the order learned from the previous frame applies (same kernel, geometry and banding); `s` is ordered behind the order kernel.
static int lpt_frame_begin(rtds_ctx* ctx, const RenderArgs& A, const void* fn, const LptBands& bands, cudaStream_t s)
{
    ctx->lpt_active = false;
    // lpt_split: single-band frames of the plain packet kernel only (no shadow rays, no materials)
    ctx->lpt_split_ok = ctx->opt.lpt_split > 0 && bands.n_bands == 1 &&
                        (fn == (const void*)render_packet_kernel<false, true> || fn == (const void*)render_packet_kernel<false, false>);
    if (!ctx->opt.lpt) return RTDS_OK;
    const int total = bands.off[bands.n_bands];
    if (total <= 0) return RTDS_OK;
    if (ctx->block_cap < total) {
        RTDS_CUDA(cudaStreamSynchronize(ctx->jit_stream));
        if (ctx->d_block_cost) cudaFree(ctx->d_block_cost);
        if (ctx->d_block_order) cudaFree(ctx->d_block_order);
        ctx->d_block_cost = nullptr; ctx->d_block_order = nullptr; ctx->block_cap = 0; ctx->block_order_valid = false;
        RTDS_CUDA(cudaMalloc(&ctx->d_block_cost, sizeof(unsigned) * (size_t)total));
        RTDS_CUDA(cudaMalloc(&ctx->d_block_order, sizeof(int) * (size_t)total));
        if (!ctx->d_heavy_list) RTDS_CUDA(cudaMalloc(&ctx->d_heavy_list, sizeof(int) * LPT_SPLIT_MAX));
        if (ctx->d_block_skip) cudaFree(ctx->d_block_skip);
        ctx->d_block_skip = nullptr;
        RTDS_CUDA(cudaMalloc(&ctx->d_block_skip, (size_t)total));
        ctx->block_cap = total;
        memset(ctx->block_key, 0, sizeof ctx->block_key);
    }
    uint64_t bh = (uint64_t)bands.n_bands;
    for (int i = 0; i <= bands.n_bands; ++i) bh = bh * 1000003ull + (uint64_t)bands.off[i];
    const uint64_t key[5] = {(uint64_t)(uintptr_t)fn, ((uint64_t)(unsigned)A.width << 32) | (unsigned)A.local_rows,
                             ((uint64_t)(unsigned)A.rank << 40) | ((uint64_t)(unsigned)A.world << 20) | (uint64_t)(unsigned)A.tile_rows,
                             ((uint64_t)(unsigned)A.block_order << 32) | (unsigned)A.spp, bh};
    RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_order_done, 0));      // the last frame's order kernel (it also clears the costs)
    if (memcmp(key, ctx->block_key, sizeof key) != 0) {
        // another geometry, banding or kernel: forget the old costs and order
        RTDS_CUDA(cudaMemsetAsync(ctx->d_block_cost, 0, sizeof(unsigned) * (size_t)total, s));
        RTDS_CUDA(cudaMemsetAsync(ctx->d_block_skip, 0, (size_t)total, s));
        RTDS_CUDA(cudaMemsetAsync(ctx->d_heavy_list, 0xff, sizeof(int) * LPT_SPLIT_MAX, s));
        memcpy(ctx->block_key, key, sizeof key);
        ctx->block_order_valid = false;
        ctx->lpt_phase = 0; ctx->lpt_choice = 1; ctx->lpt_ms[0] = ctx->lpt_ms[1] = ctx->lpt_ms[2] = 0.f;
    }
    // lpt = 1: the library times both orders once per geometry (its own kernel events) and keeps the faster one; a geometry for
    // which the learned order lost stops recording costs
    ctx->lpt_active = ctx->opt.lpt >= 2 || ctx->lpt_phase < LPT_TRIAL_FRAMES || ctx->lpt_choice != 0;
    return RTDS_OK;
}

// ... with the frame's kernel time known: baseline frame (launch order) -> trial frame (learned order) -> decision
// lpt = 1: which of the three schedules is fastest depends on how many waves of blocks a launch has (one GPU: launch order; 4 GPUs:
// the learned order; 8 GPUs: the learned order with the heaviest tiles ray-per-thread - measured on real boxes), so the library
// times them: the first LPT_TRIAL_FRAMES timed frames of a geometry cycle through the modes, the best time of each decides.
static int lpt_trial_mode(const rtds_ctx* ctx) { return ctx->lpt_phase % (ctx->lpt_split_ok ? 3 : 2); }
static void lpt_frame_timed(rtds_ctx* ctx, float ms_kernel, bool was_lpt_frame, int mode_run)
{
    if (!was_lpt_frame || ctx->opt.lpt != 1 || ctx->lpt_phase >= LPT_TRIAL_FRAMES || !(ms_kernel > 0.f)) return;
    const int mode = lpt_trial_mode(ctx);
    if (mode != mode_run) return;                          // (no order yet, or not a frame the split applies to: counts as nothing...
    float& best = ctx->lpt_ms[mode];
    best = best > 0.f ? std::min(best, ms_kernel) : ms_kernel;
    if (++ctx->lpt_phase == LPT_TRIAL_FRAMES) {
        ctx->lpt_choice = 0;
        float t = ctx->lpt_ms[0];
        for (int m = 1; m < 3; ++m)
            if (ctx->lpt_ms[m] > 0.f && ctx->lpt_ms[m] < 0.997f * t) { t = ctx->lpt_ms[m]; ctx->lpt_choice = m; }
    }
}

// ... per launch: the band's slice of the two arrays
static void lpt_band_args(rtds_ctx* ctx, RenderArgs& A, const LptBands& bands, int band)
{
    A.order = nullptr; A.cost = nullptr; A.skip = nullptr; A.heavy_list = nullptr;
    ctx->lpt_last_mode = 0;
    if (!ctx->lpt_active) return;
    const int mode = ctx->opt.lpt >= 2 ? (ctx->lpt_split_ok ? 2 : 1) : (ctx->lpt_phase < LPT_TRIAL_FRAMES ? lpt_trial_mode(ctx) : ctx->lpt_choice);
    if (ctx->block_order_valid && mode >= 1) A.order = ctx->d_block_order + bands.off[band];
    if (A.order && mode == 2 && ctx->lpt_split_ok) { A.skip = ctx->d_block_skip; A.heavy_list = ctx->d_heavy_list; }
    ctx->lpt_last_mode = A.order ? (A.skip ? 2 : 1) : 0;
    A.cost = ctx->d_block_cost + bands.off[band];
}
static int lpt_frame_end(rtds_ctx* ctx, const LptBands& bands, cudaStream_t s)
{
    if (!ctx->lpt_active) return RTDS_OK;
    RTDS_CUDA(cudaEventRecord(ctx->ev_order_go, s));
    RTDS_CUDA(cudaStreamWaitEvent(ctx->jit_stream, ctx->ev_order_go, 0));
    block_order_kernel<<<bands.n_bands, 1024, 0, ctx->jit_stream>>>(ctx->d_block_cost, ctx->d_block_order, bands,
                                                                    std::max(32, ctx->sm_count * std::max(1, ctx->opt.lpt_cap) / bands.n_bands), ctx->d_heavy_list, ctx->d_block_skip,
                                                                    bands.n_bands == 1 ? std::max(0, std::min(LPT_SPLIT_MAX, ctx->opt.lpt_split)) : 0,
                                                                    std::max(1, std::min(31, ctx->opt.lpt_bin)));
    RTDS_CUDA(cudaGetLastError());
    RTDS_CUDA(cudaEventRecord(ctx->ev_order_done, ctx->jit_stream));
    ctx->block_order_valid = true;
    return RTDS_OK;
}

int rtds_render_impl(rtds_ctx* ctx, int acc, const rtds_render_params* p, uint8_t* d_rgb_rows, int* d_hit, float* d_accum,
                     rtds_render_stats* st, const std::function<int(int, int, cudaEvent_t)>* on_band, bool global_rows)
{
    const int W = p->width, H = p->height, spp = p->aa_samples;
    if (W <= 0 || H <= 0 || spp <= 0) { rtds_set_error("render: width/height/aa_samples must be positive"); return RTDS_ERR_INVALID; }
    const int world = p->world > 0 ? p->world : 1, rank = p->rank;
    if (rank < 0 || rank >= world) { rtds_set_error("render: rank %d outside world %d", rank, world); return RTDS_ERR_INVALID; }
    const int tile_rows = p->tile_rows > 0 ? p->tile_rows : 8;
    if (ctx->n <= 0) { rtds_set_error("render: no scene"); return RTDS_ERR_NO_SCENE; }
    RTDS_TRY(rtds_finish_materials(ctx));
    const bool kdt = acc == RTDS_KDTREE;
    if (kdt && !ctx->kd.valid) { rtds_set_error("render: KDTREE requested but no KD-tree has been built"); return RTDS_ERR_NOT_BUILT; }
    const bool brute = (acc != RTDS_BVH && acc != RTDS_LBVH && !kdt);
    if (!brute && !kdt) RTDS_TRY(check_bvh(ctx, acc));

    RenderArgs A;
    A.width = W; A.height = H; A.spp = spp;
    A.wave_t = nullptr; A.wave_leaf = nullptr;
    A.rank = rank; A.world = world; A.tile_rows = tile_rows; A.lrow0 = 0;
    A.local_rows = rtds_rows_for_rank(H, tile_rows, rank, world);
    { const RayGen G0 = make_raygen(p); A.inv_w = G0.inv_w; A.inv_h = G0.inv_h; A.aspect = G0.aspect; A.angle = G0.angle; }
    A.n = ctx->n; A.sph = ctx->d_sph; A.tri = ctx->d_tris; A.prim_type = ctx->prim_type; A.mat = ctx->d_mat;
    if (!brute && !kdt) A.bvh = make_view(ctx->bvh); else A.bvh = BvhView();
    if (!ctx->opt.wide) A.bvh.wide = nullptr;      // (built behind the BVH build when the option was on then)
    if (kdt) A.kd = make_kd_view(ctx); else A.kd = KdView();
    if (p->tri_geometric && ctx->prim_type == 1) { A.prim_type = 2; if (A.bvh.prim_type == 1) A.bvh.prim_type = 2; if (A.kd.prim_type == 1) A.kd.prim_type = 2; }
    A.shade.n_lights = ctx->n_lights;
    for (int i = 0; i < ctx->n_lights; ++i) A.shade.lights[i] = ctx->lights[i];
    const bool bg0 = p->bg[0] == 0 && p->bg[1] == 0 && p->bg[2] == 0;
    A.shade.bg[0] = bg0 ? 0.6f : p->bg[0]; A.shade.bg[1] = bg0 ? 0.8f : p->bg[1]; A.shade.bg[2] = bg0 ? 1.0f : p->bg[2];
    A.shade.bias = p->bias > 0 ? p->bias : 1e-4f;
    A.shade.max_depth = p->max_depth > 0 ? p->max_depth : 2;
    A.shade.shadows = p->shadows;
    const bool full = p->shadows || ctx->has_materials;
    if (full && kdt) { rtds_set_error("render: the KDTREE path traces primary rays only (any-hit and unshaded as main.cpp:362-372, or kd_closest); shadows/materials need BVH, LBVH or NONE"); return RTDS_ERR_UNSUPPORTED; }
    const bool kd_closest = kdt && p->kd_closest;
    A.out_rgb = d_rgb_rows; A.out_hit = d_hit; A.out_accum = d_accum; A.out_global_rows = global_rows ? 1 : 0;
    A.out_vec8 = (((uintptr_t)d_rgb_rows & 7) == 0 && W % 8 == 0) ? 1 : 0;
    A.block_order = ctx->opt.block_order;
    A.counters = ctx->d_counters;
    A.order = nullptr; A.cost = nullptr; A.skip = nullptr; A.heavy_list = nullptr;

    cudaStream_t s = ctx->stream;
    int launches = 0;
    const uint64_t first_word = 4ull * p->jitter_offset;
    const size_t n_words = 4ull * (size_t)W * H * spp;
    // Optional fused strip kernel (strip option): regenerates the jitter words it needs in shared memory, so nothing is
    // expanded into HBM. MEASURED SLOWER on B200 (3.16 ms vs 1.81 + 0.24 ms on the bench frame): a block then covers
    // 312 consecutive pixels of one scanline instead of a 16x8 tile, and the L1 hit rate of the node stream — what
    // this issue-bound kernel lives on — collapses. Kept as a checked (parity-tested) negative result, off by default.
    const bool strip = !brute && !full && !kd_closest && !global_rows && spp <= 150 && ctx->opt.strip == 1;
    // One graph launch per frame (frame_graph option) for renders that stay on the device: rtds_render_device and the multi-GPU
    // shared frame. Host-buffer renders keep the banded multi-stream form below (the download overlaps the rendering there).
    // frames with shadow rays as a wavefront of two kernels (wave_primary / wave_shade): see K10c
    bool wave = full && p->shadows && ctx->opt.wavefront != 0 && ctx->opt.packet != 0 && !kdt && !brute && !p->exact && !strip && spp % PK == 0 && spp <= WAVE_THREADS && A.bvh.leaf_box_prim &&
                      A.bvh.prim_type == 0 && A.bvh.root_ref >= 0 && A.shade.max_depth >= 1 && ctx->n_lights > 0 && A.local_rows > 0;
    if (wave) wave = wave_choose(ctx, A, ctx->n_lights, st != nullptr);
    ctx->wave_this_frame = wave;
    if (wave) {
        const size_t n_loc = (size_t)A.local_rows * W * spp;
        const size_t need = n_loc * (sizeof(float) + sizeof(int)) + 1024;
        if (ctx->wave_bytes < need) {
            RTDS_CUDA(cudaStreamSynchronize(s));
            if (ctx->d_wave) cudaFree(ctx->d_wave);
            ctx->d_wave = nullptr; ctx->wave_bytes = 0;
            RTDS_CUDA(cudaMalloc(&ctx->d_wave, need));
            ctx->wave_bytes = need;
        }
        A.wave_t = (float*)ctx->d_wave;
        A.wave_leaf = (int*)(ctx->d_wave + n_loc * sizeof(float));
    }
    const bool graph_path = ctx->opt.frame_graph != 0 && !wave && !on_band && !strip && A.local_rows > 0 && st != nullptr && !d_hit && !d_accum;
    PrefetchArgs PF;
    const bool prefetch = !strip && A.local_rows > 0 && prefetch_plan(ctx, !brute && !kdt, PF);
    if (!graph_path) RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    if (prefetch && !graph_path) {
        RTDS_CUDA(cudaEventRecord(ctx->ev_pf0, s));
        RTDS_CUDA(cudaStreamWaitEvent(ctx->pf_stream, ctx->ev_pf0, 0));
        l2_prefetch_kernel<<<L2_PREFETCH_BLOCKS, 256, 0, ctx->pf_stream>>>(PF.a, PF.na, PF.b, PF.nb);
        RTDS_CUDA(cudaEventRecord(ctx->ev_pf1, ctx->pf_stream));
        launches += 1;
    }
    const uint64_t chunk_words = (uint64_t)MT_SNAP_EVERY * MT_N;
    const uint64_t c_first = first_word / chunk_words, c_last = (first_word + n_words - 1) / chunk_words;
    DirsLaunch DL;
    bool have_dirs_launch = false;
    if (strip) {
        if (c_last + 3 >= (1ull << 31)) { rtds_set_error("jitter stream position too large"); return RTDS_ERR_INVALID; }
        RTDS_TRY(ensure_snapshots(ctx, (int)(c_last + 1), &launches));
    } else {
        const RayGen G{A.angle, A.aspect, A.inv_w, A.inv_h};
        const uint64_t key[6] = {p->jitter_offset, (uint64_t)W, (uint64_t)H, (uint64_t)spp,
                                 ((uint64_t)__float_as_uint_host(G.angle) << 32) | __float_as_uint_host(G.aspect),
                                 ((uint64_t)(unsigned)rank << 40) | ((uint64_t)(unsigned)world << 20) | (uint64_t)(unsigned)tile_rows};
        if (ctx->dirs_pending && ctx->dirs_valid && memcmp(key, ctx->dirs_key, sizeof key) == 0) {
            // generated ahead of time by rtds_prefetch_dirs for exactly this frame
            RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_dirs, 0));
            launches += ctx->dirs_pending_launches;
        } else {
            if (ctx->dirs_pending) RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_dirs, 0));      // never overwrite d_dirs under a running prefetch
            if (!(p->no_jitter_regen && ctx->dirs_valid && memcmp(key, ctx->dirs_key, sizeof key) == 0)) {
                JitterOwner own{first_word, 4ull * (unsigned long long)W * spp, tile_rows, rank, world, H};
                RTDS_TRY(dirs_plan(ctx, p->jitter_offset, W, H, spp, G, own, &launches, DL));
                have_dirs_launch = true;
                if (!graph_path) {
                    if (DL.grid > 0) RTDS_CUDA(cudaLaunchKernel((const void*)mt_expand_dirs_kernel, dim3(DL.grid), dim3(MT_THREADS), DL.args, 0, s));
                    launches += 1;
                    RTDS_CUDA(cudaGetLastError());
                }
            }
        }
        ctx->dirs_pending = false;
    }
    A.dirs = ctx->d_dirs;
    A.jitter = ctx->d_jitter;
    A.jitter_rel = strip ? 0 : first_word - ctx->jitter_first_word;
    A.mt_snap = ctx->d_mt_snap;
    A.first_sample = p->jitter_offset;
    A.chunk_first = (int)c_first;
    if (graph_path) {
        const void* fn = select_render_kernel(ctx, A, p, full, kdt, brute, kd_closest);
        const unsigned lin = render_grid(ctx, fn, W, A.local_rows);
        const bool shared = global_rows && ctx->shared.frame;
        RTDS_TRY(render_frame_graph(ctx, A, fn, lin, have_dirs_launch ? &DL : nullptr, prefetch ? &PF : nullptr, shared, ctx->shared.seq, &launches));
        RTDS_CUDA(cudaStreamSynchronize(s));        // the graph ends with the counters' copy into pinned memory
        return render_stats_out(ctx, st, launches, A.local_rows, false, false);
    }
    RTDS_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * 8, s));
    if (prefetch) RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_pf1, 0));
    if (A.local_rows > 0 && strip) {
        const int blocks = (int)(c_last - c_first + 1);
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        if (kdt) render_strip_kernel<3><<<blocks, STRIP_THREADS, 0, s>>>(A);
        else if (p->exact) render_strip_kernel<0><<<blocks, STRIP_THREADS, 0, s>>>(A);
        else render_strip_kernel<1><<<blocks, STRIP_THREADS, 0, s>>>(A);
        launches += 1;
        RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
        RTDS_CUDA(cudaGetLastError());
        if (on_band) { RTDS_CUDA(cudaEventRecord(ctx->ev_band, s)); RTDS_TRY((*on_band)(0, A.local_rows, ctx->ev_band)); }
    } else if (A.local_rows > 0) {
        // Row bands (host-buffer renders only): the frame is rendered as n_bands kernels on streams of DESCENDING priority,
        // all released at once. The block scheduler drains the higher-priority band first and fills its tail with the next
        // band's blocks, so the bands finish one after the other without idle SMs, and each band's device->host copy
        // (copy stream, behind the band's event) overlaps the rendering of the remaining bands: only the last band's
        // copy is exposed. (Bands as consecutive kernels on ONE stream were measured slower: every band pays its tail.)
        const int total_rows = A.local_rows;
        int n_bands = 1;
        if (on_band && total_rows >= 512) {
            n_bands = std::max(1, std::min(RTDS_MAX_BANDS, ctx->opt.bands));
            RTDS_TRY(rtds_ensure_band_streams(ctx));
        }
        // Band boundaries (multiples of 8 rows): equal shares by default. RTDS_BAND_RATIO = r (percent) makes every band r %
        // of the one before it. MEASURED (RTDS_TRACE_FRAME=2, bench frame): shrinking bands do not shorten the ~140 us of
        // copies that trail the last kernel (4 equal bands 2.16 ms, 5 bands at 60 % 2.20-2.25, 6 at 70 % 2.16): a band only
        // completes when its longest-running blocks (image centre) are done, the lower-priority bands fill in meanwhile, and
        // the last two or three bands all end within ~50 us of each other whatever their size.
        int band_start[RTDS_MAX_BANDS + 1];
        {
            const int pct = std::max(10, std::min(100, ctx->opt.band_ratio));
            double w[RTDS_MAX_BANDS], sum = 0;
            for (int k = 0; k < n_bands; ++k) { w[k] = k ? w[k - 1] * pct / 100.0 : 1.0; sum += w[k]; }
            band_start[0] = 0;
            double acc = 0;
            for (int k = 1; k < n_bands; ++k) {
                acc += w[k - 1];
                const int r = ((int)(total_rows * (acc / sum)) + 7) & ~7;
                band_start[k] = std::min(total_rows, std::max(band_start[k - 1], r));
            }
            band_start[n_bands] = total_rows;
        }
        // lpt: every band launch starts with the blocks that were heaviest in the previous frame
        const void* fn = wave ? (const void*)wave_primary_kernel : select_render_kernel(ctx, A, p, full, kdt, brute, kd_closest);
        LptBands lb;
        lb.n_bands = n_bands; lb.off[0] = 0;
        for (int bi = 0; bi < n_bands; ++bi) lb.off[bi + 1] = lb.off[bi] + (int)render_grid(ctx, fn, W, std::max(0, band_start[bi + 1] - band_start[bi]));
        RTDS_TRY(lpt_frame_begin(ctx, A, fn, lb, s));
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        if (n_bands > 1) RTDS_CUDA(cudaEventRecord(ctx->ev_ready, s));       // directions + counters are ready behind this
        for (int bi = 0; bi < n_bands; ++bi) {
            const int r0 = band_start[bi], r1 = band_start[bi + 1];
            if (r1 <= r0) continue;
            A.lrow0 = r0;
            A.local_rows = r1;
            cudaStream_t s = ctx->stream;                       // shadows the main stream inside the band loop
            if (n_bands > 1) {
                s = ctx->band_streams[bi];
                RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_ready, 0));
            }
            const dim3 block(render_threads(fn));
            const unsigned lin = render_grid(ctx, fn, W, r1 - r0);
            lpt_band_args(ctx, A, lb, bi);
            const bool split = A.skip != nullptr;      // lpt_split: render_heavy_kernel takes the flagged tiles
            void* kargs[] = {(void*)&A};
            if (split) {
                RTDS_TRY(rtds_ensure_band_streams(ctx));
                cudaStream_t hs = ctx->band_streams[0];                // highest priority: its blocks are placed first
                RTDS_CUDA(cudaEventRecord(ctx->ev_ready, s));
                RTDS_CUDA(cudaStreamWaitEvent(hs, ctx->ev_ready, 0));
                RTDS_CUDA(cudaLaunchKernel((const void*)render_heavy_kernel, dim3(4 * LPT_SPLIT_MAX), dim3(128), kargs, 0, hs));
                RTDS_CUDA(cudaEventRecord(ctx->ev_bands[0], hs));
                launches += 1;
            }
            RTDS_CUDA(cudaLaunchKernel(fn, dim3(lin), block, kargs, 0, s));
            if (split) RTDS_CUDA(cudaStreamWaitEvent(s, ctx->ev_bands[0], 0));
            if (wave) {       // ... followed by the band's shadow rays and shading, one thread per sample
                int tw, th;
                wave_tile(spp, tw, th);
                const unsigned tiles = (unsigned)((W + tw - 1) / tw) * (unsigned)((r1 - r0 + th - 1) / th);
                const bool warp_sum = spp <= 32 && (spp & (spp - 1)) == 0;
                RTDS_CUDA(cudaLaunchKernel(warp_sum ? (const void*)wave_shade_kernel<true> : (const void*)wave_shade_kernel<false>, dim3(tiles), dim3(WAVE_THREADS), kargs, 0, s));
                launches += 1;
            }
            launches += 1;
            // the kernel-time event goes in BEFORE the band callback: a device->host copy into pageable memory blocks the
            // host, and an event recorded after it would time the copy as well
            if (n_bands == 1) {
                RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
                if (on_band) { RTDS_CUDA(cudaEventRecord(ctx->ev_band, s)); RTDS_TRY((*on_band)(r0, r1, ctx->ev_band)); }
            } else {
                cudaEvent_t done = ctx->ev_bands[bi];
                RTDS_CUDA(cudaEventRecord(done, s));
                RTDS_CUDA(cudaStreamWaitEvent(ctx->stream, done, 0));       // the main stream joins every band
            }
        }
        if (n_bands > 1) {
            RTDS_CUDA(cudaEventRecord(ctx->ev3, ctx->stream));
            // every band is queued: now the copies (a copy into pageable host memory blocks the host until it is done, so it
            // must not sit between two launches)
            for (int bi = 0; bi < n_bands; ++bi)
                if (band_start[bi + 1] > band_start[bi]) RTDS_TRY((*on_band)(band_start[bi], band_start[bi + 1], ctx->ev_bands[bi]));
        }
        A.local_rows = total_rows;
        A.lrow0 = 0;
        RTDS_TRY(lpt_frame_end(ctx, lb, ctx->stream));
        RTDS_CUDA(cudaGetLastError());
    } else {
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
    }
    if (global_rows && ctx->shared.frame) RTDS_TRY(rtds_shared_frame_signal_wait(ctx, ctx->shared.seq, &launches));
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    if (st) {
        // pinned: the read-back is on every frame's critical path
        RTDS_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
        RTDS_CUDA(cudaStreamSynchronize(s));
        return render_stats_out(ctx, st, launches, A.local_rows, on_band != nullptr, true);
    }
    return RTDS_OK;
}

// ---------------------------------------------------------------------------------------------------
// Shared frame: completion flags of the multi-GPU frame assembly (one 128-byte line per rank behind the frame)
// ---------------------------------------------------------------------------------------------------
int rtds_shared_frame_signal_wait(rtds_ctx* ctx, uint32_t seq, int* launches)
{
    cudaStream_t s = ctx->stream;
    SharedFrame& f = ctx->shared;
    frame_signal_kernel<<<1, 1, 0, s>>>(f.flags + 32 * f.rank, seq);
    *launches += 1;
    if (f.owner) {
        RTDS_CUDA(cudaMemsetAsync(ctx->d_counters + 7, 0, sizeof(unsigned long long), s));
        frame_wait_kernel<<<1, 32, 0, s>>>(f.flags, f.world, seq, 20000000000ll /* ~10 s */, (int*)(ctx->d_counters + 7));
        *launches += 1;
    }
    RTDS_CUDA(cudaGetLastError());
    return RTDS_OK;
}

int rtds_trace_impl(rtds_ctx* ctx, int acc, int exact, const float* h_o, const float* h_d, int nrays, int* h_hit, float* h_t,
                    rtds_render_stats* st)
{
    if (nrays <= 0) return RTDS_OK;
    if (ctx->n <= 0) { rtds_set_error("trace: no scene"); return RTDS_ERR_NO_SCENE; }
    const bool kdt = acc == RTDS_KDTREE;
    if (kdt && !ctx->kd.valid) { rtds_set_error("trace: KDTREE requested but no KD-tree has been built"); return RTDS_ERR_NOT_BUILT; }
    const bool brute = (acc != RTDS_BVH && acc != RTDS_LBVH && !kdt);
    if (!brute && !kdt) RTDS_TRY(check_bvh(ctx, acc));
    size_t vec = (sizeof(float) * 3 * (size_t)nrays + 255) & ~(size_t)255;
    size_t one = (sizeof(float) * (size_t)nrays + 255) & ~(size_t)255;
    RTDS_TRY(rtds_ensure_scratch(ctx, 2 * vec + 2 * one));
    char* base = (char*)ctx->d_scratch;
    float* d_o = (float*)base; float* d_d = (float*)(base + vec);
    int* d_hit = (int*)(base + 2 * vec); float* d_t = (float*)(base + 2 * vec + one);
    cudaStream_t s = ctx->stream;
    RTDS_CUDA(cudaMemcpyAsync(d_o, h_o, sizeof(float) * 3 * (size_t)nrays, cudaMemcpyHostToDevice, s));
    RTDS_CUDA(cudaMemcpyAsync(d_d, h_d, sizeof(float) * 3 * (size_t)nrays, cudaMemcpyHostToDevice, s));
    RTDS_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * 8, s));
    TraceArgs A;
    A.o = d_o; A.d = d_d; A.nrays = nrays;
    if (!brute && !kdt) A.bvh = make_view(ctx->bvh); else A.bvh = BvhView();
    if (kdt) A.kd = make_kd_view(ctx); else A.kd = KdView();
    A.sph = ctx->d_sph; A.tri = ctx->d_tris; A.prim_type = ctx->prim_type; A.n = ctx->n; A.hit = d_hit; A.t = d_t; A.counters = ctx->d_counters; A.exact = exact & 1;
    if ((exact & RTDS_TRACE_TRI_GEOMETRIC) && ctx->prim_type == 1) { A.prim_type = 2; if (A.bvh.prim_type == 1) A.bvh.prim_type = 2; if (A.kd.prim_type == 1) A.kd.prim_type = 2; }
    RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
    if (kdt && (exact & RTDS_TRACE_KD_CLOSEST)) trace_kernel<4><<<(nrays + 127) / 128, 128, 0, s>>>(A);
    else if (kdt) trace_kernel<3><<<(nrays + 127) / 128, 128, 0, s>>>(A);
    else if (brute) trace_kernel<2><<<(nrays + 127) / 128, 128, 0, s>>>(A);
    else trace_kernel<0><<<(nrays + 127) / 128, 128, 0, s>>>(A);
    RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
    RTDS_CUDA(cudaGetLastError());
    RTDS_CUDA(cudaMemcpyAsync(h_hit, d_hit, sizeof(int) * (size_t)nrays, cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaMemcpyAsync(h_t, d_t, sizeof(float) * (size_t)nrays, cudaMemcpyDeviceToHost, s));
    unsigned long long c[8];
    RTDS_CUDA(cudaMemcpyAsync(c, ctx->d_counters, sizeof c, cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    if (st) {
        memset(st, 0, sizeof *st);
        st->node_tests = c[0]; st->prim_tests = c[1]; st->node_visits = c[2]; st->rays = c[3]; st->primary_rays = c[3];
        RTDS_CUDA(cudaEventElapsedTime(&st->ms_kernel, ctx->ev2, ctx->ev3));
        st->ms_total = st->ms_kernel;
        st->kernel_launches = 1;
    }
    return RTDS_OK;
}

#ifdef RTDS_BLOCK_TIMING
extern "C" int rtds_debug_block_times(unsigned long long* out, int n_blocks)
{
    return cudaMemcpyFromSymbol(out, g_block_times, sizeof(unsigned long long) * 3 * (size_t)n_blocks) == cudaSuccess ? 0 : -2;
}
#endif
