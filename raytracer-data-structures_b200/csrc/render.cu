// render.cu — K10 (ray generation + traversal + ray/sphere test + shading + accumulate + 8-bit quantise)
// and K11 (MT19937 jitter stream).
//
// Reference semantics reproduced here (all citations relative to /root/reference/project/raytracer/):
//   render()            main.cpp:541-566   ray generation, per-pixel sample loop, accumulation order
//   random_double()     main.cpp:503-508   std::mt19937 (seed 5489) + uniform_real_distribution<double>:
//                                          generate_canonical<double,53> = (lo + hi*2^32) / 2^64
//   castRay()           main.cpp:291-500   candidate loop (strict <, first candidate wins), Phong shading
//   boxIntersect()      accelerators.h:668-690  collect every leaf whose ancestor chain passes the slab test
//   boundingBoxIntersection() accelerators.h:588-626  slab test: 6 IEEE divides, no t-range test
//   raySphereIntersect() accelerators.h:79-92  geometric solution
//   Vec3::normalize()   geometry.h:125-134 factor = (float)(1.0 / sqrt((double)n))
//   write_into_file()   main.cpp:516-528   (unsigned char)(min(1, c/aa) * 255)
// The translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true so that every float
// operation rounds exactly like the reference's SSE2 code.
#include "rtds_internal.cuh"
#include <math.h>
#include <algorithm>
#include <functional>
#include <stdlib.h>
#include <string.h>

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

// One regeneration A -> B (624 words) by a block of >= 227 threads, three dependent phases.
__device__ __forceinline__ void mt_regen(const uint32_t* __restrict__ A, uint32_t* __restrict__ B)
{
    const int t = threadIdx.x;
    if (t < 227) B[t] = A[t + MT_M] ^ mt_twist(A[t], A[t + 1]);
    __syncthreads();
    if (t < 227) B[t + 227] = B[t] ^ mt_twist(A[t + 227], A[t + 228]);
    __syncthreads();
    if (t < 169) B[t + 454] = B[t + 227] ^ mt_twist(A[t + 454], A[t + 455]);
    if (t == 169) B[623] = B[396] ^ mt_twist(A[623], B[0]);
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
                                                                    const RayGen G, const JitterOwner own)
{
    __shared__ __align__(16) uint32_t words[(MT_SNAP_EVERY + 1) * MT_N];      // slice 0 = the snapshot
    const int t = threadIdx.x;
    const unsigned long long wlo = (unsigned long long)(s0 + blockIdx.x) * MT_SNAP_EVERY * MT_N;
    if (own.world > 1) {
        const unsigned long long whi = wlo + MT_SNAP_EVERY * MT_N - 1;
        long long ylo = wlo > own.first_word ? (long long)((wlo - own.first_word) / own.row_words) : 0;
        long long yhi = whi > own.first_word ? (long long)((whi - own.first_word) / own.row_words) : 0;
        if (yhi >= own.height) yhi = own.height - 1;
        bool mine = false;
        for (long long tl = ylo / own.tile_rows; tl <= yhi / own.tile_rows; ++tl) mine |= (tl % own.world) == own.rank;
        if (!mine) return;
    }
    const uint32_t* src = snap + (size_t)(s0 + blockIdx.x) * MT_N;
    for (int i = t; i < MT_N; i += MT_THREADS) words[i] = src[i];
    __syncthreads();
    for (int r = 0; r < MT_SNAP_EVERY; ++r) mt_regen(words + r * MT_N, words + (r + 1) * MT_N);
    // all threads, no barriers: temper, canonical doubles, direction, normalise, store
    const unsigned long long as0 = wlo / 4;           // absolute stream sample of the chunk's first four words
    const uint4* w4 = reinterpret_cast<const uint4*>(words + MT_N);
    for (int i = t; i < MT_SNAP_EVERY * MT_N / 4; i += MT_THREADS) {
        const unsigned long long as = as0 + i;
        if (as >= first_sample && as - first_sample < n_samples) {
            const unsigned g = (unsigned)(as - first_sample);
            const unsigned pix = fast_div(g, div_spp);
            const unsigned py = fast_div(pix, div_width), px = pix - py * (unsigned)width;
            const uint4 w = w4[i];
            const uint4 jw = make_uint4(mt_temper(w.x), mt_temper(w.y), mt_temper(w.z), mt_temper(w.w));
            float dx, dy, dz;
            primary_dir(jw, (int)px, (int)py, G, dx, dy, dz);
            float* o = dirs + 3 * (size_t)g;
            o[0] = dx; o[1] = dy; o[2] = dz;
        }
    }
}

// ===================================================================================================
// ray / box / sphere primitives with the reference's exact operation order
// ===================================================================================================
struct Counters { unsigned node_tests, prim_tests, node_visits, rays; };

// accelerators.h:588-626 (and :628-666 for the variant returning tMin/tMax)
__device__ __forceinline__ bool slab_test(float ox, float oy, float oz, float dx, float dy, float dz, float bminx,
                                          float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float& tmin_o,
                                          float& tmax_o)
{
    float tmin = (bminx - ox) / dx;
    float tmax = (bmaxx - ox) / dx;
    if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
    float tymin = (bminy - oy) / dy;
    float tymax = (bmaxy - oy) / dy;
    if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (bminz - oz) / dz;
    float tzmax = (bmaxz - oz) / dz;
    if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    tmin_o = tmin;
    tmax_o = tmax;
    return true;
}

// accelerators.h:79-92; s = {cx,cy,cz,r^2}
__device__ __forceinline__ bool sphere_test(float ox, float oy, float oz, float dx, float dy, float dz, float4 s, float& t0,
                                            float& t1)
{
    float lx = s.x - ox, ly = s.y - oy, lz = s.z - oz;
    float tca = lx * dx + ly * dy + lz * dz;
    if (tca < 0) return false;
    float d2 = (lx * lx + ly * ly + lz * lz) - tca * tca;
    if (d2 > s.w) return false;
    float thc = sqrtf(s.w - d2);
    t0 = tca - thc;
    t1 = tca + thc;
    return true;
}

// Möller–Trumbore as written in the reference's (never compiled) MOLLER_TRUMBORE branch of
// Triangle::rayTriangleIntersect (main.cpp:138-162, `v_0` read as v0, no culling, EPS = 1e-6 main.cpp:59), plus the
// t < 0 rejection of its geometric branch (main.cpp:184). Extension: the reference never instantiates triangles.
__device__ __forceinline__ bool tri_test(float ox, float oy, float oz, float dx, float dy, float dz, float4 v0, float4 v1, float4 v2,
                                         float& t)
{
    const float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
    const float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
    const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;   // dir x v0v2
    const float det = e1x * px + e1y * py + e1z * pz;
    if (fabsf(det) < 1e-6f) return false;
    const float inv = 1 / det;
    const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
    const float u = (tx * px + ty * py + tz * pz) * inv;
    if (u < 0 || u > 1) return false;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;   // tvec x v0v1
    const float v = (dx * qx + dy * qy + dz * qz) * inv;
    if (v < 0 || u + v > 1) return false;
    t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    return !(t < 0);
}

// primitive test on an objId-indexed table (NONE loop, KD leaves)
__device__ __forceinline__ bool obj_test(int type, const float4* __restrict__ sph, const float4* __restrict__ tri, int i, float ox, float oy,
                                         float oz, float dx, float dy, float dz, float& t0, float& t1)
{
    if (type == 0) {
        float4 s = __ldg(sph + i);
        return sphere_test(ox, oy, oz, dx, dy, dz, make_float4(s.x, s.y, s.z, s.w * s.w), t0, t1);
    }
    float t;
    if (!tri_test(ox, oy, oz, dx, dy, dz, __ldg(tri + 3 * (size_t)i), __ldg(tri + 3 * (size_t)i + 1), __ldg(tri + 3 * (size_t)i + 2), t)) return false;
    t0 = t1 = t;
    return true;
}

// un-normalised surface normal at the hit: spheres P - centre (main.cpp:398), triangles v0v1 x v0v2 (main.cpp:165-168)
__device__ __forceinline__ void raw_normal(int type, const float4* __restrict__ sph_c /*centre in .xyz*/, const float4* __restrict__ tri,
                                           size_t idx, float hx, float hy, float hz, float& nx, float& ny, float& nz)
{
    if (type == 0) { float4 s = __ldg(sph_c + idx); nx = hx - s.x; ny = hy - s.y; nz = hz - s.z; }
    else {
        float4 a = __ldg(tri + 3 * idx), b = __ldg(tri + 3 * idx + 1), c = __ldg(tri + 3 * idx + 2);
        float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z, e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
        nx = e1y * e2z - e1z * e2y; ny = e1z * e2x - e1x * e2z; nz = e1x * e2y - e1y * e2x;
    }
}

// candidate update of main.cpp:350-355 / :379-384 with the reference's "first candidate wins" made
// order-independent: `key` is the candidate's position in the reference's candidate order.
__device__ __forceinline__ void candidate(float t0, float t1, int key, int leaf, float& tnear, int& best_key, int& best_leaf)
{
    if (t0 < 0) t0 = t1;
    if (t0 < tnear || (t0 == tnear && best_leaf >= 0 && key < best_key)) {
        tnear = t0; best_key = key; best_leaf = leaf;
    }
}

struct BvhView {
    const Node64* nodes;
    const float4* leaf_sph;
    const float4* leaf_tri;
    int           prim_type;
    const int*    prim_order;
    const int*    leaf_parent;
    int           root_ref;
    int           tie_by_objid;
    int           leaf_box_prim;   // sphere leaves whose box is exactly c -/+ r
    float         root_box[6];
};

constexpr int STACK_MAX = 64;

__device__ __forceinline__ bool leaf_test(const BvhView& B, int leaf, float ox, float oy, float oz, float dx, float dy, float dz, float& t0,
                                          float& t1)
{
    if (B.prim_type == 0) { float4 s = __ldg(B.leaf_sph + leaf); s.w = s.w * s.w; return sphere_test(ox, oy, oz, dx, dy, dz, s, t0, t1); }   // radius2 = r*r, accelerators.h:71
    float t;
    if (!tri_test(ox, oy, oz, dx, dy, dz, __ldg(B.leaf_tri + 3 * (size_t)leaf), __ldg(B.leaf_tri + 3 * (size_t)leaf + 1),
                  __ldg(B.leaf_tri + 3 * (size_t)leaf + 2), t))
        return false;
    t0 = t1 = t;
    return true;
}

// pruning margin for the ordered traversal: bounds the float error of raySphereIntersect's t0 against the
// true entry distance for any sphere inside the root box (DESIGN.md "Ordered traversal is exact").
__device__ __forceinline__ float prune_margin(const float rb[6], float ox, float oy, float oz)
{
    float ex = fmaxf(fabsf(rb[0] - ox), fabsf(rb[3] - ox));
    float ey = fmaxf(fabsf(rb[1] - oy), fabsf(rb[4] - oy));
    float ez = fmaxf(fabsf(rb[2] - oz), fabsf(rb[5] - oz));
    float D = sqrtf(ex * ex + ey * ey + ez * ez);
    return D * 0.00278f;
}

// Conservative slab test for INTERIOR boxes of the ordered traversal: t = (b - o) * (1/d) instead of six IEEE
// divides, with the interval widened by more than the rounding difference between the two forms, so it accepts
// every box the reference's test accepts (and a few more). Interior tests only steer the descent — whether a
// sphere becomes a candidate is decided by the EXACT test on its own leaf box (a leaf box inside a node box
// passes the reference's test only if the node box does: the per-axis intervals nest monotonically).
// Valid for finite, non-tiny direction components (checked once per ray).
constexpr float WIDE_EPS = 4.76837158e-7f;   // 2^-21 > 2^-23 (rcp.approx) + 2^-24 (mul) + 2^-24 (the divide's own rounding), x2 margin
__device__ __forceinline__ bool slab_wide(float ox, float oy, float oz, float ix, float iy, float iz, float bminx, float bminy,
                                          float bminz, float bmaxx, float bmaxy, float bmaxz, float& tmin_o, float& tmax_o)
{
    float x0 = (bminx - ox) * ix, x1 = (bmaxx - ox) * ix;
    float y0 = (bminy - oy) * iy, y1 = (bmaxy - oy) * iy;
    float z0 = (bminz - oz) * iz, z1 = (bmaxz - oz) * iz;
    float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    tmin = tmin - fabsf(tmin) * WIDE_EPS;
    tmax = tmax + fabsf(tmax) * WIDE_EPS;
    tmin_o = tmin;
    tmax_o = tmax;
    return tmin <= tmax;
}

template <bool EXACT>
__device__ __forceinline__ void traverse_bvh(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                             float& tnear, int& best_key, int& best_leaf, Counters& cnt)
{
    float tmn, tmx;
    cnt.node_tests++;
    if (!slab_test(ox, oy, oz, dx, dy, dz, B.root_box[0], B.root_box[1], B.root_box[2], B.root_box[3], B.root_box[4],
                   B.root_box[5], tmn, tmx))
        return;
    if (B.root_ref < 0) {
        float t0, t1;
        cnt.prim_tests++;
        if (leaf_test(B, 0, ox, oy, oz, dx, dy, dz, t0, t1))
            candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order) : 0, 0, tnear, best_key, best_leaf);
        return;
    }
    const float margin = EXACT ? 0.0f : prune_margin(B.root_box, ox, oy, oz);
    // reciprocal direction for the conservative interior test; rays with a zero / tiny / non-finite component
    // take the exact divide-based test everywhere
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    const bool wide_ok = !EXACT && fabsf(ix) < 1e30f && fabsf(iy) < 1e30f && fabsf(iz) < 1e30f;
    int   stack[STACK_MAX];
    float stack_t[EXACT ? 1 : STACK_MAX];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4* q = reinterpret_cast<const float4*>(B.nodes + node);
        float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        int4 q3 = __ldg(reinterpret_cast<const int4*>(q + 3));
        cnt.node_visits++;
        cnt.node_tests += 2;
        const int left = q3.x, right = q3.y;
        float tminL, tmaxL, tminR, tmaxR;
        bool hitL, hitR;
        if (!EXACT && wide_ok) {
            // interior child: the conservative test is the answer; leaf child: it is a filter — a box the wide
            // test rejects is rejected by the reference's test too, only survivors pay for the six divides
            hitL = slab_wide(ox, oy, oz, ix, iy, iz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tminL, tmaxL);
            hitR = slab_wide(ox, oy, oz, ix, iy, iz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tminR, tmaxR);
            if (hitL && left < 0) hitL = slab_test(ox, oy, oz, dx, dy, dz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tminL, tmaxL);
            if (hitR && right < 0) hitR = slab_test(ox, oy, oz, dx, dy, dz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tminR, tmaxR);
        } else {
            hitL = slab_test(ox, oy, oz, dx, dy, dz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tminL, tmaxL);
            hitR = slab_test(ox, oy, oz, dx, dy, dz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tminR, tmaxR);
        }
        if (!EXACT) {
            // NaN-safe: a comparison with NaN is false and keeps the child
            if (hitL && (tminL > tnear + margin || tmaxL < -margin)) hitL = false;
            if (hitR && (tminR > tnear + margin || tmaxR < -margin)) hitR = false;
        }
        if (hitL && left < 0) {
            int leaf = ~left;
            float t0, t1;
            cnt.prim_tests++;
            if (leaf_test(B, leaf, ox, oy, oz, dx, dy, dz, t0, t1))
                candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
            hitL = false;
        }
        if (hitR && right < 0) {
            int leaf = ~right;
            float t0, t1;
            cnt.prim_tests++;
            if (leaf_test(B, leaf, ox, oy, oz, dx, dy, dz, t0, t1))
                candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
            hitR = false;
        }
        if (hitL && hitR) {
            int nearc = left, farc = right;
            float tfar = tminR;
            if (!EXACT && tminR < tminL) { nearc = right; farc = left; tfar = tminL; }
            if (sp < STACK_MAX) {
                stack[sp] = farc;
                if (!EXACT) stack_t[sp] = tfar;
                ++sp;
            }
            node = nearc;
            continue;
        }
        if (hitL) { node = left; continue; }
        if (hitR) { node = right; continue; }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            if (!EXACT && stack_t[sp] > tnear + margin) continue;
            node = stack[sp];
            found = true;
            break;
        }
        if (!found) break;
    }
}

// The ordered traversal (exact = 0), written "while-while": interior nodes and leaves are both stack items, so the
// hot interior loop is branch-light (both children go through the conservative reciprocal test, no leaf special
// case), and a leaf is only opened when it is popped with tmin still below the current hit.
// ZERO_O: the ray starts at the origin (every primary ray, main.cpp:558): t = b * (1/d), one multiply per plane.
// ANYHIT (shadow rays, main.cpp:468-473): stop at the first candidate with t'^2 < t2max. "The nearest hit satisfies
// tNear^2 < lightDistance2" and "some candidate does" are the same predicate (every t' is >= 0), so the answer equals
// the closest-hit formulation's; subtrees that start beyond sqrt(t2max) are never opened.
// OCT: direction octant (bit0: dx < 0, bit1: dy < 0, bit2: dz < 0), a compile-time constant: the near / far plane of
// every slab is then known without min/max (t = plane * (1/d) is monotone in the plane), so a box costs six multiplies,
// one FMNMX3 for the entry and one for the exit distance. OCT < 0: generic form with per-axis min/max.
//
// Conservativeness. With 1e-30 < |1/d| < 1e30 every t_a = (b - o) * rcp(d) differs from the reference's
// t_e = (b - o) / d by at most 2^-22 relative (rcp.approx 2^-23, the multiply 2^-24, the divide's own rounding 2^-24).
//  * interior boxes / leaf filter: exit distance widened by WIDE2 = 2^-20 relative, entry distance left as computed:
//    tmin_e <= tmax_e implies tmin_a <= tmax_a (1 + 2^-20 sign-aware) in all three sign cases, so every box the
//    reference's test accepts is accepted. The pruning bound tlim is widened by the same 2^-20 instead of the entry
//    distance (accepts a superset of "tmin_a (1 - 2^-21) <= tlim").
//  * leaves whose box is the sphere's own box (c -/+ r, main.cpp:686-688): the box is rebuilt from the leaf's sphere
//    record (no parent-record reload) and tested with the same multiplies; if the NARROWED interval (entry pushed up,
//    exit pushed down by 2^-21) is still non-empty the reference's divide-based test accepts for certain; only the
//    ambiguous sliver in between (and denormal-range distances) pays for the six IEEE divides. Other trees (median
//    split with dropped ranges, triangles) run the divide test on the box stored in the parent's record.
// Cold paths of the ordered traversal, kept OUT OF LINE so that the octant copies of the hot loop stay small
// (instruction-cache footprint): arguments and results by value, the view by pointer into the kernel's
// __grid_constant__ parameter block.
struct ColdHit { float tnear; int key, leaf; unsigned node_tests, prim_tests, node_visits; };
static __device__ __noinline__ ColdHit traverse_exact_cold(const BvhView* B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                           float tnear, int key, int leaf)
{
    Counters c = {0, 0, 0, 0};
    traverse_bvh<true>(*B, ox, oy, oz, dx, dy, dz, tnear, key, leaf, c);
    return ColdHit{tnear, key, leaf, c.node_tests, c.prim_tests, c.node_visits};
}
static __device__ __noinline__ bool slab_test_cold(float ox, float oy, float oz, float dx, float dy, float dz, float bx0, float by0,
                                                   float bz0, float bx1, float by1, float bz1)
{
    float a, b;
    return slab_test(ox, oy, oz, dx, dy, dz, bx0, by0, bz0, bx1, by1, bz1, a, b);
}
// leaf whose box is NOT its sphere's own box (triangles; median-split trees with dropped ranges): the reference's
// divide-based test on the box stored in the parent's record, then the primitive
struct ColdLeaf { int pass; float t0, t1; unsigned prim_tests; };
static __device__ __noinline__ ColdLeaf leaf_parent_box_cold(const BvhView* Bp, int leaf, float ox, float oy, float oz, float dx, float dy,
                                                             float dz)
{
    const BvhView& B = *Bp;
    const int lp = __ldg(B.leaf_parent + leaf);
    const float4* q = reinterpret_cast<const float4*>(B.nodes + (lp & 0x7fffffff));
    const float4 q1 = __ldg(q + 1);
    float bx0, by0, bz0, bx1, by1, bz1;
    if (lp < 0) { const float4 q2 = __ldg(q + 2); bx0 = q1.z; by0 = q1.w; bz0 = q2.x; bx1 = q2.y; by1 = q2.z; bz1 = q2.w; }
    else { const float4 q0 = __ldg(q); bx0 = q0.x; by0 = q0.y; bz0 = q0.z; bx1 = q0.w; by1 = q1.x; bz1 = q1.y; }
    ColdLeaf r = {0, 0.f, 0.f, 0u};
    float a, b2;
    if (slab_test(ox, oy, oz, dx, dy, dz, bx0, by0, bz0, bx1, by1, bz1, a, b2)) {
        r.prim_tests = 1;
        r.pass = leaf_test(B, leaf, ox, oy, oz, dx, dy, dz, r.t0, r.t1) ? 1 : 0;
    }
    return r;
}

constexpr float WIDE2 = 9.53674316e-7f;      // 2^-20
constexpr float NARROW_EPS = 4.76837158e-7f; // 2^-21

template <int OCT>
__device__ __forceinline__ void slab_interval(float x0, float y0, float z0, float x1, float y1, float z1, float& tmin, float& tmax)
{
    // (x0,y0,z0) = t of the box's min planes, (x1,y1,z1) = t of its max planes
    if (OCT >= 0) {
        const float xn = (OCT & 1) ? x1 : x0, xf = (OCT & 1) ? x0 : x1;
        const float yn = (OCT & 2) ? y1 : y0, yf = (OCT & 2) ? y0 : y1;
        const float zn = (OCT & 4) ? z1 : z0, zf = (OCT & 4) ? z0 : z1;
        tmin = fmaxf(fmaxf(xn, yn), zn);
        tmax = fminf(fminf(xf, yf), zf);
    } else {
        tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    }
}

template <bool ZERO_O, bool ANYHIT, int OCT>
__device__ __forceinline__ void traverse_fast_loop(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                   float ix, float iy, float iz, float margin, float& tnear, int& best_key,
                                                   int& best_leaf, Counters& cnt, float t2max)
{
    const float neg_margin = -margin;
    // a subtree is opened only while its entry distance is <= tlim (kept widened by 2^-20, see above)
    float tlim = ANYHIT ? sqrtf(t2max) + margin : tnear + margin;
    tlim = __fmaf_rn(fabsf(tlim), WIDE2, tlim);
    int2 stack[STACK_MAX];        // {child ref, entry distance as bits}: one 8-byte local store / load per push / pop
    int sp = 0;
    int node = 0;
    unsigned visits = 0;
    while (true) {
        if (node >= 0) {
            const float4* q = reinterpret_cast<const float4*>(B.nodes + node);
            const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const int2 ch = __ldg(reinterpret_cast<const int2*>(q + 3));
            ++visits;
            float lx0, lx1, ly0, ly1, lz0, lz1, rx0, rx1, ry0, ry1, rz0, rz1;
            if (ZERO_O) {
                lx0 = q0.x * ix; ly0 = q0.y * iy; lz0 = q0.z * iz; lx1 = q0.w * ix; ly1 = q1.x * iy; lz1 = q1.y * iz;
                rx0 = q1.z * ix; ry0 = q1.w * iy; rz0 = q2.x * iz; rx1 = q2.y * ix; ry1 = q2.z * iy; rz1 = q2.w * iz;
            } else {
                lx0 = (q0.x - ox) * ix; ly0 = (q0.y - oy) * iy; lz0 = (q0.z - oz) * iz;
                lx1 = (q0.w - ox) * ix; ly1 = (q1.x - oy) * iy; lz1 = (q1.y - oz) * iz;
                rx0 = (q1.z - ox) * ix; ry0 = (q1.w - oy) * iy; rz0 = (q2.x - oz) * iz;
                rx1 = (q2.y - ox) * ix; ry1 = (q2.z - oy) * iy; rz1 = (q2.w - oz) * iz;
            }
            float tminL, tmaxL, tminR, tmaxR;
            slab_interval<OCT>(lx0, ly0, lz0, lx1, ly1, lz1, tminL, tmaxL);
            slab_interval<OCT>(rx0, ry0, rz0, rx1, ry1, rz1, tminR, tmaxR);
            tmaxL = __fmaf_rn(fabsf(tmaxL), WIDE2, tmaxL);
            tmaxR = __fmaf_rn(fabsf(tmaxR), WIDE2, tmaxR);
            const bool hitL = tminL <= fminf(tmaxL, tlim) && tmaxL >= neg_margin;
            const bool hitR = tminR <= fminf(tmaxR, tlim) && tmaxR >= neg_margin;
            if (hitL && hitR) {
                const bool rfirst = tminR < tminL;
                stack[sp] = make_int2(rfirst ? ch.x : ch.y, __float_as_int(rfirst ? tminL : tminR));
                sp = min(sp + 1, STACK_MAX - 1);
                node = rfirst ? ch.y : ch.x;
                continue;
            }
            if (hitL | hitR) { node = hitL ? ch.x : ch.y; continue; }
        } else {
            const int leaf = ~node;
            bool pass;
            float t0, t1;
            if (B.leaf_box_prim) {
                // sphere leaf: its box is c -/+ r (bit-identical to what the builder stored in the parent's record)
                const float4 s = __ldg(B.leaf_sph + leaf);
                const float bx0 = s.x - s.w, by0 = s.y - s.w, bz0 = s.z - s.w, bx1 = s.x + s.w, by1 = s.y + s.w, bz1 = s.z + s.w;
                float x0, y0, z0, x1, y1, z1;
                if (ZERO_O) { x0 = bx0 * ix; y0 = by0 * iy; z0 = bz0 * iz; x1 = bx1 * ix; y1 = by1 * iy; z1 = bz1 * iz; }
                else {
                    x0 = (bx0 - ox) * ix; y0 = (by0 - oy) * iy; z0 = (bz0 - oz) * iz;
                    x1 = (bx1 - ox) * ix; y1 = (by1 - oy) * iy; z1 = (bz1 - oz) * iz;
                }
                float tmn, tmx;
                slab_interval<OCT>(x0, y0, z0, x1, y1, z1, tmn, tmx);
                pass = __fmaf_rn(fabsf(tmn), NARROW_EPS, tmn) <= __fmaf_rn(-fabsf(tmx), NARROW_EPS, tmx) &&
                       fminf(fabsf(tmn), fabsf(tmx)) > 1e-30f;
                if (!pass) pass = slab_test_cold(ox, oy, oz, dx, dy, dz, bx0, by0, bz0, bx1, by1, bz1);
                if (pass) {
                    cnt.prim_tests++;
                    pass = sphere_test(ox, oy, oz, dx, dy, dz, make_float4(s.x, s.y, s.z, s.w * s.w), t0, t1);
                }
            } else {
                const ColdLeaf r = leaf_parent_box_cold(&B, leaf, ox, oy, oz, dx, dy, dz);
                cnt.prim_tests += r.prim_tests;
                pass = r.pass != 0; t0 = r.t0; t1 = r.t1;
            }
            if (pass) {
                if (ANYHIT) {
                    if (t0 < 0) t0 = t1;
                    if (t0 * t0 < t2max) { tnear = t0; best_leaf = leaf; cnt.node_visits += visits; cnt.node_tests += 2 * visits; return; }
                } else {
                    candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
                    tlim = tnear + margin;
                    tlim = __fmaf_rn(fabsf(tlim), WIDE2, tlim);
                }
            }
        }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            const int2 e = stack[sp];
            if (__int_as_float(e.y) > tlim) continue;
            node = e.x;
            found = true;
            break;
        }
        if (!found) break;
    }
    cnt.node_visits += visits;
    cnt.node_tests += 2 * visits;
}

// ZNEG: the caller guarantees dz < 0 (every primary ray: dz = -1 before normalisation), four octants instead of eight.
template <bool ZERO_O, bool ANYHIT = false, bool ZNEG = false>
__device__ __forceinline__ void traverse_fast(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                              float& tnear, int& best_key, int& best_leaf, Counters& cnt, float t2max = 0.f)
{
    // 1/d only feeds the conservative tests (its error is inside the widening): one MUFU.RCP each instead of an IEEE divide
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(dz));
    const float amin = fminf(fminf(fabsf(ix), fabsf(iy)), fabsf(iz)), amax = fmaxf(fmaxf(fabsf(ix), fabsf(iy)), fabsf(iz));
    if (B.root_ref < 0 || !(amin > 1e-30f && amax < 1e30f)) {
        // single-leaf tree, or a zero / tiny / huge / non-finite direction component: the divide-based traversal
        const ColdHit h = traverse_exact_cold(&B, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf);
        tnear = h.tnear; best_key = h.key; best_leaf = h.leaf;
        cnt.node_tests += h.node_tests; cnt.prim_tests += h.prim_tests; cnt.node_visits += h.node_visits;
        if (ANYHIT && !(best_leaf >= 0 && tnear * tnear < t2max)) best_leaf = -1;
        return;
    }
    // No separate root test: a leaf box that passes the reference's slab test lies inside the root box, which then
    // passes too (nesting), so the candidate set does not depend on it; rays that miss the scene fall out of the
    // first interior visit.
    const float margin = prune_margin(B.root_box, ox, oy, oz);
    const int oct = (dx < 0 ? 1 : 0) | (dy < 0 ? 2 : 0) | ((ZNEG || dz < 0) ? 4 : 0);
#define RTDS_OCT_CASE(o) case o: traverse_fast_loop<ZERO_O, ANYHIT, o>(B, ox, oy, oz, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt, t2max); break;
    switch (oct) {
        RTDS_OCT_CASE(4) RTDS_OCT_CASE(5) RTDS_OCT_CASE(6) RTDS_OCT_CASE(7)
        default:
            if (!ZNEG) switch (oct) { RTDS_OCT_CASE(0) RTDS_OCT_CASE(1) RTDS_OCT_CASE(2) RTDS_OCT_CASE(3) default: break; }
            break;
    }
#undef RTDS_OCT_CASE
}

// ---------------------------------------------------------------------------------------------------
// Packet traversal: the PK = 4 consecutive samples of ONE pixel walk the tree together in one thread.
// Why: the render kernel is bound by the L1 data pipe — every visit moves 56 bytes of node record into each lane's
// registers (ncu: l1tex data-pipe wavefronts 65-75 % of peak, issue slots 61 %). The samples of a pixel differ by
// sub-pixel jitter and walk nearly the same nodes, so one node load (and one stack push / pop) is shared by four
// rays; the slab arithmetic is done per ray on the loaded record.
// Why it returns the same hits: a sphere is a candidate of ray j iff its OWN leaf box passes the reference's slab
// test for ray j (leaf boxes nest in every ancestor's box), a purely leaf-local criterion. The packet descends into a
// child when ANY ray's conservative test accepts it, so each ray sees a superset of the leaves its own traversal
// would open; at a leaf every ray runs its own narrow-accept / divide test and sphere test; pruning uses each ray's
// own bound (a subtree is skipped only when no ray can still improve there), and equal-t candidates are resolved by
// the order-independent key as before. Requires a common direction octant (checked by the caller) and sphere
// leaves with their own boxes (leaf_box_prim).
// ---------------------------------------------------------------------------------------------------
constexpr int PK = 4;
template <int OCT>
__device__ __forceinline__ void traverse_packet(const BvhView& B, const float (&dx)[PK], const float (&dy)[PK], const float (&dz)[PK],
                                                const float (&ix)[PK], const float (&iy)[PK], const float (&iz)[PK], float margin,
                                                float (&tnear)[PK], int (&best_key)[PK], int (&best_leaf)[PK], Counters& cnt)
{
    const float zthr = margin * (9.5367431640625e-7f / 0.00278f);      // prune_margin = 0.00278 * corner distance
    float tlim[PK];
#pragma unroll
    for (int j = 0; j < PK; ++j) { tlim[j] = tnear[j] + margin; tlim[j] = __fmaf_rn(fabsf(tlim[j]), WIDE2, tlim[j]); }
    float tlim_max = fmaxf(fmaxf(tlim[0], tlim[1]), fmaxf(tlim[2], tlim[3]));
    int2 stack[STACK_MAX];
    int sp = 0;
    int node = 0;
    unsigned visits = 0, prim_tests = 0;
    while (true) {
        if (node >= 0) {
            const float4* q = reinterpret_cast<const float4*>(B.nodes + node);
            const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const int2 ch = __ldg(reinterpret_cast<const int2*>(q + 3));
            ++visits;
            // Boxes behind the camera. Every ray of the packet has dz < 0 (OCT bit 2). A sphere the reference's test can
            // hit has float tca >= 0, so its true tca >= -delta (delta = the rounding error of the three-term dot product,
            // <= 2^-21 * |c|), and the sphere contains the ray point at parameter tca, whose z = tca * dz <= delta: every
            // box on its root path has zmin <= delta. zthr = 2^-20 * (distance to the root box's far corner) > delta.
            // ONE compare per child for the whole packet replaces "exit distance >= -margin" per ray.
            // The same holds on x and y with the exit plane the octant selects.
            const bool frontL = q0.z <= zthr && ((OCT & 2) ? q0.y <= zthr : q1.x >= -zthr) && ((OCT & 1) ? q0.x <= zthr : q0.w >= -zthr);
            const bool frontR = q2.x <= zthr && ((OCT & 2) ? q1.w <= zthr : q2.z >= -zthr) && ((OCT & 1) ? q1.z <= zthr : q2.y >= -zthr);
            float kL[PK], kR[PK];                   // entry distance of the rays that accept the child, +inf otherwise
#pragma unroll
            for (int j = 0; j < PK; ++j) {
                float tminL, tmaxL, tminR, tmaxR;
                slab_interval<OCT>(q0.x * ix[j], q0.y * iy[j], q0.z * iz[j], q0.w * ix[j], q1.x * iy[j], q1.y * iz[j], tminL, tmaxL);
                slab_interval<OCT>(q1.z * ix[j], q1.w * iy[j], q2.x * iz[j], q2.y * ix[j], q2.z * iy[j], q2.w * iz[j], tminR, tmaxR);
                tmaxL = __fmaf_rn(fabsf(tmaxL), WIDE2, tmaxL);
                tmaxR = __fmaf_rn(fabsf(tmaxR), WIDE2, tmaxR);
                kL[j] = tminL <= fminf(tmaxL, tlim[j]) ? tminL : INFINITY;
                kR[j] = tminR <= fminf(tmaxR, tlim[j]) ? tminR : INFINITY;
            }
            // packet entry distances: min over the rays that accept the child
            const float tL = frontL ? fminf(fminf(fminf(kL[0], kL[1]), kL[2]), kL[3]) : INFINITY;
            const float tR = frontR ? fminf(fminf(fminf(kR[0], kR[1]), kR[2]), kR[3]) : INFINITY;
            const bool hitL = tL < INFINITY, hitR = tR < INFINITY;
            if (hitL && hitR) {
                const bool rfirst = tR < tL;
                stack[sp] = make_int2(rfirst ? ch.x : ch.y, __float_as_int(rfirst ? tL : tR));
                sp = min(sp + 1, STACK_MAX - 1);
                node = rfirst ? ch.y : ch.x;
                continue;
            }
            if (hitL | hitR) { node = hitL ? ch.x : ch.y; continue; }
        } else {
            const int leaf = ~node;
            const float4 s = __ldg(B.leaf_sph + leaf);
            const float bx0 = s.x - s.w, by0 = s.y - s.w, bz0 = s.z - s.w, bx1 = s.x + s.w, by1 = s.y + s.w, bz1 = s.z + s.w;
            const float4 s2 = make_float4(s.x, s.y, s.z, s.w * s.w);
            int key = leaf;
            if (B.tie_by_objid) key = __ldg(B.prim_order + leaf);
#pragma unroll
            for (int j = 0; j < PK; ++j) {
                float tmn, tmx;
                slab_interval<OCT>(bx0 * ix[j], by0 * iy[j], bz0 * iz[j], bx1 * ix[j], by1 * iy[j], bz1 * iz[j], tmn, tmx);
                bool pass = __fmaf_rn(fabsf(tmn), NARROW_EPS, tmn) <= __fmaf_rn(-fabsf(tmx), NARROW_EPS, tmx) &&
                            fminf(fabsf(tmn), fabsf(tmx)) > 1e-30f;
                // the sliver between "certainly accepted" and "rejected even by the widened interval" takes the divides
                if (!pass && tmn <= __fmaf_rn(fabsf(tmx), WIDE2, tmx))
                    pass = slab_test_cold(0.f, 0.f, 0.f, dx[j], dy[j], dz[j], bx0, by0, bz0, bx1, by1, bz1);
                if (pass) {
                    float t0, t1;
                    ++prim_tests;
                    if (sphere_test(0.f, 0.f, 0.f, dx[j], dy[j], dz[j], s2, t0, t1)) {
                        candidate(t0, t1, key, leaf, tnear[j], best_key[j], best_leaf[j]);
                        tlim[j] = tnear[j] + margin;
                        tlim[j] = __fmaf_rn(fabsf(tlim[j]), WIDE2, tlim[j]);
                    }
                }
            }
            tlim_max = fmaxf(fmaxf(tlim[0], tlim[1]), fmaxf(tlim[2], tlim[3]));
        }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            const int2 e = stack[sp];
            if (__int_as_float(e.y) > tlim_max) continue;
            node = e.x;
            found = true;
            break;
        }
        if (!found) break;
    }
    cnt.node_visits += visits;
    cnt.node_tests += 2 * PK * visits;
    cnt.prim_tests += prim_tests;
}

// single-ray fallback of the packet kernel (mixed octants / degenerate directions), out of line
static __device__ __noinline__ ColdHit trace_primary_cold(const BvhView* B, float dx, float dy, float dz)
{
    Counters c = {0, 0, 0, 0};
    float tnear = INFINITY;
    int key = 0, leaf = -1;
    traverse_fast<true, false, true>(*B, 0.f, 0.f, 0.f, dx, dy, dz, tnear, key, leaf, c);
    return ColdHit{tnear, key, leaf, c.node_tests, c.prim_tests, c.node_visits};
}

// NONE: main.cpp:376-386, spheres staged through shared memory by the whole block (all threads must call).
constexpr int NONE_CHUNK = 1024;
__device__ __forceinline__ void brute_force_block(int type, const float4* __restrict__ sph /*objId order {c,r}*/,
                                                  const float4* __restrict__ tri, int n, bool active,
                                                  float ox, float oy, float oz, float dx, float dy, float dz, float& tnear,
                                                  int& best, Counters& cnt, float4* sh)
{
    const int chunk = type == 0 ? NONE_CHUNK : NONE_CHUNK / 3;
    for (int base = 0; base < n; base += chunk) {
        int m = min(chunk, n - base);
        __syncthreads();
        if (type == 0) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                float4 s = __ldg(sph + base + i);
                sh[i] = make_float4(s.x, s.y, s.z, s.w * s.w);
            }
        } else {
            for (int i = threadIdx.x; i < 3 * m; i += blockDim.x) sh[i] = __ldg(tri + 3 * (size_t)base + i);
        }
        __syncthreads();
        if (active) {
            for (int i = 0; i < m; ++i) {
                float t0, t1;
                bool h;
                if (type == 0) h = sphere_test(ox, oy, oz, dx, dy, dz, sh[i], t0, t1);
                else { float t; h = tri_test(ox, oy, oz, dx, dy, dz, sh[3 * i], sh[3 * i + 1], sh[3 * i + 2], t); t0 = t1 = t; }
                if (h) {
                    if (t0 < 0) t0 = t1;
                    if (t0 < tnear) { tnear = t0; best = base + i; }
                }
            }
            cnt.prim_tests += m;
        }
    }
}

// ===================================================================================================
// kdtreeIntersect (accelerators.h:997-1086): any-hit, front-to-back, 64-entry todo stack
// ===================================================================================================
struct KdView {
    const rtds_kd_node* nodes;
    const int*          prim_idx;
    const float4*       sph;       // objId-indexed {c, r} (kdtreeAllSceneObjects)
    const float4*       tri;       // objId-indexed v0,v1,v2 (extension)
    int                 prim_type;
    float               bounds[6];
};

__device__ __forceinline__ bool kd_any_hit(const KdView& K, float ox, float oy, float oz, float dx, float dy, float dz, Counters& cnt)
{
    float tMin, tMax;
    cnt.node_tests++;
    if (!slab_test(ox, oy, oz, dx, dy, dz, K.bounds[0], K.bounds[1], K.bounds[2], K.bounds[3], K.bounds[4], K.bounds[5], tMin, tMax))
        return false;
    const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    const float inv[3] = {1 / dx, 1 / dy, 1 / dz};
    int   todo_node[64];
    float todo_tmin[64], todo_tmax[64];
    int todoPos = 0;
    int node = 0;
    while (true) {
        const rtds_kd_node nd = K.nodes[node];
        cnt.node_visits++;
        if ((nd.w1 & 3u) == 3u) {
            const int np = (int)nd.w2;
            if (np == 1) {
                float t0, t1;
                cnt.prim_tests++;
                if (obj_test(K.prim_type, K.sph, K.tri, (int)nd.w0, ox, oy, oz, dx, dy, dz, t0, t1)) return true;
            } else {
                for (int i = 0; i < np; ++i) {
                    int prim = __ldg(K.prim_idx + (int)nd.w0 + i);
                    float t0, t1;
                    cnt.prim_tests++;
                    if (obj_test(K.prim_type, K.sph, K.tri, prim, ox, oy, oz, dx, dy, dz, t0, t1)) return true;
                }
            }
            if (todoPos > 0) { --todoPos; node = todo_node[todoPos]; tMin = todo_tmin[todoPos]; tMax = todo_tmax[todoPos]; }
            else break;
        } else {
            const int axis = (int)(nd.w1 & 3u);
            const float split = __uint_as_float(nd.w0);
            const float tPlane = (split - o[axis]) * inv[axis];
            const bool belowFirst = (o[axis] < split) || (o[axis] == split && d[axis] <= 0);
            const int below = node + 1, above = (int)(nd.w1 >> 2);
            const int first = belowFirst ? below : above, second = belowFirst ? above : below;
            if (tPlane > tMax || tPlane <= 0) node = first;
            else if (tPlane < tMin) node = second;
            else {
                if (todoPos < 64) { todo_node[todoPos] = second; todo_tmin[todoPos] = tPlane; todo_tmax[todoPos] = tMax; ++todoPos; }
                node = first;
                tMax = tPlane;
            }
        }
    }
    return false;
}

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
    int      out_vec8;           // out_rgb is 8-byte aligned and width % 8 == 0: warps store whole 8-byte words
    int*     out_hit;            // optional, local rows x width
    float*   out_accum;          // optional, local rows x width x 3
    unsigned long long* counters;
};

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
__device__ __forceinline__ void quadrant_block(int b, int nbx, int nby, int& bx, int& by)
{
    const int hx = nbx >> 1, hy = nby >> 1, wx = nbx - hx;
    const int n0 = hx * hy, n1 = wx * hy, n2 = hx * (nby - hy);
    if (b < n0) { bx = b % hx; by = b / hx; return; }
    b -= n0;
    if (b < n1) { bx = hx + b % wx; by = b / wx; return; }
    b -= n1;
    if (b < n2) { bx = b % hx; by = hy + b / hx; return; }
    b -= n2;
    bx = hx + b % wx; by = hy + b / wx;
}

// K10: one thread per pixel (warp = 8 x 4 pixels, block = 16 x 8), samples looped in order so the per-pixel float
// accumulation matches main.cpp:553-560.
template <int MODE /*0 = BVH exact, 1 = BVH ordered, 2 = NONE, 3 = KDTREE (any-hit, unshaded: main.cpp:362-372)*/>
__global__ void __launch_bounds__(128) render_kernel(const __grid_constant__ RenderArgs A)
{
    __shared__ float4 sh_sph[MODE == 2 ? NONE_CHUNK : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    quadrant_block(blockIdx.x, (A.width + 15) / 16, (A.local_rows - A.lrow0 + 7) / 8, bx, by);
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
        } else if (active) {
            if (MODE == 0) traverse_bvh<true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
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
                if (MODE == 2) raw_normal(A.prim_type, A.sph, A.tri, (size_t)hit_obj, hx, hy, hz, nx, ny, nz);
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
static __device__ __noinline__ ShadowHit shadow_query_cold(const BvhView* B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                          float dist2)
{
    Counters c = {0, 0, 0, 0};
    float ts = INFINITY;
    int bk = 0, bl = -1;
    const float len2 = dx * dx + dy * dy + dz * dz;
    int occ;
    if (fabsf(len2 - 1.0f) < 1e-3f) {
        traverse_fast<false, true>(*B, ox, oy, oz, dx, dy, dz, ts, bk, bl, c, dist2);
        occ = bl >= 0;
    } else {
        traverse_bvh<true>(*B, ox, oy, oz, dx, dy, dz, ts, bk, bl, c);
        occ = bl >= 0 && ts * ts < dist2;
    }
    return ShadowHit{occ, c.node_tests, c.prim_tests, c.node_visits};
}

// castRay's DIFFUSE_AND_GLOSSY branch with the shadow query evaluated (main.cpp:447-494), as in render_full_kernel
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
        const ShadowHit sh = shadow_query_cold(&A.bvh, sx, sy, sz, lx, ly, lz, dist2);
        cnt.node_tests += sh.node_tests; cnt.prim_tests += sh.prim_tests; cnt.node_visits += sh.node_visits;
        const float lit = sh.occluded ? 0.0f : 1.0f;                                   // main.cpp:471-472
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
#ifndef RTDS_PK_MINB
#define RTDS_PK_MINB 6   // measured on B200: 4 blocks (112 regs) 1.40 ms, 5 (96) 1.245, 6 (80, 188 B spilled) 1.223, 8 (64) 1.238
#endif
template <bool SHADOWS /*evaluate the shadow query (extension; the reference's trace_more is a stub)*/>
__global__ void __launch_bounds__(128, RTDS_PK_MINB) render_packet_kernel(const __grid_constant__ RenderArgs A)
{
    unsigned shadow_rays = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    quadrant_block(blockIdx.x, (A.width + 15) / 16, (A.local_rows - A.lrow0 + 7) / 8, bx, by);
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int lrow = A.lrow0 + by * 8 + (warp >> 1) * 4 + (lane >> 3);
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
            float dx[PK], dy[PK], dz[PK], ix[PK], iy[PK], iz[PK], tnear[PK];
            int best_key[PK], best_leaf[PK];
            bool ok = A.bvh.root_ref >= 0;
            int oct0 = 0;
            // the packet's 4 directions are 48 contiguous, 16-byte aligned bytes (main.cpp:554-557, from mt_expand_dirs_kernel)
            const float4* dp = reinterpret_cast<const float4*>(A.dirs + 3 * (pix * A.spp + k0));
            const float4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
            dx[0] = d0.x; dy[0] = d0.y; dz[0] = d0.z; dx[1] = d0.w; dy[1] = d1.x; dz[1] = d1.y;
            dx[2] = d1.z; dy[2] = d1.w; dz[2] = d2.x; dx[3] = d2.y; dy[3] = d2.z; dz[3] = d2.w;
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
            cnt.rays += PK;
            if (ok) {
                switch (oct0) {
                    case 4: traverse_packet<4>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
                    case 5: traverse_packet<5>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
                    case 6: traverse_packet<6>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
                    default: traverse_packet<7>(A.bvh, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt); break;
                }
            } else {
#pragma unroll
                for (int j = 0; j < PK; ++j) {
                    const ColdHit h = trace_primary_cold(&A.bvh, dx[j], dy[j], dz[j]);
                    tnear[j] = h.tnear; best_leaf[j] = h.leaf;
                    cnt.node_tests += h.node_tests; cnt.prim_tests += h.prim_tests; cnt.node_visits += h.node_visits;
                }
            }
#pragma unroll
            for (int j = 0; j < PK; ++j) {
                float r, g, b;
                int hit_obj = -1;
                if (best_leaf[j] >= 0) hit_obj = __ldg(A.bvh.prim_order + best_leaf[j]);
                if (hit_obj < 0) { r = A.shade.bg[0]; g = A.shade.bg[1]; b = A.shade.bg[2]; }
                else {
                    const float4 m = __ldg(A.mat + hit_obj);
                    const float hx = 0.f + dx[j] * tnear[j], hy = 0.f + dy[j] * tnear[j], hz = 0.f + dz[j] * tnear[j];     // main.cpp:396
                    float nx, ny, nz;
                    raw_normal(0, A.bvh.leaf_sph, nullptr, (size_t)best_leaf[j], hx, hy, hz, nx, ny, nz);
                    if (SHADOWS) shade_diffuse_shadowed(A, dx[j], dy[j], dz[j], hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b, cnt, shadow_rays);
                    else shade_diffuse(A.shade, dx[j], dy[j], dz[j], hx, hy, hz, nx, ny, nz, m.x, m.y, m.z, r, g, b);
                }
                acc_r += r; acc_g += g; acc_b += b;     // sample order, main.cpp:553-560
                last_hit = hit_obj;
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
    unsigned v[5] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays, shadow_rays};
#pragma unroll
    for (int c = 0; c < (SHADOWS ? 5 : 4); ++c) {
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
                if (MODE == 0) traverse_bvh<true>(A.bvh, 0.f, 0.f, 0.f, dx, dy, dz, tnear, best_key, best_leaf, cnt);
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
        if (MODE == 0 || !(fabsf(len2 - 1.0f) < 1e-3f)) traverse_bvh<true>(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
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

template <int MODE /*0 = BVH exact, 1 = BVH ordered, 2 = NONE*/>
__global__ void __launch_bounds__(128) render_full_kernel(const __grid_constant__ RenderArgs A)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx, by;
    quadrant_block(blockIdx.x, (A.width + 15) / 16, (A.local_rows - A.lrow0 + 7) / 8, bx, by);
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
        const float bias = A.shade.bias;
        for (int k = 0; k < A.spp; ++k) {
            const float* dp = A.dirs + 3 * (pix * A.spp + k);      // main.cpp:554-557, computed by mt_expand_dirs_kernel
            float dx = __ldg(dp), dy = __ldg(dp + 1), dz = __ldg(dp + 2);
            float ox = 0, oy = 0, oz = 0;
            float r = A.shade.bg[0], g = A.shade.bg[1], b = A.shade.bg[2];
            float mult[2] = {1.f, 1.f};   // (1 - kr) of the REFLECTION_AND_REFRACTION hits on the way, outermost first
            int nmult = 0;
            for (int depth = 1;; ++depth) {
                if (depth > A.shade.max_depth) break;                        // main.cpp:311-313: sky
                float tnear;
                int hit_obj, hit_leaf;
                closest_hit<MODE>(A, ox, oy, oz, dx, dy, dz, tnear, hit_obj, hit_leaf, cnt);
                cnt.rays++;
                if (depth == 1) last_hit = hit_obj;
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
                        if (occluded<MODE>(A, sx, sy, sz, lx, ly, lz, dist2, cnt)) lit = 0.0f;   // main.cpp:471-472
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
    unsigned v[6] = {cnt.node_tests, cnt.prim_tests, cnt.node_visits, cnt.rays, shadow_rays, secondary_rays};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        unsigned long long x = v[c];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&A.counters[c], x);
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

template <int MODE /*0 BVH, 2 NONE, 3 KDTREE*/>
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
    } else if (active) {
        // the ordered traversal's pruning bound assumes a unit direction (as every ray castRay makes has)
        float len2 = dx * dx + dy * dy + dz * dz;
        bool unit = fabsf(len2 - 1.0f) < 1e-3f;
        if (A.exact || !unit) traverse_bvh<true>(A.bvh, ox, oy, oz, dx, dy, dz, tnear, best_key, best_leaf, cnt);
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
// `first_sample`; only the chunks that hold rows of `own` are filled.
static int dirs_prepare(rtds_ctx* ctx, uint64_t first_sample, int W, int H, int spp, const RayGen& G, const JitterOwner& own, int* launches,
                        cudaStream_t stream)
{
    const uint64_t n_samples = (uint64_t)W * H * spp;
    if (n_samples >= (1ull << 32)) { rtds_set_error("render: width*height*aa_samples must be below 2^32"); return RTDS_ERR_INVALID; }
    const uint64_t words_per_snap = (uint64_t)MT_SNAP_EVERY * MT_N;
    const uint64_t first_word = 4 * first_sample, n_words = 4 * n_samples;
    const uint64_t s0 = first_word / words_per_snap, s1 = (first_word + n_words - 1) / words_per_snap;
    if (s1 + 2 >= (1ull << 31)) { rtds_set_error("jitter stream position too large"); return RTDS_ERR_INVALID; }
    RTDS_TRY(ensure_snapshots(ctx, (int)(s1 + 1), launches));
    if (ctx->dirs_cap_floats < 3 * n_samples) {
        if (ctx->d_dirs) cudaFree(ctx->d_dirs);
        ctx->d_dirs = nullptr; ctx->dirs_cap_floats = 0;
        RTDS_CUDA(cudaMalloc(&ctx->d_dirs, sizeof(float) * 3 * n_samples));
        ctx->dirs_cap_floats = 3 * n_samples;
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
    mt_expand_dirs_kernel<<<(unsigned)(s1 - s0 + 1), MT_THREADS, 0, stream>>>(ctx->d_mt_snap, (int)s0, ctx->d_dirs, first_sample,
                                                                                 (unsigned)n_samples, W, spp, make_div((uint32_t)spp),
                                                                                 make_div((uint32_t)W), G, own);
    if (launches) *launches += 1;
    RTDS_CUDA(cudaGetLastError());
    ctx->dirs_key[0] = first_sample; ctx->dirs_key[1] = (uint64_t)W; ctx->dirs_key[2] = (uint64_t)H; ctx->dirs_key[3] = (uint64_t)spp;
    ctx->dirs_key[4] = ((uint64_t)__float_as_uint_host(G.angle) << 32) | __float_as_uint_host(G.aspect);
    ctx->dirs_key[5] = ((uint64_t)(unsigned)own.rank << 40) | ((uint64_t)(unsigned)own.world << 20) | (uint64_t)(unsigned)own.tile_rows;
    ctx->dirs_valid = true;
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

int rtds_prefetch_dirs(rtds_ctx* ctx, const rtds_render_params* p)
{
    const int W = p->width, H = p->height, spp = p->aa_samples;
    if (W <= 0 || H <= 0 || spp <= 0) return RTDS_OK;        // the render call reports the error
    const int world = p->world > 0 ? p->world : 1, rank = p->rank, tile_rows = p->tile_rows > 0 ? p->tile_rows : 8;
    if (rank < 0 || rank >= world) return RTDS_OK;
    if (getenv("RTDS_STRIP") && atoi(getenv("RTDS_STRIP")) == 1) return RTDS_OK;
    if (ctx->dirs_pending) RTDS_CUDA(cudaStreamSynchronize(ctx->jit_stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));            // a render still reading d_dirs
    const RayGen G = make_raygen(p);
    JitterOwner own{4ull * p->jitter_offset, 4ull * (unsigned long long)W * spp, tile_rows, rank, world, H};
    int launches = 0;
    RTDS_TRY(dirs_prepare(ctx, p->jitter_offset, W, H, spp, G, own, &launches, ctx->jit_stream));
    RTDS_CUDA(cudaEventRecord(ctx->ev_dirs, ctx->jit_stream));
    ctx->dirs_pending = true;
    ctx->dirs_pending_launches = launches;
    return RTDS_OK;
}

int rtds_render_impl(rtds_ctx* ctx, int acc, const rtds_render_params* p, uint8_t* d_rgb_rows, int* d_hit, float* d_accum,
                     rtds_render_stats* st, const std::function<int(int, int)>* on_band, bool global_rows)
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
    A.rank = rank; A.world = world; A.tile_rows = tile_rows; A.lrow0 = 0;
    A.local_rows = rtds_rows_for_rank(H, tile_rows, rank, world);
    { const RayGen G0 = make_raygen(p); A.inv_w = G0.inv_w; A.inv_h = G0.inv_h; A.aspect = G0.aspect; A.angle = G0.angle; }
    A.n = ctx->n; A.sph = ctx->d_sph; A.tri = ctx->d_tris; A.prim_type = ctx->prim_type; A.mat = ctx->d_mat;
    if (!brute && !kdt) A.bvh = make_view(ctx->bvh); else A.bvh = BvhView();
    if (kdt) A.kd = make_kd_view(ctx); else A.kd = KdView();
    A.shade.n_lights = ctx->n_lights;
    for (int i = 0; i < ctx->n_lights; ++i) A.shade.lights[i] = ctx->lights[i];
    const bool bg0 = p->bg[0] == 0 && p->bg[1] == 0 && p->bg[2] == 0;
    A.shade.bg[0] = bg0 ? 0.6f : p->bg[0]; A.shade.bg[1] = bg0 ? 0.8f : p->bg[1]; A.shade.bg[2] = bg0 ? 1.0f : p->bg[2];
    A.shade.bias = p->bias > 0 ? p->bias : 1e-4f;
    A.shade.max_depth = p->max_depth > 0 ? p->max_depth : 2;
    A.shade.shadows = p->shadows;
    const bool full = p->shadows || ctx->has_materials;
    if (full && kdt) { rtds_set_error("render: the KDTREE path is any-hit and unshaded (main.cpp:362-372); shadows/materials need BVH, LBVH or NONE"); return RTDS_ERR_UNSUPPORTED; }
    A.out_rgb = d_rgb_rows; A.out_hit = d_hit; A.out_accum = d_accum; A.out_global_rows = global_rows ? 1 : 0;
    A.out_vec8 = (((uintptr_t)d_rgb_rows & 7) == 0 && W % 8 == 0) ? 1 : 0;
    A.counters = ctx->d_counters;

    cudaStream_t s = ctx->stream;
    int launches = 0;
    RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    const uint64_t first_word = 4ull * p->jitter_offset;
    const size_t n_words = 4ull * (size_t)W * H * spp;
    // Optional fused strip kernel (RTDS_STRIP=1): regenerates the jitter words it needs in shared memory, so nothing is
    // expanded into HBM. MEASURED SLOWER on B200 (3.16 ms vs 1.81 + 0.24 ms on the bench frame): a block then covers
    // 312 consecutive pixels of one scanline instead of a 16x8 tile, and the L1 hit rate of the node stream — what
    // this issue-bound kernel lives on — collapses. Kept as a checked (parity-tested) negative result, off by default.
    const bool strip = !brute && !full && !global_rows && spp <= 150 && getenv("RTDS_STRIP") && atoi(getenv("RTDS_STRIP")) == 1;
    const uint64_t chunk_words = (uint64_t)MT_SNAP_EVERY * MT_N;
    const uint64_t c_first = first_word / chunk_words, c_last = (first_word + n_words - 1) / chunk_words;
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
                RTDS_TRY(dirs_prepare(ctx, p->jitter_offset, W, H, spp, G, own, &launches, s));
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
    RTDS_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * 8, s));
    if (A.local_rows > 0 && strip) {
        const int blocks = (int)(c_last - c_first + 1);
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        if (kdt) render_strip_kernel<3><<<blocks, STRIP_THREADS, 0, s>>>(A);
        else if (p->exact) render_strip_kernel<0><<<blocks, STRIP_THREADS, 0, s>>>(A);
        else render_strip_kernel<1><<<blocks, STRIP_THREADS, 0, s>>>(A);
        launches += 1;
        RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
        RTDS_CUDA(cudaGetLastError());
        if (on_band) RTDS_TRY((*on_band)(0, A.local_rows));
    } else if (A.local_rows > 0) {
        // Row bands: with a band callback (host-buffer render) each band's device->host copy is queued on the copy
        // stream as soon as its kernel is queued, so the frame download overlaps the rendering of the next bands.
        const int total_rows = A.local_rows;
        int n_bands = 1;
        if (on_band && total_rows >= 512) { const char* e = getenv("RTDS_BANDS"); n_bands = e ? std::max(1, atoi(e)) : 1; }
        const int band_rows = ((total_rows + n_bands - 1) / n_bands + 7) & ~7;
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        for (int r0 = 0; r0 < total_rows; r0 += band_rows) {
            const int r1 = std::min(total_rows, r0 + band_rows);
            A.lrow0 = r0;
            A.local_rows = r1;
            const dim3 block(128);
            const unsigned lin = (unsigned)((W + 15) / 16) * (unsigned)((r1 - r0 + 7) / 8);     // quadrant-major linear grid
            // four samples of a pixel per thread as one packet (see traverse_packet); RTDS_PACKET=0 turns it off
            // (shadows without reflective / refractive materials stay on the packet kernel; the shadow rays are single)
            bool packet = (!full || !ctx->has_materials) && !kdt && !brute && !p->exact && spp % PK == 0 && A.bvh.leaf_box_prim &&
                          A.shade.max_depth >= 1;
            if (const char* e = getenv("RTDS_PACKET")) packet = packet && atoi(e) != 0;
            if (packet && full) render_packet_kernel<true><<<lin, block, 0, s>>>(A);
            else if (full) {
                if (brute) render_full_kernel<2><<<lin, block, 0, s>>>(A);
                else if (p->exact) render_full_kernel<0><<<lin, block, 0, s>>>(A);
                else render_full_kernel<1><<<lin, block, 0, s>>>(A);
            }
            else if (packet) render_packet_kernel<false><<<lin, block, 0, s>>>(A);
            else if (kdt) render_kernel<3><<<lin, block, 0, s>>>(A);
            else if (brute) render_kernel<2><<<lin, block, 0, s>>>(A);
            else if (p->exact) render_kernel<0><<<lin, block, 0, s>>>(A);
            else render_kernel<1><<<lin, block, 0, s>>>(A);
            launches += 1;
            // the kernel-time event goes in BEFORE the band callback: a device->host copy into pageable memory blocks the
            // host, and an event recorded after it would time the copy as well
            if (r1 == total_rows) RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
            if (on_band) RTDS_TRY((*on_band)(r0, r1));
        }
        A.local_rows = total_rows;
        RTDS_CUDA(cudaGetLastError());
    } else {
        RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
        RTDS_CUDA(cudaEventRecord(ctx->ev3, s));
    }
    if (global_rows && ctx->shared.frame) RTDS_TRY(rtds_shared_frame_signal_wait(ctx, ctx->shared.seq, &launches));
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    if (st) {
        unsigned long long c[8];
        RTDS_CUDA(cudaMemcpyAsync(c, ctx->d_counters, sizeof c, cudaMemcpyDeviceToHost, s));
        RTDS_CUDA(cudaStreamSynchronize(s));
        memset(st, 0, sizeof *st);
        st->node_tests = c[0]; st->prim_tests = c[1]; st->node_visits = c[2];
        st->rays = c[3]; st->shadow_rays = c[4]; st->secondary_rays = c[5];
        st->primary_rays = c[3] - c[4] - c[5];
        RTDS_CUDA(cudaEventElapsedTime(&st->ms_kernel, ctx->ev2, ctx->ev3));
        RTDS_CUDA(cudaEventElapsedTime(&st->ms_total, ctx->ev0, ctx->ev1));
        st->kernel_launches = launches;
        st->rows = A.local_rows;
        st->reserved[0] = (int)(unsigned)c[7];      // shared frame: 1 + the rank the owner timed out on (0 = complete)
        if (on_band) RTDS_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    }
    return RTDS_OK;
}

// ---------------------------------------------------------------------------------------------------
// Shared frame: completion flags of the multi-GPU frame assembly (one 128-byte line per rank behind the frame)
// ---------------------------------------------------------------------------------------------------
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
    A.sph = ctx->d_sph; A.tri = ctx->d_tris; A.prim_type = ctx->prim_type; A.n = ctx->n; A.hit = d_hit; A.t = d_t; A.counters = ctx->d_counters; A.exact = exact;
    RTDS_CUDA(cudaEventRecord(ctx->ev2, s));
    if (kdt) trace_kernel<3><<<(nrays + 127) / 128, 128, 0, s>>>(A);
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
