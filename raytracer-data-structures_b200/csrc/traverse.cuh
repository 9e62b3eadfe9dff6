// traverse.cuh — ray / box / primitive tests with the reference's operation order, and the traversals built on them:
//   boundingBoxIntersection() accelerators.h:588-626   slab test: 6 IEEE divides, no t-range test
//   raySphereIntersect()      accelerators.h:79-92     geometric solution
//   boxIntersect()            accelerators.h:668-690   collect every leaf whose ancestor chain passes the slab test
//   castRay() candidate loops main.cpp:343-358, :376-386  strict <, first candidate wins
//   kdtreeIntersect()         accelerators.h:997-1086  any-hit, front-to-back
// traverse_bvh_exact = the reference's traversal; traverse_fast / traverse_packet = ordered, pruned traversals that
// return the same hits. Included by render.cu only.
#pragma once
#include "rtds_internal.cuh"

namespace {

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

// Triangle::rayTriangleIntersect as the reference COMPILES it (main.cpp:163-215: the geometric branch, MOLLER_TRUMBORE is never
// defined): plane hit with N = v0v1 x v0v2, t = (N.orig + N.v0) / N.dir (right for orig = 0 only - kept bug for bug), three
// inside-outside edge tests. Selected by rtds_render_params.tri_geometric / RTDS_TRACE_TRI_GEOMETRIC (prim_type 2 in the
// views). Pinned to the reference's own class through the compiled reference (tests/test_gpu_triangles.py).
__device__ __forceinline__ bool tri_test_geometric(float ox, float oy, float oz, float dx, float dy, float dz, float4 v0, float4 v1, float4 v2,
                                                   float& t)
{
    const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;            // v0v1
    const float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;            // v0v2
    const float Nx = ay * bz - az * by, Ny = az * bx - ax * bz, Nz = ax * by - ay * bx;
    const float nd = Nx * dx + Ny * dy + Nz * dz;
    if (fabsf(nd) < 1e-6f) return false;                                         // EPS, main.cpp:59,175
    const float dd = Nx * v0.x + Ny * v0.y + Nz * v0.z;
    t = ((Nx * ox + Ny * oy + Nz * oz) + dd) / nd;                               // main.cpp:182
    if (t < 0) return false;
    const float Px = ox + dx * t, Py = oy + dy * t, Pz = oz + dz * t;
    {
        const float px = Px - v0.x, py = Py - v0.y, pz = Pz - v0.z;
        const float cx = ay * pz - az * py, cy = az * px - ax * pz, cz = ax * py - ay * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    {
        const float ex = v2.x - v1.x, ey = v2.y - v1.y, ez = v2.z - v1.z;
        const float px = Px - v1.x, py = Py - v1.y, pz = Pz - v1.z;
        const float cx = ey * pz - ez * py, cy = ez * px - ex * pz, cz = ex * py - ey * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    {
        const float ex = v0.x - v2.x, ey = v0.y - v2.y, ez = v0.z - v2.z;
        const float px = Px - v2.x, py = Py - v2.y, pz = Pz - v2.z;
        const float cx = ey * pz - ez * py, cy = ez * px - ex * pz, cz = ex * py - ey * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    return true;
}
// type 1: Moeller-Trumbore, type 2: the reference's compiled geometric test
__device__ __forceinline__ bool tri_test_any(int type, float ox, float oy, float oz, float dx, float dy, float dz, float4 v0, float4 v1,
                                             float4 v2, float& t)
{
    return type == 2 ? tri_test_geometric(ox, oy, oz, dx, dy, dz, v0, v1, v2, t) : tri_test(ox, oy, oz, dx, dy, dz, v0, v1, v2, t);
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
    if (!tri_test_any(type, ox, oy, oz, dx, dy, dz, __ldg(tri + 3 * (size_t)i), __ldg(tri + 3 * (size_t)i + 1), __ldg(tri + 3 * (size_t)i + 2), t)) return false;
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
    int           prim_type;       // 0 spheres, 1 triangles (Moeller-Trumbore), 2 triangles (the reference's compiled geometric test)
    const int*    prim_order;
    const int*    leaf_parent;
    int           root_ref;
    int           tie_by_objid;
    int           leaf_box_prim;   // sphere leaves whose box is exactly c -/+ r
    float         root_box[6];
    const Wide4*  wide;            // 4-wide collapse of `nodes` (wide option), or nullptr
};

constexpr int STACK_MAX = 64;

__device__ __forceinline__ bool leaf_test(const BvhView& B, int leaf, float ox, float oy, float oz, float dx, float dy, float dz, float& t0,
                                          float& t1)
{
    if (B.prim_type == 0) { float4 s = __ldg(B.leaf_sph + leaf); s.w = s.w * s.w; return sphere_test(ox, oy, oz, dx, dy, dz, s, t0, t1); }   // radius2 = r*r, accelerators.h:71
    float t;
    if (!tri_test_any(B.prim_type, ox, oy, oz, dx, dy, dz, __ldg(B.leaf_tri + 3 * (size_t)leaf), __ldg(B.leaf_tri + 3 * (size_t)leaf + 1),
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

// The reference's traversal (exact = 1; boxIntersect, accelerators.h:668-690 + the candidate loop, main.cpp:343-358):
// every node whose divide-based slab test passes is visited, no ordering, no pruning, so the candidate set is the
// reference's. Also the fallback of the ordered traversal for rays it does not handle.
__device__ __forceinline__ void traverse_bvh_exact(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
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
    int stack[STACK_MAX];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4* q = reinterpret_cast<const float4*>(B.nodes + node);
        const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        const int4 q3 = __ldg(reinterpret_cast<const int4*>(q + 3));
        cnt.node_visits++;
        cnt.node_tests += 2;
        const int left = q3.x, right = q3.y;
        float tminL, tmaxL, tminR, tmaxR;
        bool hitL = slab_test(ox, oy, oz, dx, dy, dz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tminL, tmaxL);
        bool hitR = slab_test(ox, oy, oz, dx, dy, dz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tminR, tmaxR);
        if (hitL && left < 0) {
            const int leaf = ~left;
            float t0, t1;
            cnt.prim_tests++;
            if (leaf_test(B, leaf, ox, oy, oz, dx, dy, dz, t0, t1))
                candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
            hitL = false;
        }
        if (hitR && right < 0) {
            const int leaf = ~right;
            float t0, t1;
            cnt.prim_tests++;
            if (leaf_test(B, leaf, ox, oy, oz, dx, dy, dz, t0, t1))
                candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
            hitR = false;
        }
        if (hitL && hitR) {
            if (sp < STACK_MAX) stack[sp++] = right;
            node = left;
            continue;
        }
        if (hitL) { node = left; continue; }
        if (hitR) { node = right; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
}

// Cold paths of the ordered traversal, kept OUT OF LINE so that the octant copies of the hot loop stay small
// (instruction-cache footprint): arguments and results by value, the view by pointer into the kernel's
// __grid_constant__ parameter block.
struct ColdHit { float tnear; int key, leaf; unsigned node_tests, prim_tests, node_visits; };
static __device__ __noinline__ ColdHit traverse_exact_cold(const BvhView* B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                           float tnear, int key, int leaf)
{
    Counters c = {0, 0, 0, 0};
    traverse_bvh_exact(*B, ox, oy, oz, dx, dy, dz, tnear, key, leaf, c);
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

// Interior boxes of a ray with a non-zero origin (shadow and secondary rays): t = plane * (1/d) + c, ONE FFMA per plane with
// c = -(o * (1/d)) -/+ w folded in per axis (near planes get c - w, far planes c + w), instead of a subtract, a multiply and a
// separate widening of the exit distance. w bounds the distance between this t and the reference's fl(fl(plane - o) / d):
//   t = (plane * i - o * i * (1 + e4)) * (1 + e5), i = (1/d)(1 + e3), |e3| <= 2^-22 (MUFU.RCP), |e4|, |e5| <= 2^-24
//   |t - (plane - o)/d| <= |plane - o| |i| (2^-22 + 2^-24 + ..) + |o| |i| 2^-24, and the reference's value is within 2^-23 relative
//   of (plane - o)/d: together < (|plane| + |o|) |i| 2^-21; w = (max |plane| over the root box + |o|) |i| 2^-20 is twice that.
// Near values only ever decrease and far values only increase, so the test accepts every box the reference's test accepts
// (interior tests only steer: the candidate criterion is leaf-local, and the leaf test below keeps the subtract form).
struct RayAffine { float cnx, cny, cnz, cfx, cfy, cfz; };
__device__ __forceinline__ RayAffine ray_affine(const float rb[6], float ox, float oy, float oz, float ix, float iy, float iz, float& wmax)
{
    const float k = 9.53674316e-7f;     // 2^-20
    const float wx = ((fmaxf(fabsf(rb[0]), fabsf(rb[3])) + fabsf(ox)) * fabsf(ix)) * k;
    const float wy = ((fmaxf(fabsf(rb[1]), fabsf(rb[4])) + fabsf(oy)) * fabsf(iy)) * k;
    const float wz = ((fmaxf(fabsf(rb[2]), fabsf(rb[5])) + fabsf(oz)) * fabsf(iz)) * k;
    const float cx = -(ox * ix), cy = -(oy * iy), cz = -(oz * iz);
    wmax = fmaxf(fmaxf(wx, wy), wz);
    // (the roundings of c -/+ w are below 2^-24 of |c| + w, inside the factor of two)
    return RayAffine{cx - wx, cy - wy, cz - wz, cx + wx, cy + wy, cz + wz};
}

template <bool ZERO_O, bool ANYHIT, int OCT>
__device__ __forceinline__ void traverse_fast_loop(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                   float ix, float iy, float iz, float margin, float& tnear, int& best_key,
                                                   int& best_leaf, Counters& cnt, float t2max, const RayAffine& R)
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
            float tminL, tmaxL, tminR, tmaxR;
            if (ZERO_O) {
                const float lx0 = q0.x * ix, ly0 = q0.y * iy, lz0 = q0.z * iz, lx1 = q0.w * ix, ly1 = q1.x * iy, lz1 = q1.y * iz;
                const float rx0 = q1.z * ix, ry0 = q1.w * iy, rz0 = q2.x * iz, rx1 = q2.y * ix, ry1 = q2.z * iy, rz1 = q2.w * iz;
                slab_interval<OCT>(lx0, ly0, lz0, lx1, ly1, lz1, tminL, tmaxL);
                slab_interval<OCT>(rx0, ry0, rz0, rx1, ry1, rz1, tminR, tmaxR);
                tmaxL = __fmaf_rn(fabsf(tmaxL), WIDE2, tmaxL);
                tmaxR = __fmaf_rn(fabsf(tmaxR), WIDE2, tmaxR);
            } else {
                // near / far plane of each axis by the octant; widening folded into the constants (RayAffine)
                const float lxn = (OCT & 1) ? q0.w : q0.x, lxf = (OCT & 1) ? q0.x : q0.w;
                const float lyn = (OCT & 2) ? q1.x : q0.y, lyf = (OCT & 2) ? q0.y : q1.x;
                const float lzn = (OCT & 4) ? q1.y : q0.z, lzf = (OCT & 4) ? q0.z : q1.y;
                const float rxn = (OCT & 1) ? q2.y : q1.z, rxf = (OCT & 1) ? q1.z : q2.y;
                const float ryn = (OCT & 2) ? q2.z : q1.w, ryf = (OCT & 2) ? q1.w : q2.z;
                const float rzn = (OCT & 4) ? q2.w : q2.x, rzf = (OCT & 4) ? q2.x : q2.w;
                tminL = fmaxf(fmaxf(__fmaf_rn(lxn, ix, R.cnx), __fmaf_rn(lyn, iy, R.cny)), __fmaf_rn(lzn, iz, R.cnz));
                tmaxL = fminf(fminf(__fmaf_rn(lxf, ix, R.cfx), __fmaf_rn(lyf, iy, R.cfy)), __fmaf_rn(lzf, iz, R.cfz));
                tminR = fmaxf(fmaxf(__fmaf_rn(rxn, ix, R.cnx), __fmaf_rn(ryn, iy, R.cny)), __fmaf_rn(rzn, iz, R.cnz));
                tmaxR = fminf(fminf(__fmaf_rn(rxf, ix, R.cfx), __fmaf_rn(ryf, iy, R.cfy)), __fmaf_rn(rzf, iz, R.cfz));
            }
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

// The same ordered traversal over the 4-wide collapse (Wide4): one visit = seven 16-byte loads, four boxes tested; the nearest accepted
// slot is descended, the others are pushed (unsorted: the CPU model shows no gain from sorting them). Half the dependent loads per
// ray of the binary walk at equal arithmetic. Leaves, pruning and the candidate rule are the binary loop's, so are the hits.
constexpr int WIDE_STACK = 96;
template <bool ZERO_O, bool ANYHIT, int OCT>
__device__ __forceinline__ void traverse_wide_loop(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                                   float ix, float iy, float iz, float margin, float& tnear, int& best_key,
                                                   int& best_leaf, Counters& cnt, float t2max, const RayAffine& R)
{
    const float neg_margin = -margin;
    float tlim = ANYHIT ? sqrtf(t2max) + margin : tnear + margin;
    tlim = __fmaf_rn(fabsf(tlim), WIDE2, tlim);
    int2 stack[WIDE_STACK];
    int sp = 0;
    int node = 0;
    unsigned visits = 0;
    while (true) {
        if (node >= 0) {
            const float4* q = reinterpret_cast<const float4*>(B.wide + node);
            const float4 X0 = __ldg(q), X1 = __ldg(q + 1), Y0 = __ldg(q + 2), Y1 = __ldg(q + 3), Z0 = __ldg(q + 4), Z1 = __ldg(q + 5);
            const int4 rf = __ldg(reinterpret_cast<const int4*>(q + 6));
            ++visits;
            const float4 xn = (OCT & 1) ? X1 : X0, xf = (OCT & 1) ? X0 : X1;
            const float4 yn = (OCT & 2) ? Y1 : Y0, yf = (OCT & 2) ? Y0 : Y1;
            const float4 zn = (OCT & 4) ? Z1 : Z0, zf = (OCT & 4) ? Z0 : Z1;
            const float xnv[4] = {xn.x, xn.y, xn.z, xn.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
            const float ynv[4] = {yn.x, yn.y, yn.z, yn.w}, yfv[4] = {yf.x, yf.y, yf.z, yf.w};
            const float znv[4] = {zn.x, zn.y, zn.z, zn.w}, zfv[4] = {zf.x, zf.y, zf.z, zf.w};
            const int ref[4] = {rf.x, rf.y, rf.z, rf.w};
            float tmn[4];
            bool hit[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float tmx;
                if (ZERO_O) {
                    tmn[k] = fmaxf(fmaxf(xnv[k] * ix, ynv[k] * iy), znv[k] * iz);
                    tmx = fminf(fminf(xfv[k] * ix, yfv[k] * iy), zfv[k] * iz);
                    tmx = __fmaf_rn(fabsf(tmx), WIDE2, tmx);
                } else {
                    tmn[k] = fmaxf(fmaxf(__fmaf_rn(xnv[k], ix, R.cnx), __fmaf_rn(ynv[k], iy, R.cny)), __fmaf_rn(znv[k], iz, R.cnz));
                    tmx = fminf(fminf(__fmaf_rn(xfv[k], ix, R.cfx), __fmaf_rn(yfv[k], iy, R.cfy)), __fmaf_rn(zfv[k], iz, R.cfz));
                }
                hit[k] = tmn[k] <= fminf(tmx, tlim) && tmx >= neg_margin;      // an unused slot has tmn = +inf
            }
            // nearest accepted slot first
            int first = -1;
            float tfirst = INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (hit[k] && (first < 0 || tmn[k] < tfirst)) { first = k; tfirst = tmn[k]; }
            if (first >= 0) {
                int next = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k == first) next = ref[k];
                    else if (hit[k]) { stack[sp] = make_int2(ref[k], __float_as_int(tmn[k])); sp = min(sp + 1, WIDE_STACK - 1); }
                }
                node = next;
                continue;
            }
        } else {
            const int leaf = ~node;
            bool pass;
            float t0, t1;
            if (B.leaf_box_prim) {
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
                    if (t0 * t0 < t2max) { tnear = t0; best_leaf = leaf; cnt.node_visits += visits; cnt.node_tests += 4 * visits; return; }
                } else {
                    candidate(t0, t1, B.tie_by_objid ? __ldg(B.prim_order + leaf) : leaf, leaf, tnear, best_key, best_leaf);
                    tlim = tnear + margin;
                    tlim = __fmaf_rn(fabsf(tlim), WIDE2, tlim);
                }
            }
        }
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
    cnt.node_tests += 4 * visits;
}

// ZNEG: the caller guarantees dz < 0 (every primary ray: dz = -1 before normalisation), four octants instead of eight.
// WIDE: walk the 4-wide collapse (the caller has checked B.wide).
template <bool ZERO_O, bool ANYHIT = false, bool ZNEG = false, bool WIDE = false>
__device__ __forceinline__ void traverse_fast(const BvhView& B, float ox, float oy, float oz, float dx, float dy, float dz,
                                              float& tnear, int& best_key, int& best_leaf, Counters& cnt, float t2max = 0.f)
{
    // 1/d only feeds the conservative tests (its error is inside the widening): one MUFU.RCP each instead of an IEEE divide
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(dz));
    const float amin = fminf(fminf(fabsf(ix), fabsf(iy)), fabsf(iz)), amax = fmaxf(fmaxf(fabsf(ix), fabsf(iy)), fabsf(iz));
    RayAffine R = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float wmax = 0.f;
    if (!ZERO_O) R = ray_affine(B.root_box, ox, oy, oz, ix, iy, iz, wmax);
    if (B.root_ref < 0 || !(amin > 1e-30f && amax < 1e30f) || !(wmax < 1e30f)) {
        // single-leaf tree, or a zero / tiny / huge / non-finite direction component (or products out of range): the divide-based traversal
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
#define RTDS_OCT_CASE(o) case o: if (WIDE) traverse_wide_loop<ZERO_O, ANYHIT, o>(B, ox, oy, oz, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt, t2max, R); \
                                else traverse_fast_loop<ZERO_O, ANYHIT, o>(B, ox, oy, oz, dx, dy, dz, ix, iy, iz, margin, tnear, best_key, best_leaf, cnt, t2max, R); break;
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
// HULL: interior boxes are tested ONCE for the whole packet against the interval hull of the four reciprocal directions
// (near planes times [min, max] reciprocal -> a lower bound of every ray's entry distance, far planes -> an upper bound of
// every ray's exit distance) instead of once per ray. Float multiplication is monotone, so the bounds enclose each ray's
// own products and the hull test accepts every box ANY ray's own test accepts: interior tests only steer (the candidate
// criterion is leaf-local), so hits are unchanged; what changes is the work - 24 multiplies per visit instead of 48, no
// per-ray selects - against a slightly larger set of visited nodes (the hull is the box grown by about one pixel
// footprint, the samples of one pixel being at most a pixel apart).
template <int OCT, bool HULL>
__device__ __forceinline__ void traverse_packet(const BvhView& B, const float (&dx)[PK], const float (&dy)[PK], const float (&dz)[PK],
                                                const float (&ix)[PK], const float (&iy)[PK], const float (&iz)[PK], float margin,
                                                float (&tnear)[PK], int (&best_key)[PK], int (&best_leaf)[PK], Counters& cnt)
{
    const float zthr = margin * (9.5367431640625e-7f / 0.00278f);      // prune_margin = 0.00278 * corner distance
    float tlim[PK];
#pragma unroll
    for (int j = 0; j < PK; ++j) { tlim[j] = tnear[j] + margin; tlim[j] = __fmaf_rn(fabsf(tlim[j]), WIDE2, tlim[j]); }
    float tlim_max = fmaxf(fmaxf(tlim[0], tlim[1]), fmaxf(tlim[2], tlim[3]));
    // the packet's reciprocal-direction intervals (all four rays share the octant, so each interval has one sign)
    const float ixlo = fminf(fminf(ix[0], ix[1]), fminf(ix[2], ix[3])), ixhi = fmaxf(fmaxf(ix[0], ix[1]), fmaxf(ix[2], ix[3]));
    const float iylo = fminf(fminf(iy[0], iy[1]), fminf(iy[2], iy[3])), iyhi = fmaxf(fmaxf(iy[0], iy[1]), fmaxf(iy[2], iy[3]));
    const float izlo = fminf(fminf(iz[0], iz[1]), fminf(iz[2], iz[3])), izhi = fmaxf(fmaxf(iz[0], iz[1]), fmaxf(iz[2], iz[3]));
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
            float tL, tR;
            if (HULL) {
                // near / far plane of each axis as the octant selects them; lower bound of the entry distances, upper bound
                // of the exit distances over the packet
                const float lnx = (OCT & 1) ? q0.w : q0.x, lfx = (OCT & 1) ? q0.x : q0.w;
                const float lny = (OCT & 2) ? q1.x : q0.y, lfy = (OCT & 2) ? q0.y : q1.x;
                const float lnz = q1.y, lfz = q0.z;                                        // dz < 0 for the whole packet
                const float rnx = (OCT & 1) ? q2.y : q1.z, rfx = (OCT & 1) ? q1.z : q2.y;
                const float rny = (OCT & 2) ? q2.z : q1.w, rfy = (OCT & 2) ? q1.w : q2.z;
                const float rnz = q2.w, rfz = q2.x;
                const float tminL = fmaxf(fmaxf(fminf(lnx * ixlo, lnx * ixhi), fminf(lny * iylo, lny * iyhi)), fminf(lnz * izlo, lnz * izhi));
                float tmaxL = fminf(fminf(fmaxf(lfx * ixlo, lfx * ixhi), fmaxf(lfy * iylo, lfy * iyhi)), fmaxf(lfz * izlo, lfz * izhi));
                const float tminR = fmaxf(fmaxf(fminf(rnx * ixlo, rnx * ixhi), fminf(rny * iylo, rny * iyhi)), fminf(rnz * izlo, rnz * izhi));
                float tmaxR = fminf(fminf(fmaxf(rfx * ixlo, rfx * ixhi), fmaxf(rfy * iylo, rfy * iyhi)), fmaxf(rfz * izlo, rfz * izhi));
                tmaxL = __fmaf_rn(fabsf(tmaxL), WIDE2, tmaxL);
                tmaxR = __fmaf_rn(fabsf(tmaxR), WIDE2, tmaxR);
                tL = (frontL && tminL <= fminf(tmaxL, tlim_max)) ? tminL : INFINITY;
                tR = (frontR && tminR <= fminf(tmaxR, tlim_max)) ? tminR : INFINITY;
            } else {
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
            tL = frontL ? fminf(fminf(fminf(kL[0], kL[1]), kL[2]), kL[3]) : INFINITY;
            tR = frontR ? fminf(fminf(fminf(kR[0], kR[1]), kR[2]), kR[3]) : INFINITY;
            }
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
    cnt.node_tests += (HULL ? 2 : 2 * PK) * visits;      // slab tests executed: one per child and packet with HULL
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
                else { float t; h = tri_test_any(type, ox, oy, oz, dx, dy, dz, sh[3 * i], sh[3 * i + 1], sh[3 * i + 2], t); t0 = t1 = t; }
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

// Closest-hit KD traversal (extension, PARITY UNPINNED by the reference: its KD path is any-hit and unshaded,
// main.cpp:362-372). Same walk as kdtreeIntersect above (accelerators.h:1012-1083: plane distance, belowFirst rule,
// 64-entry todo stack) with the two changes PBRT's KdTreeAccel::Intersect makes to its IntersectP, which the reference
// ported: a hit shortens the ray instead of ending the walk, and the walk ends when the current cell starts beyond the
// nearest hit (`tnear < tMin`). Candidates are reduced like main.cpp:379-384 (t0 < 0 -> t1, strict <) with ties broken
// towards the smaller objId = the NONE loop's "first candidate wins", so hit ids equal the brute-force path's except for
// rays whose nearest hit lies (in float) outside every cell the primitive overlaps.
__device__ __forceinline__ void kd_closest_hit(const KdView& K, float ox, float oy, float oz, float dx, float dy, float dz, float& tnear,
                                               int& hit_obj, Counters& cnt)
{
    float tMin, tMax;
    cnt.node_tests++;
    if (!slab_test(ox, oy, oz, dx, dy, dz, K.bounds[0], K.bounds[1], K.bounds[2], K.bounds[3], K.bounds[4], K.bounds[5], tMin, tMax))
        return;
    const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    const float inv[3] = {1 / dx, 1 / dy, 1 / dz};
    int   todo_node[64];
    float todo_tmin[64], todo_tmax[64];
    int todoPos = 0;
    int node = 0;
    int best_key = 0;
    while (true) {
        if (tnear < tMin) break;
        const rtds_kd_node nd = K.nodes[node];
        cnt.node_visits++;
        if ((nd.w1 & 3u) == 3u) {
            const int np = (int)nd.w2;
            for (int i = 0; i < np; ++i) {
                const int prim = np == 1 ? (int)nd.w0 : __ldg(K.prim_idx + (int)nd.w0 + i);
                float t0, t1;
                cnt.prim_tests++;
                if (obj_test(K.prim_type, K.sph, K.tri, prim, ox, oy, oz, dx, dy, dz, t0, t1)) candidate(t0, t1, prim, prim, tnear, best_key, hit_obj);
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
}

}  // namespace
