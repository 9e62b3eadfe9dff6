// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// oracle/oracle.cpp: CPU restatement ("port") of the reference's hot path, written from the reference's
// behaviour, each function citing the reference file:line it follows (paths relative to
// /root/reference/project/raytracer/).  It is the checker for the CUDA path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (librtds.so) never links or calls it.
//
// Pinned: tests/test_oracle_vs_reference.py checks every function below against the compiled, unmodified
// reference (oracle/_ref/libref_oracle.so, built from the sources where they lie) and against the golden
// vectors the reference ships (output.ppm md5, node and test counts, Morton KATs) — see tests/golden/.
//
// Third-party arithmetic the reference's results depend on, and how it is restated here:
//   * libstdc++ 13.3 std::partition / std::nth_element (call sites accelerators.h:302,313,511,522) decide the
//     BVH's primitive order under ties.  The oracle calls the same library algorithms on compact records —
//     the permutation they produce depends only on the comparison outcomes, not on the element type.
//   * libstdc++ std::mt19937 + generate_canonical<double,53> (main.cpp:503-508): restated from the published
//     MT19937 algorithm (Matsumoto & Nishimura 1998) and [rand.util.canonical].
//   * glibc powf(x,25) (main.cpp:475): restated as an exact-product chain in double rounded once.
// Build: oracle/Makefile -> oracle/liboracle.so   (g++ -O2 -ffp-contract=off: no FMA contraction, like the
// reference's x86-64 baseline build).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace {

struct V3 { float x, y, z; };
inline float comp(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

struct Box { V3 mn, mx; };
inline Box empty_box()  // BoxBoundries() accelerators.h:108-112
{
    const float M = std::numeric_limits<float>::max();
    return Box{{M, M, M}, {-M, -M, -M}};
}
inline Box join(const Box& a, const Box& b)  // JoinBounds accelerators.h:201-214
{
    return Box{{std::min(a.mn.x, b.mn.x), std::min(a.mn.y, b.mn.y), std::min(a.mn.z, b.mn.z)},
               {std::max(a.mx.x, b.mx.x), std::max(a.mx.y, b.mx.y), std::max(a.mx.z, b.mx.z)}};
}
inline int max_axis(const Box& b)  // GeMaximumAxis accelerators.h:190-199
{
    float ex = b.mx.x - b.mn.x, ey = b.mx.y - b.mn.y, ez = b.mx.z - b.mn.z;
    if (ex > ey && ex > ez) return 0;
    else if (ey > ez) return 1;
    else return 2;
}

struct Prim { V3 c; float r; int id; Box box; };

struct LinearNode {  // accelerators.h:231-240
    float bmin[3], bmax[3];
    int32_t offset;
    uint16_t nPrimitives;
    uint8_t axis, pad;
};
static_assert(sizeof(LinearNode) == 32, "LinearBVHNode is 32 bytes");

Box prim_box(const float* s)  // main.cpp:686-688: centre -/+ radius in float
{
    return Box{{s[0] - s[3], s[1] - s[3], s[2] - s[3]}, {s[0] + s[3], s[1] + s[3], s[2] + s[3]}};
}

// Primitive table seen by the builders: type 0 = spheres (n x {cx,cy,cz,r}), type 1 = triangles (n x 9, extension:
// box = min/max of the vertices, centre = (min + max) * 0.5f).
struct Prims {
    const float* data; int type;
    Box box(int i) const
    {
        if (type == 0) return prim_box(data + 4 * (size_t)i);
        const float* t = data + 9 * (size_t)i;
        Box b;
        b.mn = V3{std::min(std::min(t[0], t[3]), t[6]), std::min(std::min(t[1], t[4]), t[7]), std::min(std::min(t[2], t[5]), t[8])};
        b.mx = V3{std::max(std::max(t[0], t[3]), t[6]), std::max(std::max(t[1], t[4]), t[7]), std::max(std::max(t[2], t[5]), t[8])};
        return b;
    }
    V3 centre(int i) const
    {
        if (type == 0) { const float* q = data + 4 * (size_t)i; return V3{q[0], q[1], q[2]}; }
        Box b = box(i);
        return V3{(b.mn.x + b.mx.x) * 0.5f, (b.mn.y + b.mx.y) * 0.5f, (b.mn.z + b.mx.z) * 0.5f};
    }
};

// ----------------------------------------------------------------------------------------------
// constructBVHNew, accelerators.h:246-337, emitting the pre-order LinearBVHNode array directly.
// Returns 0, or -6 when std::partition returns endIndex (the reference then recurses forever).
// ----------------------------------------------------------------------------------------------
struct MedianBuilder {
    std::vector<Prim>& P;
    std::vector<LinearNode>& out;
    int max_depth = 0;
    int status = 0;
    MedianBuilder(std::vector<Prim>& p, std::vector<LinearNode>& o) : P(p), out(o) {}

    void leaf(int my, int start)
    {
        out[my].offset = start; out[my].nPrimitives = 1; out[my].axis = 0; out[my].pad = 0;
    }
    int build(int start, int end, int depth)
    {
        if (status) return -1;
        int my = (int)out.size();
        out.push_back(LinearNode());
        max_depth = std::max(max_depth, depth);
        Box b = empty_box();
        for (int i = start; i < end; ++i) b = join(b, P[i].box);  // :257-260
        out[my].bmin[0] = b.mn.x; out[my].bmin[1] = b.mn.y; out[my].bmin[2] = b.mn.z;
        out[my].bmax[0] = b.mx.x; out[my].bmax[1] = b.mx.y; out[my].bmax[2] = b.mx.z;
        int n = end - start;
        if (n <= 1) { leaf(my, start); return my; }  // :265-271
        // centroidBounds is seeded with `bounds` (:275) and only grows by points inside it: == bounds
        int dim = max_axis(b);
        float lo = comp(b.mn, dim), hi = comp(b.mx, dim);
        if (hi == lo) { status = -8; return my; }  // :286-293 zero-extent range (reference pushes wrong ids); unsupported
        float pmid = (lo + hi) / 2;                // :298
        Prim* first = &P[start];
        Prim* last = &P[end - 1] + 1;
        Prim* midp = std::partition(first, last, [dim, pmid](const Prim& p) { return comp(p.c, dim) < pmid; });  // :302-306
        int mid = (int)(midp - &P[0]);
        if (mid != start && mid != end) {          // :311-319
            mid = (start + end) / 2;
            std::nth_element(first, &P[mid], last, [dim](const Prim& a, const Prim& c) { return comp(a.c, dim) < comp(c.c, dim); });
        }
        if (start == mid) { leaf(my, start); return my; }  // :321-327 (the rest of the range is dropped)
        if (mid == end) { status = -6; return my; }        // :329 recurses on the same range forever
        build(start, mid, depth + 1);
        int second = build(mid, end, depth + 1);
        out[my].offset = second; out[my].nPrimitives = 0; out[my].axis = (uint8_t)dim; out[my].pad = 0;  // :333-334
        return my;
    }
};

// ----------------------------------------------------------------------------------------------
// Morton (accelerators.h:374-394)
// ----------------------------------------------------------------------------------------------
inline unsigned expandBits(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
inline unsigned morton3D(float x, float y, float z)
{
    x = std::min(std::max(x * 1024.0f, 0.0f), 1023.0f);
    y = std::min(std::max(y * 1024.0f, 0.0f), 1023.0f);
    z = std::min(std::max(z * 1024.0f, 0.0f), 1023.0f);
    unsigned xx = expandBits((unsigned)x), yy = expandBits((unsigned)y), zz = expandBits((unsigned)z);
    return xx * 4 + yy * 2 + zz;
}
inline uint64_t spread21(uint64_t v)  // bit i -> bit 3i, one bit at a time (independent of the kernel's magic masks)
{
    uint64_t r = 0;
    for (int i = 0; i < 21; ++i) r |= ((v >> i) & 1ull) << (3 * i);
    return r;
}
inline uint64_t morton63(float x, float y, float z)
{
    x = std::min(std::max(x * 2097152.0f, 0.0f), 2097151.0f);
    y = std::min(std::max(y * 2097152.0f, 0.0f), 2097151.0f);
    z = std::min(std::max(z * 2097152.0f, 0.0f), 2097151.0f);
    return spread21((uint64_t)x) * 4 + spread21((uint64_t)y) * 2 + spread21((uint64_t)z);
}

// Top-down LBVH over sorted keys: split where the highest differing bit of the (key, position) pair changes —
// the tree Karras' parallel construction emits (NVIDIA "Thinking Parallel III", cited at accelerators.h:371,568;
// findSplit :397-449 is the reference's 8-bit attempt at it).
struct LbvhBuilder {
    const std::vector<uint64_t>& K;  // sorted keys
    const std::vector<Box>& LB;      // leaf boxes in sorted order
    std::vector<LinearNode>& out;
    int key_bits;
    int max_depth = 0;
    LbvhBuilder(const std::vector<uint64_t>& k, const std::vector<Box>& lb, std::vector<LinearNode>& o, int kb) : K(k), LB(lb), out(o), key_bits(kb) {}

    // number of leading bits shared by elements i and j of the augmented key (key, position)
    int common(int i, int j) const
    {
        if (K[i] != K[j]) return __builtin_clzll(K[i] ^ K[j]) - (64 - key_bits);
        return key_bits + __builtin_clz((unsigned)(i ^ j));
    }
    Box build(int first, int last, int depth)
    {
        int my = (int)out.size();
        out.push_back(LinearNode());
        max_depth = std::max(max_depth, depth);
        Box b;
        if (first == last) {
            b = LB[first];
            out[my].offset = first; out[my].nPrimitives = 1; out[my].axis = 0; out[my].pad = 0;
        } else {
            int cp = common(first, last);
            // last position in [first,last) that shares more than cp bits with `first`
            int lo = first, hi = last;  // invariant: common(first,lo) > cp (or lo==first), common(first,hi) == cp
            while (hi - lo > 1) {
                int m = (lo + hi) / 2;
                if (common(first, m) > cp) lo = m; else hi = m;
            }
            int split = lo;
            Box l = build(first, split, depth + 1);
            int second = (int)out.size();
            Box r = build(split + 1, last, depth + 1);
            b = join(l, r);
            int axis = 0;
            if (cp < key_bits) { int bit = key_bits - 1 - cp; axis = 2 - (bit % 3); }
            // key_bits 30 and 63 are multiples of 3, so bit%3 is the same counted in the 32/64-bit container
            out[my].offset = second; out[my].nPrimitives = 0; out[my].axis = (uint8_t)axis; out[my].pad = 0;
        }
        out[my].bmin[0] = b.mn.x; out[my].bmin[1] = b.mn.y; out[my].bmin[2] = b.mn.z;
        out[my].bmax[0] = b.mx.x; out[my].bmax[1] = b.mx.y; out[my].bmax[2] = b.mx.z;
        return b;
    }
};

// ----------------------------------------------------------------------------------------------
// Binned-SAH BVH (extension; the reference has no SAH BVH) — sequential statement of the definition in
// raytracer-data-structures_b200/csrc/sah.cu. PARITY UNPINNED by the reference; this is the oracle for K7.
// ----------------------------------------------------------------------------------------------
struct SahBuilder {
    Prims sph; std::vector<int>& order; std::vector<LinearNode>& out; int B; int max_depth = 0;
    SahBuilder(Prims s, std::vector<int>& o, std::vector<LinearNode>& n, int b) : sph(s), order(o), out(n), B(b) {}
    static float area(const Box& b)   // accelerators.h:122-125
    {
        float dx = b.mx.x - b.mn.x, dy = b.mx.y - b.mn.y, dz = b.mx.z - b.mn.z;
        return 2 * (dx * dy + dx * dz + dy * dz);
    }
    Box build(int s, int e, int depth)
    {
        int my = (int)out.size();
        out.push_back(LinearNode());
        max_depth = std::max(max_depth, depth);
        Box box;
        if (e - s == 1) {
            box = sph.box(order[s]);
            out[my].offset = s; out[my].nPrimitives = 1; out[my].axis = 0; out[my].pad = 0;
        } else {
            const float INF = INFINITY;
            float cmn[3] = {INF, INF, INF}, cmx[3] = {-INF, -INF, -INF};
            for (int p = s; p < e; ++p)
                { V3 cc = sph.centre(order[p]); for (int a = 0; a < 3; ++a) { float c = comp(cc, a); cmn[a] = std::min(cmn[a], c); cmx[a] = std::max(cmx[a], c); } }
            float ex = cmx[0] - cmn[0], ey = cmx[1] - cmn[1], ez = cmx[2] - cmn[2];
            int axis = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
            float lo = cmn[axis], hi = cmx[axis];
            std::vector<int> bin(e - s);
            std::vector<unsigned> cnt(B, 0);
            std::vector<Box> bb(B, Box{{INF, INF, INF}, {-INF, -INF, -INF}});
            for (int p = s; p < e; ++p) {
                int b = 0;
                if (hi > lo) { b = (int)((float)B * ((comp(sph.centre(order[p]), axis) - lo) / (hi - lo))); if (b > B - 1) b = B - 1; }
                bin[p - s] = b;
                cnt[b]++;
                bb[b] = join(bb[b], sph.box(order[p]));
            }
            std::vector<Box> rb(B); std::vector<unsigned> rc(B, 0);
            Box acc{{INF, INF, INF}, {-INF, -INF, -INF}}; unsigned n_acc = 0;
            for (int b = B - 1; b >= 1; --b) { if (cnt[b]) acc = join(acc, bb[b]); n_acc += cnt[b]; rb[b] = acc; rc[b] = n_acc; }
            Box lb{{INF, INF, INF}, {-INF, -INF, -INF}}; unsigned nL = 0;
            float best = INF; int best_i = -1, best_nL = 0;
            for (int i = 0; i < B - 1; ++i) {
                if (cnt[i]) lb = join(lb, bb[i]);
                nL += cnt[i];
                unsigned nR = rc[i + 1];
                if (nL == 0 || nR == 0) continue;
                float cost = (float)nL * area(lb) + (float)nR * area(rb[i + 1]);
                if (cost < best) { best = cost; best_i = i; best_nL = (int)nL; }
            }
            int mid;
            if (best_i < 0) mid = s + (e - s) / 2;
            else {
                std::vector<int> L, R;
                for (int p = s; p < e; ++p) (bin[p - s] <= best_i ? L : R).push_back(order[p]);
                std::copy(L.begin(), L.end(), order.begin() + s);
                std::copy(R.begin(), R.end(), order.begin() + s + L.size());
                mid = s + best_nL;
            }
            Box l = build(s, mid, depth + 1);
            int second = (int)out.size();
            Box r = build(mid, e, depth + 1);
            box = join(l, r);
            out[my].offset = second; out[my].nPrimitives = 0; out[my].axis = (uint8_t)axis; out[my].pad = 0;
        }
        out[my].bmin[0] = box.mn.x; out[my].bmin[1] = box.mn.y; out[my].bmin[2] = box.mn.z;
        out[my].bmax[0] = box.mx.x; out[my].bmax[1] = box.mx.y; out[my].bmax[2] = box.mx.z;
        return box;
    }
};

// ----------------------------------------------------------------------------------------------
// traversal + intersection
// ----------------------------------------------------------------------------------------------
// boundingBoxIntersection accelerators.h:588-626
inline bool slab(const float o[3], const float d[3], const float* bmin, const float* bmax)
{
    float tmin = (bmin[0] - o[0]) / d[0];
    float tmax = (bmax[0] - o[0]) / d[0];
    if (tmin > tmax) std::swap(tmin, tmax);
    float tymin = (bmin[1] - o[1]) / d[1];
    float tymax = (bmax[1] - o[1]) / d[1];
    if (tymin > tymax) std::swap(tymin, tymax);
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (bmin[2] - o[2]) / d[2];
    float tzmax = (bmax[2] - o[2]) / d[2];
    if (tzmin > tzmax) std::swap(tzmin, tzmax);
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    return true;
}

// Sphere::raySphereIntersect accelerators.h:79-92
inline bool ray_sphere(const float o[3], const float d[3], const float* s /*cx,cy,cz,r*/, float& t0, float& t1)
{
    float lx = s[0] - o[0], ly = s[1] - o[1], lz = s[2] - o[2];
    float tca = lx * d[0] + ly * d[1] + lz * d[2];
    if (tca < 0) return false;
    float d2 = (lx * lx + ly * ly + lz * lz) - tca * tca;
    float radius2 = s[3] * s[3];
    if (d2 > radius2) return false;
    float thc = std::sqrt(radius2 - d2);
    t0 = tca - thc;
    t1 = tca + thc;
    return true;
}

// Moller-Trumbore as in the reference's never-compiled MOLLER_TRUMBORE branch (main.cpp:138-162, `v_0` read as v0,
// no culling, EPS 1e-6) + the t < 0 rejection of the geometric branch (main.cpp:184). Extension: PARITY UNPINNED.
inline bool ray_triangle(const float o[3], const float d[3], const float* v /*9 floats*/, float& t)
{
    const float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
    const float e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
    const float px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    if (std::fabs(det) < 1e-6f) return false;
    const float inv = 1 / det;
    const float tx = o[0] - v[0], ty = o[1] - v[1], tz = o[2] - v[2];
    const float u = (tx * px + ty * py + tz * pz) * inv;
    if (u < 0 || u > 1) return false;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const float w = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
    if (w < 0 || u + w > 1) return false;
    t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    return !(t < 0);
}

// Triangle::rayTriangleIntersect as the reference COMPILES it (main.cpp:163-215, the geometric branch; MOLLER_TRUMBORE is never
// defined): plane hit with N = v0v1 x v0v2, t = (N.orig + N.v0) / N.dir - correct for orig = 0 only, kept bug for bug -, then the
// three inside-outside edge tests. PINNED: tests/test_oracle_vs_reference.py checks it against the reference's own class Triangle.
inline bool ray_triangle_geometric(const float o[3], const float d[3], const float* v /*9 floats*/, float& t)
{
    const float ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2];            // v0v1
    const float bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];            // v0v2
    const float Nx = ay * bz - az * by, Ny = az * bx - ax * bz, Nz = ax * by - ay * bx;
    const float nd = Nx * d[0] + Ny * d[1] + Nz * d[2];
    if (std::fabs(nd) < 1e-6f) return false;                                     // EPS, main.cpp:59,175
    const float dd = Nx * v[0] + Ny * v[1] + Nz * v[2];
    t = ((Nx * o[0] + Ny * o[1] + Nz * o[2]) + dd) / nd;                         // main.cpp:182
    if (t < 0) return false;
    const float Px = o[0] + d[0] * t, Py = o[1] + d[1] * t, Pz = o[2] + d[2] * t;
    {   // edge 0
        const float px = Px - v[0], py = Py - v[1], pz = Pz - v[2];
        const float cx = ay * pz - az * py, cy = az * px - ax * pz, cz = ax * py - ay * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    {   // edge 1
        const float ex = v[6] - v[3], ey = v[7] - v[4], ez = v[8] - v[5];
        const float px = Px - v[3], py = Py - v[4], pz = Pz - v[5];
        const float cx = ey * pz - ez * py, cy = ez * px - ex * pz, cz = ex * py - ey * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    {   // edge 2
        const float ex = v[0] - v[6], ey = v[1] - v[7], ez = v[2] - v[8];
        const float px = Px - v[6], py = Py - v[7], pz = Pz - v[8];
        const float cx = ey * pz - ez * py, cy = ez * px - ex * pz, cz = ex * py - ey * px;
        if (Nx * cx + Ny * cy + Nz * cz < 0) return false;
    }
    return true;
}

struct Scene {
    const float* sph; const float* mat; int n;
    const LinearNode* nodes; const int* prim_order; int n_nodes;
    int tie_by_objid;
    int prim_type = 0;   // 0: sph = n x 4 spheres; 1: sph = n x 9 triangles, Moller-Trumbore; 2: triangles, the reference's compiled geometric test
    bool test(const float o[3], const float d[3], int obj, float& t0, float& t1) const
    {
        if (prim_type == 0) return ray_sphere(o, d, sph + 4 * (size_t)obj, t0, t1);
        float t;
        if (!(prim_type == 2 ? ray_triangle_geometric(o, d, sph + 9 * (size_t)obj, t) : ray_triangle(o, d, sph + 9 * (size_t)obj, t))) return false;
        t0 = t1 = t;
        return true;
    }
};

// boxIntersect accelerators.h:668-690 + candidate loop main.cpp:343-358 over a flattened tree.
// tie_by_objid: candidates are visited in objId order instead of DFS order (equal-t ties only).
void closest_bvh(const Scene& S, const float o[3], const float d[3], int& hit, float& tnear, long long* cand)
{
    hit = -1; tnear = INFINITY;
    int best_key = 0;
    int stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        int ni = stack[--sp];
        const LinearNode& nd = S.nodes[ni];
        if (!slab(o, d, nd.bmin, nd.bmax)) continue;
        if (nd.nPrimitives) {
            int leafpos = nd.offset, obj = S.prim_order[leafpos];
            if (cand) ++*cand;
            float t0 = INFINITY, t1 = INFINITY;
            if (S.test(o, d, obj, t0, t1)) {
                if (t0 < 0) t0 = t1;
                int key = S.tie_by_objid ? obj : leafpos;
                if (t0 < tnear || (t0 == tnear && hit >= 0 && key < best_key)) { tnear = t0; hit = obj; best_key = key; }
            }
        } else {
            stack[sp++] = nd.offset;   // right child popped after the left subtree: DFS left -> right
            stack[sp++] = ni + 1;
        }
    }
}

// main.cpp:376-386
void closest_none(const Scene& S, const float o[3], const float d[3], int& hit, float& tnear)
{
    hit = -1; tnear = INFINITY;
    for (int i = 0; i < S.n; ++i) {
        float t0 = INFINITY, t1 = INFINITY;
        if (S.test(o, d, i, t0, t1)) {
            if (t0 < 0) t0 = t1;
            if (t0 < tnear) { tnear = t0; hit = i; }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// shading (castRay's DIFFUSE_AND_GLOSSY branch, main.cpp:394-497)
// ----------------------------------------------------------------------------------------------
inline void normalize(float v[3])  // geometry.h:125-134: factor = 1 / sqrt(n) evaluated in double
{
    float n = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (n > 0) {
        float factor = (float)(1 / std::sqrt((double)n));
        v[0] *= factor; v[1] *= factor; v[2] *= factor;
    }
}
inline float pow25(float xf)
{
    double x = xf, x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
    return (float)(x16 * x8 * x);
}
struct Light { float c[3]; float radius; float le[3]; };

// ----------------------------------------------------------------------------------------------
// castRay in full (main.cpp:291-500): material branches, recursion to depth 2, and the shadow query.
// The reference's trace_more (main.cpp:235-245) is a stub that returns false; with `shadows` the contract its
// call site states (main.cpp:468-473) is evaluated instead: nearest hit of the shadow ray, in shadow iff
// tNearShadow^2 < lightDistance2.  Shadows have no executable reference behaviour: PARITY UNPINNED for that flag.
// ----------------------------------------------------------------------------------------------
inline float clampf(float lo, float hi, float v) { return std::max(lo, std::min(hi, v)); }   // main.cpp:75-79

void fresnel(const float I[3], const float N[3], float ior, float& kr)   // main.cpp:85-103
{
    float cosi = clampf(-1, 1, I[0] * N[0] + I[1] * N[1] + I[2] * N[2]);
    float etai = 1, etat = ior;
    if (cosi > 0) std::swap(etai, etat);
    float sint = etai / etat * sqrtf(std::max(0.f, 1 - cosi * cosi));
    if (sint >= 1) kr = 1;
    else {
        float cost = sqrtf(std::max(0.f, 1 - sint * sint));
        cosi = fabsf(cosi);
        float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
        float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
        kr = (Rs * Rs + Rp * Rp) / 2;
    }
}

void refract(const float I[3], const float N[3], float ior, float out[3])   // main.cpp:223-233
{
    float cosi = clampf(-1, 1, I[0] * N[0] + I[1] * N[1] + I[2] * N[2]);
    float etai = 1, etat = ior;
    float n[3] = {N[0], N[1], N[2]};
    if (cosi < 0) cosi = -cosi;
    else { std::swap(etai, etat); n[0] = -N[0]; n[1] = -N[1]; n[2] = -N[2]; }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) { out[0] = out[1] = out[2] = 0; return; }
    float f = eta * cosi - sqrtf(k);
    for (int c = 0; c < 3; ++c) out[c] = I[c] * eta + n[c] * f;
}

struct CastCtx {
    const Scene* S; const Light* L; int nl; int shadows;
    long long rays = 0, shadow_rays = 0, secondary_rays = 0;
};

void closest(const Scene& S, const float o[3], const float d[3], int& hit, float& tnear)
{
    if (S.nodes) closest_bvh(S, o, d, hit, tnear, nullptr); else closest_none(S, o, d, hit, tnear);
}

void cast_ray(CastCtx& C, const float o[3], const float d[3], int depth, float rgb[3], int* hit_out)
{
    const Scene& S = *C.S;
    if (depth > 2) { rgb[0] = 0.6f; rgb[1] = 0.8f; rgb[2] = 1.0f; return; }                  // :311-313
    int hit; float tnear;
    closest(S, o, d, hit, tnear);
    C.rays++;
    if (hit_out) *hit_out = hit;
    if (hit < 0) { rgb[0] = 0.6f; rgb[1] = 0.8f; rgb[2] = 1.0f; return; }                     // :318,394
    const float* mat = S.mat + 4 * hit;
    float hp[3] = {o[0] + d[0] * tnear, o[1] + d[1] * tnear, o[2] + d[2] * tnear};
    float N[3];
    if (S.prim_type == 0) {
        const float* sph = S.sph + 4 * (size_t)hit;
        N[0] = hp[0] - sph[0]; N[1] = hp[1] - sph[1]; N[2] = hp[2] - sph[2];
    } else {   // main.cpp:165-168: N = v0v1 x v0v2
        const float* v = S.sph + 9 * (size_t)hit;
        float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2], e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
        N[0] = e1y * e2z - e1z * e2y; N[1] = e1z * e2x - e1x * e2z; N[2] = e1x * e2y - e1y * e2x;
    }
    normalize(N);
    if (d[0] * N[0] + d[1] * N[1] + d[2] * N[2] > 0) { N[0] = -N[0]; N[1] = -N[1]; N[2] = -N[2]; }
    const float bias = 1e-4;                                                                  // :404
    const int material = (int)mat[3];
    if (material == 1) {                                                                      // REFLECTION_AND_REFRACTION :418-434
        float rd[3];
        refract(d, N, 3, rd);
        normalize(rd);
        float ro[3];
        bool neg = rd[0] * N[0] + rd[1] * N[1] + rd[2] * N[2] < 0;
        for (int c = 0; c < 3; ++c) ro[c] = neg ? hp[c] - N[c] * bias : hp[c] + N[c] * bias;
        // the reflection ray of :428 is traced by the reference but its colour is never used; it has no effect
        float refr[3];
        if (depth + 1 <= 2) C.secondary_rays++;   // a deeper ray returns the sky untraced (:311)
        cast_ray(C, ro, rd, depth + 1, refr, nullptr);
        float kr;
        fresnel(d, N, 2, kr);
        for (int c = 0; c < 3; ++c) rgb[c] = refr[c] * (1 - kr);                              // :432
        return;
    }
    if (material == 2) {                                                                      // REFLECTION :435-446
        float kr;
        fresnel(d, N, 2, kr);
        rgb[0] = rgb[1] = rgb[2] = (1 - kr);                                                  // :445
        return;
    }
    float hc[3] = {0, 0, 0};
    const float diff[3] = {0.815f, 0.235f, 0.031f};
    for (int i = 0; i < C.nl; ++i) {
        const Light& Lt = C.L[i];
        float ld[3] = {Lt.c[0] - hp[0], Lt.c[1] - hp[1], Lt.c[2] - hp[2]};
        float dist2 = ld[0] * ld[0] + ld[1] * ld[1] + ld[2] * ld[2];                          // :465
        normalize(ld);
        float LdotN = std::max(0.f, ld[0] * N[0] + ld[1] * N[1] + ld[2] * N[2]);
        int inShadow = 0;
        if (C.shadows) {
            bool front = d[0] * N[0] + d[1] * N[1] + d[2] * N[2] < 0;                         // :455-457
            float so[3];
            for (int c = 0; c < 3; ++c) so[c] = front ? hp[c] + N[c] * bias : hp[c] - N[c] * bias;
            int sh; float ts;
            closest(S, so, ld, sh, ts);
            C.shadow_rays++; C.rays++;
            inShadow = (sh >= 0 && ts * ts < dist2) ? 1 : 0;                                  // :471-472
        }
        float I[3] = {-ld[0], -ld[1], -ld[2]};
        float s2 = 2 * (I[0] * N[0] + I[1] * N[1] + I[2] * N[2]);
        float R[3] = {I[0] - N[0] * s2, I[1] - N[1] * s2, I[2] - N[2] * s2};
        float sp = pow25(std::max(0.f, -(R[0] * d[0] + R[1] * d[1] + R[2] * d[2])));
        for (int c = 0; c < 3; ++c) {
            float amt = (Lt.le[c] * (float)(1 - inShadow)) * LdotN;                            // :473
            float spec = Lt.le[c] * sp;
            hc[c] += (amt * (diff[c] * 0.8f)) / 2.0f + spec * 0.5f;
            hc[c] += mat[c];
        }
    }
    rgb[0] = hc[0]; rgb[1] = hc[1]; rgb[2] = hc[2];
}

// ----------------------------------------------------------------------------------------------
// MT19937 + generate_canonical<double,53> (main.cpp:503-508)
// ----------------------------------------------------------------------------------------------
struct MT {
    uint32_t s[624]; int p;
    explicit MT(uint32_t seed = 5489u)
    {
        s[0] = seed;
        for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        p = 624;
    }
    void regen()
    {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
            s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        p = 0;
    }
    uint32_t next()
    {
        if (p >= 624) regen();
        uint32_t y = s[p++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    void discard(unsigned long long n) { while (n--) next(); }
    double canonical()
    {
        double lo = (double)next(), hi = (double)next();
        double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = std::nextafter(1.0, 0.0);
        return r;
    }
};

}  // namespace

extern "C" {

// createScene_new's arithmetic (main.cpp:650-717) on already-parsed vertices. out: (nv*clones + 1) x 4.
int orc_scene_from_vertices(const float* v, int nv, int clones, float* cxyz_r, float* rgb_mat)
{
    int id = 0;
    for (int clone = 0; clone < clones; ++clone) {
        float shift = clone * 20;                      // :652
        for (int i = 0; i < nv; ++i, ++id) {
            float cx = v[3 * i] * 100 + shift, cy = v[3 * i + 1] * 100 + shift, cz = v[3 * i + 2] * 100 + shift;  // :680
            cy += -10; cz += -60;                      // :681-682
            cxyz_r[4 * id] = cx; cxyz_r[4 * id + 1] = cy; cxyz_r[4 * id + 2] = cz;
            cxyz_r[4 * id + 3] = (float)(0.01 * 5);    // :679
            rgb_mat[4 * id] = 0.8f; rgb_mat[4 * id + 1] = 0.7f; rgb_mat[4 * id + 2] = 0.f; rgb_mat[4 * id + 3] = 0.f;  // :689
        }
    }
    cxyz_r[4 * id] = 0.93591022f; cxyz_r[4 * id + 1] = -105.47120094f; cxyz_r[4 * id + 2] = -43.2363205f;  // :703
    cxyz_r[4 * id + 3] = 100.f;
    rgb_mat[4 * id] = 0; rgb_mat[4 * id + 1] = 0; rgb_mat[4 * id + 2] = 0; rgb_mat[4 * id + 3] = 0;
    return id + 1;
}

// Median-split BVH over the first n_use of n spheres. nodes: capacity 2*n_use-1; prim_order: n_use.
int orc_build_bvh_p(const float* prims, int prim_type, int n_use, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth);
int orc_build_bvh(const float* cxyz_r, int n_use, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth)
{
    return orc_build_bvh_p(cxyz_r, 0, n_use, nodes, prim_order, n_nodes, max_depth);
}
int orc_build_bvh_p(const float* prims, int prim_type, int n_use, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth)
{
    Prims PV{prims, prim_type};
    std::vector<Prim> P(n_use);
    for (int i = 0; i < n_use; ++i) { P[i].c = PV.centre(i); P[i].r = 0; P[i].id = i; P[i].box = PV.box(i); }
    std::vector<LinearNode> out;
    out.reserve(2 * (size_t)n_use);
    MedianBuilder B(P, out);
    B.build(0, n_use, 0);
    if (B.status) return B.status;
    *n_nodes = (int)out.size();
    if (max_depth) *max_depth = B.max_depth;
    memcpy(nodes, out.data(), sizeof(LinearNode) * out.size());
    for (int i = 0; i < n_use; ++i) prim_order[i] = P[i].id;
    return 0;
}

int orc_build_sah_p(const float* prims, int prim_type, int n, int bins, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth);
int orc_build_sah(const float* cxyz_r, int n, int bins, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth)
{
    return orc_build_sah_p(cxyz_r, 0, n, bins, nodes, prim_order, n_nodes, max_depth);
}
int orc_build_sah_p(const float* prims, int prim_type, int n, int bins, LinearNode* nodes, int* prim_order, int* n_nodes, int* max_depth)
{
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::vector<LinearNode> out;
    out.reserve(2 * (size_t)n);
    SahBuilder B(Prims{prims, prim_type}, order, out, bins);
    B.build(0, n, 0);
    *n_nodes = (int)out.size();
    if (max_depth) *max_depth = B.max_depth;
    memcpy(nodes, out.data(), sizeof(LinearNode) * out.size());
    for (int i = 0; i < n; ++i) prim_order[i] = order[i];
    return 0;
}

void orc_morton30(const float* xyz, int n, uint32_t* codes)
{
    for (int i = 0; i < n; ++i) codes[i] = morton3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

// LBVH as the reference intends it. bits 30|63; ref_norm: (c+30)/1000 (accelerators.h:577) else scene-normalised.
int orc_build_lbvh_p(const float* prims, int prim_type, int n, int bits, int ref_norm, LinearNode* nodes, int* prim_order,
                     uint64_t* keys_sorted, int* n_nodes, int* max_depth);
int orc_build_lbvh(const float* cxyz_r, int n, int bits, int ref_norm, LinearNode* nodes, int* prim_order, uint64_t* keys_sorted,
                   int* n_nodes, int* max_depth)
{
    return orc_build_lbvh_p(cxyz_r, 0, n, bits, ref_norm, nodes, prim_order, keys_sorted, n_nodes, max_depth);
}
int orc_build_lbvh_p(const float* prims, int prim_type, int n, int bits, int ref_norm, LinearNode* nodes, int* prim_order,
                     uint64_t* keys_sorted, int* n_nodes, int* max_depth)
{
    Prims PV{prims, prim_type};
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; ++i) {
        V3 cc = PV.centre(i);
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], comp(cc, a)); hi[a] = std::max(hi[a], comp(cc, a)); }
    }
    std::vector<std::pair<uint64_t, int>> kv(n);
    for (int i = 0; i < n; ++i) {
        float q[3];
        V3 cc = PV.centre(i);
        for (int a = 0; a < 3; ++a) {
            float c = comp(cc, a);
            if (ref_norm) q[a] = (c + 30.0f) / 1000.0f;
            else { float e = hi[a] - lo[a]; q[a] = e > 0.0f ? (c - lo[a]) / e : 0.0f; }
        }
        kv[i].first = bits == 30 ? (uint64_t)morton3D(q[0], q[1], q[2]) : morton63(q[0], q[1], q[2]);
        kv[i].second = i;
    }
    std::stable_sort(kv.begin(), kv.end(), [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a.first < b.first; });
    std::vector<uint64_t> K(n);
    std::vector<Box> LB(n);
    for (int i = 0; i < n; ++i) { K[i] = kv[i].first; prim_order[i] = kv[i].second; LB[i] = PV.box(kv[i].second); if (keys_sorted) keys_sorted[i] = K[i]; }
    std::vector<LinearNode> out;
    out.reserve(2 * (size_t)n);
    LbvhBuilder B(K, LB, out, bits);
    B.build(0, n - 1, 0);
    *n_nodes = (int)out.size();
    if (max_depth) *max_depth = B.max_depth;
    memcpy(nodes, out.data(), sizeof(LinearNode) * out.size());
    return 0;
}

// Closest hits of arbitrary rays. nodes == NULL -> NONE brute force.
void orc_trace_p(const float* prims, int prim_type, int n, const LinearNode* nodes, const int* prim_order, int n_nodes, int tie_by_objid,
                 const float* o, const float* d, int nrays, int* hit, float* t, long long* candidates);
void orc_trace(const float* cxyz_r, int n, const LinearNode* nodes, const int* prim_order, int n_nodes, int tie_by_objid,
               const float* o, const float* d, int nrays, int* hit, float* t, long long* candidates)
{
    orc_trace_p(cxyz_r, 0, n, nodes, prim_order, n_nodes, tie_by_objid, o, d, nrays, hit, t, candidates);
}
void orc_trace_p(const float* prims, int prim_type, int n, const LinearNode* nodes, const int* prim_order, int n_nodes, int tie_by_objid,
                 const float* o, const float* d, int nrays, int* hit, float* t, long long* candidates)
{
    Scene S{prims, nullptr, n, nodes, prim_order, n_nodes, tie_by_objid};
    S.prim_type = prim_type;
    long long cand = 0;
    for (int r = 0; r < nrays; ++r) {
        if (nodes) closest_bvh(S, o + 3 * r, d + 3 * r, hit[r], t[r], &cand);
        else closest_none(S, o + 3 * r, d + 3 * r, hit[r], t[r]);
    }
    if (candidates) *candidates = cand;
}

// Closest-hit KD traversal over an exported KdAccelNode[] (12-byte nodes, accelerators.h:715-742; leaf primitive lists
// as kdtreePrimitiveIndices). EXTENSION, PARITY UNPINNED by the reference (its KD path is any-hit, accelerators.h:997-1086):
// the walk is that function's (root slab test returning tMin/tMax accelerators.h:628-666, tPlane / belowFirst rule and
// 64-entry todo stack :1033-1075) with the two differences PBRT's KdTreeAccel::Intersect has from its IntersectP:
// a hit shortens the ray instead of returning, and the walk stops when `tnear < tMin` of the current cell. Candidates as
// main.cpp:379-384, ties towards the smaller objId (= the NONE loop's first-candidate-wins).
struct KdNode { uint32_t w0, w1, w2; };
void orc_kd_closest(const float* prims, int prim_type, int n, const KdNode* nodes, const int* prim_idx, const float* bounds6,
                    const float* o_all, const float* d_all, int nrays, int* hit, float* t, long long* prim_tests)
{
    Scene S{prims, nullptr, n, nullptr, nullptr, 0, 1};
    S.prim_type = prim_type;
    long long tests = 0;
    for (int r = 0; r < nrays; ++r) {
        const float* o = o_all + 3 * r; const float* d = d_all + 3 * r;
        hit[r] = -1; t[r] = INFINITY;
        // boundingBoxIntersection (tMin/tMax variant), accelerators.h:628-666
        float tmin = (bounds6[0] - o[0]) / d[0], tmax = (bounds6[3] - o[0]) / d[0];
        if (tmin > tmax) std::swap(tmin, tmax);
        float tymin = (bounds6[1] - o[1]) / d[1], tymax = (bounds6[4] - o[1]) / d[1];
        if (tymin > tymax) std::swap(tymin, tymax);
        if ((tmin > tymax) || (tymin > tmax)) continue;
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = (bounds6[2] - o[2]) / d[2], tzmax = (bounds6[5] - o[2]) / d[2];
        if (tzmin > tzmax) std::swap(tzmin, tzmax);
        if ((tmin > tzmax) || (tzmin > tmax)) continue;
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        float tMin = tmin, tMax = tmax;
        const float inv[3] = {1 / d[0], 1 / d[1], 1 / d[2]};
        struct Todo { int node; float tMin, tMax; } todo[64];
        int todoPos = 0, node = 0, best = -1;
        float tnear = INFINITY;
        while (true) {
            if (tnear < tMin) break;
            const KdNode& nd = nodes[node];
            if ((nd.w1 & 3u) == 3u) {
                const int np = (int)nd.w2;
                for (int i = 0; i < np; ++i) {
                    const int prim = np == 1 ? (int)nd.w0 : prim_idx[(int)nd.w0 + i];
                    float t0 = INFINITY, t1 = INFINITY;
                    ++tests;
                    if (S.test(o, d, prim, t0, t1)) {
                        if (t0 < 0) t0 = t1;
                        if (t0 < tnear || (t0 == tnear && best >= 0 && prim < best)) { tnear = t0; best = prim; }
                    }
                }
                if (todoPos > 0) { --todoPos; node = todo[todoPos].node; tMin = todo[todoPos].tMin; tMax = todo[todoPos].tMax; }
                else break;
            } else {
                const int axis = (int)(nd.w1 & 3u);
                float split; memcpy(&split, &nd.w0, 4);
                const float tPlane = (split - o[axis]) * inv[axis];
                const bool belowFirst = (o[axis] < split) || (o[axis] == split && d[axis] <= 0);
                const int below = node + 1, above = (int)(nd.w1 >> 2);
                const int first = belowFirst ? below : above, second = belowFirst ? above : below;
                if (tPlane > tMax || tPlane <= 0) node = first;
                else if (tPlane < tMin) node = second;
                else {
                    if (todoPos < 64) { todo[todoPos].node = second; todo[todoPos].tMin = tPlane; todo[todoPos].tMax = tMax; ++todoPos; }
                    node = first;
                    tMax = tPlane;
                }
            }
        }
        hit[r] = best; t[r] = tnear;
    }
    if (prim_tests) *prim_tests = tests;
}

// ----------------------------------------------------------------------------------------------
// 4-wide collapse of a flattened binary BVH (definition for the planned wide-node traversal, DESIGN.md section 10.1;
// no reference counterpart: the reference's trees are binary, accelerators.h:131-156). Deterministic:
//   wide(b) for a binary interior node b starts from the list [left(b), right(b)] and, while the list has fewer than four
//   entries and holds an interior node, replaces IN PLACE the interior entry with the largest box surface area (first one
//   on ties; SurfaceArea as accelerators.h:122-125) by its two children - 2..4 children, DFS order preserved;
//   every interior entry c left in the list becomes wide(c); wide nodes are numbered in pre-order, root = 0.
// Child boxes are the binary nodes' own boxes (bit-identical), so the candidate criterion - "the LEAF's box passes the
// reference's slab test" - is untouched and the collect-all trace below returns closest_bvh()'s hits.
// ----------------------------------------------------------------------------------------------
struct Wide4Node {
    float bmin[4][3], bmax[4][3];
    int32_t child[4];        // >= 0: wide node index; < 0: ~leafpos (index into prim_order); unused slots: INT32_MAX
    int32_t n_children, binary_node, pad[2];
};
static_assert(sizeof(Wide4Node) == 128, "Wide4Node is 128 bytes");

int orc_collapse4_rule(const LinearNode* nodes, int n_nodes, Wide4Node* out, int cap, int rule);
int orc_collapse4(const LinearNode* nodes, int n_nodes, Wide4Node* out, int cap) { return orc_collapse4_rule(nodes, n_nodes, out, cap, 0); }
// rule 0: expand the interior entry with the largest surface area (the definition); rule 1 (model experiments only): the first
// interior entry in list order; rule 2: the interior entry with the most leaves below it (subtree size from the pre-order layout);
// rule 3: plain two-level collapse (grandchildren), no choice at all
int orc_collapse4_rule(const LinearNode* nodes, int n_nodes, Wide4Node* out, int cap, int rule)
{
    if (n_nodes <= 0 || nodes[0].nPrimitives) return 0;      // a single leaf has no interior node to widen
    std::vector<std::pair<int, int>> todo;                   // (binary node, wide index), LIFO with children pushed in reverse = pre-order
    int n_wide = 1;
    todo.push_back({0, 0});
    std::vector<Wide4Node> W(1);
    while (!todo.empty()) {
        const auto [b, w] = todo.back();
        todo.pop_back();
        int listed[4] = {b + 1, nodes[b].offset, 0, 0}, m = 2;
        auto area = [&](int c) {
            const float dx = nodes[c].bmax[0] - nodes[c].bmin[0], dy = nodes[c].bmax[1] - nodes[c].bmin[1], dz = nodes[c].bmax[2] - nodes[c].bmin[2];
            return 2 * (dx * dy + dx * dz + dy * dz);
        };
        if (rule == 3) {      // model experiment: plain two-level collapse (each child replaced by its own children once)
            int two[4], k2 = 0;
            for (int k = 0; k < 2; ++k) {
                const int c = listed[k];
                if (nodes[c].nPrimitives) two[k2++] = c; else { two[k2++] = c + 1; two[k2++] = nodes[c].offset; }
            }
            for (int k = 0; k < k2; ++k) listed[k] = two[k];
            m = k2;
        }
        while (m < 4 && rule != 3) {
            int pick = -1;
            float best = -1.f;
            for (int k = 0; k < m; ++k)
                if (!nodes[listed[k]].nPrimitives) {
                    const int c = listed[k];
                    // pre-order layout: the subtree of c ends where its right sibling-or-ancestor's right child begins; its size
                    // is recovered by walking right children down to a leaf
                    float a;
                    if (rule == 0) a = area(c);
                    else if (rule == 1) a = (float)(m - k);
                    else { int e = c; while (!nodes[e].nPrimitives) e = nodes[e].offset; a = (float)(e - c); }
                    if (a > best) { best = a; pick = k; }
                }
            if (pick < 0) break;
            const int c = listed[pick];
            for (int k = m; k > pick + 1; --k) listed[k] = listed[k - 1];
            listed[pick] = c + 1; listed[pick + 1] = nodes[c].offset;
            ++m;
        }
        Wide4Node nd;
        memset(&nd, 0, sizeof nd);
        nd.n_children = m; nd.binary_node = b;
        for (int k = 0; k < 4; ++k) {
            if (k >= m) { nd.child[k] = INT32_MAX; continue; }
            const LinearNode& c = nodes[listed[k]];
            for (int a = 0; a < 3; ++a) { nd.bmin[k][a] = c.bmin[a]; nd.bmax[k][a] = c.bmax[a]; }
            if (c.nPrimitives) nd.child[k] = ~c.offset;
            else { nd.child[k] = n_wide++; W.push_back(Wide4Node()); }
        }
        W[w] = nd;
        // provisional indices are handed out in listing order; the pre-order numbering follows below
        for (int k = m - 1; k >= 0; --k) if (nd.child[k] >= 0 && nd.child[k] != INT32_MAX) todo.push_back({listed[k], nd.child[k]});
    }
    // renumber to pre-order (DFS, children in listing order)
    std::vector<int> order, new_index(n_wide, -1);
    std::vector<int> st{0};
    while (!st.empty()) {
        const int w = st.back(); st.pop_back();
        new_index[w] = (int)order.size(); order.push_back(w);
        for (int k = W[w].n_children - 1; k >= 0; --k) if (W[w].child[k] >= 0 && W[w].child[k] != INT32_MAX) st.push_back(W[w].child[k]);
    }
    if (out) {
        if (cap < n_wide) return -n_wide;
        for (int i = 0; i < n_wide; ++i) {
            Wide4Node nd = W[order[i]];
            for (int k = 0; k < nd.n_children; ++k) if (nd.child[k] >= 0) nd.child[k] = new_index[nd.child[k]];
            out[i] = nd;
        }
    }
    return n_wide;
}

// boxIntersect's collect-all walk (accelerators.h:668-690) + the candidate loop (main.cpp:343-358) over the 4-wide tree:
// every child box that passes the reference's slab test is opened; candidates are reduced with the same order-independent
// key as closest_bvh (leaf position, or objId with tie_by_objid).
void orc_trace_wide4(const float* prims, int prim_type, int n, const Wide4Node* wide, int n_wide, const int* prim_order, int tie_by_objid,
                     const float* o_all, const float* d_all, int nrays, int* hit, float* t, long long* box_tests)
{
    Scene S{prims, nullptr, n, nullptr, prim_order, 0, tie_by_objid};
    S.prim_type = prim_type;
    long long tests = 0;
    for (int r = 0; r < nrays; ++r) {
        const float* o = o_all + 3 * r; const float* d = d_all + 3 * r;
        int best = -1, best_key = 0; float tnear = INFINITY;
        int stack[256]; int sp = 0;
        if (n_wide > 0) stack[sp++] = 0;
        while (sp) {
            const Wide4Node& nd = wide[stack[--sp]];
            for (int k = 0; k < nd.n_children; ++k) {
                ++tests;
                if (!slab(o, d, nd.bmin[k], nd.bmax[k])) continue;
                if (nd.child[k] >= 0) { if (sp < 256) stack[sp++] = nd.child[k]; continue; }
                const int leafpos = ~nd.child[k], obj = prim_order[leafpos];
                float t0 = INFINITY, t1 = INFINITY;
                if (S.test(o, d, obj, t0, t1)) {
                    if (t0 < 0) t0 = t1;
                    const int key = tie_by_objid ? obj : leafpos;
                    if (t0 < tnear || (t0 == tnear && best >= 0 && key < best_key)) { tnear = t0; best = obj; best_key = key; }
                }
            }
        }
        hit[r] = best; t[r] = tnear;
    }
    if (box_tests) *box_tests = tests;
}


void orc_jitter(double* out, int n, unsigned long long first)
{
    MT g;
    g.discard(2 * first);
    for (int i = 0; i < n; ++i) out[i] = g.canonical();
}

// render() rows [y0,y1) (main.cpp:541-566) + write_into_file's quantisation (main.cpp:521-523).
// lights: m x {cx,cy,cz,radius,r,g,b}. nodes == NULL -> NONE. Optional outputs may be NULL.
void orc_render_rows_ex(const float* cxyz_r, const float* rgb_mat, int n, const LinearNode* nodes, const int* prim_order, int n_nodes,
                        int tie_by_objid, const float* lights7, int m, int width, int height, int spp, int y0, int y1,
                        int shadows, uint8_t* rgb8, int* hit_out, float* accum, float* dirs, long long* ray_counts3)
{
    Scene S{cxyz_r, rgb_mat, n, nodes, prim_order, n_nodes, tie_by_objid & 1};
    S.prim_type = (tie_by_objid >> 8) & 3;   // bits 8-9 of the flag word: the primitive table holds triangles (n x 9); 2 = geometric test
    std::vector<Light> L(m);
    for (int i = 0; i < m; ++i) {
        const float* l = lights7 + 7 * i;
        L[i] = Light{{l[0], l[1], l[2]}, l[3], {l[4], l[5], l[6]}};
    }
    CastCtx CC;
    CC.S = &S; CC.L = L.data(); CC.nl = m; CC.shadows = shadows;
    MT gen;
    gen.discard(4ull * (unsigned long long)y0 * width * spp);
    float invWidth = 1 / float(width), invHeight = 1 / float(height);         // :544
    float fov = 30, aspectratio = width / float(height);                       // :545
    float angle = (float)std::tan(3.141592653589793 * 0.5 * fov / 180.);       // :546
    const float o[3] = {0, 0, 0};
    size_t k = 0, kd = 0;
    for (unsigned y = (unsigned)y0; y < (unsigned)y1; ++y) {
        for (unsigned x = 0; x < (unsigned)width; ++x, ++k) {
            float px[3] = {0, 0, 0};
            int last = -1;
            for (int s = 0; s < spp; ++s) {
                float xx = (float)((2 * ((x + gen.canonical()) * invWidth) - 1) * angle * aspectratio);  // :554
                float yy = (float)((1 - 2 * ((y + gen.canonical()) * invHeight)) * angle);               // :555
                float d[3] = {xx, yy, -1};
                normalize(d);
                if (dirs) { dirs[kd++] = d[0]; dirs[kd++] = d[1]; dirs[kd++] = d[2]; }
                int hit = -1;
                float c[3];
                cast_ray(CC, o, d, 1, c, &hit);                                                           // :558
                px[0] += c[0]; px[1] += c[1]; px[2] += c[2];
                last = hit;
            }
            if (hit_out) hit_out[k] = last;
            if (accum) { accum[3 * k] = px[0]; accum[3 * k + 1] = px[1]; accum[3 * k + 2] = px[2]; }
            if (rgb8) {
                float fs = (float)(uint32_t)spp;
                rgb8[3 * k]     = (unsigned char)(std::min(float(1), px[0] / fs) * 255);
                rgb8[3 * k + 1] = (unsigned char)(std::min(float(1), px[1] / fs) * 255);
                rgb8[3 * k + 2] = (unsigned char)(std::min(float(1), px[2] / fs) * 255);
            }
        }
    }
    if (ray_counts3) { ray_counts3[0] = CC.rays; ray_counts3[1] = CC.shadow_rays; ray_counts3[2] = CC.secondary_rays; }
}

void orc_render_rows(const float* cxyz_r, const float* rgb_mat, int n, const LinearNode* nodes, const int* prim_order, int n_nodes,
                     int tie_by_objid, const float* lights7, int m, int width, int height, int spp, int y0, int y1,
                     uint8_t* rgb8, int* hit_out, float* accum, float* dirs)
{
    orc_render_rows_ex(cxyz_r, rgb_mat, n, nodes, prim_order, n_nodes, tie_by_objid, lights7, m, width, height, spp, y0, y1, 0,
                       rgb8, hit_out, accum, dirs, nullptr);
}

}  // extern "C"
