// lbvh.cu — the LBVH the reference intends (accelerators.h:371,568 cite NVIDIA's "Thinking Parallel III"):
//   K1 scene/centroid bounds reduce      (JoinBounds / JoinBoxPopintBounds, accelerators.h:201-229)
//   K2 30-bit / 63-bit Morton codes      (expandBits / morton3D, accelerators.h:374-394; 63-bit is an extension)
//   K3 onesweep radix sort               (sort.cu)
//   K4 Karras hierarchy emission         (the intended behaviour of findSplit, accelerators.h:397-449)
//   K5 atomic bottom-up AABB refit       (node bounds = union of leaf boxBoundries, accelerators.h:257-260)
// plus the pre-order flattening into the reference's LinearBVHNode layout (accelerators.h:231-240).
#include "rtds_internal.cuh"
#include <chrono>
#include <math.h>
#include <string.h>
#include <stdlib.h>

int rtds_bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches);

namespace {

// ---------------------------------------------------------------------------------------------------
// K1: bounds.  out[0..5] = centre min.xyz / max.xyz, out[6..11] = AABB (c -/+ r) min / max, as ordered uints.
// ---------------------------------------------------------------------------------------------------
__global__ void bounds_init_kernel(unsigned* out)
{
    int i = threadIdx.x;
    if (i < 12) out[i] = ((i % 6) < 3) ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) bounds_kernel(const PrimView pv, int n, unsigned* __restrict__ out)
{
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float c[3], mn[3], mx[3];
        prim_fetch(pv, i, c, mn, mx);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            cmin[a] = fminf(cmin[a], c[a]);
            cmax[a] = fmaxf(cmax[a], c[a]);
            bmin[a] = fminf(bmin[a], mn[a]);
            bmax[a] = fmaxf(bmax[a], mx[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            cmin[a] = fminf(cmin[a], __shfl_xor_sync(0xffffffffu, cmin[a], o));
            cmax[a] = fmaxf(cmax[a], __shfl_xor_sync(0xffffffffu, cmax[a], o));
            bmin[a] = fminf(bmin[a], __shfl_xor_sync(0xffffffffu, bmin[a], o));
            bmax[a] = fmaxf(bmax[a], __shfl_xor_sync(0xffffffffu, bmax[a], o));
        }
    }
    // one set of 12 atomics per BLOCK: all warps hammering the same 12 words made this kernel atomic-bound (77 us for 17 MB)
    __shared__ unsigned sh[8][12];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sh[warp][a] = f2ord(cmin[a]); sh[warp][3 + a] = f2ord(cmax[a]);
            sh[warp][6 + a] = f2ord(bmin[a]); sh[warp][9 + a] = f2ord(bmax[a]);
        }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const bool is_min = (threadIdx.x % 6) < 3;
        unsigned v = sh[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = is_min ? min(v, sh[w][threadIdx.x]) : max(v, sh[w][threadIdx.x]);
        if (is_min) atomicMin(&out[threadIdx.x], v); else atomicMax(&out[threadIdx.x], v);
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: Morton codes
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned expand_bits10(unsigned v)  // accelerators.h:374-381
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ unsigned morton30(float x, float y, float z)  // accelerators.h:385-394
{
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    unsigned xx = expand_bits10((unsigned)x);
    unsigned yy = expand_bits10((unsigned)y);
    unsigned zz = expand_bits10((unsigned)z);
    return xx * 4 + yy * 2 + zz;
}

__device__ __forceinline__ unsigned long long expand_bits21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__device__ __forceinline__ unsigned long long morton63(float x, float y, float z)
{
    x = fminf(fmaxf(x * 2097152.0f, 0.0f), 2097151.0f);
    y = fminf(fmaxf(y * 2097152.0f, 0.0f), 2097151.0f);
    z = fminf(fmaxf(z * 2097152.0f, 0.0f), 2097151.0f);
    return expand_bits21((unsigned long long)x) * 4 + expand_bits21((unsigned long long)y) * 2 +
           expand_bits21((unsigned long long)z);
}

// ref_norm: (centre + 30) / 1000 (accelerators.h:577); else (centre - cmin) / (cmax - cmin) per axis.
template <typename K>
__global__ void __launch_bounds__(256)
morton_kernel(const PrimView pv, int n, const unsigned* __restrict__ bounds_ord, int ref_norm,
              K* __restrict__ keys, uint32_t* __restrict__ vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c_[3], mn_[3], mx_[3];
    prim_fetch(pv, i, c_, mn_, mx_);
    const float3 s = make_float3(c_[0], c_[1], c_[2]);
    float x, y, z;
    if (ref_norm) {
        x = (s.x + 30.0f) / 1000.0f; y = (s.y + 30.0f) / 1000.0f; z = (s.z + 30.0f) / 1000.0f;
    } else {
        float lo[3] = {ord2f(bounds_ord[0]), ord2f(bounds_ord[1]), ord2f(bounds_ord[2])};
        float hi[3] = {ord2f(bounds_ord[3]), ord2f(bounds_ord[4]), ord2f(bounds_ord[5])};
        float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        x = ex > 0.0f ? (s.x - lo[0]) / ex : 0.0f;
        y = ey > 0.0f ? (s.y - lo[1]) / ey : 0.0f;
        z = ez > 0.0f ? (s.z - lo[2]) / ez : 0.0f;
    }
    if (sizeof(K) == 4) keys[i] = (K)morton30(x, y, z);
    else keys[i] = (K)morton63(x, y, z);
    vals[i] = (uint32_t)i;
}

__global__ void morton30_points_kernel(const float* __restrict__ xyz, int n, uint32_t* __restrict__ codes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) codes[i] = morton30(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

// ---------------------------------------------------------------------------------------------------
// K4: Karras hierarchy emission over the sorted keys.  Duplicate keys are disambiguated by position.
// ---------------------------------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ int delta(const K* __restrict__ keys, int n, int i, K ki, int j)
{
    if (j < 0 || j >= n) return -1;
    K kj = keys[j];
    if (ki == kj) return (int)(sizeof(K) * 8) + __clz(i ^ j);
    return sizeof(K) == 4 ? __clz((unsigned)(ki ^ kj)) : __clzll((long long)(ki ^ kj));
}

template <typename K>
__global__ void __launch_bounds__(256)
karras_kernel(const K* __restrict__ keys, int n, Node64* __restrict__ nodes, int* __restrict__ leaf_parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    K ki = keys[i];
    int d = (delta(keys, n, i, ki, i + 1) - delta(keys, n, i, ki, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, ki, i - d);
    int lmax = 2;
    while (delta(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, ki, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    nodes[i].left = left;
    nodes[i].right = right;
    // axis of the first differing Morton bit (bit 3k+2 = x, 3k+1 = y, 3k = z); equal keys -> 0
    int axis = 0;
    if (dnode < (int)(sizeof(K) * 8)) {
        int b = (int)(sizeof(K) * 8) - 1 - dnode;
        axis = 2 - (b % 3);
    }
    nodes[i].axis = axis;
    if (i == 0) nodes[0].parent = -1;
    if (left < 0) leaf_parent[~left] = i; else nodes[left].parent = i * 2;
    if (right < 0) leaf_parent[~right] = i | 0x80000000; else nodes[right].parent = i * 2 + 1;
}

// ---------------------------------------------------------------------------------------------------
// K5: bottom-up refit.  One thread per leaf writes its box into its parent's child slot; the second
// thread to arrive at an interior node (atomic counter) unions the two slots and carries on upward.
// Node64::parent holds parent*2+side for interior nodes (side 1 = right child), -1 for the root.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_child_box(Node64* nd, int side, const float mn[3], const float mx[3])
{
    float* pmin = side ? nd->rmin : nd->lmin;
    float* pmax = side ? nd->rmax : nd->lmax;
#pragma unroll
    for (int a = 0; a < 3; ++a) { __stcg(pmin + a, mn[a]); __stcg(pmax + a, mx[a]); }
}

__global__ void __launch_bounds__(256)
refit_kernel(const PrimView pv, const uint32_t* __restrict__ sorted_ids, int n, Node64* nodes,
             const int* __restrict__ leaf_parent, float4* __restrict__ leaf_sph, float4* __restrict__ leaf_tri, int* __restrict__ prim_order,
             unsigned* __restrict__ counters, float* __restrict__ root_box, int* __restrict__ max_depth)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned height = 0;     // of the subtree this thread has finished; travels through the arrival counter (bits 8..)
    int prim = (int)sorted_ids[j];
    prim_store_leaf(pv, prim, j, leaf_sph, leaf_tri);
    prim_order[j] = prim;
    float cc[3], mn[3], mx[3];
    prim_fetch(pv, prim, cc, mn, mx);
    if (n == 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { root_box[a] = mn[a]; root_box[3 + a] = mx[a]; }
        return;
    }
    int lp = leaf_parent[j];
    int parent = lp & 0x7fffffff, side = (lp >> 31) & 1;
    while (true) {
        Node64* nd = nodes + parent;
        store_child_box(nd, side, mn, mx);
        __threadfence();
        unsigned old = atomicAdd(&counters[parent], 1u + (height << 8));
        if (old == 0) return;  // the sibling subtree is not finished: its last thread continues
        height = 1u + max(height, old >> 8);
        __threadfence();
        const float* omin = side ? nd->lmin : nd->rmin;
        const float* omax = side ? nd->lmax : nd->rmax;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __ldcg(omin + a));
            mx[a] = fmaxf(mx[a], __ldcg(omax + a));
        }
        int pe = __ldcg(&nd->parent);
        if (pe < 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { root_box[a] = mn[a]; root_box[3 + a] = mx[a]; }
            if (max_depth) *max_depth = (int)height;      // depth of the deepest leaf (root = 0): bounds the traversal stack
            return;
        }
        parent = pe >> 1;
        side = pe & 1;
    }
}

// depth of the deepest leaf (root = depth 0): bounds the traversal stack
__global__ void __launch_bounds__(256)
depth_kernel(const Node64* __restrict__ nodes, const int* __restrict__ leaf_parent, int n, int* __restrict__ max_depth)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int depth = 0;
    if (j < n && n > 1) {
        int p = leaf_parent[j] & 0x7fffffff;
        depth = 1;
        while (true) {
            int pe = nodes[p].parent;
            if (pe < 0) break;
            p = pe >> 1;
            ++depth;
        }
    }
    for (int o = 16; o; o >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, o));
    if ((threadIdx.x & 31) == 0 && depth > 0) atomicMax(max_depth, depth);
}

// ---------------------------------------------------------------------------------------------------
// Pre-order flattening (K9).  Leaf order is DFS order in every builder of this library, so a node whose
// subtree starts at leaf f and that lies in the LEFT subtree of `lv` of its ancestors has pre-order index
// 2*f + lv (f leaves and f - (#right turns) interior nodes precede it, plus its ancestors).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int first_leaf(const Node64* __restrict__ nodes, int ref)
{
    while (ref >= 0) ref = nodes[ref].left;
    return ~ref;
}

__global__ void __launch_bounds__(256)
preorder_kernel(const Node64* __restrict__ nodes, const int* __restrict__ leaf_parent, int n,
                const float* __restrict__ root_box, rtds_linear_bvh_node* __restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n_internal = n - 1;
    if (t >= 2 * n - 1) return;
    rtds_linear_bvh_node r;
    if (n == 1) {
        for (int a = 0; a < 3; ++a) { r.bmin[a] = root_box[a]; r.bmax[a] = root_box[3 + a]; }
        r.offset = 0; r.nPrimitives = 1; r.axis = 0; r.pad = 0;
        out[0] = r;
        return;
    }
    int pe;          // parent*2+side of the node this thread handles
    int f;           // first leaf of its subtree
    bool leaf = t >= n_internal;
    if (leaf) {
        int j = t - n_internal;
        int lp = leaf_parent[j];
        pe = ((lp & 0x7fffffff) << 1) | ((lp >> 31) & 1);
        f = j;
        r.offset = j; r.nPrimitives = 1; r.axis = 0; r.pad = 0;
    } else {
        pe = nodes[t].parent;
        f = first_leaf(nodes, t);
        r.nPrimitives = 0; r.axis = (uint8_t)nodes[t].axis; r.pad = 0;
    }
    // own box: from the parent's child slot (root: root_box)
    if (pe < 0) {
        for (int a = 0; a < 3; ++a) { r.bmin[a] = root_box[a]; r.bmax[a] = root_box[3 + a]; }
    } else {
        const Node64* pn = nodes + (pe >> 1);
        const float* mn = (pe & 1) ? pn->rmin : pn->lmin;
        const float* mx = (pe & 1) ? pn->rmax : pn->lmax;
        for (int a = 0; a < 3; ++a) { r.bmin[a] = mn[a]; r.bmax[a] = mx[a]; }
    }
    int lv = 0;
    for (int q = pe; q >= 0; q = nodes[q >> 1].parent) lv += (q & 1) ? 0 : 1;
    int my = 2 * f + lv;
    if (!leaf) {
        int fr = first_leaf(nodes, nodes[t].right);
        r.offset = 2 * fr + lv;  // secondChildOffset: the right child is in the left subtree of the same ancestors
    }
    out[my] = r;
}

// ---------------------------------------------------------------------------------------------------
// Layout pass: renumber the interior nodes in depth-first pre-order (new index = first leaf of the subtree + number
// of ancestors the node is a LEFT descendant of), so a node's left child is the next record and a subtree is one
// contiguous run of memory. Topology, boxes and leaf order are untouched (exports are identical); only the node
// stream's locality changes — what matters once the tree no longer fits L2 (7 M primitives = 448 MB of nodes).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
preorder_index_kernel(const Node64* __restrict__ nodes, int n_internal, int* __restrict__ new_index)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    int f = first_leaf(nodes, i);
    int lv = 0;
    for (int q = nodes[i].parent; q >= 0; q = nodes[q >> 1].parent) lv += (q & 1) ? 0 : 1;
    new_index[i] = f + lv;
}

__global__ void __launch_bounds__(256)
reorder_nodes_kernel(const Node64* __restrict__ in, int n_internal, const int* __restrict__ new_index, Node64* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    Node64 nd = in[i];
    if (nd.left >= 0) nd.left = new_index[nd.left];
    if (nd.right >= 0) nd.right = new_index[nd.right];
    if (nd.parent >= 0) nd.parent = new_index[nd.parent >> 1] * 2 + (nd.parent & 1);
    out[new_index[i]] = nd;
}

__global__ void __launch_bounds__(256)
reorder_leaf_parent_kernel(int* __restrict__ leaf_parent, int n, const int* __restrict__ new_index)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int lp = leaf_parent[j];
    leaf_parent[j] = new_index[lp & 0x7fffffff] | (lp & 0x80000000);
}

template <typename K>
__global__ void widen_keys_kernel(const K* __restrict__ in, int n, uint64_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint64_t)in[i];
}

template <typename K>
int build_true(rtds_ctx* ctx, const rtds_build_params* p, rtds_build_stats* st, int key_bits)
{
    const int n = ctx->n;
    DeviceBvh& b = ctx->bvh;
    int launches = 0;
    const auto tr0 = std::chrono::steady_clock::now();
    auto tr_us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count(); };
    double tr[4] = {0, 0, 0, 0};
    // scratch: bounds[12] | counters[n] | max_depth | keys, keys_tmp | vals, vals_tmp
    size_t off_bounds = 0;
    size_t off_counters = 256;
    size_t off_depth = off_counters + sizeof(unsigned) * (size_t)n;
    off_depth = (off_depth + 255) & ~(size_t)255;
    size_t off_keys = off_depth + 256;
    size_t off_keys_tmp = off_keys + ((sizeof(K) * (size_t)n + 255) & ~(size_t)255);
    size_t off_vals = off_keys_tmp + ((sizeof(K) * (size_t)n + 255) & ~(size_t)255);
    size_t off_vals_tmp = off_vals + ((sizeof(uint32_t) * (size_t)n + 255) & ~(size_t)255);
    size_t off_reorder = off_vals_tmp + ((sizeof(uint32_t) * (size_t)n + 255) & ~(size_t)255);
    size_t total = off_reorder + (((size_t)n * 4 + 255) & ~(size_t)255) + sizeof(Node64) * (size_t)n + 256;
    RTDS_TRY(rtds_ensure_scratch(ctx, total));
    char* base = (char*)ctx->d_scratch;
    unsigned* d_bounds = (unsigned*)(base + off_bounds);
    unsigned* d_counters = (unsigned*)(base + off_counters);
    int* d_depth = (int*)(base + off_depth);
    K* d_keys = (K*)(base + off_keys);
    K* d_keys_tmp = (K*)(base + off_keys_tmp);
    uint32_t* d_vals = (uint32_t*)(base + off_vals);
    uint32_t* d_vals_tmp = (uint32_t*)(base + off_vals_tmp);

    RTDS_TRY(rtds_alloc_bvh_for(ctx, b, n));
    if (ctx->keys_capacity < n) {
        if (ctx->d_keys_sorted) cudaFree(ctx->d_keys_sorted);
        ctx->d_keys_sorted = nullptr;
        RTDS_CUDA(cudaMalloc(&ctx->d_keys_sorted, sizeof(uint64_t) * (size_t)n));
        ctx->keys_capacity = n;
    }
    ctx->n_keys = n;

    cudaStream_t s = ctx->stream;
    tr[0] = tr_us();
    RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    RTDS_TRACE_RECORD(ctx, 4, s);
    const int T = 256;
    const int G = (n + T - 1) / T;
    bounds_init_kernel<<<1, 32, 0, s>>>(d_bounds);
    const PrimView pv = rtds_prim_view(ctx);
    bounds_kernel<<<min(G, ctx->sm_count * 8), T, 0, s>>>(pv, n, d_bounds);
    morton_kernel<K><<<G, T, 0, s>>>(pv, n, d_bounds, p ? p->morton_ref_norm : 0, d_keys, d_vals);
    launches += 3;
    tr[1] = tr_us();
    RTDS_TRY((sizeof(K) == 4)
                 ? rtds_onesweep_sort_u32(ctx, (uint32_t*)d_keys, d_vals, (uint32_t*)d_keys_tmp, d_vals_tmp, n, key_bits, &launches)
                 : rtds_onesweep_sort_u64(ctx, (uint64_t*)d_keys, d_vals, (uint64_t*)d_keys_tmp, d_vals_tmp, n, key_bits, &launches));
    RTDS_TRY(rtds_zero_async(d_counters, sizeof(unsigned) * (size_t)n, s, &launches));     // kernels, not cudaMemsetAsync: see rtds_zero_async
    RTDS_TRY(rtds_zero_async(d_depth, sizeof(int), s, &launches));
    float* d_root_box = (float*)(d_bounds + 16);
    if (n > 1) {
        karras_kernel<K><<<(n - 1 + T - 1) / T, T, 0, s>>>(d_keys, n, b.nodes, b.leaf_parent);
        launches += 1;
    }
    refit_kernel<<<G, T, 0, s>>>(pv, d_vals, n, b.nodes, b.leaf_parent, b.leaf_sph, b.leaf_tri, b.prim_order, d_counters, d_root_box, d_depth);
    b.n_prims = n;
    RTDS_TRY(rtds_bvh_reorder_preorder(ctx, b, base + off_reorder, &launches));
    widen_keys_kernel<K><<<G, T, 0, s>>>(d_keys, n, ctx->d_keys_sorted);
    launches += 2;
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    RTDS_TRACE_RECORD(ctx, 5, s);
    RTDS_CUDA(cudaGetLastError());
    // root box + depth come back into PINNED memory: a copy into pageable memory is staged by the driver, ~15 us each
    // on the rtds_frame critical path
    float* h_box = reinterpret_cast<float*>(ctx->h_counters + 9);
    int* h_depth = reinterpret_cast<int*>(ctx->h_counters + 12);
    RTDS_CUDA(cudaMemcpyAsync(h_box, d_root_box, sizeof(float) * 6, cudaMemcpyDeviceToHost, s));
    tr[2] = tr_us();
    RTDS_CUDA(cudaMemcpyAsync(h_depth, d_depth, sizeof(int), cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    tr[3] = tr_us();
    for (int i = 0; i < 6; ++i) b.root_box[i] = h_box[i];
    const int depth = *h_depth;
    if (ctx->trace_ev[0]) fprintf(stderr, "[lbvh build host] allocs done %.0f us | 3 launches in %.0f | all enqueued %.0f | synced %.0f\n", tr[0], tr[1], tr[2], tr[3]);
    b.n_prims = n;
    b.n_internal = n - 1;
    b.root_ref = n > 1 ? 0 : ~0;
    b.tie_by_objid = 1;
    b.leaf_box_prim = ctx->prim_type == 0;
    b.max_depth = depth;
    b.valid = true;
    float ms = 0;
    RTDS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (st) {
        st->n_prims = n;
        st->total_nodes = 2 * n - 1;
        st->alloc_nodes = 2 * n - 1;
        st->max_depth = depth;
        st->kernel_launches = launches;
        st->ms = ms;
    }
    return RTDS_OK;
}

}  // namespace

// AABB of the scene: out12 = centre min/max (6), box min/max (6)
int rtds_scene_bounds(rtds_ctx* ctx, float out12[12])
{
    RTDS_TRY(rtds_ensure_scratch(ctx, 4096));
    unsigned* d_bounds = (unsigned*)((char*)ctx->d_scratch + ((ctx->scratch_bytes - 512) & ~(size_t)255));   // aligned tail of the scratch area
    bounds_init_kernel<<<1, 32, 0, ctx->stream>>>(d_bounds);
    bounds_kernel<<<min((ctx->n + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(rtds_prim_view(ctx), ctx->n, d_bounds);
    RTDS_CUDA(cudaGetLastError());
    unsigned h[12];
    RTDS_CUDA(cudaMemcpyAsync(h, d_bounds, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 12; ++i) {
        unsigned u = h[i];
        unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        memcpy(&out12[i], &b, 4);
    }
    return RTDS_OK;
}

// Bottom-up refit of a finished topology (nodes[].left/right/parent, leaf_parent[]) whose leaves hold the primitives
// d_ids[leafpos]: fills every child-box slot, leaf_sph, prim_order and root_box. d_counters: n zeroed unsigneds.
int rtds_bvh_refit(rtds_ctx* ctx, DeviceBvh& b, const uint32_t* d_ids, int n, unsigned* d_counters, float* d_root_box, int* d_depth)
{
    RTDS_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(unsigned) * (size_t)n, ctx->stream));
    refit_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(rtds_prim_view(ctx), d_ids, n, b.nodes, b.leaf_parent, b.leaf_sph, b.leaf_tri,
                                                           b.prim_order, d_counters, d_root_box, d_depth);
    RTDS_CUDA(cudaGetLastError());
    return RTDS_OK;
}

// ---------------------------------------------------------------------------------------------------
// K12: 4-wide collapse (`wide` option). One thread per interior node: its depth parity by walking the parent links (the builders
// all write Node64::parent); a node at EVEN depth gathers its grandchildren's boxes from its two children's records - a child
// that is a leaf contributes itself with the box its parent holds for it. Nodes at odd depth are absorbed (their slot stays unused).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wide_collapse_kernel(const Node64* __restrict__ nodes, int n_internal, Wide4* __restrict__ wide)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    int depth = 0;
    for (int pe = __ldg(&nodes[i].parent); pe >= 0; pe = __ldg(&nodes[pe >> 1].parent)) ++depth;
    if (depth & 1) return;
    const Node64 nd = nodes[i];
    Wide4 w;
    int k = 0;
    auto put = [&](int ref, const float* mn, const float* mx) {
        w.x0[k] = mn[0]; w.y0[k] = mn[1]; w.z0[k] = mn[2]; w.x1[k] = mx[0]; w.y1[k] = mx[1]; w.z1[k] = mx[2]; w.ref[k] = ref; ++k;
    };
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int c = side ? nd.right : nd.left;
        if (c >= 0) {
            const Node64 ch = nodes[c];
            put(ch.left, ch.lmin, ch.lmax);
            put(ch.right, ch.rmin, ch.rmax);
        } else {
            put(c, side ? nd.rmin : nd.lmin, side ? nd.rmax : nd.lmax);
        }
    }
    for (; k < 4; ++k) {
        w.x0[k] = w.y0[k] = w.z0[k] = INFINITY; w.x1[k] = w.y1[k] = w.z1[k] = -INFINITY; w.ref[k] = WIDE_EMPTY;
    }
    w.pad[0] = w.pad[1] = w.pad[2] = w.pad[3] = 0;
    wide[i] = w;
}

static int bvh_collapse_wide(rtds_ctx* ctx, DeviceBvh& b, int* launches)
{
    b.wide_valid = false;
    const int ni = b.n_prims - 1;
    if (!ctx->opt.wide || ni < 1) return RTDS_OK;
    if (b.wide_capacity < ni) {
        RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (b.wide) cudaFree(b.wide);
        b.wide = nullptr; b.wide_capacity = 0;
        RTDS_CUDA(cudaMalloc(&b.wide, sizeof(Wide4) * (size_t)ni));
        b.wide_capacity = ni;
    }
    wide_collapse_kernel<<<(ni + 255) / 256, 256, 0, ctx->stream>>>(b.nodes, ni, b.wide);
    RTDS_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    b.wide_valid = true;
    return RTDS_OK;
}

// Layout passes behind every BVH build (all three builders end here): the optional pre-order renumbering, then the optional
// 4-wide collapse. Renumbers b.nodes in depth-first pre-order (see preorder_index_kernel); needs n_internal * (4 + 64) bytes of
// scratch at `scratch` (not overlapping anything still in use).
static int bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches);
int rtds_bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches)
{
    RTDS_TRY(bvh_reorder_preorder(ctx, b, scratch, launches));
    return bvh_collapse_wide(ctx, b, launches);
}
static int bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches)
{
    const int ni = b.n_prims - 1;
    if (ni <= 1) return RTDS_OK;
    // Off by default: measured +0.09 ms per million primitives of build time for <= 1 % of traversal time, on the
    // L2-resident 1 M-primitive tree and on the 7 M-primitive one alike (RTDS_NODE_ORDER=preorder turns it on).
    if (!ctx->opt.node_preorder) return RTDS_OK;
    int* new_index = (int*)scratch;
    Node64* tmp = (Node64*)((char*)scratch + (((size_t)ni * 4 + 255) & ~(size_t)255));
    cudaStream_t s = ctx->stream;
    const int G = (ni + 255) / 256;
    preorder_index_kernel<<<G, 256, 0, s>>>(b.nodes, ni, new_index);
    reorder_nodes_kernel<<<G, 256, 0, s>>>(b.nodes, ni, new_index, tmp);
    reorder_leaf_parent_kernel<<<(b.n_prims + 255) / 256, 256, 0, s>>>(b.leaf_parent, b.n_prims, new_index);
    RTDS_CUDA(cudaMemcpyAsync(b.nodes, tmp, sizeof(Node64) * (size_t)ni, cudaMemcpyDeviceToDevice, s));
    RTDS_CUDA(cudaGetLastError());
    if (launches) *launches += 3;
    return RTDS_OK;
}

int rtds_bvh_compute_depth(rtds_ctx* ctx, DeviceBvh& b, int* depth_out)
{
    int* d_depth = (int*)(ctx->d_counters + 7);   // last slot of the context's counter block
    RTDS_CUDA(cudaMemsetAsync(d_depth, 0, sizeof(int), ctx->stream));
    depth_kernel<<<(b.n_prims + 255) / 256, 256, 0, ctx->stream>>>(b.nodes, b.leaf_parent, b.n_prims, d_depth);
    RTDS_CUDA(cudaGetLastError());
    RTDS_CUDA(cudaMemcpyAsync(depth_out, d_depth, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
    return RTDS_OK;
}

int rtds_build_lbvh_true(rtds_ctx* ctx, const rtds_build_params* p, rtds_build_stats* st)
{
    int bits = (p && p->morton_bits) ? p->morton_bits : 30;
    if (bits == 30) return build_true<uint32_t>(ctx, p, st, 30);
    if (bits == 63) return build_true<uint64_t>(ctx, p, st, 63);
    rtds_set_error("morton_bits must be 30 or 63 (got %d)", bits);
    return RTDS_ERR_INVALID;
}

int rtds_morton30_device(rtds_ctx* ctx, const float* h_xyz, int n, uint32_t* h_codes)
{
    if (n <= 0) return RTDS_OK;
    size_t bytes_in = sizeof(float) * 3 * (size_t)n, bytes_out = sizeof(uint32_t) * (size_t)n;
    RTDS_TRY(rtds_ensure_scratch(ctx, bytes_in + bytes_out + 512));
    float* d_in = (float*)ctx->d_scratch;
    uint32_t* d_out = (uint32_t*)((char*)ctx->d_scratch + ((bytes_in + 255) & ~(size_t)255));
    RTDS_CUDA(cudaMemcpyAsync(d_in, h_xyz, bytes_in, cudaMemcpyHostToDevice, ctx->stream));
    morton30_points_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_in, n, d_out);
    RTDS_CUDA(cudaGetLastError());
    RTDS_CUDA(cudaMemcpyAsync(h_codes, d_out, bytes_out, cudaMemcpyDeviceToHost, ctx->stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
    return RTDS_OK;
}

int rtds_bvh_preorder_export(rtds_ctx* ctx, rtds_linear_bvh_node* h_nodes, int cap_nodes, int* n_nodes,
                             int* h_prim_order, int cap_prims, int* n_prims)
{
    DeviceBvh& b = ctx->bvh;
    if (!b.valid) { rtds_set_error("export_bvh: no BVH/LBVH has been built"); return RTDS_ERR_NOT_BUILT; }
    const int n = b.n_prims, total = 2 * n - 1;
    if (n_nodes) *n_nodes = total;
    if (n_prims) *n_prims = n;
    if ((h_nodes && cap_nodes < total) || (h_prim_order && cap_prims < n)) {
        rtds_set_error("export_bvh: need %d nodes / %d prims", total, n);
        return RTDS_ERR_CAPACITY;
    }
    if (h_nodes) {
        size_t bytes = sizeof(rtds_linear_bvh_node) * (size_t)total;
        RTDS_TRY(rtds_ensure_scratch(ctx, bytes + 256));
        float* d_root = (float*)ctx->d_scratch;
        rtds_linear_bvh_node* d_out = (rtds_linear_bvh_node*)((char*)ctx->d_scratch + 256);
        RTDS_CUDA(cudaMemcpyAsync(d_root, b.root_box, sizeof(float) * 6, cudaMemcpyHostToDevice, ctx->stream));
        preorder_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(b.nodes, b.leaf_parent, n, d_root, d_out);
        RTDS_CUDA(cudaGetLastError());
        RTDS_CUDA(cudaMemcpyAsync(h_nodes, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (h_prim_order)
        RTDS_CUDA(cudaMemcpyAsync(h_prim_order, b.prim_order, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RTDS_CUDA(cudaStreamSynchronize(ctx->stream));
    return RTDS_OK;
}
