// median.cu — K6: the reference's BVH builder, constructBVHNew (accelerators.h:246-337), on the GPU, bit-exact:
// same topology, same AABB bit patterns, same primitive order — including what libstdc++'s std::partition
// (stl_algo.h:1472-1496) and std::nth_element (__introselect, stl_algo.h:1957-1980) do under ties.
//
// The recursion is run level by level.  Every node ("task") owns a range [s,e) of the permuted scene vector:
//   1. bounds = union of the range's AABBs (:257-260); written straight into the parent's child slot — the
//      reference computes a node's box over its whole range BEFORE any drop, so no bottom-up refit is used;
//   2. dim = longest axis (:279), pmid (:298), std::partition by centre[dim] < pmid (:302-306);
//   3. if the partition is not degenerate, std::nth_element at (s+e)/2 by centre[dim] (:311-319);
//   4. partition == start: the node becomes a leaf holding scene[s], the rest of the range is DROPPED (:321-327);
//      partition == end: the reference recurses forever -> RTDS_ERR_DEGENERATE;
//   5. children [s,mid) and [mid,e).
//
// libstdc++'s Hoare passes are sequential scans, but their net effect is a pairing of "stop" positions that
// only depends on prefix counts, so a thread block reproduces them with compactions:
//   __partition(pred):  F = ascending !pred positions left of the final cut, T = ascending pred positions right
//                       of it, |F| = |T| = k; element F[i] swaps with T[k-1-i].
//   __unguarded_partition(pivot v) on [lo,hi): L = ascending positions with !(a<v), R = descending positions with
//                       !(v<a); k = #{i : L[i] < R[i]}; swap L[i]<->R[i] for i<k; cut = k ? min(L[k], R[k-1]) : L[0].
// __move_median_to_first, __insertion_sort (<= 3 elements) and the depth-limit fallback (__heap_select +
// iter_swap) are executed literally by one thread.  Ranges of <= `small` elements are finished by ONE thread
// running the literal sequential algorithms on its whole subtree.
#include "rtds_internal.cuh"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <cooperative_groups.h>

namespace {

struct Task { int s, e, parent_enc; };   // parent_enc = parent*2+side, -1 for the root

struct MedianArgs {
    PrimView pv;             // objId -> centre / AABB
    int*   perm;             // scene position -> objId
    float* key;              // scene position -> centre[dim] of the task that owns the position
    int*   la;               // scratch lists, indexed by scene position
    int*   lb;
    int*   leaf_info;        // scene position -> parent_enc+2 for leaves, 0 otherwise
    Node64* nodes;
    float* root_box;         // [6]
    int*   counters;         // [0] node counter, [1] error, [2] tasks in list A, [3] tasks in list B, [4] root_ref
    Task*  tasks_a;
    Task*  tasks_b;
};

enum { C_NODES = 0, C_ERR = 1, C_TA = 2, C_TB = 3, C_ROOT = 4 };

// ---------------------------------------------------------------------------------------------------
// literal sequential libstdc++ (one thread), on (key, perm) pairs addressed by scene position
// ---------------------------------------------------------------------------------------------------
struct Elem { float k; int id; };
__device__ __forceinline__ Elem ld(const float* key, const int* perm, int p) { return Elem{key[p], perm[p]}; }
__device__ __forceinline__ void st(float* key, int* perm, int p, Elem v) { key[p] = v.k; perm[p] = v.id; }
__device__ __forceinline__ void swp(float* key, int* perm, int a, int b)
{
    float k = key[a]; key[a] = key[b]; key[b] = k;
    int i = perm[a]; perm[a] = perm[b]; perm[b] = i;
}

// stl_algo.h:1472-1496
__device__ int seq_partition(float* key, int* perm, int first, int last, float pmid)
{
    while (true) {
        while (true) {
            if (first == last) return first;
            else if (key[first] < pmid) ++first;
            else break;
        }
        --last;
        while (true) {
            if (first == last) return first;
            else if (!(key[last] < pmid)) --last;
            else break;
        }
        swp(key, perm, first, last);
        ++first;
    }
}

// stl_algo.h:82-103
__device__ void seq_move_median_to_first(float* key, int* perm, int result, int a, int b, int c)
{
    float ka = key[a], kb = key[b], kc = key[c];
    if (ka < kb) {
        if (kb < kc) swp(key, perm, result, b);
        else if (ka < kc) swp(key, perm, result, c);
        else swp(key, perm, result, a);
    } else if (ka < kc) swp(key, perm, result, a);
    else if (kb < kc) swp(key, perm, result, c);
    else swp(key, perm, result, b);
}

// stl_algo.h:1870-1888
__device__ int seq_unguarded_partition(float* key, int* perm, int first, int last, int pivot)
{
    const float v = key[pivot];   // the pivot position (first-1) is never touched by the loop
    while (true) {
        while (key[first] < v) ++first;
        --last;
        while (v < key[last]) --last;
        if (!(first < last)) return first;
        swp(key, perm, first, last);
        ++first;
    }
}

// stl_algo.h:1790-1830
__device__ void seq_insertion_sort(float* key, int* perm, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        Elem val = ld(key, perm, i);
        if (val.k < key[first]) {
            for (int j = i; j > first; --j) st(key, perm, j, ld(key, perm, j - 1));   // move_backward
            st(key, perm, first, val);
        } else {
            int lastp = i, next = i - 1;
            while (val.k < key[next]) { st(key, perm, lastp, ld(key, perm, next)); lastp = next; --next; }
            st(key, perm, lastp, val);
        }
    }
}

// stl_heap.h: __push_heap / __adjust_heap / __make_heap / __pop_heap, offsets relative to `first`
__device__ void seq_push_heap(float* key, int* perm, int first, int hole, int top, Elem value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && key[first + parent] < value.k) {
        st(key, perm, first + hole, ld(key, perm, first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    st(key, perm, first + hole, value);
}
__device__ void seq_adjust_heap(float* key, int* perm, int first, int hole, int len, Elem value)
{
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (key[first + second] < key[first + (second - 1)]) second--;
        st(key, perm, first + hole, ld(key, perm, first + second));
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        st(key, perm, first + hole, ld(key, perm, first + (second - 1)));
        hole = second - 1;
    }
    seq_push_heap(key, perm, first, hole, top, value);
}
// stl_algo.h:1624-1634
__device__ void seq_heap_select(float* key, int* perm, int first, int middle, int last)
{
    const int len = middle - first;
    if (len >= 2) {                       // __make_heap
        int parent = (len - 2) / 2;
        while (true) {
            Elem value = ld(key, perm, first + parent);
            seq_adjust_heap(key, perm, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; ++i)
        if (key[i] < key[first]) {        // __pop_heap(first, middle, i)
            Elem value = ld(key, perm, i);
            st(key, perm, i, ld(key, perm, first));
            seq_adjust_heap(key, perm, first, 0, len, value);
        }
}

// stl_algo.h:1957-1980
__device__ void seq_introselect(float* key, int* perm, int first, int nth, int last, int depth_limit)
{
    while (last - first > 3) {
        if (depth_limit == 0) {
            seq_heap_select(key, perm, first, nth + 1, last);
            swp(key, perm, first, nth);
            return;
        }
        --depth_limit;
        int mid = first + (last - first) / 2;
        seq_move_median_to_first(key, perm, first, first + 1, mid, last - 1);
        int cut = seq_unguarded_partition(key, perm, first + 1, last, first);
        if (cut <= nth) first = cut; else last = cut;
    }
    seq_insertion_sort(key, perm, first, last);
}

__device__ __forceinline__ int lg2(int n) { return 31 - __clz(n); }   // std::__lg

// GeMaximumAxis, accelerators.h:190-199
__device__ __forceinline__ int max_axis(const float mn[3], const float mx[3])
{
    float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    if (ex > ey && ex > ez) return 0;
    else if (ey > ez) return 1;
    else return 2;
}

__device__ __forceinline__ void write_box(const MedianArgs& A, int parent_enc, const float mn[3], const float mx[3], int i)
{
    // thread-parallel helper: i in [0,6) writes one float
    float v = i < 3 ? mn[i] : mx[i - 3];
    if (parent_enc < 0) A.root_box[i] = v;
    else {
        Node64* nd = A.nodes + (parent_enc >> 1);
        float* dst = (parent_enc & 1) ? (i < 3 ? nd->rmin + i : nd->rmax + (i - 3)) : (i < 3 ? nd->lmin + i : nd->lmax + (i - 3));
        *dst = v;
    }
}

__device__ __forceinline__ void link_child(const MedianArgs& A, int parent_enc, int ref)
{
    if (parent_enc < 0) A.counters[C_ROOT] = ref;
    else if (parent_enc & 1) A.nodes[parent_enc >> 1].right = ref;
    else A.nodes[parent_enc >> 1].left = ref;
}

__device__ __forceinline__ void push_task(Task* next, int* next_count, int s, int e, int parent_enc)
{
    int i = atomicAdd(next_count, 1);
    next[i] = Task{s, e, parent_enc};
}

// a size-1 child of interior node `node`: its box and its leaf record
__device__ __forceinline__ void emit_single_leaf(const MedianArgs& A, int pos, int node, int side)
{
    float cc[3], mn[3], mx[3];
    prim_fetch(A.pv, A.perm[pos], cc, mn, mx);
    for (int i = 0; i < 6; ++i) write_box(A, node * 2 + side, mn, mx, i);
    A.leaf_info[pos] = node * 2 + side + 2;
}

// ---------------------------------------------------------------------------------------------------
// one thread finishes a whole subtree with the literal sequential algorithms
// ---------------------------------------------------------------------------------------------------
__device__ void subtree_sequential(const MedianArgs& A, Task root)
{
    Task stack[40];
    int sp = 0;
    stack[sp++] = root;
    while (sp) {
        Task t = stack[--sp];
        const int s = t.s, e = t.e, m = e - s;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int p = s; p < e; ++p) {
            float cc[3], pmn[3], pmx[3];
            prim_fetch(A.pv, A.perm[p], cc, pmn, pmx);
            for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], pmn[a]); mx[a] = fmaxf(mx[a], pmx[a]); }
        }
        for (int i = 0; i < 6; ++i) write_box(A, t.parent_enc, mn, mx, i);
        if (m <= 1) { A.leaf_info[s] = t.parent_enc + 2; continue; }
        const int dim = max_axis(mn, mx);
        const float lo = mn[dim], hi = mx[dim];
        if (hi == lo) { atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_UNSUPPORTED); return; }
        const float pmid = (lo + hi) / 2;
        for (int p = s; p < e; ++p) {
            float cc[3], pmn[3], pmx[3];
            prim_fetch(A.pv, A.perm[p], cc, pmn, pmx);
            A.key[p] = cc[dim];
        }
        int mid = seq_partition(A.key, A.perm, s, e, pmid);
        if (mid != s && mid != e) {
            mid = (s + e) / 2;
            seq_introselect(A.key, A.perm, s, mid, e, lg2(m) * 2);
        }
        if (mid == s) { A.leaf_info[s] = t.parent_enc + 2; continue; }          // :321-327
        if (mid == e) { atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_DEGENERATE); return; }
        const int node = atomicAdd(&A.counters[C_NODES], 1);
        A.nodes[node].axis = dim;
        A.nodes[node].parent = t.parent_enc;
        link_child(A, t.parent_enc, node);
        // right first so the left subtree is finished first (order is irrelevant for the result)
        if (e - mid == 1) emit_single_leaf(A, mid, node, 1); else stack[sp++] = Task{mid, e, node * 2 + 1};
        if (mid - s == 1) emit_single_leaf(A, s, node, 0); else stack[sp++] = Task{s, mid, node * 2};
    }
}

__global__ void __launch_bounds__(64) median_small_kernel(const MedianArgs A, const Task* __restrict__ tasks, const int* __restrict__ count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count || A.counters[C_ERR]) return;
    subtree_sequential(A, tasks[i]);
}

// ---------------------------------------------------------------------------------------------------
// block-parallel task
// ---------------------------------------------------------------------------------------------------
template <int BLOCK>
struct BlockShared {
    int   warp_sums[BLOCK / 32];
    float red[6][BLOCK / 32];
    float box[6];
    int   ival[4];
};

// ascending compaction of the positions p in [lo,hi) with flag(p) into out[0..count)
template <int BLOCK, typename F>
__device__ int block_compact(BlockShared<BLOCK>& S, int lo, int hi, F flag, int* __restrict__ out)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int total = 0;
    for (int base = lo; base < hi; base += BLOCK) {
        const int p = base + tid;
        const bool f = p < hi && flag(p);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) S.warp_sums[warp] = __popc(b);
        __syncthreads();
        int woff = 0, chunk = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) { int c = S.warp_sums[w]; woff += (w < warp) ? c : 0; chunk += c; }
        if (f) out[total + woff + __popc(b & ((1u << lane) - 1u))] = p;
        total += chunk;
        __syncthreads();
    }
    return total;
}

template <int BLOCK>
__device__ int block_sum(BlockShared<BLOCK>& S, int v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) S.warp_sums[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) t += S.warp_sums[w];
    __syncthreads();
    return t;
}

// __introselect (stl_algo.h:1957-1980) on [first,last) by one thread block; `s` is the owning task's start (the
// scratch lists live at la+s / lb+s). Also the tail of the cooperative kernel's rounds once a range is small.
template <int BLOCK>
__device__ void block_introselect(BlockShared<BLOCK>& S, const MedianArgs& A, int s, int first, int last, int nth, int depth_limit)
{
    const int tid = threadIdx.x;
    float* key = A.key;
    int* perm = A.perm;
    while (last - first > 3) {
        if (depth_limit == 0) {
            if (tid == 0) { atomicAdd(&A.counters[7], 1); atomicAdd(&A.counters[8], last - first); }
            if (tid == 0) { seq_heap_select(key, perm, first, nth + 1, last); swp(key, perm, first, nth); }
            first = last = nth;   // done (skip the insertion sort below, as the reference returns here)
            break;
        }
        --depth_limit;
        if (tid == 0) seq_move_median_to_first(key, perm, first, first + 1, first + (last - first) / 2, last - 1);
        __syncthreads();
        const float v = key[first];
        const int lo_p = first + 1;
        int* L = A.la + s;
        int* R = A.lb + s;   // ascending; R_desc[i] = R[nR-1-i]
        const int nL = block_compact(S, lo_p, last, [&](int p) { return !(key[p] < v); }, L);
        const int nR = block_compact(S, lo_p, last, [&](int p) { return !(v < key[p]); }, R);
        const int nmin = min(nL, nR);
        int kc = 0;
        for (int i = tid; i < nmin; i += BLOCK) kc += (L[i] < R[nR - 1 - i]) ? 1 : 0;
        const int k = block_sum(S, kc);
        int cut;
        if (k == 0) cut = L[0];
        else cut = min(k < nL ? L[k] : INT_MAX, R[nR - k]);
        __syncthreads();   // everyone has read L/R heads before the swaps move keys (lists are position lists: safe)
        for (int i = tid; i < k; i += BLOCK) swp(key, perm, L[i], R[nR - 1 - i]);
        __syncthreads();
        if (cut <= nth) first = cut; else last = cut;
    }
    if (tid == 0 && last - first > 0) seq_insertion_sort(key, perm, first, last);
    __syncthreads();
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) median_block_kernel(const MedianArgs A, const Task* __restrict__ tasks,
                                                             const int* __restrict__ count, Task* __restrict__ next,
                                                             int* __restrict__ next_count)
{
    __shared__ BlockShared<BLOCK> S;
    if ((int)blockIdx.x >= *count || A.counters[C_ERR]) return;
    const Task t = tasks[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = t.s, e = t.e, m = e - s;
    float* key = A.key;
    int* perm = A.perm;

    // 1. bounds over the whole range (:257-260)
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int p = s + tid; p < e; p += BLOCK) {
            float cc[3], pmn[3], pmx[3];
            prim_fetch(A.pv, perm[p], cc, pmn, pmx);
#pragma unroll
            for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], pmn[a]); mx[a] = fmaxf(mx[a], pmx[a]); }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int o = 16; o; o >>= 1) {
                mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            }
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { S.red[a][warp] = mn[a]; S.red[3 + a][warp] = mx[a]; }
        }
        __syncthreads();
        if (tid < 6) {
            float v = S.red[tid][0];
            for (int w = 1; w < BLOCK / 32; ++w) v = tid < 3 ? fminf(v, S.red[tid][w]) : fmaxf(v, S.red[tid][w]);
            S.box[tid] = v;
        }
        __syncthreads();
    }
    float mn[3] = {S.box[0], S.box[1], S.box[2]}, mx[3] = {S.box[3], S.box[4], S.box[5]};
    if (tid < 6) write_box(A, t.parent_enc, mn, mx, tid);
    if (m <= 1) { if (tid == 0) A.leaf_info[s] = t.parent_enc + 2; return; }
    const int dim = max_axis(mn, mx);
    const float lo = mn[dim], hi = mx[dim];
    if (hi == lo) { if (tid == 0) atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_UNSUPPORTED); return; }
    const float pmid = (lo + hi) / 2;

    // 2. keys + std::partition(centre[dim] < pmid)
    int cT = 0;
    for (int p = s + tid; p < e; p += BLOCK) {
        float cc[3], pmn[3], pmx[3];
        prim_fetch(A.pv, perm[p], cc, pmn, pmx);
        float k = cc[dim];
        key[p] = k;
        cT += (k < pmid) ? 1 : 0;
    }
    cT = block_sum(S, cT);     // (also a barrier: keys are visible block-wide)
    if (cT == 0) { if (tid == 0) A.leaf_info[s] = t.parent_enc + 2; return; }                 // :321-327 drop
    if (cT == m) { if (tid == 0) atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_DEGENERATE); return; }
    {
        const int cut = s + cT;
        int* F = A.la + s;
        int* T = A.lb + s;
        const int k = block_compact(S, s, cut, [&](int p) { return !(key[p] < pmid); }, F);
        block_compact(S, cut, e, [&](int p) { return key[p] < pmid; }, T);
        for (int i = tid; i < k; i += BLOCK) swp(key, perm, F[i], T[k - 1 - i]);
        __syncthreads();
    }

    // 3. std::nth_element(first, first + m/2, last): __introselect
    const int nth = (s + e) / 2;
    block_introselect<BLOCK>(S, A, s, s, e, nth, lg2(m) * 2);

    // 4. node + children
    if (tid == 0) {
        const int node = atomicAdd(&A.counters[C_NODES], 1);
        A.nodes[node].axis = dim;
        A.nodes[node].parent = t.parent_enc;
        link_child(A, t.parent_enc, node);
        const int mid = nth;
        if (mid - s == 1) emit_single_leaf(A, s, node, 0); else push_task(next, next_count, s, mid, node * 2);
        if (e - mid == 1) emit_single_leaf(A, mid, node, 1); else push_task(next, next_count, mid, e, node * 2 + 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// cooperative kernel for the few huge ranges at the top of the tree: the WHOLE grid works on one task at a time.
// Same algorithm as median_block_kernel; every streaming pass is split over the blocks (contiguous slices, so lists
// stay in ascending position order), with grid-wide barriers between the dependent steps. Ranges that have shrunk
// below COOP_TAIL are finished by block 0 with the block-level code.
// ---------------------------------------------------------------------------------------------------
constexpr int COOP_BLOCK = 1024;
constexpr int COOP_TAIL = 16384;

struct CoopScratch {          // global, zeroed per launch
    unsigned box[6];          // ordered uints
    int      cnt[4];          // general counters
    int      blk[2][1024];    // per-block counts of the two lists
};

__device__ __forceinline__ void slice_of(int lo, int hi, int& a, int& b)
{
    const long long len = hi - lo;
    const long long chunk = (len + gridDim.x - 1) / gridDim.x;
    a = (int)min((long long)hi, lo + chunk * blockIdx.x);
    b = (int)min((long long)hi, lo + chunk * (blockIdx.x + 1));
}

// two simultaneous ascending compactions over [lo,hi): list 0 -> out0, list 1 -> out1; returns counts
template <typename F0, typename F1>
__device__ void grid_compact2(cooperative_groups::grid_group& grid, BlockShared<COOP_BLOCK>& S, CoopScratch* G, int lo, int hi, F0 f0, F1 f1,
                              int* __restrict__ out0, int* __restrict__ out1, int& n0, int& n1)
{
    int a, b;
    slice_of(lo, hi, a, b);
    int c0 = 0, c1 = 0;
    for (int p = a + threadIdx.x; p < b; p += COOP_BLOCK) { c0 += f0(p) ? 1 : 0; c1 += f1(p) ? 1 : 0; }
    c0 = block_sum(S, c0);
    c1 = block_sum(S, c1);
    if (threadIdx.x == 0) { G->blk[0][blockIdx.x] = c0; G->blk[1][blockIdx.x] = c1; }
    grid.sync();
    int o0 = 0, o1 = 0, t0 = 0, t1 = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += COOP_BLOCK) {
        int x0 = G->blk[0][i], x1 = G->blk[1][i];
        t0 += x0; t1 += x1;
        if (i < (int)blockIdx.x) { o0 += x0; o1 += x1; }
    }
    o0 = block_sum(S, o0); o1 = block_sum(S, o1); t0 = block_sum(S, t0); t1 = block_sum(S, t1);
    block_compact(S, a, b, f0, out0 + o0);
    block_compact(S, a, b, f1, out1 + o1);
    n0 = t0; n1 = t1;
    grid.sync();
}

__global__ void __launch_bounds__(COOP_BLOCK) median_coop_kernel(const MedianArgs A, const Task* __restrict__ tasks, const int* __restrict__ count,
                                                                 Task* __restrict__ next, int* __restrict__ next_count, CoopScratch* G)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ BlockShared<COOP_BLOCK> S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool lead = blockIdx.x == 0 && tid == 0;
    float* key = A.key;
    int* perm = A.perm;
    const int n_tasks = *count;
    for (int ti = 0; ti < n_tasks; ++ti) {
        if (A.counters[C_ERR]) return;          // uniform: set before a grid barrier, read after it
        const Task t = tasks[ti];
        const int s = t.s, e = t.e, m = e - s;
        if (lead) { for (int a = 0; a < 3; ++a) { G->box[a] = 0xffffffffu; G->box[3 + a] = 0u; } G->cnt[0] = 0; G->cnt[1] = 0; }
        grid.sync();
        // 1. bounds
        {
            int a, b;
            slice_of(s, e, a, b);
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int p = a + tid; p < b; p += COOP_BLOCK) {
                float cc[3], pmn[3], pmx[3];
                prim_fetch(A.pv, perm[p], cc, pmn, pmx);
#pragma unroll
                for (int x = 0; x < 3; ++x) { mn[x] = fminf(mn[x], pmn[x]); mx[x] = fmaxf(mx[x], pmx[x]); }
            }
#pragma unroll
            for (int x = 0; x < 3; ++x)
                for (int o = 16; o; o >>= 1) {
                    mn[x] = fminf(mn[x], __shfl_xor_sync(0xffffffffu, mn[x], o));
                    mx[x] = fmaxf(mx[x], __shfl_xor_sync(0xffffffffu, mx[x], o));
                }
            if (lane == 0 && a < b) {
#pragma unroll
                for (int x = 0; x < 3; ++x) { atomicMin(&G->box[x], f2ord(mn[x])); atomicMax(&G->box[3 + x], f2ord(mx[x])); }
            }
            (void)warp;
        }
        grid.sync();
        float mn[3] = {ord2f(G->box[0]), ord2f(G->box[1]), ord2f(G->box[2])}, mx[3] = {ord2f(G->box[3]), ord2f(G->box[4]), ord2f(G->box[5])};
        if (blockIdx.x == 0 && tid < 6) write_box(A, t.parent_enc, mn, mx, tid);
        const int dim = max_axis(mn, mx);
        const float lo = mn[dim], hi = mx[dim];
        if (hi == lo) { if (lead) atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_UNSUPPORTED); grid.sync(); continue; }
        const float pmid = (lo + hi) / 2;
        // 2. keys + count of pred
        {
            int a, b;
            slice_of(s, e, a, b);
            int c = 0;
            for (int p = a + tid; p < b; p += COOP_BLOCK) {
                float cc[3], pmn[3], pmx[3];
                prim_fetch(A.pv, perm[p], cc, pmn, pmx);
                const float k = cc[dim];
                key[p] = k;
                c += (k < pmid) ? 1 : 0;
            }
            c = block_sum(S, c);
            if (tid == 0 && c) atomicAdd(&G->cnt[0], c);
        }
        grid.sync();
        const int cT = G->cnt[0];
        if (cT == 0) { if (lead) A.leaf_info[s] = t.parent_enc + 2; grid.sync(); continue; }              // drop (:321-327)
        if (cT == m) { if (lead) atomicCAS(&A.counters[C_ERR], 0, RTDS_ERR_DEGENERATE); grid.sync(); continue; }
        // std::partition
        {
            const int cut = s + cT;
            int* F = A.la + s;
            int* T = A.lb + s;
            int nF, nT;
            // one pass over [s,e): positions left of the cut that fail pred -> F, positions right of it that pass -> T
            grid_compact2(grid, S, G, s, e, [&](int p) { return p < cut && !(key[p] < pmid); }, [&](int p) { return p >= cut && key[p] < pmid; },
                          F, T, nF, nT);
            const int k = nF;
            for (int i = blockIdx.x * COOP_BLOCK + tid; i < k; i += gridDim.x * COOP_BLOCK) swp(key, perm, F[i], T[k - 1 - i]);
            grid.sync();
        }
        // std::nth_element
        const int nth = (s + e) / 2;
        int first = s, last = e, depth_limit = lg2(m) * 2;
        bool done = false;
        while (last - first > COOP_TAIL) {
            if (depth_limit == 0) {
                if (lead) { atomicAdd(&A.counters[7], 1); atomicAdd(&A.counters[8], last - first); }
                if (lead) { seq_heap_select(key, perm, first, nth + 1, last); swp(key, perm, first, nth); }
                done = true;
                break;
            }
            --depth_limit;
            if (lead) { seq_move_median_to_first(key, perm, first, first + 1, first + (last - first) / 2, last - 1); G->cnt[1] = 0; }
            grid.sync();
            const float v = key[first];
            int* L = A.la + s;
            int* R = A.lb + s;
            int nL, nR;
            grid_compact2(grid, S, G, first + 1, last, [&](int p) { return !(key[p] < v); }, [&](int p) { return !(v < key[p]); }, L, R, nL, nR);
            const int nmin = min(nL, nR);
            int kc = 0;
            for (int i = blockIdx.x * COOP_BLOCK + tid; i < nmin; i += gridDim.x * COOP_BLOCK) kc += (L[i] < R[nR - 1 - i]) ? 1 : 0;
            kc = block_sum(S, kc);
            if (tid == 0 && kc) atomicAdd(&G->cnt[1], kc);
            grid.sync();
            const int k = G->cnt[1];
            int cut;
            if (k == 0) cut = L[0];
            else cut = min(k < nL ? L[k] : INT_MAX, R[nR - k]);
            for (int i = blockIdx.x * COOP_BLOCK + tid; i < k; i += gridDim.x * COOP_BLOCK) swp(key, perm, L[i], R[nR - 1 - i]);
            grid.sync();
            if (cut <= nth) first = cut; else last = cut;
        }
        if (!done && blockIdx.x == 0) block_introselect<COOP_BLOCK>(S, A, s, first, last, nth, depth_limit);
        // node + children
        if (lead) {
            const int node = atomicAdd(&A.counters[C_NODES], 1);
            A.nodes[node].axis = dim;
            A.nodes[node].parent = t.parent_enc;
            link_child(A, t.parent_enc, node);
            const int mid = nth;
            if (mid - s == 1) emit_single_leaf(A, s, node, 0); else push_task(next, next_count, s, mid, node * 2);
            if (e - mid == 1) emit_single_leaf(A, mid, node, 1); else push_task(next, next_count, mid, e, node * 2 + 1);
        }
        grid.sync();
    }
}

// ---------------------------------------------------------------------------------------------------
// finalize: rank the kept leaves (DFS order = scene-position order), fill leaf arrays and child refs
// ---------------------------------------------------------------------------------------------------
constexpr int FIN_BLOCK = 256, FIN_ITEMS = 8, FIN_TILE = FIN_BLOCK * FIN_ITEMS;

__global__ void __launch_bounds__(FIN_BLOCK) leaf_count_kernel(const int* __restrict__ leaf_info, int n, int* __restrict__ tile_counts)
{
    __shared__ int ws[FIN_BLOCK / 32];
    int c = 0;
    const int base = blockIdx.x * FIN_TILE;
    for (int i = threadIdx.x; i < FIN_TILE; i += FIN_BLOCK) { int p = base + i; c += (p < n && leaf_info[p] != 0) ? 1 : 0; }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < FIN_BLOCK / 32; ++w) t += ws[w]; tile_counts[blockIdx.x] = t; }
}

__global__ void tile_scan_kernel(int* tile_counts, int tiles, int* total)
{
    // tiles <= a few thousand: one thread is enough
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < tiles; ++i) { int c = tile_counts[i]; tile_counts[i] = run; run += c; }
        *total = run;
    }
}

__global__ void __launch_bounds__(FIN_BLOCK)
leaf_emit_kernel(const int* __restrict__ leaf_info, const int* __restrict__ perm, const PrimView pv, int n,
                 const int* __restrict__ tile_offsets, Node64* nodes, float4* __restrict__ leaf_sph, float4* __restrict__ leaf_tri,
                 int* __restrict__ prim_order, int* __restrict__ leaf_parent, int* __restrict__ counters)
{
    __shared__ int ws[FIN_BLOCK / 32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = tile_offsets[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * FIN_TILE;
    for (int it = 0; it < FIN_ITEMS; ++it) {
        const int p = base + it * FIN_BLOCK + threadIdx.x;
        const int info = p < n ? leaf_info[p] : 0;
        const bool f = info != 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) ws[warp] = __popc(b);
        __syncthreads();
        int woff = 0, chunk = 0;
        for (int w = 0; w < FIN_BLOCK / 32; ++w) { int c = ws[w]; woff += (w < warp) ? c : 0; chunk += c; }
        const int start = running;
        if (f) {
            const int r = start + woff + __popc(b & ((1u << lane) - 1u));
            const int obj = perm[p];
            prim_store_leaf(pv, obj, r, leaf_sph, leaf_tri);
            prim_order[r] = obj;
            const int pe = info - 2;
            if (pe < 0) { leaf_parent[r] = 0; counters[C_ROOT] = ~r; }
            else {
                leaf_parent[r] = (pe >> 1) | ((pe & 1) ? 0x80000000 : 0);
                if (pe & 1) nodes[pe >> 1].right = ~r; else nodes[pe >> 1].left = ~r;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) running = start + chunk;
        __syncthreads();
    }
}

__global__ void iota_kernel(int* p, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace

int rtds_bvh_compute_depth(rtds_ctx* ctx, DeviceBvh& b, int* depth_out);   // lbvh.cu
int rtds_bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches);

int rtds_build_median(rtds_ctx* ctx, int n_use, rtds_build_stats* st)
{
    const int n = n_use;
    DeviceBvh& b = ctx->bvh;
    b.valid = false;
    int small = 64;
    if (ctx->opt.median_small > 0) small = ctx->opt.median_small;   // test hook (median_small option): huge -> all sequential
    RTDS_TRY(rtds_alloc_bvh_for(ctx, b, n));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const int fin_tiles = (n + FIN_TILE - 1) / FIN_TILE;
    size_t o_perm = 0, o_key = o_perm + al(4 * (size_t)n), o_la = o_key + al(4 * (size_t)n), o_lb = o_la + al(4 * (size_t)n),
           o_info = o_lb + al(4 * (size_t)n), o_ta = o_info + al(4 * (size_t)n), o_tb = o_ta + al(sizeof(Task) * ((size_t)n / 2 + 2)),
           o_cnt = o_tb + al(sizeof(Task) * ((size_t)n / 2 + 2)), o_root = o_cnt + 256, o_tiles = o_root + 256,
           o_coop = o_tiles + al(4 * (size_t)fin_tiles + 4), total = o_coop + al(sizeof(CoopScratch));
    total = std::max(total, al(4 * (size_t)n) + sizeof(Node64) * (size_t)n + 1024);   // the final layout pass reuses the area
    RTDS_TRY(rtds_ensure_scratch(ctx, total));
    char* base = (char*)ctx->d_scratch;
    MedianArgs A;
    A.pv = rtds_prim_view(ctx);
    A.perm = (int*)(base + o_perm); A.key = (float*)(base + o_key); A.la = (int*)(base + o_la); A.lb = (int*)(base + o_lb);
    A.leaf_info = (int*)(base + o_info); A.nodes = b.nodes; A.root_box = (float*)(base + o_root); A.counters = (int*)(base + o_cnt);
    A.tasks_a = (Task*)(base + o_ta); A.tasks_b = (Task*)(base + o_tb);
    int* d_tiles = (int*)(base + o_tiles);

    cudaStream_t s = ctx->stream;
    int launches = 0;
    RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    RTDS_CUDA(cudaMemsetAsync(A.leaf_info, 0, 4 * (size_t)n, s));
    RTDS_CUDA(cudaMemsetAsync(A.counters, 0, 256, s));
    iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(A.perm, n);
    ++launches;
    const Task root{0, n, -1};
    const int one = 1;
    RTDS_CUDA(cudaMemcpyAsync(A.tasks_a, &root, sizeof root, cudaMemcpyHostToDevice, s));
    RTDS_CUDA(cudaMemcpyAsync(A.counters + C_TA, &one, sizeof one, cudaMemcpyHostToDevice, s));

    // level loop: range sizes at level L are <= ceil(n / 2^L) (smaller after drops), task count <= 2^L
    Task* cur = A.tasks_a; Task* nxt = A.tasks_b;
    int ccur = C_TA, cnxt = C_TB;
    long long max_size = n, max_tasks = 1;
    int coop_threshold = 65536;     // ranges larger than this get the whole grid, one after another
    if (ctx->opt.median_coop != 0) coop_threshold = ctx->opt.median_coop > 0 ? ctx->opt.median_coop : (1 << 30);
    int coop_blocks = 0;
    {
        int per_sm = 0;
        RTDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, median_coop_kernel, COOP_BLOCK, 0));
        coop_blocks = std::min(ctx->sm_count * std::max(per_sm, 0), 1024);
        int can = 0;
        RTDS_CUDA(cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, ctx->device));
        if (!can) coop_blocks = 0;
    }
    CoopScratch* d_coop = (CoopScratch*)(base + o_coop);
    while (max_size > small) {
        RTDS_CUDA(cudaMemsetAsync(A.counters + cnxt, 0, sizeof(int), s));
        const int grid = (int)std::min<long long>(max_tasks, (long long)n / 2 + 1);
        if (max_size > coop_threshold && coop_blocks > 0 && small < coop_threshold) {
            RTDS_CUDA(cudaMemsetAsync(d_coop, 0, sizeof(CoopScratch), s));
            const Task* a_tasks = cur; const int* a_count = A.counters + ccur; Task* a_next = nxt; int* a_ncount = A.counters + cnxt;
            void* args[] = {(void*)&A, (void*)&a_tasks, (void*)&a_count, (void*)&a_next, (void*)&a_ncount, (void*)&d_coop};
            RTDS_CUDA(cudaLaunchCooperativeKernel((void*)median_coop_kernel, dim3(coop_blocks), dim3(COOP_BLOCK), args, 0, s));
        }
        else if (max_size > 4096) median_block_kernel<1024><<<grid, 1024, 0, s>>>(A, cur, A.counters + ccur, nxt, A.counters + cnxt);
        else if (max_size > 512) median_block_kernel<256><<<grid, 256, 0, s>>>(A, cur, A.counters + ccur, nxt, A.counters + cnxt);
        else median_block_kernel<64><<<grid, 64, 0, s>>>(A, cur, A.counters + ccur, nxt, A.counters + cnxt);
        ++launches;
        std::swap(cur, nxt);
        std::swap(ccur, cnxt);
        max_size = (max_size + 1) / 2;
        max_tasks *= 2;
    }
    {
        const int grid_threads = (int)std::min<long long>(max_tasks, (long long)n / 2 + 1);
        median_small_kernel<<<(grid_threads + 63) / 64, 64, 0, s>>>(A, cur, A.counters + ccur);
        ++launches;
    }
    leaf_count_kernel<<<fin_tiles, FIN_BLOCK, 0, s>>>(A.leaf_info, n, d_tiles);
    tile_scan_kernel<<<1, 32, 0, s>>>(d_tiles, fin_tiles, A.counters + 5);
    leaf_emit_kernel<<<fin_tiles, FIN_BLOCK, 0, s>>>(A.leaf_info, A.perm, A.pv, n, d_tiles, b.nodes, b.leaf_sph, b.leaf_tri, b.prim_order,
                                                      b.leaf_parent, A.counters);
    launches += 3;
    RTDS_CUDA(cudaGetLastError());
    int h_cnt[12];
    RTDS_CUDA(cudaMemcpyAsync(h_cnt, A.counters, sizeof h_cnt, cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaMemcpyAsync(b.root_box, A.root_box, sizeof(float) * 6, cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    if (h_cnt[C_ERR] == RTDS_ERR_DEGENERATE) {
        rtds_set_error("median-split build: std::partition returned endIndex for a range — the reference recurses forever on this input (accelerators.h:311-330)");
        return RTDS_ERR_DEGENERATE;
    }
    if (h_cnt[C_ERR] != 0) {
        rtds_set_error("median-split build: a range has zero extent on its longest axis (accelerators.h:286-293 pushes unrelated ids); unsupported");
        return RTDS_ERR_UNSUPPORTED;
    }
    if (ctx->opt.median_debug) fprintf(stderr, "[median] heap-select fallbacks: %d (elements %d)\n", h_cnt[7], h_cnt[8]);
    const int n_leaves = h_cnt[5], n_internal = h_cnt[C_NODES];
    if (n_leaves != n_internal + 1) { rtds_set_error("median-split build: internal error (%d leaves, %d interior)", n_leaves, n_internal); return RTDS_ERR_CUDA; }
    b.n_prims = n_leaves;
    b.n_internal = n_internal;
    b.root_ref = h_cnt[C_ROOT];
    b.tie_by_objid = 0;
    b.leaf_box_prim = (ctx->prim_type == 0 && n_leaves == n) ? 1 : 0;   // no dropped range: every leaf is one sphere with its own box
    // layout pass; the level-loop scratch (keys, lists, task arrays: > 68 bytes per primitive from the start) is free now
    RTDS_TRY(rtds_bvh_reorder_preorder(ctx, b, base, &launches));
    int depth = 0;
    RTDS_TRY(rtds_bvh_compute_depth(ctx, b, &depth));
    ++launches;
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    b.max_depth = depth;
    b.valid = true;
    float ms = 0;
    RTDS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (st) {
        st->n_prims = n_leaves;
        st->total_nodes = n_leaves + n_internal;
        st->alloc_nodes = st->total_nodes;
        st->max_depth = depth;
        st->kernel_launches = launches;
        st->ms = ms;
    }
    return RTDS_OK;
}
