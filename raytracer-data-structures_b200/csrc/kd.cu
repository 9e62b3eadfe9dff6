// kd.cu — K8: the reference's KD-tree builder (buildTree / constructKDTreeNew, accelerators.h:815-988, a port of
// PBRT-v3's KdTreeAccel) on the GPU, level-synchronous:
//   per level, for all nodes at once:  leaf test (:843) -> edges of every primitive on the node's axis (:865-870)
//   -> ONE onesweep sort of all edges by (node, t, Start<End) (:873-879) -> SAH cost of every edge from prefix
//   counts (:882-912, same float expression, so the same costs bit for bit) -> per-node first minimum (atomicMin on
//   {cost bits, position}) -> retry on the next axis (:915-919) -> leaf / interior decision (:920-925)
//   -> classification into the children's primitive lists by segmented compaction (:928-934).
// The finished tree is relabelled into the reference's depth-first layout (left child = index+1, right child =
// aboveChild) and emitted as 12-byte KdAccelNode records + kdtreePrimitiveIndices (accelerators.h:715-773).
//
// Effective parameters are the reference's GLOBALS (accelerators.h:767-770): maxPrims 5, isectCost 80,
// traversalCost 1, emptyBonus (char)0.5f = 0 — not the arguments main() passes (:951-954 shadows them).
// std::sort is unstable: among edges with equal (t, type) the reference's order is implementation-defined, so which
// primitive sits at a tied split position may differ; costs, split positions and counts do not depend on it.
// Parity for KD is therefore judged on hit results — node counts are reported next to the reference's.
#include "rtds_internal.cuh"
#include "scan.cuh"
#include <math.h>
#include <algorithm>

int rtds_scene_bounds(rtds_ctx* ctx, float out12[12]);   // lbvh.cu

namespace {

struct KdNode {                 // build-time node, breadth-first order
    float bmin[3], bmax[3];
    int   parent, is_right;
    int   prim_off, prim_cnt;   // list in the current level's index buffer
    int   depth_left, bad_refines;
    int   state;                // 0 pending, 1 leaf (counted in totalKdNodes), 2 leaf (not counted), 3 interior
    int   axis, retries;
    float split;
    int   left, right;          // BFS ids of the children
    int   slot;                 // index among the level's active nodes, -1 if not active
    int   edge_start;           // first edge of this node in the level's edge array
    int   best_off;             // position of the chosen edge inside the node's sorted edges
    int   n0, n1;
    int   leaf_off;             // offset of the leaf's list in the leaf pool
    int   size, dfs;            // subtree size, depth-first index
    int   pad;
};

struct KdParams { int max_prims; float isect_cost; float traversal_cost; float empty_bonus; };

using rtds_scan::Scanner;
using rtds_scan::SC_TILE;

// ---------------------------------------------------------------------------------------------------
// level kernels
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int kd_max_axis(const KdNode& nd)   // GeMaximumAxis, accelerators.h:190-199
{
    float ex = nd.bmax[0] - nd.bmin[0], ey = nd.bmax[1] - nd.bmin[1], ez = nd.bmax[2] - nd.bmin[2];
    if (ex > ey && ex > ez) return 0;
    else if (ey > ez) return 1;
    else return 2;
}

// :843 termination test; flags[i] = 1 for nodes that go on to the split search, cnt2[i] = their edge count
__global__ void kd_level_begin(KdNode* nodes, int lvl_begin, int lvl_n, KdParams P, int* flags, int* cnt2, int* leafcnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_n) return;
    KdNode& nd = nodes[lvl_begin + i];
    bool leaf = nd.prim_cnt <= P.max_prims || nd.depth_left == 0;
    if (leaf) { nd.state = 1; nd.slot = -1; }
    else { nd.state = 0; nd.axis = kd_max_axis(nd); nd.retries = 0; nd.best_off = -1; }
    flags[i] = leaf ? 0 : 1;
    cnt2[i] = leaf ? 0 : 2 * nd.prim_cnt;
    leafcnt[i] = leaf ? nd.prim_cnt : 0;
}

__global__ void kd_assign_slots(KdNode* nodes, int lvl_begin, int lvl_n, const int* flags, const int* slot_scan, const int* edge_scan,
                                int* slot_to_node)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_n) return;
    if (flags[i]) {
        KdNode& nd = nodes[lvl_begin + i];
        nd.slot = slot_scan[i];
        nd.edge_start = edge_scan[i];
        slot_to_node[nd.slot] = lvl_begin + i;
    }
}

// :865-870 — one thread per list entry of the level
__global__ void kd_gen_edges(const KdNode* __restrict__ nodes, const int* __restrict__ idx, const int* __restrict__ owner, int n_entries,
                             const PrimView pv, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_entries) return;
    const KdNode& nd = nodes[owner[j]];
    if (nd.state != 0) return;
    int prim = idx[j];
    float cc[3], mn[3], mx[3];
    prim_fetch(pv, prim, cc, mn, mx);
    float lo = mn[nd.axis], hi = mx[nd.axis];
    int e = nd.edge_start + 2 * (j - nd.prim_off);
    unsigned long long slot = (unsigned long long)nd.slot << 33;
    keys[e] = slot | ((unsigned long long)f2ord(lo) << 1) | 0ull;      // Start
    keys[e + 1] = slot | ((unsigned long long)f2ord(hi) << 1) | 1ull;  // End
    vals[e] = (uint32_t)prim;
    vals[e + 1] = (uint32_t)prim;
}

__global__ void kd_start_flags(const unsigned long long* __restrict__ keys, int n_edges, int* __restrict__ flags)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) flags[e] = (keys[e] & 1ull) ? 0 : 1;
}

__global__ void kd_reset_best(unsigned long long* best, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) best[i] = ~0ull; }

// :882-912 — cost of splitting at every edge, first minimum per node
__global__ void kd_sweep(const KdNode* __restrict__ nodes, const int* __restrict__ slot_to_node, const unsigned long long* __restrict__ keys,
                         const int* __restrict__ start_scan, int n_edges, KdParams P, unsigned long long* __restrict__ best)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned long long k = keys[e];
    int slot = (int)(k >> 33);
    const KdNode& nd = nodes[slot_to_node[slot]];
    const int axis = nd.axis;
    const bool is_end = (k & 1ull) != 0;
    const float edgeT = ord2f((unsigned)((k >> 1) & 0xffffffffull));
    const int pos = e - nd.edge_start;
    const int starts_before = start_scan[e] - start_scan[nd.edge_start];
    const int nBelow = starts_before;
    const int ends_incl = (pos + 1) - (starts_before + (is_end ? 0 : 1));
    const int nAbove = nd.prim_cnt - ends_incl;
    if (edgeT > nd.bmin[axis] && edgeT < nd.bmax[axis]) {
        const float d[3] = {nd.bmax[0] - nd.bmin[0], nd.bmax[1] - nd.bmin[1], nd.bmax[2] - nd.bmin[2]};
        const float totalSA = 2 * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2]);     // accelerators.h:122-125
        const float invTotalSA = 1 / totalSA;
        const int o0 = (axis + 1) % 3, o1 = (axis + 2) % 3;
        float belowSA = 2 * (d[o0] * d[o1] + (edgeT - nd.bmin[axis]) * (d[o0] + d[o1]));
        float aboveSA = 2 * (d[o0] * d[o1] + (nd.bmax[axis] - edgeT) * (d[o0] + d[o1]));
        float pBelow = belowSA * invTotalSA;
        float pAbove = aboveSA * invTotalSA;
        float eb = (nAbove == 0 || nBelow == 0) ? P.empty_bonus : 0;
        float cost = P.traversal_cost + P.isect_cost * (1 - eb) * (pBelow * nBelow + pAbove * nAbove);
        if (cost < INFINITY) {   // `cost < bestCost` with bestCost = INFINITY initially; NaN never wins
            // costs are non-negative here: IEEE bits order like the values; equal costs -> lowest position
            unsigned long long key = ((unsigned long long)__float_as_uint(cost < 0 ? 0.0f : cost) << 32) | (unsigned)pos;
            atomicMin(&best[slot], key);
        }
    }
}

// :915-925 — retry / leaf / interior
__global__ void kd_decide(KdNode* nodes, const int* __restrict__ slot_to_node, int n_slots, const unsigned long long* __restrict__ best,
                          const unsigned long long* __restrict__ keys, const int* __restrict__ start_scan, KdParams P, int final_round,
                          int* n_retry, int* interior_flags, int* child_entries)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    KdNode& nd = nodes[slot_to_node[s]];
    if (nd.state != 0) { if (final_round) { interior_flags[s] = nd.state == 3; child_entries[s] = nd.state == 3 ? nd.n0 + nd.n1 : 0; } return; }
    unsigned long long b = best[s];
    bool found = b != ~0ull;
    if (!found && nd.retries < 2 && !final_round) {
        nd.retries++;
        nd.axis = (nd.axis + 1) % 3;
        atomicAdd(n_retry, 1);
        return;
    }
    float bestCost = found ? __uint_as_float((unsigned)(b >> 32)) : INFINITY;
    float oldCost = P.isect_cost * float(nd.prim_cnt);
    int bad = nd.bad_refines;
    if (bestCost > oldCost) ++bad;
    bool leaf = (bestCost > 4 * oldCost && nd.prim_cnt < 16) || !found || bad == 3;
    if (leaf) { nd.state = 2; }
    else {
        int pos = (int)(b & 0xffffffffull);
        int e = nd.edge_start + pos;
        unsigned long long k = keys[e];
        bool is_end = (k & 1ull) != 0;
        int starts_before = start_scan[e] - start_scan[nd.edge_start];
        int ends_incl = (pos + 1) - (starts_before + (is_end ? 0 : 1));
        nd.state = 3;
        nd.best_off = pos;
        nd.split = ord2f((unsigned)((k >> 1) & 0xffffffffull));
        nd.n0 = starts_before;                 // Start edges before bestOffset (:929-931)
        nd.n1 = nd.prim_cnt - ends_incl;       // End edges after bestOffset (:932-934)
        nd.bad_refines = bad;
    }
    // the caller re-runs this kernel with final_round = 1 to publish flags once no node is left pending
}

// children flags per edge: left = Start before best, right = End after best
__global__ void kd_child_flags(const KdNode* __restrict__ nodes, const int* __restrict__ slot_to_node, const unsigned long long* __restrict__ keys,
                               int n_edges, int* __restrict__ lflags, int* __restrict__ rflags)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned long long k = keys[e];
    const KdNode& nd = nodes[slot_to_node[(int)(k >> 33)]];
    int pos = e - nd.edge_start;
    bool is_end = (k & 1ull) != 0;
    bool interior = nd.state == 3;
    lflags[e] = (interior && !is_end && pos < nd.best_off) ? 1 : 0;
    rflags[e] = (interior && is_end && pos > nd.best_off) ? 1 : 0;
}

// create the children of interior nodes (BFS ids next_begin + 2*rank, +1) and their list ranges in the next buffer
__global__ void kd_make_children(KdNode* nodes, const int* __restrict__ slot_to_node, int n_slots, const int* __restrict__ interior_scan,
                                 const int* __restrict__ entry_scan, int next_begin)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    int ni = slot_to_node[s];
    KdNode& nd = nodes[ni];
    if (nd.state != 3) return;
    int l = next_begin + 2 * interior_scan[s], r = l + 1;
    nd.left = l; nd.right = r;
    KdNode a = nd, b = nd;
    a.parent = ni; a.is_right = 0; a.prim_off = entry_scan[s]; a.prim_cnt = nd.n0; a.depth_left = nd.depth_left - 1;
    a.bmax[nd.axis] = nd.split;                                   // bounds0.max[bestAxis] = tSplit (:938-939)
    b.parent = ni; b.is_right = 1; b.prim_off = entry_scan[s] + nd.n0; b.prim_cnt = nd.n1; b.depth_left = nd.depth_left - 1;
    b.bmin[nd.axis] = nd.split;
    a.state = b.state = 0; a.left = a.right = b.left = b.right = -1; a.slot = b.slot = -1; a.leaf_off = b.leaf_off = -1;
    nodes[l] = a;
    nodes[r] = b;
}

__global__ void kd_scatter_children(const KdNode* __restrict__ nodes, const int* __restrict__ slot_to_node, const unsigned long long* __restrict__ keys,
                                    const uint32_t* __restrict__ vals, int n_edges, const int* __restrict__ lscan, const int* __restrict__ rscan,
                                    const int* __restrict__ lflags, const int* __restrict__ rflags, int* __restrict__ idx_next, int* __restrict__ owner_next)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const KdNode& nd = nodes[slot_to_node[(int)(keys[e] >> 33)]];
    if (nd.state != 3) return;
    if (lflags[e]) {
        int dst = nodes[nd.left].prim_off + (lscan[e] - lscan[nd.edge_start]);
        idx_next[dst] = (int)vals[e];
        owner_next[dst] = nd.left;
    }
    if (rflags[e]) {
        int dst = nodes[nd.right].prim_off + (rscan[e] - rscan[nd.edge_start]);
        idx_next[dst] = (int)vals[e];
        owner_next[dst] = nd.right;
    }
}

// leaves decided at this level copy their lists into the persistent pool
__global__ void kd_leaf_counts(const KdNode* nodes, int lvl_begin, int lvl_n, int* cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lvl_n) { const KdNode& nd = nodes[lvl_begin + i]; cnt[i] = (nd.state == 1 || nd.state == 2) ? nd.prim_cnt : 0; }
}
__global__ void kd_leaf_offsets(KdNode* nodes, int lvl_begin, int lvl_n, const int* scan, int pool_base)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lvl_n) { KdNode& nd = nodes[lvl_begin + i]; if (nd.state == 1 || nd.state == 2) nd.leaf_off = pool_base + scan[i]; }
}
__global__ void kd_leaf_copy(const KdNode* __restrict__ nodes, const int* __restrict__ idx, const int* __restrict__ owner, int n_entries, int* __restrict__ pool)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_entries) return;
    const KdNode& nd = nodes[owner[j]];
    if (nd.state == 1 || nd.state == 2) pool[nd.leaf_off + (j - nd.prim_off)] = idx[j];
}

// ---------------------------------------------------------------------------------------------------
// depth-first relabel + emission in the reference's layout
// ---------------------------------------------------------------------------------------------------
__global__ void kd_sizes(KdNode* nodes, int lvl_begin, int lvl_n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_n) return;
    KdNode& nd = nodes[lvl_begin + i];
    nd.size = nd.state == 3 ? 1 + nodes[nd.left].size + nodes[nd.right].size : 1;
}
__global__ void kd_dfs(KdNode* nodes, int lvl_begin, int lvl_n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_n) return;
    KdNode& nd = nodes[lvl_begin + i];
    if (nd.parent < 0) nd.dfs = 0;
    if (nd.state == 3) {
        nodes[nd.left].dfs = nd.dfs + 1;
        nodes[nd.right].dfs = nd.dfs + 1 + nodes[nd.left].size;
    }
}
__global__ void kd_leaf_sizes_by_dfs(const KdNode* nodes, int n_nodes, int* arr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const KdNode& nd = nodes[i];
    arr[nd.dfs] = (nd.state != 3 && nd.prim_cnt > 1) ? nd.prim_cnt : 0;
}
// KdAccelNode::InitLeaf / InitInterior (accelerators.h:718-761)
__global__ void kd_emit(const KdNode* __restrict__ nodes, int n_nodes, const int* __restrict__ off_by_dfs, const int* __restrict__ pool,
                        rtds_kd_node* __restrict__ out, int* __restrict__ out_idx, int* __restrict__ counted)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const KdNode& nd = nodes[i];
    rtds_kd_node r;
    if (nd.state == 3) {
        r.w0 = __float_as_uint(nd.split);
        r.w1 = (unsigned)nd.axis | ((unsigned)nodes[nd.right].dfs << 2);
        r.w2 = 0;
        atomicAdd(counted, 1);                       // ++totalKdNodes (:944)
    } else {
        int np = nd.prim_cnt;
        r.w1 = 3u | ((unsigned)np << 2);
        r.w2 = (unsigned)np;
        if (np == 0) r.w0 = 0;
        else if (np == 1) r.w0 = (unsigned)pool[nd.leaf_off];
        else {
            int off = off_by_dfs[nd.dfs];
            r.w0 = (unsigned)off;
            for (int k = 0; k < np; ++k) out_idx[off + k] = pool[nd.leaf_off + k];
        }
        if (nd.state == 1) atomicAdd(counted, 1);    // ++totalKdNodes only on the termination-test path (:843-847)
    }
    out[nd.dfs] = r;
}

__global__ void kd_init_root(KdNode* nodes, const float* bounds6, int n, int depth, int* idx, int* owner)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        KdNode r;
        memset(&r, 0, sizeof r);
        for (int a = 0; a < 3; ++a) { r.bmin[a] = bounds6[a]; r.bmax[a] = bounds6[3 + a]; }
        r.parent = -1; r.prim_off = 0; r.prim_cnt = n; r.depth_left = depth; r.bad_refines = 0; r.left = r.right = -1; r.slot = -1; r.leaf_off = -1;
        nodes[0] = r;
    }
    if (i < n) { idx[i] = i; owner[i] = 0; }
}

template <typename T> T* carve(char*& p, size_t count)
{
    T* r = (T*)p;
    p += (sizeof(T) * count + 255) & ~(size_t)255;
    return r;
}

}  // namespace

static int build_kd_attempt(rtds_ctx* ctx, const rtds_build_params* bp, rtds_build_stats* st, size_t entries_per_prim)
{
    const int n = ctx->n;
    KdParams P;
    P.max_prims = (bp && bp->kd_max_prims > 0) ? bp->kd_max_prims : 5;
    P.isect_cost = (bp && bp->kd_isect_cost > 0) ? (float)bp->kd_isect_cost : 80.0f;
    P.traversal_cost = (bp && bp->kd_traversal_cost > 0) ? (float)bp->kd_traversal_cost : 1.0f;
    P.empty_bonus = (bp && bp->kd_empty_bonus > 0) ? bp->kd_empty_bonus : 0.0f;
    int max_depth = bp ? bp->kd_max_depth : 0;
    if (max_depth <= 0) {   // accelerators.h:957-958: round(8 + 1.3f * Log2Int(n))
        int lg = 0; { unsigned v = (unsigned)n; while (v > 1) { v >>= 1; ++lg; } }
        max_depth = (int)std::round(8 + 1.3f * lg);
    }
    // kdtreeIntersect's todo stack holds 64 entries (accelerators.h:1008: KdToDo todo[64]) and so does the device traversal: a
    // deeper tree could silently lose subtrees, so it is refused here instead of rendered wrongly
    if (max_depth > 64) {
        rtds_set_error("build: kd_max_depth %d exceeds the 64-entry traversal stack (accelerators.h:1008)", max_depth);
        return RTDS_ERR_UNSUPPORTED;
    }
    rtds_free_kd(ctx->kd);
    cudaStream_t s = ctx->stream;
    int launches = 0;

    // capacities: list entries per level are bounded in practice by a small multiple of n (PBRT reserves (depth+1)*n)
    const size_t cap_entries = std::max<size_t>((size_t)n * entries_per_prim, 1 << 16);
    const size_t cap_edges = 2 * cap_entries;
    const size_t cap_nodes = std::max<size_t>((size_t)n * std::min<size_t>(entries_per_prim, 48), 1 << 12);
    const int scan_tiles = (int)((cap_edges + SC_TILE - 1) / SC_TILE) + 1;
    size_t bytes = 0;
    auto add = [&](size_t b) { bytes += (b + 255) & ~(size_t)255; };
    add(sizeof(KdNode) * cap_nodes);
    add(4 * cap_entries); add(4 * cap_entries); add(4 * cap_entries); add(4 * cap_entries);  // idx/owner x2
    add(8 * cap_edges); add(8 * cap_edges); add(4 * cap_edges); add(4 * cap_edges);          // keys x2, vals x2
    add(4 * cap_edges); add(4 * cap_edges); add(4 * cap_edges);                              // flag/scan buffers
    add(4 * cap_edges); add(4 * cap_edges);
    add(4 * cap_nodes); add(4 * cap_nodes); add(4 * cap_nodes); add(4 * cap_nodes); add(4 * cap_nodes); add(4 * cap_nodes);
    add(8 * cap_nodes);                                                                       // best
    add(4 * cap_entries);                                                                     // leaf pool
    add(4 * (size_t)scan_tiles); add(1024);
    RTDS_TRY(rtds_ensure_scratch(ctx, bytes + 4096));
    char* p = (char*)ctx->d_scratch;
    KdNode* nodes = carve<KdNode>(p, cap_nodes);
    int* idx[2] = {carve<int>(p, cap_entries), carve<int>(p, cap_entries)};
    int* owner[2] = {carve<int>(p, cap_entries), carve<int>(p, cap_entries)};
    unsigned long long* keys = carve<unsigned long long>(p, cap_edges);
    unsigned long long* keys_tmp = carve<unsigned long long>(p, cap_edges);
    uint32_t* vals = carve<uint32_t>(p, cap_edges);
    uint32_t* vals_tmp = carve<uint32_t>(p, cap_edges);
    int* eflag = carve<int>(p, cap_edges);
    int* start_scan = carve<int>(p, cap_edges);
    int* lflags = carve<int>(p, cap_edges);
    int* rflags = carve<int>(p, cap_edges);
    int* escan2 = carve<int>(p, cap_edges);
    int* nflag = carve<int>(p, cap_nodes);
    int* ncnt = carve<int>(p, cap_nodes);
    int* nscan_a = carve<int>(p, cap_nodes);
    int* nscan_b = carve<int>(p, cap_nodes);
    int* slot_to_node = carve<int>(p, cap_nodes);
    int* nleaf = carve<int>(p, cap_nodes);
    unsigned long long* best = carve<unsigned long long>(p, cap_nodes);
    int* pool = carve<int>(p, cap_entries);
    int* scan_sums_buf = carve<int>(p, scan_tiles);
    int* d_small = carve<int>(p, 64);   // [0] scan total, [1] retry counter, [2] counted nodes
    Scanner scan{ctx, scan_sums_buf, scan_tiles, d_small};

    RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    float b12[12];
    RTDS_TRY(rtds_scene_bounds(ctx, b12));
    launches += 2;
    float* d_bounds6 = (float*)(d_small + 16);
    RTDS_CUDA(cudaMemcpyAsync(d_bounds6, b12 + 6, sizeof(float) * 6, cudaMemcpyHostToDevice, s));
    kd_init_root<<<(n + 255) / 256, 256, 0, s>>>(nodes, d_bounds6, n, max_depth, idx[0], owner[0]);
    ++launches;

    std::vector<std::pair<int, int>> levels;   // (begin, count) in BFS order
    int lvl_begin = 0, lvl_n = 1, n_entries = n, cur = 0, total_nodes = 1, pool_used = 0;
    const int T = 256;
    auto G = [&](long long m) { return (unsigned)((m + T - 1) / T); };
    auto read_total = [&](int* out) -> int {
        RTDS_CUDA(cudaMemcpyAsync(out, d_small, sizeof(int), cudaMemcpyDeviceToHost, s));
        RTDS_CUDA(cudaStreamSynchronize(s));
        return RTDS_OK;
    };
    while (lvl_n > 0) {
        levels.emplace_back(lvl_begin, lvl_n);
        kd_level_begin<<<G(lvl_n), T, 0, s>>>(nodes, lvl_begin, lvl_n, P, nflag, ncnt, nleaf);
        ++launches;
        int n_slots = 0, n_edges = 0;
        RTDS_TRY(scan.run(nflag, nscan_a, lvl_n, &launches));
        RTDS_TRY(read_total(&n_slots));
        RTDS_TRY(scan.run(ncnt, nscan_b, lvl_n, &launches));
        RTDS_TRY(read_total(&n_edges));
        int next_n = 0, next_entries = 0;
        if (n_slots > 0) {
            kd_assign_slots<<<G(lvl_n), T, 0, s>>>(nodes, lvl_begin, lvl_n, nflag, nscan_a, nscan_b, slot_to_node);
            ++launches;
            int slot_bits = 1; while ((1 << slot_bits) < n_slots) ++slot_bits;
            for (int round = 0; round < 3; ++round) {
                kd_gen_edges<<<G(n_entries), T, 0, s>>>(nodes, idx[cur], owner[cur], n_entries, rtds_prim_view(ctx), keys, vals);
                ++launches;
                RTDS_TRY(rtds_onesweep_sort_u64(ctx, (uint64_t*)keys, vals, (uint64_t*)keys_tmp, vals_tmp, n_edges, 33 + slot_bits, &launches));
                kd_start_flags<<<G(n_edges), T, 0, s>>>(keys, n_edges, eflag);
                RTDS_TRY(scan.run(eflag, start_scan, n_edges, &launches));
                kd_reset_best<<<G(n_slots), T, 0, s>>>(best, n_slots);
                kd_sweep<<<G(n_edges), T, 0, s>>>(nodes, slot_to_node, keys, start_scan, n_edges, P, best);
                RTDS_CUDA(cudaMemsetAsync(d_small + 1, 0, sizeof(int), s));
                kd_decide<<<G(n_slots), T, 0, s>>>(nodes, slot_to_node, n_slots, best, keys, start_scan, P, 0, d_small + 1, nflag, ncnt);
                launches += 4;
                int n_retry = 0;
                RTDS_CUDA(cudaMemcpyAsync(&n_retry, d_small + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
                RTDS_CUDA(cudaStreamSynchronize(s));
                if (n_retry == 0) break;
                // Later rounds regenerate the edges of the still-pending nodes only; a decided node's segment stays
                // sorted in place (the sort is stable and groups by slot), so its best_off keeps pointing at its edge.
            }
            // publish interior flags / child entry counts
            kd_decide<<<G(n_slots), T, 0, s>>>(nodes, slot_to_node, n_slots, best, keys, start_scan, P, 1, d_small + 1, nflag, ncnt);
            ++launches;
            RTDS_TRY(scan.run(nflag, nscan_a, n_slots, &launches));
            int n_interior = 0;
            RTDS_TRY(read_total(&n_interior));
            RTDS_TRY(scan.run(ncnt, nscan_b, n_slots, &launches));
            RTDS_TRY(read_total(&next_entries));
            next_n = 2 * n_interior;
            if ((size_t)(total_nodes + next_n) > cap_nodes || (size_t)next_entries > cap_entries) {
                rtds_set_error("kd build: capacity exceeded (%d nodes, %d list entries)", total_nodes + next_n, next_entries);
                return RTDS_ERR_CAPACITY;
            }
            if (n_interior > 0) {
                kd_make_children<<<G(n_slots), T, 0, s>>>(nodes, slot_to_node, n_slots, nscan_a, nscan_b, lvl_begin + lvl_n);
                kd_child_flags<<<G(n_edges), T, 0, s>>>(nodes, slot_to_node, keys, n_edges, lflags, rflags);
                launches += 2;
                RTDS_TRY(scan.run(lflags, start_scan, n_edges, &launches));
                RTDS_TRY(scan.run(rflags, escan2, n_edges, &launches));
                kd_scatter_children<<<G(n_edges), T, 0, s>>>(nodes, slot_to_node, keys, vals, n_edges, start_scan, escan2, lflags, rflags,
                                                              idx[cur ^ 1], owner[cur ^ 1]);
                ++launches;
            }
        }
        // leaves of this level -> pool
        kd_leaf_counts<<<G(lvl_n), T, 0, s>>>(nodes, lvl_begin, lvl_n, nleaf);
        RTDS_TRY(scan.run(nleaf, nscan_a, lvl_n, &launches));
        int leaf_entries = 0;
        RTDS_TRY(read_total(&leaf_entries));
        if ((size_t)(pool_used + leaf_entries) > cap_entries) { rtds_set_error("kd build: leaf pool capacity exceeded"); return RTDS_ERR_CAPACITY; }
        kd_leaf_offsets<<<G(lvl_n), T, 0, s>>>(nodes, lvl_begin, lvl_n, nscan_a, pool_used);
        kd_leaf_copy<<<G(n_entries), T, 0, s>>>(nodes, idx[cur], owner[cur], n_entries, pool);
        launches += 3;
        pool_used += leaf_entries;
        lvl_begin += lvl_n;
        lvl_n = next_n;
        total_nodes += next_n;
        n_entries = next_entries;
        cur ^= 1;
    }
    RTDS_CUDA(cudaGetLastError());

    // depth-first relabel
    for (int l = (int)levels.size() - 1; l >= 0; --l) kd_sizes<<<G(levels[l].second), T, 0, s>>>(nodes, levels[l].first, levels[l].second);
    for (size_t l = 0; l < levels.size(); ++l) kd_dfs<<<G(levels[l].second), T, 0, s>>>(nodes, levels[l].first, levels[l].second);
    launches += 2 * (int)levels.size();
    int* by_dfs = eflag;
    int* off_by_dfs = start_scan;
    kd_leaf_sizes_by_dfs<<<G(total_nodes), T, 0, s>>>(nodes, total_nodes, by_dfs);
    RTDS_TRY(scan.run(by_dfs, off_by_dfs, total_nodes, &launches));
    int n_idx = 0;
    RTDS_TRY(read_total(&n_idx));
    DeviceKd& kd = ctx->kd;
    RTDS_CUDA(cudaMalloc(&kd.nodes, sizeof(rtds_kd_node) * (size_t)total_nodes));
    RTDS_CUDA(cudaMalloc(&kd.prim_idx, sizeof(int) * (size_t)std::max(n_idx, 1)));
    RTDS_CUDA(cudaMemsetAsync(d_small + 2, 0, sizeof(int), s));
    kd_emit<<<G(total_nodes), T, 0, s>>>(nodes, total_nodes, off_by_dfs, pool, kd.nodes, kd.prim_idx, d_small + 2);
    launches += 2;
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    RTDS_CUDA(cudaGetLastError());
    int counted = 0;
    RTDS_CUDA(cudaMemcpyAsync(&counted, d_small + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    kd.n_nodes = total_nodes;
    kd.total_nodes = counted;
    kd.n_idx = n_idx;
    for (int a = 0; a < 6; ++a) kd.bounds[a] = b12[6 + a];
    kd.valid = true;
    float ms = 0;
    RTDS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (st) {
        st->n_prims = n;
        st->total_nodes = counted;
        st->alloc_nodes = total_nodes;
        st->max_depth = max_depth;
        st->kernel_launches = launches;
        st->ms = ms;
    }
    return RTDS_OK;
}

// List entries per level are bounded by (maxDepth+1)*n in PBRT's own allocation (accelerators.h:976); in practice
// a few n for the reference's small spheres, more when primitives straddle many planes: grow and retry.
int rtds_build_kd(rtds_ctx* ctx, const rtds_build_params* bp, rtds_build_stats* st)
{
    int rc = RTDS_ERR_CAPACITY;
    for (size_t mult : {(size_t)12, (size_t)40, (size_t)128}) {
        rc = build_kd_attempt(ctx, bp, st, mult);
        if (rc != RTDS_ERR_CAPACITY) break;
    }
    return rc;
}
