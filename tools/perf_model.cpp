// tools/perf_model.cpp — PERFORMANCE MODEL, not an oracle of reference behaviour and not product code.
// CPU restatement of the ordered packet traversal (hull test, binary or 4-wide tree, optional quantised child boxes) that
// COUNTS visits / tests / bytes; it was used to decide which traversal variants were worth building.
// It reuses the oracle's traversal primitives by including the oracle's translation unit (a dev tool may; the product never does).
// Built on demand by tools/wide4_model.py / tools/packet_size_model.py:  g++ -O2 -ffp-contract=off -shared -fPIC tools/perf_model.cpp
#include "../oracle/oracle.cpp"

extern "C" {
// ----------------------------------------------------------------------------------------------
// MODEL (design evidence, not a parity oracle): the ordered packet traversal of csrc/traverse.cuh (four rays of one pixel,
// origin 0, one hull test per child and packet, children ordered by the packet's entry bound, pruning against
// max_j(tnear_j + margin)) run on the CPU over the binary tree and over its 4-wide collapse, to COUNT dependent node visits
// for DESIGN.md section 10.1. Reciprocals are IEEE here (rcp.approx on the device) - irrelevant for counts. The hits it
// returns are checked against closest_bvh by the test that drives it.
// ----------------------------------------------------------------------------------------------
struct PacketStats { long long packets, interior_visits, leaf_visits, box_tests, prim_tests, max_stack; };

namespace {
struct PacketState {
    const Scene* S;
    int NR = 4;                     // rays per packet (<= 16)
    float d[16][3], inv[16][3], ilo[3], ihi[3];
    float tnear[16]; int best[16], best_key[16];
    float margin, tlim_max;
    void update_tlim()
    {
        tlim_max = -INFINITY;
        for (int j = 0; j < NR; ++j) { float t = tnear[j] + margin; t = std::fabs(t) * 9.53674316e-7f + t; tlim_max = std::max(tlim_max, t); }
    }
    // hull test of one box: lower bound of the entry distances (returned), accept flag
    bool hull(const float* bmin, const float* bmax, float& tmin_lo) const
    {
        float lo = -INFINITY, hi = INFINITY;
        for (int a = 0; a < 3; ++a) {
            const float p0 = bmin[a] * ilo[a], p1 = bmin[a] * ihi[a], p2 = bmax[a] * ilo[a], p3 = bmax[a] * ihi[a];
            lo = std::max(lo, std::min(std::min(p0, p1), std::min(p2, p3)));     // every ray's near-plane distance is above this
            hi = std::min(hi, std::max(std::max(p0, p1), std::max(p2, p3)));     // every ray's far-plane distance is below this
        }
        hi = std::fabs(hi) * 9.53674316e-7f + hi;
        tmin_lo = lo;
        return lo <= std::min(hi, tlim_max) && hi >= -margin;
    }
    void leaf(int leafpos, PacketStats& st)
    {
        const int obj = S->prim_order[leafpos];
        const float* s = S->sph + 4 * (size_t)obj;
        const float bmin[3] = {s[0] - s[3], s[1] - s[3], s[2] - s[3]}, bmax[3] = {s[0] + s[3], s[1] + s[3], s[2] + s[3]};
        const float o[3] = {0, 0, 0};
        for (int j = 0; j < NR; ++j) {
            if (!slab(o, d[j], bmin, bmax)) continue;              // the reference's own test decides candidacy (leaf-local)
            ++st.prim_tests;
            float t0 = INFINITY, t1 = INFINITY;
            if (S->test(o, d[j], obj, t0, t1)) {
                if (t0 < 0) t0 = t1;
                const int key = S->tie_by_objid ? obj : leafpos;
                if (t0 < tnear[j] || (t0 == tnear[j] && best[j] >= 0 && key < best_key[j])) { tnear[j] = t0; best[j] = obj; best_key[j] = key; }
            }
        }
        update_tlim();
    }
};
}  // namespace

void orc_packet_model(const float* cxyz_r, int n, const LinearNode* nodes, int n_nodes, const Wide4Node* wide, int n_wide, const int* prim_order,
                      int tie_by_objid, const float* dirs /*packets x NR x 3*/, int n_packets, int use_wide, int* hit /*packets x NR*/,
                      PacketStats* out)
{
    const int order_mode = (use_wide >> 4) & 15;   // 0: accepted children fully sorted by entry bound; 1: nearest first, the rest in list order
    const int quant_bits = (use_wide >> 8) & 255;  // > 0 (wide tree only): child boxes quantised OUTWARD to this many bits per plane relative to
                                                   // the union of the node's child boxes - what a compressed node would hold (conservative)
    const int NR = ((use_wide >> 16) & 31) ? ((use_wide >> 16) & 31) : 4;      // rays per packet, default 4, <= 16
    use_wide &= 15;
    Scene S{cxyz_r, nullptr, n, nodes, prim_order, n_nodes, tie_by_objid};
    PacketStats st = {0, 0, 0, 0, 0, 0};
    float ex = 0, ey = 0, ez = 0;
    for (int a = 0; a < 3; ++a) {
        const float e = std::max(std::fabs(nodes[0].bmin[a]), std::fabs(nodes[0].bmax[a]));
        (a == 0 ? ex : a == 1 ? ey : ez) = e;
    }
    const float margin = std::sqrt(ex * ex + ey * ey + ez * ez) * 0.00278f;      // prune_margin, traverse.cuh
    struct Item { int ref; float t; };
    std::vector<Item> stack;
    for (int p = 0; p < n_packets; ++p) {
        PacketState P;
        P.S = &S; P.margin = margin; P.NR = NR;
        bool same_oct = true;
        for (int j = 0; j < NR; ++j)
            for (int a = 0; a < 3; ++a) {
                P.d[j][a] = dirs[((size_t)p * NR + j) * 3 + a];
                P.inv[j][a] = 1.0f / P.d[j][a];
                if ((P.d[j][a] < 0) != (P.d[0][a] < 0)) same_oct = false;
            }
        for (int j = 0; j < NR; ++j) { P.tnear[j] = INFINITY; P.best[j] = -1; P.best_key[j] = 0; }
        if (!same_oct) {      // the kernel traces such pixels ray by ray; leave them out of the counts
            for (int j = 0; j < NR; ++j) { float t; const float o[3] = {0, 0, 0}; closest_bvh(S, o, P.d[j], hit[NR * p + j], t, nullptr); }
            continue;
        }
        ++st.packets;
        for (int a = 0; a < 3; ++a) {
            P.ilo[a] = P.ihi[a] = P.inv[0][a];
            for (int j = 1; j < NR; ++j) { P.ilo[a] = std::min(P.ilo[a], P.inv[j][a]); P.ihi[a] = std::max(P.ihi[a], P.inv[j][a]); }
        }
        P.update_tlim();
        stack.clear();
        // refs: binary tree: node index (leaf when nPrimitives); wide tree: >= 0 wide node, < 0 ~leafpos
        int cur = 0;
        bool have = true;
        while (have) {
            const bool is_leaf = use_wide ? cur < 0 : nodes[cur].nPrimitives != 0;
            if (is_leaf) {
                ++st.leaf_visits;
                P.leaf(use_wide ? ~cur : nodes[cur].offset, st);
            } else {
                ++st.interior_visits;
                Item acc[4]; int m = 0;
                if (use_wide) {
                    const Wide4Node& nd = wide[cur];
                    float pmin[3], pext[3];
                    if (quant_bits) {
                        for (int a = 0; a < 3; ++a) {
                            float lo = INFINITY, hi = -INFINITY;
                            for (int k = 0; k < nd.n_children; ++k) { lo = std::min(lo, nd.bmin[k][a]); hi = std::max(hi, nd.bmax[k][a]); }
                            pmin[a] = lo; pext[a] = hi - lo;
                        }
                    }
                    for (int k = 0; k < nd.n_children; ++k) {
                        float t, qmin[3], qmax[3];
                        const float* bmn = nd.bmin[k]; const float* bmx = nd.bmax[k];
                        if (quant_bits) {
                            const double levels = (double)((1u << quant_bits) - 1);
                            for (int a = 0; a < 3; ++a) {
                                if (pext[a] > 0) {
                                    const double step = (double)pext[a] / levels;
                                    qmin[a] = (float)(pmin[a] + std::floor(((double)nd.bmin[k][a] - pmin[a]) / step) * step);
                                    qmax[a] = (float)(pmin[a] + std::ceil(((double)nd.bmax[k][a] - pmin[a]) / step) * step);
                                    qmin[a] = std::min(qmin[a], nd.bmin[k][a]); qmax[a] = std::max(qmax[a], nd.bmax[k][a]);   // float rounding: stay outside
                                } else { qmin[a] = nd.bmin[k][a]; qmax[a] = nd.bmax[k][a]; }
                            }
                            bmn = qmin; bmx = qmax;
                        }
                        ++st.box_tests;
                        if (P.hull(bmn, bmx, t)) acc[m++] = Item{nd.child[k], t};
                    }
                } else {
                    const int kids[2] = {cur + 1, nodes[cur].offset};
                    for (int k = 0; k < 2; ++k) { float t; ++st.box_tests; if (P.hull(nodes[kids[k]].bmin, nodes[kids[k]].bmax, t)) acc[m++] = Item{kids[k], t}; }
                }
                if (order_mode == 0) std::stable_sort(acc, acc + m, [](const Item& a, const Item& b) { return a.t < b.t; });
                else if (m > 1) { int kmin = 0; for (int k = 1; k < m; ++k) if (acc[k].t < acc[kmin].t) kmin = k; std::swap(acc[0], acc[kmin]); }
                for (int k = m - 1; k >= 1; --k) stack.push_back(acc[k]);          // nearest of the deferred ones on top
                st.max_stack = std::max<long long>(st.max_stack, (long long)stack.size());
                if (m) { cur = acc[0].ref; continue; }
            }
            have = false;
            while (!stack.empty()) {
                const Item e = stack.back(); stack.pop_back();
                if (e.t > P.tlim_max) continue;
                cur = e.ref; have = true;
                break;
            }
        }
        for (int j = 0; j < NR; ++j) hit[NR * p + j] = P.best[j];
    }
    if (out) *out = st;
}
}  // extern "C"
