// bvh_builder.cpp -- host SAH builder producing the device BVH layout (FlatBvh, host_scene.h).
//
// Takes the place of create_bvh_tree / create_bvh_node / try_split (bvh.rs:26-144) for the device path.  Node
// order and shape are NOT observable through render_scene (only nearest-hit results and box containment are),
// so this is a different tree on purpose:
//   * the same objective family -- full-sweep surface-area heuristic over the three axes on EPS-padded triangle
//     boxes (aabb.rs:53-65), half-area x*y+y*z+z*x like Aabb::area (aabb.rs:32-38) -- but with a traversal-cost
//     term and leaves that may be split below 4 triangles when that is cheaper for the GPU traversal;
//   * O(n log^2 n): one sort per axis per level, boxes cached (the reference recomputes them in the comparator);
//   * output is a pre-order array of child-pair nodes with float boxes rounded outward.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "host_scene.h"

namespace rtb {
namespace {

const double kEps = 0.00001;  // geometry.rs:49
const double kInf = std::numeric_limits<double>::infinity();

struct Box {
    double mn[3], mx[3];
    void reset() { for (int a = 0; a < 3; ++a) { mn[a] = kInf; mx[a] = -kInf; } }
    void grow(const Box& o) { for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], o.mn[a]); mx[a] = std::max(mx[a], o.mx[a]); } }
    double half_area() const { double x = mx[0] - mn[0], y = mx[1] - mn[1], z = mx[2] - mn[2]; return x * y + y * z + z * x; }
};

struct TempNode { Box box; int left = -1, right = -1, first = 0, count = 0; };

Box from_d(const BoxD& d) { Box b; for (int a = 0; a < 3; ++a) { b.mn[a] = d.mn[a]; b.mx[a] = d.mx[a]; } return b; }

float round_down(double x) { float f = (float)x; if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity()); return f; }
float round_up(double x) { float f = (float)x; if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity()); return f; }

struct Build {
    const BvhBuildParams& p;
    std::vector<Box> boxes;        // per original triangle id (only those in `order` are valid)
    std::vector<int32_t> order;    // being permuted
    std::vector<TempNode> nodes;
    std::vector<double> right_area;
    std::vector<double> weight;    // per original id: what one more box of this primitive costs (1 for a triangle, the triangle count of a subtree); empty = all 1

    int build(int first, int count, int depth, int* max_depth) {
        TempNode n;
        n.box.reset();
        for (int i = first; i < first + count; ++i) n.box.grow(boxes[(size_t)order[(size_t)i]]);
        n.first = first; n.count = count;
        *max_depth = std::max(*max_depth, depth);
        int self = (int)nodes.size();
        nodes.push_back(n);
        if (count <= 1) return self;

        double best_cost = kInf; int best_axis = -1, best_split = -1;
        right_area.resize((size_t)count);
        if (p.force_median) {                              // balanced fallback: median of the centres along the longest axis, leaves of <= max_leaf
            if (count <= p.max_leaf_size && count <= 8) return self;
            int ax = 0;
            for (int a = 1; a < 3; ++a) if (n.box.mx[a] - n.box.mn[a] > n.box.mx[ax] - n.box.mn[ax]) ax = a;
            sort_axis(first, count, ax);
            int l = build(first, count / 2, depth + 1, max_depth);
            int r = build(first + count / 2, count - count / 2, depth + 1, max_depth);
            nodes[(size_t)self].left = l; nodes[(size_t)self].right = r;
            return self;
        }
        for (int axis = 0; axis < (p.size_split ? 4 : 3); ++axis) {             // axis 3: primitives ordered by box area, largest first ("the walls vs the mesh")
            sort_axis(first, count, axis);
            Box acc; acc.reset();
            for (int i = count - 1; i > 0; --i) { acc.grow(boxes[(size_t)order[(size_t)(first + i)]]); right_area[(size_t)i] = acc.half_area(); }
            acc.reset();
            double total_w = (double)count, left_w = 0.0;
            if (!weight.empty()) { total_w = 0.0; for (int i = 0; i < count; ++i) total_w += weight[(size_t)order[(size_t)(first + i)]]; }
            for (int i = 0; i + 1 < count; ++i) {
                acc.grow(boxes[(size_t)order[(size_t)(first + i)]]);
                left_w += weight.empty() ? 1.0 : weight[(size_t)order[(size_t)(first + i)]];
                double cost = left_w * acc.half_area() + (total_w - left_w) * right_area[(size_t)(i + 1)];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i + 1; }
            }
        }
        double parent_area = n.box.half_area();
        double split_cost = p.traversal_cost * parent_area + best_cost;
        double leaf_cost = (double)count * parent_area;
        bool can_leaf = count <= p.max_leaf_size && count <= 8;
        if (can_leaf && leaf_cost <= split_cost) return self;
        if (best_axis < 0 || !(best_cost < kInf)) {  // degenerate boxes (NaN/inf input): median split on x
            best_axis = 0; best_split = count / 2;
        }
        sort_axis(first, count, best_axis);
        int l = build(first, best_split, depth + 1, max_depth);
        int r = build(first + best_split, count - best_split, depth + 1, max_depth);
        nodes[(size_t)self].left = l; nodes[(size_t)self].right = r;
        return self;
    }

    // Bottom-up alternative for SMALL sets (O(n^3), n <= 512): greedy agglomerative clustering -- merge the two clusters whose union has the
    // smallest surface area (Walter et al. 2008) -- then the optimal leaf / inner decision per node by dynamic programming over the
    // same cost model as the sweep (leaf: count * area; inner: traversal_cost * area + children).  Fills `nodes` (pre-order) and `order`.
    void build_agglomerative(int* max_depth) {
        struct A { Box box; int left, right, count, prim; double cost; bool leaf; };
        const int n = (int)order.size();
        std::vector<A> t;
        std::vector<int> active;
        for (int i = 0; i < n; ++i) { A a; a.box = boxes[(size_t)order[(size_t)i]]; a.left = a.right = -1; a.count = 1; a.prim = order[(size_t)i]; a.cost = a.box.half_area(); a.leaf = true; t.push_back(a); active.push_back(i); }
        while (active.size() > 1) {
            double best = kInf; size_t bi = 0, bj = 1;
            for (size_t i = 0; i < active.size(); ++i)
                for (size_t j = i + 1; j < active.size(); ++j) {
                    Box u = t[(size_t)active[i]].box; u.grow(t[(size_t)active[j]].box);
                    const double a = u.half_area();                       // (the increase of area, or area x count, measured worse: DESIGN.md section 13)
                    if (a < best) { best = a; bi = i; bj = j; }
                }
            A m; m.box = t[(size_t)active[bi]].box; m.box.grow(t[(size_t)active[bj]].box);
            m.left = active[bi]; m.right = active[bj]; m.count = t[(size_t)m.left].count + t[(size_t)m.right].count; m.prim = -1;
            const double area = m.box.half_area();
            const double inner = p.traversal_cost * area + t[(size_t)m.left].cost + t[(size_t)m.right].cost;
            const double as_leaf = (double)m.count * area;
            m.leaf = m.count <= p.max_leaf_size && m.count <= 8 && as_leaf <= inner;
            m.cost = m.leaf ? as_leaf : inner;
            t.push_back(m);
            active[bi] = (int)t.size() - 1;
            active.erase(active.begin() + (long)bj);
        }
        int root = active[0];
        // Re-insertion (Bittner, Hapala, Havran 2013): take a subtree out, close the gap with its sibling, and put it back where the sum
        // of the inner-node areas grows least (the old place is among the candidates, so a step never makes the tree worse); repeated
        // over all nodes, largest area first, until a pass gains < 0.1 %.  Brute-force candidate search: these sets are small.
        if (p.reinsertion && n >= 4) {
            std::vector<int> parent(t.size(), -1);
            for (size_t i = 0; i < t.size(); ++i) if (t[i].prim < 0) { parent[(size_t)t[i].left] = (int)i; parent[(size_t)t[i].right] = (int)i; }
            auto refit_up = [&](int x) {
                for (; x >= 0; x = parent[(size_t)x]) {
                    A& a = t[(size_t)x];
                    a.box = t[(size_t)a.left].box; a.box.grow(t[(size_t)a.right].box);
                    a.count = t[(size_t)a.left].count + t[(size_t)a.right].count;
                }
            };
            auto total_area = [&]() { double c = 0.0; for (size_t i = 0; i < t.size(); ++i) if (t[i].prim < 0) c += t[i].box.half_area(); return c; };
            double before = total_area();
            for (int pass = 0; pass < 24; ++pass) {
                std::vector<int> cand;
                for (size_t i = 0; i < t.size(); ++i) if ((int)i != root && parent[i] >= 0) cand.push_back((int)i);
                std::sort(cand.begin(), cand.end(), [&](int x, int y) { const double ax = t[(size_t)x].box.half_area(), ay = t[(size_t)y].box.half_area(); return ax != ay ? ax > ay : x < y; });
                for (int sidx : cand) {
                    const int pnode = parent[(size_t)sidx];
                    if (pnode < 0) continue;
                    const int sib = t[(size_t)pnode].left == sidx ? t[(size_t)pnode].right : t[(size_t)pnode].left;
                    const int g = parent[(size_t)pnode];
                    // detach: sib takes the place of pnode
                    parent[(size_t)sib] = g;
                    if (g < 0) root = sib; else { if (t[(size_t)g].left == pnode) t[(size_t)g].left = sib; else t[(size_t)g].right = sib; refit_up(g); }
                    parent[(size_t)sidx] = -1; parent[(size_t)pnode] = -1;
                    // best position: every node x of the remaining tree
                    const Box sb = t[(size_t)sidx].box;
                    double best = kInf; int bx = -1;
                    std::vector<std::pair<int, double>> stck(1, std::make_pair(root, 0.0));      // (node, area increase induced in its ancestors)
                    while (!stck.empty()) {
                        const std::pair<int, double> it = stck.back(); stck.pop_back();
                        const A& x = t[(size_t)it.first];
                        Box u = x.box; u.grow(sb);
                        const double direct = u.half_area();
                        if (it.second + direct < best) { best = it.second + direct; bx = it.first; }
                        if (x.prim < 0) {
                            const double induced = it.second + direct - x.box.half_area();
                            if (induced < best) { stck.push_back(std::make_pair(x.left, induced)); stck.push_back(std::make_pair(x.right, induced)); }   // branch and bound
                        }
                    }
                    if (bx < 0) bx = root;                                                         // (cannot happen with finite boxes)
                    // insert: pnode becomes the parent of (bx, sidx) in bx's place
                    const int bp = parent[(size_t)bx];
                    t[(size_t)pnode].left = bx; t[(size_t)pnode].right = sidx;
                    parent[(size_t)bx] = pnode; parent[(size_t)sidx] = pnode; parent[(size_t)pnode] = bp;
                    if (bp < 0) root = pnode; else { if (t[(size_t)bp].left == bx) t[(size_t)bp].left = pnode; else t[(size_t)bp].right = pnode; }
                    refit_up(pnode);
                }
                const double after = total_area();
                if (!(after < before * 0.999)) break;
                before = after;
            }
            // leaf / inner decision again, bottom-up
            std::vector<std::pair<int, int>> po(1, std::make_pair(root, 0));
            while (!po.empty()) {
                const int x = po.back().first;
                A& a = t[(size_t)x];
                if (a.prim >= 0) { po.pop_back(); continue; }
                if (po.back().second == 0) { po.back().second = 1; const int l = a.left, r = a.right; po.push_back(std::make_pair(l, 0)); po.push_back(std::make_pair(r, 0)); continue; }
                const double area = a.box.half_area();
                const double inner = p.traversal_cost * area + t[(size_t)a.left].cost + t[(size_t)a.right].cost, as_leaf = (double)a.count * area;
                a.leaf = a.count <= p.max_leaf_size && a.count <= 8 && as_leaf <= inner;
                a.cost = a.leaf ? as_leaf : inner;
                po.pop_back();
            }
        }
        // pre-order emission into TempNode / order
        nodes.clear();
        std::vector<int32_t> new_order;
        struct It { int a, parent, side, depth; };
        std::vector<It> st(1, It{root, -1, 0, 1});
        while (!st.empty()) {
            const It it = st.back(); st.pop_back();
            const A& a = t[(size_t)it.a];
            TempNode nd; nd.box = a.box; nd.count = a.count; nd.first = (int)new_order.size();
            *max_depth = std::max(*max_depth, it.depth);
            const int self = (int)nodes.size();
            nodes.push_back(nd);
            if (it.parent >= 0) { if (it.side == 0) nodes[(size_t)it.parent].left = self; else nodes[(size_t)it.parent].right = self; }
            if (a.leaf) {                                  // gather the primitives below in depth-first order
                std::vector<int> g(1, it.a);
                while (!g.empty()) { const int x = g.back(); g.pop_back(); if (t[(size_t)x].prim >= 0) new_order.push_back(t[(size_t)x].prim); else { g.push_back(t[(size_t)x].right); g.push_back(t[(size_t)x].left); } }
            } else {
                // `first` of an inner node = where its left-most primitive will land: the left child is emitted next (stack order), so the running size is right
                st.push_back(It{a.right, self, 1, it.depth + 1});
                st.push_back(It{a.left, self, 0, it.depth + 1});
            }
        }
        order = new_order;
    }

    void sort_axis(int first, int count, int axis) {
        if (axis == 3) {
            std::sort(order.begin() + first, order.begin() + first + count, [this](int32_t x, int32_t y) {
                const double ax = boxes[(size_t)x].half_area(), ay = boxes[(size_t)y].half_area();
                if (ax != ay) return ax > ay;
                return x < y;
            });
            return;
        }
        std::sort(order.begin() + first, order.begin() + first + count, [this, axis](int32_t x, int32_t y) {
            double cx = boxes[(size_t)x].mn[axis] + boxes[(size_t)x].mx[axis], cy = boxes[(size_t)y].mn[axis] + boxes[(size_t)y].mx[axis];
            if (cx != cy) return cx < cy;
            return x < y;  // deterministic order for equal centres
        });
    }
};

void put_box(FlatBvh* out, int node, int slot, const Box& b) {
    float* A = &out->box_a[(size_t)node * 4]; float* B = &out->box_b[(size_t)node * 4]; float* C = &out->box_c[(size_t)node * 4];
    float* xy = slot == 0 ? A : B;
    xy[0] = round_down(b.mn[0]); xy[1] = round_up(b.mx[0]); xy[2] = round_down(b.mn[1]); xy[3] = round_up(b.mx[1]);
    C[slot * 2 + 0] = round_down(b.mn[2]); C[slot * 2 + 1] = round_up(b.mx[2]);
}

int32_t leaf_ref(int first, int count) { return ~(int32_t)(((uint32_t)first << 3) | (uint32_t)(count - 1)); }

int emit(const Build& b, int t, FlatBvh* out) {
    int self = out->n_nodes++;
    out->box_a.resize((size_t)out->n_nodes * 4); out->box_b.resize((size_t)out->n_nodes * 4); out->box_c.resize((size_t)out->n_nodes * 4);
    out->child.resize((size_t)out->n_nodes * 2);
    int kids[2] = {b.nodes[(size_t)t].left, b.nodes[(size_t)t].right};
    for (int s = 0; s < 2; ++s) {
        const TempNode& c = b.nodes[(size_t)kids[s]];
        put_box(out, self, s, c.box);
        int32_t ref;
        if (c.left < 0) { ref = leaf_ref(c.first, c.count); out->n_leaves++; out->max_leaf = std::max(out->max_leaf, c.count); }
        else ref = emit(b, kids[s], out);
        out->child[(size_t)self * 2 + (size_t)s] = ref;
    }
    return self;
}

}  // namespace

BoxD tri_box_d(const double* v) {  // aabb.rs:53-65: min/max of the vertices -/+ EPS
    BoxD b;
    for (int a = 0; a < 3; ++a) {
        b.mn[a] = std::min(std::min(v[a], v[3 + a]), v[6 + a]) - kEps;
        b.mx[a] = std::max(std::max(v[a], v[3 + a]), v[6 + a]) + kEps;
    }
    return b;
}

void quat_to_matrix(const double* q, double* m) {
    const double i = q[0], j = q[1], k = q[2], w = q[3];
    m[0] = 1 - 2 * (j * j + k * k); m[1] = 2 * (i * j - k * w);     m[2] = 2 * (i * k + j * w);
    m[3] = 2 * (i * j + k * w);     m[4] = 1 - 2 * (i * i + k * k); m[5] = 2 * (j * k - i * w);
    m[6] = 2 * (i * k - j * w);     m[7] = 2 * (j * k + i * w);     m[8] = 1 - 2 * (i * i + j * j);
}

void build_bvh(const std::vector<BoxD>& boxes, const std::vector<int32_t>& ids, const BvhBuildParams& p, FlatBvh* out, const std::vector<double>* weights) {
    *out = FlatBvh();
    Build b{p, {}, ids, {}, {}, {}};
    if (weights) b.weight = *weights;
    int32_t max_id = -1;
    for (int32_t id : ids) max_id = std::max(max_id, id);
    b.boxes.resize((size_t)(max_id + 1));
    for (int32_t id : ids) {
        Box bx = from_d(boxes[(size_t)id]);
        bool finite = true;                                   // a primitive with NaN / inf / astronomically large coordinates cannot be hit; it must not poison
        for (int a = 0; a < 3; ++a) finite = finite && std::isfinite(bx.mn[a]) && std::isfinite(bx.mx[a]) && std::fabs(bx.mn[a]) < 1e30 && std::fabs(bx.mx[a]) < 1e30;   // the comparisons of the builders
        if (!finite) for (int a = 0; a < 3; ++a) { bx.mn[a] = 0.0; bx.mx[a] = 0.0; }
        b.boxes[(size_t)id] = bx;
    }
    int max_depth = 0;
    if (p.agglomerative && !p.force_median && ids.size() >= 2 && ids.size() <= 512) b.build_agglomerative(&max_depth);
    else if (!ids.empty()) b.build(0, (int)ids.size(), 1, &max_depth);
    out->tri_order = b.order;
    out->depth = max_depth;
    if (!ids.empty() && b.nodes[0].left >= 0) {
        emit(b, 0, out);
        return;
    }
    // Root is a leaf (or the set is empty): wrap it in one pair node whose second child is a far-away point box
    // that no practical ray reaches (its leaf re-tests triangle 0, which cannot change the nearest hit).
    out->n_nodes = 1;
    out->box_a.assign(4, 0.f); out->box_b.assign(4, 0.f); out->box_c.assign(4, 0.f); out->child.assign(2, 0);
    Box far; for (int a = 0; a < 3; ++a) { far.mn[a] = 1e30; far.mx[a] = 1e30; }
    if (ids.empty()) {
        put_box(out, 0, 0, far); put_box(out, 0, 1, far);
        out->child[0] = leaf_ref(0, 1); out->child[1] = leaf_ref(0, 1);   // caller supplies one degenerate triangle
        out->n_leaves = 0; out->max_leaf = 0; out->depth = 1;
    } else {
        put_box(out, 0, 0, b.nodes[0].box); put_box(out, 0, 1, far);
        out->child[0] = leaf_ref(0, (int)ids.size()); out->child[1] = leaf_ref(0, 1);
        out->n_leaves = 1; out->max_leaf = (int)ids.size(); out->depth = 2;
    }
}

// GPU LBVH -> better top of the tree.  The Morton-code hierarchy is good locally and poor globally (it splits space, not surface
// area), and the top levels are the ones every ray walks.  Cut the tree into ~n_clusters subtrees (largest surface area first),
// rebuild the part above the cut with the host's full-sweep SAH over the subtree boxes (weighted by their triangle counts) and
// graft the untouched subtrees below it.  Milliseconds for a few thousand clusters; leaves and tri_order do not change.
void regraft_top_sah(FlatBvh* bvh, int n_clusters, const BvhBuildParams& p) {
    const int N = bvh->n_nodes;
    n_clusters = std::min(n_clusters, N / 3);                 // the cut needs room below it
    if (n_clusters < 8) return;
    auto slot_box = [bvh](int node, int s) {
        const float* A = &bvh->box_a[(size_t)node * 4]; const float* B = &bvh->box_b[(size_t)node * 4]; const float* C = &bvh->box_c[(size_t)node * 4];
        const float* xy = s == 0 ? A : B;
        BoxD b;
        b.mn[0] = xy[0]; b.mx[0] = xy[1]; b.mn[1] = xy[2]; b.mx[1] = xy[3]; b.mn[2] = C[s * 2]; b.mx[2] = C[s * 2 + 1];
        return b;
    };
    // triangles below every node (iterative post-order; the GPU builder's node order is not specified)
    std::vector<int32_t> tris((size_t)N, -1);
    {
        std::vector<int32_t> st(1, 0);
        while (!st.empty()) {
            const int32_t n = st.back();
            int32_t sum = 0; bool ready = true;
            for (int c = 0; c < 2; ++c) {
                const int32_t r = bvh->child[(size_t)n * 2 + (size_t)c];
                if (r < 0) sum += (int32_t)((uint32_t)(~r) & 7u) + 1;
                else if (tris[(size_t)r] < 0) { st.push_back(r); ready = false; }
                else sum += tris[(size_t)r];
            }
            if (ready) { tris[(size_t)n] = sum; st.pop_back(); }
        }
    }
    struct Elem { BoxD box; int32_t ref; double area; };
    auto elem = [&](int node, int s) {
        Elem e; e.box = slot_box(node, s); e.ref = bvh->child[(size_t)node * 2 + (size_t)s];
        const double x = e.box.mx[0] - e.box.mn[0], y = e.box.mx[1] - e.box.mn[1], z = e.box.mx[2] - e.box.mn[2];
        e.area = x * y + y * z + z * x;
        if (!(e.area >= 0.0)) e.area = 0.0;                   // NaN boxes (non-finite input) must not break the heap order
        return e;
    };
    std::vector<Elem> done;                                   // leaves met on the way: part of the frontier as they are
    auto cmp = [](const Elem& a, const Elem& b) { return a.area < b.area; };
    std::vector<Elem> heap;
    std::vector<char> dropped((size_t)N, 0);
    dropped[0] = 1;
    for (int s = 0; s < 2; ++s) { Elem e = elem(0, s); if (e.ref >= 0) { heap.push_back(e); std::push_heap(heap.begin(), heap.end(), cmp); } else done.push_back(e); }
    while (!heap.empty() && (int)(heap.size() + done.size()) < n_clusters) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        const Elem top = heap.back(); heap.pop_back();
        dropped[(size_t)top.ref] = 1;
        for (int s = 0; s < 2; ++s) { Elem e = elem(top.ref, s); if (e.ref >= 0) { heap.push_back(e); std::push_heap(heap.begin(), heap.end(), cmp); } else done.push_back(e); }
    }
    std::vector<Elem> frontier = done;
    frontier.insert(frontier.end(), heap.begin(), heap.end());
    const int K = (int)frontier.size();
    if (K < 8) return;
    std::vector<BoxD> boxes((size_t)K); std::vector<int32_t> ids((size_t)K); std::vector<double> w((size_t)K);
    for (int k = 0; k < K; ++k) {
        boxes[(size_t)k] = frontier[(size_t)k].box; ids[(size_t)k] = k;
        const int32_t r = frontier[(size_t)k].ref;
        w[(size_t)k] = r >= 0 ? (double)tris[(size_t)r] : (double)(((uint32_t)(~r) & 7u) + 1);
    }
    BvhBuildParams tp = p; tp.max_leaf_size = 1;              // a leaf of the top tree is exactly one subtree
    FlatBvh T;
    build_bvh(boxes, ids, tp, &T, &w);
    // new numbering: the top tree first, then the surviving nodes in their old order
    std::vector<int32_t> map((size_t)N, -1);
    int32_t next = T.n_nodes;
    for (int n = 0; n < N; ++n) if (!dropped[(size_t)n]) map[(size_t)n] = next++;
    FlatBvh out;
    out.n_nodes = next; out.n_leaves = bvh->n_leaves; out.max_leaf = bvh->max_leaf; out.tri_order = bvh->tri_order;
    out.box_a.resize((size_t)next * 4); out.box_b.resize((size_t)next * 4); out.box_c.resize((size_t)next * 4); out.child.resize((size_t)next * 2);
    for (int n = 0; n < T.n_nodes; ++n) {
        for (int k = 0; k < 4; ++k) { out.box_a[(size_t)n * 4 + k] = T.box_a[(size_t)n * 4 + k]; out.box_b[(size_t)n * 4 + k] = T.box_b[(size_t)n * 4 + k]; out.box_c[(size_t)n * 4 + k] = T.box_c[(size_t)n * 4 + k]; }
        for (int c = 0; c < 2; ++c) {
            int32_t r = T.child[(size_t)n * 2 + (size_t)c];
            if (r < 0) {                                                        // top leaf = frontier element
                const int32_t fr = frontier[(size_t)T.tri_order[(size_t)((uint32_t)(~r) >> 3)]].ref;
                r = fr >= 0 ? map[(size_t)fr] : fr;
            }
            out.child[(size_t)n * 2 + (size_t)c] = r;
        }
    }
    for (int n = 0; n < N; ++n) {
        if (dropped[(size_t)n]) continue;
        const size_t m = (size_t)map[(size_t)n];
        for (int k = 0; k < 4; ++k) { out.box_a[m * 4 + k] = bvh->box_a[(size_t)n * 4 + k]; out.box_b[m * 4 + k] = bvh->box_b[(size_t)n * 4 + k]; out.box_c[m * 4 + k] = bvh->box_c[(size_t)n * 4 + k]; }
        for (int c = 0; c < 2; ++c) { const int32_t r = bvh->child[(size_t)n * 2 + (size_t)c]; out.child[m * 2 + (size_t)c] = r >= 0 ? map[(size_t)r] : r; }
    }
    // depth of the grafted tree (a leaf below a node at depth d sits at d + 1, like the host builder counts)
    int depth = 0;
    std::vector<std::pair<int32_t, int>> st(1, std::make_pair(0, 1));
    while (!st.empty()) {
        const std::pair<int32_t, int> it = st.back(); st.pop_back();
        for (int c = 0; c < 2; ++c) {
            const int32_t r = out.child[(size_t)it.first * 2 + (size_t)c];
            if (r >= 0) st.push_back(std::make_pair(r, it.second + 1)); else depth = std::max(depth, it.second + 1);
        }
    }
    out.depth = depth;
    *bvh = out;
}

int validate_flat_bvh(const FlatBvh& bvh, const std::vector<BoxD>& boxes) {
    int bad = 0;
    auto slot_box = [&bvh](int node, int s, float mn[3], float mx[3]) {
        const float* A = &bvh.box_a[(size_t)node * 4]; const float* B = &bvh.box_b[(size_t)node * 4]; const float* C = &bvh.box_c[(size_t)node * 4];
        const float* xy = s == 0 ? A : B;
        mn[0] = xy[0]; mx[0] = xy[1]; mn[1] = xy[2]; mx[1] = xy[3]; mn[2] = C[s * 2]; mx[2] = C[s * 2 + 1];
    };
    for (int n = 0; n < bvh.n_nodes; ++n) {
        for (int s = 0; s < 2; ++s) {
            float mn[3], mx[3];
            slot_box(n, s, mn, mx);
            int32_t ref = bvh.child[(size_t)n * 2 + (size_t)s];
            if (mn[0] > 1e29f) continue;  // the far-away filler box
            if (ref < 0) {
                uint32_t code = (uint32_t)~ref; int first = (int)(code >> 3), count = (int)(code & 7u) + 1;
                if (bvh.tri_order.empty()) continue;
                for (int i = first; i < first + count; ++i) {
                    Box tb = from_d(boxes[(size_t)bvh.tri_order[(size_t)i]]);
                    for (int a = 0; a < 3; ++a) if (!((double)mn[a] <= tb.mn[a] && (double)mx[a] >= tb.mx[a])) { ++bad; break; }
                }
            } else {
                for (int cs = 0; cs < 2; ++cs) {
                    float cmn[3], cmx[3];
                    slot_box(ref, cs, cmn, cmx);
                    if (cmn[0] > 1e29f) continue;
                    for (int a = 0; a < 3; ++a) if (!(mn[a] <= cmn[a] && mx[a] >= cmx[a])) { ++bad; break; }
                }
            }
        }
    }
    return bad;
}

}  // namespace rtb
