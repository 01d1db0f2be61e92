// Host-side proxy for the traversal work of a tree: builds the trees the library builds (bvh_builder.cpp, no CUDA), walks them with the device's
// rule (ordered by entry distance, pruned by the best hit, leaves tested with a plain f64 triangle test) for camera rays and cosine-distributed
// bounce rays from random surface points, and prints box tests / triangle tests per ray.  Development aid for builder changes when no GPU is at hand:
//   g++ -O2 -std=c++17 -I../../../raytracing-course-2024_b200/csrc -I../../../include bvh_proxy.cpp ../../../raytracing-course-2024_b200/csrc/gltf_loader.cpp \
//       ../../../raytracing-course-2024_b200/csrc/bvh_builder.cpp -o _build/bvh_proxy && _build/bvh_proxy ../../../scenes/practice7_2.gltf
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <array>
#include "host_scene.h"
using namespace rtb;
struct V { double x, y, z; };
static V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { double l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }
static bool tri_hit(const double* v, V o, V d, double& t) {
    V a{v[0], v[1], v[2]}, e1 = sub(V{v[3], v[4], v[5]}, a), e2 = sub(V{v[6], v[7], v[8]}, a);
    V p = cross(d, e2); double det = dot(e1, p); if (det == 0) return false;
    V s = sub(o, a); double u = dot(s, p) / det; if (u < 0 || u > 1) return false;
    V q = cross(s, e1); double w = dot(d, q) / det; if (w < 0 || u + w > 1) return false;
    t = dot(e2, q) / det; return t > 1e-7;
}
struct Counts { double box = 0, tri = 0, rays = 0; };
static void walk(const FlatBvh& b, const std::vector<double>& tv, V o, V d, Counts& c) {
    double inv[3] = {1 / (std::fabs(d.x) < 1e-20 ? 1e-20 : d.x), 1 / (std::fabs(d.y) < 1e-20 ? 1e-20 : d.y), 1 / (std::fabs(d.z) < 1e-20 ? 1e-20 : d.z)};
    double oo[3] = {o.x, o.y, o.z};
    double best = 1e300;
    std::vector<int32_t> st; int32_t cur = 0;
    c.rays += 1;
    for (;;) {
        if (cur >= 0) {
            const float* A = &b.box_a[(size_t)cur * 4]; const float* B = &b.box_b[(size_t)cur * 4]; const float* C = &b.box_c[(size_t)cur * 4];
            double lo[2][3] = {{A[0], A[2], C[0]}, {B[0], B[2], C[2]}}, hi[2][3] = {{A[1], A[3], C[1]}, {B[1], B[3], C[3]}};
            double tn[2], tf[2]; bool h[2];
            for (int s = 0; s < 2; ++s) {
                double t0 = 0, t1 = best;
                for (int a = 0; a < 3; ++a) { double x = (lo[s][a] - oo[a]) * inv[a], y = (hi[s][a] - oo[a]) * inv[a]; if (x > y) std::swap(x, y); t0 = std::max(t0, x); t1 = std::min(t1, y); }
                tn[s] = t0; tf[s] = t1; h[s] = t0 <= t1;
            }
            c.box += 2;
            const int32_t r0 = b.child[(size_t)cur * 2], r1 = b.child[(size_t)cur * 2 + 1];
            if (h[0] && h[1]) { if (tn[1] < tn[0]) { st.push_back(r0); cur = r1; } else { st.push_back(r1); cur = r0; } continue; }
            if (h[0]) { cur = r0; continue; }
            if (h[1]) { cur = r1; continue; }
        } else {
            const uint32_t code = (uint32_t)~cur; const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
            for (int i = first; i < first + n; ++i) { double t; c.tri += 1; if (tri_hit(&tv[(size_t)b.tri_order[(size_t)i] * 9], o, d, t) && t < best) best = t; }
        }
        if (st.empty()) break;
        cur = st.back(); st.pop_back();
    }
}
int main(int argc, char** argv) {
    HostScene h; LoadError err;
    if (argc < 2 || !load_gltf_scene(argv[1], 64, 64, 1, &h, &err)) { printf("usage: bvh_proxy scene.gltf [rays]\n"); return 1; }
    const int n = h.n_tris(), n_rays = argc > 2 ? atoi(argv[2]) : 20000;
    std::vector<BoxD> boxes((size_t)n); std::vector<int32_t> ids((size_t)n);
    for (int i = 0; i < n; ++i) { boxes[(size_t)i] = tri_box_d(&h.tri_v[(size_t)i * 9]); ids[(size_t)i] = i; }
    // rays: bounce rays from area-weighted random surface points, cosine-distributed about the face normal (either side)
    std::mt19937_64 rng(7); std::uniform_real_distribution<double> U(0, 1);
    std::vector<double> cdf((size_t)n); double acc = 0;
    for (int i = 0; i < n; ++i) { const double* v = &h.tri_v[(size_t)i * 9]; V a{v[0], v[1], v[2]}; V c = cross(sub(V{v[3], v[4], v[5]}, a), sub(V{v[6], v[7], v[8]}, a)); acc += 0.5 * std::sqrt(dot(c, c)); cdf[(size_t)i] = acc; }
    std::vector<V> ro, rd;
    for (int k = 0; k < n_rays; ++k) {
        const int i = (int)(std::lower_bound(cdf.begin(), cdf.end(), U(rng) * acc) - cdf.begin());
        const double* v = &h.tri_v[(size_t)std::min(i, n - 1) * 9];
        double u = U(rng), w = U(rng); if (u + w > 1) { u = 1 - u; w = 1 - w; }
        V a{v[0], v[1], v[2]}, e1 = sub(V{v[3], v[4], v[5]}, a), e2 = sub(V{v[6], v[7], v[8]}, a);
        V nn = norm(cross(e1, e2)); if (U(rng) < 0.5) nn = V{-nn.x, -nn.y, -nn.z};
        V p{a.x + u * e1.x + w * e2.x + 1e-5 * nn.x, a.y + u * e1.y + w * e2.y + 1e-5 * nn.y, a.z + u * e1.z + w * e2.z + 1e-5 * nn.z};
        const double z = 1 - 2 * U(rng), r = std::sqrt(std::max(0.0, 1 - z * z)), ph = 6.283185307179586 * U(rng);
        V dd = norm(V{r * std::cos(ph) + nn.x, r * std::sin(ph) + nn.y, z + nn.z});
        ro.push_back(p); rd.push_back(dd);
    }
    auto report = [&](const char* label, const FlatBvh& b) {
        Counts c; for (size_t k = 0; k < ro.size(); ++k) walk(b, h.tri_v, ro[k], rd[k], c);
        printf("%-46s nodes %6d depth %3d  box tests / ray %7.3f  triangle tests / ray %6.3f  validate %d\n", label, b.n_nodes, b.depth, c.box / c.rays, c.tri / c.rays, validate_flat_bvh(b, boxes));
        fflush(stdout);
    };
    BvhBuildParams p; p.max_leaf_size = 2; p.traversal_cost = 2.0;
    { BvhBuildParams q = p; q.agglomerative = false; FlatBvh b; build_bvh(boxes, ids, q, &b); report("top-down SAH sweep", b);
      if (n > 512) { for (int k : {8, 16, 32, 64, 96, 128, 192, 256, 512}) { FlatBvh t = b; regraft_top_sah(&t, k, p); char l[64]; snprintf(l, sizeof l, "sweep + bottom-up top %d", k); report(l, t); } } }
    if (n <= 512) { FlatBvh b; build_bvh(boxes, ids, p, &b); report("bottom-up + re-insertion", b); BvhBuildParams q = p; q.reinsertion = false; FlatBvh c; build_bvh(boxes, ids, q, &c); report("bottom-up", c); }
    if (argc > 3 && std::strcmp(argv[3], "split") == 0) {
        // Potential of SPATIAL SPLITS: every triangle whose box area exceeds `frac` of the scene box is cut into 4^L sub-triangles by midpoint
        // subdivision; each piece becomes a REFERENCE (its own box, the original triangle behind it).  The library's builders run over the references;
        // the walk tests the original triangle (a triangle reached through two leaves is tested twice, as a real implementation would).
        BoxD sb = boxes[0];
        for (int i = 1; i < n; ++i) for (int a = 0; a < 3; ++a) { sb.mn[a] = std::min(sb.mn[a], boxes[(size_t)i].mn[a]); sb.mx[a] = std::max(sb.mx[a], boxes[(size_t)i].mx[a]); }
        auto harea = [](const BoxD& b) { const double x = b.mx[0] - b.mn[0], y = b.mx[1] - b.mn[1], z = b.mx[2] - b.mn[2]; return x * y + y * z + z * x; };
        const double scene_area = harea(sb);
        for (double frac : {1.0, 0.05, 0.01, 0.002}) for (int L : {1, 2, 3}) {
            if (frac == 1.0 && L > 1) continue;
            std::vector<BoxD> rb; std::vector<int32_t> rtri; std::vector<double> rv;     // references: box, original triangle, vertices (for the f64 test)
            for (int i = 0; i < n; ++i) {
                std::vector<std::array<double, 9>> parts(1); for (int k = 0; k < 9; ++k) parts[0][(size_t)k] = h.tri_v[(size_t)i * 9 + (size_t)k];
                if (harea(boxes[(size_t)i]) > frac * scene_area)
                    for (int l = 0; l < L; ++l) {
                        std::vector<std::array<double, 9>> nx;
                        for (auto& t : parts) {
                            double m[9]; for (int k = 0; k < 3; ++k) { m[k] = 0.5 * (t[(size_t)k] + t[(size_t)(3 + k)]); m[3 + k] = 0.5 * (t[(size_t)(3 + k)] + t[(size_t)(6 + k)]); m[6 + k] = 0.5 * (t[(size_t)(6 + k)] + t[(size_t)k]); }
                            nx.push_back({t[0], t[1], t[2], m[0], m[1], m[2], m[6], m[7], m[8]}); nx.push_back({m[0], m[1], m[2], t[3], t[4], t[5], m[3], m[4], m[5]});
                            nx.push_back({m[6], m[7], m[8], m[3], m[4], m[5], t[6], t[7], t[8]}); nx.push_back({m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]});
                        }
                        parts.swap(nx);
                    }
                for (auto& t : parts) { rb.push_back(tri_box_d(t.data())); rtri.push_back(i); }
            }
            std::vector<int32_t> rid(rb.size()); for (size_t k = 0; k < rid.size(); ++k) rid[k] = (int32_t)k;
            FlatBvh b; build_bvh(rb, rid, p, &b);
            if ((int)rb.size() > 512) regraft_top_sah(&b, 128, p);
            for (auto& x : b.tri_order) x = rtri[(size_t)x];                             // leaves now name original triangles
            Counts c; for (size_t k = 0; k < ro.size(); ++k) walk(b, h.tri_v, ro[k], rd[k], c);
            printf("split triangles above %5.3f of the scene box area, %d level(s): %6zu references (%d triangles), nodes %6d depth %3d  box tests / ray %7.3f  triangle tests / ray %6.3f\n",
                   frac, L, rb.size(), n, b.n_nodes, b.depth, c.box / c.rays, c.tri / c.rays);
            fflush(stdout);
        }
        return 0;
    }
    if (argc > 3 && std::strcmp(argv[3], "search") == 0 && n <= 512) {
        // How far is the library's tree from a LOCAL optimum of the measured work itself?  Hill climbing over subtree swaps (two child slots that are
        // not on one root path exchange their contents, the boxes above are refitted), objective = box tests + 2.1 * triangle tests on a ray sample.
        FlatBvh b; build_bvh(boxes, ids, p, &b);
        const size_t n_eval = std::min<size_t>(ro.size(), 6000);
        auto cost = [&](const FlatBvh& t, double* bx, double* tr) { Counts c; for (size_t k = 0; k < n_eval; ++k) walk(t, h.tri_v, ro[k], rd[k], c); if (bx) *bx = c.box / c.rays; if (tr) *tr = c.tri / c.rays; return (c.box + 2.1 * c.tri) / c.rays; };
        auto get = [&](FlatBvh& t, int node, int s, float* bx6, int32_t* ref) {
            const float* xy = s == 0 ? &t.box_a[(size_t)node * 4] : &t.box_b[(size_t)node * 4];
            bx6[0] = xy[0]; bx6[1] = xy[1]; bx6[2] = xy[2]; bx6[3] = xy[3]; bx6[4] = t.box_c[(size_t)node * 4 + s * 2]; bx6[5] = t.box_c[(size_t)node * 4 + s * 2 + 1]; *ref = t.child[(size_t)node * 2 + s]; };
        auto put = [&](FlatBvh& t, int node, int s, const float* bx6, int32_t ref) {
            float* xy = s == 0 ? &t.box_a[(size_t)node * 4] : &t.box_b[(size_t)node * 4];
            xy[0] = bx6[0]; xy[1] = bx6[1]; xy[2] = bx6[2]; xy[3] = bx6[3]; t.box_c[(size_t)node * 4 + s * 2] = bx6[4]; t.box_c[(size_t)node * 4 + s * 2 + 1] = bx6[5]; t.child[(size_t)node * 2 + s] = ref; };
        auto parents = [&](const FlatBvh& t, std::vector<int>& par, std::vector<int>& side) {
            par.assign((size_t)t.n_nodes, -1); side.assign((size_t)t.n_nodes, 0);
            for (int i = 0; i < t.n_nodes; ++i) for (int s = 0; s < 2; ++s) { const int32_t r = t.child[(size_t)i * 2 + s]; if (r >= 0) { par[(size_t)r] = i; side[(size_t)r] = s; } } };
        auto refit = [&](FlatBvh& t, int node, const std::vector<int>& par, const std::vector<int>& side) {
            for (int x = node; par[(size_t)x] >= 0; x = par[(size_t)x]) {
                float a6[6], b6[6], u[6]; int32_t r;
                get(t, x, 0, a6, &r); get(t, x, 1, b6, &r);
                for (int k = 0; k < 6; k += 2) { u[k] = std::min(a6[k], b6[k]); u[k + 1] = std::max(a6[k + 1], b6[k + 1]); }
                put(t, par[(size_t)x], side[(size_t)x], u, x);
            } };
        double bx, tr; double best = cost(b, &bx, &tr);
        printf("search start: cost %.3f (box %.3f, tri %.3f)\n", best, bx, tr);
        std::mt19937_64 g2(11);
        const int iters = argc > 4 ? atoi(argv[4]) : 60000;
        int accepted = 0;
        for (int it = 0; it < iters; ++it) {
            std::vector<int> par, side; parents(b, par, side);
            const int i = (int)(g2() % (uint64_t)b.n_nodes), j = (int)(g2() % (uint64_t)b.n_nodes), si = (int)(g2() & 1), sj = (int)((g2() >> 1) & 1);
            if (i == j) continue;
            // reject if one slot's subtree contains the other node (ancestor relation)
            auto is_anc = [&](int anc_node, int anc_side, int x) { for (int y = x; par[(size_t)y] >= 0; y = par[(size_t)y]) if (par[(size_t)y] == anc_node && side[(size_t)y] == anc_side) return true; return false; };
            if (is_anc(i, si, j) || is_anc(j, sj, i)) continue;
            FlatBvh t = b;
            float bi[6], bj[6]; int32_t ri, rj;
            get(t, i, si, bi, &ri); get(t, j, sj, bj, &rj);
            put(t, i, si, bj, rj); put(t, j, sj, bi, ri);
            std::vector<int> par2, side2; parents(t, par2, side2);
            refit(t, i, par2, side2); refit(t, j, par2, side2);
            const double c2 = cost(t, nullptr, nullptr);
            if (c2 < best) { best = c2; b = t; ++accepted; }
        }
        // honest evaluation on the FULL ray set (the search saw the first n_eval rays only)
        Counts c; for (size_t k = n_eval; k < ro.size(); ++k) walk(b, h.tri_v, ro[k], rd[k], c);
        printf("search end: %d swaps accepted, cost on the training rays %.3f; held-out rays: box %.3f tri %.3f; validate %d\n", accepted, best, c.box / c.rays, c.tri / c.rays, validate_flat_bvh(b, boxes));
        FlatBvh b0; build_bvh(boxes, ids, p, &b0); Counts c0; for (size_t k = n_eval; k < ro.size(); ++k) walk(b0, h.tri_v, ro[k], rd[k], c0);
        printf("library tree on the same held-out rays: box %.3f tri %.3f\n", c0.box / c0.rays, c0.tri / c0.rays);
        return 0;
    }
    if (n > 512) {
        BvhBuildParams q = p; q.agglomerative = false; q.size_split = true; FlatBvh b; build_bvh(boxes, ids, q, &b); report("sweep with the size order as 4th axis", b);
        FlatBvh t = b; regraft_top_sah(&t, 128, p); report("  + bottom-up top 128", t);
    }
    return 0;
}
