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
    if (n > 512) {
        BvhBuildParams q = p; q.agglomerative = false; q.size_split = true; FlatBvh b; build_bvh(boxes, ids, q, &b); report("sweep with the size order as 4th axis", b);
        FlatBvh t = b; regraft_top_sah(&t, 128, p); report("  + bottom-up top 128", t);
    }
    return 0;
}
