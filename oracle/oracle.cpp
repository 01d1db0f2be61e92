// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under raytracing-course-2024_b200/
// may include, link or dlopen this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, and only as the checker / the CPU stand-in.
//
// A CPU f64 restatement of the reference's per-pixel path-tracing loop (metametamoon/raytracing-course-2024):
//   src/rendering.rs:21-262, src/geometry.rs:5-251, src/bvh.rs:26-322, src/aabb.rs:5-106,
//   src/distributions.rs:53-298, src/utils.rs:3-21.
// Each function cites the lines it follows.  The arithmetic keeps the reference's operation order (compile with
// -ffp-contract=off: rustc does not fuse multiply-adds).
//
// PARITY STATUS: UNPINNED.  The reference ships no golden vectors, known-answer tests or images for this path
//  and its Rust toolchain is absent here, so this restatement is pinned only by the
// hand-derived known answers of SURVEY.md A.3 (tests/test_oracle_kat.py) and by its own invariants.
//
// Two kinds of content, marked at every function:
//   [REF]       follows the cited reference lines (triangles, boxes, Object3D transforms, AABBs, BVH, samplers, BRDF);
//   [OWN SPEC]  has NO counterpart at reference HEAD: the text-scene shapes PLANE and ELLIPSOID, the DIELECTRIC
//               material and ellipsoid light sampling.  The reference's parser and those shapes were deleted before
//               HEAD (main.rs:48 keeps `// let scene = parse_file_content(file_lines);`, scene.rs:18,37 and
//               geometry.rs:23 keep the dead fields `ior`, `infinite_primitives`, `is_outer_to_inner`).  Their
//               semantics are declared in DESIGN.md section 12 and restated here so that the device has a checker.
//
// Third-party arithmetic restated from its published definition (crate versions are caret ranges, no lockfile):
//   nalgebra 0.32 : Vector3 dot/cross/normalize (v / |v|), Matrix3::try_inverse (cofactors / det, None iff
//                   det == 0), UnitQuaternion * Vector3 (t = 2 q.v x p;  p' = t w + q.v x t + p; the identity
//                   quaternion is an exact no-op), conjugate (negated vector part).
//   rand 0.8 / rand_xoshiro 0.6 : xoshiro256** seeded through SplitMix64, gen::<f64>() = (u64 >> 11) * 2^-53,
//                   gen_range(0.0..1.0) = [1,2)-mantissa trick, gen_range(0..n) = widening-multiply rejection.
//   rand_distr 0.4 Normal : the ziggurat tables are not available here; a Marsaglia polar normal is used
//                   instead (same distribution, different stream).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

typedef double Fp;                                   // geometry.rs:5
static const Fp EPS = 0.00001;                       // geometry.rs:49
static const Fp FP_PI = 3.14159265358979323846;      // geometry.rs:6
static const Fp FP_INF = std::numeric_limits<Fp>::infinity();

// ------------------------------------------------------------------------------------------------ vectors
struct V3 { Fp x, y, z; };
static inline V3 v3(Fp x, Fp y, Fp z) { V3 r = {x, y, z}; return r; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, Fp s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(Fp s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, Fp s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 cmul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 cdiv(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline Fp dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline Fp norm_squared(V3 a) { return dot(a, a); }
static inline Fp norm(V3 a) { return std::sqrt(norm_squared(a)); }
static inline V3 normalize(V3 a) { return a / norm(a); }
static inline V3 vinf(V3 a, V3 b) { return v3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
static inline V3 vsup(V3 a, V3 b) { return v3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
static inline Fp comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline Fp powi2(Fp x) { return x * x; }
static inline Fp powi5(Fp x) { Fp x2 = x * x; return x2 * x2 * x; }   // f64::powi(5) (llvm.powi: x * (x^2)^2)

// utils.rs:3-21
static inline Fp safe_sqrt(Fp x) { return std::sqrt(std::max(0.0, x)); }
static inline Fp chi_plus(Fp x) { return x > 0.0 ? 1.0 : 0.0; }
static inline bool almost_equal_vecs(V3 a, V3 b) { return norm(a - b) < EPS; }
static inline bool almost_equal_floats(Fp a, Fp b) { return std::fabs(a - b) < EPS; }

// ------------------------------------------------------------------------------------------------ types
struct Ray { V3 origin, direction; };                                               // geometry.rs:13-16
struct Intersection { Fp offset; V3 normal_geometry, normal_shading; bool is_outer_to_inner; };  // :19-24
struct Material { V3 base_color_factor; Fp metallic_factor, metallic_roughness; };  // scene.rs:6-11
struct Aabb {                                                                        // aabb.rs:5-51
    V3 min, max;
    static Aabb empty() { Aabb a; a.min = v3(FP_INF, FP_INF, FP_INF); a.max = v3(-FP_INF, -FP_INF, -FP_INF); return a; }
    Aabb extend_aabb(const Aabb& o) const { Aabb r; r.min = vinf(min, o.min); r.max = vsup(max, o.max); return r; }
    Fp area() const { V3 d = max - min; return d.x * d.y + d.y * d.z + d.z * d.x; }
    bool contains(const Aabb& o) const {
        for (int c = 0; c < 3; ++c) {
            if (comp(o.min, c) < comp(min, c)) return false;
            if (comp(o.max, c) > comp(max, c)) return false;
        }
        return true;
    }
};
// geometry.rs:27-46 Shape3D + Object3D and scene.rs:13-20 Primitive, flattened into one struct.  [REF] kinds: TRIANGLE, BOX
// (half sizes s).  [OWN SPEC] kinds: ELLIPSOID (radii in s), PLANE (unit normal in s; lives in Scene::infinite_primitives).
// rotation = nalgebra UnitQuaternion (i, j, k, w).  mat_kind 0 = the metallic-roughness Material of scene.rs:6-11,
// 1 = [OWN SPEC] dielectric (uses the reference's dead field `ior`, scene.rs:18).  orig_id = index in load order.
enum { SHAPE_TRIANGLE = 0, SHAPE_BOX = 1, SHAPE_ELLIPSOID = 2, SHAPE_PLANE = 3 };
enum { MAT_PBR = 0, MAT_DIELECTRIC = 1 };
struct Quat { Fp i, j, k, w; };
struct Primitive {
    V3 a, b, c, a_norm, b_norm, c_norm;   // Shape3D::Triangle
    int kind; bool identity;              // identity: rotation == 1 and position == 0 -> the Object3D transforms are exact no-ops
    int mat_kind, orig_id;
    Material material; V3 emission; Aabb aabb;
    V3 s;                                 // Shape3D::Box { s } | ellipsoid radii | plane normal
    V3 position; Quat rotation; Fp ior;   // Object3D position / rotation, Primitive.ior
};

struct Counters {
    uint64_t node_tests, tri_tests, segments, vertices, attempts, light_node_tests, light_tri_tests;
    uint64_t vndf_assert_fail, nan_pixels, samples, attempt_cap_hits;
};

// ------------------------------------------------------------------------------------------------ geometry
// [REF] geometry.rs:93-138  intersect_with_triangle: solve [b-a, c-a, -d] (u,v,t)^T = o - a with Matrix3::try_inverse.
static inline __attribute__((always_inline)) bool intersect_with_triangle(const Ray& ray, Fp upper_bound, const Primitive& p, Intersection* out,
                                                                          Fp* uo = nullptr, Fp* vo = nullptr) {
    V3 c0 = p.b - p.a, c1 = p.c - p.a, c2 = -ray.direction;
    Fp m11 = c0.x, m21 = c0.y, m31 = c0.z;
    Fp m12 = c1.x, m22 = c1.y, m32 = c1.z;
    Fp m13 = c2.x, m23 = c2.y, m33 = c2.z;
    Fp minor_m12_m23 = m22 * m33 - m32 * m23;
    Fp minor_m11_m23 = m21 * m33 - m31 * m23;
    Fp minor_m11_m22 = m21 * m32 - m31 * m22;
    Fp det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0.0) return false;
    Fp i11 = minor_m12_m23 / det, i12 = (m13 * m32 - m33 * m12) / det, i13 = (m12 * m23 - m22 * m13) / det;
    Fp i21 = -minor_m11_m23 / det, i22 = (m11 * m33 - m31 * m13) / det, i23 = (m13 * m21 - m23 * m11) / det;
    Fp i31 = minor_m11_m22 / det, i32 = (m12 * m31 - m32 * m11) / det, i33 = (m11 * m22 - m21 * m12) / det;
    V3 rhs = ray.origin - p.a;
    Fp u = i11 * rhs.x + i12 * rhs.y + i13 * rhs.z;
    Fp v = i21 * rhs.x + i22 * rhs.y + i23 * rhs.z;
    Fp t = i31 * rhs.x + i32 * rhs.y + i33 * rhs.z;
    if (0.0 <= u && 0.0 <= v && u + v <= 1.0 && 0.0 < t && t < upper_bound) {
        V3 outer_normal = normalize(cross(p.b - p.a, p.c - p.a));
        V3 ns = normalize(p.a_norm + (p.b_norm - p.a_norm) * u + (p.c_norm - p.a_norm) * v);
        bool front = dot(outer_normal, ray.direction) < 0.0;
        out->offset = t;
        out->normal_shading = front ? ns : -ns;
        out->normal_geometry = front ? outer_normal : -outer_normal;
        out->is_outer_to_inner = front;
        if (uo) *uo = u;
        if (vo) *vo = v;
        return true;
    }
    return false;
}

static inline void sort2(Fp x, Fp y, Fp* lo, Fp* hi) { if (x < y) { *lo = x; *hi = y; } else { *lo = y; *hi = x; } }   // geometry.rs:71-77
static inline Fp signum(Fp x) { return x != x ? x : (std::signbit(x) ? -1.0 : 1.0); }   // f64::signum: +-1 also for +-0, NaN for NaN

// [REF] geometry.rs:140-194 intersect_with_box: slab test in object space with `d + 0.001*EPS` denominators; up to two
// hits, entry (outer -> inner) first, each kept iff 0 < t < upper_bound; face normal chosen by `s - |p| < EPS` with the
// x, y, z priority of :161-169 (z is the fall-through).  Returns the number of hits written to out[0..2).
static int intersect_with_box(const Ray& ray, Fp upper_bound, V3 s, Intersection* out) {
    Fp tx0, tx1, ty0, ty1, tz0, tz1;
    sort2((-s.x - ray.origin.x) / (ray.direction.x + 0.001 * EPS), (s.x - ray.origin.x) / (ray.direction.x + 0.001 * EPS), &tx0, &tx1);
    sort2((-s.y - ray.origin.y) / (ray.direction.y + 0.001 * EPS), (s.y - ray.origin.y) / (ray.direction.y + 0.001 * EPS), &ty0, &ty1);
    sort2((-s.z - ray.origin.z) / (ray.direction.z + 0.001 * EPS), (s.z - ray.origin.z) / (ray.direction.z + 0.001 * EPS), &tz0, &tz1);
    Fp t_min = std::max(tx0, std::max(ty0, tz0));
    Fp t_max = std::min(tx1, std::min(ty1, tz1));
    int n = 0;
    if (t_min <= t_max) {
        auto calculate_norm = [s](V3 p) {
            if (s.x - std::fabs(p.x) < EPS) return v3(signum(p.x), 0.0, 0.0);
            if (s.y - std::fabs(p.y) < EPS) return v3(0.0, signum(p.y), 0.0);
            return v3(0.0, 0.0, signum(p.z));
        };
        if (t_min > 0.0 && t_min < upper_bound) {
            V3 nm = calculate_norm(ray.origin + ray.direction * t_min);
            out[n].offset = t_min; out[n].normal_geometry = nm; out[n].normal_shading = nm; out[n].is_outer_to_inner = true; ++n;
        }
        if (t_max > 0.0 && t_max < upper_bound) {
            V3 nm = calculate_norm(ray.origin + ray.direction * t_max);
            out[n].offset = t_max; out[n].normal_geometry = -nm; out[n].normal_shading = -nm; out[n].is_outer_to_inner = false; ++n;
        }
    }
    return n;
}

// [OWN SPEC] (no ellipsoid exists at reference HEAD) -- shaped like intersect_with_box: object
// space, up to two hits, entry first, each kept iff 0 < t < upper_bound.  |(o + t d) / r|^2 = 1 solved with the half-b
// form; outward normal = normalize(p / r^2), negated (is_outer_to_inner = false) on the exit hit.
static int intersect_with_ellipsoid(const Ray& ray, Fp upper_bound, V3 r, Intersection* out) {
    V3 od = cdiv(ray.origin, r), dd = cdiv(ray.direction, r);
    Fp a = dot(dd, dd), hb = dot(od, dd), c = dot(od, od) - 1.0;
    Fp disc = hb * hb - a * c;
    int n = 0;
    if (!(disc >= 0.0)) return 0;
    Fp sq = std::sqrt(disc);
    Fp t1 = (-hb - sq) / a, t2 = (-hb + sq) / a;
    auto outward = [r](V3 p) { return normalize(cdiv(cdiv(p, r), r)); };
    if (t1 > 0.0 && t1 < upper_bound) {
        V3 nm = outward(ray.origin + ray.direction * t1);
        out[n].offset = t1; out[n].normal_geometry = nm; out[n].normal_shading = nm; out[n].is_outer_to_inner = true; ++n;
    }
    if (t2 > 0.0 && t2 < upper_bound) {
        V3 nm = outward(ray.origin + ray.direction * t2);
        out[n].offset = t2; out[n].normal_geometry = -nm; out[n].normal_shading = -nm; out[n].is_outer_to_inner = false; ++n;
    }
    return n;
}

// [OWN SPEC] (no plane exists at reference HEAD): the plane through the object origin with unit normal nrm, object space.
// t = -(o.n)/(d.n), kept iff 0 < t < upper_bound; the normal faces the ray, is_outer_to_inner = the ray arrives from the
// side the normal points to.  d.n == 0 gives +-inf / NaN, which the comparisons reject.
static int intersect_with_plane(const Ray& ray, Fp upper_bound, V3 nrm, Intersection* out) {
    Fp dn = dot(ray.direction, nrm);
    Fp t = -dot(ray.origin, nrm) / dn;
    if (t > 0.0 && t < upper_bound) {
        bool front = dn < 0.0;
        out[0].offset = t; out[0].normal_geometry = front ? nrm : -nrm; out[0].normal_shading = out[0].normal_geometry; out[0].is_outer_to_inner = front;
        return 1;
    }
    return 0;
}

// [REF] geometry.rs:79-91 intersect_all_points: dispatch on the shape ([OWN SPEC] arms for ellipsoid / plane).
static int intersect_all_points(const Ray& ray, const Primitive& p, Fp upper_bound, Intersection* out, Fp* uo = nullptr, Fp* vo = nullptr) {
    switch (p.kind) {
        case SHAPE_BOX: return intersect_with_box(ray, upper_bound, p.s, out);
        case SHAPE_ELLIPSOID: return intersect_with_ellipsoid(ray, upper_bound, p.s, out);
        case SHAPE_PLANE: return intersect_with_plane(ray, upper_bound, p.s, out);
        default: return intersect_with_triangle(ray, upper_bound, p, out, uo, vo) ? 1 : 0;
    }
}

// nalgebra 0.32 `UnitQuaternion * Vector3` (= transform_vector):  t = 2 (q.v x p);  p' = t * q.w + q.v x t + p.
static inline V3 quat_transform(const Quat& q, V3 p) {
    V3 qv = v3(q.i, q.j, q.k);
    V3 t = cross(qv, p) * 2.0;
    V3 c = cross(qv, t);
    return t * q.w + c + p;
}
static inline Quat quat_conjugate(const Quat& q) { Quat r = {-q.i, -q.j, -q.k, q.w}; return r; }

// [REF] geometry.rs:226-251 intersect_ray_with_object3d_all_points: ray into object space (origin - position, both
// rotated by the CONJUGATE quaternion), all hits with upper bound +inf, normal_geometry rotated back -- normal_shading
// is NOT rotated back (:245-249 touch normal_geometry only).  Identity objects skip the rotations (exact no-ops).
static int intersect_ray_with_object3d_all_points(const Ray& ray, const Primitive& p, Intersection* out, Ray* rotated_ray, Fp upper_bound = FP_INF,
                                                  Fp* uo = nullptr, Fp* vo = nullptr) {
    Ray rr;
    if (p.identity) rr = ray;
    else {
        Quat qc = quat_conjugate(p.rotation);
        rr.origin = quat_transform(qc, ray.origin - p.position);
        rr.direction = quat_transform(qc, ray.direction);
    }
    int n = intersect_all_points(rr, p, upper_bound, out, uo, vo);
    if (!p.identity) for (int k = 0; k < n; ++k) out[k].normal_geometry = quat_transform(p.rotation, out[k].normal_geometry);
    if (rotated_ray) *rotated_ray = rr;
    return n;
}
// [REF] geometry.rs:196-223 intersect_ray_with_object3d: same transform, FIRST hit below min_dist (geometry.rs:51-58).
static bool intersect_ray_with_object3d(const Ray& ray, const Primitive& p, Fp min_dist, Intersection* out) {
    Intersection hits[2];
    int n = intersect_ray_with_object3d_all_points(ray, p, hits, nullptr, min_dist);
    if (n == 0) return false;
    *out = hits[0];
    return true;
}

// geometry.rs:140-194 intersect_with_box reduced to what bvh.rs:146-166 consumes: the FIRST hit of the list
// (entry if 0 < t_min < upper, else exit if 0 < t_max < upper) with its is_outer_to_inner flag.  Normals of the
// box hit are never read by the BVH code and are not computed here.
static bool box_first_hit(const Ray& local, V3 s, Fp upper_bound, Fp* t_out, bool* outer_to_inner) {
    Fp tx0, tx1, ty0, ty1, tz0, tz1;
    sort2((-s.x - local.origin.x) / (local.direction.x + 0.001 * EPS), (s.x - local.origin.x) / (local.direction.x + 0.001 * EPS), &tx0, &tx1);
    sort2((-s.y - local.origin.y) / (local.direction.y + 0.001 * EPS), (s.y - local.origin.y) / (local.direction.y + 0.001 * EPS), &ty0, &ty1);
    sort2((-s.z - local.origin.z) / (local.direction.z + 0.001 * EPS), (s.z - local.origin.z) / (local.direction.z + 0.001 * EPS), &tz0, &tz1);
    Fp t_min = std::max(tx0, std::max(ty0, tz0));
    Fp t_max = std::min(tx1, std::min(ty1, tz1));
    if (t_min <= t_max) {
        if (t_min > 0.0 && t_min < upper_bound) { *t_out = t_min; *outer_to_inner = true; return true; }
        if (t_max > 0.0 && t_max < upper_bound) { *t_out = t_max; *outer_to_inner = false; return true; }
    }
    return false;
}

// bvh.rs:157-166 get_aabb_intersection (and :146-155 intersects == .is_some()):  the AABB becomes an
// Object3D{Box{s=(max-min)/2}, position=(max+min)/2, rotation=identity} and goes through
// geometry.rs:196-223 (origin - position; the identity quaternion rotation is an exact no-op).
static bool get_aabb_intersection(const Ray& ray, const Aabb& aabb, Fp* t_out, bool* outer_to_inner) {
    V3 s = (aabb.max - aabb.min) * 0.5;
    V3 position = (aabb.max + aabb.min) * 0.5;
    Ray local; local.origin = ray.origin - position; local.direction = ray.direction;
    return box_first_hit(local, s, FP_INF, t_out, outer_to_inner);
}

static inline V3 reflect_vec(V3 v, V3 n) { Fp projection = dot(v, n); return -v + (2.0 * projection) * n; }   // geometry.rs:65-69

// ------------------------------------------------------------------------------------------------ aabb.rs
// [REF] aabb.rs:53-65 calculate_aabb_for_shape: EPS-padded local bounds ([OWN SPEC] arm: ellipsoid = +-radii; planes are
// never boxed, they live in infinite_primitives).
static Aabb calculate_aabb_for_shape(const Primitive& p) {
    V3 eps = v3(EPS, EPS, EPS);
    Aabb r;
    if (p.kind == SHAPE_TRIANGLE) { r.min = vinf(vinf(p.a, p.b), p.c) - eps; r.max = vsup(vsup(p.a, p.b), p.c) + eps; }
    else { r.min = -p.s - eps; r.max = p.s + eps; }
    return r;
}
// [REF] aabb.rs:75-94 calculate_aabb_for_object: bounds of the 8 rotated + translated corners of the shape box.  For an
// identity object the transform is an exact no-op, so min/max are the padded shape bounds themselves.
static Aabb calculate_aabb_for_object(const Primitive& p) {
    Aabb sh = calculate_aabb_for_shape(p);
    if (p.identity) return sh;
    Aabb r = Aabb::empty();
    for (int xm = 0; xm < 2; ++xm)
        for (int ym = 0; ym < 2; ++ym)
            for (int zm = 0; zm < 2; ++zm) {
                V3 point = v3(xm ? sh.max.x : sh.min.x, ym ? sh.max.y : sh.min.y, zm ? sh.max.z : sh.min.z);   // order irrelevant: min/max are commutative
                V3 op = quat_transform(p.rotation, point) + p.position;
                r.min = vinf(r.min, op); r.max = vsup(r.max, op);
            }
    return r;
}
static Aabb calculate_aabb(const Primitive* s, size_t n) {      // aabb.rs:96-106
    Aabb r = Aabb::empty();
    for (size_t i = 0; i < n; ++i) r = r.extend_aabb(calculate_aabb_for_object(s[i]));
    return r;
}

// ------------------------------------------------------------------------------------------------ bvh.rs
static const size_t NO_CHILD = (size_t)-1;
struct BvhNode { Aabb aabb; size_t left_child_index, right_child_index, content_start, content_length; };   // :11-17
struct BvhTree { std::vector<Primitive> primitives; std::vector<BvhNode> nodes; };                            // :21-24

// bvh.rs:87-144 try_split: full-sweep SAH over 3 axes; Rust's sort_by is a stable sort and the four sorts are
// applied one after another to the same slice, so the tie order of each depends on the previous one.
static bool try_split(const Aabb& aabb, Primitive* prims, size_t n, size_t* first_len) {
    if (n <= 4) return false;
    Fp best_cost = FP_INF; int best_coord = 0; size_t best_index = NO_CHILD;
    std::vector<Fp> left_prefixes, right_prefixes;
    for (int coord = 0; coord < 3; ++coord) {
        std::stable_sort(prims, prims + n, [coord](const Primitive& x, const Primitive& y) {
            Aabb ax = calculate_aabb_for_object(x), ay = calculate_aabb_for_object(y);
            return comp(ax.min + ax.max, coord) < comp(ay.min + ay.max, coord);
        });
        left_prefixes.clear(); right_prefixes.clear();
        Aabb l = Aabb::empty(), r = Aabb::empty();
        for (size_t i = 0; i + 1 < n; ++i) {
            l = l.extend_aabb(prims[i].aabb);
            r = r.extend_aabb(prims[n - 1 - i].aabb);
            left_prefixes.push_back(l.area());
            right_prefixes.push_back(r.area());
        }
        for (size_t i = 0; i + 1 < n; ++i) {
            Fp left_cost = (Fp)(i + 1) * left_prefixes[i];
            Fp right_cost = (Fp)(n - (i + 1)) * right_prefixes[n - (i + 1) - 1];
            Fp cost = left_cost + right_cost;
            if (cost < best_cost) { best_cost = cost; best_coord = coord; best_index = i; }
        }
    }
    Fp trivial_cost = aabb.area() * (Fp)n;
    if (trivial_cost < best_cost) return false;
    std::stable_sort(prims, prims + n, [best_coord](const Primitive& x, const Primitive& y) {
        Aabb ax = calculate_aabb_for_object(x), ay = calculate_aabb_for_object(y);
        return comp(ax.min + ax.max, best_coord) < comp(ay.min + ay.max, best_coord);
    });
    size_t count = best_index + 1;
    if (count >= 1 && (n - count) >= 1) { *first_len = count; return true; }
    return false;
}

// bvh.rs:47-84 create_bvh_node: children pushed before the parent (post-order, root last).
static size_t create_bvh_node(std::vector<BvhNode>& result, std::vector<Primitive>& prims, size_t start, size_t length) {
    Aabb aabb = calculate_aabb(prims.data() + start, length);
    size_t first_len = 0;
    BvhNode node; node.aabb = aabb; node.content_start = start; node.content_length = length;
    if (try_split(aabb, prims.data() + start, length, &first_len)) {
        node.left_child_index = create_bvh_node(result, prims, start, first_len);
        node.right_child_index = create_bvh_node(result, prims, start + first_len, length - first_len);
    } else {
        node.left_child_index = NO_CHILD; node.right_child_index = NO_CHILD;
    }
    result.push_back(node);
    return result.size() - 1;
}
static void create_bvh_tree(BvhTree* tree, const std::vector<Primitive>& prims) {   // bvh.rs:26-40
    tree->primitives = prims;
    tree->nodes.clear();
    create_bvh_node(tree->nodes, tree->primitives, 0, tree->primitives.size());
}

// bvh.rs:299-322 validate_bvh; returns the number of violated asserts instead of panicking.
static int validate_bvh(const BvhTree& t) {
    int bad = 0;
    for (const BvhNode& n : t.nodes) {
        if (n.left_child_index == NO_CHILD) {
            for (size_t i = n.content_start; i < n.content_start + n.content_length; ++i) {
                Aabb pa = calculate_aabb_for_object(t.primitives[i]);
                if (!(n.aabb.min.x <= pa.min.x && n.aabb.min.y <= pa.min.y && n.aabb.min.z <= pa.min.z &&
                      n.aabb.max.x >= pa.max.x && n.aabb.max.y >= pa.max.y && n.aabb.max.z >= pa.max.z)) ++bad;
            }
        } else {
            if (!n.aabb.contains(t.nodes[n.left_child_index].aabb)) ++bad;
            if (!n.aabb.contains(t.nodes[n.right_child_index].aabb)) ++bad;
        }
    }
    return bad;
}

// bvh.rs:168-172: all hits of one primitive (ArrayVec<_, 2>; hits[0] is the one that competes), the primitive, the object-space ray.
struct BvhIntersection { Intersection hits[2]; int n_hits; const Primitive* primitive; Fp u, v; Ray rotated_ray; };

// bvh.rs:249-297 nearest: unordered DFS (left, then right), prune iff best < t_box_first && entering, leaf
// prims tested with upper = +inf; a primitive competes with its FIRST hit only (points.0[0], bvh.rs:269) under strict `<`.
static void nearest_impl(const Ray& ray, const BvhTree& t, size_t idx, BvhIntersection* res, bool* found, Fp* shortest, Counters* c) {
    const BvhNode& node = t.nodes[idx];
    Fp tb; bool outer;
    ++c->node_tests;
    if (!get_aabb_intersection(ray, node.aabb, &tb, &outer)) return;
    if (*shortest < tb && outer) return;
    if (node.left_child_index == NO_CHILD) {
        for (size_t i = node.content_start; i < node.content_start + node.content_length; ++i) {
            const Primitive& prim = t.primitives[i];
            ++c->tri_tests;
            if (prim.kind == SHAPE_TRIANGLE && prim.identity) {          // what every glTF primitive is: the transforms are exact no-ops (same arithmetic as below)
                Intersection h; Fp u, v;
                if (intersect_with_triangle(ray, FP_INF, prim, &h, &u, &v) && h.offset < *shortest) {
                    *shortest = h.offset; res->hits[0] = h; res->n_hits = 1; res->primitive = &prim; res->u = u; res->v = v; *found = true;
                }
                continue;
            }
            Intersection h[2]; Fp u = 1.0 / 3.0, v = 1.0 / 3.0; Ray rr;
            int n = intersect_ray_with_object3d_all_points(ray, prim, h, &rr, FP_INF, &u, &v);
            if (n > 0 && h[0].offset < *shortest) {
                *shortest = h[0].offset; res->hits[0] = h[0]; res->hits[1] = h[1]; res->n_hits = n; res->rotated_ray = rr;
                res->primitive = &prim; res->u = u; res->v = v; *found = true;
            }
        }
    } else {
        nearest_impl(ray, t, node.left_child_index, res, found, shortest, c);
        nearest_impl(ray, t, node.right_child_index, res, found, shortest, c);
    }
}
static bool intersect_with_bvh_nearest_point(const Ray& ray, const BvhTree& t, BvhIntersection* res, Counters* c) {   // :231-247
    bool found = false; Fp nearest = FP_INF;
    if (t.nodes.empty()) return false;
    nearest_impl(ray, t, t.nodes.size() - 1, res, &found, &nearest, c);
    return found;
}

// bvh.rs:190-229 all points (no pruning), used on the light BVH by the pdf: every primitive with at least one hit
// contributes ALL its hits (both faces of a box).
static void all_points_impl(const Ray& ray, const BvhTree& t, size_t idx, std::vector<BvhIntersection>* out, Counters* c) {
    const BvhNode& node = t.nodes[idx];
    Fp tb; bool outer;
    ++c->light_node_tests;
    if (!get_aabb_intersection(ray, node.aabb, &tb, &outer)) return;       // intersects() :146-155
    if (node.left_child_index == NO_CHILD) {
        for (size_t i = node.content_start; i < node.content_start + node.content_length; ++i) {
            const Primitive& prim = t.primitives[i];
            ++c->light_tri_tests;
            if (prim.kind == SHAPE_TRIANGLE && prim.identity) {          // identity triangle: no transform, one hit, the local point is never read
                Intersection h; Fp u, v;
                if (intersect_with_triangle(ray, FP_INF, prim, &h, &u, &v)) {
                    out->emplace_back();
                    BvhIntersection& bi = out->back();
                    bi.hits[0] = h; bi.n_hits = 1; bi.primitive = &prim; bi.u = u; bi.v = v;
                }
                continue;
            }
            BvhIntersection bi; bi.u = bi.v = 1.0 / 3.0;
            bi.n_hits = intersect_ray_with_object3d_all_points(ray, prim, bi.hits, &bi.rotated_ray, FP_INF, &bi.u, &bi.v);
            if (bi.n_hits > 0) { bi.primitive = &prim; out->push_back(bi); }
        }
    } else {
        all_points_impl(ray, t, node.left_child_index, out, c);
        all_points_impl(ray, t, node.right_child_index, out, c);
    }
}

// ------------------------------------------------------------------------------------------------ RNG
struct Rng {
    uint64_t s[4]; bool has_spare; Fp spare;
    static uint64_t splitmix(uint64_t* st) { uint64_t z = (*st += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    void seed_from_u64(uint64_t seed) { uint64_t st = seed; for (int i = 0; i < 4; ++i) s[i] = splitmix(&st); has_spare = false; spare = 0; }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {                                                    // xoshiro256**
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    Fp gen_f64() { return (Fp)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }                  // rng.gen::<f64>()
    Fp gen_range01() { uint64_t b = (next_u64() >> 12) | 0x3FF0000000000000ull; Fp f; std::memcpy(&f, &b, 8); return f - 1.0; }   // gen_range(0.0..1.0)
    size_t gen_index(size_t n) {                                             // gen_range(0..n), rand 0.8 UniformInt::sample_single
        uint64_t range = (uint64_t)n;
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            unsigned __int128 m = (unsigned __int128)next_u64() * range;
            if ((uint64_t)m <= zone) return (size_t)(m >> 64);
        }
    }
    Fp gen_range(Fp low, Fp high) {                                          // gen_range(low..high), rand 0.8 UniformFloat::sample_single
        Fp scale = high - low;
        for (;;) { Fp res = gen_range01() * scale + low; if (res < high) return res; }
    }
    bool gen_bool_half() { return next_u64() < (1ull << 63); }                // gen_bool(0.5): Bernoulli p_int = 2^63, sample = (u64 < p_int)
    Fp normal() {                                                            // Normal::new(0,1).sample(rng) -- polar method, see header
        if (has_spare) { has_spare = false; return spare; }
        for (;;) {
            Fp a = 2.0 * gen_f64() - 1.0, b = 2.0 * gen_f64() - 1.0, q = a * a + b * b;
            if (q > 0.0 && q < 1.0) { Fp f = std::sqrt(-2.0 * std::log(q) / q); spare = b * f; has_spare = true; return a * f; }
        }
    }
};

// ------------------------------------------------------------------------------------------------ scene
struct Scene {                                                                // scene.rs:22-39
    int width, height, ray_depth, samples;
    V3 bg_color, camera_position, camera_forward, camera_right, camera_up;
    Fp camera_fov_x, camera_fov_y;
    BvhTree bvh_finite_primitives, bvh_light_sources;
    std::vector<Primitive> infinite_primitives;                               // scene.rs:37 (planes; empty for glTF input)
    bool has_lights;
    int max_attempts;                                                         // NOT in the reference: 0 = unbounded rejection loop (rendering.rs:102-110)
};

// ------------------------------------------------------------------------------------------------ distributions.rs
// :53-68 CosineWeightedDistribution
static V3 cosine_sample_from_sphere(V3 n, V3 uniform_unit) { return normalize(uniform_unit + n); }          // :62
static V3 cosine_sample(V3 n, Rng* rng) { Fp a = rng->normal(), b = rng->normal(), c = rng->normal(); return cosine_sample_from_sphere(n, normalize(v3(a, b, c))); }
static Fp cosine_pdf(V3 n, V3 l) { return std::max(0.0, dot(normalize(l), n)) / FP_PI; }                     // :65-67

// [REF] :70-81 get_local_pdf: 1 / surface area (box: 8 (sx sy + sy sz + sz sx); triangle: |e1 x e2| / 2).
// [OWN SPEC] ellipsoid arm: the sampler maps a uniform unit-sphere point u to r (.) u, whose area density at the
// object-space point p = r (.) u is 1 / (4 pi sqrt((u.x ry rz)^2 + (rx u.y rz)^2 + (rx ry u.z)^2)).
static Fp get_local_pdf(const Primitive& p, V3 local_point) {
    if (p.kind == SHAPE_BOX) { Fp area = (p.s.x * p.s.y + p.s.y * p.s.z + p.s.z * p.s.x) * 8.0; return 1.0 / area; }
    if (p.kind == SHAPE_ELLIPSOID) {
        V3 u = cdiv(local_point, p.s), r = p.s;
        Fp j = std::sqrt(powi2(u.x * r.y * r.z) + powi2(r.x * u.y * r.z) + powi2(r.x * r.y * u.z));
        return 1.0 / (4.0 * FP_PI * j);
    }
    Fp area = norm(cross(p.b - p.a, p.c - p.a)) * 0.5;
    return 1.0 / area;
}

// [REF] :121-124: object space -> world (rotation.transform_vector(local) + position), direction = normalize(global - point).
static V3 light_direction_from_local(const Primitive& p, V3 point, V3 local) {
    V3 global = p.identity ? local : quat_transform(p.rotation, local) + p.position;
    return normalize(global - point);
}
// [REF] :111-125 DirectLightSamplingDistribution::sample_unit_vector, triangle arm, with the two uniforms explicit.
static V3 light_sample_from_uv(const Primitive& p, V3 point, Fp u, Fp v) {
    if (!(u + v < 1.0)) { u = 1.0 - u; v = 1.0 - v; }
    V3 local = p.a + (p.b - p.a) * u + (p.c - p.a) * v;
    return light_direction_from_local(p, point, local);
}
// [REF] :86-110 box arm with the draws explicit: x in [0, wx+wy+wz) picks the face pair by area, sign picks the face,
// (c1, c2) are the two in-face coordinates in draw order (gen_range(-s.y..s.y) then (-s.z..s.z) for an x face, ...).
static V3 light_sample_box_from(const Primitive& p, V3 point, Fp x, Fp rnd_sign, Fp c1, Fp c2) {
    V3 s = p.s;
    Fp wx = 4.0 * s.y * s.z, wy = 4.0 * s.x * s.z;
    V3 local;
    if (x < wx) local = v3(s.x * rnd_sign, c1, c2);
    else if (x < wx + wy) local = v3(c1, s.y * rnd_sign, c2);
    else local = v3(c1, c2, s.z * rnd_sign);
    return light_direction_from_local(p, point, local);
}
// [OWN SPEC] ellipsoid arm: local point = radii (.) (uniform point on the unit sphere).
static V3 light_sample_ellipsoid_from(const Primitive& p, V3 point, V3 sphere_unit) { return light_direction_from_local(p, point, cmul(p.s, sphere_unit)); }

// [REF] :83-125 DirectLightSamplingDistribution::sample_unit_vector with the reference's draw order per arm.
static V3 direct_light_sample(const Primitive& p, V3 point, Rng* rng) {
    if (p.kind == SHAPE_BOX) {
        V3 s = p.s;
        Fp wx = 4.0 * s.y * s.z, wy = 4.0 * s.x * s.z, wz = 4.0 * s.x * s.y;
        Fp x = rng->gen_range(0.0, wx + wy + wz);
        Fp rnd_sign = rng->gen_bool_half() ? 1.0 : -1.0;
        Fp c1, c2;
        if (x < wx) { c1 = rng->gen_range(-s.y, s.y); c2 = rng->gen_range(-s.z, s.z); }
        else if (x < wx + wy) { c1 = rng->gen_range(-s.x, s.x); c2 = rng->gen_range(-s.z, s.z); }
        else { c1 = rng->gen_range(-s.x, s.x); c2 = rng->gen_range(-s.y, s.y); }
        return light_sample_box_from(p, point, x, rnd_sign, c1, c2);
    }
    if (p.kind == SHAPE_ELLIPSOID) {                                        // [OWN SPEC]: three normals -> unit sphere, like :55-61
        Fp a = rng->normal(), b = rng->normal(), c = rng->normal();
        return light_sample_ellipsoid_from(p, point, normalize(v3(a, b, c)));
    }
    Fp u = rng->gen_range01(), v = rng->gen_range01();
    return light_sample_from_uv(p, point, u, v);
}
// :150-158 MultipleLightSamplingDistribution::sample_unit_vector
static V3 multiple_light_sample(const Scene& sc, V3 point, Rng* rng) {
    size_t len = sc.bvh_light_sources.primitives.size();
    size_t idx = rng->gen_index(len);
    return direct_light_sample(sc.bvh_light_sources.primitives[idx], point, rng);
}
// :160-184 MultipleLightSamplingDistribution::pdf: every hit of every pierced light contributes
// local_pdf |p - point|^2 / |ng . omega| (both faces, no occlusion); the sum is divided by the light count.
static Fp multiple_light_pdf(const Scene& sc, V3 point, V3 l, Counters* c) {
    Ray ray; ray.origin = point; ray.direction = l;
    std::vector<BvhIntersection> hits;
    all_points_impl(ray, sc.bvh_light_sources, sc.bvh_light_sources.nodes.size() - 1, &hits, c);
    Fp pdf = 0.0;
    for (const BvhIntersection& bi : hits) {
        Fp sum = 0.0;
        for (int k = 0; k < bi.n_hits; ++k) {
            const Intersection& x = bi.hits[k];
            V3 global = ray.origin + x.offset * ray.direction;
            const bool flat = bi.primitive->kind == SHAPE_TRIANGLE;            // the local point matters for ellipsoids only
            Fp local_pdf = get_local_pdf(*bi.primitive, flat ? v3(0, 0, 0) : bi.rotated_ray.origin + bi.rotated_ray.direction * x.offset);
            V3 vec = global - point;
            V3 omega = normalize(vec);
            sum += local_pdf * (norm_squared(vec) / std::fabs(dot(x.normal_geometry, omega)));
        }
        pdf += sum;
    }
    return pdf / (Fp)sc.bvh_light_sources.primitives.size();
}

// :204-298 VndfDistribution
struct Frame { V3 t1, t2, n; };
static Frame vndf_frame(V3 n) {                                              // :265-268 / :277-280
    Frame f;
    f.t1 = normalize(cross(n, normalize(v3(0.234, 0.1234, 0.97686))));
    f.t2 = normalize(cross(n, f.t1));
    f.n = n;
    return f;
}
static V3 to_local(const Frame& f, V3 v) { return v3(dot(f.t1, v), dot(f.t2, v), dot(f.n, v)); }            // m.transpose() * v
static V3 to_global(const Frame& f, V3 v) { return f.t1 * v.x + f.t2 * v.y + f.n * v.z; }                   // m * v (column axpy)
static V3 sample_ggx_vndf(V3 v_local, Fp alpha, Fp U1, Fp U2) {              // :209-234
    V3 Vh = normalize(v3(alpha * v_local.x, alpha * v_local.y, v_local.z));
    Fp lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    V3 T1 = lensq > 0.0 ? v3(-Vh.y, Vh.x, 0.0) * std::sqrt(1.0 / lensq) : v3(1.0, 0.0, 0.0);
    V3 T2 = cross(Vh, T1);
    Fp r = std::sqrt(U1);
    Fp phi = 2.0 * FP_PI * U2;
    Fp t1 = r * std::cos(phi);
    Fp t2 = r * std::sin(phi);
    Fp s = 0.5 * (1.0 + Vh.z);
    t2 = (1.0 - s) * std::sqrt(1.0 - t1 * t1) + s * t2;
    V3 Nh = t1 * T1 + t2 * T2 + std::sqrt(std::max(0.0, 1.0 - t1 * t1 - t2 * t2)) * Vh;
    return normalize(v3(alpha * Nh.x, alpha * Nh.y, std::max(0.0, Nh.z)));
}
static Fp vndf_g1(V3 v, Fp alpha) { Fp under = 1.0 + powi2(alpha) * (powi2(v.x) + powi2(v.y)) / powi2(v.z); Fp lambda = (-1.0 + std::sqrt(under)) / 2.0; return 1.0 / (1.0 + lambda); }   // :236-243
static Fp vndf_dn(V3 n, Fp alpha) { Fp a2 = powi2(alpha); Fp den = FP_PI * a2 * powi2(n.x * n.x / a2 + n.y * n.y / a2 + n.z * n.z); return 1.0 / den; }                                    // :245-252
static Fp vndf_dv(V3 n, V3 v, Fp alpha) { Fp num = vndf_g1(v, alpha) * std::max(0.0, dot(v, n)) * vndf_dn(n, alpha); return num / dot(v, v3(0, 0, 1)); }                                    // :255-260
static V3 vndf_sample_from_u(V3 n, V3 v, Fp roughness, Fp U1, Fp U2, Counters* c) {   // :264-274
    Frame f = vndf_frame(n);
    V3 local_norm = sample_ggx_vndf(to_local(f, v), powi2(roughness), U1, U2);
    V3 global_norm = to_global(f, local_norm);
    V3 result = reflect_vec(v, global_norm);
    if (c && !almost_equal_vecs(global_norm, normalize(result + v))) ++c->vndf_assert_fail;   // assert! at :272
    return normalize(result);
}
// NOT in the reference: the device's VNDF sampler.  It draws the same visible-normal distribution as sample_ggx_vndf
// (:209-234) with the spherical-cap construction (Dupuy & Benyoub 2023) in a branchless tangent frame (Duff et al. 2017);
// for the isotropic alpha of the reference the distribution does not depend on the tangent frame, so it is
// distribution-equivalent to vndf_sample_from_u (streams differ anyway: the device uses Philox).  f64 restatement used
// to check the device on explicit uniforms; tests/test_oracle_kat.py checks it against the reference sampler statistically.
static V3 vndf_cap_sample_from_u(V3 n, V3 v, Fp roughness, Fp U1, Fp U2) {
    Fp sign = std::copysign(1.0, n.z), a = -1.0 / (sign + n.z), b = n.x * n.y * a;
    V3 t1 = v3(1.0 + sign * n.x * n.x * a, sign * b, -sign * n.x), t2 = v3(b, sign + n.y * n.y * a, -n.y);
    Fp alpha = powi2(roughness);
    V3 vl = v3(dot(t1, v), dot(t2, v), dot(n, v));
    V3 Vh = normalize(v3(alpha * vl.x, alpha * vl.y, vl.z));
    Fp phi = 2.0 * FP_PI * U1;
    Fp z = (1.0 - U2) * (1.0 + Vh.z) - Vh.z;
    Fp r = std::sqrt(std::min(1.0, std::max(0.0, 1.0 - z * z)));
    V3 Nh = v3(r * std::cos(phi) + Vh.x, r * std::sin(phi) + Vh.y, z + Vh.z);
    V3 Ne = normalize(v3(alpha * Nh.x, alpha * Nh.y, std::max(0.0, Nh.z)));
    V3 m = t1 * Ne.x + t2 * Ne.y + n * Ne.z;
    return normalize(reflect_vec(v, m));
}
static Fp vndf_pdf(V3 n, V3 l, V3 v, Fp roughness) {                          // :276-297 (asserts at :283-288 not evaluated)
    Frame f = vndf_frame(n);
    V3 vl = to_local(f, v), ll = to_local(f, l);
    V3 n_i = normalize(ll + vl);
    Fp alpha = powi2(roughness);
    return vndf_dv(n_i, vl, alpha) / (4.0 * dot(vl, n_i));
}

// :187-202 MixDistribution over [Cosine, Vndf] (+ MultipleLight iff lights exist, rendering.rs:23-31)
static V3 mix_sample(const Scene& sc, V3 point, V3 n, V3 v, const Material& m, Rng* rng, Counters* c) {
    size_t len = sc.has_lights ? 3 : 2;
    size_t idx = rng->gen_index(len);
    if (idx == 0) return cosine_sample(n, rng);
    if (idx == 1) { Fp U1 = rng->gen_range01(), U2 = rng->gen_range01(); return vndf_sample_from_u(n, v, m.metallic_roughness, U1, U2, c); }
    return multiple_light_sample(sc, point, rng);
}
static Fp mix_pdf(const Scene& sc, V3 point, V3 n, V3 l, V3 v, const Material& m, Counters* c) {
    Fp ans = cosine_pdf(n, l);
    ans += vndf_pdf(n, l, v, m.metallic_roughness);
    if (sc.has_lights) ans += multiple_light_pdf(sc, point, l, c);
    return ans / (Fp)(sc.has_lights ? 3 : 2);
}

// ------------------------------------------------------------------------------------------------ rendering.rs
static V3 fresnel_term(V3 f0, V3 f90, V3 h, V3 l) { return f0 + (f90 - f0) * powi5(1.0 - std::fabs(dot(h, l))); }   // :129-131

static Fp specular_brdf(V3 l, V3 n, V3 v, V3 h, const Material& m) {          // :157-184 (norm asserts :159-161 omitted)
    Fp alpha = powi2(m.metallic_roughness);
    Fp hn = dot(h, n);
    Fp d = (powi2(alpha) * chi_plus(hn)) / (FP_PI * powi2((powi2(alpha) - 1.0) * hn * hn + 1.0));
    auto g1 = [alpha](V3 n_, V3 x) {
        Fp nx = dot(n_, x);
        Fp a = (nx * chi_plus(nx)) / (alpha * safe_sqrt(1.0 - powi2(nx)));
        Fp under = 1.0 + 1.0 / (a * a);
        Fp lambda = 0.5 * (std::sqrt(under) - 1.0);
        return 1.0 / (1.0 + lambda);
    };
    Fp g = g1(n, l) * g1(n, v);
    return d * g / (4.0 * dot(l, n) * dot(v, n));
}
static V3 brdf(V3 l, V3 n, V3 v, const Material& m) {                         // :133-155
    V3 h = normalize(l + v);
    V3 diffuse = m.base_color_factor / FP_PI;
    Fp spec = specular_brdf(l, n, v, h, m);
    V3 one = v3(1.0, 1.0, 1.0);
    V3 metal = spec * fresnel_term(m.base_color_factor, one, h, l);
    V3 f = fresnel_term(v3(0.04, 0.04, 0.04), one, h, l);
    V3 diel = spec * f + cmul(diffuse, one - f);
    return metal * m.metallic_factor + diel * (1.0 - m.metallic_factor);
}

static Ray primary_ray(const Scene& sc, int x, int y, Fp xi1, Fp xi2) {       // :71-84 with the two uniforms explicit
    Fp real_x = (Fp)x + xi1, real_y = (Fp)y + xi2;
    Fp w = (Fp)sc.width, h = (Fp)sc.height;
    Fp px = (2.0 * real_x / w - 1.0) * std::tan(sc.camera_fov_x * 0.5);
    Fp py = -(2.0 * real_y / h - 1.0) * std::tan(sc.camera_fov_y * 0.5);
    Fp pz = 1.0;
    V3 dir = px * sc.camera_right + py * sc.camera_up + pz * sc.camera_forward;
    Ray r; r.origin = sc.camera_position; r.direction = normalize(dir);
    return r;
}

// [REF] rendering.rs:201-226 intersect_ray_with_scene: nearest BVH hit (first hit of the winning primitive), then a linear
// scan of infinite_primitives with the running upper bound and strict `<`.
static bool intersect_ray_with_scene(const Ray& ray, const Scene& sc, BvhIntersection* bi, Counters* c) {
    Fp latest = FP_INF;
    bool found = intersect_with_bvh_nearest_point(ray, sc.bvh_finite_primitives, bi, c);
    if (found) latest = bi->hits[0].offset;
    for (const Primitive& plane : sc.infinite_primitives) {
        Intersection x;
        ++c->tri_tests;
        if (intersect_ray_with_object3d(ray, plane, latest, &x) && x.offset < latest) {
            latest = x.offset; bi->hits[0] = x; bi->n_hits = 1; bi->primitive = &plane; bi->u = bi->v = 1.0 / 3.0; found = true;
        }
    }
    return found;
}

// [OWN SPEC] direction choice of the smooth dielectric (see get_ray_color): returns true for refraction.
static bool dielectric_sample(V3 n, V3 v, Fp ior, bool outer_to_inner, Fp u, V3* dir) {
    Fp eta = outer_to_inner ? 1.0 / ior : ior;
    Fp cos1 = dot(n, v);
    Fp sin2 = eta * std::sqrt(std::max(0.0, 1.0 - cos1 * cos1));
    bool reflect = true;
    Fp cos2 = 0.0;
    if (sin2 < 1.0) {
        cos2 = std::sqrt(1.0 - sin2 * sin2);
        Fp r0 = powi2((eta - 1.0) / (eta + 1.0));
        Fp refl = r0 + (1.0 - r0) * powi5(1.0 - cos1);
        reflect = u < refl;
    }
    *dir = reflect ? normalize(reflect_vec(v, n)) : normalize((-v) * eta + n * (eta * cos1 - cos2));
    return !reflect;
}

static V3 get_ray_color(const Ray& ray, const Scene& sc, int depth, Rng* rng, Counters* c) {   // :86-127
    if (depth <= 0) return v3(0, 0, 0);
    BvhIntersection bi;
    ++c->segments;
    if (!intersect_ray_with_scene(ray, sc, &bi, c)) return sc.bg_color;       // :96, :125
    ++c->vertices;
    const Primitive& prim = *bi.primitive;
    V3 corrected_point = ray.origin + ray.direction * (bi.hits[0].offset - EPS);
    V3 total = prim.emission;
    V3 n = bi.hits[0].normal_geometry;
    V3 v = -normalize(ray.direction);
    if (prim.mat_kind == MAT_DIELECTRIC) {
        // [OWN SPEC] (reference HEAD has no transmissive material, only the dead fields `ior` scene.rs:18
        // and `is_outer_to_inner` geometry.rs:23): smooth dielectric, delta BSDF, no rejection loop.  eta = n_from / n_to with
        // the outside index 1; Schlick reflectance; total internal reflection reflects; one uniform picks reflection (< R) or
        // refraction; the refracted ray starts EPS BEHIND the surface and is tinted by the base colour when it ENTERS.
        ++c->attempts;
        V3 dir;
        bool reflect = !dielectric_sample(n, v, prim.ior, bi.hits[0].is_outer_to_inner, rng->gen_f64(), &dir);
        Ray next; V3 weight = v3(1, 1, 1);
        next.direction = dir;
        if (reflect) next.origin = corrected_point;
        else {
            next.origin = ray.origin + ray.direction * (bi.hits[0].offset + EPS);
            if (bi.hits[0].is_outer_to_inner) weight = prim.material.base_color_factor;
        }
        return total + cmul(get_ray_color(next, sc, depth - 1, rng, c), weight);
    }
    V3 l; Fp pdf;
    for (int attempt = 1;; ++attempt) {                                       // rejection loop :102-110
        ++c->attempts;
        l = mix_sample(sc, corrected_point, n, v, prim.material, rng, c);
        pdf = mix_pdf(sc, corrected_point, n, l, v, prim.material, c);
        if (pdf > 0.0 && dot(l, bi.hits[0].normal_shading) > 0.0) break;
        // the reference spins for ever when no direction can pass (e.g. an object rotated by ~180 degrees: normal_shading is
        // not rotated back); sc.max_attempts > 0 ends the path like the device's attempt cap does (0 = unbounded, the reference)
        if (sc.max_attempts > 0 && attempt >= sc.max_attempts) { ++c->attempt_cap_hits; return total; }
    }
    Ray next; next.origin = corrected_point; next.direction = l;
    V3 refl = get_ray_color(next, sc, depth - 1, rng, c);
    V3 f = brdf(l, n, v, prim.material);
    return total + cmul(refl, f) * dot(l, n) * (1.0 / pdf);
}

static inline Fp clamp01(Fp x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }   // f64::clamp: NaN stays NaN
static V3 aces_tonemap(V3 x) {                                                // :236-248
    const Fp a = 2.51, b = 0.03, cc = 2.43, d = 0.59, e = 0.14;
    V3 unit = v3(1, 1, 1);
    V3 r = cdiv(cmul(x, a * x + b * unit), cmul(x, cc * x + d * unit) + e * unit);
    return v3(clamp01(r.x), clamp01(r.y), clamp01(r.z));
}
static inline uint8_t round_as_u8(Fp x) {                                     // f64::round (half away from zero) then saturating `as u8`, NaN -> 0
    Fp r = std::round(x);
    if (!(r == r)) return 0;
    if (r <= 0.0) return 0;
    if (r >= 255.0) return 255;
    return (uint8_t)r;
}
static void color_to_pixel(V3 c, uint8_t* out) {                              // :250-262
    V3 t = aces_tonemap(c);
    out[0] = round_as_u8(std::pow(t.x, 1.0 / 2.2) * 255.0);
    out[1] = round_as_u8(std::pow(t.y, 1.0 / 2.2) * 255.0);
    out[2] = round_as_u8(std::pow(t.z, 1.0 / 2.2) * 255.0);
}

// ------------------------------------------------------------------------------------------------ C ABI (ctypes)
extern "C" {

struct OrSceneDesc {
    int32_t width, height, samples, ray_depth;
    double bg_color[3], camera_position[3], camera_forward[3], camera_right[3], camera_up[3];
    double camera_fov_x, camera_fov_y;
    int32_t n_tris, _pad;
    const double* tri_v;          // n x 9  (a, b, c)
    const double* tri_n;          // n x 9  (a_norm, b_norm, c_norm)
    const double* tri_material;   // n x 5  (base r g b, metallic, roughness)
    const double* tri_emission;   // n x 3
    // general primitives (all may be null = triangles with identity transforms, what the glTF loader emits):
    const int32_t* kind;          // n: SHAPE_* ; for kind != TRIANGLE the first 3 doubles of tri_v are s / radii / plane normal
    const double* position;       // n x 3  Object3D.position
    const double* rotation;       // n x 4  Object3D.rotation as (i, j, k, w)
    const double* ior;            // n      Primitive.ior (scene.rs:18)
    const int32_t* mat_kind;      // n: MAT_PBR | MAT_DIELECTRIC
    int32_t max_attempts, _pad2;  // 0 = unbounded rejection loop (the reference)
};
struct OrInfo { int64_t n_nodes, n_leaves, depth, n_lights, n_light_nodes, validate_failures; };
struct OrStats { uint64_t node_tests, tri_tests, segments, vertices, attempts, light_node_tests, light_tri_tests, vndf_assert_fail, nan_pixels, samples; double seconds; uint64_t attempt_cap_hits; };

static V3 rd3(const double* p) { return v3(p[0], p[1], p[2]); }
static void wr3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

void* or_scene_create(const OrSceneDesc* d) {
    Scene* sc = new Scene();
    sc->width = d->width; sc->height = d->height; sc->samples = d->samples; sc->ray_depth = d->ray_depth;
    sc->bg_color = rd3(d->bg_color); sc->camera_position = rd3(d->camera_position); sc->camera_forward = rd3(d->camera_forward);
    sc->camera_right = rd3(d->camera_right); sc->camera_up = rd3(d->camera_up);
    sc->camera_fov_x = d->camera_fov_x; sc->camera_fov_y = d->camera_fov_y;
    sc->max_attempts = d->max_attempts;
    std::vector<Primitive> finite, lights;
    for (int i = 0; i < d->n_tris; ++i) {
        Primitive p;
        p.kind = d->kind ? d->kind[i] : SHAPE_TRIANGLE;
        p.s = rd3(d->tri_v + 9 * i);
        p.position = d->position ? rd3(d->position + 3 * i) : v3(0, 0, 0);
        Quat q = {0, 0, 0, 1};
        if (d->rotation) { q.i = d->rotation[4 * i]; q.j = d->rotation[4 * i + 1]; q.k = d->rotation[4 * i + 2]; q.w = d->rotation[4 * i + 3]; }
        p.rotation = q;
        p.identity = q.i == 0 && q.j == 0 && q.k == 0 && q.w == 1 && p.position.x == 0 && p.position.y == 0 && p.position.z == 0;
        p.ior = d->ior ? d->ior[i] : 1.0;
        p.mat_kind = d->mat_kind ? d->mat_kind[i] : MAT_PBR;
        p.a = rd3(d->tri_v + 9 * i); p.b = rd3(d->tri_v + 9 * i + 3); p.c = rd3(d->tri_v + 9 * i + 6);
        p.a_norm = rd3(d->tri_n + 9 * i); p.b_norm = rd3(d->tri_n + 9 * i + 3); p.c_norm = rd3(d->tri_n + 9 * i + 6);
        p.material.base_color_factor = rd3(d->tri_material + 5 * i);
        p.material.metallic_factor = d->tri_material[5 * i + 3];
        p.material.metallic_roughness = d->tri_material[5 * i + 4];
        p.emission = rd3(d->tri_emission + 3 * i);
        p.orig_id = i;
        if (p.kind == SHAPE_PLANE) { p.aabb = Aabb::empty(); sc->infinite_primitives.push_back(p); continue; }   // scene.rs:37; [OWN SPEC] planes are never lights
        p.aabb = calculate_aabb_for_object(p);                         // gltf_to_scene.rs:234
        finite.push_back(p);
        if (norm(p.emission) > EPS) lights.push_back(p);               // gltf_to_scene.rs:240
    }
    create_bvh_tree(&sc->bvh_finite_primitives, finite);               // gltf_to_scene.rs:72
    create_bvh_tree(&sc->bvh_light_sources, lights);                   // gltf_to_scene.rs:77
    sc->has_lights = !sc->bvh_light_sources.primitives.empty();        // rendering.rs:27
    return sc;
}
void or_scene_destroy(void* h) { delete (Scene*)h; }

static int64_t tree_depth(const BvhTree& t, size_t idx) {
    const BvhNode& n = t.nodes[idx];
    if (n.left_child_index == NO_CHILD) return 1;
    return 1 + std::max(tree_depth(t, n.left_child_index), tree_depth(t, n.right_child_index));
}
void or_scene_info(void* h, OrInfo* o) {
    Scene* sc = (Scene*)h;
    const BvhTree& t = sc->bvh_finite_primitives;
    o->n_nodes = (int64_t)t.nodes.size(); o->n_leaves = 0;
    for (const BvhNode& n : t.nodes) if (n.left_child_index == NO_CHILD) ++o->n_leaves;
    o->depth = t.nodes.empty() ? 0 : tree_depth(t, t.nodes.size() - 1);
    o->n_lights = (int64_t)sc->bvh_light_sources.primitives.size();
    o->n_light_nodes = (int64_t)sc->bvh_light_sources.nodes.size();
    o->validate_failures = validate_bvh(t) + validate_bvh(sc->bvh_light_sources);     // rendering.rs:22
}
// BVH-ordered original triangle ids (bvh_finite_primitives.primitives after the build's sorts).
void or_scene_bvh_order(void* h, int32_t* ids) {
    Scene* sc = (Scene*)h;
    for (size_t i = 0; i < sc->bvh_finite_primitives.primitives.size(); ++i) ids[i] = sc->bvh_finite_primitives.primitives[i].orig_id;
}
// Node dump: per node min[3], max[3], then left, right, start, length as doubles (-1 for no child).
void or_scene_bvh_nodes(void* h, double* out) {
    Scene* sc = (Scene*)h;
    const BvhTree& t = sc->bvh_finite_primitives;
    for (size_t i = 0; i < t.nodes.size(); ++i) {
        const BvhNode& n = t.nodes[i]; double* o = out + 10 * i;
        wr3(o, n.aabb.min); wr3(o + 3, n.aabb.max);
        o[6] = n.left_child_index == NO_CHILD ? -1.0 : (double)n.left_child_index;
        o[7] = n.right_child_index == NO_CHILD ? -1.0 : (double)n.right_child_index;
        o[8] = (double)n.content_start; o[9] = (double)n.content_length;
    }
}

// render_scene (rendering.rs:21-69) over rows y0, y0+y_step, ... < y1 (the reference renders 0..height).  One
// xoshiro256** stream per row seeded with (width*y) as u64 (:50-51); `seed` != 0 perturbs it so independent
// runs exist for noise-floor estimates (seed 0 == the reference seeding).  Rows are handed out dynamically to
// n_threads std::threads (the reference: rayon par_iter over rows).  Outputs are full-frame sized; rows not
// rendered are left untouched.  mean/var: per-pixel linear radiance mean and per-channel sample variance.
int or_render(void* h, uint64_t seed, int n_threads, int y0, int y1, int y_step, uint8_t* rgb, double* mean, double* var, OrStats* stats) {
    Scene* sc = (Scene*)h;
    if (sc->samples <= 0 || sc->width <= 0 || sc->height <= 0) return 1;
    if (y_step <= 0) y_step = 1;
    if (y0 < 0) y0 = 0;
    if (y1 > sc->height) y1 = sc->height;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::vector<int> rows;
    for (int y = y0; y < y1; y += y_step) rows.push_back(y);
    std::atomic<size_t> next(0);
    std::vector<Counters> counters(n_threads);
    auto t_begin = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        Counters c; std::memset(&c, 0, sizeof(c));
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= rows.size()) break;
            int y = rows[k];
            Rng rng; rng.seed_from_u64((uint64_t)(int64_t)(int32_t)(sc->width * y) ^ (seed * 0x9E3779B97F4A7C15ull));
            for (int x = 0; x < sc->width; ++x) {
                V3 total = v3(0, 0, 0), sq = v3(0, 0, 0);
                for (int s = 0; s < sc->samples; ++s) {
                    Fp xi1 = rng.gen_f64(); Fp xi2 = rng.gen_f64();
                    Ray r = primary_ray(*sc, x, y, xi1, xi2);
                    V3 col = get_ray_color(r, *sc, sc->ray_depth, &rng, &c);
                    total = s == 0 ? col : total + col;
                    sq = sq + cmul(col, col);
                    ++c.samples;
                }
                V3 color = total / (Fp)sc->samples;
                size_t pix = (size_t)y * sc->width + x;
                if (!(color.x == color.x && color.y == color.y && color.z == color.z)) ++c.nan_pixels;
                if (rgb) color_to_pixel(color, rgb + 3 * pix);
                if (mean) wr3(mean + 3 * pix, color);
                if (var) {
                    Fp n = (Fp)sc->samples, dn = n > 1 ? n - 1 : 1;
                    wr3(var + 3 * pix, v3((sq.x - n * color.x * color.x) / dn, (sq.y - n * color.y * color.y) / dn, (sq.z - n * color.z * color.z) / dn));
                }
            }
        }
        counters[tid] = c;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) th.emplace_back(worker, i);
    worker(0);
    for (auto& t : th) t.join();
    auto t_end = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (const Counters& c : counters) {
            stats->node_tests += c.node_tests; stats->tri_tests += c.tri_tests; stats->segments += c.segments; stats->vertices += c.vertices;
            stats->attempts += c.attempts; stats->light_node_tests += c.light_node_tests; stats->light_tri_tests += c.light_tri_tests;
            stats->vndf_assert_fail += c.vndf_assert_fail; stats->nan_pixels += c.nan_pixels; stats->samples += c.samples;
            stats->attempt_cap_hits += c.attempt_cap_hits;
        }
        stats->seconds = std::chrono::duration<double>(t_end - t_begin).count();
    }
    return 0;
}

// Nearest hit of n rays (o[3], d[3] each) through intersect_ray_with_scene (BVH, then the infinite primitives).
// tri_id = original (load-order) primitive index or -1; t of the winner; (u, v) its barycentrics (1/3, 1/3 for shapes
// without edges); second_t = nearest FIRST hit among all OTHER primitives (brute force, +inf if none) so callers can
// recognise ties (SURVEY.md 8d parity rule).
void or_trace_primary(void* h, const double* rays, int64_t n, int32_t* tri_id, double* t, double* u, double* v, double* second_t, OrStats* stats) {
    Scene* sc = (Scene*)h;
    Counters c; std::memset(&c, 0, sizeof(c));
    const BvhTree& tree = sc->bvh_finite_primitives;
    for (int64_t i = 0; i < n; ++i) {
        Ray r; r.origin = rd3(rays + 6 * i); r.direction = rd3(rays + 6 * i + 3);
        BvhIntersection bi; ++c.segments;
        bool hit = intersect_ray_with_scene(r, *sc, &bi, &c);
        tri_id[i] = hit ? bi.primitive->orig_id : -1;
        t[i] = hit ? bi.hits[0].offset : FP_INF;
        if (u) u[i] = hit ? bi.u : 0.0;
        if (v) v[i] = hit ? bi.v : 0.0;
        if (second_t) {
            Fp best = FP_INF;
            auto consider = [&](const Primitive& p) {
                if (hit && p.orig_id == bi.primitive->orig_id) return;
                Intersection x[2];
                if (intersect_ray_with_object3d_all_points(r, p, x, nullptr) > 0 && x[0].offset < best) best = x[0].offset;
            };
            for (const Primitive& p : tree.primitives) consider(p);
            for (const Primitive& p : sc->infinite_primitives) consider(p);
            second_t[i] = best;
        }
    }
    if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->node_tests = c.node_tests; stats->tri_tests = c.tri_tests; stats->segments = c.segments; }
}

// Full hit record for n rays: t, ng[3], ns[3], orig id, is_outer_to_inner (as doubles) -> 9 doubles per ray (t = inf on miss).
void or_trace_hits(void* h, const double* rays, int64_t n, double* out) {
    Scene* sc = (Scene*)h; Counters c; std::memset(&c, 0, sizeof(c));
    for (int64_t i = 0; i < n; ++i) {
        Ray r; r.origin = rd3(rays + 6 * i); r.direction = rd3(rays + 6 * i + 3);
        BvhIntersection bi; double* o = out + 9 * i;
        if (intersect_ray_with_scene(r, *sc, &bi, &c)) {
            o[0] = bi.hits[0].offset; wr3(o + 1, bi.hits[0].normal_geometry); wr3(o + 4, bi.hits[0].normal_shading); o[7] = (double)bi.primitive->orig_id;
            o[8] = bi.hits[0].is_outer_to_inner ? 1.0 : 0.0;
        } else { o[0] = FP_INF; for (int k = 1; k < 7; ++k) o[k] = 0.0; o[7] = -1.0; o[8] = 0.0; }
    }
}

void or_primary_rays(void* h, const int32_t* xy, const double* xi, int64_t n, double* rays_out) {
    Scene* sc = (Scene*)h;
    for (int64_t i = 0; i < n; ++i) {
        Ray r = primary_ray(*sc, xy[2 * i], xy[2 * i + 1], xi[2 * i], xi[2 * i + 1]);
        wr3(rays_out + 6 * i, r.origin); wr3(rays_out + 6 * i + 3, r.direction);
    }
}

// ---- unit functions, batched.  mat = (base r g b, metallic, roughness).
void or_brdf(const double* l, const double* n, const double* v, const double* mat, int64_t cnt, double* out) {
    for (int64_t i = 0; i < cnt; ++i) {
        Material m; m.base_color_factor = rd3(mat + 5 * i); m.metallic_factor = mat[5 * i + 3]; m.metallic_roughness = mat[5 * i + 4];
        wr3(out + 3 * i, brdf(rd3(l + 3 * i), rd3(n + 3 * i), rd3(v + 3 * i), m));
    }
}
void or_specular_brdf(const double* l, const double* n, const double* v, const double* hh, const double* rough, int64_t cnt, double* out) {
    for (int64_t i = 0; i < cnt; ++i) { Material m; m.base_color_factor = v3(1, 1, 1); m.metallic_factor = 1; m.metallic_roughness = rough[i]; out[i] = specular_brdf(rd3(l + 3 * i), rd3(n + 3 * i), rd3(v + 3 * i), rd3(hh + 3 * i), m); }
}
void or_pdf_cosine(const double* n, const double* l, int64_t cnt, double* out) { for (int64_t i = 0; i < cnt; ++i) out[i] = cosine_pdf(rd3(n + 3 * i), rd3(l + 3 * i)); }
void or_pdf_vndf(const double* n, const double* l, const double* v, const double* rough, int64_t cnt, double* out) { for (int64_t i = 0; i < cnt; ++i) out[i] = vndf_pdf(rd3(n + 3 * i), rd3(l + 3 * i), rd3(v + 3 * i), rough[i]); }
void or_pdf_light(void* h, const double* point, const double* l, int64_t cnt, double* out) {
    Scene* sc = (Scene*)h; Counters c; std::memset(&c, 0, sizeof(c));
    for (int64_t i = 0; i < cnt; ++i) out[i] = sc->has_lights ? multiple_light_pdf(*sc, rd3(point + 3 * i), rd3(l + 3 * i), &c) : 0.0;
}
void or_pdf_mix(void* h, const double* point, const double* n, const double* l, const double* v, const double* mat, int64_t cnt, double* out) {
    Scene* sc = (Scene*)h; Counters c; std::memset(&c, 0, sizeof(c));
    for (int64_t i = 0; i < cnt; ++i) {
        Material m; m.base_color_factor = rd3(mat + 5 * i); m.metallic_factor = mat[5 * i + 3]; m.metallic_roughness = mat[5 * i + 4];
        out[i] = mix_pdf(*sc, rd3(point + 3 * i), rd3(n + 3 * i), rd3(l + 3 * i), rd3(v + 3 * i), m, &c);
    }
}
void or_sample_cosine(const double* n, const double* sphere_unit, int64_t cnt, double* out) { for (int64_t i = 0; i < cnt; ++i) wr3(out + 3 * i, cosine_sample_from_sphere(rd3(n + 3 * i), rd3(sphere_unit + 3 * i))); }
void or_sample_vndf(const double* n, const double* v, const double* rough, const double* u12, int64_t cnt, double* out) {
    for (int64_t i = 0; i < cnt; ++i) wr3(out + 3 * i, vndf_sample_from_u(rd3(n + 3 * i), rd3(v + 3 * i), rough[i], u12[2 * i], u12[2 * i + 1], nullptr));
}
void or_sample_vndf_cap(const double* n, const double* v, const double* rough, const double* u12, int64_t cnt, double* out) {
    for (int64_t i = 0; i < cnt; ++i) wr3(out + 3 * i, vndf_cap_sample_from_u(rd3(n + 3 * i), rd3(v + 3 * i), rough[i], u12[2 * i], u12[2 * i + 1]));
}
// light_idx indexes the light list in LOAD order (ascending original primitive id).  draws: 4 doubles per sample --
// triangle: (u, v, -, -); box: (x in [0,1) scaled by wx+wy+wz, sign +-1, c1, c2 in [-1,1) scaled by the face's half sizes);
// ellipsoid: (sphere_unit xyz, -).
void or_sample_light(void* h, const int32_t* light_idx, const double* point, const double* draws, int64_t cnt, double* out) {
    Scene* sc = (Scene*)h;
    std::vector<const Primitive*> by_load;
    for (const Primitive& p : sc->bvh_light_sources.primitives) by_load.push_back(&p);
    std::sort(by_load.begin(), by_load.end(), [](const Primitive* a, const Primitive* b) { return a->orig_id < b->orig_id; });
    for (int64_t i = 0; i < cnt; ++i) {
        const Primitive& p = *by_load[light_idx[i]];
        const double* q = draws + 4 * i;
        V3 pt = rd3(point + 3 * i), l;
        if (p.kind == SHAPE_BOX) {
            V3 s = p.s;
            Fp wx = 4.0 * s.y * s.z, wy = 4.0 * s.x * s.z, wz = 4.0 * s.x * s.y;
            Fp x = q[0] * (wx + wy + wz);
            Fp h1 = x < wx ? s.y : s.x, h2 = x < wx + wy ? s.z : s.y;
            l = light_sample_box_from(p, pt, x, q[1], q[2] * h1, q[3] * h2);
        } else if (p.kind == SHAPE_ELLIPSOID) l = light_sample_ellipsoid_from(p, pt, rd3(q));
        else l = light_sample_from_uv(p, pt, q[0], q[1]);
        wr3(out + 3 * i, l);
    }
}
// Object-space intersection of one shape: kind, params (9 doubles like tri_v), ray -> n hits, per hit t, normal[3], outer flag (5 doubles).
int or_intersect_shape(int32_t kind, const double* params, const double* o, const double* d, double upper, double* out10) {
    Primitive p; std::memset(&p, 0, sizeof(p));
    p.kind = kind; p.s = rd3(params); p.a = rd3(params); p.b = rd3(params + 3); p.c = rd3(params + 6);
    V3 ng = cross(p.b - p.a, p.c - p.a); p.a_norm = p.b_norm = p.c_norm = ng;
    p.identity = true; p.rotation = Quat{0, 0, 0, 1};
    Ray r; r.origin = rd3(o); r.direction = rd3(d);
    Intersection x[2];
    int n = intersect_all_points(r, p, upper, x);
    for (int k = 0; k < n; ++k) { out10[5 * k] = x[k].offset; wr3(out10 + 5 * k + 1, x[k].normal_geometry); out10[5 * k + 4] = x[k].is_outer_to_inner ? 1.0 : 0.0; }
    return n;
}
// [OWN SPEC] dielectric direction choice: n, v, (ior, outer, u) per sample -> dir xyz, refracted flag.
void or_dielectric(const double* n, const double* v, const double* iou, int64_t cnt, double* out4) {
    for (int64_t i = 0; i < cnt; ++i) {
        V3 dir;
        bool refr = dielectric_sample(rd3(n + 3 * i), rd3(v + 3 * i), iou[3 * i], iou[3 * i + 1] > 0.0, iou[3 * i + 2], &dir);
        wr3(out4 + 4 * i, dir); out4[4 * i + 3] = refr ? 1.0 : 0.0;
    }
}
void or_quat_transform(const double* q_ijkw, const double* v, int32_t conjugate, double* out) {
    Quat q = {q_ijkw[0], q_ijkw[1], q_ijkw[2], q_ijkw[3]};
    wr3(out, quat_transform(conjugate ? quat_conjugate(q) : q, rd3(v)));
}
// calculate_aabb_for_object (aabb.rs:75-94) of one primitive given like or_scene_create's arrays -> min[3], max[3].
void or_object_aabb(int32_t kind, const double* params, const double* position, const double* q_ijkw, double* out6) {
    Primitive p; std::memset(&p, 0, sizeof(p));
    p.kind = kind; p.s = rd3(params); p.a = rd3(params); p.b = rd3(params + 3); p.c = rd3(params + 6);
    p.position = rd3(position); p.rotation = Quat{q_ijkw[0], q_ijkw[1], q_ijkw[2], q_ijkw[3]}; p.identity = false;
    Aabb a = calculate_aabb_for_object(p);
    wr3(out6, a.min); wr3(out6 + 3, a.max);
}
void or_color_to_pixel(const double* rgb, int64_t cnt, uint8_t* out) { for (int64_t i = 0; i < cnt; ++i) color_to_pixel(rd3(rgb + 3 * i), out + 3 * i); }
int or_intersect_triangle(const double* o, const double* d, const double* abc, double* tuv) {
    Primitive p; std::memset(&p, 0, sizeof(p));
    p.a = rd3(abc); p.b = rd3(abc + 3); p.c = rd3(abc + 6); p.a_norm = p.b_norm = p.c_norm = v3(0, 0, 1);
    Ray r; r.origin = rd3(o); r.direction = rd3(d);
    Intersection x; Fp u, v;
    if (!intersect_with_triangle(r, FP_INF, p, &x, &u, &v)) return 0;
    tuv[0] = x.offset; tuv[1] = u; tuv[2] = v; return 1;
}
int or_aabb_first_hit(const double* o, const double* d, const double* mn, const double* mx, double* t, int32_t* outer) {
    Ray r; r.origin = rd3(o); r.direction = rd3(d);
    Aabb a; a.min = rd3(mn); a.max = rd3(mx);
    Fp tt; bool oo;
    if (!get_aabb_intersection(r, a, &tt, &oo)) return 0;
    *t = tt; *outer = oo ? 1 : 0; return 1;
}
// xoshiro256** known-answer hook: first `cnt` outputs after seed_from_u64(seed).
void or_rng_u64(uint64_t seed, int64_t cnt, uint64_t* out) { Rng r; r.seed_from_u64(seed); for (int64_t i = 0; i < cnt; ++i) out[i] = r.next_u64(); }

}  // extern "C"
