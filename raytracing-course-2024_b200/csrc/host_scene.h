// host_scene.h -- host-side scene types shared by the loader, the BVH builder and the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rtb {

// The reference's `Scene` (src/scene.rs:22-39) flattened, f64 like the reference (geometry.rs:5), load order.
struct HostScene {
    int32_t width = 0, height = 0, samples = 0, ray_depth = 6;
    double bg_color[3] = {0, 0, 0};
    double camera_position[3] = {0, 0, 0}, camera_forward[3] = {0, 0, 0}, camera_right[3] = {0, 0, 0}, camera_up[3] = {0, 0, 0};
    double camera_fov_x = 0, camera_fov_y = 0;
    std::vector<double> tri_v;         // n x 9
    std::vector<double> tri_n;         // n x 9
    std::vector<double> tri_material;  // n x 5  (base rgb, metallic, roughness)
    std::vector<double> tri_emission;  // n x 3
    // General primitives (Object3D + Shape3D::Box, geometry.rs:27-46; text-scene shapes and the dielectric material: own spec).
    // All empty = every primitive is a world-space triangle with the metallic-roughness material (what the glTF loader emits).
    // For kind != triangle the first three doubles of the tri_v row are the box half sizes / ellipsoid radii / plane normal.
    std::vector<int32_t> kind;         // n: RT_SHAPE_*
    std::vector<double> position;      // n x 3
    std::vector<double> rotation;      // n x 4: unit quaternion (i, j, k, w)
    std::vector<double> ior;           // n
    std::vector<int32_t> mat_kind;     // n: RT_MATERIAL_*
    int32_t n_tris() const { return (int32_t)(tri_v.size() / 9); }   // number of primitives of any kind
    bool general() const { return !kind.empty(); }
};

// 3x3 rotation matrix (row major) of a unit quaternion (i, j, k, w): nalgebra UnitQuaternion::transform_vector as a matrix.
void quat_to_matrix(const double* q, double* m9);

// Loader status: ok, or a message + an RT_ERR_* class.
struct LoadError { int code = 0; std::string message; };

// main.rs:45-47: gltf::import(path) + convert_gltf_to_scene(..., width, height, samples).
bool load_gltf_scene(const std::string& path, int32_t width, int32_t height, int32_t samples, HostScene* out, LoadError* err);
// The course's text scene format (text_loader.cpp; no parser exists at reference HEAD -- own spec).
// width / height / samples > 0 override DIMENSIONS / SAMPLES.
bool load_text_scene(const std::string& path, int32_t width, int32_t height, int32_t samples, HostScene* out, LoadError* err);

// ---- flattened device-layout BVH (built on the host, bvh_builder.cpp) ---------------------------------------
// Binary BVH stored as "child-pair" nodes: inner node i keeps the boxes of BOTH children, so one visit tests two
// boxes and descends without fetching the children's own nodes.  A child reference is >= 0 for an inner node
// index and < 0 for a leaf: ~((first_tri << 3) | (count - 1)), first_tri indexing the BVH-ordered triangle
// arrays, count in 1..8.  Node 0 is the root.  Boxes are the reference's EPS-padded triangle bounds
// (aabb.rs:53-65) rounded OUTWARD to float, so the device slab test is conservative w.r.t. the f64 boxes.
struct FlatBvh {
    std::vector<float> box_a;     // n_nodes x 4: c0.min.x, c0.max.x, c0.min.y, c0.max.y
    std::vector<float> box_b;     // n_nodes x 4: c1.min.x, c1.max.x, c1.min.y, c1.max.y
    std::vector<float> box_c;     // n_nodes x 4: c0.min.z, c0.max.z, c1.min.z, c1.max.z
    std::vector<int32_t> child;   // n_nodes x 2
    std::vector<int32_t> tri_order;  // BVH order -> original triangle id
    int32_t n_nodes = 0, n_leaves = 0, depth = 0, max_leaf = 0;
};

struct BoxD { double mn[3], mx[3]; };
// aabb.rs:53-65 for a triangle: min/max of the vertices -/+ EPS.
BoxD tri_box_d(const double* v9);

struct BvhBuildParams {
    int max_leaf_size = 4;       // bvh.rs:89 (n <= 4 never splits)
    double traversal_cost = 1.0; // cost of one box test relative to one triangle test in the SAH
    bool size_split = false;     // experiment: the sweep also tries the order by box area (largest first) as a fourth axis
    bool force_median = false;   // object-median splits on the longest axis instead of any cost heuristic (fallback for degenerate input)
    bool reinsertion = true;     // (with agglomerative) subtree re-insertion passes after the clustering
    bool agglomerative = true;   // sets of <= 512 primitives: greedy bottom-up clustering instead of the top-down sweep (bvh_builder.cpp)
};

// boxes: per primitive id (load order), EPS-padded like aabb.rs:53-94; ids: which primitives to include (the finite ones, or the lights).
// weights (optional, per primitive id): SAH weight of a primitive (default 1; regraft_top_sah passes the triangle count of a subtree).
void build_bvh(const std::vector<BoxD>& boxes, const std::vector<int32_t>& ids, const BvhBuildParams& p, FlatBvh* out, const std::vector<double>* weights = nullptr);
// Rebuilds the TOP of a tree (above a cut of ~n_clusters subtrees, largest surface area first) with the SAH sweep and grafts the
// subtrees back: the repair of the GPU LBVH's globally poor splits.  n_clusters is capped at a third of the node count; < 8: no-op.
void regraft_top_sah(FlatBvh* bvh, int n_clusters, const BvhBuildParams& p);
// GPU builder (rt_bvh_gpu.cu): Morton codes + radix sort + Karras hierarchy + refit + collapse to leaves <= max_leaf_size.
// Same output format; lower tree quality than the SAH sweep, milliseconds instead of seconds on large meshes.  Returns
// false (with *err) when CUDA fails or the set is too small to bother (<= 2 * max_leaf_size triangles).
bool build_bvh_gpu(const double* tri_v, int n_total_tris, const std::vector<int32_t>& ids, const BvhBuildParams& p, int device, FlatBvh* out,
                   double* build_ms, std::string* err);
// validate_bvh (bvh.rs:299-322) on the flattened tree: every leaf box contains its triangles' EPS-padded
// boxes, every inner pair box contains the boxes stored in the child node.  Returns the number of violations.
int validate_flat_bvh(const FlatBvh& bvh, const std::vector<BoxD>& boxes);

}  // namespace rtb
