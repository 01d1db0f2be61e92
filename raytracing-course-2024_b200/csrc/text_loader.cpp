// text_loader.cpp -- the course's text scene format -> HostScene.
//
// OWN SPEC.  Reference HEAD has no parser for this format: src/main.rs:45 calls gltf::import unconditionally and :48 keeps
// only the comment `// let scene = parse_file_content(file_lines);`.  The grammar is recovered from the scene files the
// reference still ships (scenes/practice3_1.txt:1-25, practice3_5.txt:1-52, practice3_2.txt:26 ROTATION, practice3_3.txt:47
// METALLIC, practice3_4.txt:47-48 DIELECTRIC / IOR, working.txt TRIANGLE); the mapping onto the reference's HEAD data model
// (scene.rs:6-39, geometry.rs:27-46) is declared in DESIGN.md section 12 and restated in oracle/text_ref.py, against which
// this loader is compared value by value (tests/test_text_scene.py).
//
// Grammar: line oriented, whitespace separated, blank lines and unknown keywords ignored.
//   header     DIMENSIONS w h | RAY_DEPTH n | SAMPLES n | BG_COLOR r g b | CAMERA_POSITION x y z | CAMERA_RIGHT x y z |
//              CAMERA_UP x y z | CAMERA_FORWARD x y z | CAMERA_FOV_X radians
//   primitive  NEW_PRIMITIVE then any of: PLANE nx ny nz | ELLIPSOID rx ry rz | BOX sx sy sz | TRIANGLE a b c (9 numbers) |
//              POSITION x y z | ROTATION x y z w | COLOR r g b | EMISSION r g b | METALLIC | DIELECTRIC | IOR x
// Mapping:
//   fov_y = 2 atan(tan(fov_x / 2) h / w);  ROTATION (x y z w) -> unit quaternion (i, j, k, w), normalised;
//   PLANE normals normalised, planes become Scene::infinite_primitives (scene.rs:37);
//   TRIANGLE vertex normals = its face normal;
//   Material (scene.rs:6-11): COLOR -> base_color_factor (default 0); default -> metallic 0, roughness 1; METALLIC -> metallic 1,
//   roughness 0.03 (the glTF loader's floor, gltf_to_scene.rs:221); DIELECTRIC -> RT_MATERIAL_DIELECTRIC with IOR (scene.rs:18).
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/rt_api.h"
#include "host_scene.h"

namespace rtb {
namespace {

struct TextPrim {
    int kind = -1;
    double shape[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double position[3] = {0, 0, 0}, rotation[4] = {0, 0, 0, 1}, color[3] = {0, 0, 0}, emission[3] = {0, 0, 0};
    bool metallic = false, dielectric = false;
    double ior = 1.0;
};

// Reads up to `n` numbers from the tokens after the keyword; false if a token is not a number.
bool numbers(const std::vector<std::string>& tok, size_t n, double* out, size_t* got) {
    size_t k = 0;
    for (; k < n && k + 1 < tok.size(); ++k) {
        const char* s = tok[k + 1].c_str();
        char* end = nullptr;
        out[k] = std::strtod(s, &end);
        if (end == s || *end != '\0') return false;
    }
    if (got) *got = k;
    return true;
}

}  // namespace

bool load_text_scene(const std::string& path, int32_t width, int32_t height, int32_t samples, HostScene* out, LoadError* err) {
    std::ifstream f(path);
    if (!f) { err->code = RT_ERR_IO; err->message = "cannot open " + path; return false; }
    double dims[2] = {0, 0}, ray_depth = 1, spp = 1, bg[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, right[3] = {1, 0, 0}, up[3] = {0, 1, 0}, fwd[3] = {0, 0, -1};
    double fov_x = 3.14159265358979323846 / 2;
    std::vector<TextPrim> prims;
    TextPrim* cur = nullptr;
    std::string line;
    int line_no = 0;
    auto bad = [&](const std::string& what) { err->code = RT_ERR_FORMAT; err->message = path + ":" + std::to_string(line_no) + ": " + what; return false; };
    while (std::getline(f, line)) {
        ++line_no;
        std::istringstream ss(line);
        std::vector<std::string> tok;
        for (std::string t; ss >> t;) tok.push_back(t);
        if (tok.empty()) continue;
        const std::string& key = tok[0];
        size_t got = 0;
        auto want = [&](size_t n, double* dst) { return numbers(tok, n, dst, &got) && got == n; };
        if (key == "NEW_PRIMITIVE") { prims.emplace_back(); cur = &prims.back(); }
        else if (key == "DIMENSIONS") { if (!want(2, dims)) return bad("DIMENSIONS needs 2 numbers"); }
        else if (key == "RAY_DEPTH") { if (!want(1, &ray_depth)) return bad("RAY_DEPTH needs 1 number"); }
        else if (key == "SAMPLES") { if (!want(1, &spp)) return bad("SAMPLES needs 1 number"); }
        else if (key == "BG_COLOR") { if (!want(3, bg)) return bad("BG_COLOR needs 3 numbers"); }
        else if (key == "CAMERA_POSITION") { if (!want(3, pos)) return bad("CAMERA_POSITION needs 3 numbers"); }
        else if (key == "CAMERA_RIGHT") { if (!want(3, right)) return bad("CAMERA_RIGHT needs 3 numbers"); }
        else if (key == "CAMERA_UP") { if (!want(3, up)) return bad("CAMERA_UP needs 3 numbers"); }
        else if (key == "CAMERA_FORWARD") { if (!want(3, fwd)) return bad("CAMERA_FORWARD needs 3 numbers"); }
        else if (key == "CAMERA_FOV_X") { if (!want(1, &fov_x)) return bad("CAMERA_FOV_X needs 1 number"); }
        else if (!cur) continue;
        else if (key == "PLANE") { cur->kind = RT_SHAPE_PLANE; if (!want(3, cur->shape)) return bad("PLANE needs 3 numbers"); }
        else if (key == "ELLIPSOID") { cur->kind = RT_SHAPE_ELLIPSOID; if (!want(3, cur->shape)) return bad("ELLIPSOID needs 3 numbers"); }
        else if (key == "BOX") { cur->kind = RT_SHAPE_BOX; if (!want(3, cur->shape)) return bad("BOX needs 3 numbers"); }
        else if (key == "TRIANGLE") { cur->kind = RT_SHAPE_TRIANGLE; if (!want(9, cur->shape)) return bad("TRIANGLE needs 9 numbers"); }
        else if (key == "POSITION") { if (!want(3, cur->position)) return bad("POSITION needs 3 numbers"); }
        else if (key == "ROTATION") { if (!want(4, cur->rotation)) return bad("ROTATION needs 4 numbers"); }
        else if (key == "COLOR") { if (!want(3, cur->color)) return bad("COLOR needs 3 numbers"); }
        else if (key == "EMISSION") { if (!want(3, cur->emission)) return bad("EMISSION needs 3 numbers"); }
        else if (key == "METALLIC") cur->metallic = true;
        else if (key == "DIELECTRIC") cur->dielectric = true;
        else if (key == "IOR") { if (!want(1, &cur->ior)) return bad("IOR needs 1 number"); }
    }
    HostScene& h = *out;
    h = HostScene();
    h.width = width > 0 ? width : (int32_t)dims[0];
    h.height = height > 0 ? height : (int32_t)dims[1];
    h.samples = samples > 0 ? samples : (int32_t)spp;
    h.ray_depth = (int32_t)ray_depth;
    if (h.width <= 0 || h.height <= 0) { err->code = RT_ERR_FORMAT; err->message = path + ": no usable DIMENSIONS"; return false; }
    for (int a = 0; a < 3; ++a) { h.bg_color[a] = bg[a]; h.camera_position[a] = pos[a]; h.camera_right[a] = right[a]; h.camera_up[a] = up[a]; h.camera_forward[a] = fwd[a]; }
    h.camera_fov_x = fov_x;
    h.camera_fov_y = 2.0 * std::atan(std::tan(fov_x * 0.5) * (double)h.height / (double)h.width);
    for (const TextPrim& p : prims) {
        if (p.kind < 0) continue;                                    // NEW_PRIMITIVE without a shape line
        double shape[9];
        for (int k = 0; k < 9; ++k) shape[k] = p.shape[k];
        double nrm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (p.kind == RT_SHAPE_PLANE) {
            const double len = std::sqrt(shape[0] * shape[0] + shape[1] * shape[1] + shape[2] * shape[2]);
            for (int a = 0; a < 3; ++a) shape[a] = shape[a] / len;
        } else if (p.kind == RT_SHAPE_TRIANGLE) {
            double e1[3], e2[3];
            for (int a = 0; a < 3; ++a) { e1[a] = shape[3 + a] - shape[a]; e2[a] = shape[6 + a] - shape[a]; }
            double ng[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            const double len = std::sqrt(ng[0] * ng[0] + ng[1] * ng[1] + ng[2] * ng[2]);
            for (int a = 0; a < 3; ++a) { ng[a] = ng[a] / len; nrm[a] = nrm[3 + a] = nrm[6 + a] = ng[a]; }
        }
        const double* q = p.rotation;
        const double qlen = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        h.kind.push_back(p.kind);
        h.tri_v.insert(h.tri_v.end(), shape, shape + 9);
        h.tri_n.insert(h.tri_n.end(), nrm, nrm + 9);
        for (int a = 0; a < 4; ++a) h.rotation.push_back(q[a] / qlen);
        h.position.insert(h.position.end(), p.position, p.position + 3);
        const double mat[5] = {p.color[0], p.color[1], p.color[2], p.metallic ? 1.0 : 0.0, p.metallic ? 0.03 : 1.0};
        h.tri_material.insert(h.tri_material.end(), mat, mat + 5);
        h.tri_emission.insert(h.tri_emission.end(), p.emission, p.emission + 3);
        h.mat_kind.push_back(p.dielectric ? RT_MATERIAL_DIELECTRIC : RT_MATERIAL_PBR);
        h.ior.push_back(p.dielectric ? p.ior : 1.0);
    }
    return true;
}

}  // namespace rtb
