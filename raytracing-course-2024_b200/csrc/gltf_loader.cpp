// gltf_loader.cpp -- glTF 2.0 (.gltf + .bin / data: URIs, or .glb) -> HostScene with the reference's rules.
//
// Replaces gltf::import (main.rs:45) + convert_gltf_to_scene / read_primitives (gltf_to_scene.rs:21-256), minus
// the BVH build (done for the device layout in bvh_builder.cpp).  Behaviour kept on purpose:
//   * JSON numbers become f32 first (the `gltf` crate deserialises f32), then widen to f64;
//   * a TRS node's matrix is T*R*S evaluated in f32 (gltf crate), widened afterwards (gltf_to_scene.rs:107-122);
//   * EVERY node of the document is visited with the identity parent, children are visited again through the
//     recursion (gltf_to_scene.rs:42-52, 245-255) -- a child mesh is therefore instanced twice, like the reference;
//   * only the first primitive of a mesh (:148); indices required (todo!() at :151-153 -> RT_ERR_FORMAT);
//   * positions through the f64 4x4 with a perspective divide (:172-180); normals rotated by the accumulated
//     unit quaternion local*parent, scale ignored (:112-117, 192-194); face normal if NORMAL is absent (:184-200);
//   * camera basis = matrix columns, forward = -Z column, fov_x = aspect*yfov in f32 (:134-143);
//   * roughness >= 0.03 (:221), emission = emissiveFactor * KHR_materials_emissive_strength (:223-231),
//     ray_depth 6 and black background (:65, :73).
// Compile with -ffp-contract=off: the reference's arithmetic is unfused.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/rt_api.h"
#include "host_scene.h"
#include "json_min.h"

namespace rtb {
namespace {

struct Mat4 { double m[4][4]; };  // row-major: m[row][col]
struct Quat { double w, i, j, k; };

Mat4 identity4() { Mat4 r; for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) r.m[a][b] = a == b ? 1.0 : 0.0; return r; }
Mat4 mul4(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = a.m[i][0] * b.m[0][j];
            s += a.m[i][1] * b.m[1][j];
            s += a.m[i][2] * b.m[2][j];
            s += a.m[i][3] * b.m[3][j];
            r.m[i][j] = s;
        }
    return r;
}
// nalgebra Quaternion * Quaternion (gltf_to_scene.rs:112-117)
Quat qmul(const Quat& a, const Quat& b) {
    Quat r;
    r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
    r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
    r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
    r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
    return r;
}
Quat qnormalize(const Quat& q) {  // Unit::new_normalize, coords stored [i, j, k, w]
    double n = std::sqrt(q.i * q.i + q.j * q.j + q.k * q.k + q.w * q.w);
    Quat r = {q.w / n, q.i / n, q.j / n, q.k / n};
    return r;
}
// UnitQuaternion::transform_vector: t = (q.v x v) * 2; (t*w + q.v x t) + v
void qrotate(const Quat& q, const double v[3], double out[3]) {
    double t[3] = {(q.j * v[2] - q.k * v[1]) * 2.0, (q.k * v[0] - q.i * v[2]) * 2.0, (q.i * v[1] - q.j * v[0]) * 2.0};
    double c[3] = {q.j * t[2] - q.k * t[1], q.k * t[0] - q.i * t[2], q.i * t[1] - q.j * t[0]};
    for (int a = 0; a < 3; ++a) out[a] = (t[a] * q.w + c[a]) + v[a];
}

float f32_of(const rtjson::Value& v) { return (float)v.num; }

// gltf crate Transform::Decomposed{t, r, s}.matrix(): T * R * S in f32.
Mat4 trs_matrix_f32(const float t[3], const float r[4], const float s[3]) {
    float x = r[0], y = r[1], z = r[2], w = r[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z;
    float yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    float sx2 = x2 * w, sy2 = y2 * w, sz2 = z2 * w;
    float one = 1.0f;
    float c0[3] = {one - yy2 - zz2, xy2 + sz2, xz2 - sy2};
    float c1[3] = {xy2 - sz2, one - xx2 - zz2, yz2 + sx2};
    float c2[3] = {xz2 + sy2, yz2 - sx2, one - xx2 - yy2};
    Mat4 m;
    std::memset(&m, 0, sizeof(m));
    for (int a = 0; a < 3; ++a) {
        m.m[a][0] = (double)(float)(c0[a] * s[0]);
        m.m[a][1] = (double)(float)(c1[a] * s[1]);
        m.m[a][2] = (double)(float)(c2[a] * s[2]);
        m.m[a][3] = (double)t[a];
    }
    m.m[3][3] = 1.0;
    return m;
}

bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

bool base64_decode(const std::string& in, size_t start, std::string* out) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        std::memset(T, -1, sizeof(T));
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) T[(unsigned char)a[i]] = (int8_t)i;
        init = true;
    }
    unsigned acc = 0; int bits = 0;
    for (size_t p = start; p < in.size(); ++p) {
        unsigned char c = (unsigned char)in[p];
        if (c == '=') break;
        if (T[c] < 0) continue;
        acc = (acc << 6) | (unsigned)T[c]; bits += 6;
        if (bits >= 8) { bits -= 8; out->push_back((char)((acc >> bits) & 0xFF)); }
    }
    return true;
}

struct Doc {
    rtjson::ValuePtr root;
    std::vector<std::string> buffers;
    std::string dir;
};

struct AccessorView { const unsigned char* base; size_t stride; size_t count; int component_type; int ncomp; };

int ncomp_of(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    if (t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    return 0;
}
int comp_size(int ct) { switch (ct) { case 5120: case 5121: return 1; case 5122: case 5123: return 2; case 5125: case 5126: return 4; default: return 0; } }

// Every index / size of the document is taken through Value::index (non-negative integral < 2^53, else the file is rejected -- the
// gltf crate's deserialiser rejects such files too) and every range check is written without overflow.
bool accessor_view(const Doc& d, long long idx, AccessorView* v, std::string* why) {
    const rtjson::Value* accs = d.root->get("accessors");
    if (!accs || idx < 0 || (size_t)idx >= accs->size()) { *why = "accessor index out of range"; return false; }
    const rtjson::Value& acc = accs->at((size_t)idx);
    if (acc.has("sparse")) { *why = "sparse accessors are not supported"; return false; }
    const long long bvi = acc.index("bufferView");
    const rtjson::Value* bvs = d.root->get("bufferViews");
    if (!bvs || bvi < 0 || (size_t)bvi >= bvs->size()) { *why = "accessor without a valid bufferView"; return false; }
    const rtjson::Value& bv = bvs->at((size_t)bvi);
    const long long bi = bv.index("buffer");
    if (bi < 0 || (size_t)bi >= d.buffers.size()) { *why = "bufferView.buffer out of range"; return false; }
    const long long ct = acc.index("componentType");
    v->component_type = ct >= 0 && ct < 65536 ? (int)ct : 0;
    v->ncomp = ncomp_of(acc.string("type", ""));
    const int cs = comp_size(v->component_type);
    if (cs == 0 || v->ncomp == 0) { *why = "unknown accessor type"; return false; }
    const long long bv_off = bv.has("byteOffset") ? bv.index("byteOffset") : 0, acc_off = acc.has("byteOffset") ? acc.index("byteOffset") : 0;
    const long long bv_stride = bv.has("byteStride") ? bv.index("byteStride") : 0, count = acc.index("count");
    const long long bv_len = bv.has("byteLength") ? bv.index("byteLength") : -1;
    if (bv_off < 0 || acc_off < 0 || bv_stride < 0 || count < 0 || (bv.has("byteLength") && bv_len < 0)) { *why = "accessor / bufferView with a negative or non-integral size"; return false; }
    const std::string& buf = d.buffers[(size_t)bi];
    const size_t elem = (size_t)cs * (size_t)v->ncomp;
    const size_t stride = bv_stride == 0 ? elem : (size_t)bv_stride;
    if (stride < elem && bv_stride != 0) { *why = "bufferView.byteStride smaller than the element"; return false; }
    // the view [bv_off, bv_off + bv_len) must lie inside the buffer, the accessor inside the view
    if ((size_t)bv_off > buf.size()) { *why = "bufferView starts past its buffer"; return false; }
    size_t avail = buf.size() - (size_t)bv_off;
    if (bv_len >= 0) { if ((size_t)bv_len > avail) { *why = "bufferView exceeds its buffer"; return false; } avail = (size_t)bv_len; }
    if ((size_t)acc_off > avail) { *why = "accessor starts past its bufferView"; return false; }
    avail -= (size_t)acc_off;
    if (count > 0 && (avail < elem || (size_t)(count - 1) > (avail - elem) / stride)) { *why = "accessor exceeds its bufferView"; return false; }
    v->count = (size_t)count;
    v->stride = stride;
    v->base = (const unsigned char*)buf.data() + (size_t)bv_off + (size_t)acc_off;
    return true;
}

struct Builder {
    const Doc& d;
    HostScene* sc;
    LoadError* err;
    int depth_guard = 0;

    bool fail(int code, const std::string& msg) { if (err->code == 0) { err->code = code; err->message = msg; } return false; }

    // gltf_to_scene.rs:215-231 with the glTF defaults of the gltf crate's material getters
    void material_of(const rtjson::Value& prim, double mat[5], double emission[3]) {
        const rtjson::Value* m = nullptr;
        const rtjson::Value* mats = d.root->get("materials");
        long long mi = prim.index("material");
        if (mats && mi >= 0 && (size_t)mi < mats->size()) m = &mats->at((size_t)mi);
        float base[3] = {1.0f, 1.0f, 1.0f}, metallic = 1.0f, rough = 1.0f, ef[3] = {0.0f, 0.0f, 0.0f};
        double strength = 1.0;
        if (m) {
            if (const rtjson::Value* pbr = m->get("pbrMetallicRoughness")) {
                if (const rtjson::Value* b = pbr->get("baseColorFactor")) for (int a = 0; a < 3 && (size_t)a < b->size(); ++a) base[a] = f32_of(b->at((size_t)a));
                if (const rtjson::Value* v = pbr->get("metallicFactor")) metallic = f32_of(*v);
                if (const rtjson::Value* v = pbr->get("roughnessFactor")) rough = f32_of(*v);
            }
            if (const rtjson::Value* e = m->get("emissiveFactor")) for (int a = 0; a < 3 && (size_t)a < e->size(); ++a) ef[a] = f32_of(e->at((size_t)a));
            if (const rtjson::Value* ex = m->get("extensions"))
                if (const rtjson::Value* es = ex->get("KHR_materials_emissive_strength")) {
                    float s = 1.0f;
                    if (const rtjson::Value* v = es->get("emissiveStrength")) s = f32_of(*v);
                    strength = (double)s;
                }
        }
        mat[0] = (double)base[0]; mat[1] = (double)base[1]; mat[2] = (double)base[2];
        mat[3] = (double)metallic;
        mat[4] = std::fmax((double)rough, 0.03);
        for (int a = 0; a < 3; ++a) emission[a] = (double)ef[a] * strength;
    }

    bool node_transform(const rtjson::Value& node, Mat4* local, Quat* rot) {
        if (const rtjson::Value* mv = node.get("matrix")) {
            if (mv->size() != 16) return fail(RT_ERR_FORMAT, "node.matrix must have 16 entries");
            float cm[16];
            for (int a = 0; a < 16; ++a) cm[a] = f32_of(mv->at((size_t)a));
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) local->m[r][c] = (double)cm[c * 4 + r];
            // gltf crate 1.x `Transform::Matrix { matrix }.decomposed()` (scene/mod.rs; restated from its published source -- the crate
            // is not vendored): ALL in f32.  i = upper 3x3 by columns; sx = |i.x|, sy = |i.y|, sz = signum(det i) * |i.z| (a mirror goes to
            // the z scale only); columns multiplied by 1/s; rotation = Quaternion::from_matrix(i) with the cgmath branch order.
            float x[3] = {cm[0], cm[1], cm[2]}, y[3] = {cm[4], cm[5], cm[6]}, z[3] = {cm[8], cm[9], cm[10]};
            auto mag = [](const float* v) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); };
            const float det = (x[0] * (y[1] * z[2] - z[1] * y[2]) - y[0] * (x[1] * z[2] - z[1] * x[2])) + z[0] * (x[1] * y[2] - y[1] * x[2]);
            const float sx = mag(x), sy = mag(y), sz = std::copysign(1.0f, det) * mag(z);
            const float ix = 1.0f / sx, iy = 1.0f / sy, iz = 1.0f / sz;
            for (int a = 0; a < 3; ++a) { x[a] = x[a] * ix; y[a] = y[a] * iy; z[a] = z[a] * iz; }
            const float trace = (x[0] + y[1]) + z[2];
            float q[4];
            if (trace >= 0.0f) {
                float s = std::sqrt(1.0f + trace);
                q[3] = 0.5f * s; s = 0.5f / s;
                q[0] = (y[2] - z[1]) * s; q[1] = (z[0] - x[2]) * s; q[2] = (x[1] - y[0]) * s;
            } else if (x[0] > y[1] && x[0] > z[2]) {
                float s = std::sqrt(((x[0] - y[1]) - z[2]) + 1.0f);
                q[0] = 0.5f * s; s = 0.5f / s;
                q[1] = (y[0] + x[1]) * s; q[2] = (x[2] + z[0]) * s; q[3] = (y[2] - z[1]) * s;
            } else if (y[1] > z[2]) {
                float s = std::sqrt(((y[1] - x[0]) - z[2]) + 1.0f);
                q[1] = 0.5f * s; s = 0.5f / s;
                q[2] = (z[1] + y[2]) * s; q[0] = (y[0] + x[1]) * s; q[3] = (z[0] - x[2]) * s;
            } else {
                float s = std::sqrt(((z[2] - x[0]) - y[1]) + 1.0f);
                q[2] = 0.5f * s; s = 0.5f / s;
                q[0] = (x[2] + z[0]) * s; q[1] = (z[1] + y[2]) * s; q[3] = (x[1] - y[0]) * s;
            }
            rot->i = (double)q[0]; rot->j = (double)q[1]; rot->k = (double)q[2]; rot->w = (double)q[3];
            return true;
        }
        float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        if (const rtjson::Value* v = node.get("translation")) for (int a = 0; a < 3 && (size_t)a < v->size(); ++a) t[a] = f32_of(v->at((size_t)a));
        if (const rtjson::Value* v = node.get("rotation")) for (int a = 0; a < 4 && (size_t)a < v->size(); ++a) r[a] = f32_of(v->at((size_t)a));
        if (const rtjson::Value* v = node.get("scale")) for (int a = 0; a < 3 && (size_t)a < v->size(); ++a) s[a] = f32_of(v->at((size_t)a));
        *local = trs_matrix_f32(t, r, s);
        rot->i = (double)r[0]; rot->j = (double)r[1]; rot->k = (double)r[2]; rot->w = (double)r[3];
        return true;
    }

    // gltf_to_scene.rs:97-256 read_primitives
    bool read_primitives(size_t node_idx, const Mat4& transformation, const Quat& rotation) {
        if (++depth_guard > 256) return fail(RT_ERR_FORMAT, "node hierarchy too deep (cycle?)");
        const rtjson::Value* nodes = d.root->get("nodes");
        if (!nodes || node_idx >= nodes->size()) return fail(RT_ERR_FORMAT, "node index out of range");
        const rtjson::Value& node = nodes->at(node_idx);
        Mat4 local; Quat r;
        if (!node_transform(node, &local, &r)) return false;
        Quat current_rotation = qnormalize(qmul(r, rotation));
        Mat4 m = mul4(transformation, local);

        if (node.has("camera")) {
            const rtjson::Value* cams = d.root->get("cameras");
            long long ci = node.index("camera");
            if (!cams || ci < 0 || (size_t)ci >= cams->size()) return fail(RT_ERR_FORMAT, "camera index out of range");
            const rtjson::Value& cam = cams->at((size_t)ci);
            const rtjson::Value* persp = cam.get("perspective");
            if (cam.string("type", "") != "perspective" || !persp)
                return fail(RT_ERR_FORMAT, "non-perspective camera (the reference hits todo!() at gltf_to_scene.rs:131-133)");
            float yfov = (float)persp->number("yfov", 0.0);
            float aspect = persp->has("aspectRatio") ? (float)persp->number("aspectRatio", 1.0) : 1.0f;
            sc->camera_fov_y = (double)yfov;
            sc->camera_fov_x = (double)(float)(aspect * yfov);
            for (int a = 0; a < 3; ++a) {
                sc->camera_position[a] = m.m[a][3] / m.m[3][3];
                sc->camera_forward[a] = -m.m[a][2];
                sc->camera_up[a] = m.m[a][1];
                sc->camera_right[a] = m.m[a][0];
            }
        }

        if (node.has("mesh")) {
            const rtjson::Value* meshes = d.root->get("meshes");
            long long mi = node.index("mesh");
            if (!meshes || mi < 0 || (size_t)mi >= meshes->size()) return fail(RT_ERR_FORMAT, "mesh index out of range");
            const rtjson::Value* prims = meshes->at((size_t)mi).get("primitives");
            if (!prims || prims->size() == 0) return fail(RT_ERR_FORMAT, "mesh without primitives");
            const rtjson::Value& prim = prims->at(0);                                   // :148
            if (!prim.has("indices"))
                return fail(RT_ERR_FORMAT, "non-indexed mesh (the reference hits todo!() at gltf_to_scene.rs:151-153)");
            const rtjson::Value* attrs = prim.get("attributes");
            if (!attrs || !attrs->has("POSITION")) return fail(RT_ERR_FORMAT, "Missing positions!");
            std::string why;
            AccessorView iv, pv, nv;
            if (!accessor_view(d, prim.index("indices"), &iv, &why)) return fail(RT_ERR_FORMAT, "indices: " + why);
            if (!accessor_view(d, attrs->index("POSITION"), &pv, &why)) return fail(RT_ERR_FORMAT, "POSITION: " + why);
            if (pv.component_type != 5126 || pv.ncomp != 3) return fail(RT_ERR_FORMAT, "POSITION must be float VEC3");
            bool has_normals = attrs->has("NORMAL");
            if (has_normals) {
                if (!accessor_view(d, attrs->index("NORMAL"), &nv, &why)) return fail(RT_ERR_FORMAT, "NORMAL: " + why);
                if (nv.component_type != 5126 || nv.ncomp != 3) return fail(RT_ERR_FORMAT, "NORMAL must be float VEC3");
            }
            if (iv.component_type != 5121 && iv.component_type != 5123 && iv.component_type != 5125) return fail(RT_ERR_FORMAT, "indices must be u8/u16/u32");

            // world-space positions and rotated normals per vertex (the reference recomputes them per corner;
            // the arithmetic per vertex is identical)
            std::vector<double> wpos(pv.count * 3), wnrm;
            for (size_t v = 0; v < pv.count; ++v) {
                float p[3];
                std::memcpy(p, pv.base + v * pv.stride, 12);
                double x = (double)p[0], y = (double)p[1], z = (double)p[2], o[4];
                for (int a = 0; a < 4; ++a) {                                            // column-axpy order
                    double s = m.m[a][0] * x;
                    s = s + m.m[a][1] * y;
                    s = s + m.m[a][2] * z;
                    s = s + m.m[a][3];
                    o[a] = s;
                }
                wpos[v * 3 + 0] = o[0] / o[3]; wpos[v * 3 + 1] = o[1] / o[3]; wpos[v * 3 + 2] = o[2] / o[3];
            }
            if (has_normals) {
                wnrm.resize(nv.count * 3);
                for (size_t v = 0; v < nv.count; ++v) {
                    float p[3];
                    std::memcpy(p, nv.base + v * nv.stride, 12);
                    double in[3] = {(double)p[0], (double)p[1], (double)p[2]};
                    qrotate(current_rotation, in, &wnrm[v * 3]);
                }
            }
            double mat[5], emission[3];
            material_of(prim, mat, emission);
            size_t ntri = iv.count / 3;                                                  // chunks_exact(3)
            for (size_t t = 0; t < ntri; ++t) {
                size_t idx[3];
                for (int c = 0; c < 3; ++c) {
                    const unsigned char* p = iv.base + (t * 3 + (size_t)c) * iv.stride;
                    if (iv.component_type == 5121) idx[c] = *p;
                    else if (iv.component_type == 5123) { uint16_t u; std::memcpy(&u, p, 2); idx[c] = u; }
                    else { uint32_t u; std::memcpy(&u, p, 4); idx[c] = u; }
                    if (idx[c] >= pv.count || (has_normals && idx[c] >= nv.count)) return fail(RT_ERR_FORMAT, "vertex index out of range");
                }
                const double* a = &wpos[idx[0] * 3]; const double* b = &wpos[idx[1] * 3]; const double* c = &wpos[idx[2] * 3];
                for (int k = 0; k < 3; ++k) sc->tri_v.push_back(a[k]);
                for (int k = 0; k < 3; ++k) sc->tri_v.push_back(b[k]);
                for (int k = 0; k < 3; ++k) sc->tri_v.push_back(c[k]);
                if (has_normals) {
                    for (int c2 = 0; c2 < 3; ++c2) for (int k = 0; k < 3; ++k) sc->tri_n.push_back(wnrm[idx[c2] * 3 + (size_t)k]);
                } else {                                                                 // :184 default_normal
                    double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
                    double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                    double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                    for (int c2 = 0; c2 < 3; ++c2) for (int k = 0; k < 3; ++k) sc->tri_n.push_back(n[k] / len);
                }
                for (int k = 0; k < 5; ++k) sc->tri_material.push_back(mat[k]);
                for (int k = 0; k < 3; ++k) sc->tri_emission.push_back(emission[k]);
            }
        }
        if (const rtjson::Value* ch = node.get("children"))                               // :245-255
            for (size_t k = 0; k < ch->size(); ++k)
            {
                const long long ci = ch->at(k).as_index();
                if (ci < 0) return fail(RT_ERR_FORMAT, "node.children entry is not a non-negative integer");
                if (!read_primitives((size_t)ci, m, current_rotation)) return false;
            }
        --depth_guard;
        return true;
    }
};

}  // namespace

bool load_gltf_scene(const std::string& path, int32_t width, int32_t height, int32_t samples, HostScene* out, LoadError* err) {
    err->code = 0; err->message.clear();
    std::string text;
    if (!read_file(path, &text)) { err->code = RT_ERR_IO; err->message = "cannot read " + path; return false; }
    Doc d;
    size_t slash = path.find_last_of('/');
    d.dir = slash == std::string::npos ? std::string(".") : path.substr(0, slash);
    std::string glb_bin;
    bool is_glb = text.size() >= 12 && std::memcmp(text.data(), "glTF", 4) == 0;
    try {
        if (is_glb) {
            size_t p = 12; std::string json_chunk;
            while (p + 8 <= text.size()) {
                uint32_t len, type;
                std::memcpy(&len, text.data() + p, 4); std::memcpy(&type, text.data() + p + 4, 4);
                p += 8;
                if (p + len > text.size()) break;
                if (type == 0x4E4F534Au) json_chunk = text.substr(p, len);
                else if (type == 0x004E4942u && glb_bin.empty()) glb_bin = text.substr(p, len);
                p += len;
            }
            d.root = rtjson::parse(json_chunk);
        } else {
            d.root = rtjson::parse(text);
        }
    } catch (const std::exception& e) { err->code = RT_ERR_FORMAT; err->message = std::string("glTF JSON: ") + e.what(); return false; }
    if (!d.root->is_object()) { err->code = RT_ERR_FORMAT; err->message = "glTF root is not an object"; return false; }

    if (const rtjson::Value* bufs = d.root->get("buffers")) {
        for (size_t b = 0; b < bufs->size(); ++b) {
            const rtjson::Value& bv = bufs->at(b);
            std::string uri = bv.string("uri", "");
            std::string data;
            if (uri.empty()) { data = glb_bin; }
            else if (uri.compare(0, 5, "data:") == 0) {
                size_t comma = uri.find(',');
                if (comma == std::string::npos) { err->code = RT_ERR_FORMAT; err->message = "malformed data: URI"; return false; }
                base64_decode(uri, comma + 1, &data);
            } else if (!read_file(d.dir + "/" + uri, &data)) { err->code = RT_ERR_IO; err->message = "cannot read buffer " + d.dir + "/" + uri; return false; }
            d.buffers.push_back(std::move(data));
        }
    }

    HostScene sc;
    sc.width = width; sc.height = height; sc.samples = samples; sc.ray_depth = 6;          // gltf_to_scene.rs:73
    Builder b{d, &sc, err};
    Mat4 ident = identity4();
    Quat qi = {1.0, 0.0, 0.0, 0.0};
    try {
        if (const rtjson::Value* nodes = d.root->get("nodes"))
            for (size_t n = 0; n < nodes->size(); ++n) {                                      // :42-52: ALL nodes
                b.depth_guard = 0;
                if (!b.read_primitives(n, ident, qi)) return false;
            }
    } catch (const std::exception& e) { err->code = RT_ERR_FORMAT; err->message = std::string("glTF: ") + e.what(); return false; }
    *out = std::move(sc);
    return true;
}

}  // namespace rtb
