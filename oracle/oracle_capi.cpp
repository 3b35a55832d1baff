// oracle_capi.cpp — C binding of the CPU ORACLE (test infrastructure only; see rtc_oracle.hpp).
//
// Exports (a) the same `sg_*` scene-building entry points as the product's host library
// (include/rtc_scene.h) so one scene script builds in both, with sg_camera_render being the CPU
// restatement of Camera::render (camera.rs:76-91); and (b) `orc_*` probes that expose every sub-function
// the reference's unit tests pin.
#include "rtc_oracle.hpp"

#include <chrono>
#include <map>
#include <sstream>

#include "../include/rtc_scene.h"

#include <atomic>
#include <mutex>
#include <thread>

using namespace orc;

namespace orc {
uint64_t g_next_id = 1;
}

static thread_local std::string g_err;
static int fail(const std::string& m) {
    g_err = m;
    return -1;
}

struct sg_ctx {
    std::vector<std::shared_ptr<Pattern>> patterns;
    std::vector<std::shared_ptr<UVPattern>> uvs;
    std::vector<Canvas> canvases;
    std::vector<Material> materials;
    std::vector<std::unique_ptr<Shape>> owned;  // slot is released when a parent takes the shape
    std::vector<Shape*> shapes;                 // handle -> shape (stays valid after adoption)
    std::map<const Shape*, int> handle_of;
    std::vector<World> worlds;
    std::vector<Camera> cameras;
    int threads = 1;

    int add_shape(std::unique_ptr<Shape> s) {
        int h = (int)shapes.size();
        shapes.push_back(s.get());
        handle_of[s.get()] = h;
        owned.push_back(std::move(s));
        return h;
    }
    int handle(const Shape* s) {
        auto it = handle_of.find(s);
        if (it != handle_of.end()) return it->second;
        int h = (int)shapes.size();
        shapes.push_back(const_cast<Shape*>(s));
        owned.push_back(nullptr);
        handle_of[s] = h;
        return h;
    }
    Shape* shape(int h) { return (h >= 0 && h < (int)shapes.size()) ? shapes[h] : nullptr; }
};

static Tuple pt(const float* p) { return point(p[0], p[1], p[2]); }
static Tuple vec(const float* p) { return vector(p[0], p[1], p[2]); }
static Color col(const float* p) { return {p[0], p[1], p[2]}; }
static void put3(float* o, Tuple t) { o[0] = t.x, o[1] = t.y, o[2] = t.z; }
static void put3(float* o, Color c) { o[0] = c.r, o[1] = c.g, o[2] = c.b; }

extern "C" {

const char* sg_last_error(void) { return g_err.c_str(); }

void sg_translation(float x, float y, float z, float out[16]) { translation(x, y, z).to16(out); }
void sg_scaling(float x, float y, float z, float out[16]) { scaling(x, y, z).to16(out); }
void sg_rotation_x(float r, float out[16]) { rotation_x(r).to16(out); }
void sg_rotation_y(float r, float out[16]) { rotation_y(r).to16(out); }
void sg_rotation_z(float r, float out[16]) { rotation_z(r).to16(out); }
void sg_shearing(float xy, float xz, float yx, float yz, float zx, float zy, float out[16]) {
    shearing(xy, xz, yx, yz, zx, zy).to16(out);
}
void sg_view_transform(const float from[3], const float to[3], const float up[3], float out[16]) {
    view_transform(pt(from), pt(to), vec(up)).to16(out);
}
void sg_matmul(const float a[16], const float b[16], float out[16]) { (Matrix::from16(a) * Matrix::from16(b)).to16(out); }
void sg_inverse(const float m[16], float out[16]) { Matrix::from16(m).inverse().to16(out); }
void sg_transpose(const float m[16], float out[16]) { Matrix::from16(m).transpose().to16(out); }
float sg_determinant(const float m[16]) { return Matrix::from16(m).determinant(); }

sg_ctx* sg_create(void) { return new sg_ctx(); }
void sg_destroy(sg_ctx* c) { delete c; }

int sg_pattern_new(sg_ctx* c, int kind, const float a[3], const float b[3]) {
    if (kind < SG_PAT_STRIPES || kind > SG_PAT_TEST) return fail("sg_pattern_new: bad kind");
    auto p = std::make_shared<Pattern>();
    p->kind = kind;
    if (a) p->a = col(a);
    if (b) p->b = col(b);
    p->distance = p->b - p->a;  // gradient.rs:23, sine_2d.rs:22
    c->patterns.push_back(p);
    return (int)c->patterns.size() - 1;
}
int sg_pattern_set_transform(sg_ctx* c, int pattern, const float m[16]) {
    if (pattern < 0 || pattern >= (int)c->patterns.size()) return fail("bad pattern handle");
    c->patterns[pattern]->set_transformation(Matrix::from16(m));
    return 0;
}
int sg_uv_pattern_new(sg_ctx* c, int kind, const float* p, int n) {
    if (kind == SG_UV_CHECKERS) {
        if (n != 8) return fail("UVCheckers needs 8 params");
        auto u = std::make_shared<UVCheckers>();
        u->width = p[0], u->height = p[1], u->a = col(p + 2), u->b = col(p + 5);
        c->uvs.push_back(u);
    } else if (kind == SG_UV_ALIGN_CHECK) {
        if (n != 15) return fail("AlignCheck needs 15 params");
        auto u = std::make_shared<AlignCheck>();
        u->main = col(p), u->ul = col(p + 3), u->ur = col(p + 6), u->bl = col(p + 9), u->br = col(p + 12);
        c->uvs.push_back(u);
    } else {
        return fail("bad uv pattern kind");
    }
    return (int)c->uvs.size() - 1;
}
int sg_canvas_new(sg_ctx* c, int w, int h, const float* rgb) {
    if (w < 0 || h < 0) return fail("bad canvas size");
    Canvas cv((size_t)w, (size_t)h);
    if (rgb)
        for (size_t i = 0; i < cv.data.size(); i++) cv.data[i] = Color{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
    c->canvases.push_back(std::move(cv));
    return (int)c->canvases.size() - 1;
}
int sg_canvas_from_ppm(sg_ctx* c, const char* text, int64_t n) {
    try {
        c->canvases.push_back(canvas_from_ppm(std::string(text, (size_t)n)));
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    return (int)c->canvases.size() - 1;
}
int sg_canvas_size(sg_ctx* c, int cv, int* w, int* h) {
    if (cv < 0 || cv >= (int)c->canvases.size()) return fail("bad canvas handle");
    *w = (int)c->canvases[cv].width, *h = (int)c->canvases[cv].height;
    return 0;
}
int sg_canvas_pixels(sg_ctx* c, int cv, float* out) {
    if (cv < 0 || cv >= (int)c->canvases.size()) return fail("bad canvas handle");
    const Canvas& k = c->canvases[cv];
    for (size_t i = 0; i < k.data.size(); i++) out[3 * i] = k.data[i].r, out[3 * i + 1] = k.data[i].g, out[3 * i + 2] = k.data[i].b;
    return 0;
}
int64_t sg_canvas_to_ppm(sg_ctx* c, int cv, char* out, int64_t capacity) {
    if (cv < 0 || cv >= (int)c->canvases.size()) return fail("bad canvas handle");
    std::string ppm = c->canvases[cv].to_ppm();
    if (out && capacity > 0) memcpy(out, ppm.data(), (size_t)std::min<int64_t>(capacity, (int64_t)ppm.size()));
    return (int64_t)ppm.size();
}
int sg_uv_image_new(sg_ctx* c, int cv) {
    if (cv < 0 || cv >= (int)c->canvases.size()) return fail("bad canvas handle");
    auto u = std::make_shared<UVImage>();
    u->canvas = c->canvases[cv];
    c->uvs.push_back(u);
    return (int)c->uvs.size() - 1;
}
int sg_texture_map_new(sg_ctx* c, int uv, int mapping) {
    if (uv < 0 || uv >= (int)c->uvs.size()) return fail("bad uv handle");
    if (mapping < 0 || mapping > 2) return fail("bad mapping");
    auto p = std::make_shared<Pattern>();
    p->kind = SG_PAT_TEXTURE_MAP;
    p->uv = c->uvs[uv];
    p->mapping = mapping;
    c->patterns.push_back(p);
    return (int)c->patterns.size() - 1;
}
int sg_cubic_map_new(sg_ctx* c, const int uv[6]) {
    auto p = std::make_shared<Pattern>();
    p->kind = SG_PAT_CUBIC_MAP;
    for (int i = 0; i < 6; i++) {
        if (uv[i] < 0 || uv[i] >= (int)c->uvs.size()) return fail("bad uv handle");
        p->faces[i] = c->uvs[uv[i]];
    }
    c->patterns.push_back(p);
    return (int)c->patterns.size() - 1;
}

int sg_material_new(sg_ctx* c, const float p[10], int pattern) {
    Material m;
    m.color = col(p);
    m.ambient = p[3], m.diffuse = p[4], m.specular = p[5], m.shininess = p[6];
    m.reflective = p[7], m.transparency = p[8], m.refractive_index = p[9];
    if (pattern >= 0) {
        if (pattern >= (int)c->patterns.size()) return fail("bad pattern handle");
        m.pattern = c->patterns[pattern];
    }
    c->materials.push_back(m);
    return (int)c->materials.size() - 1;
}

int sg_shape_new(sg_ctx* c, int kind) {
    switch (kind) {
        case SG_SPHERE: return c->add_shape(std::make_unique<Sphere>());
        case SG_PLANE: return c->add_shape(std::make_unique<Plane>());
        case SG_CUBE: return c->add_shape(std::make_unique<Cube>());
        case SG_CYLINDER: return c->add_shape(std::make_unique<Cylinder>());
        case SG_CONE: return c->add_shape(std::make_unique<Cone>());
        case SG_GROUP: return c->add_shape(std::make_unique<GroupShape>());
        case SG_TEST_SHAPE: return c->add_shape(std::make_unique<TestShape>());
    }
    return fail("sg_shape_new: bad kind");
}
int sg_triangle_new(sg_ctx* c, const float p1[3], const float p2[3], const float p3[3]) {
    return c->add_shape(std::make_unique<Triangle>(pt(p1), pt(p2), pt(p3)));
}
int sg_smooth_triangle_new(sg_ctx* c, const float p[9], const float n[9]) {
    return c->add_shape(std::make_unique<SmoothTriangle>(pt(p), pt(p + 3), pt(p + 6), vec(n), vec(n + 3), vec(n + 6)));
}
static std::unique_ptr<Shape> take(sg_ctx* c, int h) {
    if (h < 0 || h >= (int)c->owned.size() || !c->owned[h]) return nullptr;
    return std::move(c->owned[h]);
}
int sg_csg_new(sg_ctx* c, int op, int s1, int s2) {
    if (op < 0 || op > 2) return fail("bad csg op");
    auto a = take(c, s1);
    auto b = take(c, s2);
    if (!a || !b) return fail("sg_csg_new: child missing or already owned by a parent");
    return c->add_shape(std::make_unique<CSG>(op, std::move(a), std::move(b)));
}
int sg_shape_clone(sg_ctx* c, int h) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    return c->add_shape(s->clone());
}
int sg_shape_set_transform(sg_ctx* c, int h, const float m[16]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->set_transformation(Matrix::from16(m));
    return 0;
}
int sg_shape_set_material(sg_ctx* c, int h, int m) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    if (m < 0 || m >= (int)c->materials.size()) return fail("bad material handle");
    s->set_material(c->materials[m]);
    return 0;
}
int sg_shape_set_casts_shadow(sg_ctx* c, int h, int v) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->set_casts_shadow(v != 0);
    return 0;
}
int sg_shape_set_bounds(sg_ctx* c, int h, float lo, float hi, int closed) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    if (auto* cy = dynamic_cast<Cylinder*>(s)) {
        cy->minimum_y = lo, cy->maximum_y = hi, cy->closed = closed != 0;
    } else if (auto* co = dynamic_cast<Cone*>(s)) {
        co->minimum_y = lo, co->maximum_y = hi, co->closed = closed != 0;
    } else {
        return fail("sg_shape_set_bounds: not a cylinder or cone");
    }
    return 0;
}
int sg_group_add_child(sg_ctx* c, int g, int child) {
    auto* grp = dynamic_cast<GroupShape*>(c->shape(g));
    if (!grp) return fail("not a group");
    auto ch = take(c, child);
    if (!ch) return fail("sg_group_add_child: child missing or already owned by a parent");
    grp->add_child(std::move(ch));
    return 0;
}
int sg_shape_divide(sg_ctx* c, int h, int threshold) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->divide((size_t)threshold);
    return 0;
}

// obj_parser.rs:250-264
static void normalize_vertices(std::vector<Tuple>& v) {
    BoundingBox b;
    for (size_t i = 1; i < v.size(); i++) b.add_point(v[i]);
    Tuple span = b.max - b.min;
    float scale = rmax(span.x, rmax(span.y, span.z)) / 2.f;
    for (size_t i = 1; i < v.size(); i++) {
        v[i].x = (v[i].x - (b.min.x + span.x / 2.f)) / scale;
        v[i].y = (v[i].y - (b.min.y + span.y / 2.f)) / scale;
        v[i].z = (v[i].z - (b.min.z + span.z / 2.f)) / scale;
    }
}
// obj_parser.rs:100-216 (+ parse_face :224-247, fan_triangulation :267-293, take_all_as_group :33-55)
int sg_parse_obj(sg_ctx* c, const char* text, int64_t n_bytes) {
    std::vector<Tuple> vertices{point(0, 0, 0)}, normals{point(0, 0, 0)};
    std::vector<std::pair<std::string, std::unique_ptr<GroupShape>>> groups;  // declaration order
    GroupShape* current = nullptr;
    bool normalized = false;
    auto find_group = [&](const std::string& name) -> int {
        for (size_t i = 0; i < groups.size(); i++)
            if (groups[i].first == name) return (int)i;
        return -1;
    };
    auto insert_group = [&](const std::string& name) {
        int i = find_group(name);
        if (i >= 0) {
            groups[i].second = std::make_unique<GroupShape>();  // HashMap::insert replaces
            return groups[i].second.get();
        }
        groups.emplace_back(name, std::make_unique<GroupShape>());
        return groups.back().second.get();
    };
    std::istringstream in(std::string(text, (size_t)n_bytes));
    std::string line;
    int index = 0;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string head;
        if (!(ls >> head)) {
            index++;
            continue;
        }
        if (head == "v" || head == "vn") {
            if (head == "v" && normalized) return fail("vertex after first face at line " + std::to_string(index));
            std::vector<float> co;
            std::string tok;
            while (ls >> tok) {
                char* end = nullptr;
                float f = strtof(tok.c_str(), &end);
                if (end == tok.c_str() || *end) return fail("bad float at line " + std::to_string(index));
                co.push_back(f);
            }
            if (co.size() != 3) return fail("wrong number of coordinates at line " + std::to_string(index));
            if (head == "v")
                vertices.push_back(point(co[0], co[1], co[2]));
            else
                normals.push_back(vector(co[0], co[1], co[2]));
        } else if (head == "f") {
            if (!normalized) {
                normalize_vertices(vertices);
                normalized = true;
            }
            struct Spec {
                size_t vertex;
                bool has_normal;
            };
            std::vector<Spec> specs;
            std::string tok;
            while (ls >> tok) {
                std::vector<std::string> parts;
                size_t start = 0;
                while (true) {
                    size_t slash = tok.find('/', start);
                    parts.push_back(tok.substr(start, slash == std::string::npos ? std::string::npos : slash - start));
                    if (slash == std::string::npos) break;
                    start = slash + 1;
                }
                if (parts[0].empty()) return fail("Missing vertex index");
                for (auto& p : parts)
                    for (char ch : p)
                        if (ch < '0' || ch > '9') return fail("bad face index at line " + std::to_string(index));
                Spec s;
                s.vertex = (size_t)strtoull(parts[0].c_str(), nullptr, 10);
                s.has_normal = parts.size() > 2 && !parts[2].empty();
                specs.push_back(s);
            }
            if (specs.size() < 3) return fail("not enough vertices for a face at line " + std::to_string(index));
            if (!current) current = insert_group("");
            bool smooth = specs[0].has_normal;
            for (size_t i = 1; i + 1 < specs.size(); i++) {
                size_t a = specs[0].vertex, b = specs[i].vertex, d = specs[i + 1].vertex;
                if (a >= vertices.size() || b >= vertices.size() || d >= vertices.size())
                    return fail("vertex index out of range at line " + std::to_string(index));
                std::unique_ptr<Shape> tri;
                if (smooth) {
                    // the reference indexes normals with the VERTEX index (obj_parser.rs:283-285)
                    if (a >= normals.size() || b >= normals.size() || d >= normals.size())
                        return fail("normal index out of range at line " + std::to_string(index));
                    tri = std::make_unique<SmoothTriangle>(vertices[a], vertices[b], vertices[d], normals[a], normals[b], normals[d]);
                } else {
                    tri = std::make_unique<Triangle>(vertices[a], vertices[b], vertices[d]);
                }
                current->add_child(std::move(tri));
            }
        } else if (head == "g") {
            std::string name;
            if (!(ls >> name)) return fail("Missing group name on line " + std::to_string(index));
            current = insert_group(name);
        }
        index++;
    }
    if (groups.empty()) return fail("OBJ has no groups");
    if (groups.size() == 1) return c->add_shape(std::move(groups[0].second));
    auto all = std::make_unique<GroupShape>();
    for (auto& g : groups) all->add_child(std::move(g.second));
    return c->add_shape(std::move(all));
}

int sg_shape_kind(sg_ctx* c, int h) {
    Shape* s = c->shape(h);
    return s ? s->kind : fail("bad shape handle");
}
int sg_shape_get_transform(sg_ctx* c, int h, float out[16]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->t.to16(out);
    return 0;
}
int sg_shape_get_inverse(sg_ctx* c, int h, float out[16]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->t_inverse.to16(out);
    return 0;
}
int sg_shape_get_inverse_transpose(sg_ctx* c, int h, float out[16]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    s->t_inverse_transpose.to16(out);
    return 0;
}
int sg_shape_bounding_box(sg_ctx* c, int h, float mn[3], float mx[3]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    BoundingBox b = s->bounding_box();
    put3(mn, b.min), put3(mx, b.max);
    return 0;
}
int sg_shape_parent_space_bounding_box(sg_ctx* c, int h, float mn[3], float mx[3]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    BoundingBox b = s->parent_space_bounding_box();
    put3(mn, b.min), put3(mx, b.max);
    return 0;
}
int sg_group_child_count(sg_ctx* c, int g) {
    Shape* s = c->shape(g);
    if (auto* grp = dynamic_cast<GroupShape*>(s)) return (int)grp->children.size();
    if (dynamic_cast<CSG*>(s)) return 2;
    return fail("not a group");
}
int sg_group_child(sg_ctx* c, int g, int i) {
    Shape* s = c->shape(g);
    if (auto* grp = dynamic_cast<GroupShape*>(s)) {
        if (i < 0 || i >= (int)grp->children.size()) return fail("child index out of range");
        return c->handle(grp->children[i].get());
    }
    if (auto* csg = dynamic_cast<CSG*>(s)) return c->handle(i == 0 ? csg->s1.get() : csg->s2.get());
    return fail("not a group");
}
int sg_triangle_get(sg_ctx* c, int h, float out[12]) {
    Shape* s = c->shape(h);
    const Triangle* t = dynamic_cast<Triangle*>(s);
    if (auto* st = dynamic_cast<SmoothTriangle*>(s)) t = &st->base;
    if (!t) return fail("not a triangle");
    put3(out, t->p1), put3(out + 3, t->e1), put3(out + 6, t->e2), put3(out + 9, t->normal);
    return 0;
}

int sg_world_new(sg_ctx* c) {
    c->worlds.emplace_back();
    return (int)c->worlds.size() - 1;
}
int sg_world_add_object(sg_ctx* c, int w, int h) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    c->worlds[w].objects.push_back(s);
    return 0;
}
int sg_world_set_point_light(sg_ctx* c, int w, const float pos[3], const float intensity[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    c->worlds[w].light = Light::point_light(pt(pos), col(intensity));
    c->worlds[w].has_light = true;
    return 0;
}
int sg_world_set_rect_light(sg_ctx* c, int w, const float intensity[3], const float corner[3], const float u_vec[3],
                            int u_steps, const float v_vec[3], int v_steps, const float* jitter, int n_jitter,
                            uint64_t seed) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    if (u_steps < 1 || v_steps < 1) return fail("light steps must be >= 1");
    Light l = Light::rectangle(col(intensity), pt(corner), vec(u_vec), u_steps, vec(v_vec), v_steps);
    if (n_jitter > 0) l.jitter.assign(jitter, jitter + n_jitter);
    l.seed = seed;
    c->worlds[w].light = l;
    c->worlds[w].has_light = true;
    return 0;
}
int sg_camera_new(sg_ctx* c, uint32_t w, uint32_t h, float fov, const float m[16]) {
    c->cameras.emplace_back(w, h, fov, Matrix::from16(m));
    return (int)c->cameras.size() - 1;
}

// SURVEY.md Appendix E — frozen algorithmic flop table
static double flops_of(const Counters& k, uint64_t pixels) {
    double f = 0;
    f += 33.0 * k.xforms + 26.0 * k.aabb;
    const double pf[7] = {30, 2, 29, 45, 50, 46, 0};
    for (int i = 0; i < 7; i++) f += pf[i] * k.prim[i];
    f += 36.0 * k.primary + 85.0 * k.shades + 75.0 * k.shades + 41.0 * k.patterns;
    f += 13.0 * k.shadow + 14.0 * k.cells + 20.0 * k.schlick + 25.0 * k.refr_dirs + 9.0 * k.combines;
    f += 3.0 * pixels;
    return f;
}
static void add_counters(Counters& a, const Counters& b) {
    a.primary += b.primary, a.secondary += b.secondary, a.shadow += b.shadow, a.shades += b.shades;
    a.xforms += b.xforms, a.aabb += b.aabb, a.patterns += b.patterns, a.cells += b.cells;
    a.schlick += b.schlick, a.refr_dirs += b.refr_dirs, a.combines += b.combines;
    for (int i = 0; i < 7; i++) a.prim[i] += b.prim[i];
}

void orc_set_threads(sg_ctx* c, int n) { c->threads = n < 1 ? 1 : n; }
int orc_max_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// Camera::render (camera.rs:76-91) restricted to rows y0, y0+ystep, ... < min(y1, height-1).
// y0=0, y1=height, ystep=1 is the reference's loop.  Rows run on `threads` host threads (the reference
// is serial; threads>1 is the courtesy all-core baseline and does not change any pixel).
int orc_camera_render_rows(sg_ctx* c, int cam, int w, int depth, uint32_t y0, uint32_t y1, uint32_t ystep,
                           float* out_rgb, uint8_t* out_u8, sg_stats* stats) {
    if (cam < 0 || cam >= (int)c->cameras.size()) return fail("bad camera handle");
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    const Camera& camera = c->cameras[cam];
    const World& world = c->worlds[w];
    if (!world.has_light) return fail("World light should be set");  // world.rs:66
    if (ystep == 0) return fail("ystep must be >= 1");
    for (const Shape* s : world.objects) s->prefill_caches();
    const uint32_t W = camera.width, H = camera.height;
    if (y1 > H) y1 = H;
    size_t npx = (size_t)W * H;
    if (y0 == 0 && ystep == 1) {
        if (out_rgb) memset(out_rgb, 0, npx * 3 * sizeof(float));  // canvas.rs:19-25
        if (out_u8) memset(out_u8, 0, npx * 3);
    }
    Counters total;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t y_end = (H == 0) ? 0 : std::min(y1, H - 1);  // camera.rs:80
    int64_t nrows = (y_end > y0) ? (int64_t)((y_end - y0 + ystep - 1) / ystep) : 0;
    std::atomic<int64_t> next_row{0};
    std::mutex merge;
    auto worker = [&]() {
        Counters local;
        for (;;) {
            int64_t ri = next_row.fetch_add(1);
            if (ri >= nrows) break;
            uint32_t y = y0 + (uint32_t)ri * ystep;
            for (uint32_t x = 0; x + 1 < W; x++) {  // camera.rs:81
                Ray ray = camera.ray_for_pixel(x, y);
                local.primary++;
                PathCtx ctx;
                ctx.pixel = y * W + x;
                Color col = world.color_at(ray, depth, ctx, &local);
                size_t o = ((size_t)y * W + x) * 3;
                if (out_rgb) out_rgb[o] = col.r, out_rgb[o + 1] = col.g, out_rgb[o + 2] = col.b;
                if (out_u8) out_u8[o] = scale_color(col.r), out_u8[o + 1] = scale_color(col.g), out_u8[o + 2] = scale_color(col.b);
            }
        }
        std::lock_guard<std::mutex> lock(merge);
        add_counters(total, local);
    };
    if (c->threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int i = 0; i < c->threads; i++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->primary_rays = total.primary;
        stats->secondary_rays = total.secondary;
        stats->shadow_rays = total.shadow;
        stats->shades = total.shades;
        stats->flops = flops_of(total, total.primary);
        stats->ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->ms_total = stats->ms;
    }
    return 0;
}
int sg_camera_render(sg_ctx* c, int cam, int w, int depth, float* out_rgb, uint8_t* out_u8, sg_stats* stats) {
    return orc_camera_render_rows(c, cam, w, depth, 0, UINT32_MAX, 1, out_rgb, out_u8, stats);
}

// ------------------------------------------------------------------ probes (unit-test surface)
static int emit(sg_ctx* c, const std::vector<Intersection>& xs, float* out_t, int* out_obj, float* out_uv, int cap) {
    int n = (int)xs.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (out_t) out_t[i] = xs[i].distance;
        if (out_obj) out_obj[i] = c->handle(xs[i].object);
        if (out_uv) out_uv[2 * i] = xs[i].u, out_uv[2 * i + 1] = xs[i].v;
    }
    return n;
}
int orc_shape_intersect(sg_ctx* c, int h, const float o[3], const float d[3], int local, float* out_t, int* out_obj,
                        float* out_uv, int cap) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    std::vector<Intersection> xs;
    Ray r(pt(o), vec(d));
    if (local)
        s->local_intersect(r, xs, nullptr);
    else
        s->intersect(r, xs, nullptr);
    return emit(c, xs, out_t, out_obj, out_uv, cap);
}
int orc_shape_normal_at(sg_ctx* c, int h, const float p[3], int local, float u, float v, float out[3]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    Intersection hit{1.f, s, u, v};
    put3(out, local ? s->local_norm_at(pt(p), hit) : s->normal_at(pt(p), hit));
    return 0;
}
int orc_shape_world_to_object(sg_ctx* c, int h, const float p[3], float out[3]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    put3(out, s->t_inverse * pt(p));
    return 0;
}
int orc_shape_normal_to_world(sg_ctx* c, int h, const float n[3], float out[3]) {
    Shape* s = c->shape(h);
    if (!s) return fail("bad shape handle");
    put3(out, s->normal_to_world(vec(n)));
    return 0;
}
int orc_test_shape_saved_ray(sg_ctx* c, int h, float out[6]) {
    auto* s = dynamic_cast<TestShape*>(c->shape(h));
    if (!s || !s->has_saved_ray) return fail("no saved ray");
    put3(out, s->saved_ray.origin), put3(out + 3, s->saved_ray.direction);
    return 0;
}
int orc_shape_includes(sg_ctx* c, int a, int b) {
    Shape *sa = c->shape(a), *sb = c->shape(b);
    if (!sa || !sb) return fail("bad shape handle");
    return sa->includes(sb) ? 1 : 0;
}
int orc_world_intersect(sg_ctx* c, int w, const float o[3], const float d[3], float* out_t, int* out_obj, int cap) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    std::vector<Intersection> xs;
    c->worlds[w].intersect(Ray(pt(o), vec(d)), xs, nullptr);
    return emit(c, xs, out_t, out_obj, nullptr, cap);
}
// Intersection::hit over a hand-built list (intersection.rs:30-35); returns index or -1
int orc_hit_index(const float* ts, int n) {
    std::vector<Intersection> xs;
    for (int i = 0; i < n; i++) xs.push_back({ts[i], nullptr});
    const Intersection* h = hit(xs);
    return h ? (int)(h - xs.data()) : -1;
}
int orc_color_at(sg_ctx* c, int w, const float o[3], const float d[3], int remaining, float out[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    if (!c->worlds[w].has_light) return fail("World light should be set");
    put3(out, c->worlds[w].color_at(Ray(pt(o), vec(d)), remaining, PathCtx(), nullptr));
    return 0;
}
// out[24] = point, eye, normal, reflect, over, under (3 each), inside, n1, n2, distance, pad, pad
static bool build_comps(sg_ctx* c, const float o[3], const float d[3], int n, const float* ts, const int* shapes,
                        const float* uvs, int hit_index, Ray& r, Comps& comps) {
    std::vector<Intersection> xs;
    for (int i = 0; i < n; i++) {
        Shape* s = c->shape(shapes[i]);
        if (!s) return false;
        if (auto* st = dynamic_cast<SmoothTriangle*>(s)) {
            (void)st;  // direct precompute on a SmoothTriangle keeps the outer object (smooth_triangle.rs:96-108)
        }
        xs.push_back({ts[i], s, uvs ? uvs[2 * i] : 0.f, uvs ? uvs[2 * i + 1] : 0.f});
    }
    if (hit_index < 0 || hit_index >= n) return false;
    r = Ray(pt(o), vec(d));
    comps = precompute_values(r, xs[hit_index], xs);
    return true;
}
int orc_precompute(sg_ctx* c, const float o[3], const float d[3], int n, const float* ts, const int* shapes,
                   const float* uvs, int hit_index, float out[24]) {
    Ray r;
    Comps k;
    if (!build_comps(c, o, d, n, ts, shapes, uvs, hit_index, r, k)) return fail("orc_precompute: bad arguments");
    put3(out, k.point), put3(out + 3, k.eye_vector), put3(out + 6, k.surface_normal), put3(out + 9, k.reflection_vector);
    put3(out + 12, k.over_point), put3(out + 15, k.under_point);
    out[18] = k.inside ? 1.f : 0.f, out[19] = k.n1, out[20] = k.n2, out[21] = k.distance, out[22] = out[23] = 0.f;
    return 0;
}
// what: 0 shade_hit, 1 reflected_color, 2 refracted_color, 3 schlick (out[0])
int orc_comps_eval(sg_ctx* c, int w, const float o[3], const float d[3], int n, const float* ts, const int* shapes,
                   int hit_index, int remaining, int what, float out[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    Ray r;
    Comps k;
    if (!build_comps(c, o, d, n, ts, shapes, nullptr, hit_index, r, k)) return fail("orc_comps_eval: bad arguments");
    const World& world = c->worlds[w];
    switch (what) {
        case 0: put3(out, world.shade_hit(k, remaining, PathCtx(), nullptr)); break;
        case 1: put3(out, world.reflected_color(k, remaining, PathCtx(), nullptr)); break;
        case 2: put3(out, world.refracted_color(k, remaining, PathCtx(), nullptr)); break;
        default: out[0] = out[1] = out[2] = schlick_reflectance(k); break;
    }
    return 0;
}
int orc_is_shadowed(sg_ctx* c, int w, const float light_pos[3], const float p[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    return c->worlds[w].is_shadowed(pt(light_pos), pt(p), nullptr) ? 1 : 0;
}
float orc_intensity_at(sg_ctx* c, int w, const float p[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return NAN;
    return c->worlds[w].light.intensity_at(pt(p), c->worlds[w], PathCtx(), nullptr);
}
// out[10] = u_vec(3) v_vec(3) position(3) cells
int orc_light_info(sg_ctx* c, int w, float out[10]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    const Light& l = c->worlds[w].light;
    put3(out, l.u_vec), put3(out + 3, l.v_vec), put3(out + 6, l.position);
    out[9] = (float)l.cells;
    return 0;
}
// point_on_light consuming the light's own jitter table from `cursor` (rectangle_light.rs:60-66)
int orc_point_on_light(sg_ctx* c, int w, int u, int v, int cursor, float out[3]) {
    if (w < 0 || w >= (int)c->worlds.size()) return fail("bad world handle");
    const Light& l = c->worlds[w].light;
    if (l.jitter.empty()) return fail("light has no jitter table");
    float j1 = l.jitter[cursor % l.jitter.size()], j2 = l.jitter[(cursor + 1) % l.jitter.size()];
    put3(out, l.point_on_light(u, v, j1, j2));
    return 0;
}
// phong_lighting with an explicit PointLight-like (position, intensity) pair (phong_lighting.rs:77-271);
// material < 0 means "the shape's own material"
int orc_phong(sg_ctx* c, int shape, int material, const float light_pos[3], const float light_rgb[3], const float p[3],
              const float eye[3], const float normal[3], float light_intensity, float out[3]) {
    Shape* s = c->shape(shape);
    if (!s) return fail("bad shape handle");
    const Material* m = &s->m;
    if (material >= 0) {
        if (material >= (int)c->materials.size()) return fail("bad material handle");
        m = &c->materials[material];
    }
    Light l = Light::point_light(pt(light_pos), col(light_rgb));
    put3(out, phong_lighting(*s, *m, l, pt(p), vec(eye), vec(normal), light_intensity));
    return 0;
}
int orc_pattern_color_at(sg_ctx* c, int pattern, int shape, const float p[3], float out[3]) {
    if (pattern < 0 || pattern >= (int)c->patterns.size()) return fail("bad pattern handle");
    if (shape < 0) {
        put3(out, c->patterns[pattern]->color_at_world(pt(p)));
    } else {
        Shape* s = c->shape(shape);
        if (!s) return fail("bad shape handle");
        put3(out, c->patterns[pattern]->color_at_object(pt(p), *s));
    }
    return 0;
}
int orc_uv_pattern_color_at(sg_ctx* c, int uv, float u, float v, float out[3]) {
    if (uv < 0 || uv >= (int)c->uvs.size()) return fail("bad uv handle");
    try {
        put3(out, c->uvs[uv]->color_at(u, v));
    } catch (const std::exception& e) {  // UVImage outside [0, 1]: the reference panics (canvas.rs:35)
        return fail(std::string("index out of bounds: ") + e.what());
    }
    return 0;
}
void orc_uv_map(int mapping, const float p[3], float out[2]) {
    if (mapping == 0)
        map_spherical(pt(p), out[0], out[1]);
    else if (mapping == 1)
        map_planar(pt(p), out[0], out[1]);
    else
        map_cylindrical(pt(p), out[0], out[1]);
}
int orc_face_from_point(const float p[3]) { return (int)face_from_point(pt(p)); }
void orc_cube_uv(int face, const float p[3], float out[2]) { cube_uv((Face)face, pt(p), out[0], out[1]); }

int orc_camera_ray(sg_ctx* c, int cam, uint32_t x, uint32_t y, float out[6]) {
    if (cam < 0 || cam >= (int)c->cameras.size()) return fail("bad camera handle");
    Ray r = c->cameras[cam].ray_for_pixel(x, y);
    put3(out, r.origin), put3(out + 3, r.direction);
    return 0;
}
int orc_camera_info(sg_ctx* c, int cam, float out[3]) {
    if (cam < 0 || cam >= (int)c->cameras.size()) return fail("bad camera handle");
    out[0] = c->cameras[cam].pixel_size, out[1] = c->cameras[cam].half_width, out[2] = c->cameras[cam].half_height;
    return 0;
}
int orc_scale_color(float v) { return (int)scale_color(v); }

void orc_bbox_transform(const float mn[3], const float mx[3], const float m[16], float out_min[3], float out_max[3]) {
    BoundingBox b;
    b.min = pt(mn), b.max = pt(mx);
    BoundingBox t = b.transform(Matrix::from16(m));
    put3(out_min, t.min), put3(out_max, t.max);
}
int orc_bbox_intersects(const float mn[3], const float mx[3], const float o[3], const float d[3]) {
    BoundingBox b;
    b.min = pt(mn), b.max = pt(mx);
    return b.intersects(Ray(pt(o), vec(d))) ? 1 : 0;
}
void orc_bbox_split(const float mn[3], const float mx[3], float out[12]) {
    BoundingBox b, l, r;
    b.min = pt(mn), b.max = pt(mx);
    b.split(l, r);
    put3(out, l.min), put3(out + 3, l.max), put3(out + 6, r.min), put3(out + 9, r.max);
}
int orc_bbox_contains_box(const float mn[3], const float mx[3], const float omn[3], const float omx[3]) {
    BoundingBox b, o;
    b.min = pt(mn), b.max = pt(mx), o.min = pt(omn), o.max = pt(omx);
    return b.contains_bounding_box(o) ? 1 : 0;
}
int orc_bbox_contains_point(const float mn[3], const float mx[3], const float p[3]) {
    BoundingBox b;
    b.min = pt(mn), b.max = pt(mx);
    return b.contains_point(pt(p)) ? 1 : 0;
}
void orc_reflect(const float in[3], const float n[3], float out[3]) { put3(out, reflect(vec(in), vec(n))); }
void orc_mat_mul_tuple(const float m[16], const float t[4], float out[4]) {
    Tuple r = Matrix::from16(m) * Tuple{t[0], t[1], t[2], t[3]};
    out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}
int orc_csg_allowed(int op, int hit_s1, int in_s1, int in_s2) { return CSG::intersection_allowed(op, hit_s1, in_s1, in_s2) ? 1 : 0; }
// CSG::filter_intersections over a hand-built sorted list of (t, shape) (csg.rs:271-296); returns count
int orc_csg_filter(sg_ctx* c, int csg, int n, const float* ts, const int* shapes, float* out_t, int cap) {
    auto* g = dynamic_cast<CSG*>(c->shape(csg));
    if (!g) return fail("not a csg");
    std::vector<Intersection> xs, out;
    for (int i = 0; i < n; i++) {
        Shape* s = c->shape(shapes[i]);
        if (!s) return fail("bad shape handle");
        xs.push_back({ts[i], s});
    }
    g->filter_intersections(xs, out);
    for (int i = 0; i < (int)out.size() && i < cap; i++) out_t[i] = out[i].distance;
    return (int)out.size();
}

}  // extern "C"
