/* rtc_scene.h — C binding of the host-side scene model (World / Camera / Shape / Material / Pattern /
 * Light / Canvas) that mirrors the reference's `lib` crate API one-to-one.
 *
 * The reference (garfieldnate/ray_tracer_challenge) is a Rust crate; its public API for the hot path is
 *   Camera::new / Camera::render            lib/src/camera.rs:23-56, 76-91
 *   World { objects, light }                lib/src/world.rs:18-21
 *   Shape::{set_transformation, set_material, set_casts_shadow}   lib/src/shape/shape.rs:31-48
 *   GroupShape::{add_child, divide}         lib/src/shape/group.rs:39-44, 158-172
 *   CSG::new                                lib/src/shape/csg.rs:27-35
 *   Material::builder                       lib/src/material.rs:19-51
 *   patterns                                lib/src/pattern/ *.rs
 *   PointLight::new / RectangleLight::new   lib/src/light/point_light.rs:12-18, rectangle_light.rs:34-59
 *   transformations                         lib/src/transformations.rs:4-68
 *
 * Rust is not available in this image, so the host side is C++ (ray_tracer_challenge_b200/csrc/host/)
 * and this header is the flat binding that Python (ctypes) and C callers use to drive it.  The very same
 * set of `sg_*` entry points is exported by the CPU oracle (oracle/librtc_oracle.so) so a scene written
 * once can be built in both; the two libraries share NO code.
 *
 * Conventions: every matrix is 16 floats, row-major (m[r*4+c]), f32 like the reference
 * (lib/src/matrix.rs:9-12).  Handles are small non-negative ints scoped to one sg_ctx.  Functions that
 * return int return a handle (>= 0) or a negative error code; sg_last_error() gives the message.
 */
#ifndef RTC_SCENE_H
#define RTC_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sg_ctx sg_ctx;

/* shape kinds (lib/src/shape/ *.rs) */
enum {
    SG_SPHERE = 0,
    SG_PLANE = 1,
    SG_CUBE = 2,
    SG_CYLINDER = 3,
    SG_CONE = 4,
    SG_TRIANGLE = 5,
    SG_SMOOTH_TRIANGLE = 6,
    SG_GROUP = 7,
    SG_CSG = 8,
    SG_TEST_SHAPE = 9 /* shape/test_shape.rs — the reference's test double; accepted by the oracle only */
};

/* CSG operators (lib/src/shape/csg.rs:10-15) */
enum { SG_CSG_UNION = 0, SG_CSG_INTERSECTION = 1, SG_CSG_DIFFERENCE = 2 };

/* pattern kinds (lib/src/pattern/ *.rs); SG_PAT_TEST is pattern.rs:66-88 (needed by world.rs:716-744) */
enum {
    SG_PAT_STRIPES = 0,
    SG_PAT_GRADIENT = 1,
    SG_PAT_RINGS = 2,
    SG_PAT_CHECKERS = 3,
    SG_PAT_SINE2D = 4,
    SG_PAT_TEST = 5,
    SG_PAT_TEXTURE_MAP = 6,
    SG_PAT_CUBIC_MAP = 7
};

/* uv patterns and mappings (lib/src/pattern/uv.rs) */
enum { SG_UV_CHECKERS = 0, SG_UV_ALIGN_CHECK = 1, SG_UV_IMAGE = 2 };
enum { SG_MAP_SPHERICAL = 0, SG_MAP_PLANAR = 1, SG_MAP_CYLINDRICAL = 2 };

const char* sg_last_error(void);

/* ---- transformations.rs:4-68 and matrix.rs:86-103,134-143,201-212 (pure functions) ---- */
void sg_translation(float x, float y, float z, float out[16]);
void sg_scaling(float x, float y, float z, float out[16]);
void sg_rotation_x(float radians, float out[16]);
void sg_rotation_y(float radians, float out[16]);
void sg_rotation_z(float radians, float out[16]);
void sg_shearing(float xy, float xz, float yx, float yz, float zx, float zy, float out[16]);
void sg_view_transform(const float from[3], const float to[3], const float up[3], float out[16]);
void sg_matmul(const float a[16], const float b[16], float out[16]);
void sg_inverse(const float m[16], float out[16]);
void sg_transpose(const float m[16], float out[16]);
float sg_determinant(const float m[16]);

/* ---- context ---- */
sg_ctx* sg_create(void);
void sg_destroy(sg_ctx*);

/* ---- patterns ---- */
int sg_pattern_new(sg_ctx*, int kind, const float a[3], const float b[3]);
int sg_pattern_set_transform(sg_ctx*, int pattern, const float m[16]);
/* UVCheckers: params = {width, height, a.rgb, b.rgb} (8); AlignCheck: {main, ul, ur, bl, br}.rgb (15) */
int sg_uv_pattern_new(sg_ctx*, int kind, const float* params, int n_params);
int sg_texture_map_new(sg_ctx*, int uv_pattern, int mapping);
/* order Front, Back, Left, Right, Up, Down (uv.rs:229-249) */
int sg_cubic_map_new(sg_ctx*, const int uv_patterns[6]);

/* ---- canvases as data (canvas.rs) and image textures (uv.rs:346-377) ----
 * sg_canvas_new: Canvas::new (black) or, with rgb != NULL, width*height*3 f32 pixels row-major.
 * sg_canvas_from_ppm: canvas_from_ppm (canvas.rs:119-182): P3 text, comment / empty lines skipped, triplets may span
 *   lines, values divided by the file's scale.  Errors (negative return, sg_last_error()) carry the reference's
 *   ParseError kind: "IncorrectFormat: ...", "MalformedDimensionHeader: ...", "ParseIntError: ...".
 * sg_canvas_to_ppm: Canvas::to_ppm (canvas.rs:58-96), byte for byte: "P3", "w h", "255", then rows of 8-bit values
 *   (truncating scale_color) wrapped at 70 columns.  Returns the length; writes min(length, capacity) bytes.
 * sg_uv_image_new: UVImage::new(canvas) -> a uv pattern handle for sg_texture_map_new / sg_cubic_map_new. */
int sg_canvas_new(sg_ctx*, int width, int height, const float* rgb);
int sg_canvas_from_ppm(sg_ctx*, const char* text, int64_t n_bytes);
int sg_canvas_size(sg_ctx*, int canvas, int* width, int* height);
int sg_canvas_pixels(sg_ctx*, int canvas, float* out_rgb);
int64_t sg_canvas_to_ppm(sg_ctx*, int canvas, char* out, int64_t capacity);
int sg_uv_image_new(sg_ctx*, int canvas);

/* ---- materials (material.rs:19-51) ----
 * params = {color.r, color.g, color.b, ambient, diffuse, specular, shininess, reflective, transparency,
 *           refractive_index}; pattern = handle or -1 */
int sg_material_new(sg_ctx*, const float params[10], int pattern);

/* ---- shapes ---- */
int sg_shape_new(sg_ctx*, int kind); /* sphere, plane, cube, cylinder, cone, group */
int sg_triangle_new(sg_ctx*, const float p1[3], const float p2[3], const float p3[3]);
int sg_smooth_triangle_new(sg_ctx*, const float p[9], const float n[9]);
int sg_csg_new(sg_ctx*, int op, int s1, int s2);
int sg_shape_clone(sg_ctx*, int shape); /* deep clone with fresh ids (object_id.rs:27-31) */
int sg_shape_set_transform(sg_ctx*, int shape, const float m[16]);
int sg_shape_set_material(sg_ctx*, int shape, int material);
int sg_shape_set_casts_shadow(sg_ctx*, int shape, int casts_shadow);
int sg_shape_set_bounds(sg_ctx*, int shape, float minimum_y, float maximum_y, int closed); /* cylinder / cone */
int sg_group_add_child(sg_ctx*, int group, int child);
int sg_shape_divide(sg_ctx*, int shape, int threshold);
/* parse OBJ text (obj_parser.rs:100-216) and take_all_as_group (obj_parser.rs:33-55); groups are taken in
 * order of first declaration (the reference's HashMap order is unspecified). */
int sg_parse_obj(sg_ctx*, const char* text, int64_t n_bytes);

/* introspection used by the tests */
int sg_shape_kind(sg_ctx*, int shape);
int sg_shape_get_transform(sg_ctx*, int shape, float out[16]);
int sg_shape_get_inverse(sg_ctx*, int shape, float out[16]);
int sg_shape_get_inverse_transpose(sg_ctx*, int shape, float out[16]);
int sg_shape_bounding_box(sg_ctx*, int shape, float out_min[3], float out_max[3]);
int sg_shape_parent_space_bounding_box(sg_ctx*, int shape, float out_min[3], float out_max[3]);
int sg_group_child_count(sg_ctx*, int group);
int sg_group_child(sg_ctx*, int group, int index); /* returns the child's handle */
int sg_triangle_get(sg_ctx*, int shape, float out12[12]); /* p1, e1, e2, normal */

/* ---- world, lights, camera ---- */
int sg_world_new(sg_ctx*);
int sg_world_add_object(sg_ctx*, int world, int shape);
int sg_world_set_point_light(sg_ctx*, int world, const float position[3], const float intensity[3]);
/* RectangleLight::new (rectangle_light.rs:34-59).  u_vec / v_vec are the FULL edges.  jitter: a cyclic
 * table (test/utils.rs:15-24) of n_jitter values, consumed as `for v { for u { j_u, j_v } }`
 * (rectangle_light.rs:60-66,76-88) and restarted at every intensity_at call; n_jitter == 0 selects the
 * counter-based generator with `seed` (the stand-in for thread_rng, rectangle_light.rs:46). */
int sg_world_set_rect_light(sg_ctx*, int world, const float intensity[3], const float corner[3],
                            const float u_vec[3], int u_steps, const float v_vec[3], int v_steps,
                            const float* jitter, int n_jitter, uint64_t seed);
int sg_camera_new(sg_ctx*, uint32_t width, uint32_t height, float field_of_view, const float transform[16]);

/* Ray / work counters shared by the oracle and the device path. */
typedef struct sg_stats {
    uint64_t primary_rays;
    uint64_t secondary_rays; /* reflect + refract color_at calls */
    uint64_t shadow_rays;    /* is_shadowed calls */
    uint64_t shades;         /* shade_hit calls */
    double flops;            /* algorithmic FP32 flops, SURVEY.md Appendix E */
    double ms;               /* device (or CPU) time spent in the render itself */
    double ms_total;         /* including host<->device copies */
} sg_stats;

/* Camera::render (camera.rs:76-91).  out_rgb: width*height*3 f32, row-major, last row/column left
 * black; out_u8 (optional): Canvas::scale_color (canvas.rs:39-43) of the same pixels. */
int sg_camera_render(sg_ctx*, int camera, int world, int depth, float* out_rgb, uint8_t* out_u8,
                     sg_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
