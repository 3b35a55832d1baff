//! Source a maintainer adds to the reference's `lib` crate (lib/src/) to get `Camera::render_b200`, the drop-in
//! sibling of `Camera::render` (camera.rs:76-91).  NOT COMPILED HERE (no Rust toolchain in the build image); the
//! same flattening is implemented and tested in C++ in ray_tracer_challenge_b200/csrc/host/rtc_host.hpp
//! (class Flattener) — this file is its line-for-line Rust counterpart.
//!
//! Three additions to the reference are needed because its fields are private:
//!   1. `trait Shape { fn flatten(&self, out: &mut FlatScene, parent: i32) -> i32; }` implemented per shape
//!      (leaves push an RtcPrim; GroupShape / CSG push an RtcNode and recurse in child / (s1, s2) order);
//!   2. `trait Pattern { fn lower(&self, out: &mut FlatScene) -> i32; }` and `trait Light { fn lower(&self, scene) }`;
//!   3. this method on Camera (camera.rs), which can read its own private fields.
use rtc_b200_sys as sys;

pub struct FlatScene {
    pub prims: Vec<sys::RtcPrim>,
    pub nodes: Vec<sys::RtcNode>,
    pub refs: Vec<i32>,
    pub materials: Vec<sys::RtcMaterial>,
    pub patterns: Vec<sys::RtcPattern>,
    pub uvs: Vec<sys::RtcUvPattern>,
    /// One entry per distinct UVImage canvas (uv.rs:346-377): its f32 pixels row-major, kept alive until commit.
    pub texture_pixels: Vec<Vec<f32>>,
    pub texture_sizes: Vec<(u32, u32)>,
}

// UVImage lowering (inside uv.rs, where `canvas` is visible): the canvas travels once as an RtcTexture and the uv
// pattern refers to it by index (params[0]).
impl UVImage {
    fn lower_uv(&self, out: &mut FlatScene) -> i32 {
        let (w, h) = (self.canvas.width, self.canvas.height);
        let mut px = Vec::with_capacity(w * h * 3);
        for y in 0..h {
            for x in 0..w {
                let c = self.canvas.pixel_at(x, y);
                px.extend_from_slice(&[c.r, c.g, c.b]);
            }
        }
        out.texture_pixels.push(px);
        out.texture_sizes.push((w as u32, h as u32));
        let mut params = [0f32; 15];
        params[0] = (out.texture_pixels.len() - 1) as f32;
        out.uvs.push(sys::RtcUvPattern { kind: sys::RTC_UV_IMAGE, params });
        (out.uvs.len() - 1) as i32
    }
}

fn mat16(m: &Matrix) -> [f32; 16] {
    let mut o = [0f32; 16];
    for r in 0..4 {
        for c in 0..4 {
            o[r * 4 + c] = m.data[r][c];
        }
    }
    o
}

// Example leaf implementation (Sphere, sphere.rs): every leaf follows the same shape.
impl Sphere {
    fn flatten_leaf(&self, out: &mut FlatScene, parent: i32) -> i32 {
        let b = self.parent_space_bounding_box(); // shape.rs:162-164
        out.prims.push(sys::RtcPrim {
            type_: sys::RTC_SPHERE,
            material: out.material_index(self.material()),
            casts_shadow: self.casts_shadow() as i32,
            parent,
            inv: mat16(self.transformation_inverse()),
            params: [0.0; 12],
            bbox_min: [b.min.x, b.min.y, b.min.z],
            bbox_max: [b.max.x, b.max.y, b.max.z],
        });
        (out.prims.len() - 1) as i32
    }
}

impl Camera {
    /// Same signature and result as `render` (camera.rs:76): World by value, depth, owned Canvas.
    pub fn render_b200(&self, world: World, reflection_recursion_depth: i16) -> Canvas {
        let mut flat = FlatScene::new();
        for o in &world.objects {
            o.flatten(&mut flat, -1); // depth-first: the tie-break order of world.rs:58
        }
        let (w, h) = (self.width_pixels as usize, self.height_pixels as usize);
        let mut rgb = vec![0f32; w * h * 3];
        unsafe {
            let mut scene = std::ptr::null_mut();
            check(sys::rtc_scene_create(&mut scene));
            check(sys::rtc_set_camera(scene, self.width_pixels, self.height_pixels, self.half_width_world,
                                      self.half_height_world, self.pixel_size, mat16(&self.transform_inverse).as_ptr()));
            check(sys::rtc_set_primitives(scene, flat.prims.len() as u32, flat.prims.as_ptr()));
            check(sys::rtc_set_nodes(scene, flat.nodes.len() as u32, flat.nodes.as_ptr(), flat.refs.len() as u32, flat.refs.as_ptr()));
            check(sys::rtc_set_materials(scene, flat.materials.len() as u32, flat.materials.as_ptr()));
            check(sys::rtc_set_patterns(scene, flat.patterns.len() as u32, flat.patterns.as_ptr(), flat.uvs.len() as u32, flat.uvs.as_ptr()));
            let textures: Vec<sys::RtcTexture> = flat.texture_pixels.iter().zip(&flat.texture_sizes)
                .map(|(px, &(w, h))| sys::RtcTexture { width: w, height: h, rgb: px.as_ptr() }).collect();
            check(sys::rtc_set_textures(scene, textures.len() as u32, textures.as_ptr()));
            world.light.as_ref().expect("World light should be set").lower(scene); // world.rs:66
            check(sys::rtc_scene_commit(scene, 0, std::ptr::null())); // every visible GPU
            let mut stats = sys::RtcStats::default();
            check(sys::rtc_render(scene, reflection_recursion_depth as i32, rgb.as_mut_ptr(), std::ptr::null_mut(), &mut stats));
            sys::rtc_scene_destroy(scene);
        }
        let mut canvas = Canvas::new(w, h);
        for y in 0..h {
            for x in 0..w {
                let i = (y * w + x) * 3;
                canvas.write_pixel(x, y, Color::new(rgb[i], rgb[i + 1], rgb[i + 2]));
            }
        }
        canvas
    }
}

fn check(rc: i32) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::rtc_last_error()) };
        panic!("rtc-b200: {}", msg.to_string_lossy()); // the reference's error style is panic (world.rs:66)
    }
}
