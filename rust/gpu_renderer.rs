//! `src/renderer/gpu.rs` for the `ray_tracing` crate: the GPU drop-in for `step_by_step::ThreadPoolRenderer`
//! (src/renderer/step_by_step.rs:37-121) behind `pub trait Renderer` (src/renderer/mod.rs:47-56).
//! UNCOMPILED (no Rust toolchain in the authoring image); the tested equivalent is
//! `renderer::GpuRenderer` in rs_pathtracing_b200/csrc/host/ray_tracing.cpp.
use std::sync::{Arc, RwLock};

use ray_tracing_b200_sys as sys;

use crate::{
    algebra::Vector3d,
    camera::{ray_caster::ImageParams, Camera},
    renderer::Renderer,
    world::Scene,
};

/// Flat structure-of-arrays description of `Scene.world`, filled by `Describe::describe` on every shape
/// after `add_random_spheres` (src/world/json_models.rs:44).  Rows 0..2 of
/// `InversableTransform.{direct, inverse}` (src/algebra/transform.rs:16-23) go in verbatim.
#[derive(Default)]
pub struct FlatScene {
    pub kind: Vec<u8>, pub flags: Vec<u8>,
    pub inverse: Vec<f64>, pub direct: Vec<f64>, pub params: Vec<f64>, pub material: Vec<u32>,
    pub materials: Vec<sys::rt_material>, pub textures: Vec<sys::rt_texture>,
    pub images: Vec<(u32, u32, Vec<u8>)>, pub noise: Vec<sys::rt_perlin>,
}
/// implemented by Sphere, Cube, Rectangle, RayMarchingShape (+ each ShapeFunction), every Material and Texture
pub trait Describe { fn describe(&self, flat: &mut FlatScene); }

fn v(a: &Vector3d) -> sys::rt_vec3 { sys::rt_vec3 { x: a.x, y: a.y, z: a.z } }
fn check(rc: i32) { if rc != sys::RT_OK { panic!("rt_b200: {}", sys::last_error()) } }  // the reference unwraps too

pub struct GpuRenderer { scene: *mut sys::rt_scene, depth: u32, seed: u64, progressive: bool }

impl GpuRenderer {
    /// same constructor convention as ThreadPoolRenderer::new (step_by_step.rs:37); `thread_number` is ignored
    pub fn new(scene: Arc<RwLock<Scene>>, _thread_number: u32, depth: u32) -> Self {
        let s = scene.read().unwrap();
        let f: &FlatScene = s.flat();
        let images: Vec<sys::rt_image> = f.images.iter()
            .map(|(w, h, px)| sys::rt_image { width: *w, height: *h, rgba: px.as_ptr() }).collect();
        let desc = sys::rt_scene_desc {
            n_shapes: f.kind.len() as u32, kind: f.kind.as_ptr(), flags: f.flags.as_ptr(),
            inverse: f.inverse.as_ptr(), direct: f.direct.as_ptr(), params: f.params.as_ptr(),
            material: f.material.as_ptr(),
            n_materials: f.materials.len() as u32, materials: f.materials.as_ptr(),
            n_textures: f.textures.len() as u32, textures: f.textures.as_ptr(),
            n_images: images.len() as u32, images: images.as_ptr(),
            n_noise: f.noise.len() as u32, noise: f.noise.as_ptr(),
        };
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::rt_scene_create(&desc, 0, &mut h) });   // copies the description
        GpuRenderer { scene: h, depth, seed: 0, progressive: false }
    }
    /// Progressive accumulation (rt_render_set_accumulate): while on, start_rendering calls with an unchanged
    /// camera ADD their samples to the frame instead of starting over -- what RendererState::render
    /// (main_raylib.rs:204-237) would call with samples_number = 1 per UI frame while the camera is still.
    pub fn set_progressive(&mut self, on: bool) {
        self.progressive = on;
        check(unsafe { sys::rt_render_set_accumulate(self.scene, on as i32) });
    }
}

impl Renderer for GpuRenderer {
    fn start_rendering(&mut self, camera: Arc<RwLock<Camera>>, img: &ImageParams, samples_number: u32) {
        let c = camera.read().unwrap();
        let cam = sys::rt_camera {
            position: v(c.position()), direction: v(c.direction()), up: v(c.up()), right: v(c.rigth()),
            fov_rad: c.fov(), focal_length: c.focal_length(),
        };
        let p = sys::rt_render_params {
            image: sys::rt_image_params { width: img.width, height: img.height },
            samples_number, max_depth: self.depth, seed: self.seed,
            shard_count: 1, shard_index: 0, tile_width: 0, tile_height: 0,
        };
        if !self.progressive { self.seed = self.seed.wrapping_add(1); } // a fresh stream per frame, like thread_rng
                                                                       // (an accumulating frame keeps its key)
        check(unsafe { sys::rt_render_start(self.scene, &cam, &p) });  // returns immediately
    }
    /// non-blocking, partial pixels may arrive before completion (step_by_step.rs:101-121);
    /// needs #[repr(C)] on algebra::Vector3d (3 x f64, src/algebra/mod.rs:23-28)
    fn render_step(&mut self, buffer: &mut Vec<Vector3d>) -> bool {
        let mut done = 0;
        check(unsafe { sys::rt_render_poll(self.scene, buffer.as_mut_ptr() as *mut sys::rt_vec3, &mut done) });
        done != 0
    }
    fn stop_rendering(&mut self) { check(unsafe { sys::rt_render_stop(self.scene) }); }
}

impl Drop for GpuRenderer { fn drop(&mut self) { unsafe { sys::rt_scene_destroy(self.scene) } } }
