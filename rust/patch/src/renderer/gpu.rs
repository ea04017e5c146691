//! `src/renderer/gpu.rs` (new file of the patch): the GPU drop-in for `step_by_step::ThreadPoolRenderer`
//! (src/renderer/step_by_step.rs:37-121) behind `pub trait Renderer` (src/renderer/mod.rs:47-56).
//! UNCOMPILED (no Rust toolchain in the authoring image); the tested equivalent is
//! `renderer::GpuRenderer` in rs_pathtracing_b200/csrc/host/ray_tracing.cpp.
use std::sync::{Arc, RwLock};

use ray_tracing_b200_sys as sys;

use crate::{
    algebra::Vector3d,
    camera::{ray_caster::ImageParams, Camera},
    renderer::Renderer,
    world::{flat::FlatScene, Scene},
};

fn v(a: &Vector3d) -> sys::rt_vec3 { sys::rt_vec3 { x: a.x, y: a.y, z: a.z } }
fn check(rc: i32) { if rc != sys::RT_OK { panic!("rt_b200: {}", sys::last_error()) } }  // the reference unwraps too

pub struct GpuRenderer { scene: *mut sys::rt_scene, depth: u32, seed: u64, progressive: bool }

impl GpuRenderer {
    /// same constructor convention as ThreadPoolRenderer::new (step_by_step.rs:37); `thread_number` is ignored
    pub fn new(scene: Arc<RwLock<Scene>>, thread_number: u32, depth: u32) -> Self {
        Self::with_devices(scene, thread_number, depth, &[0])
    }
    /// the same renderer over several GPUs of the box: ONE handle (rt_scene_create_multi), frames sharded by
    /// interleaved tiles over all of them, the assembled frame identical to the single-GPU one
    pub fn with_devices(scene: Arc<RwLock<Scene>>, _thread_number: u32, depth: u32, devices: &[i32]) -> Self {
        let s = scene.read().unwrap();
        let f: &FlatScene = s.flat();
        let mut images = Vec::new();
        let desc = f.desc(&mut images);
        let mut h = std::ptr::null_mut();
        // (the library copies the description)
        check(unsafe { sys::rt_scene_create_multi(&desc, devices.len() as i32, devices.as_ptr(), &mut h) });
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
