//! `src/world/flat.rs` (new file of the patch): the flat structure-of-arrays description of a `Scene`'s shape list
//! that `include/rt_b200.h` (`rt_scene_desc`) takes, and the registries that turn `MaterialPtr`s and nested
//! `Box<dyn Texture>`s into table indices.
//!
//! Filled by `Shape::describe` on every shape of the list IN LIST ORDER, after `add_random_spheres`
//! (src/world/json_models.rs:44) and BEFORE `Scene::new` moves the list into its `BvhNode`
//! (src/world/mod.rs:35): the kernels reproduce `ShapeCollection::ray_intersect`'s "later shape wins ties"
//! (src/world/shapes/mod.rs:587-596), which needs the original indices.
//!
//! UNCOMPILED: the authoring image has no Rust toolchain.  The tested equivalent is `FlatScene` / `Builder` in
//! rs_pathtracing_b200/csrc/host/ray_tracing.cpp; tests/test_rust_patch.py checks that this file and the patch
//! produce every RT_SHAPE_* / RT_SURF_* / RT_MAT_* / RT_TEX_* tag and the same params[8] slots as that mirror.
use std::{collections::HashMap, sync::Arc};

use ray_tracing_b200_sys as sys;

use super::{material::MaterialPtr, shapes::Shape};
use crate::algebra::transform::InversableTransform;

/// params[8] of a shape row (RT_SHAPE_PARAMS):
///   Sphere, Cube       all zero
///   Rectangle          [x0, y0, x1, y1, 0, 0, 0, 0]
///   Torus              [radius, tube_radius, 0, 0, 0, 0, 0, 0]
///   RayMarchingShape   [surface kind (RT_SURF_*), step, depth, a, b, c, d, sphere_radius]
///                      (Heart: a..sphere_radius = 0, its bound is the fixed ellipsoid of Heart::new;
///                       Sine / Star: a; DupinCyclide: a, b, c, d; Hunt / Cushion: only sphere_radius)
#[derive(Default)]
pub struct FlatScene {
    pub kind: Vec<u8>,
    pub flags: Vec<u8>,
    pub inverse: Vec<f64>,   // [n][12]: rows 0..2 of InversableTransform.inverse (src/algebra/transform.rs:16-23)
    pub direct: Vec<f64>,    // [n][12]: rows 0..2 of InversableTransform.direct
    pub params: Vec<f64>,    // [n][8]
    pub material: Vec<u32>,
    pub materials: Vec<sys::rt_material>,
    pub textures: Vec<sys::rt_texture>,
    pub images: Vec<(u32, u32, Vec<u8>)>,   // (width, height, RGBA8 texels)
    pub noise: Vec<sys::rt_perlin>,
    seen_materials: HashMap<*const (), u32>,   // Arc pointer identity -> index into `materials`
}

impl FlatScene {
    /// the whole list, in order
    pub fn from_shapes(shapes: &[Box<dyn Shape>]) -> Self {
        let mut flat = FlatScene::default();
        for shape in shapes {
            shape.describe(&mut flat);
        }
        flat
    }

    /// one shape row; called by the `Shape::describe` impls
    pub fn push_shape(&mut self, kind: u8, flags: u8, transform: &InversableTransform, params: [f64; 8],
                      material: &MaterialPtr) {
        self.kind.push(kind);
        self.flags.push(flags);
        for row in 0..3 {
            self.inverse.extend_from_slice(&transform.inverse.0[row]);
            self.direct.extend_from_slice(&transform.direct.0[row]);
        }
        self.params.extend_from_slice(&params);
        let index = self.material_index(material);
        self.material.push(index);
    }

    /// index of a material in the table: the same `Arc` (a named material shared by several shapes, or one of
    /// add_random_spheres' per-sphere materials) is described once
    pub fn material_index(&mut self, material: &MaterialPtr) -> u32 {
        let key = Arc::as_ptr(material) as *const ();
        if let Some(&index) = self.seen_materials.get(&key) {
            return index;
        }
        let row = material.describe(self);   // pushes the material's texture tree first
        self.materials.push(row);
        let index = self.materials.len() as u32 - 1;
        self.seen_materials.insert(key, index);
        index
    }

    /// a texture row; children (CheckerTexture / UVChecker: odd, even) are pushed by the caller first
    pub fn push_texture(&mut self, kind: u32, color: [f64; 3], odd: u32, even: u32, image: u32) -> u32 {
        self.textures.push(sys::rt_texture {
            kind, odd, even, image,
            color: sys::rt_vec3 { x: color[0], y: color[1], z: color[2] },
        });
        self.textures.len() as u32 - 1
    }

    /// the `rt_scene_desc` view of the tables (valid while `self` and `images` live)
    pub fn desc<'a>(&'a self, images: &'a mut Vec<sys::rt_image>) -> sys::rt_scene_desc {
        images.clear();
        images.extend(self.images.iter().map(|(w, h, px)| sys::rt_image { width: *w, height: *h, rgba: px.as_ptr() }));
        sys::rt_scene_desc {
            n_shapes: self.kind.len() as u32, kind: self.kind.as_ptr(), flags: self.flags.as_ptr(),
            inverse: self.inverse.as_ptr(), direct: self.direct.as_ptr(), params: self.params.as_ptr(),
            material: self.material.as_ptr(),
            n_materials: self.materials.len() as u32, materials: self.materials.as_ptr(),
            n_textures: self.textures.len() as u32, textures: self.textures.as_ptr(),
            n_images: images.len() as u32, images: images.as_ptr(),
            n_noise: self.noise.len() as u32, noise: self.noise.as_ptr(),
        }
    }
}
