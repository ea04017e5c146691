// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU (FP64) restatement of the rs-pathtracing hot path: primary-ray generation, nearest hit
// over the shape list, ray marching, material scatter / textures, the ray_color integrator and
// the threaded per-pixel accumulation.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product (rs_pathtracing_b200/) never
// does.  Every function cites the reference file:line it follows (paths are into
// dkarpushkin/rs-pathtracing).  Compile with -O2 -ffp-contract=off -fno-fast-math: Rust never
// contracts a*b+c into an FMA, so neither may this.
//
// PARITY PINNING: the reference's own tests pin only the transform / AABB / camera arithmetic
// (src/algebra/transform.rs:637-691, src/world/shapes/mod.rs:880-899, src/camera/mod.rs:315-343);
// those four known-answer tests are reproduced in tests/test_oracle_kat.py.  For nearest-hit, t,
// normal, scatter and pixel colour the reference holds no golden vector and cannot be built here
// (no Rust toolchain) => for those outputs this oracle is "parity unpinned": its only authority is
// the source text it restates line by line.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include "../include/rt_b200.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// algebra::Vector3d — src/algebra/mod.rs:23-28 and its operator impls :223-517
// ---------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }       // :223-269
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }       // :271-317
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                            // :495-517
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // :319-349 (`*` is dot)
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }         // :351-373
inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }         // :375-397
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }         // :399-421
inline V3 divide(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }          // :143-150
inline V3 product(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }         // :135-141
inline V3 cross(V3 a, V3 b) {                                                       // :99-105
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double squared_length(V3 a) { return dot(a, a); }                            // :112-115
inline double length(V3 a) { return std::sqrt(squared_length(a)); }                 // :117-120
inline V3 normalize(V3 a) { return a / length(a); }                                 // :107-110
inline bool approx_equal(double a, double b) { return std::fabs(a - b) < 1e-15; }   // :14-17
inline bool is_zero(V3 a) {                                                         // :156-158
    return approx_equal(a.x, 0.0) && approx_equal(a.y, 0.0) && approx_equal(a.z, 0.0);
}
// f64::min / f64::max ignore a NaN operand, exactly like C fmin / fmax             // :168-198
inline V3 vmin(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline double min_component(V3 a) { return std::fmin(std::fmin(a.x, a.y), a.z); }
inline double max_component(V3 a) { return std::fmax(std::fmax(a.x, a.y), a.z); }
V3 reflect(V3 v, V3 n);                                                             // :122-125
V3 refract(V3 v, V3 n, double ratio);                                               // :127-133

// ---------------------------------------------------------------------------------------------
// algebra::transform — src/algebra/transform.rs
// ---------------------------------------------------------------------------------------------
struct Mat4 {
    double m[4][4];
};
Mat4 mat_mul(const Mat4& a, const Mat4& b);          // transform.rs:553-570
Mat4 mat_translate(V3 v);                            // :316-323
Mat4 mat_scale(V3 v);                                // :325-332
Mat4 mat_rotate_roll(double deg);                    // :364-372
Mat4 mat_rotate_pitch(double deg);                   // :374-382
Mat4 mat_rotate_yaw(double deg);                     // :384-392
Mat4 mat_rotate(V3 deg);                             // :334-358
Mat4 mat_rotate_inverse(V3 deg);                     // :360-362
V3 transform_point(const Mat4& m, V3 p);             // :394-409
V3 transform_vector(const Mat4& m, V3 v);            // :411-417
V3 transform_normal(const Mat4& m, V3 n);            // :419-425
void inversable_transform_new(V3 translate, V3 rotate, V3 scale, Mat4* direct, Mat4* inverse);  // :16-23
void aabb_transform(V3 mn, V3 mx, const Mat4& m, V3* out_min, V3* out_max);  // shapes/mod.rs:93-108

// ---------------------------------------------------------------------------------------------
// world::ray — src/world/ray.rs
// ---------------------------------------------------------------------------------------------
struct Ray {
    V3 origin, direction;
};
inline Ray ray_new(V3 o, V3 d) { return Ray{o, normalize(d)}; }   // ray.rs:12-17

struct Hit {                     // RayHit, ray.rs:21-29
    V3 point;
    V3 normal;
    double distance;
    bool is_front_face;
    double u, v;
    int32_t shape;               // index into the flat list (stands in for the material reference)
};

// ---------------------------------------------------------------------------------------------
// camera — src/camera/mod.rs:71-88 and src/camera/ray_caster.rs:30-48,77-81
// ---------------------------------------------------------------------------------------------
rt_camera camera_new(V3 position, V3 direction, V3 up, double focal_length, double fov_rad);
struct RayCaster {
    V3 camera_position, camera_right, camera_up, left_top;
    double pixel_resolution;
    uint32_t width, height;
};
RayCaster raycaster_new(const rt_camera& cam, rt_image_params img);
Ray raycaster_get_ray(const RayCaster& rc, double x, double y);

// ---------------------------------------------------------------------------------------------
// counters the oracle reports with every run
// ---------------------------------------------------------------------------------------------
struct alignas(64) Counters {  // one cache line each: per-thread instances must not false-share
    uint64_t segments = 0, shape_tests = 0, march_steps = 0, march_rays = 0, aabb_tests = 0;
    void add(const Counters& o) {
        segments += o.segments; shape_tests += o.shape_tests; march_steps += o.march_steps;
        march_rays += o.march_rays; aabb_tests += o.aabb_tests;
    }
};

// ---------------------------------------------------------------------------------------------
// the scene as the oracle sees it: the flat description, decoded once
// ---------------------------------------------------------------------------------------------
struct Shape {
    int kind;
    bool inverse_normal;
    Mat4 direct, inverse;
    double p[RT_SHAPE_PARAMS];
    uint32_t material;
};
struct BvhNode {                 // shapes/mod.rs:621-625
    int left, right;             // >= 0: node index; < 0: ~shape index; right == INT32_MIN: none
    V3 bb_min, bb_max;
};
struct Scene {
    std::vector<Shape> shapes;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    struct Image { uint32_t w, h; std::vector<uint8_t> rgba; };
    std::vector<Image> images;
    std::vector<rt_perlin> noise;  // algebra::noise::Perlin tables
    std::vector<BvhNode> bvh;    // built on demand by build_bvh
    int bvh_root = -1;
};
bool scene_from_desc(const rt_scene_desc* d, Scene* out);
void shape_bounding_box(const Shape& s, V3* mn, V3* mx);     // get_bounding_box impls
void build_bvh(Scene& sc, uint64_t seed);                    // BvhNode::new, shapes/mod.rs:663-729

// Shape::ray_hit_transformed for one shape — shapes/mod.rs:112-124
bool shape_ray_hit(const Shape& s, int index, const Ray& ray, double min_t, double max_t, Hit* hit,
                   Counters* c);
// ShapeCollection::ray_intersect — shapes/mod.rs:573-597 (THE nearest-hit contract)
bool collection_ray_intersect(const Scene& sc, const Ray& ray, double min_t, double max_t, Hit* hit,
                              Counters* c);
// BvhNode::ray_hit — shapes/mod.rs:628-651 (what Scene::new really builds, world/mod.rs:35)
bool bvh_ray_hit(const Scene& sc, const Ray& ray, double min_t, double max_t, Hit* hit, Counters* c);

// ShapeFunction::intersect_bound of a ray-marched shape on an object-space ray (ray_marching.rs:135-145, 213-225)
bool march_intersect_bound(const Shape& s, const Ray& local, double* start, double* end);
// solve_quantic_equation (src/algebra/equation.rs:17-67): the four complex roots of a x^4 + b x^3 + c x^2 + d x + e
void solve_quantic_equation(double a, double b, double c, double d, double e, double re_out[4], double im_out[4]);

// Perlin::noise / turb — src/algebra/noise.rs:43-86
double perlin_noise(const rt_perlin& pn, V3 p);
double perlin_turb(const rt_perlin& pn, V3 p, int depth);

// implicit surfaces — ray_marching.rs:134-520
double surface_func(const double* p8, V3 p);
V3 surface_gradient(const double* p8, V3 p);

// ---------------------------------------------------------------------------------------------
// RNG.  The reference draws from rand::thread_rng (unseeded, unreproducible); only the
// distributions are restated.  Two generators:
//   PHILOX  — the counter-based schedule shared with the GPU core: stream of doubles keyed by
//             (seed; pixel, sample, event), event 0 = pixel jitter, event k+1 = scatter at hit k.
//             With it the oracle follows the SAME paths as the GPU renderer.
//   XOSHIRO — an unrelated sequential generator, for statistically independent renders.
// ---------------------------------------------------------------------------------------------
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
enum RngMode { RNG_PHILOX = 0, RNG_XOSHIRO = 1 };
struct PathRng {
    RngMode mode;
    // philox
    uint32_t key[2], pixel, sample, event, next_index;
    double cache[2];
    // xoshiro256**
    uint64_t s[4];
    void seed_xoshiro(uint64_t seed, uint64_t stream);
    void begin_event(uint32_t ev) { event = ev; next_index = 0; }
    double next();                       // uniform [0,1), 53 bits
};
V3 random_range(PathRng& r, double mn, double mx);    // Vector3d::random, algebra/mod.rs:59-66
V3 random_in_unit_sphere(PathRng& r);                 // :77-84
V3 random_unit(PathRng& r);                           // :86-88

// ---------------------------------------------------------------------------------------------
// materials / textures / integrator
// ---------------------------------------------------------------------------------------------
V3 texture_value(const Scene& sc, uint32_t tex, double u, double v, V3 p);          // texture.rs:17-116
// Material::scatter; returns false when the material does not scatter                material.rs:42-115
bool material_scatter(const Scene& sc, const rt_material& m, const Ray& ray, const Hit& hit,
                      PathRng& rng, Ray* scattered, V3* attenuation);
V3 material_emitted(const Scene& sc, const rt_material& m, double u, double v, V3 p);  // :123-127
V3 background(const Ray& ray);                                                       // world/mod.rs:199-202
// ray_color — renderer/mod.rs:23-45 (kept recursive like the reference)
V3 ray_color(const Scene& sc, bool use_bvh, const Ray& ray, uint32_t depth, PathRng& rng,
             uint32_t hit_number, Counters* c);

struct RenderOptions {
    rt_image_params image;
    uint32_t samples_number, max_depth;
    uint64_t seed;
    RngMode rng;
    bool use_bvh;
    uint32_t threads;
    // bounded sample for timing: only pixels with (x % stride_x == 0 && y % stride_y == 0) are
    // traced (the others stay untouched); 1/1 = the whole image
    uint32_t stride_x, stride_y;
};
// ThreadPoolRenderer (step_by_step.rs:37-121) + dispatcher/worker threads (renderer/mod.rs:66-155):
// one serial dispatcher producing chunks of w*h/threads/8 pixels, `threads` workers, per-pixel mean.
// Returns wall seconds from start_rendering to the last worker's None (main_raylib.rs:225-237).
double render(const Scene& sc, const rt_camera& cam, const RenderOptions& opt, rt_vec3* buffer,
              Counters* counters);

}  // namespace orc
