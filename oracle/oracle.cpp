// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  CPU FP64 restatement of the reference's
// hot path.  Expressions are transcribed with the reference's operand order and associativity;
// build with -ffp-contract=off so nothing is fused.
#include "oracle.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

namespace orc {

static const double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI

// ---------------------------------------------------------------------------------------------
// algebra/mod.rs
// ---------------------------------------------------------------------------------------------
V3 reflect(V3 v, V3 n) {  // algebra/mod.rs:122-125
    V3 b = dot(v, n) * n;  // (self * normal) * normal  -> f64 * &Vector3d
    return v - (2.0 * b);
}

V3 refract(V3 v, V3 n, double ratio) {  // algebra/mod.rs:127-133
    double cos_theta = dot(-v, n);
    V3 r_out_perp = ratio * (v + cos_theta * n);
    V3 r_out_parallel = -(std::sqrt(std::fabs(1.0 - squared_length(r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}

// ---------------------------------------------------------------------------------------------
// algebra/transform.rs
// ---------------------------------------------------------------------------------------------
Mat4 mat_mul(const Mat4& a, const Mat4& b) {  // transform.rs:553-570
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] +
                        a.m[i][3] * b.m[3][j];
    return r;
}
Mat4 mat_translate(V3 v) {  // :316-323
    return Mat4{{{1.0, 0.0, 0.0, v.x}, {0.0, 1.0, 0.0, v.y}, {0.0, 0.0, 1.0, v.z}, {0.0, 0.0, 0.0, 1.0}}};
}
Mat4 mat_scale(V3 v) {  // :325-332
    return Mat4{{{v.x, 0.0, 0.0, 0.0}, {0.0, v.y, 0.0, 0.0}, {0.0, 0.0, v.z, 0.0}, {0.0, 0.0, 0.0, 1.0}}};
}
static inline double to_radians(double deg) { return deg * (PI / 180.0); }  // f64::to_radians
Mat4 mat_rotate_roll(double deg) {  // :364-372
    double r = to_radians(deg);
    return Mat4{{{1.0, 0.0, 0.0, 0.0},
                 {0.0, std::cos(r), -std::sin(r), 0.0},
                 {0.0, std::sin(r), std::cos(r), 0.0},
                 {0.0, 0.0, 0.0, 1.0}}};
}
Mat4 mat_rotate_pitch(double deg) {  // :374-382
    double r = to_radians(deg);
    return Mat4{{{std::cos(r), 0.0, std::sin(r), 0.0},
                 {0.0, 1.0, 0.0, 0.0},
                 {-std::sin(r), 0.0, std::cos(r), 0.0},
                 {0.0, 0.0, 0.0, 1.0}}};
}
Mat4 mat_rotate_yaw(double deg) {  // :384-392
    double r = to_radians(deg);
    return Mat4{{{std::cos(r), -std::sin(r), 0.0, 0.0},
                 {std::sin(r), std::cos(r), 0.0, 0.0},
                 {0.0, 0.0, 1.0, 0.0},
                 {0.0, 0.0, 0.0, 1.0}}};
}
Mat4 mat_rotate(V3 d) {  // :357  roll * pitch * yaw, left-associated
    return mat_mul(mat_mul(mat_rotate_roll(d.x), mat_rotate_pitch(d.y)), mat_rotate_yaw(d.z));
}
Mat4 mat_rotate_inverse(V3 d) {  // :360-362  yaw * pitch * roll
    return mat_mul(mat_mul(mat_rotate_yaw(d.z), mat_rotate_pitch(d.y)), mat_rotate_roll(d.x));
}
V3 transform_point(const Mat4& t, V3 p) {  // :394-409
    return {p.x * t.m[0][0] + p.y * t.m[0][1] + p.z * t.m[0][2] + t.m[0][3],
            p.x * t.m[1][0] + p.y * t.m[1][1] + p.z * t.m[1][2] + t.m[1][3],
            p.x * t.m[2][0] + p.y * t.m[2][1] + p.z * t.m[2][2] + t.m[2][3]};
}
V3 transform_vector(const Mat4& t, V3 v) {  // :411-417
    return {v.x * t.m[0][0] + v.y * t.m[0][1] + v.z * t.m[0][2],
            v.x * t.m[1][0] + v.y * t.m[1][1] + v.z * t.m[1][2],
            v.x * t.m[2][0] + v.y * t.m[2][1] + v.z * t.m[2][2]};
}
V3 transform_normal(const Mat4& t, V3 n) {  // :419-425 (transpose of the 3x3)
    return {n.x * t.m[0][0] + n.y * t.m[1][0] + n.z * t.m[2][0],
            n.x * t.m[0][1] + n.y * t.m[1][1] + n.z * t.m[2][1],
            n.x * t.m[0][2] + n.y * t.m[1][2] + n.z * t.m[2][2]};
}
void inversable_transform_new(V3 translate, V3 rotate, V3 scale, Mat4* direct, Mat4* inverse) {  // :16-23
    *direct = mat_mul(mat_mul(mat_translate(translate), mat_rotate(rotate)), mat_scale(scale));
    *inverse = mat_mul(mat_mul(mat_scale(v3(1.0 / scale.x, 1.0 / scale.y, 1.0 / scale.z)),
                               mat_rotate_inverse(v3(-rotate.x, -rotate.y, -rotate.z))),
                       mat_translate(v3(-translate.x, -translate.y, -translate.z)));
}
void aabb_transform(V3 mn, V3 mx, const Mat4& m, V3* out_min, V3* out_max) {  // shapes/mod.rs:93-108
    V3 c[2] = {mn, mx};
    V3 lo = v3(INFINITY, INFINITY, INFINITY), hi = v3(-INFINITY, -INFINITY, -INFINITY);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                V3 p = transform_point(m, v3(c[i].x, c[j].y, c[k].z));
                lo = vmin(lo, p);
                hi = vmax(hi, p);
            }
    *out_min = lo;
    *out_max = hi;
}

// ---------------------------------------------------------------------------------------------
// camera
// ---------------------------------------------------------------------------------------------
static inline rt_vec3 to_rt(V3 a) { return rt_vec3{a.x, a.y, a.z}; }
static inline V3 from_rt(rt_vec3 a) { return V3{a.x, a.y, a.z}; }

rt_camera camera_new(V3 position, V3 direction, V3 up_vector, double focal_length, double fov) {
    // camera/mod.rs:71-88
    V3 right_vec = normalize(cross(direction, up_vector));
    rt_camera c;
    c.position = to_rt(position);
    c.direction = to_rt(normalize(direction));
    c.up = to_rt(normalize(cross(right_vec, direction)));
    c.right = to_rt(right_vec);
    c.fov_rad = fov;
    c.focal_length = focal_length;
    return c;
}

RayCaster raycaster_new(const rt_camera& cam, rt_image_params img) {  // ray_caster.rs:30-48
    V3 pos = from_rt(cam.position), dir = from_rt(cam.direction), right = from_rt(cam.right),
       up = from_rt(cam.up);
    V3 center = pos + cam.focal_length * dir;
    double aspect_ratio = (double)img.width / (double)img.height;
    double viewport_width = std::tan(cam.fov_rad / 2.0) * cam.focal_length * 2.0;
    double viewport_height = viewport_width / aspect_ratio;
    RayCaster rc;
    rc.left_top = center - right * (viewport_width / 2.0) + up * (viewport_height / 2.0);
    rc.pixel_resolution = viewport_width / (double)img.width;
    rc.camera_position = pos;
    rc.camera_right = right;
    rc.camera_up = up;
    rc.width = img.width;
    rc.height = img.height;
    return rc;
}

Ray raycaster_get_ray(const RayCaster& rc, double x, double y) {  // ray_caster.rs:77-81,109-112
    V3 dir = rc.left_top + (rc.pixel_resolution * x) * rc.camera_right -
             (rc.pixel_resolution * y) * rc.camera_up;
    return ray_new(rc.camera_position, dir - rc.camera_position);
}

// ---------------------------------------------------------------------------------------------
// scene decoding
// ---------------------------------------------------------------------------------------------
bool scene_from_desc(const rt_scene_desc* d, Scene* out) {
    out->shapes.resize(d->n_shapes);
    for (uint32_t i = 0; i < d->n_shapes; i++) {
        Shape& s = out->shapes[i];
        s.kind = d->kind[i];
        s.inverse_normal = (d->flags[i] & RT_SHAPE_FLAG_INVERSE_NORMAL) != 0;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) {
                s.direct.m[r][c] = d->direct[i * 12 + r * 4 + c];
                s.inverse.m[r][c] = d->inverse[i * 12 + r * 4 + c];
            }
        for (int c = 0; c < 4; c++) s.direct.m[3][c] = s.inverse.m[3][c] = (c == 3) ? 1.0 : 0.0;
        for (int k = 0; k < RT_SHAPE_PARAMS; k++) s.p[k] = d->params[i * RT_SHAPE_PARAMS + k];
        s.material = d->material[i];
        if (s.material >= d->n_materials) return false;
    }
    out->materials.assign(d->materials, d->materials + d->n_materials);
    out->textures.assign(d->textures, d->textures + d->n_textures);
    out->images.resize(d->n_images);
    for (uint32_t i = 0; i < d->n_images; i++) {
        out->images[i].w = d->images[i].width;
        out->images[i].h = d->images[i].height;
        out->images[i].rgba.assign(d->images[i].rgba,
                                   d->images[i].rgba + (size_t)4 * d->images[i].width * d->images[i].height);
    }
    out->noise.assign(d->noise, d->noise + (d->noise ? d->n_noise : 0));
    return true;
}

// ---------------------------------------------------------------------------------------------
// implicit surfaces — ray_marching.rs
// ---------------------------------------------------------------------------------------------
double surface_func(const double* q, V3 p) {
    switch ((int)q[0]) {
        case RT_SURF_HEART: {  // :147-155
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double z3 = z2 * p.z;
            double a = x2 + (9.0 / 4.0) * y2 + z2 - 1.0;
            return a * a * a - x2 * z3 - (9.0 / 80.0) * y2 * z3;
        }
        case RT_SURF_SINE: {  // :203-211
            double a_ = q[3];
            return a_ * a_ * (p.x - p.y - p.z) * (p.x + p.y - p.z) * (p.x - p.y + p.z) *
                       (p.x + p.y + p.z) +
                   4.0 * p.x * p.x * p.y * p.y * p.z * p.z;
        }
        case RT_SURF_STAR: {  // :268-274
            double a_ = q[3];
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double c = x2 + y2 + z2 - 1.0;
            return a_ * (x2 * y2 + x2 * z2 + y2 * z2) + (c * c * c);
        }
        case RT_SURF_DUPIN: {  // :340-345
            double a_ = q[3], b_ = q[4], c_ = q[5], d_ = q[6];
            double b2 = b_ * b_;
            double e = p.x * p.x + p.y * p.y + p.z * p.z + b2 - d_ * d_;
            double f = a_ * p.x - c_ * d_;
            return e * e - 4.0 * (f * f + b2 * p.y * p.y);
        }
        case RT_SURF_HUNTS: {  // :399-406
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a = x2 + y2 + z2 - 13.0;
            double b = 3.0 * x2 + y2 - 4.0 * z2 - 12.0;
            return 4.0 * a * a * a + 27.0 * b * b;
        }
        case RT_SURF_CUSHION: {  // :464-478
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a = x2 - p.z;
            return z2 * x2 - z2 * z2 - 2.0 * p.z * x2 + 2.0 * p.z * z2 + x2 - z2 - a * a - y2 * y2 -
                   2.0 * x2 * y2 - y2 * z2 + 2.0 * y2 * p.z + y2;
        }
    }
    return NAN;
}

V3 surface_gradient(const double* q, V3 p) {
    switch ((int)q[0]) {
        case RT_SURF_HEART: {  // :157-168
            double a = p.x * p.x + (9.0 / 4.0) * p.y * p.y + p.z * p.z - 1.0;
            a = 3.0 * a * a;
            double z2 = p.z * p.z;
            double z3 = z2 * p.z;
            return v3(2.0 * p.x * (a - z3), (9.0 / 2.0) * p.y * (a - 0.05 * z3),
                      2.0 * p.z * (a - p.z * (1.5 * p.x * p.x + (27.0 / 40.0) * p.y * p.y)));
        }
        case RT_SURF_SINE: {  // :227-237
            double a_ = q[3];
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a2 = a_ * a_;
            return v3(4.0 * p.x * (a2 * (x2 - y2 - z2) + 2.0 * y2 * z2),
                      8.0 * x2 * p.y * z2 - 4.0 * a2 * p.y * (x2 - y2 + z2),
                      8.0 * x2 * y2 * p.z - 4.0 * a2 * p.z * (x2 + y2 - z2));
        }
        case RT_SURF_STAR: {  // :290-300
            double a_ = q[3];
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double c = x2 + y2 + z2 - 1.0;
            return v3(2.0 * a_ * p.x * (y2 + z2) + 6.0 * p.x * c * c,
                      2.0 * a_ * p.y * (x2 + z2) + 6.0 * p.y * c * c,
                      2.0 * a_ * p.z * (x2 + y2) + 6.0 * p.z * c * c);
        }
        case RT_SURF_DUPIN: {  // :361-369
            double a_ = q[3], b_ = q[4], c_ = q[5], d_ = q[6];
            double b2 = b_ * b_;
            double e = 4.0 * (p.x * p.x + p.y * p.y + p.z * p.z + b2 - d_ * d_);
            return v3(e * p.x - 8.0 * a_ * (a_ * p.x - c_ * d_), e * p.y - 8.0 * b2 * p.y, e * p.z);
        }
        case RT_SURF_HUNTS: {  // :422-434
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a = x2 + y2 + z2 - 13.0;
            double b = 3.0 * x2 + y2 - 4.0 * (z2 + 3.0);
            return v3(24.0 * p.x * a * a + 324.0 * p.x * b, 12.0 * p.y * (2.0 * a * a + 9.0 * b),
                      24.0 * p.z * (a * a - 18.0 * b));
        }
        case RT_SURF_CUSHION: {  // :494-504
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            return v3(2.0 * p.x * (-2.0 * x2 - 2.0 * y2 + z2 + 1.0),
                      -2.0 * p.y * (2.0 * x2 + 2.0 * y2 + z2 - 2.0 * p.z - 1.0),
                      2.0 * p.z * (x2 - 2.0 * z2 + 3.0 * p.z - 2.0) - 2.0 * p.y * (p.z - 1.0));
        }
    }
    return v3(NAN, NAN, NAN);
}

static inline bool surface_has_uv(int kind) {  // uv(): (0,0) for Heart/Sine/Star, (p.x,p.y) otherwise
    return kind == RT_SURF_DUPIN || kind == RT_SURF_HUNTS || kind == RT_SURF_CUSHION;
}

// algebra/equation.rs:5-15
static bool solve_quadratic_equation(double a, double half_b, double c, double* x1, double* x2) {
    double d = half_b * half_b - a * c;
    double d_sqrt = std::sqrt(d);
    if (d < 0.0) return false;
    if (d == 0.0) {
        *x1 = -half_b;
        *x2 = -half_b;
        return true;
    }
    *x1 = (-half_b - d_sqrt) / a;
    *x2 = (-half_b + d_sqrt) / a;
    return true;
}

// ShapeFunction::intersect_bound — ray_marching.rs:135-145 (Heart), :213-225 etc. (sphere bound)
static bool intersect_bound(const double* q, V3 origin, V3 dir, double* start, double* end) {
    double x1, x2;
    if ((int)q[0] == RT_SURF_HEART) {
        double sr = 1.45;  // Heart::new, :126-131
        V3 radius = v3(sr, sr / 2.05, sr);
        V3 o = divide(origin, radius);
        V3 d = divide(dir, radius);
        if (!solve_quadratic_equation(dot(d, d), dot(d, o), dot(o, o) - 1.0, &x1, &x2)) return false;
    } else {
        double R = q[7];
        if (!solve_quadratic_equation(dot(dir, dir), dot(dir, origin), dot(origin, origin) - R * R, &x1,
                                      &x2))
            return false;
    }
    if (x1 < 0.0 && x2 < 0.0) return false;
    *start = std::fmax(x1, 0.0);
    *end = std::fmax(x2, 0.0);
    return true;
}

// ---------------------------------------------------------------------------------------------
// object-space intersections.  They fill an object-space Hit: point, UN-normalised normal, t, u, v;
// ray_hit_new below applies RayHit::new.
// ---------------------------------------------------------------------------------------------
struct ObjHit {
    V3 p, n;
    double t, u, v;
};

static bool rectangle_intersect(const Shape& s, const Ray& ray, double min_t, double max_t, ObjHit* h) {
    // shapes/mod.rs:181-204
    double x0 = s.p[0], y0 = s.p[1], x1 = s.p[2], y1 = s.p[3];
    double t = -ray.origin.z / ray.direction.z;
    if (t < min_t || t > max_t) return false;
    V3 p = ray.origin + ray.direction * t;
    if (p.x < x0 || p.x > x1 || p.y < y0 || p.y > y1) return false;
    h->u = (p.x - x0) / (x1 - x0);
    h->v = (p.y - y0) / (y1 - y0);
    h->p = p;
    h->n = v3(0.0, 0.0, 1.0);
    h->t = t;
    return true;
}

static bool cube_intersect(const Shape&, const Ray& ray, double min_t, double max_t, ObjHit* h) {
    // shapes/mod.rs:250-285
    V3 min_p = v3(-1.0, -1.0, -1.0), max_p = v3(1.0, 1.0, 1.0);
    V3 t_lower = divide(min_p - ray.origin, ray.direction);
    V3 t_upper = divide(max_p - ray.origin, ray.direction);
    V3 t_mins = vmin(t_lower, t_upper);
    V3 t_maxes = vmax(t_lower, t_upper);
    double t_box_min = std::fmax(max_component(t_mins), min_t);
    double t_box_max = std::fmin(min_component(t_maxes), max_t);
    if (t_box_min > t_box_max || t_box_min > max_t) return false;
    V3 p = ray.origin + t_box_min * ray.direction;
    V3 p_abs = vabs(p);
    double max_c = max_component(p_abs);
    if (max_c == p_abs.x) {
        h->n = v3(p.x, 0.0, 0.0); h->u = p.y; h->v = p.z;
    } else if (max_c == p_abs.y) {
        h->n = v3(0.0, p.y, 0.0); h->u = p.x; h->v = p.z;
    } else if (max_c == p_abs.z) {
        h->n = v3(0.0, 0.0, p.z); h->u = p.x; h->v = p.y;
    } else {
        // the reference panics here (max_c is NaN); the port and the GPU core report a miss
        return false;
    }
    h->p = p;
    h->t = t_box_min;
    return true;
}

static bool sphere_intersect(const Shape& s, const Ray& ray, double min_t, double max_t, ObjHit* h) {
    // shapes/mod.rs:330-374
    V3 origin = ray.origin, dir = ray.direction;
    double a = dot(dir, dir);
    double half_b = dot(dir, origin);
    double c = dot(origin, origin) - 1.0;
    double d = half_b * half_b - a * c;
    double x;
    if (d < 0.0) {
        return false;
    } else if (d == 0.0) {
        x = -half_b * a;  // sic: multiplied, and accepted without a range check
    } else {
        x = (-half_b - std::sqrt(d)) / a;
        if (x < min_t || x > max_t) {
            x = (-half_b + std::sqrt(d)) / a;
            if (x < min_t || x > max_t) return false;
        }
    }
    V3 p = origin + dir * x;
    h->n = s.inverse_normal ? -p : p;
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    h->u = phi / (2.0 * PI);
    h->v = theta / PI;
    h->p = p;
    h->t = x;
    return true;
}

bool march_intersect_bound(const Shape& s, const Ray& local, double* start, double* end) {
    return intersect_bound(s.p, local.origin, local.direction, start, end);
}

static bool march_intersect(const Shape& s, const Ray& ray, double min_t, double max_t, ObjHit* h,
                            Counters* c) {
    // ray_marching.rs:20-74
    V3 origin = ray.origin, dir = ray.direction;
    double start, end;
    if (!intersect_bound(s.p, origin, dir, &start, &end)) return false;
    if (c) c->march_rays++;
    double step = s.p[1];
    int depth = (int)s.p[2];
    double t = start;
    V3 p = origin + t * dir;
    double r = surface_func(s.p, p);
    for (int it = 0; it < depth; it++) {
        bool finished = false;
        for (;;) {
            if (t > end || t < start) return false;
            t += step;
            V3 sd = step * dir;
            p.x += sd.x; p.y += sd.y; p.z += sd.z;
            double next = surface_func(s.p, p);
            if (c) c->march_steps++;
            if (approx_equal(next, 0.0)) { finished = true; break; }
            if ((r < 0.0 && next > 0.0) || (r > 0.0 && next < 0.0)) {
                step *= -0.01;
                r = next;
                break;
            }
            r = next;
        }
        if (finished) break;
    }
    if (t < min_t || t > max_t) return false;
    V3 hp = origin + dir * t;
    h->p = hp;
    h->n = surface_gradient(s.p, hp);
    if (surface_has_uv((int)s.p[0])) { h->u = hp.x; h->v = hp.y; } else { h->u = 0.0; h->v = 0.0; }
    h->t = t;
    return true;
}


// ---------------------------------------------------------------------------------------------
// Torus -- shapes/mod.rs:403-494, on solve_quantic_equation -- algebra/equation.rs:17-67.
// The solver runs on num::Complex<f64> (crate `num ^0.4`, Cargo.toml:21 -> num-complex 0.4.x; not vendored).  Its
// arithmetic is restated here from that crate's published source: +, -, scalar * and / componentwise; complex *
// (re re' - im im', re im' + im re'); complex / via norm_sqr = re'^2 + im'^2; sqrt / cbrt by cases (im == 0, re == 0,
// else polar form hypot / atan2 -> root of r, angle / n -> r cos, r sin), signs of zero respected.
// ---------------------------------------------------------------------------------------------
namespace {
struct Cx {
    double re, im;
};
inline Cx cx(double re) { return Cx{re, 0.0}; }                       // f64::into()
inline Cx operator+(Cx a, Cx b) { return Cx{a.re + b.re, a.im + b.im}; }
inline Cx operator-(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
inline Cx operator-(Cx a) { return Cx{-a.re, -a.im}; }
inline Cx operator*(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx operator/(Cx a, Cx b) {
    double norm_sqr = b.re * b.re + b.im * b.im;
    double re = a.re * b.re + a.im * b.im;
    double im = a.im * b.re - a.re * b.im;
    return Cx{re / norm_sqr, im / norm_sqr};
}
inline Cx operator*(double k, Cx a) { return Cx{k * a.re, k * a.im}; }   // impl Mul<Complex<f64>> for f64
inline Cx operator/(Cx a, double k) { return Cx{a.re / k, a.im / k}; }
inline Cx operator+(double k, Cx a) { return Cx{k + a.re, a.im}; }
inline Cx cx_from_polar(double r, double theta) { return Cx{r * std::cos(theta), r * std::sin(theta)}; }
inline Cx cx_sqrt(Cx z) {
    if (z.im == 0.0) {
        if (!std::signbit(z.re)) return Cx{std::sqrt(z.re), z.im};
        double im = std::sqrt(-z.re);                                   // sqrt(r e^(i pi)) = i sqrt(r)
        return !std::signbit(z.im) ? Cx{0.0, im} : Cx{0.0, -im};
    }
    if (z.re == 0.0) {
        double x = std::sqrt(std::fabs(z.im) / 2.0);                    // sqrt(r e^(i pi/2)) = sqrt(r/2)(1 + i)
        return !std::signbit(z.im) ? Cx{x, x} : Cx{x, -x};
    }
    return cx_from_polar(std::sqrt(std::hypot(z.re, z.im)), std::atan2(z.im, z.re) / 2.0);
}
inline Cx cx_cbrt(Cx z) {
    if (z.im == 0.0) {
        if (!std::signbit(z.re)) return Cx{std::cbrt(z.re), z.im};
        double re = std::cbrt(-z.re) / 2.0;                             // cbrt(r e^(i pi)) = cbrt(r) e^(i pi/3)
        double im = std::sqrt(3.0) * re;
        return !std::signbit(z.im) ? Cx{re, im} : Cx{re, -im};
    }
    if (z.re == 0.0) {
        double im = std::cbrt(std::fabs(z.im)) / 2.0;                   // cbrt(r e^(i pi/2)) = cbrt(r) e^(i pi/6)
        double re = std::sqrt(3.0) * im;
        return !std::signbit(z.im) ? Cx{re, im} : Cx{re, -im};
    }
    return cx_from_polar(std::cbrt(std::hypot(z.re, z.im)), std::atan2(z.im, z.re) / 3.0);
}
}  // namespace

// equation.rs:17-67
void solve_quantic_equation(double a_, double b_, double c_, double d_, double e_, double re_out[4], double im_out[4]) {
    Cx a = cx(a_), b = cx(b_) / a, c = cx(c_) / a, d = cx(d_) / a, e = cx(e_) / a;
    Cx b2 = b * b;
    Cx alpha = c - (3.0 / 8.0) * b2;
    Cx beta = (b2 * b) / 8.0 - (b * c) / 2.0 + d;
    Cx gamma = (-3.0 / 256.0) * b2 * b2 + b2 * c / 16.0 - b * d / 4.0 + e;
    Cx alpha2 = alpha * alpha;
    Cx t = -b / 4.0;
    Cx roots[4];
    if (approx_equal(beta.re, 0.0) && approx_equal(beta.im, 0.0)) {
        Cx r = cx_sqrt(alpha2 - 4.0 * gamma);
        Cx r1 = cx_sqrt((-alpha + r) / 2.0);
        Cx r2 = cx_sqrt((-alpha - r) / 2.0);
        roots[0] = t + r1; roots[1] = t - r1; roots[2] = t + r2; roots[3] = t - r2;
    } else {
        Cx p = -(alpha2 / 12.0 + gamma);
        Cx q = -alpha2 * alpha / 108.0 + alpha * gamma / 3.0 - beta * beta / 8.0;
        Cx r = -q / 2.0 + cx_sqrt(q * q / 4.0 + p * p * p / 27.0);
        Cx u = cx_cbrt(r);
        Cx y = (-5.0 / 6.0) * alpha + u;
        if (approx_equal(u.re, 0.0) && approx_equal(u.im, 0.0)) y = y - cx_cbrt(q);
        else y = y - p / (3.0 * u);
        Cx w = cx_sqrt(alpha + 2.0 * y);
        Cx r1 = cx_sqrt(-(3.0 * alpha + 2.0 * y + 2.0 * beta / w));
        Cx r2 = cx_sqrt(-(3.0 * alpha + 2.0 * y - 2.0 * beta / w));
        roots[0] = t + (w - r1) / 2.0; roots[1] = t + (w + r1) / 2.0;
        roots[2] = t + (-w - r2) / 2.0; roots[3] = t + (-w + r2) / 2.0;
    }
    for (int k = 0; k < 4; k++) { re_out[k] = roots[k].re; im_out[k] = roots[k].im; }
}

static bool torus_intersect(const Shape& s, const Ray& ray, double min_t, double max_t, ObjHit* h) {
    // shapes/mod.rs:429-476
    const double PI = 3.14159265358979323846264338327950288;
    V3 origin = ray.origin, dir = ray.direction;
    double radius = s.p[0], tube_radius = s.p[1];
    double t = 4.0 * radius * radius;
    double g = t * (dir.x * dir.x + dir.y * dir.y);
    double hh = 2.0 * t * (origin.x * dir.x + origin.y * dir.y);
    double i = t * (origin.x * origin.x + origin.y * origin.y);
    double j = dot(dir, dir);
    double k = 2.0 * dot(origin, dir);
    double l = dot(origin, origin) + radius * radius - tube_radius * tube_radius;
    double a = j * j;
    double b = 2.0 * j * k;
    double c = 2.0 * j * l + k * k - g;
    double d = 2.0 * k * l - hh;
    double e = l * l - i;
    double re[4], im[4];
    solve_quantic_equation(a, b, c, d, e, re, im);
    double min_root = INFINITY;
    for (int r = 0; r < 4; r++)
        if (approx_equal(im[r], 0.0) && re[r] < min_root) min_root = re[r];
    if (std::isinf(min_root) || min_root < min_t || min_root > max_t) return false;
    V3 p = origin + min_root * dir;
    h->p = p;
    h->n = p - normalize(v3(p.x, p.y, 0.0)) * radius;
    double theta = std::asin(p.z / tube_radius);
    double phi = std::acos(p.z / (radius + tube_radius * std::cos(theta))) + PI;
    h->u = phi / (2.0 * PI);
    h->v = theta / PI;
    h->t = min_root;
    return true;
}

// Shape::ray_hit_transformed (shapes/mod.rs:112-124) around RayHit::new / set_normal (ray.rs:32-64)
bool shape_ray_hit(const Shape& s, int index, const Ray& ray, double min_t, double max_t, Hit* hit,
                   Counters* c) {
    if (c) c->shape_tests++;
    Ray local;  // InversableTransform::inverse_transform_ray, transform.rs:32-37 (no renormalisation)
    local.origin = transform_point(s.inverse, ray.origin);
    local.direction = transform_vector(s.inverse, ray.direction);
    ObjHit oh;
    bool ok = false;
    switch (s.kind) {
        case RT_SHAPE_SPHERE: ok = sphere_intersect(s, local, min_t, max_t, &oh); break;
        case RT_SHAPE_CUBE: ok = cube_intersect(s, local, min_t, max_t, &oh); break;
        case RT_SHAPE_RECTANGLE: ok = rectangle_intersect(s, local, min_t, max_t, &oh); break;
        case RT_SHAPE_MARCH: ok = march_intersect(s, local, min_t, max_t, &oh, c); break;
        case RT_SHAPE_TORUS: ok = torus_intersect(s, local, min_t, max_t, &oh); break;
    }
    if (!ok) return false;
    V3 n_obj = normalize(oh.n);                        // RayHit::new, ray.rs:42-45
    hit->point = transform_point(s.direct, oh.p);      // shapes/mod.rs:117
    V3 n_w = transform_normal(s.inverse, n_obj);       // shapes/mod.rs:118
    bool front = dot(n_w, ray.direction) < 0.0;        // set_normal, ray.rs:60-64
    hit->normal = normalize(front ? n_w : -n_w);
    hit->is_front_face = front;
    hit->distance = oh.t;
    hit->u = oh.u;
    hit->v = oh.v;
    hit->shape = index;
    return true;
}

bool collection_ray_intersect(const Scene& sc, const Ray& ray, double min_t, double max_t, Hit* out,
                              Counters* c) {
    // shapes/mod.rs:587-596
    double min_distance = max_t;
    bool any = false;
    Hit h;
    for (size_t i = 0; i < sc.shapes.size(); i++) {
        if (shape_ray_hit(sc.shapes[i], (int)i, ray, min_t, min_distance, &h, c)) {
            min_distance = h.distance;
            *out = h;
            any = true;
        }
    }
    return any;
}

// ---------------------------------------------------------------------------------------------
// BVH — shapes/mod.rs:18-109 (AABB), :621-729 (BvhNode)
// ---------------------------------------------------------------------------------------------
void shape_bounding_box(const Shape& s, V3* mn, V3* mx) {
    V3 lo, hi;
    switch (s.kind) {
        case RT_SHAPE_RECTANGLE:  // :214-220
            lo = v3(s.p[0], s.p[1], -0.0001);
            hi = v3(s.p[2], s.p[3], 0.0001);
            break;
        case RT_SHAPE_MARCH:  // ray_marching.rs:84-91 + get_bounds
            if ((int)s.p[0] == RT_SURF_HEART) {
                double sr = 1.45;
                hi = v3(sr, sr / 2.05, sr);
            } else {
                hi = v3(s.p[7], s.p[7], s.p[7]);
            }
            lo = -hi;
            break;
        case RT_SHAPE_TORUS: {  // :486-493
            double a = s.p[0] + s.p[1];
            lo = v3(-a, -a, -s.p[1]);
            hi = v3(a, a, s.p[1]);
            break;
        }
        default:  // Sphere :384-398, Cube :295-301
            lo = v3(-1.0, -1.0, -1.0);
            hi = v3(1.0, 1.0, 1.0);
    }
    aabb_transform(lo, hi, s.direct, mn, mx);
}

static bool aabb_ray_hit(V3 mn, V3 mx, const Ray& ray, double min_t, double max_t) {  // :68-79
    V3 t_lower = divide(mn - ray.origin, ray.direction);
    V3 t_upper = divide(mx - ray.origin, ray.direction);
    V3 t_mins = vmin(t_lower, t_upper);
    V3 t_maxes = vmax(t_lower, t_upper);
    double t_box_min = std::fmax(max_component(t_mins), min_t);
    double t_box_max = std::fmin(min_component(t_maxes), max_t);
    return t_box_min <= t_box_max;
}

static const int NO_CHILD = INT_MIN;

static uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int bvh_build_rec(Scene& sc, std::vector<int> ids, const std::vector<V3>& mins,
                         const std::vector<V3>& maxs, uint64_t& rng) {
    // BvhNode::new, :663-729.  The reference draws the axis from thread_rng().gen_range(0..2)
    // (x or y only) and sorts with a never-Equal comparator; a seeded generator stands in.
    int axis = (int)(splitmix64(rng) >> 63);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
        return axis == 0 ? mins[a].x < mins[b].x : mins[a].y < mins[b].y;
    });
    size_t n = ids.size();
    BvhNode node;
    if (n == 1) {
        node.left = ~ids[0];
        node.right = NO_CHILD;
    } else if (n == 2) {
        // shapes.swap_remove(0) twice: first takes element 0 (and moves the last into slot 0),
        // second takes that moved element => (ids[0], ids[1])
        node.left = ~ids[0];
        node.right = ~ids[1];
    } else {
        std::vector<int> l(ids.begin(), ids.begin() + n / 2), r(ids.begin() + n / 2, ids.end());
        node.left = bvh_build_rec(sc, l, mins, maxs, rng);
        node.right = bvh_build_rec(sc, r, mins, maxs, rng);
    }
    auto child_box = [&](int ch, V3* mn, V3* mx) {
        if (ch >= 0) { *mn = sc.bvh[ch].bb_min; *mx = sc.bvh[ch].bb_max; }
        else { *mn = mins[~ch]; *mx = maxs[~ch]; }
    };
    V3 lmn, lmx;
    child_box(node.left, &lmn, &lmx);
    if (node.right != NO_CHILD) {
        V3 rmn, rmx;
        child_box(node.right, &rmn, &rmx);
        node.bb_min = vmin(lmn, rmn);
        node.bb_max = vmax(lmx, rmx);
    } else {
        node.bb_min = lmn;
        node.bb_max = lmx;
    }
    sc.bvh.push_back(node);
    return (int)sc.bvh.size() - 1;
}

void build_bvh(Scene& sc, uint64_t seed) {
    sc.bvh.clear();
    sc.bvh_root = -1;
    size_t n = sc.shapes.size();
    if (n == 0) return;
    std::vector<V3> mins(n), maxs(n);
    std::vector<int> ids(n);
    for (size_t i = 0; i < n; i++) {
        shape_bounding_box(sc.shapes[i], &mins[i], &maxs[i]);
        ids[i] = (int)i;
    }
    uint64_t rng = seed;
    sc.bvh_root = bvh_build_rec(sc, ids, mins, maxs, rng);
}

static bool bvh_child_hit(const Scene& sc, int ch, const Ray& ray, double min_t, double max_t, Hit* out,
                          Counters* c);

static bool bvh_node_hit(const Scene& sc, int ni, const Ray& ray, double min_t, double max_t, Hit* out,
                         Counters* c) {
    const BvhNode& nd = sc.bvh[ni];
    if (c) c->aabb_tests++;
    if (!aabb_ray_hit(nd.bb_min, nd.bb_max, ray, min_t, max_t)) return false;  // :628-634
    // ray_intersect, :636-651
    Hit lh;
    bool left = bvh_child_hit(sc, nd.left, ray, min_t, max_t, &lh, c);
    if (nd.right != NO_CHILD) {
        if (left) {
            Hit rh;
            if (bvh_child_hit(sc, nd.right, ray, min_t, lh.distance, &rh, c)) *out = rh;
            else *out = lh;
            return true;
        }
        return bvh_child_hit(sc, nd.right, ray, min_t, max_t, out, c);
    }
    if (left) *out = lh;
    return left;
}

static bool bvh_child_hit(const Scene& sc, int ch, const Ray& ray, double min_t, double max_t, Hit* out,
                          Counters* c) {
    if (ch >= 0) return bvh_node_hit(sc, ch, ray, min_t, max_t, out, c);
    return shape_ray_hit(sc.shapes[~ch], ~ch, ray, min_t, max_t, out, c);
}

bool bvh_ray_hit(const Scene& sc, const Ray& ray, double min_t, double max_t, Hit* out, Counters* c) {
    if (sc.bvh_root < 0) return false;
    return bvh_node_hit(sc, sc.bvh_root, ray, min_t, max_t, out, c);
}

// ---------------------------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------------------------
void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    // Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32, 10 rounds
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

void PathRng::seed_xoshiro(uint64_t seed, uint64_t stream) {
    uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (int i = 0; i < 4; i++) s[i] = splitmix64(x);
}

double PathRng::next() {
    if (mode == RNG_PHILOX) {
        uint32_t i = next_index++;
        if ((i & 1u) == 0) {
            uint32_t ctr[4] = {pixel, sample, event, i >> 1}, out[4];
            philox4x32_10(ctr, key, out);
            uint64_t a = ((uint64_t)out[1] << 32) | out[0];
            uint64_t b = ((uint64_t)out[3] << 32) | out[2];
            cache[0] = (double)(a >> 11) * (1.0 / 9007199254740992.0);
            cache[1] = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        }
        return cache[i & 1u];
    }
    // xoshiro256** (Blackman & Vigna)
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return (double)(result >> 11) * (1.0 / 9007199254740992.0);
}

V3 random_range(PathRng& r, double mn, double mx) {  // algebra/mod.rs:59-66 (gen_range(min..=max) x3)
    double x = mn + (mx - mn) * r.next();
    double y = mn + (mx - mn) * r.next();
    double z = mn + (mx - mn) * r.next();
    return v3(x, y, z);
}
V3 random_in_unit_sphere(PathRng& r) {  // :77-84
    for (;;) {
        V3 v = random_range(r, -1.0, 1.0);
        if (squared_length(v) <= 1.0) return v;
    }
}
V3 random_unit(PathRng& r) { return normalize(random_in_unit_sphere(r)); }  // :86-88

// ---------------------------------------------------------------------------------------------
// textures / materials
// ---------------------------------------------------------------------------------------------
// Perlin::noise — src/algebra/noise.rs:43-73
double perlin_noise(const rt_perlin& pn, V3 p) {
    auto as_i32 = [](double v) -> int32_t {  // Rust `as i32`: saturating, NaN -> 0
        if (std::isnan(v)) return 0;
        if (v <= -2147483648.0) return INT32_MIN;
        if (v >= 2147483647.0) return INT32_MAX;
        return (int32_t)v;
    };
    int32_t x = as_i32(std::floor(p.x)), y = as_i32(std::floor(p.y)), z = as_i32(std::floor(p.z));
    double u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
    double u2 = u * u * (3.0 - 2.0 * u);
    double v2 = v * v * (3.0 - 2.0 * v);
    double w2 = w * w * (3.0 - 2.0 * w);
    double acc = 0.0;  // Iterator::sum::<f64>() folds from 0.0
    for (int i = 0; i < 2; i++)          // multi_cartesian_product of three 0..2 ranges: last one fastest
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                uint32_t ix = (uint32_t)(i + (int64_t)x) & 255u, iy = (uint32_t)(j + (int64_t)y) & 255u,
                         iz = (uint32_t)(k + (int64_t)z) & 255u;
                const rt_vec3& c = pn.ranvec[pn.perm_x[ix] ^ pn.perm_y[iy] ^ pn.perm_z[iz]];
                double fi = (double)i, fj = (double)j, fk = (double)k;
                double dotp = c.x * (u - fi) + c.y * (v - fj) + c.z * (w - fk);
                double term = (fi * u2 + (double)(1 - i) * (1.0 - u2)) * (fj * v2 + (double)(1 - j) * (1.0 - v2)) *
                              (fk * w2 + (double)(1 - k) * (1.0 - w2)) * dotp;
                acc = acc + term;
            }
    return acc;
}
// Perlin::turb — :75-86.  `self.noise(&p)`: the scan never uses temp_p, every octave samples p itself.
double perlin_turb(const rt_perlin& pn, V3 p, int depth) {
    double weight = 1.0, acc = 0.0;
    for (int i = 0; i < depth; i++) {
        double ret = weight * perlin_noise(pn, p);
        weight *= 0.5;
        acc = acc + ret;
    }
    return std::fabs(acc);
}

V3 texture_value(const Scene& sc, uint32_t tex, double u, double v, V3 p) {
    for (int depth = 0; depth <= RT_TEX_MAX_DEPTH; depth++) {
        const rt_texture& t = sc.textures[tex];
        switch (t.kind) {
            case RT_TEX_SOLID:  // texture.rs:16-20
                return from_rt(t.color);
            case RT_TEX_CHECKER: {  // :40-51
                double sines = std::sin(t.color.x * p.x) * std::sin(t.color.y * p.y) * std::sin(t.color.z * p.z);
                tex = sines < 0.0 ? t.odd : t.even;
                break;
            }
            case RT_TEX_UV_CHECKER: {  // :78-88 (note v pairs with multipliers.0)
                double sines = std::sin(v * t.color.x * PI) * std::sin(u * t.color.y * PI);
                tex = sines < 0.0 ? t.odd : t.even;
                break;
            }
            case RT_TEX_IMAGE: {  // :98-117
                const Scene::Image& im = sc.images[t.image];
                double uu = std::isnan(u) ? u : std::fmin(std::fmax(u, 0.0), 1.0);  // f64::clamp keeps NaN
                double vc = std::isnan(v) ? v : std::fmin(std::fmax(v, 0.0), 1.0);
                double vv = 1.0 - vc;
                double fx = uu * (double)im.w, fy = vv * (double)im.h;
                // `as u32` saturates and maps NaN to 0; get_pixel would panic at x == width (u == 1):
                // the port clamps to the last texel instead (documented deviation)
                uint32_t x = std::isnan(fx) ? 0u : (fx <= 0.0 ? 0u : (fx >= 4294967295.0 ? 4294967295u : (uint32_t)fx));
                uint32_t y = std::isnan(fy) ? 0u : (fy <= 0.0 ? 0u : (fy >= 4294967295.0 ? 4294967295u : (uint32_t)fy));
                if (x >= im.w) x = im.w - 1;
                if (y >= im.h) y = im.h - 1;
                const uint8_t* px = &im.rgba[((size_t)y * im.w + x) * 4];
                double color_scale = 1.0 / 255.0;
                return v3((double)px[0] * color_scale, (double)px[1] * color_scale, (double)px[2] * color_scale);
            }
            case RT_TEX_NOISE: {  // texture.rs:61-67
                double k = 0.5 * (1.0 + std::sin(t.color.x * p.z + 10.0 * perlin_turb(sc.noise[t.image], p, 7)));
                return v3(k * 1.0, k * 1.0, k * 1.0);
            }
            default:
                return v3(0.0, 0.0, 0.0);
        }
    }
    return v3(0.0, 0.0, 0.0);
}

static double reflectance(double cosine, double ref_index) {  // material.rs:84-88
    double r0 = (1.0 - ref_index) / (1.0 + ref_index);
    r0 = r0 * r0;
    double x = 1.0 - cosine;
    double x2 = x * x;            // powi(5) lowers to x * ((x*x)*(x*x))
    double x5 = x * (x2 * x2);
    return r0 + (1.0 - r0) * x5;
}

bool material_scatter(const Scene& sc, const rt_material& m, const Ray& ray, const Hit& hit, PathRng& rng,
                      Ray* scattered, V3* attenuation) {
    switch (m.kind) {
        case RT_MAT_LAMBERTIAN: {  // material.rs:42-53
            V3 direction = hit.normal + random_unit(rng);
            if (is_zero(direction)) direction = hit.normal;
            *scattered = ray_new(hit.point, direction);
            *attenuation = texture_value(sc, m.texture, hit.u, hit.v, hit.point);
            return true;
        }
        case RT_MAT_METAL: {  // :64-75
            V3 reflected = reflect(ray.direction, hit.normal);
            V3 direction = (m.scalar == 0.0) ? reflected : reflected + m.scalar * random_in_unit_sphere(rng);
            *scattered = ray_new(hit.point, direction);
            *attenuation = texture_value(sc, m.texture, hit.u, hit.v, hit.point);
            return true;
        }
        case RT_MAT_DIELECTRIC: {  // :93-115
            double refract_ratio = hit.is_front_face ? 1.0 / m.scalar : m.scalar;
            double cos_theta = dot(-ray.direction, hit.normal);
            double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
            V3 direction;
            // `||` short-circuits: the uniform is only drawn when total internal reflection is excluded
            if (refract_ratio * sin_theta > 1.0 || reflectance(cos_theta, refract_ratio) > rng.next())
                direction = reflect(ray.direction, hit.normal);
            else
                direction = refract(ray.direction, hit.normal, refract_ratio);
            *scattered = ray_new(hit.point, direction);
            *attenuation = v3(1.0, 1.0, 1.0);
            return true;
        }
        default:  // DiffuseLight, EmptyMaterial: Material::scatter default, :24-26
            return false;
    }
}

V3 material_emitted(const Scene& sc, const rt_material& m, double u, double v, V3 p) {
    if (m.kind == RT_MAT_DIFFUSE_LIGHT) return texture_value(sc, m.texture, u, v, p);  // :123-127
    return v3(0.0, 0.0, 0.0);                                                         // :28-30
}

V3 background(const Ray& ray) {  // world/mod.rs:199-202
    double t = 0.5 * (ray.direction.y + 1.0);
    return (1.0 - t) * v3(1.0, 1.0, 1.0) + t * v3(0.5, 0.7, 1.0);
}

V3 ray_color(const Scene& sc, bool use_bvh, const Ray& ray, uint32_t depth, PathRng& rng,
             uint32_t hit_number, Counters* c) {
    // renderer/mod.rs:23-45
    Hit hit;
    if (c) c->segments++;
    bool found = use_bvh ? bvh_ray_hit(sc, ray, 0.001, INFINITY, &hit, c)
                         : collection_ray_intersect(sc, ray, 0.001, INFINITY, &hit, c);
    if (found) {
        if (depth == 0) return v3(0.0, 0.0, 0.0);
        const rt_material& m = sc.materials[sc.shapes[hit.shape].material];
        Ray scattered;
        V3 attenuation;
        rng.begin_event(hit_number + 1);
        if (material_scatter(sc, m, ray, hit, rng, &scattered, &attenuation))
            return product(attenuation, ray_color(sc, use_bvh, scattered, depth - 1, rng, hit_number + 1, c));
        return material_emitted(sc, m, hit.u, hit.v, hit.point);
    }
    return background(ray);
}

// ---------------------------------------------------------------------------------------------
// threaded renderer — renderer/mod.rs:66-155 + renderer/step_by_step.rs:37-121
// ---------------------------------------------------------------------------------------------
namespace {
struct PixelRays {
    uint32_t index;
    std::vector<Ray> rays;
};
typedef std::vector<PixelRays> InputChunk;                     // InputDataVec
typedef std::vector<std::pair<uint32_t, V3>> OutputChunk;      // OutputDataVec

template <class T>
struct Channel {  // std::sync::mpsc channel behind Arc<Mutex<..>>
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::pair<bool, T>> q;  // (is_some, value)
    void send(bool some, T&& v) {
        {
            std::lock_guard<std::mutex> l(mu);
            q.emplace_back(some, std::move(v));
        }
        cv.notify_one();
    }
    std::pair<bool, T> recv() {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return !q.empty(); });
        auto r = std::move(q.front());
        q.pop_front();
        return r;
    }
};
}  // namespace

double render(const Scene& sc, const rt_camera& cam, const RenderOptions& opt, rt_vec3* buffer,
              Counters* counters) {
    const uint32_t width = opt.image.width, height = opt.image.height;
    const uint32_t threads = opt.threads ? opt.threads : 1;
    const uint32_t sx = opt.stride_x ? opt.stride_x : 1, sy = opt.stride_y ? opt.stride_y : 1;
    Channel<InputChunk> input;
    Channel<OutputChunk> output;
    std::vector<Counters> per_thread(threads);

    auto t0 = std::chrono::steady_clock::now();

    // new_dispatcher_thread, renderer/mod.rs:66-90: ONE thread generates every primary ray
    std::thread dispatcher([&] {
        // chunk = w*h/threads/8 pixels (renderer/mod.rs:74); for a strided (bounded) sample the same
        // rule is applied to the pixels actually traced, so the workers stay equally loaded
        uint32_t traced = ((width + sx - 1) / sx) * ((height + sy - 1) / sy);
        size_t chunk_size = (size_t)(traced / threads / 8);
        if (chunk_size == 0) chunk_size = 1;
        RayCaster rc = raycaster_new(cam, opt.image);
        PathRng rng;
        rng.mode = opt.rng;
        rng.key[0] = (uint32_t)opt.seed;
        rng.key[1] = (uint32_t)(opt.seed >> 32);
        if (opt.rng == RNG_XOSHIRO) rng.seed_xoshiro(opt.seed, 0);
        InputChunk chunk;
        for (uint32_t y = 0; y < height; y++) {          // cartesian_product(0..height, 0..width)
            if (y % sy) continue;
            for (uint32_t x = 0; x < width; x++) {
                if (x % sx) continue;
                PixelRays pr;
                pr.index = x + y * width;
                pr.rays.reserve(opt.samples_number);
                for (uint32_t s = 0; s < opt.samples_number; s++) {   // ray_caster.rs:103-115
                    rng.pixel = pr.index;
                    rng.sample = s;
                    rng.begin_event(0);
                    double u = rng.next();
                    double v = rng.next();
                    pr.rays.push_back(raycaster_get_ray(rc, (double)x + u, (double)y + v));
                }
                chunk.push_back(std::move(pr));
                if (chunk.size() == chunk_size) {
                    input.send(true, std::move(chunk));
                    chunk = InputChunk();
                }
            }
        }
        if (!chunk.empty()) input.send(true, std::move(chunk));
        for (uint32_t i = 0; i < threads; i++) input.send(false, InputChunk());
    });

    // new_worker_thread, renderer/mod.rs:92-125
    std::vector<std::thread> workers;
    for (uint32_t tid = 0; tid < threads; tid++) {
        workers.emplace_back([&, tid] {
            Counters* c = counters ? &per_thread[tid] : nullptr;
            for (;;) {
                auto msg = input.recv();
                if (!msg.first) {
                    output.send(false, OutputChunk());
                    return;  // the reference parks on a Condvar here
                }
                OutputChunk result;  // trace_pixel_samples_group, :127-149
                result.reserve(msg.second.size());
                for (const PixelRays& pr : msg.second) {  // trace_pixel_samples, :151-155
                    PathRng rng;
                    rng.mode = opt.rng;
                    rng.key[0] = (uint32_t)opt.seed;
                    rng.key[1] = (uint32_t)(opt.seed >> 32);
                    rng.pixel = pr.index;
                    if (opt.rng == RNG_XOSHIRO) rng.seed_xoshiro(opt.seed, (uint64_t)pr.index + 1);
                    V3 acc = v3(0.0, 0.0, 0.0);
                    for (size_t s = 0; s < pr.rays.size(); s++) {
                        rng.sample = (uint32_t)s;
                        V3 col = ray_color(sc, opt.use_bvh, pr.rays[s], opt.max_depth, rng, 0, c);
                        acc.x += col.x; acc.y += col.y; acc.z += col.z;
                    }
                    double ln = (double)pr.rays.size();
                    result.emplace_back(pr.index, acc / ln);
                }
                output.send(true, std::move(result));
            }
        });
    }

    // render_step loop, step_by_step.rs:101-121 (blocking here; the reference polls per UI frame)
    uint32_t num_finished = 0;
    while (num_finished < threads) {
        auto msg = output.recv();
        if (!msg.first) {
            num_finished++;
            continue;
        }
        for (auto& pc : msg.second) buffer[pc.first] = rt_vec3{pc.second.x, pc.second.y, pc.second.z};
    }
    auto t1 = std::chrono::steady_clock::now();
    dispatcher.join();
    for (auto& w : workers) w.join();
    if (counters)
        for (auto& c : per_thread) counters->add(c);
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace orc
