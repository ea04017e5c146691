// Device-side FP64 arithmetic of the path-tracing core.
//
// The parity contract (bit-exact nearest-hit index, t and normal identical to the reference's f64
// code) holds only if every expression is evaluated with the reference's operand order and WITHOUT
// fused multiply-add: this translation unit must be compiled with -fmad=false.  Where an FMA is
// wanted (conservative culling, never the exact tests) it is written explicitly as fma()/__fmaf_rn.
// Citations are file:line into dkarpushkin/rs-pathtracing.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rt_b200.h"

namespace rt {

// algebra::Vector3d, src/algebra/mod.rs:23-28
struct D3 {
    double x, y, z;
};
__host__ __device__ __forceinline__ D3 mk(double x, double y, double z) { return D3{x, y, z}; }
__host__ __device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a) { return {-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ D3 operator*(double s, D3 a) { return {a.x * s, a.y * s, a.z * s}; }
// Vector / scalar (mod.rs:299-317: three IEEE divisions by the same divisor).  nvcc's inline division takes a
// ~100-instruction slow path whenever the numerator is zero -- two components of every axis-aligned wall normal
// -- which made `normalize` a quarter of k_shade.  div3_exact returns the same three correctly rounded
// quotients from ONE correctly rounded reciprocal y = RN(1 / s) (Markstein: with r = a - s q computed exactly
// by an FMA, q' = RN(q + r y) is RN(a / s) as soon as q is within one ulp of a / s; the first correction makes
// q = RN(a y) faithful, the second one correctly rounded).  The FMAs are explicit, so -fmad=false does not
// touch them.  Operands outside [2^-500, 2^500] (where r or q could leave the normal range) and s == 0 / non-finite s
// take the plain divisions.  A negative divisor is handled as (-a) / |s| -- round-to-nearest is symmetric, and the
// sign of a zero quotient comes out as IEEE's.  Host build: rt_div3_exact, tests/test_div3_exact.py.
// div2_exact: the same for two numerators (Cube::ray_intersect divides -1 - o and 1 - o by the same d component,
// the marching bounds o and d by the same radius: `divide` was 12 % of k_extend's instructions).
__host__ __device__ __forceinline__ int fp_exponent_field(double v) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(v) >> 20) & 0x7ff;
#else
    unsigned long long b;
    memcpy(&b, &v, sizeof b);
    return (int)((b >> 52) & 0x7ff);
#endif
}
__host__ __device__ __forceinline__ bool div3_operand_ok(double v) {
    return (unsigned)(fp_exponent_field(v) - 523) <= 1000u || v == 0.0;
}
__host__ __device__ __forceinline__ double quotient_by_rcp(double a, double s, double y) {
    double q = a * y;
    double r = fma(-s, q, a);
    q = fma(r, y, q);
    r = fma(-s, q, a);
    q = fma(r, y, q);
    return a == 0.0 ? a : q;
}
__host__ __device__ __forceinline__ void div3_exact(double ax, double ay, double az, double s, double& qx, double& qy,
                                                    double& qz) {
    if ((unsigned)(fp_exponent_field(s) - 523) <= 1000u && div3_operand_ok(ax) && div3_operand_ok(ay) &&
        div3_operand_ok(az)) {
        const double sa = fabs(s);
        const bool neg = s < 0.0;
#ifdef __CUDA_ARCH__
        const double y = __drcp_rn(sa);
#else
        const double y = 1.0 / sa;
#endif
        qx = quotient_by_rcp(neg ? -ax : ax, sa, y);
        qy = quotient_by_rcp(neg ? -ay : ay, sa, y);
        qz = quotient_by_rcp(neg ? -az : az, sa, y);
        return;
    }
    qx = ax / s;
    qy = ay / s;
    qz = az / s;
}
__host__ __device__ __forceinline__ void div2_exact(double a, double b, double s, double& qa, double& qb) {
    if ((unsigned)(fp_exponent_field(s) - 523) <= 1000u && div3_operand_ok(a) && div3_operand_ok(b)) {
        const double sa = fabs(s);
        const bool neg = s < 0.0;
#ifdef __CUDA_ARCH__
        const double y = __drcp_rn(sa);
#else
        const double y = 1.0 / sa;
#endif
        qa = quotient_by_rcp(neg ? -a : a, sa, y);
        qb = quotient_by_rcp(neg ? -b : b, sa, y);
        return;
    }
    qa = a / s;
    qb = b / s;
}
__device__ __forceinline__ D3 operator/(D3 a, double s) {
    D3 q;
    div3_exact(a.x, a.y, a.z, s, q.x, q.y, q.z);
    return q;
}
// `*` between vectors is the dot product: (x*x' + y*y') + z*z'   (mod.rs:319-349)
__host__ __device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 hadamard(D3 a, D3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // product, :135-141
__device__ __forceinline__ D3 divide(D3 a, D3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }    // :143-150
__device__ __forceinline__ double length(D3 a) { return sqrt(dot(a, a)); }                       // :117-120
__device__ __forceinline__ D3 normalize(D3 a) { return a / length(a); }                          // :107-110
__host__ __device__ __forceinline__ bool approx_zero(double a) { return fabs(a - 0.0) < 1e-15; }          // approx_equal(a, 0.0), :14-17
// f64::min/max ignore a NaN operand; so do fmin/fmax (:168-198)
__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }
__device__ __forceinline__ double min3(double a, double b, double c) { return fmin(fmin(a, b), c); }

__device__ __forceinline__ D3 reflect(D3 v, D3 n) {  // :122-125
    D3 b = dot(v, n) * n;
    return v - (2.0 * b);
}
__device__ __forceinline__ D3 refract(D3 v, D3 n, double ratio) {  // :127-133
    double cos_theta = dot(-v, n);
    D3 r_out_perp = ratio * (v + cos_theta * n);
    D3 r_out_parallel = -(sqrt(fabs(1.0 - dot(r_out_perp, r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}

// Transform::transform_point / transform_vector / transform_normal on rows 0..2 of the 4x4
// (src/algebra/transform.rs:394-425).  `m` is 12 doubles, row-major 3x4.
__host__ __device__ __forceinline__ D3 xf_point(const double* m, D3 p) {
    return {p.x * m[0] + p.y * m[1] + p.z * m[2] + m[3], p.x * m[4] + p.y * m[5] + p.z * m[6] + m[7],
            p.x * m[8] + p.y * m[9] + p.z * m[10] + m[11]};
}
__host__ __device__ __forceinline__ D3 xf_vector(const double* m, D3 v) {
    return {v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[4] + v.y * m[5] + v.z * m[6],
            v.x * m[8] + v.y * m[9] + v.z * m[10]};
}
__device__ __forceinline__ D3 xf_normal(const double* m, D3 n) {  // transpose of the 3x3
    return {n.x * m[0] + n.y * m[4] + n.z * m[8], n.x * m[1] + n.y * m[5] + n.z * m[9],
            n.x * m[2] + n.y * m[6] + n.z * m[10]};
}

// ------------------------------------------------------------------------------------------------
// implicit surfaces — src/world/shapes/ray_marching.rs:134-520.  KIND is a compile-time constant so
// that the marching loop contains exactly one polynomial; q = params[8] of the shape.  The scalar
// type is a template parameter: T = double is the reference's arithmetic (same operand order, no
// FMA), T = Dual gives the derivative along the ray for the safe-skip bound, and the host evaluates
// the same text with interval jets to bound the second derivatives (march_bounds.hpp).
// ------------------------------------------------------------------------------------------------
template <int KIND, typename T>
__host__ __device__ __forceinline__ auto surface_func_t(const double* q, T px, T py, T pz) {
    // `auto` everywhere: with T = a degree-tracking polynomial (rt_march.cuh) every product has its own type
    if constexpr (KIND == RT_SURF_HEART) {  // :147-155
        auto x2 = px * px;
        auto y2 = py * py;
        auto z2 = pz * pz;
        auto z3 = z2 * pz;
        auto a = x2 + (9.0 / 4.0) * y2 + z2 - 1.0;
        return a * a * a - x2 * z3 - (9.0 / 80.0) * y2 * z3;
    } else if constexpr (KIND == RT_SURF_SINE) {  // :203-211
        double a_ = q[3];
        return a_ * a_ * (px - py - pz) * (px + py - pz) * (px - py + pz) * (px + py + pz) +
               4.0 * px * px * py * py * pz * pz;
    } else if constexpr (KIND == RT_SURF_STAR) {  // :268-274
        double a_ = q[3];
        auto x2 = px * px;
        auto y2 = py * py;
        auto z2 = pz * pz;
        auto c = x2 + y2 + z2 - 1.0;
        return a_ * (x2 * y2 + x2 * z2 + y2 * z2) + (c * c * c);
    } else if constexpr (KIND == RT_SURF_DUPIN) {  // :340-345
        double a_ = q[3], b_ = q[4], c_ = q[5], d_ = q[6];
        double b2 = b_ * b_;
        auto e = px * px + py * py + pz * pz + b2 - d_ * d_;
        auto f = a_ * px - c_ * d_;
        return e * e - 4.0 * (f * f + b2 * py * py);
    } else if constexpr (KIND == RT_SURF_HUNTS) {  // :399-406
        auto x2 = px * px;
        auto y2 = py * py;
        auto z2 = pz * pz;
        auto a = x2 + y2 + z2 - 13.0;
        auto b = 3.0 * x2 + y2 - 4.0 * z2 - 12.0;
        return 4.0 * a * a * a + 27.0 * b * b;
    } else {  // RT_SURF_CUSHION, :464-478
        auto x2 = px * px;
        auto y2 = py * py;
        auto z2 = pz * pz;
        auto a = x2 - pz;
        return z2 * x2 - z2 * z2 - 2.0 * pz * x2 + 2.0 * pz * z2 + x2 - z2 - a * a - y2 * y2 - 2.0 * x2 * y2 -
               y2 * z2 + 2.0 * y2 * pz + y2;
    }
}

template <int KIND>
__host__ __device__ __forceinline__ double surface_func(const double* q, D3 p) {
    return surface_func_t<KIND, double>(q, p.x, p.y, p.z);
}

// value + derivative along the ray (forward-mode AD); products use the plain product rule
struct Dual {
    double v, d;
};
__host__ __device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__host__ __device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__host__ __device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__host__ __device__ __forceinline__ Dual operator+(Dual a, double b) { return {a.v + b, a.d}; }
__host__ __device__ __forceinline__ Dual operator-(Dual a, double b) { return {a.v - b, a.d}; }
__host__ __device__ __forceinline__ Dual operator*(Dual a, double b) { return {a.v * b, a.d * b}; }
__host__ __device__ __forceinline__ Dual operator*(double a, Dual b) { return {a * b.v, a * b.d}; }
__host__ __device__ __forceinline__ Dual operator+(double a, Dual b) { return {a + b.v, b.d}; }
__host__ __device__ __forceinline__ Dual operator-(double a, Dual b) { return {a - b.v, -b.d}; }

__device__ inline D3 surface_gradient(const double* q, D3 p) {
    switch ((int)q[0]) {
        case RT_SURF_HEART: {  // :157-168
            double a = p.x * p.x + (9.0 / 4.0) * p.y * p.y + p.z * p.z - 1.0;
            a = 3.0 * a * a;
            double z2 = p.z * p.z;
            double z3 = z2 * p.z;
            return mk(2.0 * p.x * (a - z3), (9.0 / 2.0) * p.y * (a - 0.05 * z3),
                      2.0 * p.z * (a - p.z * (1.5 * p.x * p.x + (27.0 / 40.0) * p.y * p.y)));
        }
        case RT_SURF_SINE: {  // :227-237
            double a_ = q[3];
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a2 = a_ * a_;
            return mk(4.0 * p.x * (a2 * (x2 - y2 - z2) + 2.0 * y2 * z2),
                      8.0 * x2 * p.y * z2 - 4.0 * a2 * p.y * (x2 - y2 + z2),
                      8.0 * x2 * y2 * p.z - 4.0 * a2 * p.z * (x2 + y2 - z2));
        }
        case RT_SURF_STAR: {  // :290-300
            double a_ = q[3];
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double c = x2 + y2 + z2 - 1.0;
            return mk(2.0 * a_ * p.x * (y2 + z2) + 6.0 * p.x * c * c, 2.0 * a_ * p.y * (x2 + z2) + 6.0 * p.y * c * c,
                      2.0 * a_ * p.z * (x2 + y2) + 6.0 * p.z * c * c);
        }
        case RT_SURF_DUPIN: {  // :361-369
            double a_ = q[3], b_ = q[4], c_ = q[5], d_ = q[6];
            double b2 = b_ * b_;
            double e = 4.0 * (p.x * p.x + p.y * p.y + p.z * p.z + b2 - d_ * d_);
            return mk(e * p.x - 8.0 * a_ * (a_ * p.x - c_ * d_), e * p.y - 8.0 * b2 * p.y, e * p.z);
        }
        case RT_SURF_HUNTS: {  // :422-434
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            double a = x2 + y2 + z2 - 13.0;
            double b = 3.0 * x2 + y2 - 4.0 * (z2 + 3.0);
            return mk(24.0 * p.x * a * a + 324.0 * p.x * b, 12.0 * p.y * (2.0 * a * a + 9.0 * b),
                      24.0 * p.z * (a * a - 18.0 * b));
        }
        default: {  // RT_SURF_CUSHION, :494-504
            double x2 = p.x * p.x;
            double y2 = p.y * p.y;
            double z2 = p.z * p.z;
            return mk(2.0 * p.x * (-2.0 * x2 - 2.0 * y2 + z2 + 1.0),
                      -2.0 * p.y * (2.0 * x2 + 2.0 * y2 + z2 - 2.0 * p.z - 1.0),
                      2.0 * p.z * (x2 - 2.0 * z2 + 3.0 * p.z - 2.0) - 2.0 * p.y * (p.z - 1.0));
        }
    }
}

#define RT_MARCH_BUDGET 200000000ull

// ---- Torus (src/world/shapes/mod.rs:403-494) on solve_quantic_equation (src/algebra/equation.rs:17-67) ----------
// num::Complex<f64> arithmetic as the crate `num ^0.4` (num-complex 0.4.x) defines it: +, -, scalar * and /
// componentwise; complex * = (re re' - im im', re im' + im re'); complex / through norm_sqr; sqrt / cbrt by cases
// (im == 0, re == 0, else the polar form), signs of zero respected.  hypot / atan2 / cos / sin / cbrt are CUDA's
// (1-2 ulp): the roots agree with a glibc build to ~1e-15 relative, the `|im| < 1e-15` acceptance test of the
// caller does not always (Torus).
struct Cx {
    double re, im;
};
__device__ __forceinline__ Cx cxr(double re) { return Cx{re, 0.0}; }
__device__ __forceinline__ Cx operator+(Cx a, Cx b) { return Cx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cx operator-(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cx operator-(Cx a) { return Cx{-a.re, -a.im}; }
__device__ __forceinline__ Cx operator*(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx operator/(Cx a, Cx b) {
    const double norm_sqr = b.re * b.re + b.im * b.im;
    const double re = a.re * b.re + a.im * b.im;
    const double im = a.im * b.re - a.re * b.im;
    return Cx{re / norm_sqr, im / norm_sqr};
}
__device__ __forceinline__ Cx operator*(double k, Cx a) { return Cx{k * a.re, k * a.im}; }
__device__ __forceinline__ Cx operator/(Cx a, double k) { return Cx{a.re / k, a.im / k}; }
__device__ __forceinline__ Cx cx_from_polar(double r, double theta) { return Cx{r * cos(theta), r * sin(theta)}; }
static __device__ __noinline__ Cx cx_sqrt(Cx z) {
    if (z.im == 0.0) {
        if (!signbit(z.re)) return Cx{sqrt(z.re), z.im};
        const double im = sqrt(-z.re);
        return !signbit(z.im) ? Cx{0.0, im} : Cx{0.0, -im};
    }
    if (z.re == 0.0) {
        const double x = sqrt(fabs(z.im) / 2.0);
        return !signbit(z.im) ? Cx{x, x} : Cx{x, -x};
    }
    return cx_from_polar(sqrt(hypot(z.re, z.im)), atan2(z.im, z.re) / 2.0);
}
static __device__ __noinline__ Cx cx_cbrt(Cx z) {
    if (z.im == 0.0) {
        if (!signbit(z.re)) return Cx{cbrt(z.re), z.im};
        const double re = cbrt(-z.re) / 2.0;
        const double im = sqrt(3.0) * re;
        return !signbit(z.im) ? Cx{re, im} : Cx{re, -im};
    }
    if (z.re == 0.0) {
        const double im = cbrt(fabs(z.im)) / 2.0;
        const double re = sqrt(3.0) * im;
        return !signbit(z.im) ? Cx{re, im} : Cx{re, -im};
    }
    return cx_from_polar(cbrt(hypot(z.re, z.im)), atan2(z.im, z.re) / 3.0);
}
// equation.rs:17-67
static __device__ __noinline__ void solve_quartic(double a_, double b_, double c_, double d_, double e_, Cx roots[4]) {
    const Cx a = cxr(a_), b = cxr(b_) / a, c = cxr(c_) / a, d = cxr(d_) / a, e = cxr(e_) / a;
    const Cx b2 = b * b;
    const Cx alpha = c - (3.0 / 8.0) * b2;
    const Cx beta = (b2 * b) / 8.0 - (b * c) / 2.0 + d;
    const Cx gamma = (-3.0 / 256.0) * b2 * b2 + b2 * c / 16.0 - b * d / 4.0 + e;
    const Cx alpha2 = alpha * alpha;
    const Cx t = -b / 4.0;
    if (approx_zero(beta.re) && approx_zero(beta.im)) {
        const Cx r = cx_sqrt(alpha2 - 4.0 * gamma);
        const Cx r1 = cx_sqrt((-alpha + r) / 2.0);
        const Cx r2 = cx_sqrt((-alpha - r) / 2.0);
        roots[0] = t + r1; roots[1] = t - r1; roots[2] = t + r2; roots[3] = t - r2;
    } else {
        const Cx p = -(alpha2 / 12.0 + gamma);
        const Cx q = -alpha2 * alpha / 108.0 + alpha * gamma / 3.0 - beta * beta / 8.0;
        const Cx r = -q / 2.0 + cx_sqrt(q * q / 4.0 + p * p * p / 27.0);
        const Cx u = cx_cbrt(r);
        Cx y = (-5.0 / 6.0) * alpha + u;
        if (approx_zero(u.re) && approx_zero(u.im)) y = y - cx_cbrt(q);
        else y = y - p / (3.0 * u);
        const Cx w = cx_sqrt(alpha + 2.0 * y);
        const Cx r1 = cx_sqrt(-(3.0 * alpha + 2.0 * y + 2.0 * beta / w));
        const Cx r2 = cx_sqrt(-(3.0 * alpha + 2.0 * y - 2.0 * beta / w));
        roots[0] = t + (w - r1) / 2.0; roots[1] = t + (w + r1) / 2.0;
        roots[2] = t + (-w - r2) / 2.0; roots[3] = t + (-w + r2) / 2.0;
    }
}
// Torus::ray_intersect, shapes/mod.rs:429-452: the candidate t (the smallest root accepted as real), range-checked
static __device__ __noinline__ bool torus_candidate(const double* q, D3 origin, D3 dir, double min_t, double max_t, double& t_out) {
    const double radius = q[0], tube_radius = q[1];
    const double t = 4.0 * radius * radius;
    const double g = t * (dir.x * dir.x + dir.y * dir.y);
    const double h = 2.0 * t * (origin.x * dir.x + origin.y * dir.y);
    const double i = t * (origin.x * origin.x + origin.y * origin.y);
    const double j = dot(dir, dir);
    const double k = 2.0 * dot(origin, dir);
    const double l = dot(origin, origin) + radius * radius - tube_radius * tube_radius;
    const double a = j * j;
    const double b = 2.0 * j * k;
    const double c = 2.0 * j * l + k * k - g;
    const double d = 2.0 * k * l - h;
    const double e = l * l - i;
    Cx roots[4];
    solve_quartic(a, b, c, d, e, roots);
    double min_root = INFINITY;
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (approx_zero(roots[r].im) && roots[r].re < min_root) min_root = roots[r].re;
    if (isinf(min_root) || min_root < min_t || min_root > max_t) return false;
    t_out = min_root;
    return true;
}

// solve_quadratic_equation, src/algebra/equation.rs:5-15
__host__ __device__ __forceinline__ bool solve_quadratic(double a, double half_b, double c, double& x1, double& x2) {
    double d = half_b * half_b - a * c;
    if (d < 0.0) return false;
    if (d == 0.0) {
        x1 = -half_b;
        x2 = -half_b;
        return true;
    }
    double d_sqrt = sqrt(d);  // NaN d falls through to here, exactly like the reference
    x1 = (-half_b - d_sqrt) / a;
    x2 = (-half_b + d_sqrt) / a;
    return true;
}

// ShapeFunction::intersect_bound — ray_marching.rs:135-145 (Heart: ellipsoid), :213-225 (sphere)
__host__ __device__ __forceinline__ bool march_bound(const double* q, D3 o, D3 d, double& start, double& end) {
    double x1, x2;
    if ((int)q[0] == RT_SURF_HEART) {
        const double sr = 1.45;  // Heart::new, :126-131
        D3 radius = mk(sr, sr / 2.05, sr);
        D3 os, ds;   // divide(o, radius), divide(d, radius): six IEEE quotients from three reciprocals
        div2_exact(o.x, d.x, radius.x, os.x, ds.x);
        div2_exact(o.y, d.y, radius.y, os.y, ds.y);
        div2_exact(o.z, d.z, radius.z, os.z, ds.z);
        if (!solve_quadratic(dot(ds, ds), dot(ds, os), dot(os, os) - 1.0, x1, x2)) return false;
    } else {
        double R = q[7];
        if (!solve_quadratic(dot(d, d), dot(d, o), dot(o, o) - R * R, x1, x2)) return false;
    }
    if (x1 < 0.0 && x2 < 0.0) return false;
    start = fmax(x1, 0.0);
    end = fmax(x2, 0.0);
    return true;
}

// RayMarchingShape::ray_intersect's loops, ray_marching.rs:27-57.  Returns the candidate t or
// false; `evals` counts shape_func evaluations (statistics only).
template <int KIND>
__device__ __forceinline__ bool march_loop(const double* q, D3 o, D3 d, double start, double end, double min_t,
                                           double max_t, double& t_out, unsigned long long& evals) {
    double step = q[1];
    int depth = (int)q[2];
    double t = start;
    D3 p = o + t * d;
    double r = surface_func<KIND>(q, p);
    unsigned long long n = 0;
    for (int it = 0; it < depth; it++) {
        bool finished = false;
        D3 sd = step * d;  // `step * dir` is loop-invariant until the step changes
        for (;;) {
            // n > RT_MARCH_BUDGET: t + step == t (step underflowed against t); the reference would spin
            // forever, a kernel must not: report a miss
            if (t > end || t < start || n > RT_MARCH_BUDGET) {
                evals += n;
                return false;
            }
            t += step;
            p.x += sd.x;
            p.y += sd.y;
            p.z += sd.z;
            double next = surface_func<KIND>(q, p);
            n++;
            if (approx_zero(next)) {
                finished = true;
                break;
            }
            if ((r < 0.0 && next > 0.0) || (r > 0.0 && next < 0.0)) {
                step *= -0.01;
                r = next;
                break;
            }
            r = next;
        }
        if (finished) break;
    }
    evals += n;
    if (t < min_t || t > max_t) return false;
    t_out = t;
    return true;
}

__device__ inline bool march_candidate(const double* q, D3 o, D3 d, double min_t, double max_t, double& t,
                                       unsigned long long& evals) {
    double start, end;
    if (!march_bound(q, o, d, start, end)) return false;
    switch ((int)q[0]) {
        case RT_SURF_HEART: return march_loop<RT_SURF_HEART>(q, o, d, start, end, min_t, max_t, t, evals);
        case RT_SURF_SINE: return march_loop<RT_SURF_SINE>(q, o, d, start, end, min_t, max_t, t, evals);
        case RT_SURF_STAR: return march_loop<RT_SURF_STAR>(q, o, d, start, end, min_t, max_t, t, evals);
        case RT_SURF_DUPIN: return march_loop<RT_SURF_DUPIN>(q, o, d, start, end, min_t, max_t, t, evals);
        case RT_SURF_HUNTS: return march_loop<RT_SURF_HUNTS>(q, o, d, start, end, min_t, max_t, t, evals);
        default: return march_loop<RT_SURF_CUSHION>(q, o, d, start, end, min_t, max_t, t, evals);
    }
}

// ------------------------------------------------------------------------------------------------
// analytic shapes, object space.  Each returns the candidate t (the rest of the RayHit is a pure
// function of (shape, t, ray) and is rebuilt for the winner only by finalize_hit()).
// ------------------------------------------------------------------------------------------------

// Sphere::ray_intersect, src/world/shapes/mod.rs:330-356
__device__ __forceinline__ bool sphere_candidate(D3 o, D3 d, double min_t, double max_t, double& t,
                                                 bool* degenerate = nullptr) {
    double a = dot(d, d);
    double half_b = dot(d, o);
    double c = dot(o, o) - 1.0;
    double disc = half_b * half_b - a * c;
    if (disc < 0.0) return false;
    if (disc == 0.0) {
        t = -half_b * a;  // sic (:343-344): multiplied, and accepted with no range check
        if (degenerate) *degenerate = true;
        return true;
    }
    double sq = sqrt(disc);
    double x = (-half_b - sq) / a;
    if (x < min_t || x > max_t) {
        x = (-half_b + sq) / a;
        if (x < min_t || x > max_t) return false;
    }
    t = x;
    return true;
}

// Cube::ray_intersect, :250-263
__device__ __forceinline__ bool cube_candidate(D3 o, D3 d, double min_t, double max_t, double& t) {
    // divide(mk(-1, -1, -1) - o, d) and divide(mk(1, 1, 1) - o, d): six IEEE quotients from three reciprocals
    D3 t_lower, t_upper;
    div2_exact(-1.0 - o.x, 1.0 - o.x, d.x, t_lower.x, t_upper.x);
    div2_exact(-1.0 - o.y, 1.0 - o.y, d.y, t_lower.y, t_upper.y);
    div2_exact(-1.0 - o.z, 1.0 - o.z, d.z, t_lower.z, t_upper.z);
    double t_box_min = fmax(max3(fmin(t_lower.x, t_upper.x), fmin(t_lower.y, t_upper.y), fmin(t_lower.z, t_upper.z)), min_t);
    double t_box_max = fmin(min3(fmax(t_lower.x, t_upper.x), fmax(t_lower.y, t_upper.y), fmax(t_lower.z, t_upper.z)), max_t);
    if (t_box_min > t_box_max || t_box_min > max_t) return false;
    // the reference panics when the hit point's largest |component| is NaN (:279-281); that can only
    // happen for a NaN ray, which the core reports as a miss
    D3 p = o + t_box_min * d;
    double m = max3(fabs(p.x), fabs(p.y), fabs(p.z));
    if (!(m == fabs(p.x) || m == fabs(p.y) || m == fabs(p.z))) return false;
    t = t_box_min;
    return true;
}

// Rectangle::ray_intersect, :181-190
__device__ __forceinline__ bool rect_candidate(const double* q, D3 o, D3 d, double min_t, double max_t, double& t) {
    double x = -o.z / d.z;
    if (x < min_t || x > max_t) return false;
    D3 p = o + d * x;
    if (p.x < q[0] || p.x > q[2] || p.y < q[1] || p.y > q[3]) return false;
    t = x;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  One stream of doubles per
// (seed; pixel, sample, event): event 0 = pixel jitter, event k+1 = scatter at the k-th hit.
// Double i of a stream is half (i & 1) of block (i >> 1); 53 bits each.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct PathRng {
    uint32_t k0, k1, pixel, sample, event, next_index;
    double cache1;
    __device__ __forceinline__ void begin_event(uint32_t ev) {
        event = ev;
        next_index = 0;
    }
    __device__ __forceinline__ double next() {
        uint32_t i = next_index++;
        if (i & 1u) return cache1;
        uint32_t o[4];
        philox4x32_10(pixel, sample, event, i >> 1, k0, k1, o);
        unsigned long long a = ((unsigned long long)o[1] << 32) | o[0];
        unsigned long long b = ((unsigned long long)o[3] << 32) | o[2];
        cache1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        return (double)(a >> 11) * (1.0 / 9007199254740992.0);
    }
};

// Vector3d::random(min, max), src/algebra/mod.rs:59-66 (x, y, z drawn in that order)
__device__ __forceinline__ D3 random_range(PathRng& r, double mn, double mx) {
    double x = mn + (mx - mn) * r.next();
    double y = mn + (mx - mn) * r.next();
    double z = mn + (mx - mn) * r.next();
    return mk(x, y, z);
}
// random_in_unit_sphere, :77-84 (rejection, accepts |v|^2 <= 1)
__device__ __forceinline__ D3 random_in_unit_sphere(PathRng& r) {
    for (;;) {
        D3 v = random_range(r, -1.0, 1.0);
        if (dot(v, v) <= 1.0) return v;
    }
}
__device__ __forceinline__ D3 random_unit(PathRng& r) { return normalize(random_in_unit_sphere(r)); }  // :86-88

// Try number k of random_in_unit_sphere on a fresh event stream, computed out of sequence: the candidate is
// doubles 3k, 3k+1, 3k+2 of the stream (blocks 3k/2 and 3k/2 + 1), mapped by random_range(-1, 1) -- the very
// values the sequential loop above would draw in its iteration k.
__device__ __forceinline__ D3 ball_try(uint32_t pixel, uint32_t sample, uint32_t event, uint32_t k, uint32_t k0,
                                       uint32_t k1) {
    const uint32_t i0 = 3u * k, b0 = i0 >> 1;
    uint32_t A[4], B[4];
    philox4x32_10(pixel, sample, event, b0, k0, k1, A);
    philox4x32_10(pixel, sample, event, b0 + 1u, k0, k1, B);
    const unsigned long long a_lo = ((unsigned long long)A[1] << 32) | A[0], a_hi = ((unsigned long long)A[3] << 32) | A[2];
    const unsigned long long b_lo = ((unsigned long long)B[1] << 32) | B[0], b_hi = ((unsigned long long)B[3] << 32) | B[2];
    const bool odd = (i0 & 1u) != 0u;
    const unsigned long long w0 = odd ? a_hi : a_lo, w1 = odd ? b_lo : a_hi, w2 = odd ? b_hi : b_lo;
    const double S53 = 1.0 / 9007199254740992.0;
    const double u0 = (double)(w0 >> 11) * S53, u1 = (double)(w1 >> 11) * S53, u2 = (double)(w2 >> 11) * S53;
    return mk(-1.0 + (1.0 - -1.0) * u0, -1.0 + (1.0 - -1.0) * u1, -1.0 + (1.0 - -1.0) * u2);
}

// random_in_unit_sphere for a whole warp.  The sequential rejection loop keeps a warp busy for its slowest lane
// (5.7 rounds on average for 32 lanes that need 1.9 tries each: a third of the lanes active, 38 % of k_shade's
// instructions).  Here every lane first evaluates its own try 0; after that ALL lanes work for the lanes still
// rejected -- with nr of them left, lane j evaluates try (next + j % T) of the (j / T)-th rejected lane,
// T = 32 / nr -- and each owner takes its first accepted candidate.  Same tries, same order of preference, same
// arithmetic: the result is the sequential loop's, bit for bit.  Must be called by all 32 lanes;
// s_id: 32 uint4 of shared memory private to the warp.  `event`, k0, k1 are warp-uniform.
__device__ __forceinline__ D3 coop_random_in_unit_sphere(bool need, uint32_t pixel, uint32_t sample, uint32_t event,
                                                         uint32_t k0, uint32_t k1, uint4* s_id) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    D3 v = mk(0.0, 0.0, 0.0);
    bool pending = need;
    if (need) {
        v = ball_try(pixel, sample, event, 0u, k0, k1);
        pending = !(dot(v, v) <= 1.0);
    }
    uint32_t next_try = 1u;
    unsigned rej = __ballot_sync(FULL, pending);
    while (rej) {
        const int nr = __popc(rej);
        const int T = 32 / nr;  // helpers per rejected lane
        const int rank = __popc(rej & ((1u << lane) - 1u));
        if (pending) s_id[rank] = make_uint4(pixel, sample, next_try, 0u);
        __syncwarp();
        const int orank = lane / T, off = lane - orank * T;
        D3 w = mk(0.0, 0.0, 0.0);
        bool acc = false;
        if (orank < nr) {
            const uint4 id = s_id[orank];
            w = ball_try(id.x, id.y, event, id.z + (uint32_t)off, k0, k1);
            acc = dot(w, w) <= 1.0;
        }
        const unsigned accm = __ballot_sync(FULL, acc);
        __syncwarp();  // the reads of s_id are done before the next round overwrites it
        int src = lane;
        bool got = false;
        if (pending) {
            const unsigned mine = (accm >> (rank * T)) & (T == 32 ? FULL : ((1u << T) - 1u));
            if (mine) {
                src = rank * T + __ffs(mine) - 1;
                got = true;
            } else {
                next_try += (uint32_t)T;
            }
        }
        const double x = __shfl_sync(FULL, w.x, src), y = __shfl_sync(FULL, w.y, src), z = __shfl_sync(FULL, w.z, src);
        if (got) {
            v = mk(x, y, z);
            pending = false;
        }
        rej = __ballot_sync(FULL, pending);
    }
    return v;
}

}  // namespace rt
