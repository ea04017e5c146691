// Conservative FP32 bounding-ball cull in front of the exact FP64 shape tests.
//
// ShapeCollection::ray_intersect (src/world/shapes/mod.rs:573-597) tests every shape for every ray in
// FP64 (33 flop ray -> object space + 19 flop unit-sphere discriminant, Sphere::ray_intersect :330-356).
// A shape may be skipped for a ray iff the reference's test would return None for EVERY max_t; then
// skipping it leaves the sequential loop's state untouched and the result is bit-identical.
//
// Host side (cull_entry): for a Sphere / Cube with inverse rows [M | tv] the object-space ball |p'| <= ext
// (ext = 1 / sqrt 3) is, in world space, inside the ball  centre C = -M^-1 tv,  radius R = ext * ||M^-1||_2
// (Gershgorin bound of the spectral norm; exact for translate*rotate*scale).  It is derived from the
// INVERSE rows because those are what the exact test multiplies the ray with.
//
// Device side (cull_pass), FP32 with explicit FMAs, direction renormalised in FP32:
//     oc = C - o,  b = oc.d,  p = oc - b d,
//     reject  iff  |p|^2 > rhs,      rhs = 1.1 R^2 + 3B(|C|^2 + |o|^2)
// i.e. the test is on the LINE, not the half-line: a sphere behind the origin must not be culled, because
// the reference's D == 0 branch (src/world/shapes/mod.rs:343-344) accepts a tangent line without any
// range check, wherever the tangent point lies (tests: test_cull_is_conservative_on_fixture_scenes).
// Why this is conservative (u = 2^-24):
//   * every FP32 quantity above carries an absolute error <= 16u(|C| + |o| + |oc|) (input rounding, the
//     5u error of the renormalised direction, 3 FMA roundings), so a line that truly touches the ball has
//     |p_computed| <= R + delta with delta^2 <= 3*256 u^2 (|C|^2+|o|^2+|oc|^2) = 2.7e-12 (...), and
//     (R + delta)^2 <= 1.1 R^2 + 11 delta^2 <= 1.1 R^2 + 3e-11 (...);  |oc|^2 <= 2|C|^2 + 2|o|^2;
//   * a line whose world distance from C exceeds sqrt(1.1) R misses the object-space ball by >= 10 % of
//     ext^2 in the reference's discriminant, whose FP64 rounding error is <= ~1e3 * 2^-53 * kappa^2 *
//     (|C|^2+|o|^2+|oc|^2)/R^2 relative to it (kappa = cond(M)); shapes with kappa > 30 are never culled,
//     which keeps that term below 1e-10 (...) as well;
//   * B = 2e-10 >= 3e-11 + 1e-10;
//   * NaN / Inf operands make the comparisons false -> the exact test runs.
// Rectangles are never culled: a ray lying in their plane gives t = NaN, which the reference accepts
// (SURVEY A.6), whatever its distance from the rectangle.  Ray-marched shapes are handled by march_needed.
// Evidence: RT_ISECT_VERIFY runs the literal loop beside the culled one on every ray and tests every
// culled (ray, shape) pair exactly; tests/test_gpu_intersect.py requires zero disagreements.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rt_b200.h"

namespace rt {

#define RT_CULL_B 2e-10
#define RT_CULL_MAX_KAPPA 30.0

// ---- host: one table entry per shape --------------------------------------------------------------
// (cx, cy, cz, A): A = 1.1 R^2 + 3B|C|^2 rounded up; A = +inf: never culled; A = -inf: never tested
// (padding and ray-marched shapes, masked out by the chunk's valid bits anyway).
inline float4 cull_entry(const double* m, int kind) {
    const float INF = INFINITY;
    float4 never = make_float4(0.f, 0.f, 0.f, INF);
    if (kind == RT_SHAPE_MARCH) return make_float4(0.f, 0.f, 0.f, -INF);
    if (kind != RT_SHAPE_SPHERE && kind != RT_SHAPE_CUBE) return never;
    const double a[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}};
    const double tv[3] = {m[3], m[7], m[11]};
    double c[3][3];  // cofactors
    c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    c[0][1] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    c[0][2] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    c[1][0] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    c[1][2] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    c[2][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    c[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    const double det = a[0][0] * c[0][0] + a[0][1] * c[0][1] + a[0][2] * c[0][2];
    if (!(fabs(det) > 0.0) || !isfinite(det)) return never;
    double inv[3][3];  // M^-1 = adj(M) / det, adj = cofactor^T
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) inv[r][k] = c[k][r] / det;
    auto gersh = [](const double x[3][3]) {  // >= lambda_max(x^T x) = ||x||_2^2
        double g[3][3];
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < 3; k++) g[r][k] = x[0][r] * x[0][k] + x[1][r] * x[1][k] + x[2][r] * x[2][k];
        double best = 0.0;
        for (int r = 0; r < 3; r++) best = fmax(best, fabs(g[r][0]) + fabs(g[r][1]) + fabs(g[r][2]));
        return best;
    };
    const double n_m = gersh(a), n_inv = gersh(inv);
    if (!isfinite(n_m) || !isfinite(n_inv)) return never;
    const double kappa = sqrt(n_m * n_inv);
    if (!(kappa <= RT_CULL_MAX_KAPPA)) return never;
    double C[3];
    for (int r = 0; r < 3; r++) C[r] = -(inv[r][0] * tv[0] + inv[r][1] * tv[1] + inv[r][2] * tv[2]);
    const double ext2 = kind == RT_SHAPE_CUBE ? 3.0 : 1.0;
    const double R2 = ext2 * n_inv * (1.0 + 1e-9);
    const double cc = C[0] * C[0] + C[1] * C[1] + C[2] * C[2];
    const double A = 1.1 * R2 + 3.0 * RT_CULL_B * cc;
    if (!isfinite(A) || !isfinite(C[0]) || !isfinite(C[1]) || !isfinite(C[2])) return never;
    float Af = (float)A;
    if ((double)Af < A) Af = nextafterf(Af, INF);
    return make_float4((float)C[0], (float)C[1], (float)C[2], Af);
}

// ---- device -----------------------------------------------------------------------------------------
struct CullRay {
    float ox, oy, oz, dx, dy, dz, rhs0;  // rhs0 = 3B|o|^2
};

__device__ __forceinline__ CullRay make_cull_ray(double ox, double oy, double oz, double dx, double dy, double dz) {
    CullRay r;
    r.ox = __double2float_rn(ox);
    r.oy = __double2float_rn(oy);
    r.oz = __double2float_rn(oz);
    float x = __double2float_rn(dx), y = __double2float_rn(dy), z = __double2float_rn(dz);
    float s = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(s));
    if (!(s > 1e-30f && s < 1e30f)) inv = NAN;  // zero / denormal / huge / NaN direction: nothing is culled
    r.dx = __fmul_rn(x, inv);
    r.dy = __fmul_rn(y, inv);
    r.dz = __fmul_rn(z, inv);
    float oo = __fmaf_rn(r.oz, r.oz, __fmaf_rn(r.oy, r.oy, __fmul_rn(r.ox, r.ox)));
    r.rhs0 = __fmul_rn((float)(3.0 * RT_CULL_B), oo);
    return r;
}

// true: the exact test must run.  s = (C, A) from cull_entry.
__device__ __forceinline__ bool cull_pass(const CullRay& r, float4 s) {
    float ocx = __fsub_rn(s.x, r.ox), ocy = __fsub_rn(s.y, r.oy), ocz = __fsub_rn(s.z, r.oz);
    float b = __fmaf_rn(ocz, r.dz, __fmaf_rn(ocy, r.dy, __fmul_rn(ocx, r.dx)));
    float px = __fmaf_rn(-b, r.dx, ocx), py = __fmaf_rn(-b, r.dy, ocy), pz = __fmaf_rn(-b, r.dz, ocz);
    float p2 = __fmaf_rn(pz, pz, __fmaf_rn(py, py, __fmul_rn(px, px)));
    float rhs = __fadd_rn(s.w, r.rhs0);
    return !(p2 > rhs);
}

}  // namespace rt
