// Rigorous bounds on the first and second derivatives of an implicit surface polynomial over its
// marching region, computed once per scene on the host.
//
// The exact-skip marcher (rt_march.cuh) may skip m fixed steps without evaluating the surface
// function only if it can PROVE that no sample among them changes sign or comes within 1e-15 of
// zero.  It models f along the ray as an exact univariate polynomial; what the model ignores is the
// rounding drift of the accumulated sample positions, at most a few 1e-11 of the object's size.  To
// turn that displacement into a bound on |f| it needs G >= sup |grad f| over the region: the
// polynomial is evaluated here with second-order interval jets (value, gradient, Hessian as
// intervals, outward rounded) on a grid of cells covering the (inflated) bounding ellipsoid; G and H
// are the largest Frobenius bounds of gradient and Hessian (H is reported for diagnostics).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "rt_math.cuh"

namespace rt {
namespace bounds {

struct Iv {
    double lo, hi;
};
inline double dn(double x) { return std::nextafter(x, -std::numeric_limits<double>::infinity()); }
inline double up(double x) { return std::nextafter(x, std::numeric_limits<double>::infinity()); }
inline Iv iv(double a) { return Iv{a, a}; }
inline Iv operator+(Iv a, Iv b) { return Iv{dn(a.lo + b.lo), up(a.hi + b.hi)}; }
inline Iv operator-(Iv a, Iv b) { return Iv{dn(a.lo - b.hi), up(a.hi - b.lo)}; }
inline Iv operator-(Iv a) { return Iv{-a.hi, -a.lo}; }
inline Iv operator*(Iv a, Iv b) {
    double p1 = a.lo * b.lo, p2 = a.lo * b.hi, p3 = a.hi * b.lo, p4 = a.hi * b.hi;
    return Iv{dn(std::min(std::min(p1, p2), std::min(p3, p4))), up(std::max(std::max(p1, p2), std::max(p3, p4)))};
}
inline double mag(Iv a) { return std::max(std::fabs(a.lo), std::fabs(a.hi)); }

// second-order jet: value, gradient (x,y,z), Hessian (xx, xy, xz, yy, yz, zz)
struct Jet2 {
    Iv v, g[3], h[6];
};
inline int hidx(int i, int j) {
    if (i > j) std::swap(i, j);
    static const int t[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    return t[i][j];
}
inline Jet2 constant(double c) {
    Jet2 r;
    r.v = iv(c);
    for (auto& x : r.g) x = iv(0.0);
    for (auto& x : r.h) x = iv(0.0);
    return r;
}
inline Jet2 variable(Iv range, int axis) {
    Jet2 r = constant(0.0);
    r.v = range;
    r.g[axis] = iv(1.0);
    return r;
}
inline Jet2 operator+(const Jet2& a, const Jet2& b) {
    Jet2 r;
    r.v = a.v + b.v;
    for (int i = 0; i < 3; i++) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < 6; i++) r.h[i] = a.h[i] + b.h[i];
    return r;
}
inline Jet2 operator-(const Jet2& a, const Jet2& b) {
    Jet2 r;
    r.v = a.v - b.v;
    for (int i = 0; i < 3; i++) r.g[i] = a.g[i] - b.g[i];
    for (int i = 0; i < 6; i++) r.h[i] = a.h[i] - b.h[i];
    return r;
}
inline Jet2 operator*(const Jet2& a, const Jet2& b) {
    Jet2 r;
    r.v = a.v * b.v;
    for (int i = 0; i < 3; i++) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            r.h[hidx(i, j)] = a.h[hidx(i, j)] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[hidx(i, j)];
    return r;
}
inline Jet2 operator+(const Jet2& a, double b) { return a + constant(b); }
inline Jet2 operator-(const Jet2& a, double b) { return a - constant(b); }
inline Jet2 operator*(const Jet2& a, double b) { return a * constant(b); }
inline Jet2 operator+(double a, const Jet2& b) { return constant(a) + b; }
inline Jet2 operator-(double a, const Jet2& b) { return constant(a) - b; }
inline Jet2 operator*(double a, const Jet2& b) { return constant(a) * b; }

// Frobenius bounds of the gradient (G) and of the Hessian (H) over one cell
template <int KIND>
inline void bounds_cell(const double* q, Iv x, Iv y, Iv z, double* G, double* H) {
    Jet2 f = surface_func_t<KIND, Jet2>(q, variable(x, 0), variable(y, 1), variable(z, 2));
    double s = 0.0, g = 0.0;
    for (int i = 0; i < 3; i++) {
        double m = mag(f.g[i]);
        g += m * m;
        for (int j = 0; j < 3; j++) {
            double mm = mag(f.h[hidx(i, j)]);
            s += mm * mm;
        }
    }
    *G = std::sqrt(g) * (1.0 + 1e-12);
    *H = std::sqrt(s) * (1.0 + 1e-12);
}

constexpr double kRegionInflate = 1.06;  // samples may overshoot the bound by one step; see march guard

// magnitude arithmetic: every operation returns an upper bound of the sum of the absolute values of the
// terms it combines, so evaluating the polynomial text on it bounds the "absolute-value polynomial"
// F >= sum |monomial| over the region.  The rounding error of evaluating f in floating point with n
// operations is at most gamma_n * F (running error analysis).
struct Mag {
    double v;
};
inline Mag operator+(Mag a, Mag b) { return Mag{a.v + b.v}; }
inline Mag operator-(Mag a, Mag b) { return Mag{a.v + b.v}; }
inline Mag operator*(Mag a, Mag b) { return Mag{a.v * b.v}; }
inline Mag operator+(Mag a, double b) { return Mag{a.v + std::fabs(b)}; }
inline Mag operator-(Mag a, double b) { return Mag{a.v + std::fabs(b)}; }
inline Mag operator*(Mag a, double b) { return Mag{a.v * std::fabs(b)}; }
inline Mag operator+(double a, Mag b) { return Mag{std::fabs(a) + b.v}; }
inline Mag operator-(double a, Mag b) { return Mag{std::fabs(a) + b.v}; }
inline Mag operator*(double a, Mag b) { return Mag{std::fabs(a) * b.v}; }

// radii of the marching bound (ShapeFunction::intersect_bound / get_bounds)
inline void bound_radii(const double* q, double r[3]) {
    if ((int)q[0] == RT_SURF_HEART) {
        const double sr = 1.45;
        r[0] = sr; r[1] = sr / 2.05; r[2] = sr;
    } else {
        r[0] = r[1] = r[2] = q[7];
    }
}

// F >= sum of |monomials| of f anywhere in the inflated bounding box of the marching region
inline double region_magnitude(const double* q) {
    const double INF = std::numeric_limits<double>::infinity();
    double r[3];
    bound_radii(q, r);
    for (int a = 0; a < 3; a++)
        if (!(r[a] > 0.0) || !std::isfinite(r[a])) return INF;
    Mag x{r[0] * kRegionInflate}, y{r[1] * kRegionInflate}, z{r[2] * kRegionInflate};
    Mag f;
    switch ((int)q[0]) {
        case RT_SURF_HEART: f = surface_func_t<RT_SURF_HEART, Mag>(q, x, y, z); break;
        case RT_SURF_SINE: f = surface_func_t<RT_SURF_SINE, Mag>(q, x, y, z); break;
        case RT_SURF_STAR: f = surface_func_t<RT_SURF_STAR, Mag>(q, x, y, z); break;
        case RT_SURF_DUPIN: f = surface_func_t<RT_SURF_DUPIN, Mag>(q, x, y, z); break;
        case RT_SURF_HUNTS: f = surface_func_t<RT_SURF_HUNTS, Mag>(q, x, y, z); break;
        default: f = surface_func_t<RT_SURF_CUSHION, Mag>(q, x, y, z); break;
    }
    if (!(f.v == f.v)) return INF;
    return f.v * (1.0 + 1e-9);
}

// G >= sup |grad f| and H >= sup |u^T Hess f u| over the inflated bounding ellipsoid; +inf when the
// parameters are unusable (the marcher then never skips and behaves exactly like the plain loop)
inline void region_bounds(const double* q, double* G_out, double* H_out) {
    const double INF = std::numeric_limits<double>::infinity();
    *G_out = *H_out = INF;
    double r[3];
    bound_radii(q, r);
    for (int a = 0; a < 3; a++)
        if (!(r[a] > 0.0) || !std::isfinite(r[a])) return;
    const int N = 16;
    double best = 0.0, best_g = 0.0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            for (int k = 0; k < N; k++) {
                int c[3] = {i, j, k};
                Iv box[3];
                double near2 = 0.0;  // squared normalised distance of the cell's nearest point to the centre
                for (int a = 0; a < 3; a++) {
                    double R = r[a] * kRegionInflate;
                    box[a].lo = -R + 2.0 * R * c[a] / N;
                    box[a].hi = -R + 2.0 * R * (c[a] + 1) / N;
                    double nearest = (box[a].lo > 0.0) ? box[a].lo : (box[a].hi < 0.0 ? box[a].hi : 0.0);
                    near2 += (nearest / R) * (nearest / R);
                }
                if (near2 > 1.0 + 1e-9) continue;  // cell entirely outside the inflated ellipsoid
                double h, g;
                switch ((int)q[0]) {
                    case RT_SURF_HEART: bounds_cell<RT_SURF_HEART>(q, box[0], box[1], box[2], &g, &h); break;
                    case RT_SURF_SINE: bounds_cell<RT_SURF_SINE>(q, box[0], box[1], box[2], &g, &h); break;
                    case RT_SURF_STAR: bounds_cell<RT_SURF_STAR>(q, box[0], box[1], box[2], &g, &h); break;
                    case RT_SURF_DUPIN: bounds_cell<RT_SURF_DUPIN>(q, box[0], box[1], box[2], &g, &h); break;
                    case RT_SURF_HUNTS: bounds_cell<RT_SURF_HUNTS>(q, box[0], box[1], box[2], &g, &h); break;
                    default: bounds_cell<RT_SURF_CUSHION>(q, box[0], box[1], box[2], &g, &h); break;
                }
                if (!(h == h) || !(g == g)) return;
                best = std::max(best, h);
                best_g = std::max(best_g, g);
            }
    *G_out = best_g;
    *H_out = best;
}

}  // namespace bounds
}  // namespace rt
