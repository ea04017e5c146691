// Exact-skip ray marching.
//
// RayMarchingShape::ray_intersect (src/world/shapes/ray_marching.rs:20-74) walks the ray in fixed
// steps of 0.01 in t — up to 23 925 strictly sequential iterations in cornell_box.json — looking for
// the first sign change of the surface polynomial, then refines it `depth` times with step *= -0.01.
// The candidate t it returns depends on every rounded partial sum of `t += step; p += step*dir`, so
// the loop cannot simply be replaced by a root finder.  This file reproduces its result BIT FOR BIT
// while executing a small fraction of the iterations.  Two independent facts make that possible:
//
//  (1) Skipping m steps of an accumulator exactly (advance_exact).  While a += s stays inside one
//      binade [2^e, 2^(e+1)), every step adds the SAME multiple D of ulp(a): the exact sum is
//      (T + S) ulp + rem with a constant remainder, so round-to-nearest always resolves it the same
//      way.  D is computed from the bit patterns of s and the binade, an exact tie is resolved by its
//      parity rule, and the steps that cross a binade boundary — the only ones whose rounding is
//      irregular — are executed literally.  m steps therefore cost O(number of binades crossed)
//      integer operations, and the result is the very double the reference's loop would hold.
//
//  (2) Proving that the skipped samples are uneventful (safe_extent).  Restricted to the ray the
//      surface function is a univariate polynomial g(tau) = f(p0 + tau d) of degree <= 6; its Taylor
//      coefficients are obtained once per ray by evaluating the same polynomial text on truncated
//      Taylor series.  Starting from a point where |g| > M, hops of length
//          dtau = 2b / (|g'| + sqrt(g'^2 + 2 B2 b)),   b = |g| - M,   B2 >= max |g''| on the chord
//      stay inside {|g| >= M, same sign} by Taylor's theorem.  M bounds (64-fold) everything the
//      exact-arithmetic model ignores: the drift of the accumulated t and p against p0 + tau d (at
//      most half an ulp per step, N steps) seen through the largest slope of g on the chord, plus
//      the rounding of evaluating f.  On every skipped sample the reference's loop would therefore
//      have found no sign change, no |f| < 1e-15 and no range violation: it would only have executed
//      `r = next`.
//
// After a skip the surface function is evaluated at the landing sample like the reference does
// (`r = next`) and compared with the polynomial's prediction; a mismatch (the model does not apply to
// this ray) restores the saved state and finishes the ray with the plain loop.  Wherever skipping is
// not worthwhile or not provable the reference's literal step is taken.
// tests/test_gpu_intersect.py compares t bit-exactly with the oracle for all six surfaces.
#pragma once
#include "rt_math.cuh"

namespace rt {

// ---- (1) exact multi-step advance ---------------------------------------------------------------
// returns the value `a` holds after m iterations of `a = a + s` in IEEE double arithmetic
__device__ __forceinline__ double advance_exact(double a, double s, long long m) {
    const long long MANT = 0x000fffffffffffffLL;
    if (s == 0.0) return m > 0 ? a + s : a;
    const long long sb = __double_as_longlong(fabs(s));
    const int sexp = (int)(sb >> 52);
    const long long Ms = (sb & MANT) | (1LL << 52);
    const bool s_neg = s < 0.0;
    while (m > 0) {
        const long long bits = __double_as_longlong(a);
        const int exp = (int)((bits >> 52) & 0x7ff);
        const long long mant = bits & MANT;
        const bool a_neg = bits < 0;
        const int k = exp - sexp;           // ulp(a) = 2^k ulp(s):  s / ulp(a) = Ms / 2^k
        const bool up = (a_neg == s_neg);   // the magnitude grows
        // literal step wherever the regular-progression argument does not apply: zero / subnormal /
        // non-finite operands, a step comparable to the value itself, or the bottom of a binade
        // approached from above (the grid below it is finer)
        bool literal = exp == 0 || exp == 0x7ff || sexp == 0 || sexp == 0x7ff || k < 1 || (!up && mant == 0);
        long long D = 0;
        if (!literal && k <= 54) {
            const long long S = Ms >> k;
            const long long rem = Ms & ((1LL << k) - 1);
            const long long half = 1LL << (k - 1);
            if (rem > half) D = S + 1;
            else if (rem < half) D = S;
            else if (mant & 1) literal = true;   // exact tie from an odd mantissa: one literal step makes it even
            else D = S + (S & 1);                // exact tie from an even mantissa: round-half-even lands on even again
        }                                        // k > 54: |s| < ulp(a)/4, the sum rounds back to a (D = 0)
        if (literal) {
            a = a + s;
            m--;
            continue;
        }
        if (D == 0) return a;  // a + s == a for every remaining step
        // steps that provably stay inside this binade.  The quotient is taken in floating point
        // (operands < 2^53 are exact; the -2 absorbs its rounding): 64-bit integer division is far
        // more expensive on the GPU.  Going down, the landing mantissa must stay >= 1.
        double roomf = (up ? (double)(MANT - mant) : (double)(mant - 1)) / (double)D - 2.0;
        long long room = roomf > 0.0 ? (long long)roomf : 0;
        long long take = room < m ? room : m;
        if (take > 0) {
            a = __longlong_as_double(up ? bits + take * D : bits - take * D);
            m -= take;
        }
        if (m > 0) {  // next to the binade edge: literal steps carry it across
            a = a + s;
            m--;
        }
    }
    return a;
}

// ---- (2) the surface function along the ray as a univariate polynomial ---------------------------
#define RT_POLY_N 7  // degree <= 6 for every ShapeFunction of the reference
struct Jet {
    double c[RT_POLY_N];
};
__device__ __forceinline__ Jet jet_lin(double v, double d) {
    Jet r;
    r.c[0] = v;
    r.c[1] = d;
#pragma unroll
    for (int i = 2; i < RT_POLY_N; i++) r.c[i] = 0.0;
    return r;
}
__device__ __forceinline__ Jet operator+(Jet a, Jet b) {
#pragma unroll
    for (int i = 0; i < RT_POLY_N; i++) a.c[i] += b.c[i];
    return a;
}
__device__ __forceinline__ Jet operator-(Jet a, Jet b) {
#pragma unroll
    for (int i = 0; i < RT_POLY_N; i++) a.c[i] -= b.c[i];
    return a;
}
__device__ __forceinline__ Jet operator*(Jet a, Jet b) {
    Jet r;
#pragma unroll
    for (int k = 0; k < RT_POLY_N; k++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i <= k; i++) s = fma(a.c[i], b.c[k - i], s);
        r.c[k] = s;
    }
    return r;
}
__device__ __forceinline__ Jet operator+(Jet a, double b) { a.c[0] += b; return a; }
__device__ __forceinline__ Jet operator-(Jet a, double b) { a.c[0] -= b; return a; }
__device__ __forceinline__ Jet operator+(double a, Jet b) { b.c[0] += a; return b; }
__device__ __forceinline__ Jet operator-(double a, Jet b) {
#pragma unroll
    for (int i = 0; i < RT_POLY_N; i++) b.c[i] = -b.c[i];
    b.c[0] += a;
    return b;
}
__device__ __forceinline__ Jet operator*(Jet a, double b) {
#pragma unroll
    for (int i = 0; i < RT_POLY_N; i++) a.c[i] *= b;
    return a;
}
__device__ __forceinline__ Jet operator*(double a, Jet b) { return b * a; }

struct RayPoly {
    double c[RT_POLY_N];  // g(tau) = sum c[k] tau^k, tau = t - t0
    double t0;            // the value t had at the expansion point
    double M;             // {|g| >= M} is the provably uneventful region
    double B2;            // >= max |g''| for tau in [0, tau_hi]
    double tau_hi;        // the model covers tau in [0, tau_hi]
    __device__ __forceinline__ void eval(double tau, double& g, double& dg) const {
        double v = c[RT_POLY_N - 1], d = 0.0;
#pragma unroll
        for (int k = RT_POLY_N - 2; k >= 0; k--) {
            d = fma(d, tau, v);
            v = fma(v, tau, c[k]);
        }
        g = v;
        dg = d;
    }
};

// Taylor expansion of f along the ray around the current sample (t, p); n_steps = an upper bound of the
// number of steps the reference can still take on this ray (drift bound), dlen = |d|.
template <int KIND>
__device__ __forceinline__ void expand_ray(const double* q, D3 p, D3 d, double t, double t_end, double tau_hi,
                                           double n_steps, double G, RayPoly& P) {
    Jet g = surface_func_t<KIND, Jet>(q, jet_lin(p.x, d.x), jet_lin(p.y, d.y), jet_lin(p.z, d.z));
    double scale = 0.0, g1 = 0.0, b2 = 0.0;
    double pw = 1.0;  // tau_hi^k
#pragma unroll
    for (int k = 0; k < RT_POLY_N; k++) {
        P.c[k] = g.c[k];
        scale += fabs(g.c[k]) * pw;
        pw *= tau_hi;
    }
    pw = 1.0;
#pragma unroll
    for (int k = 1; k < RT_POLY_N; k++) {
        g1 += (double)k * fabs(g.c[k]) * pw;             // >= max |g'|
        if (k >= 2) b2 += (double)(k * (k - 1)) * fabs(g.c[k]) * (pw / tau_hi);  // >= max |g''|
        pw *= tau_hi;
    }
    // drift of the accumulators against the exact line p0 + tau d: at most half an ulp per step.  A
    // drift of t shifts the sample along the ray (seen through max |g'| <= g1); a drift e of p changes
    // f by at most G |e|, G >= sup |grad f| over the marching region (march_bounds.hpp).
    const double EPS = 1.1102230246251565e-16;  // 2^-53
    double tmax = fmax(fabs(t), fabs(t_end)) + tau_hi;
    double drift_t = n_steps * EPS * tmax;
    double pmax = fmax(fmax(fabs(p.x), fabs(p.y)), fabs(p.z)) + tau_hi * fmax(fmax(fabs(d.x), fabs(d.y)), fabs(d.z));
    double drift_p = 1.7320508075688772 * n_steps * EPS * pmax;
    P.t0 = t;
    P.M = 64.0 * (g1 * drift_t + G * drift_p) + 1e-9 * scale + 1e-300;
    P.B2 = b2 * (1.0 + 1e-9);
    P.tau_hi = tau_hi;
}

// furthest tau from `tau` in direction dir (+1 / -1), not beyond tau_lim, such that the whole stretch
// is provably inside {|g| >= M, sign(g) constant}
__device__ __forceinline__ double safe_extent(const RayPoly& P, double tau, double dir, double tau_lim, double min_hop) {
    double g, dg;
    P.eval(tau, g, dg);
    for (int hop = 0; hop < 64; hop++) {
        double b = fabs(g) - P.M;
        if (!(b > 0.0)) break;
        double a = fabs(dg);
        double dt = 2.0 * b / (a + sqrt(a * a + 2.0 * P.B2 * b));
        dt *= 0.999;  // rounding of the bound arithmetic itself
        double nt = tau + dir * dt;
        if ((dir > 0.0) ? (nt >= tau_lim) : (nt <= tau_lim)) return tau_lim;
        if (!(dt > min_hop)) break;
        tau = nt;
        P.eval(tau, g, dg);
    }
    return tau;
}

#define RT_MARCH_MIN_JUMP 8

// RayMarchingShape::ray_intersect's loops (ray_marching.rs:27-57) with exact skipping.
// G = gradient bound of the surface over its marching region; a non-finite G (or a chord of few
// steps) gives exactly the plain loop.
template <int KIND>
__device__ __forceinline__ bool march_loop_skip(const double* q, D3 o, D3 d, double start, double end, double min_t,
                                                double max_t, double G, double& t_out, unsigned long long& evals) {
    double step = q[1];
    const int depth = (int)q[2];
    double t = start;
    D3 p = o + t * d;
    double r = surface_func<KIND>(q, p);
    unsigned long long n = 0;
    const double step0 = step;
    // skipping pays when the chord holds many steps (NaN-proof comparisons)
    bool skip_ok = (G == G) && G < 1e300 && step > 0.0 && (end - start) > 64.0 * step && (end - start) < 1e300;
    bool have_poly = false;
    RayPoly P;
    for (int it = 0; it < depth; it++) {
        bool finished = false;
        D3 sd = step * d;  // `step * dir`, loop-invariant until the step changes
        const double abs_step = fabs(step);
        const double dir = step > 0.0 ? 1.0 : -1.0;
        int cooldown = 0, backoff = 4;
        for (;;) {
            if (t > end || t < start || n > RT_MARCH_BUDGET) {
                evals += n;
                return false;
            }
            if (skip_ok && cooldown == 0) {
                if (!have_poly) {
                    // expand around the current sample (the first one); covers every later sample of the ray
                    double tau_hi = (end - t) + 4.0 * step0;
                    expand_ray<KIND>(q, p, d, t, end, tau_hi, tau_hi / step0 + 1024.0, G, P);
                    n += 8;  // cost of the expansion in evaluation-equivalents (statistics only)
                    have_poly = true;
                }
                const double tau = t - P.t0;
                // the range checks must not fire on skipped samples: stay 2 steps inside [start, end]
                double tau_lim = (dir > 0.0 ? end : start) - P.t0 - dir * 2.0 * abs_step;
                tau_lim = fmin(fmax(tau_lim, 0.0), P.tau_hi);
                double ts = safe_extent(P, tau, dir, tau_lim, abs_step);
                double mf = (ts - tau) * dir / abs_step * (1.0 - 1e-9) - 2.0;
                if (mf >= (double)RT_MARCH_MIN_JUMP) {
                    const long long m = (long long)fmin(mf, 1.0e15);
                    const double st = t;
                    const D3 sp = p;
                    t = advance_exact(t, step, m);
                    p.x = advance_exact(p.x, sd.x, m);
                    p.y = advance_exact(p.y, sd.y, m);
                    p.z = advance_exact(p.z, sd.z, m);
                    double land = surface_func<KIND>(q, p);  // the reference's `r = next` at the landing sample
                    n++;
                    double gp, dgp;
                    P.eval(t - P.t0, gp, dgp);
                    // self-check: the landing value must be what the polynomial predicts and keep the sign
                    bool same_sign = ((land > 0.0) == (r > 0.0)) && land != 0.0;
                    if (same_sign && fabs(land - gp) <= 0.25 * P.M && fabs(land) >= 0.5 * P.M) {
                        r = land;
                        backoff = 4;
                        continue;
                    }
                    t = st;  // the model does not describe this ray: undo and finish it with the plain loop
                    p = sp;
                    skip_ok = false;
                } else {
                    // inside the |g| < M zone or next to a range limit: plain steps, retry later
                    cooldown = backoff;
                    backoff = min(backoff * 2, 64);
                }
            }
            if (cooldown > 0) cooldown--;
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

__device__ inline bool march_candidate_skip(const double* q, D3 o, D3 d, double start, double end, double min_t,
                                            double max_t, double G, double& t, unsigned long long& evals) {
    switch ((int)q[0]) {
        case RT_SURF_HEART: return march_loop_skip<RT_SURF_HEART>(q, o, d, start, end, min_t, max_t, G, t, evals);
        case RT_SURF_SINE: return march_loop_skip<RT_SURF_SINE>(q, o, d, start, end, min_t, max_t, G, t, evals);
        case RT_SURF_STAR: return march_loop_skip<RT_SURF_STAR>(q, o, d, start, end, min_t, max_t, G, t, evals);
        case RT_SURF_DUPIN: return march_loop_skip<RT_SURF_DUPIN>(q, o, d, start, end, min_t, max_t, G, t, evals);
        case RT_SURF_HUNTS: return march_loop_skip<RT_SURF_HUNTS>(q, o, d, start, end, min_t, max_t, G, t, evals);
        default: return march_loop_skip<RT_SURF_CUSHION>(q, o, d, start, end, min_t, max_t, G, t, evals);
    }
}

}  // namespace rt
