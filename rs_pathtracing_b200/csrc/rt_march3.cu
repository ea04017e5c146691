// rt_march3.cu -- K3 of the wavefront renderer, third generation: exact-skip marching with a POOL OF RAYS PER WARP.
//
// What the profiles of k_march (one ray per lane, rt_march_kernels.cu) say (profiles/r2b_*): 7-13 of 32 lanes
// active.  Half of its warp instructions are the exact advance of a jump (advance_exact: an accumulator that
// crosses zero walks ~15 binades, the others one) at 2.5-6 lanes, a quarter the hop planning (1-100 hops per ray)
// at ~6 lanes: the lanes of a warp sit in different phases of their rays, and inside a phase the trip counts of
// the loops differ by an order of magnitude -- a warp executes the union of all of it.
//
// Here a ray is not tied to a lane.  Each warp owns RT_M3_R ray RECORDS in shared memory (35 doubles each: the
// reference loop's state, the ray's polynomial model, the planned jump) and its 32 lanes are workers: the warp
// repeatedly picks the phase most of its rays are waiting for and runs that phase's SERVICE over the list of those
// rays,
//     TRANS    finish a marched shape (candidate t against the path's best), start the next one (bound, first
//              sample, polynomial expansion along the ray) or retire the record and write the path's result
//     PLAN     how many iterations can provably be skipped (Marcher::plan_*: the hop loop)
//     ADV      the exact advance of t, p.x, p.y, p.z by that many steps (advance_iter), 4 tasks per ray
//     LAND     the reference's `r = next` at the landing sample + the model self-check
//     LIT      the reference's literal steps
// and the services whose work per ray varies (PLAN: hops, ADV: binades, LIT: steps) hand a lane the NEXT ray of
// the list the moment it finishes one (dynamic pick-up, one ballot per trip), so that a trip of the loop always
// runs with as many lanes as there are rays left.  Everything is warp-private: no block barrier, no atomics
// except the one on the global queue head, no inter-warp waiting.
//
// The arithmetic per ray is Marcher's (rt_march.cuh), unchanged: the t this kernel returns is the reference
// loop's, bit for bit (tests/test_gpu_intersect.py, test_gpu_render.py::test_alternative_schedules_give_the_same_frame
// compare it with k_march and the oracle).  One deliberate difference in the model's bookkeeping: the landing
// self-check evaluates the ray polynomial P at tau = t_landing - t0 instead of the Taylor-shifted copy the planner
// used (7 doubles less per record); both are the same polynomial, the evaluation error of either is covered by
// err0 (rt_march.cuh, expand_ray).
// Compile with -fmad=false (see rt_math.cuh).
#include <cuda_runtime.h>

#include <algorithm>

#include "rt_queues.cuh"

#ifndef RT_M3_R
#define RT_M3_R 48          // ray records per warp
#endif
#define RT_M3_RS 35         // doubles per record (odd: consecutive records start in different banks)
#ifndef RT_M3_WARPS
#define RT_M3_WARPS 4
#endif
#ifndef RT_M3_MIN_BLOCKS
#define RT_M3_MIN_BLOCKS 4
#endif
#ifndef RT_M3_LIT_CAP
#define RT_M3_LIT_CAP 24    // literal steps per LIT service call (a ray that needs thousands must not hold the warp)
#endif
#ifndef RT_M3_REFILL_MIN
#define RT_M3_REFILL_MIN 12 // refill when this many records are free (or nothing is live)
#endif

enum { M3_FREE = 0, M3_TRANS = 1, M3_PLAN = 2, M3_ADV = 3, M3_LAND = 4, M3_LIT = 5, M3_NPH = 6 };

// record layout (doubles)
enum {
    F_T = 0, F_R = 1, F_STEP = 2, F_P = 3, F_D = 6, F_START = 9, F_END = 10, F_BEST = 11,
    F_STATE = 12,   // (it | cooldown << 8 | backoff << 16 | flags << 24, evaluations)
    F_PATH = 13,    // (path slot, march-queue entry)
    F_MASK = 14,    // (remaining shape mask, winner)
    F_SHAPE = 15,   // (march-list index k or 0xffffffff, shape index)
    F_C = 16,       // P.c[0..6]
    F_P0 = 23,      // P.p0
    F_ERR0 = 26, F_DRIFT1 = 27,
    F_M = 28, F_MJ = 29,      // the planned jump: uncertainty band, number of iterations
    F_NT = 30, F_NP = 31,     // its landing sample (34 = one spare)
};
#define M3_FLAG_SKIP_OK 1u
#define M3_FLAG_MORE 2u

__device__ __forceinline__ uint2 m3_u2(const double* rec, int f) { return *reinterpret_cast<const uint2*>(rec + f); }
__device__ __forceinline__ void m3_set_u2(double* rec, int f, uint32_t x, uint32_t y) {
    *reinterpret_cast<uint2*>(rec + f) = make_uint2(x, y);
}
__device__ __forceinline__ uint8_t m3_phase_of(int ph) {
    return ph == RT_PHASE_ATTEMPT ? M3_PLAN : ph == RT_PHASE_LITERAL ? M3_LIT : M3_TRANS;
}

// the loop state literal() / phase() / land touch
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_load_core(const DevScene& S, const double* rec, Marcher<KIND, COUNT>& m, bool& more) {
    m.t = rec[F_T]; m.r = rec[F_R]; m.step = rec[F_STEP];
    m.p = mk(rec[F_P], rec[F_P + 1], rec[F_P + 2]);
    m.d = mk(rec[F_D], rec[F_D + 1], rec[F_D + 2]);
    m.start = rec[F_START]; m.end = rec[F_END];
    const uint2 w = m3_u2(rec, F_STATE), ks = m3_u2(rec, F_SHAPE);
    m.it = (int)(w.x & 0xffu);
    m.cooldown = (int)((w.x >> 8) & 0xffu);
    m.backoff = (int)((w.x >> 16) & 0xffu);
    m.skip_ok = ((w.x >> 24) & M3_FLAG_SKIP_OK) != 0;
    more = ((w.x >> 24) & M3_FLAG_MORE) != 0;
    m.have_poly = true;   // (expanded when the shape is started; only read when skip_ok)
    m.n = w.y;
    m.q = S.params + RT_SHAPE_PARAMS * (int)ks.y;
    m.step0 = m.q[1];
    m.depth = (int)m.q[2];
    m.sd = m.step * m.d;
    if (COUNT) m.prof[0] = m.prof[1] = m.prof[2] = m.prof[3] = 0;
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_store_state(double* rec, const Marcher<KIND, COUNT>& m, bool more) {
    const uint32_t flags = (m.skip_ok ? M3_FLAG_SKIP_OK : 0u) | (more ? M3_FLAG_MORE : 0u);
    const uint32_t nn = m.n > 0xffffffffull ? 0xffffffffu : (uint32_t)m.n;
    m3_set_u2(rec, F_STATE, (uint32_t)m.it | ((uint32_t)m.cooldown << 8) | ((uint32_t)m.backoff << 16) | (flags << 24), nn);
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_store_sample(double* rec, const Marcher<KIND, COUNT>& m) {
    rec[F_T] = m.t; rec[F_R] = m.r; rec[F_STEP] = m.step;
    rec[F_P] = m.p.x; rec[F_P + 1] = m.p.y; rec[F_P + 2] = m.p.z;
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_load_poly(const double* rec, Marcher<KIND, COUNT>& m) {
    constexpr int DEG = Marcher<KIND, COUNT>::DEG;
#pragma unroll
    for (int i = 0; i <= DEG; i++) m.P.c[i] = rec[F_C + i];
    m.P.t0 = rec[F_START];                              // the model is expanded at the first sample, t = start
    m.P.p0 = mk(rec[F_P0], rec[F_P0 + 1], rec[F_P0 + 2]);
    m.P.tau_hi = (m.end - m.start) + 4.0 * m.step0;    // as in plan_begin
    m.P.err0 = rec[F_ERR0]; m.P.drift1 = rec[F_DRIFT1];
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_add_prof(DevCounters& c, const Marcher<KIND, COUNT>& m) {
    if (COUNT) {
#pragma unroll
        for (int k = 0; k < 4; k++) c.march_prof[k] += m.prof[k];
    }
}

// the list of the records in phase `want` (this warp's), in slot order; returns its length
__device__ __forceinline__ int m3_build_list(const uint8_t* s_ph, uint8_t* s_list, int want, int lane) {
    int len = 0;
#pragma unroll
    for (int row = 0; row < (RT_M3_R + 31) / 32; row++) {
        const int slot = row * 32 + lane;
        const bool hit = slot < RT_M3_R && s_ph[slot] == want;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (hit) s_list[len + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)slot;
        len += __popc(mask);
    }
    __syncwarp();
    return len;
}

template <int KIND, bool COUNT>
__global__ void __launch_bounds__(32 * RT_M3_WARPS, RT_M3_MIN_BLOCKS)
k_march3(DevScene S, uint32_t kind_mask, PathQueue in, HitQueue hq, const uint32_t* __restrict__ march_count,
         uint32_t* head, DevCounters* g_counters) {
    typedef Marcher<KIND, COUNT> M;
    constexpr int DEG = M::DEG;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    double* const recs = reinterpret_cast<double*>(rt_smem_raw) + (size_t)warp * RT_M3_R * RT_M3_RS;
    uint8_t* const bytes = rt_smem_raw + (size_t)RT_M3_WARPS * RT_M3_R * RT_M3_RS * sizeof(double);
    uint8_t* const s_ph = bytes + warp * (2 * RT_M3_R);
    uint8_t* const s_list = s_ph + RT_M3_R;
    DevCounters c = {};
    const uint32_t n = *march_count;
    bool exhausted = n == 0;
    for (int i = lane; i < RT_M3_R; i += 32) s_ph[i] = M3_FREE;
    __syncwarp();

    for (;;) {
        // ---- how many records wait for what (scalar counters: no dynamically indexed local array) ------------
        int n_free = 0, want = M3_TRANS, want_cnt = 0;
        unsigned free_mask[(RT_M3_R + 31) / 32];
        {
            int ph_row[(RT_M3_R + 31) / 32];
#pragma unroll
            for (int row = 0; row < (RT_M3_R + 31) / 32; row++) {
                const int slot = row * 32 + lane;
                ph_row[row] = slot < RT_M3_R ? (int)s_ph[slot] : -1;
                free_mask[row] = __ballot_sync(FULL, ph_row[row] == M3_FREE);
                n_free += __popc(free_mask[row]);
            }
#pragma unroll
            for (int k = M3_TRANS; k < M3_NPH; k++) {
                int ck = 0;
#pragma unroll
                for (int row = 0; row < (RT_M3_R + 31) / 32; row++) ck += __popc(__ballot_sync(FULL, ph_row[row] == k));
                if (ck > want_cnt) {   // the phase most records wait for (ties: the earlier phase)
                    want_cnt = ck;
                    want = k;
                }
            }
        }
        const int live = RT_M3_R - n_free;
        // ---- refill the free records from the march queue --------------------------------------------------
        if (!exhausted && n_free > 0 && (live == 0 || n_free >= RT_M3_REFILL_MIN)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(head, (uint32_t)n_free);
            base = __shfl_sync(FULL, base, 0);
            int before = 0;
#pragma unroll
            for (int row = 0; row < (RT_M3_R + 31) / 32; row++) {
                const int slot = row * 32 + lane;
                if ((free_mask[row] >> lane) & 1u) {
                    const uint32_t j = base + (uint32_t)(before + __popc(free_mask[row] & lt));
                    if (j < n) {
                        const uint32_t mask = hq.mq_mask[j] & kind_mask;
                        if (mask != 0) {
                            const uint32_t pslot = hq.mq_slot[j];
                            double* rec = recs + slot * RT_M3_RS;
                            rec[F_BEST] = hq.t[pslot];
                            m3_set_u2(rec, F_PATH, pslot, j);
                            m3_set_u2(rec, F_MASK, mask, (uint32_t)hq.index[pslot]);
                            m3_set_u2(rec, F_SHAPE, 0xffffffffu, 0u);
                            s_ph[slot] = M3_TRANS;
                        }
                    }
                }
                before += __popc(free_mask[row]);
            }
            if (base + (uint32_t)n_free >= n) exhausted = true;
            __syncwarp();
            continue;
        }
        if (live == 0) break;   // (exhausted)
        const int len = m3_build_list(s_ph, s_list, want, lane);

        if (want == M3_TRANS) {
            // ---- finish a marched shape / start the next one / retire the record ---------------------------
            for (int base = 0; base < len; base += 32) {
                const int idx = base + lane;
                if (idx < len) {
                    const int slot = s_list[idx];
                    double* rec = recs + slot * RT_M3_RS;
                    const uint2 pe = m3_u2(rec, F_PATH), mw = m3_u2(rec, F_MASK), ks = m3_u2(rec, F_SHAPE);
                    const uint32_t pslot = pe.x, entry = pe.y;
                    uint32_t mask = mw.x;
                    int winner = (int)mw.y;
                    double best = rec[F_BEST];
                    bool retired = false;
                    if (ks.x != 0xffffffffu) {   // a shape has just been marched to its end
                        const uint2 w = m3_u2(rec, F_STATE);
                        const int it = (int)(w.x & 0xffu);
                        const int shape = (int)ks.y;
                        const int depth = (int)S.params[RT_SHAPE_PARAMS * shape + 2];
                        const double t = rec[F_T];
                        if (COUNT) {
                            c.march_steps += w.y;
                            if (w.y > 2048) c.march_long_rays++;
                            if (w.y > c.march_max_evals) c.march_max_evals = w.y;
                        }
                        if (it >= depth && !(t < 0.001)) {   // finish() == DONE; ray_marching.rs:55 with max_t = +inf
                            if (t != t) {   // NaN candidate: replay now, and hide the entry from later kind passes
                                replay_brute(S, in, hq, pslot);
                                hq.mq_mask[entry] = 0;
                                retired = true;
                            } else if (t < best || (t == best && shape > winner)) {
                                best = t;
                                winner = shape;
                            }
                        }
                    }
                    bool started = false;
                    while (!retired && mask != 0 && !started) {
                        const int k = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int shape = S.march_index[k];
                        const double* q = S.params + RT_SHAPE_PARAMS * shape;
                        D3 ro = mk(in.ox[pslot], in.oy[pslot], in.oz[pslot]);
                        D3 rd = mk(in.dx[pslot], in.dy[pslot], in.dz[pslot]);
                        D3 o, d;
                        double start, end_c;
                        if (march_needed(S, S.inv + 12 * shape, q, ro, rd, best, o, d, start, end_c)) {
                            M m;
                            m.begin(q, o, d, start, end_c, S.march_G[k], S.march_F[k]);
                            if (COUNT) c.march_rays++;
                            if (m.skip_ok) {   // the model along the ray, once per (ray, shape): plan_begin's expansion
                                expand_ray<KIND>(q, m.p, m.d, m.t, m.end, (m.end - m.t) + 4.0 * m.step0, m.G, m.F, m.P);
                                m.n += 2;
#pragma unroll
                                for (int i = 0; i <= DEG; i++) rec[F_C + i] = m.P.c[i];
                                rec[F_P0] = m.P.p0.x; rec[F_P0 + 1] = m.P.p0.y; rec[F_P0 + 2] = m.P.p0.z;
                                rec[F_ERR0] = m.P.err0; rec[F_DRIFT1] = m.P.drift1;
                            }
                            m3_store_sample(rec, m);
                            rec[F_D] = m.d.x; rec[F_D + 1] = m.d.y; rec[F_D + 2] = m.d.z;
                            rec[F_START] = m.start; rec[F_END] = m.end;
                            m3_store_state(rec, m, false);
                            m3_set_u2(rec, F_SHAPE, (uint32_t)k, (uint32_t)shape);
                            s_ph[slot] = m3_phase_of(m.phase());
                            started = true;
                        }
                    }
                    if (started) {
                        rec[F_BEST] = best;
                        m3_set_u2(rec, F_MASK, mask, (uint32_t)winner);
                    } else {
                        if (!retired) {
                            hq.t[pslot] = best;
                            hq.index[pslot] = winner;
                        }
                        s_ph[slot] = M3_FREE;
                    }
                }
            }
        } else if (want == M3_PLAN) {
            // ---- jump planning: one hop per trip, a lane that finishes its ray picks up the next of the list ---
            M m;
            typename M::Plan pl;
            bool has = false, more_dummy;
            int slot = 0, next = 0;
            for (;;) {
                const unsigned need = __ballot_sync(FULL, !has);
                if (next < len && need) {
                    const int idx = next + __popc(need & lt);
                    if (!has && idx < len) {
                        slot = s_list[idx];
                        const double* rec = recs + slot * RT_M3_RS;
                        m3_load_core(S, rec, m, more_dummy);
                        const uint2 ks = m3_u2(rec, F_SHAPE);
                        m.G = S.march_G[ks.x];
                        m.F = S.march_F[ks.x];
                        m3_load_poly(rec, m);
                        m.plan_begin(pl);
                        has = true;
                    }
                    next = min(len, next + __popc(need));
                }
                if (__ballot_sync(FULL, has) == 0) break;
                if (has && !m.plan_hop(pl)) {
                    const long long mj = m.plan_end(pl);
                    double* rec = recs + slot * RT_M3_RS;
                    rec[F_M] = pl.M;
                    *reinterpret_cast<long long*>(rec + F_MJ) = mj;
                    m3_store_state(rec, m, pl.more);   // (cooldown / backoff when nothing can be skipped)
                    s_ph[slot] = mj > 0 ? M3_ADV : m3_phase_of(m.phase());
                    m3_add_prof(c, m);
                    has = false;
                }
            }
        } else if (want == M3_ADV) {
            // ---- the exact advance: 4 tasks per ray (t, p.x, p.y, p.z), one binade per trip, dynamic pick-up ---
            const int tasks = 4 * len;
            double res = 0.0, s = 0.0;
            long long left = 0;
            bool has = false;
            int slot = 0, comp = 0, next = 0;
            for (;;) {
                const unsigned need = __ballot_sync(FULL, !has);
                if (next < tasks && need) {
                    const int id = next + __popc(need & lt);
                    if (!has && id < tasks) {
                        slot = s_list[id >> 2];
                        comp = id & 3;
                        const double* rec = recs + slot * RT_M3_RS;
                        const double step = rec[F_STEP];
                        res = comp == 0 ? rec[F_T] : rec[F_P + comp - 1];
                        s = comp == 0 ? step : rec[F_D + comp - 1] * step;   // `step * dir` (Marcher::sd)
                        left = *reinterpret_cast<const long long*>(rec + F_MJ);
                        has = true;
                    }
                    next = min(tasks, next + __popc(need));
                }
                if (__ballot_sync(FULL, has) == 0) break;
                if (has) {
                    advance_iter(res, s, left);
                    if (left <= 0) {
                        recs[slot * RT_M3_RS + F_NT + comp] = res;
                        has = false;
                    }
                }
            }
            __syncwarp();
            for (int idx = lane; idx < len; idx += 32) s_ph[s_list[idx]] = M3_LAND;
        } else if (want == M3_LAND) {
            // ---- the reference's `r = next` at the landing sample + the model self-check --------------------
            for (int base = 0; base < len; base += 32) {
                const int idx = base + lane;
                if (idx < len) {
                    const int slot = s_list[idx];
                    double* rec = recs + slot * RT_M3_RS;
                    M m;
                    bool more;
                    m3_load_core(S, rec, m, more);
                    const double nt = rec[F_NT];
                    const D3 np = mk(rec[F_NP], rec[F_NP + 1], rec[F_NP + 2]);
                    const double Mb = rec[F_M];
                    const double land = surface_func<KIND>(m.q, np);   // attempt_land, with P evaluated at nt - t0
                    m.n++;
                    if (COUNT) c.march_prof[2]++;
                    const double x = nt - m.start;
                    double gp = rec[F_C + DEG];
#pragma unroll
                    for (int k = DEG - 1; k >= 0; k--) gp = fma(gp, x, rec[F_C + k]);
                    const bool same_sign = ((land > 0.0) == (m.r > 0.0)) && land != 0.0;
                    if (same_sign && fabs(land - gp) <= 0.25 * Mb && fabs(land) >= 0.5 * Mb) {
                        m.t = nt;
                        m.p = np;
                        m.r = land;
                        m.backoff = 4;
                        m.cooldown = more ? 0 : RT_MARCH_LAND_COOLDOWN;
                        m3_store_sample(rec, m);
                    } else {
                        m.skip_ok = false;   // the model does not describe this ray: finish it with the plain loop
                    }
                    m3_store_state(rec, m, false);
                    s_ph[slot] = m3_phase_of(m.phase());
                }
            }
        } else {
            // ---- the reference's literal steps: one per trip, dynamic pick-up, at most RT_M3_LIT_CAP trips -------
            M m;
            bool has = false, more_dummy;
            int slot = 0, next = 0;
            for (int trip = 0;; trip++) {
                const unsigned need = __ballot_sync(FULL, !has);
                if (next < len && need && trip < RT_M3_LIT_CAP) {
                    const int idx = next + __popc(need & lt);
                    if (!has && idx < len) {
                        slot = s_list[idx];
                        m3_load_core(S, recs + slot * RT_M3_RS, m, more_dummy);
                        has = true;
                    }
                    next = min(len, next + __popc(need));
                }
                if (__ballot_sync(FULL, has) == 0) break;
                if (has) {
                    m.literal();
                    const int ph = m.phase();
                    if (ph != RT_PHASE_LITERAL || trip + 1 >= RT_M3_LIT_CAP) {
                        double* rec = recs + slot * RT_M3_RS;
                        m3_store_sample(rec, m);
                        m3_store_state(rec, m, false);
                        s_ph[slot] = m3_phase_of(ph);
                        m3_add_prof(c, m);
                        if (COUNT) m.prof[0] = m.prof[1] = 0;
                        has = false;
                    }
                }
                if (trip + 1 >= RT_M3_LIT_CAP) break;
            }
        }
        __syncwarp();
    }
    if (COUNT) flush_counters(c, g_counters);
}

static size_t march3_smem() {
    return (size_t)RT_M3_WARPS * RT_M3_R * RT_M3_RS * sizeof(double) + (size_t)RT_M3_WARPS * 2 * RT_M3_R + 16;
}

template <int K_>
static int occupancy3() {
    const size_t smem = march3_smem();
    cudaFuncSetAttribute(k_march3<K_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_march3<K_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_march3<K_, false>, 32 * RT_M3_WARPS, smem);
    return b;
}

int rt_march3_occupancy(size_t* smem3) {
    *smem3 = march3_smem();
    int b = occupancy3<RT_SURF_HEART>();
    b = std::min(b, occupancy3<RT_SURF_SINE>());
    b = std::min(b, occupancy3<RT_SURF_STAR>());
    b = std::min(b, occupancy3<RT_SURF_DUPIN>());
    b = std::min(b, occupancy3<RT_SURF_HUNTS>());
    b = std::min(b, occupancy3<RT_SURF_CUSHION>());
    return std::max(b, 1);
}

template <int K_>
static void launch3(const MarchLaunch& ml) {
    if (ml.count)
        k_march3<K_, true><<<ml.grid3, 32 * RT_M3_WARPS, ml.smem3, ml.stream>>>(ml.ds, ml.kind_mask, ml.in, ml.hq, ml.march_count,
                                                                               ml.head, ml.counters);
    else
        k_march3<K_, false><<<ml.grid3, 32 * RT_M3_WARPS, ml.smem3, ml.stream>>>(ml.ds, ml.kind_mask, ml.in, ml.hq, ml.march_count,
                                                                                ml.head, ml.counters);
}

void rt_launch_march3(const MarchLaunch& ml) {
    switch (ml.kind) {
        case RT_SURF_HEART: launch3<RT_SURF_HEART>(ml); break;
        case RT_SURF_SINE: launch3<RT_SURF_SINE>(ml); break;
        case RT_SURF_STAR: launch3<RT_SURF_STAR>(ml); break;
        case RT_SURF_DUPIN: launch3<RT_SURF_DUPIN>(ml); break;
        case RT_SURF_HUNTS: launch3<RT_SURF_HUNTS>(ml); break;
        default: launch3<RT_SURF_CUSHION>(ml); break;
    }
}
