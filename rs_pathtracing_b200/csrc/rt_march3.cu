// rt_march3.cu -- K3 of the wavefront renderer, third generation: exact-skip marching with a POOL OF RAYS PER SM
// and a work queue per phase.
//
// What the profiles of k_march (one ray per lane, rt_march_kernels.cu) say (profiles/r2b_*): 7-13 of 32 lanes
// active.  Half of its warp instructions are the exact advance of a jump (advance_exact: an accumulator that
// crosses zero walks ~15 binades, the others one) at 2.5-6 lanes, a quarter the hop planning (1-100 hops per ray)
// at ~6 lanes: the lanes of a warp sit in different phases of their rays, and inside a phase the trip counts of
// the loops differ by an order of magnitude -- a warp executes the union of all of it.
//
// Here a ray is not tied to a lane, nor to a warp.  One persistent CTA per SM (16 warps) owns RT_M3_R ray RECORDS
// in shared memory (37 doubles each: the reference loop's state, the ray's polynomial model, the planned jump) and
// one QUEUE per phase a ray can wait for,
//     TRANS    finish a marched shape (candidate t against the path's best), start the next one (bound, first
//              sample, polynomial expansion along the ray) or retire the record and write the path's result
//     PLAN     how many iterations can provably be skipped (Marcher::plan_*: the hop loop)
//     ADV      the exact advance of t, p.x, p.y, p.z by that many steps (advance_iter): 4 tasks per ray
//     LAND     the reference's `r = next` at the landing sample + the model self-check
//     LIT      the reference's literal steps
// plus the queue of FREE records.  Every warp loops: take the fullest queue and run that phase's SERVICE -- 32 lanes
// work on 32 queue entries, and in the services whose work per entry varies (PLAN: hops, ADV: binades, LIT: steps)
// a lane that finishes its entry pops the next one at once (dynamic pick-up), so that a trip of the service loop
// runs with (nearly) all lanes as long as the queue has entries; a finished entry is pushed to the queue of its
// next phase, where any warp may pick it up.  A first version with one pool PER WARP (32-96 records) was measured at
// 7 lanes per instruction in the services: the lists were shorter than the warp (profiles/r2f_*); with ~670
// records per SM each queue holds hundreds.
//
// Synchronisation: the queues are rings in shared memory with atomic head / tail counters; an entry is the record
// number + 1, 0 = empty slot, written after a __threadfence_block() that orders the record's fields before it and
// cleared by the consumer.  No block barrier after start-up, no lock; the only waits are for a ring slot that a
// peer is in the middle of filling or clearing (a handful of instructions), each bounded by a spin limit that
// raises an error state (reported through rt_stats.verify_false_culls, which every test requires to be 0)
// instead of hanging.
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
#define RT_M3_R 672         // ray records per CTA (one CTA per SM): 672 x 296 B = 194 KB of the SM's 227 KB
#endif
#define RT_M3_RS 37         // doubles per record (odd: consecutive records start in different banks)
#ifndef RT_M3_THREADS
#define RT_M3_THREADS 512
#endif
#define RT_M3_QCAP 1024     // ring capacity of the record queues (power of two >= RT_M3_R)
#define RT_M3_ACAP 4096     // ring capacity of the ADV task queue (power of two >= 4 RT_M3_R)
#ifndef RT_M3_LIT_CAP
#define RT_M3_LIT_CAP 24    // literal steps per pick-up (a ray that needs thousands goes back to the end of the queue)
#endif
#define RT_M3_SPIN_LIMIT (1 << 22)
static_assert(RT_M3_R <= RT_M3_QCAP && 4 * RT_M3_R <= RT_M3_ACAP && 4 * RT_M3_R < 65535, "queue capacities");

enum { Q_FREE = 0, Q_TRANS = 1, Q_PLAN = 2, Q_LAND = 3, Q_LIT = 4, Q_ADV = 5, Q_COUNT = 6 };

// record layout (doubles)
enum {
    F_T = 0, F_R = 1, F_STEP = 2, F_P = 3, F_D = 6, F_START = 9, F_END = 10, F_BEST = 11,
    F_STATE = 12,   // (it | cooldown << 8 | backoff << 16 | flags << 24, evaluations)
    F_PATH = 13,    // (path slot, march-queue entry)
    F_MASK = 14,    // (remaining shape mask, winner)
    F_SHAPE = 15,   // (march-list index k | depth << 8, or 0xffffffff before the first shape; shape index)
    F_C = 16,       // P.c[0..6]
    F_P0 = 23,      // P.p0
    F_ERR0 = 26, F_DRIFT1 = 27,
    F_M = 28, F_MJ = 29,      // the planned jump: uncertainty band, number of iterations
    F_NT = 30, F_NP = 31,     // its landing sample
    F_STEP0 = 34, F_G = 35,   // the shape's step and gradient bound (so that no service starts with a global load)
    F_SPARE = 36,
};
#define M3_FLAG_SKIP_OK 1u
#define M3_FLAG_MORE 2u

// shared memory map (bytes from rt_smem_raw)
#define M3_OFF_RING16 (RT_M3_R * RT_M3_RS * 8)                       // 5 rings of RT_M3_QCAP uint16
#define M3_OFF_RINGA (M3_OFF_RING16 + 5 * RT_M3_QCAP * 2)            // ADV ring, RT_M3_ACAP uint16
#define M3_OFF_CTRL (M3_OFF_RINGA + RT_M3_ACAP * 2)                  // head[6], tail[6], live, exhausted, error, pad
#define M3_OFF_ADVCNT (M3_OFF_CTRL + 16 * 4)                         // int per record: ADV tasks still running
#define M3_SMEM_BYTES (M3_OFF_ADVCNT + RT_M3_R * 4)

// Everything in shared memory is addressed THROUGH THE EXTERN ARRAY ITSELF (not through derived pointers), so that
// the compiler keeps the address space and emits LDS / STS / ATOMS.
struct Rec {
    int base;   // index of the record's first double
    __device__ __forceinline__ double& operator[](int f) const { return reinterpret_cast<double*>(rt_smem_raw)[base + f]; }
    __device__ __forceinline__ uint2& u2(int f) const { return reinterpret_cast<uint2*>(rt_smem_raw)[base + f]; }
    __device__ __forceinline__ long long& i64(int f) const { return reinterpret_cast<long long*>(rt_smem_raw)[base + f]; }
};
#define REC(slot_) (Rec{(int)(slot_) * RT_M3_RS})
#define M3_CTRL(i_) (reinterpret_cast<unsigned*>(rt_smem_raw + M3_OFF_CTRL)[i_])
#define M3_VCTRL(i_) (reinterpret_cast<volatile unsigned*>(rt_smem_raw + M3_OFF_CTRL)[i_])
#define M3_HEAD(q_) (q_)
#define M3_TAIL(q_) (6 + (q_))
#define M3_LIVE 12
#define M3_EXHAUSTED 13
#define M3_ERROR 14
#define M3_ADVCNT(slot_) (reinterpret_cast<int*>(rt_smem_raw + M3_OFF_ADVCNT)[slot_])

__device__ __forceinline__ volatile unsigned short* m3_ring(int q) {
    return reinterpret_cast<volatile unsigned short*>(rt_smem_raw + (q == Q_ADV ? M3_OFF_RINGA : M3_OFF_RING16 + q * RT_M3_QCAP * 2));
}
__device__ __forceinline__ unsigned m3_ring_mask(int q) { return q == Q_ADV ? RT_M3_ACAP - 1 : RT_M3_QCAP - 1; }
__device__ __forceinline__ int m3_queue_size(int q) { return (int)(M3_VCTRL(M3_TAIL(q)) - M3_VCTRL(M3_HEAD(q))); }

// Warp-collective: the lanes with `wants` try to take one entry each from queue q.  Returns true for the lanes that
// got one (in `id`).  Never waits for entries that are not there.
__device__ __forceinline__ bool m3_pop(int q, bool wants, unsigned& id) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned need = __ballot_sync(FULL, wants);
    if (need == 0) return false;
    unsigned h = 0;
    int n = 0;
    if (lane == 0) {
        const int want_n = __popc(need);
        for (int tries = 0; tries < 4; tries++) {
            const unsigned hh = M3_VCTRL(M3_HEAD(q)), tt = M3_VCTRL(M3_TAIL(q));
            const int avail = (int)(tt - hh);
            if (avail <= 0) break;
            const int k = min(want_n, avail);
            if (atomicCAS(&M3_CTRL(M3_HEAD(q)), hh, hh + (unsigned)k) == hh) {
                h = hh;
                n = k;
                break;
            }
        }
    }
    n = __shfl_sync(FULL, n, 0);
    if (n == 0) return false;
    h = __shfl_sync(FULL, h, 0);
    const int rank = __popc(need & ((1u << lane) - 1u));
    const bool got = wants && rank < n;
    if (got) {
        volatile unsigned short* ring = m3_ring(q);
        const unsigned idx = (h + (unsigned)rank) & m3_ring_mask(q);
        unsigned v = ring[idx];
        for (int spin = 0; v == 0 && spin < RT_M3_SPIN_LIMIT; spin++) v = ring[idx];   // the producer is writing it
        if (v == 0) M3_VCTRL(M3_ERROR) = 1;
        ring[idx] = 0;
        id = v - 1u;
    }
    __threadfence_block();   // (acquire: the record's fields were written before the entry)
    return got;
}

// Warp-collective: the lanes with `valid` append `id` to queue q.
__device__ __forceinline__ void m3_push(int q, bool valid, unsigned id) {
    const unsigned FULL = 0xffffffffu;
    const unsigned m = __ballot_sync(FULL, valid);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    __threadfence_block();   // (release: this lane's writes to the record before the entry)
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&M3_CTRL(M3_TAIL(q)), (unsigned)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (valid) {
        volatile unsigned short* ring = m3_ring(q);
        const unsigned idx = (base + (unsigned)__popc(m & ((1u << lane) - 1u))) & m3_ring_mask(q);
        int spin = 0;
        while (ring[idx] != 0 && spin < RT_M3_SPIN_LIMIT) spin++;   // a consumer is still clearing the slot's last entry
        if (spin >= RT_M3_SPIN_LIMIT) M3_VCTRL(M3_ERROR) = 1;
        ring[idx] = (unsigned short)(id + 1u);
    }
}
// a ray that goes to ADV: its 4 tasks (record * 4 + component)
__device__ __forceinline__ void m3_push_adv(bool valid, unsigned slot) {
    const unsigned FULL = 0xffffffffu;
    const unsigned m = __ballot_sync(FULL, valid);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    if (valid) M3_ADVCNT(slot) = 4;
    __threadfence_block();
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&M3_CTRL(M3_TAIL(Q_ADV)), 4u * (unsigned)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (valid) {
        volatile unsigned short* ring = m3_ring(Q_ADV);
        const unsigned first = base + 4u * (unsigned)__popc(m & ((1u << lane) - 1u));
#pragma unroll
        for (unsigned c = 0; c < 4; c++) {
            const unsigned idx = (first + c) & (RT_M3_ACAP - 1);
            int spin = 0;
            while (ring[idx] != 0 && spin < RT_M3_SPIN_LIMIT) spin++;
            if (spin >= RT_M3_SPIN_LIMIT) M3_VCTRL(M3_ERROR) = 1;
            ring[idx] = (unsigned short)(slot * 4u + c + 1u);
        }
    }
}

__device__ __forceinline__ int m3_queue_of(int ph) {   // Marcher::phase() -> the queue of the service that continues
    return ph == RT_PHASE_ATTEMPT ? Q_PLAN : ph == RT_PHASE_LITERAL ? Q_LIT : Q_TRANS;
}

// the loop state literal() / phase() / land touch
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_load_core(const DevScene& S, Rec rec, Marcher<KIND, COUNT>& m, bool& more) {
    m.t = rec[F_T]; m.r = rec[F_R]; m.step = rec[F_STEP];
    m.p = mk(rec[F_P], rec[F_P + 1], rec[F_P + 2]);
    m.d = mk(rec[F_D], rec[F_D + 1], rec[F_D + 2]);
    m.start = rec[F_START]; m.end = rec[F_END];
    const uint2 w = rec.u2(F_STATE), ks = rec.u2(F_SHAPE);
    m.it = (int)(w.x & 0xffu);
    m.cooldown = (int)((w.x >> 8) & 0xffu);
    m.backoff = (int)((w.x >> 16) & 0xffu);
    m.skip_ok = ((w.x >> 24) & M3_FLAG_SKIP_OK) != 0;
    more = ((w.x >> 24) & M3_FLAG_MORE) != 0;
    m.have_poly = true;   // (expanded when the shape is started; only read when skip_ok)
    m.plan_miss_ok = false;   // t lives in the record: only begin()'s hull proof declares misses here
    m.local_model = false;    // (the record keeps the model of the chord's start: a refinement plan re-expands every time)
    m.n = w.y;
    m.q = S.params + RT_SHAPE_PARAMS * (int)ks.y;
    m.step0 = rec[F_STEP0];
    m.depth = (int)((ks.x >> 8) & 0xffu);
    m.G = rec[F_G];
    m.sd = m.step * m.d;
    if (COUNT) m.prof[0] = m.prof[1] = m.prof[2] = m.prof[3] = m.prof[4] = m.prof[5] = m.prof[6] = m.prof[7] = 0;
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_store_state(Rec rec, const Marcher<KIND, COUNT>& m, bool more) {
    const uint32_t flags = (m.skip_ok ? M3_FLAG_SKIP_OK : 0u) | (more ? M3_FLAG_MORE : 0u);
    const uint32_t nn = m.n > 0xffffffffull ? 0xffffffffu : (uint32_t)m.n;
    rec.u2(F_STATE) = make_uint2((uint32_t)m.it | ((uint32_t)m.cooldown << 8) | ((uint32_t)m.backoff << 16) | (flags << 24), nn);
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_store_sample(Rec rec, const Marcher<KIND, COUNT>& m) {
    rec[F_T] = m.t; rec[F_R] = m.r; rec[F_STEP] = m.step;
    rec[F_P] = m.p.x; rec[F_P + 1] = m.p.y; rec[F_P + 2] = m.p.z;
}
template <int KIND, bool COUNT>
__device__ __forceinline__ void m3_load_poly(Rec rec, Marcher<KIND, COUNT>& m) {
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
        for (int k = 0; k < 8; k++) c.march_prof[k] += m.prof[k];
    }
}

template <int KIND, bool COUNT>
__global__ void __launch_bounds__(RT_M3_THREADS, 1)
k_march3(DevScene S, uint32_t kind_mask, PathQueue in, HitQueue hq, const uint32_t* __restrict__ march_count,
         uint32_t* head, DevCounters* g_counters) {
    typedef Marcher<KIND, COUNT> M;
    constexpr int DEG = M::DEG;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    DevCounters c = {};
    const uint32_t n = *march_count;
    // ---- start-up: every record is free ----------------------------------------------------------------------
    for (int i = threadIdx.x; i < 5 * RT_M3_QCAP; i += RT_M3_THREADS)
        reinterpret_cast<unsigned short*>(rt_smem_raw + M3_OFF_RING16)[i] = i < RT_M3_R ? (unsigned short)(i + 1) : 0;
    for (int i = threadIdx.x; i < RT_M3_ACAP; i += RT_M3_THREADS) reinterpret_cast<unsigned short*>(rt_smem_raw + M3_OFF_RINGA)[i] = 0;
    if (threadIdx.x < 16) M3_CTRL(threadIdx.x) = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        M3_CTRL(M3_TAIL(Q_FREE)) = RT_M3_R;
        M3_CTRL(M3_EXHAUSTED) = n == 0 ? 1u : 0u;
    }
    __syncthreads();

    unsigned idle_rounds = 0;
    for (;;) {
        // ---- the fullest queue (lane q reads queue q's length) -----------------------------------------------
        int my_size = lane < Q_COUNT ? m3_queue_size(lane) : 0;
        if (lane == Q_ADV) my_size = (my_size + 3) >> 2;          // tasks -> rays
        const int n_free = __shfl_sync(FULL, my_size, Q_FREE);
        const bool exhausted = M3_VCTRL(M3_EXHAUSTED) != 0;
        int want = Q_TRANS, want_size = __shfl_sync(FULL, my_size, Q_TRANS);
#pragma unroll
        for (int q = Q_PLAN; q < Q_COUNT; q++) {
            const int sz = __shfl_sync(FULL, my_size, q);
            if (sz > want_size) {
                want_size = sz;
                want = q;
            }
        }
        // ---- refill free records from the march queue: whenever a warp's worth is free, or nothing else to do ---
        if (!exhausted && n_free > 0 && (n_free >= 32 || want_size == 0)) {
            unsigned slot = 0;
            const bool got = m3_pop(Q_FREE, true, slot);
            const unsigned gm = __ballot_sync(FULL, got);
            if (gm != 0) {
                const int k = __popc(gm);
                uint32_t base = 0;
                if (lane == 0) {
                    atomicAdd(&M3_CTRL(M3_LIVE), (unsigned)k);
                    base = atomicAdd(head, (uint32_t)k);
                }
                base = __shfl_sync(FULL, base, 0);
                bool filled = false;
                if (got) {
                    const uint32_t j = base + (uint32_t)__popc(gm & ((1u << lane) - 1u));
                    if (j < n) {
                        const uint32_t mask = hq.mq_mask[j] & kind_mask;
                        if (mask != 0) {
                            const uint32_t pslot = hq.mq_slot[j];
                            const Rec rec = REC(slot);
                            rec[F_BEST] = hq.t[pslot];
                            rec.u2(F_PATH) = make_uint2(pslot, j);
                            rec.u2(F_MASK) = make_uint2(mask, (uint32_t)hq.index[pslot]);
                            rec.u2(F_SHAPE) = make_uint2(0xffffffffu, 0u);
                            filled = true;
                        }
                    }
                }
                const unsigned back = __ballot_sync(FULL, got && !filled);
                if (base + (uint32_t)k >= n && lane == 0) M3_VCTRL(M3_EXHAUSTED) = 1;
                m3_push(Q_TRANS, filled, slot);
                m3_push(Q_FREE, got && !filled, slot);
                if (back != 0 && lane == 0) atomicSub(&M3_CTRL(M3_LIVE), (unsigned)__popc(back));
            }
            idle_rounds = 0;
            continue;
        }
        if (want_size == 0) {
            if (exhausted && M3_VCTRL(M3_LIVE) == 0) break;
            if (M3_VCTRL(M3_ERROR) != 0 || ++idle_rounds > (1u << 24)) {   // never hang: give up loudly
                M3_VCTRL(M3_ERROR) = 1;
                break;
            }
            __nanosleep(400);
            continue;
        }
        idle_rounds = 0;

        if (want == Q_TRANS) {
            // ---- finish a marched shape / start the next one / retire the record ---------------------------
            for (;;) {
                unsigned slot = 0;
                const bool got = m3_pop(Q_TRANS, true, slot);
                if (__ballot_sync(FULL, got) == 0) break;
                int dest = -1;   // queue the record goes to next
                if (got) {
                    const Rec rec = REC(slot);
                    const uint2 pe = rec.u2(F_PATH), mw = rec.u2(F_MASK), ks = rec.u2(F_SHAPE);
                    const uint32_t pslot = pe.x, entry = pe.y;
                    uint32_t mask = mw.x;
                    int winner = (int)mw.y;
                    double best = rec[F_BEST];
                    bool retired = false;
                    if (ks.x != 0xffffffffu) {   // a shape has just been marched to its end
                        const uint2 w = rec.u2(F_STATE);
                        const int it = (int)(w.x & 0xffu);
                        const int shape = (int)ks.y;
                        const int depth = (int)((ks.x >> 8) & 0xffu);
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
                    while (!retired && mask != 0 && dest < 0) {
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
                            if (m.skip_ok) {   // the model along the ray, once per (ray, shape): expanded by begin()
#pragma unroll
                                for (int i = 0; i <= DEG; i++) rec[F_C + i] = m.P.c[i];
                                rec[F_P0] = m.P.p0.x; rec[F_P0 + 1] = m.P.p0.y; rec[F_P0 + 2] = m.P.p0.z;
                                rec[F_ERR0] = m.P.err0; rec[F_DRIFT1] = m.P.drift1;
                            }
                            m3_store_sample(rec, m);
                            rec[F_D] = m.d.x; rec[F_D + 1] = m.d.y; rec[F_D + 2] = m.d.z;
                            rec[F_START] = m.start; rec[F_END] = m.end;
                            rec[F_STEP0] = m.step0; rec[F_G] = m.G;
                            m3_store_state(rec, m, false);
                            rec.u2(F_SHAPE) = make_uint2((uint32_t)k | ((uint32_t)(m.depth & 0xff) << 8), (uint32_t)shape);
                            dest = m3_queue_of(m.phase());
                        }
                    }
                    if (dest >= 0) {
                        rec[F_BEST] = best;
                        rec.u2(F_MASK) = make_uint2(mask, (uint32_t)winner);
                    } else {
                        if (!retired) {
                            hq.t[pslot] = best;
                            hq.index[pslot] = winner;
                        }
                        dest = Q_FREE;
                    }
                }
                const unsigned freed = __ballot_sync(FULL, dest == Q_FREE);
                m3_push(Q_PLAN, dest == Q_PLAN, slot);
                m3_push(Q_LIT, dest == Q_LIT, slot);
                m3_push(Q_TRANS, dest == Q_TRANS, slot);
                m3_push(Q_FREE, dest == Q_FREE, slot);
                if (freed != 0 && lane == 0) atomicSub(&M3_CTRL(M3_LIVE), (unsigned)__popc(freed));
            }
        } else if (want == Q_PLAN) {
            // ---- jump planning: one hop per trip; a lane that finishes its ray pops the next one --------------
            M m;
            typename M::Plan pl;
            bool has = false, more_dummy;
            unsigned slot = 0;
            for (;;) {
                unsigned id = 0;
                if (m3_pop(Q_PLAN, !has, id)) {
                    slot = id;
                    const Rec rec = REC(slot);
                    m3_load_core(S, rec, m, more_dummy);
                    m3_load_poly(rec, m);
                    m.plan_begin(pl);
                    has = true;
                }
                if (__ballot_sync(FULL, has) == 0) break;
                int dest = -1;
                if (has && !m.plan_hop(pl)) {
                    const long long mj = m.plan_end(pl);
                    const Rec rec = REC(slot);
                    rec[F_M] = pl.M;
                    rec.i64(F_MJ) = mj;
                    m3_store_state(rec, m, pl.more);   // (cooldown / backoff when nothing can be skipped)
                    dest = mj > 0 ? Q_ADV : m3_queue_of(m.phase());
                    m3_add_prof(c, m);
                    has = false;
                }
                if (__ballot_sync(FULL, dest >= 0) != 0) {
                    m3_push_adv(dest == Q_ADV, slot);
                    m3_push(Q_LIT, dest == Q_LIT, slot);
                    m3_push(Q_TRANS, dest == Q_TRANS, slot);
                }
            }
        } else if (want == Q_ADV) {
            // ---- the exact advance: one task = one accumulator of one ray, one binade per trip ----------------
            double res = 0.0, s = 0.0;
            long long left = 0;
            bool has = false;
            unsigned slot = 0, comp = 0;
            for (;;) {
                unsigned id = 0;
                if (m3_pop(Q_ADV, !has, id)) {
                    slot = id >> 2;
                    comp = id & 3u;
                    const Rec rec = REC(slot);
                    const double step = rec[F_STEP];
                    res = comp == 0 ? rec[F_T] : rec[F_P + comp - 1];
                    s = comp == 0 ? step : rec[F_D + comp - 1] * step;   // `step * dir` (Marcher::sd)
                    left = rec.i64(F_MJ);
                    has = true;
                }
                if (__ballot_sync(FULL, has) == 0) break;
                bool last = false;
                if (has) {
                    advance_iter(res, s, left);
                    if (left <= 0) {
                        REC(slot)[F_NT + comp] = res;
                        __threadfence_block();
                        last = atomicSub(&M3_ADVCNT(slot), 1) == 1;   // the ray's fourth accumulator has arrived
                        has = false;
                    }
                }
                m3_push(Q_LAND, last, slot);
            }
        } else if (want == Q_LAND) {
            // ---- the reference's `r = next` at the landing sample + the model self-check --------------------
            for (;;) {
                unsigned slot = 0;
                const bool got = m3_pop(Q_LAND, true, slot);
                if (__ballot_sync(FULL, got) == 0) break;
                int dest = -1;
                if (got) {
                    const Rec rec = REC(slot);
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
                    dest = m3_queue_of(m.phase());
                }
                m3_push(Q_PLAN, dest == Q_PLAN, slot);
                m3_push(Q_LIT, dest == Q_LIT, slot);
                m3_push(Q_TRANS, dest == Q_TRANS, slot);
            }
        } else {
            // ---- the reference's literal steps: one per trip, at most RT_M3_LIT_CAP per pick-up ------------------
            M m;
            bool has = false, more_dummy;
            unsigned slot = 0;
            int steps = 0;
            for (;;) {
                unsigned id = 0;
                if (m3_pop(Q_LIT, !has, id)) {
                    slot = id;
                    m3_load_core(S, REC(slot), m, more_dummy);
                    steps = 0;
                    has = true;
                }
                if (__ballot_sync(FULL, has) == 0) break;
                int dest = -1;
                if (has) {
                    m.literal();
                    const int ph = m.phase();
                    if (ph != RT_PHASE_LITERAL || ++steps >= RT_M3_LIT_CAP) {
                        const Rec rec = REC(slot);
                        m3_store_sample(rec, m);
                        m3_store_state(rec, m, false);
                        dest = m3_queue_of(ph);
                        m3_add_prof(c, m);
                        has = false;
                    }
                }
                if (__ballot_sync(FULL, dest >= 0) != 0) {
                    m3_push(Q_PLAN, dest == Q_PLAN, slot);
                    m3_push(Q_TRANS, dest == Q_TRANS, slot);
                    m3_push(Q_LIT, dest == Q_LIT, slot);
                }
            }
        }
    }
    // loud: every test requires verify_false_culls == 0
    if (M3_VCTRL(M3_ERROR) != 0 && threadIdx.x == 0) atomicAdd(&g_counters->verify_false_culls, 1ull << 40);
    if (COUNT) flush_counters(c, g_counters);
}

template <int K_>
static void prepare3() {
    cudaFuncSetAttribute(k_march3<K_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3_SMEM_BYTES);
    cudaFuncSetAttribute(k_march3<K_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3_SMEM_BYTES);
}

int rt_march3_occupancy(size_t* smem3) {
    *smem3 = M3_SMEM_BYTES;
    prepare3<RT_SURF_HEART>();
    prepare3<RT_SURF_SINE>();
    prepare3<RT_SURF_STAR>();
    prepare3<RT_SURF_DUPIN>();
    prepare3<RT_SURF_HUNTS>();
    prepare3<RT_SURF_CUSHION>();
    return 1;   // one persistent CTA per SM
}

template <int K_>
static void launch3(const MarchLaunch& ml) {
    if (ml.count)
        k_march3<K_, true><<<ml.grid3, RT_M3_THREADS, ml.smem3, ml.stream>>>(ml.ds, ml.kind_mask, ml.in, ml.hq, ml.march_count,
                                                                            ml.head, ml.counters);
    else
        k_march3<K_, false><<<ml.grid3, RT_M3_THREADS, ml.smem3, ml.stream>>>(ml.ds, ml.kind_mask, ml.in, ml.hq, ml.march_count,
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
