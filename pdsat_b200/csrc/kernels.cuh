// kernels.cuh — sm_100a kernels of the PDSAT hot path.
//
//   fill_words_kernel   sample values -> bit-sliced assumption words (one 64-bit word per
//                       (64-sample group, set position): bit s = value in sample s)
//   mt19937_bits_kernel boost/std mt19937 + bool_rand (src_common/addit_func.cpp:193-196)
//                       as a packed bit stream, 624 draws per round per stream segment
//   bcp_kernel          bit-sliced unit propagation: one warp owns one group of 64 samples,
//                       the assignment lives in shared memory as one 64-bit plane per
//                       literal, lanes evaluate different clauses of an occurrence list
//                       (semantics of Solver::propagate, src_common/minisat/core/Solver.cc:599-662,
//                       to the fixpoint / first conflict)
//
// No tensor cores: nothing here is a contraction (integer / bit work bound by on-chip
// memory and the LSU).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pdsat {

// ------------------------------------------------------------------ device views
struct DbDev {
    const uint4* rec;         // 16-byte clause records: 8 x uint16 FALSE-plane codes
    const uint32_t* occ_off;  // [2*n_live+1]
    const uint32_t* occ;      // first-record index of every clause containing the literal
    int32_t n_live;
    int32_t n_lit;            // 2*n_live
    int32_t n_base_assigned;
    int32_t unsat;            // level-0 conflict: every sample is a conflict
    int32_t count_bits;       // bits needed to count to n_live
};

struct PointDev {
    uint64_t n_samples;
    uint64_t group_start;  // first global group of this point
    uint64_t word_off;     // offset of its assumption words
    uint64_t seed;
    uint64_t first_sample;
    uint64_t stream_off;   // offset (32-bit words) of its packed value stream
    uint64_t stream_skip;  // leading bits of the stream to skip
    uint32_t k;
    uint32_t var_off;      // offset into the set-code table
    uint32_t mode;
    uint32_t n_assumed_live;  // set variables that are not fixed at level 0
};

static constexpr uint32_t SETCODE_FIXED = 0x80000000u;  // set var already assigned at level 0; bit0 = its lbool

struct BatchDev {
    const PointDev* points;
    int32_t n_points;
    uint64_t n_groups;
    uint64_t n_words;
    const uint32_t* setcode;
    uint64_t* words;
    const uint32_t* stream;
    unsigned long long* cost;   // n_points * 8
    uint64_t* conflict_bits;    // per global group (optional)
    uint64_t* full_bits;        // per global group (optional)
    uint16_t* trail;            // per global group * 64 (optional)
    uint64_t* assign;           // per global group * n_lit (optional)
    unsigned long long* model_count;
    uint32_t* model_bits;       // max_models * model_words
    uint32_t* model_point;
    uint64_t* model_sample;
    int32_t max_models;
    int32_t model_words;
};

__device__ __forceinline__ int find_point(const PointDev* pts, int n, uint64_t g) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pts[mid].group_start <= g) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// ------------------------------------------------------------------ fill
// One thread per assumption word.  Word (group gi of point p, set position j) collects the
// value of set variable j in samples 64*gi .. 64*gi+63.
__global__ void fill_words_kernel(BatchDev b) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= b.n_words) return;
    // locate the point by word offset
    int lo = 0, hi = b.n_points - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (b.points[mid].word_off <= w) lo = mid;
        else hi = mid - 1;
    }
    const PointDev& pt = b.points[lo];
    uint64_t rel = w - pt.word_off;
    uint64_t gi = rel / pt.k;
    uint32_t j = (uint32_t)(rel - gi * pt.k);
    uint64_t s0 = gi * 64;
    uint64_t out = 0;
    if (pt.mode == 1) {  // PDSAT_MODE_RANDOM_COUNTER
        uint64_t g = pt.first_sample / 64 + gi;
        out = splitmix64(pt.seed + 0x9E3779B97F4A7C15ull * (g * 65536ull + j + 1));
    } else if (pt.mode == 2) {  // PDSAT_MODE_ENUM
        uint64_t lint = pt.first_sample + s0;
        if (j < 64) {
#pragma unroll 8
            for (int s = 0; s < 64; s++) out |= (((lint + s) >> j) & 1ull) << s;
        }
    } else {  // packed sample-major value stream (mt19937 bits or caller's explicit values)
        const uint32_t* st = b.stream + pt.stream_off;
        uint64_t n = pt.n_samples - s0;
        int ns = n < 64 ? (int)n : 64;
        uint64_t d = pt.stream_skip + s0 * pt.k + j;
        for (int s = 0; s < ns; s++, d += pt.k) out |= (uint64_t)((__ldg(&st[d >> 5]) >> (d & 31)) & 1u) << s;
    }
    b.words[w] = out;
}

// ------------------------------------------------------------------ mt19937
struct MtSegment {
    uint64_t out_off;   // 32-bit word offset of this segment's first output bit (round aligned)
    uint64_t n_rounds;  // rounds of 624 draws to produce (always even)
    uint32_t state_idx; // slot of the generator state that precedes the first round
    uint32_t pad;
};

// One jump-ahead: states[dst] = g(A) states[src], g = polys[poly] (x^(624*2^t) mod phi as
// 19968 bits).  out[w] = XOR over the set bits i of g of x[i + w], where x[0..623] is the
// source window and x continues by the mt19937 recurrence (mt_jump.hpp).
struct MtJumpTask {
    uint32_t src, dst, poly, pad;
};

static constexpr int MT_JUMP_THREADS = 640;
static constexpr int MT_JUMP_ROUNDS = 33;  // 33 * 624 = 20592 >= 19937 + 623 words
static constexpr int MT_JUMP_SMEM = (MT_JUMP_ROUNDS * 624 + 624) * 4;

__global__ void __launch_bounds__(MT_JUMP_THREADS) mt19937_jump_kernel(const uint32_t* __restrict__ polys,
                                                                       uint32_t* __restrict__ states,
                                                                       const MtJumpTask* __restrict__ tasks) {
    extern __shared__ uint32_t jsm[];
    uint32_t* x = jsm;                         // MT_JUMP_ROUNDS * 624 words of the sequence
    uint32_t* g = jsm + MT_JUMP_ROUNDS * 624;  // 624 words of polynomial bits
    const int tid = threadIdx.x;
    const MtJumpTask t = tasks[blockIdx.x];
    const uint32_t* src = states + (size_t)t.src * 624;
    const uint32_t* gp = polys + (size_t)t.poly * 624;
    if (tid < 624) {
        x[tid] = src[tid];
        g[tid] = gp[tid];
    }
    __syncthreads();
    // run the recurrence forward: x[n+624] = x[n+397] ^ twist(x[n], x[n+1])
    for (int r = 0; r < MT_JUMP_ROUNDS - 1; r++) {
        uint32_t* cur = x + r * 624;
        const int base[3] = {0, 227, 454};
        const int cnt[3] = {227, 227, 170};
#pragma unroll
        for (int ph = 0; ph < 3; ph++) {
            if (tid < cnt[ph]) {
                const int n = base[ph] + tid;
                const uint32_t y = (cur[n] & 0x80000000u) | (cur[n + 1] & 0x7fffffffu);
                cur[n + 624] = cur[n + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncthreads();
        }
    }
    if (tid < 624) {
        uint32_t acc = 0;
        for (int c = 0; c < 624; c++) {
            uint32_t gw = g[c];
            const uint32_t* xb = x + c * 32 + tid;
            while (gw) {
                const int b = __ffs(gw) - 1;
                gw &= gw - 1;
                acc ^= xb[b];
            }
        }
        states[(size_t)t.dst * 624 + tid] = acc;
    }
}

// One CTA (256 threads) per stream segment.  states[state_base + blockIdx.x] is the 624-word
// generator state *before* the twist that produces the segment's first round.  Emits, per
// draw, the bool_rand bit ((tempered >> 31) == 0), packed little-endian into 32-bit words.
__global__ void __launch_bounds__(256) mt19937_bits_kernel(const uint32_t* __restrict__ states,
                                                           const MtSegment* __restrict__ segs,
                                                           uint32_t* __restrict__ out) {
    __shared__ uint32_t st[624];
    __shared__ uint32_t pairw[40];
    const int tid = threadIdx.x;
    const MtSegment seg = segs[blockIdx.x];
    const uint32_t* s0 = states + (size_t)seg.state_idx * 624;
    for (int i = tid; i < 624; i += 256) st[i] = s0[i];
    if (tid < 40) pairw[tid] = 0;
    __syncthreads();
    uint32_t* dst = out + seg.out_off;
    for (uint64_t r = 0; r < seg.n_rounds; r++) {
        // twist: x[i] <- x[i+397] ^ twist(x[i], x[i+1]) in three dependency-free phases
        const int base[3] = {0, 227, 454};
        const int cnt[3] = {227, 227, 170};
#pragma unroll
        for (int ph = 0; ph < 3; ph++) {
            uint32_t nv = 0;
            int i = base[ph] + tid;
            if (tid < cnt[ph]) {
                uint32_t y = (st[i] & 0x80000000u) | (st[(i + 1) % 624] & 0x7fffffffu);
                nv = st[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncthreads();
            if (tid < cnt[ph]) st[i] = nv;
            __syncthreads();
        }
        // temper + bool_rand bit, 32 draws per ballot
        const int odd = (int)(r & 1);
#pragma unroll
        for (int m = 0; m < 3; m++) {
            int t = tid + 256 * m;
            uint32_t bit = 0;
            if (t < 624) {
                uint32_t y = st[t];
                y ^= y >> 11;
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= y >> 18;
                bit = (y >> 31) == 0;
            }
            uint32_t word = __ballot_sync(0xffffffffu, bit);
            if ((tid & 31) == 0 && t < 624) {
                int c = t >> 5;  // chunk 0..19
                if (!odd) {
                    if (c < 19) pairw[c] = word;
                    else atomicOr(&pairw[19], word & 0xffffu);
                } else {
                    atomicOr(&pairw[19 + c], word << 16);
                    if (c < 19) atomicOr(&pairw[20 + c], word >> 16);
                }
            }
        }
        __syncthreads();
        if (odd) {
            if (tid < 39) {
                dst[(r >> 1) * 39 + tid] = pairw[tid];
                pairw[tid] = 0;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ BCP
struct WarpCtx {
    uint64_t* S;        // [n_lit + 2] one plane per literal: samples where it is TRUE
    uint16_t* queue;    // ring of literals whose occurrence lists must be visited
    uint32_t* inq;      // bitset: literal currently queued
    uint64_t* dead;     // samples in conflict
    uint32_t* qtail;
    uint32_t* implied;  // set when BCP implied anything in this group
    uint32_t qmask;
};

__device__ __forceinline__ void push_lit(const WarpCtx& c, uint32_t l) {
    uint32_t bit = 1u << (l & 31);
    uint32_t old = atomicOr(&c.inq[l >> 5], bit);
    if (!(old & bit)) {
        uint32_t pos = atomicAdd(c.qtail, 1u);
        c.queue[pos & c.qmask] = (uint16_t)l;
    }
}

// Evaluate one clause on the 64 samples of the group.  Slots hold the code of each
// literal's FALSE plane.  ge1 / ge2 = samples with >= 1 / >= 2 non-false literals.
__device__ __forceinline__ void visit_clause(const uint4* __restrict__ rec, uint32_t ci, const WarpCtx& c,
                                             uint64_t live) {
    uint64_t ge1 = 0, ge2 = 0;
    uint32_t cur = ci;
    bool more;
    do {
        more = false;
        const uint4 r = __ldg(&rec[cur]);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t m = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            if (m >= 0xFFFEu) {
                if (m == 0xFFFEu) {
                    more = true;
                    cur++;
                }
                break;
            }
            uint64_t nf = ~c.S[m];
            ge2 |= ge1 & nf;
            ge1 |= nf;
        }
    } while (more);
    const uint64_t confl = ~ge1 & live;
    const uint64_t unit = ge1 & ~ge2 & live;
    if (confl) atomicOr((unsigned long long*)c.dead, (unsigned long long)confl);
    if (unit) {
        // rare path: find, per sample, the one non-false literal and make it true
        cur = ci;
        do {
            more = false;
            const uint4 r = __ldg(&rec[cur]);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t m = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
                if (m >= 0xFFFEu) {
                    if (m == 0xFFFEu) {
                        more = true;
                        cur++;
                    }
                    break;
                }
                uint64_t imp = unit & ~c.S[m] & ~c.S[m ^ 1];
                if (imp) {
                    uint64_t old = atomicOr((unsigned long long*)&c.S[m ^ 1], (unsigned long long)imp);
                    if (imp & ~old) {
                        push_lit(c, m ^ 1);
                        *c.implied = 1;
                    }
                }
            }
        } while (more);
    }
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src) << 32) |
           __shfl_sync(0xffffffffu, (uint32_t)v, src);
}
__device__ __forceinline__ uint64_t shfl_down64(uint64_t v, int d) {
    return ((uint64_t)__shfl_down_sync(0xffffffffu, (uint32_t)(v >> 32), d) << 32) |
           __shfl_down_sync(0xffffffffu, (uint32_t)v, d);
}

static constexpr int MAX_COUNT_BITS = 13;

struct WarpAcc {
    uint64_t n, n_conflict, n_full, sum, sumsq, n_impl;
    int point;
};

__device__ __forceinline__ void flush_acc(WarpAcc& a, unsigned long long* cost, int lane) {
    if (a.point >= 0 && lane == 0 && a.n) {
        unsigned long long* c = cost + (size_t)a.point * 8;
        atomicAdd(&c[0], (unsigned long long)a.n);
        if (a.n_conflict) atomicAdd(&c[1], (unsigned long long)a.n_conflict);
        if (a.n_full) atomicAdd(&c[2], (unsigned long long)a.n_full);
        atomicAdd(&c[3], (unsigned long long)a.sum);
        atomicAdd(&c[4], (unsigned long long)a.sumsq);
        if (a.n_impl) atomicAdd(&c[5], (unsigned long long)a.n_impl);
    }
    a.n = a.n_conflict = a.n_full = a.sum = a.sumsq = a.n_impl = 0;
}

// Persistent kernel: grid = #SMs x CTAs/SM, each warp strides over the groups.
__global__ void __launch_bounds__(1024, 1)
bcp_kernel(DbDev db, BatchDev b, int warp_bytes, int qcap) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    unsigned char* base = smem + (size_t)warp * warp_bytes;
    WarpCtx c;
    c.S = (uint64_t*)base;
    const int n_planes = db.n_lit + 2;
    c.queue = (uint16_t*)(c.S + n_planes);
    c.inq = (uint32_t*)(c.queue + qcap);
    const int inq_words = (db.n_lit + 31) >> 5;
    c.dead = (uint64_t*)(c.inq + ((inq_words + 1) & ~1));
    c.qtail = (uint32_t*)(c.dead + 1);
    c.implied = c.qtail + 1;
    c.qmask = (uint32_t)qcap - 1;

    WarpAcc acc;
    acc.n = acc.n_conflict = acc.n_full = acc.sum = acc.sumsq = acc.n_impl = 0;
    acc.point = -1;

    const uint64_t stride = (uint64_t)gridDim.x * wpc;
    for (uint64_t g = (uint64_t)blockIdx.x * wpc + warp; g < b.n_groups; g += stride) {
        const int p = find_point(b.points, b.n_points, g);
        if (p != acc.point) {
            flush_acc(acc, b.cost, lane);
            acc.point = p;
        }
        const PointDev* pt = &b.points[p];
        const uint64_t gi = g - pt->group_start;
        const uint64_t s0 = gi * 64;
        const uint64_t left = pt->n_samples - s0;
        const uint64_t valid = left >= 64 ? ~0ull : ((1ull << left) - 1);
        const uint32_t k = pt->k;

        // ---- reset the group's state
        for (int i = lane; i < n_planes; i += 32) c.S[i] = 0;
        for (int i = lane; i < inq_words; i += 32) c.inq[i] = 0;
        if (lane == 0) {
            *c.dead = db.unsat ? valid : 0;
            *c.qtail = 0;
            *c.implied = 0;
        }
        __syncwarp();

        // ---- assumptions (src_mpi/mpi_predicter.cpp:3236-3242: true bit => positive literal)
        if (!db.unsat) {
            const uint64_t* w = b.words + pt->word_off + gi * k;
            const uint32_t* code = b.setcode + pt->var_off;
            for (uint32_t j = lane; j < k; j += 32) {
                const uint32_t cd = code[j];
                const uint64_t bits = w[j];
                if (cd & SETCODE_FIXED) {
                    // variable already assigned at level 0: a contradicting value is a
                    // conflict (addClause_ drops the false literal: empty clause)
                    const uint64_t bad = ((cd & 1u) ? bits : ~bits) & valid;
                    if (bad) atomicOr((unsigned long long*)c.dead, (unsigned long long)bad);
                } else {
                    const uint64_t tp = bits & valid, tn = ~bits & valid;
                    c.S[cd] = tp;
                    c.S[cd ^ 1] = tn;
                    if (tp) push_lit(c, cd);
                    if (tn) push_lit(c, cd ^ 1);
                }
            }
        }
        __syncwarp();

        // ---- propagate to the fixpoint
        uint32_t qh = 0;
        while (true) {
            const uint32_t qt = *(volatile uint32_t*)c.qtail;
            const uint64_t deadv = *(volatile uint64_t*)c.dead;
            if (qh == qt || (deadv & valid) == valid) break;
            const uint32_t l = c.queue[qh & c.qmask];
            qh++;
            __syncwarp();
            if (lane == 0) atomicAnd(&c.inq[l >> 5], ~(1u << (l & 31)));
            __syncwarp();
            const uint64_t live = valid & ~deadv;
            const uint32_t nl = l ^ 1;  // this literal just became false somewhere
            const uint32_t beg = __ldg(&db.occ_off[nl]), end = __ldg(&db.occ_off[nl + 1]);
            for (uint32_t i = beg + lane; i < end; i += 32) visit_clause(db.rec, __ldg(&db.occ[i]), c, live);
            __syncwarp();
        }
        __syncwarp();

        // ---- per-sample results
        const uint64_t dead = *c.dead & valid;
        const uint64_t okm = valid & ~dead;
        const bool any_impl = *c.implied != 0;
        uint32_t cnt0, cnt1;  // assigned live variables of samples lane, lane+32
        if (!any_impl) {
            cnt0 = cnt1 = pt->n_assumed_live;
        } else {
            uint64_t cb[MAX_COUNT_BITS];
#pragma unroll
            for (int q = 0; q < MAX_COUNT_BITS; q++) cb[q] = 0;
            for (int v = lane; v < db.n_live; v += 32) {
                uint64_t carry = c.S[2 * v] | c.S[2 * v + 1];
#pragma unroll
                for (int q = 0; q < MAX_COUNT_BITS; q++) {
                    uint64_t t = cb[q] & carry;
                    cb[q] ^= carry;
                    carry = t;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                uint64_t carry = 0;
#pragma unroll
                for (int q = 0; q < MAX_COUNT_BITS; q++) {
                    uint64_t o = shfl_down64(cb[q], off);
                    uint64_t x = cb[q] ^ o;
                    uint64_t s = x ^ carry;
                    carry = (cb[q] & o) | (carry & x);
                    cb[q] = s;
                }
            }
            cnt0 = cnt1 = 0;
#pragma unroll
            for (int q = 0; q < MAX_COUNT_BITS; q++) {
                uint64_t t = shfl64(cb[q], 0);
                cnt0 |= (uint32_t)((t >> lane) & 1ull) << q;
                cnt1 |= (uint32_t)((t >> (lane + 32)) & 1ull) << q;
            }
        }
        const bool ok0 = (okm >> lane) & 1ull, ok1 = (okm >> (lane + 32)) & 1ull;
        const uint32_t full_lo = __ballot_sync(0xffffffffu, ok0 && cnt0 == (uint32_t)db.n_live);
        const uint32_t full_hi = __ballot_sync(0xffffffffu, ok1 && cnt1 == (uint32_t)db.n_live);
        const uint64_t fullm = ((uint64_t)full_hi << 32) | full_lo;
        const uint32_t t0 = db.n_base_assigned + cnt0, t1 = db.n_base_assigned + cnt1;
        uint32_t lsum = (ok0 ? t0 : 0) + (ok1 ? t1 : 0);
        uint32_t lsq = (ok0 ? t0 * t0 : 0) + (ok1 ? t1 * t1 : 0);
        uint32_t limp = (ok0 && cnt0 > pt->n_assumed_live ? 1 : 0) + (ok1 && cnt1 > pt->n_assumed_live ? 1 : 0);
        lsum = __reduce_add_sync(0xffffffffu, lsum);
        lsq = __reduce_add_sync(0xffffffffu, lsq);
        limp = __reduce_add_sync(0xffffffffu, limp);
        acc.n += __popcll(valid);
        acc.n_conflict += __popcll(dead);
        acc.n_full += __popcll(fullm);
        acc.sum += lsum;
        acc.sumsq += lsq;
        acc.n_impl += limp;

        if (b.conflict_bits && lane == 0) {
            b.conflict_bits[g] = dead;
            b.full_bits[g] = fullm;
        }
        if (b.trail) {
            b.trail[g * 64 + lane] = ok0 ? (uint16_t)t0 : (uint16_t)0;
            b.trail[g * 64 + 32 + lane] = ok1 ? (uint16_t)t1 : (uint16_t)0;
        }
        if (b.assign) {
            uint64_t* dst = b.assign + g * (uint64_t)db.n_lit;
            for (int i = lane; i < db.n_lit; i += 32) dst[i] = c.S[i];
        }
        if (fullm) {
            // models found by BCP alone (mpi_predicter.cpp:755-762 b_SAT_set_array)
            uint64_t rem = fullm;
            while (rem) {
                const int s = __ffsll((long long)rem) - 1;
                rem &= rem - 1;
                unsigned long long idx = 0;
                if (lane == 0) idx = atomicAdd(b.model_count, 1ull);
                idx = shfl64(idx, 0);
                if (idx < (unsigned long long)b.max_models) {
                    for (int v0 = 0; v0 < db.n_live; v0 += 32) {
                        const int v = v0 + lane;
                        const uint32_t bit = v < db.n_live ? (uint32_t)((c.S[2 * v] >> s) & 1ull) : 0u;
                        const uint32_t word = __ballot_sync(0xffffffffu, bit);
                        if (lane == 0) b.model_bits[idx * b.model_words + (v0 >> 5)] = word;
                    }
                    if (lane == 0) {
                        b.model_point[idx] = (uint32_t)p;
                        b.model_sample[idx] = s0 + s;
                    }
                }
            }
        }
        __syncwarp();
    }
    flush_acc(acc, b.cost, lane);
}

}  // namespace pdsat
