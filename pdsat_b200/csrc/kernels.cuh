// kernels.cuh — sm_100a kernels of the PDSAT hot path.
//
//   fill_words_kernel   sample values -> bit-sliced assumption words (one 64-bit word per
//                       (64-sample group, set position): bit s = value in sample s)
//   mt19937_bits_kernel boost/std mt19937 + bool_rand (src_common/addit_func.cpp:193-196)
//                       as a packed bit stream, 624 draws per round per stream segment
//   bcp_kernel          bit-sliced unit propagation: one warp owns one group of 64 samples,
//                       the assignment lives in shared memory as one 64-bit plane per
//                       literal, lanes evaluate different gates / clauses of an occurrence list
//                       (semantics of Solver::propagate, src_common/minisat/core/Solver.cc:599-662,
//                       to the fixpoint / first conflict); the index blob of the database is
//                       staged into shared memory and the assumption words are streamed in by
//                       bulk async copies (cp.async.bulk + mbarrier)
//
// No tensor cores: nothing here is a contraction (integer / bit work bound by on-chip
// memory and the LSU).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pdsat {

// ------------------------------------------------------------------ device views
struct DbDev {
    const uint4* rec;         // 16-byte clause records: 8 x uint16 FALSE-plane codes (chains for > 8 literals)
    const uint4* occrec;      // per literal: the RECORDS of the clauses containing it, inline
                              // (a clause of > 8 literals appears as an indirect record -> rec[])
    // index blob (host_db.hpp): [occ_off u32[2*n_live+1]] [gocc_off u32[n_live+1]] [gocc u16[]]
    // [gate records uint4[]], sections 16-byte aligned; staged into shared memory when it fits
    const unsigned char* blob;
    uint32_t blob_bytes;
    uint32_t off_gocc_off, off_gocc, off_grec, off_locc_off, off_locc, off_lrec;
    int32_t blob_in_smem;
    int32_t n_live;
    int32_t n_lit;            // 2*n_live
    int32_t n_base_assigned;
    int32_t unsat;            // level-0 conflict: every sample is a conflict
    int32_t n_gates;
    int32_t n_rec;            // short-clause occurrence entries (0: no clause phase)
    int32_t n_long;           // long clauses handled lazily (0: none)
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
    uint32_t cand_off;        // first-wave candidate records (gates first, then clauses)
    uint32_t cand_cnt;
    uint32_t kp;              // words per group in `words`: k rounded up to even (16-byte tiles)
    uint32_t flags;           // POINT_*
    uint64_t block_off;       // ENUM_BLOCKS: first prefix of this point in block_prefix
    uint32_t block_bits;
    uint32_t pad_;
    uint64_t run_start;       // first global RUN of this point
    uint32_t tile;            // groups per run: 1, or BatchDev::tile_groups for a streaming point
    uint32_t pad2_;
};
static constexpr uint32_t POINT_HAS_FIXED = 1u;  // some set variable is assigned at level 0
// BCP cannot start in this point (no candidate record, no level-0-fixed set variable, not every live
// variable assumed, no assignment output; or the database itself is refuted): the host decides,
// gives the point runs of several groups and the kernel streams through them
static constexpr uint32_t POINT_STREAMING = 2u;

static constexpr uint32_t SETCODE_FIXED = 0x80000000u;  // set var already assigned at level 0; bit0 = its lbool

// Multi-GPU cost reduce fused into the BCP kernel (an all-gather + local sum over NVLink peer
// memory).  Every rank owns a receive buffer that its peers map through CUDA IPC:
//     recv[parity][source rank][2 * n_words] packets of 8 bytes = {uint32 data, uint32 tag}
// (each 64-bit accumulator travels as two packets).  The last CTA of a rank's kernel stores the
// rank's per-point sums, tagged with the step number, into its slot of EVERY rank's buffer with
// plain 8-byte stores - a naturally aligned 8-byte store is delivered whole, so a packet whose
// tag matches carries valid data: no fence, no flag, no atomics, and stale packets of older
// steps never match, so nothing is ever cleared (the scheme of NCCL's LL protocol).  The SAME CTA
// then polls its own buffer until all ranks' packets of this step are in, adds them in rank
// order (integers: any order gives the same sum) and publishes the totals - there is no second
// kernel.  Buffers alternate by step parity, so a rank that runs ahead never overwrites packets
// a slower rank is still reading (it cannot get two steps ahead: its own gather needs the
// slower rank's packets of the step in between).
static constexpr int MAX_PEERS = 16;
struct PeerDev {
    unsigned long long* shared[MAX_PEERS];  // rank r's receive buffer as mapped in this process
    int32_t n_ranks;
    int32_t me;
    int32_t parity;
    int32_t n_words;                        // accumulators = n_points * 8
    uint32_t tag;                           // step number + 1 (never 0)
    unsigned int* timed_out;                // set when a peer's packets did not arrive within PEER_TIMEOUT_NS
    // Deferred exchange.  Launch t only ADDS into its own scratch (acc[t & 1]); nothing is sent
    // and nobody waits at the end of a launch.  CTA 0 of launch t+1 - while the other CTAs already
    // work - first gathers the packets of step t-1 (sent by the peers one launch ago), then pushes
    // the sums of step t and clears their scratch.  When the host asks for the costs,
    // peer_finish_kernel does the same two things and gathers step t as well (the ranks ask at the
    // same step).  Order matters: gather before push.  A rank's push of step t overwrites the
    // parity that carried step t-2; it comes after the rank has seen every peer's packets of t-1,
    // which every peer sends after ITS gather of t-2 - so two parities suffice.
    int32_t gather_prev;                    // 1: gather (g_parity, g_tag) first
    int32_t g_parity;
    uint32_t g_tag;
    int32_t push_prev;                      // 1: then push acc_prev as (parity, tag) and clear it
    unsigned long long* acc_prev;
};
// a rank that never launches its step (crashed, or the caller broke the "same calls on every
// rank" rule) must not leave the others spinning for ever: the wait gives up and the host
// reports an error from pdsat_batch_costs
static constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct BatchDev {
    const PointDev* points;
    int32_t n_points;
    uint64_t n_groups;
    uint64_t n_words;
    const uint32_t* setcode;
    const uint4* cand;          // first-wave candidate clause records, all points
    uint64_t* words;
    const uint32_t* stream;
    const uint64_t* block_prefix;  // ENUM_BLOCKS prefixes, all points
    unsigned long long* cost;   // n_points * 8: the published accumulators (output)
    unsigned long long* acc;    // n_points * 8: scratch the warps add into; zero between launches (the
                                // last CTA publishes it into `cost` and clears it: no memset per step)
    unsigned int* cta_done;     // ticket counter (last-CTA detection)
    unsigned long long* zero_next;  // direct path (no peers, library-owned buffers): acc IS the output
                                    // of this launch, CTA 0 clears zero_next (= the next launch's) and
                                    // nobody publishes; nullptr: ticket path
    int32_t no_ticket;          // direct path or peers: the launch ends with the CTAs' adds (no last-CTA pass)
    PointDev p0;                // point 0 again, in the kernel parameters: the first tile of a one-point
                                // batch is requested without a global load in front of it
    uint64_t* conflict_bits;    // per global group (optional)
    uint64_t* full_bits;        // per global group (optional)
    uint32_t* trail;            // per global group * 64 (optional)
    uint64_t* assign;           // per global group * n_lit (optional)
    unsigned long long* model_count;
    uint32_t* model_bits;       // max_models * model_words
    uint32_t* model_point;
    uint64_t* model_sample;
    int32_t max_models;
    int32_t model_words;
    // per-warp ring of assumption-word tiles (bulk async copies): stages x stage_bytes; a tile
    // holds the words of one run = PointDev::tile consecutive groups of a point (one bulk copy)
    int32_t ring_stages;
    int32_t ring_stage_bytes;
    int32_t tile_groups;
    uint64_t n_runs;
    // fused cost all-reduce over NVLink peer memory (0 ranks = off, see PeerDev)
    PeerDev peer;
};

__device__ __forceinline__ int find_point_run(const PointDev* pts, int n, uint64_t r) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pts[mid].run_start <= r) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

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
// value of set variable j in samples 64*gi .. 64*gi+63.  A group owns kp = k rounded up to even
// words, so that every group tile is 16-byte aligned for the bulk copies of bcp_kernel.
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
    uint64_t gi = rel / pt.kp;
    uint32_t j = (uint32_t)(rel - gi * pt.kp);
    uint64_t s0 = gi * 64;
    uint64_t out = 0;
    if (j >= pt.k) {  // padding word of an odd k
    } else if (pt.mode == 1) {  // PDSAT_MODE_RANDOM_COUNTER
        uint64_t g = pt.first_sample / 64 + gi;
        out = splitmix64(pt.seed + 0x9E3779B97F4A7C15ull * (g * 65536ull + j + 1));
    } else if (pt.mode == 2 || pt.mode == 4) {  // PDSAT_MODE_ENUM / PDSAT_MODE_ENUM_BLOCKS
        uint64_t lint = pt.first_sample + s0;
        if (pt.mode == 4)  // blocks of >= 64 samples: the 64 samples of a group share their prefix
            lint = (b.block_prefix[pt.block_off + (s0 >> pt.block_bits)] << pt.block_bits) | (s0 & ((1ull << pt.block_bits) - 1));
        if (j < 64) {
#pragma unroll 8
            for (int s = 0; s < 64; s++) out |= (((lint + s) >> j) & 1ull) << s;
        }
    } else {  // packed sample-major value stream: fill_stream_kernel writes these words
        return;
    }
    b.words[w] = out;
}

// 32 x 32 bit transpose across a warp: lane r holds row r; afterwards bit r of lane c is the
// old bit c of lane r (recursive swap of the off-diagonal blocks, 5 shuffles).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
    }
    return x;
}

// Assumption words of the points whose values come as a packed sample-major bit stream
// (mt19937 bits, or the caller's explicit values): draw (s, j) = bit s*k + j of the point's
// window.  One warp per sample group: lane s pulls the k-bit row of samples s and s+32 out of
// the stream 32 columns at a time, two warp transposes turn the 64 x 32 block into the 32
// output words of those columns (bit s of word j = value of set variable j in sample s).
static constexpr int FILL_STREAM_WARPS = 8;
__global__ void __launch_bounds__(FILL_STREAM_WARPS * 32) fill_stream_kernel(BatchDev b) {
    const uint64_t g = (uint64_t)blockIdx.x * FILL_STREAM_WARPS + (threadIdx.x >> 5);
    if (g >= b.n_groups) return;
    const int lane = threadIdx.x & 31;
    const PointDev& pt = b.points[b.n_points == 1 ? 0 : find_point(b.points, b.n_points, g)];
    if (pt.mode != 0 && pt.mode != 3) return;  // PDSAT_MODE_RANDOM_MT19937 / PDSAT_MODE_EXPLICIT
    const uint64_t gi = g - pt.group_start;
    const uint64_t s0 = gi * 64;
    const uint64_t left = pt.n_samples - s0;
    const uint32_t ns = left < 64 ? (uint32_t)left : 64u;
    const uint32_t k = pt.k, kp = pt.kp;
    const uint64_t d0 = pt.stream_skip + s0 * k;                   // first bit of the group in the stream
    const uint32_t* st = b.stream + pt.stream_off + (d0 >> 5);      // its word; inside the group 32-bit offsets do
    const uint32_t sh0 = (uint32_t)(d0 & 31);
    uint64_t* dst = b.words + pt.word_off + gi * kp;
    for (uint32_t c = 0; c < kp; c += 32) {
        const uint32_t nb = c < k ? (k - c < 32 ? k - c : 32u) : 0u;  // columns of this block
        const uint32_t keep = nb < 32 ? (1u << nb) - 1u : 0xffffffffu;
        uint32_t half[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t s = lane + 32 * h;
            uint32_t row = 0;
            if (s < ns && nb) {
                const uint32_t d = sh0 + s * k + c;  // < 32 + 64 * k + k: fits 32 bits for every k the ABI allows
                const uint32_t sh = d & 31;
                const uint32_t lo = __ldg(st + (d >> 5));
                const uint32_t hi = sh + nb > 32 ? __ldg(st + (d >> 5) + 1) : 0u;
                row = __funnelshift_r(lo, hi, sh) & keep;
            }
            half[h] = warp_transpose32(row, lane);
        }
        if (c + lane < kp) dst[c + lane] = (uint64_t)half[0] | ((uint64_t)half[1] << 32);
    }
}

// ------------------------------------------------------------------ mt19937
struct MtSegment {
    uint64_t out_off;   // 32-bit word offset of this segment's first output bit (round aligned)
    uint64_t n_rounds;  // rounds of 624 draws to produce (always even)
    uint32_t state_idx; // slot of the generator state that precedes the first round
    uint32_t pad;
};

// One jump-ahead: state[dst] = g(A) state[src], g = polys[poly] (x^J mod phi as 19937 bits):
// out[w] = XOR over the set bits i of g of x[i + w], where x[0..623] is the source window and
// x continues by the mt19937 recurrence (mt_jump.hpp).  Two kernels per level of the jump tree:
//   mt19937_seq_kernel   one CTA per SOURCE state runs the recurrence and writes x[0 .. MT_SEQ_WORDS)
//   mt19937_corr_kernel  MT_PARTS CTAs per jump; CTA q correlates the polynomial bits
//                        [q*MT_CORR_CTA_BITS, (q+1)*MT_CORR_CTA_BITS) against its stretch of x
// Several jumps of one level start from the same source (the tree has radix 8): they share the
// sequence.  A state slot is stored as MT_PARTS partial vectors whose XOR is the state (CTA q
// writes partial vector q of dst: no atomics, no zeroing; consumers XOR the parts on load).
struct MtJumpTask {
    uint32_t seq;   // which sequence of the level's mt19937_seq_kernel launch
    uint32_t dst;   // state slot written
    uint32_t poly;  // slot of the jump polynomial
    uint32_t pad;
};

static constexpr int MT_PARTS = 10;           // CTAs per jump = partial vectors per state slot
static constexpr int MT_CORR_V = 20;          // output words per lane: one WARP covers 32*20 >= 624 outputs
static constexpr int MT_CORR_WARPS = 10;      // warps per CTA, each with its own polynomial bits
static constexpr int MT_CORR_BLOCKS = 10;     // 20-bit blocks per warp (a block = 10 steps of 2 bits)
static constexpr int MT_CORR_WARP_BITS = MT_CORR_BLOCKS * 20;               // 200
static constexpr int MT_CORR_CTA_BITS = MT_CORR_WARPS * MT_CORR_WARP_BITS;  // 2000; x MT_PARTS = 20000 >= 19937
static constexpr int MT_CORR_X_WORDS = MT_CORR_CTA_BITS + 32 * MT_CORR_V + 4;  // stretch of x one CTA reads
static constexpr int MT_SEQ_ROUNDS = 34;
static constexpr int MT_SEQ_WORDS = MT_SEQ_ROUNDS * 624;  // 21216 >= 20000 + 644
static constexpr int MT_SEQ_THREADS = 256;
static constexpr int MT_SEQ_SMEM = MT_SEQ_WORDS * 4;
static_assert(MT_PARTS * MT_CORR_CTA_BITS >= 19937, "polynomial bits not covered");
static_assert((MT_PARTS - 1) * MT_CORR_CTA_BITS + MT_CORR_X_WORDS <= MT_SEQ_WORDS, "sequence too short");

__device__ __forceinline__ size_t mt_state_off(uint32_t slot, int part) {
    return ((size_t)slot * MT_PARTS + part) * 624;
}

// x[n + 624] = x[n + 397] ^ twist(x[n], x[n + 1]), 227 words per phase (x[n + 397] of the
// last word of a phase is the newest word of the phase before).  The sequence stays in shared
// memory until the end.
__global__ void __launch_bounds__(MT_SEQ_THREADS) mt19937_seq_kernel(const uint32_t* __restrict__ states,
                                                                     const uint32_t* __restrict__ srcs,
                                                                     uint32_t* __restrict__ seq) {
    extern __shared__ __align__(16) uint32_t sx[];
    const int tid = threadIdx.x;
    const uint32_t src = srcs[blockIdx.x];
    for (int i = tid; i < 624; i += MT_SEQ_THREADS) {
        uint32_t v = 0;
#pragma unroll
        for (int p = 0; p < MT_PARTS; p++) v ^= states[mt_state_off(src, p) + i];
        sx[i] = v;
    }
    __syncthreads();
    for (int n0 = 0; n0 < MT_SEQ_WORDS - 624; n0 += 227) {
        const int n = n0 + tid;
        if (tid < 227 && n < MT_SEQ_WORDS - 624) {
            const uint32_t a = sx[n], b = sx[n + 1], m = sx[n + 397];
            const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
            sx[n + 624] = m ^ (y >> 1) ^ ((0u - (b & 1u)) & 0x9908b0dfu);
        }
        __syncthreads();
    }
    uint4* out = reinterpret_cast<uint4*>(seq + (size_t)blockIdx.x * MT_SEQ_WORDS);
    const uint4* in = reinterpret_cast<const uint4*>(sx);
    for (int i = tid; i < MT_SEQ_WORDS / 4; i += MT_SEQ_THREADS) out[i] = in[i];
}

// acc[v] ^= XOR over the set bits j of the 2-bit digit P of e[v + j], where e[0..19] is the
// lane's register window (element i in r[(2*S + i) % 20], S = step inside the 20-bit block)
// and e[20] is the first of the two words entering it.  Two bits per step keep the unrolled
// body (10 steps x 3 cases) small enough for the instruction cache: with 4-bit digits the
// warps spent 8 of 10 cycles waiting for instructions.
template <int P, int S>
__device__ __forceinline__ void mt_corr_apply(uint32_t (&acc)[MT_CORR_V], const uint32_t (&r)[MT_CORR_V], uint32_t in0) {
#pragma unroll
    for (int v = 0; v < MT_CORR_V; v++) {
        uint32_t a = acc[v];
        if (P & 1) a ^= r[(2 * S + v) % 20];
        if (P & 2) a ^= v + 1 < 20 ? r[(2 * S + v + 1) % 20] : in0;
        acc[v] = a;
    }
}

template <int S>
__device__ __forceinline__ void mt_corr_step(uint32_t digit, uint32_t (&acc)[MT_CORR_V], uint32_t (&r)[MT_CORR_V],
                                             uint32_t in0, uint32_t in1) {
    // the digit is warp-uniform: one uniform branch, then one XOR per output word
    if (digit == 1) mt_corr_apply<1, S>(acc, r, in0);
    else if (digit == 2) mt_corr_apply<2, S>(acc, r, in0);
    else if (digit == 3) mt_corr_apply<3, S>(acc, r, in0);
    r[(2 * S) % 20] = in0;
    r[(2 * S + 1) % 20] = in1;
}

__global__ void __launch_bounds__(MT_CORR_WARPS * 32, 2) mt19937_corr_kernel(const uint32_t* __restrict__ polys,
                                                                             uint32_t* __restrict__ states,
                                                                             const MtJumpTask* __restrict__ tasks,
                                                                             const uint32_t* __restrict__ seq) {
    __shared__ __align__(16) uint32_t x[MT_CORR_X_WORDS];
    __shared__ uint32_t g[MT_CORR_CTA_BITS / 32 + 2];
    __shared__ uint32_t red[MT_CORR_WARPS][MT_CORR_V * 33];  // [warp][v * 33 + lane]: conflict-free both ways
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int task = blockIdx.x / MT_PARTS, q = blockIdx.x % MT_PARTS;
    const MtJumpTask t = tasks[task];
    const int bit0 = q * MT_CORR_CTA_BITS;  // first polynomial bit of this CTA (a multiple of 4)
    {
        const uint4* src = reinterpret_cast<const uint4*>(seq + (size_t)t.seq * MT_SEQ_WORDS + bit0);
        uint4* dst = reinterpret_cast<uint4*>(x);
        for (int i = tid; i < MT_CORR_X_WORDS / 4; i += MT_CORR_WARPS * 32) dst[i] = __ldg(src + i);
    }
    const int gw0 = bit0 >> 5;
    if (tid < MT_CORR_CTA_BITS / 32 + 2) g[tid] = gw0 + tid < 624 ? polys[(size_t)t.poly * 624 + gw0 + tid] : 0u;
    __syncthreads();

    uint32_t acc[MT_CORR_V], r[MT_CORR_V];
    const uint32_t* xs = x + warp * MT_CORR_WARP_BITS + lane * MT_CORR_V;  // the lane's window at the warp's first bit
#pragma unroll
    for (int v = 0; v < MT_CORR_V; v += 4) {
        const uint4 w = *reinterpret_cast<const uint4*>(xs + v);
        r[v] = w.x, r[v + 1] = w.y, r[v + 2] = w.z, r[v + 3] = w.w;
        acc[v] = acc[v + 1] = acc[v + 2] = acc[v + 3] = 0;
    }
    for (int blk = 0; blk < MT_CORR_BLOCKS; blk++) {
        const int rel = (bit0 & 31) + warp * MT_CORR_WARP_BITS + blk * 20;  // bit offset inside g[]
        const uint32_t dw = __funnelshift_r(g[rel >> 5], g[(rel >> 5) + 1], rel & 31);
        const uint4* xin = reinterpret_cast<const uint4*>(xs + blk * 20 + MT_CORR_V);
        uint4 t4 = xin[0];
        mt_corr_step<0>(dw & 3u, acc, r, t4.x, t4.y);
        mt_corr_step<1>((dw >> 2) & 3u, acc, r, t4.z, t4.w);
        t4 = xin[1];
        mt_corr_step<2>((dw >> 4) & 3u, acc, r, t4.x, t4.y);
        mt_corr_step<3>((dw >> 6) & 3u, acc, r, t4.z, t4.w);
        t4 = xin[2];
        mt_corr_step<4>((dw >> 8) & 3u, acc, r, t4.x, t4.y);
        mt_corr_step<5>((dw >> 10) & 3u, acc, r, t4.z, t4.w);
        t4 = xin[3];
        mt_corr_step<6>((dw >> 12) & 3u, acc, r, t4.x, t4.y);
        mt_corr_step<7>((dw >> 14) & 3u, acc, r, t4.z, t4.w);
        t4 = xin[4];
        mt_corr_step<8>((dw >> 16) & 3u, acc, r, t4.x, t4.y);
        mt_corr_step<9>((dw >> 18) & 3u, acc, r, t4.z, t4.w);
    }
#pragma unroll
    for (int v = 0; v < MT_CORR_V; v++) red[warp][v * 33 + lane] = acc[v];
    __syncthreads();
    for (int w = tid; w < 624; w += MT_CORR_WARPS * 32) {
        const int idx = (w % MT_CORR_V) * 33 + w / MT_CORR_V;  // output w = word v of lane w / 20
        uint32_t o = 0;
#pragma unroll
        for (int ww = 0; ww < MT_CORR_WARPS; ww++) o ^= red[ww][idx];
        states[mt_state_off(t.dst, q) + w] = o;
    }
}

// one level of jumps: n_src sequences, n jumps reading them (MtJumpTask::seq < n_src)
static inline void mt19937_launch_jumps(const uint32_t* polys, uint32_t* states, const uint32_t* srcs, size_t n_src,
                                        const MtJumpTask* tasks, size_t n, uint32_t* seq, cudaStream_t st) {
    mt19937_seq_kernel<<<(unsigned)n_src, MT_SEQ_THREADS, MT_SEQ_SMEM, st>>>(states, srcs, seq);
    mt19937_corr_kernel<<<(unsigned)(n * MT_PARTS), MT_CORR_WARPS * 32, 0, st>>>(polys, states, tasks, seq);
}

// ---- warp-register mt19937: one WARP per stream segment, the 624-word state lives in 20
// registers per lane (R[c] of lane l = mt[32c + l]); a round is 20 chunks of 32 words updated
// in place, neighbours fetched with warp shuffles - no shared memory, no block barriers:
//     mt[k] <- mt[(k+397) % 624] ^ twist(mt[k], mt[(k+1) % 624])
// Per draw only the bool_rand bit is needed: bool_rand == !(tempered >> 31), and bit 31 of the
// tempered word is parity(y & 0x89010000) of the raw word y (bits 31, 27, 24, 16 - follows
// from the four tempering shifts/masks).  Bits are packed little-endian, 39 words per 2 rounds.
// The loop is bound by the ALU pipe (one warp instruction per two cycles per scheduler), so
// the update keeps that pipe's share small: one bit-select LOP3 for the twist input, the
// matrix term and the parity as integer multiplies (FMA pipe), the ballot word stored by lane
// 0 as it is produced instead of being routed to a lane with a compare + select per chunk.
struct MtWarp {
    uint32_t R[20];
    uint32_t carry;  // pending low half-word of an odd round
    uint32_t* dst;   // output words of the current round pair
    bool store;      // lane 0 of a live segment
    int lane;

    template <int IDX>
    __device__ __forceinline__ void emit(uint32_t word) {
        if (store) dst[IDX] = word;
    }
    // neighbours of chunk C: b = mt[k+1], m = mt[(k+397) % 624] for k = 32C + lane
    template <int C>
    __device__ __forceinline__ void gather(uint32_t& b, uint32_t& m) const {
        if (C < 19) {
            const uint32_t v = lane >= 1 ? R[C] : R[(C + 1) % 20];
            b = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
        } else {  // k = 608..623: the last word pairs with the NEW mt[0]
            const uint32_t v = lane == 0 ? R[0] : R[19];
            b = __shfl_sync(0xffffffffu, v, (lane + 1) & 15);
        }
        if (C < 7) {  // k + 397 < 624: old words of chunks C+12, C+13
            const uint32_t v = lane >= 13 ? R[(C + 12) % 20] : R[(C + 13) % 20];
            m = __shfl_sync(0xffffffffu, v, (lane + 13) & 31);
        } else if (C == 7) {  // straddles the wrap: old mt[621..623], then new mt[0..28]
            const uint32_t m_old = __shfl_sync(0xffffffffu, R[19], (lane + 13) & 31);
            const uint32_t m_new = __shfl_sync(0xffffffffu, R[0], (lane + 29) & 31);
            m = lane < 3 ? m_old : m_new;
        } else {  // k - 227: new words of chunks C-8, C-7
            const uint32_t v = lane >= 29 ? R[(C + 12) % 20] : R[(C + 13) % 20];
            m = __shfl_sync(0xffffffffu, v, (lane + 29) & 31);
        }
    }
    template <int C, int ODD>
    __device__ __forceinline__ void update(uint32_t b, uint32_t m) {
        uint32_t y;  // (mt[k] & 0x80000000) | (mt[k+1] & 0x7fffffff) as one bit-select
        asm("lop3.b32 %0, %1, %2, 0x80000000, 0xE4;" : "=r"(y) : "r"(R[C]), "r"(b));
        uint32_t mag;  // (mt[k+1] & 1) ? 0x9908b0df : 0 as a multiply (inline PTX: the compiler would turn it into compare + select)
        asm("mul.lo.u32 %0, %1, 0x9908b0df;" : "=r"(mag) : "r"(b & 1u));
        const uint32_t nv = m ^ (y >> 1) ^ mag;
        R[C] = nv;
        // parity of bits 31, 27, 24, 16 = bit 31 of the product with 2^0 + 2^4 + 2^7 + 2^15: the
        // four bits meet at position 31, the other partial products fall on distinct lower bits
        int32_t par;
        asm("mul.lo.s32 %0, %1, 0x8091;" : "=r"(par) : "r"(nv & 0x89010000u));
        uint32_t w = __ballot_sync(0xffffffffu, par >= 0);
        if (C == 19) w &= 0xffffu;
        if (!ODD) {
            if (C < 19) emit<C>(w);
            else carry = w;
        } else {
            emit<19 + C>(carry | (w << 16));
            carry = w >> 16;
        }
    }
    // A chunk's neighbours are 1 and 397 words ahead (= 227 behind, new): chunk c needs the NEW
    // words of chunks c-8 and c-7 and the OLD words of chunk c+1.  So the round is a software
    // pipeline of distance 7: right after chunk j is updated the shuffles of chunk j+7 are issued,
    // and their results are consumed seven updates later - the shuffle latency is hidden by
    // independent work even with one warp per scheduler.
    template <int ODD>
    __device__ __forceinline__ void round() {
        uint32_t b[7], m[7];
        gather<0>(b[0], m[0]); gather<1>(b[1], m[1]); gather<2>(b[2], m[2]); gather<3>(b[3], m[3]);
        gather<4>(b[4], m[4]); gather<5>(b[5], m[5]); gather<6>(b[6], m[6]);
#define PDSAT_MT_STEP(J)                     \
    update<J, ODD>(b[(J) % 7], m[(J) % 7]); \
    if ((J) + 7 < 20) gather<((J) + 7 < 20 ? (J) + 7 : 0)>(b[(J) % 7], m[(J) % 7]);
        PDSAT_MT_STEP(0) PDSAT_MT_STEP(1) PDSAT_MT_STEP(2) PDSAT_MT_STEP(3) PDSAT_MT_STEP(4)
        PDSAT_MT_STEP(5) PDSAT_MT_STEP(6) PDSAT_MT_STEP(7) PDSAT_MT_STEP(8) PDSAT_MT_STEP(9)
        PDSAT_MT_STEP(10) PDSAT_MT_STEP(11) PDSAT_MT_STEP(12) PDSAT_MT_STEP(13) PDSAT_MT_STEP(14)
        PDSAT_MT_STEP(15) PDSAT_MT_STEP(16) PDSAT_MT_STEP(17) PDSAT_MT_STEP(18) PDSAT_MT_STEP(19)
#undef PDSAT_MT_STEP
    }
};

static constexpr int MT_BITS_WARPS = 4;  // warps (= segments) per CTA

// n_rounds (even) is the same for every segment and passed by value so that the loop trip
// count is provably uniform (no divergence around the warp collectives); surplus warps of the
// last CTA redo the last segment and skip the stores.
__global__ void __launch_bounds__(MT_BITS_WARPS * 32) mt19937_bits_kernel(const uint32_t* __restrict__ states,
                                                                          const MtSegment* __restrict__ segs,
                                                                          int n_segs, uint32_t n_rounds,
                                                                          uint32_t* __restrict__ out) {
    const int seg_raw = blockIdx.x * MT_BITS_WARPS + (threadIdx.x >> 5);
    const bool live = seg_raw < n_segs;
    const MtSegment seg = segs[live ? seg_raw : n_segs - 1];
    MtWarp g;
    g.lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 20; c++) {
        uint32_t v = 0;
        if (32 * c + g.lane < 624) {
#pragma unroll
            for (int p = 0; p < MT_PARTS; p++) v ^= states[mt_state_off(seg.state_idx, p) + 32 * c + g.lane];
        }
        g.R[c] = v;
    }
    g.carry = 0;
    g.store = live && g.lane == 0;
    g.dst = out + seg.out_off;
    for (uint32_t r = 0; r < n_rounds; r += 2) {
        g.round<0>();
        g.round<1>();
        g.dst += 39;
    }
}

// ------------------------------------------------------------------ bulk async copies (TMA) + mbarrier
// 1-D bulk copies global -> shared issued by ONE thread and tracked by an mbarrier
// (cp.async.bulk, SASS UBLKCP): used to stage the index blob of the clause database into shared
// memory and to stream the assumption words of the next groups behind the propagation.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same on 32-bit shared-window addresses (hot loops: no generic -> shared conversion)
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

// ------------------------------------------------------------------ BCP
// Per-warp working set in shared memory (one warp owns one group of 64 samples):
//   S[l]      one plane per literal: samples in which literal l is TRUE.  S[2v], S[2v+1] (the
//             TRUE and FALSE plane of variable v) are adjacent: one vector load fetches both
//   fill      bitset: literals that gained samples since they were last visited - the frontier
//             under construction (making a literal true = atomicOr on its plane + one atomicOr here)
//   drain     bitset: the frontier being visited (the two swap roles at every level: the
//             propagation is level-synchronous, breadth first)
//   tset      bitset: literals implied in this group (undo list, per-sample trail counts)
//   ring      multi-stage ring of assumption-word tiles filled by bulk async copies
// The plane type PT is a template parameter: uint64_t = the 64 samples of a group in one pass;
// uint32_t = two passes of 32 samples each over half-size planes.  Half-size planes halve the
// shared memory per warp (twice the warps per SM for large instances: ODLS 5 -> 11, A5/2 1 -> 3)
// and make every bit operation a native 32-bit one; the price is that a literal implied in both
// halves is visited twice.  The host picks per database (pdsat_cuda.cu launch_shape).
template <typename PT>
struct WarpCtx {
    PT* S;
    uint32_t* fill;
    uint32_t* drain;
    uint32_t* tset;
    PT* dead;  // samples in conflict
};

template <typename PT>
struct Plane;
template <>
struct Plane<uint64_t> {
    typedef ulonglong2 Pair;
    static constexpr int W = 64;
    static __device__ __forceinline__ int popc(uint64_t v) { return __popcll(v); }
    static __device__ __forceinline__ int ffs(uint64_t v) { return __ffsll((long long)v); }
    // OR into shared memory as two native 32-bit ATOMS.OR (a 64-bit shared atomicOr compiles to
    // a compare-and-swap spin loop); returns the previous value of the halves touched
    static __device__ __forceinline__ uint64_t or_shared(uint64_t* p, uint64_t v) {
        uint32_t* q = reinterpret_cast<uint32_t*>(p);
        const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        uint32_t olo = 0, ohi = 0;
        if (lo) olo = atomicOr(q, lo);
        if (hi) ohi = atomicOr(q + 1, hi);
        return ((uint64_t)ohi << 32) | olo;
    }
};
template <>
struct Plane<uint32_t> {
    typedef uint2 Pair;
    static constexpr int W = 32;
    static __device__ __forceinline__ int popc(uint32_t v) { return __popc(v); }
    static __device__ __forceinline__ int ffs(uint32_t v) { return __ffs((int)v); }
    static __device__ __forceinline__ uint32_t or_shared(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
};

// make literal l true in the samples `bits` (Solver.cc:579-585 uncheckedEnqueue, for up to 64
// samples at once).  Planes only gain bits; a literal that gained a bit joins the next frontier,
// so every clause / gate that can react to it is visited (again).
// A sample in which the complement is (or becomes) true as well is a CONFLICT: two reasons
// implied opposite values, the later reason's clause is falsified (Solver.cc:641-646).  The check
// is made here, after the own write: of two racing writers the later one sees the earlier one's
// plane (shared-memory operations of a warp complete in order), so no such sample is missed.
// Clauses would notice the contradiction by themselves (all their literals are false); two gates
// that read the variable through opposite polarities would each see a value that suits them.
template <typename PT>
__device__ __forceinline__ void set_true(const WarpCtx<PT>& c, uint32_t l, PT bits) {
    const PT old = Plane<PT>::or_shared(&c.S[l], bits);
    const PT contra = bits & *(const volatile PT*)&c.S[l ^ 1u];
    if (contra) Plane<PT>::or_shared(c.dead, contra);
    if (bits & ~old) atomicOr(&c.fill[l >> 5], 1u << (l & 31));
}

// variable with TRUE-plane code tcode (even): true in the samples t, false in the samples f
__device__ __forceinline__ void set_pair(const WarpCtx<uint64_t>& c, uint32_t tcode, uint64_t t, uint64_t f) {
    uint32_t* q = reinterpret_cast<uint32_t*>(c.S + tcode);  // T.lo T.hi F.lo F.hi
    const uint32_t tl = (uint32_t)t, th = (uint32_t)(t >> 32), fl = (uint32_t)f, fh = (uint32_t)(f >> 32);
    uint32_t nt = 0, nf = 0;
    if (tl) nt |= tl & ~atomicOr(q, tl);
    if (th) nt |= th & ~atomicOr(q + 1, th);
    if (fl) nf |= fl & ~atomicOr(q + 2, fl);
    if (fh) nf |= fh & ~atomicOr(q + 3, fh);
    const volatile uint64_t* now = c.S + tcode;
    const uint64_t contra = (t & now[1]) | (f & now[0]);
    if (contra) Plane<uint64_t>::or_shared(c.dead, contra);
    const uint32_t m = (nt ? 1u : 0u) | (nf ? 2u : 0u);
    if (m) atomicOr(&c.fill[tcode >> 5], m << (tcode & 31));  // both literals sit in one word
}
__device__ __forceinline__ void set_pair(const WarpCtx<uint32_t>& c, uint32_t tcode, uint32_t t, uint32_t f) {
    uint32_t* q = c.S + tcode;  // T F
    uint32_t nt = 0, nf = 0;
    if (t) nt = t & ~atomicOr(q, t);
    if (f) nf = f & ~atomicOr(q + 1, f);
    const volatile uint32_t* now = q;
    const uint32_t contra = (t & now[1]) | (f & now[0]);
    if (contra) atomicOr(c.dead, contra);
    const uint32_t m = (nt ? 1u : 0u) | (nf ? 2u : 0u);
    if (m) atomicOr(&c.fill[tcode >> 5], m << (tcode & 31));
}

__device__ __forceinline__ uint32_t rec_slot(const uint4& r, int j) {
    const uint32_t w = j < 2 ? r.x : (j < 4 ? r.y : (j < 6 ? r.z : r.w));
    return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// (S[code & ~1], S[code | 1]) with one vector shared load
template <typename PT>
__device__ __forceinline__ void ld_pair(const PT* S, uint32_t code, PT& even, PT& odd) {
    const typename Plane<PT>::Pair v = *reinterpret_cast<const typename Plane<PT>::Pair*>(S + (code & ~1u));
    even = v.x;
    odd = v.y;
}

// Record kinds (slot 7): plain clause (a plane code), REC_CONT_D / REC_INDIRECT_D (clauses of more
// than 8 literals: INDIRECT = first word is the index of the chain in rec[], CONT links the chain),
// REC_GATE_XOR_D | p, REC_GATE_ANDXOR_D | p (gate records, see host_db.hpp).
static constexpr uint32_t REC_CONT_D = 0xFFFEu, REC_INDIRECT_D = 0xFFFDu;
static constexpr uint32_t REC_GATE_XOR_D = 0xFFF0u, REC_GATE_ANDXOR_D = 0xFFF2u, REC_TAG_MIN_D = 0xFFF0u;

// Clause slots hold the code m of the literal's FALSE plane; its TRUE plane is m ^ 1.
//   ge1 / ge2 = samples with >= 1 / >= 2 non-false literals,  sat = samples with a true literal
// Unused slots point at the pad plane pair (FALSE plane all ones, TRUE plane zero).
template <typename PT>
__device__ __forceinline__ void accumulate_slot(const PT* S, uint32_t m, PT& ge1, PT& ge2, PT& sat) {
    PT e, o;
    ld_pair(S, m, e, o);
    const PT f = (m & 1u) ? o : e, t = (m & 1u) ? e : o;
    const PT nf = ~f;
    ge2 |= ge1 & nf;
    ge1 |= nf;
    sat |= t;
}

// in the samples of `unit` the clause has exactly one non-false literal and it is undefined: make
// it true (Solver.cc:641-648 uncheckedEnqueue(first, cr))
template <typename PT>
__device__ __forceinline__ void imply_slot(const WarpCtx<PT>& c, uint32_t m, PT unit) {
    const PT imp = unit & ~c.S[m];
    if (imp) set_true(c, m ^ 1u, imp);
}

template <typename PT>
__device__ __forceinline__ void eval_clause(const uint4& r0, const uint4* __restrict__ rec, const WarpCtx<PT>& c, PT live,
                                            uint32_t pad2) {
    PT ge1 = 0, ge2 = 0, sat = 0;
    const bool indirect = rec_slot(r0, 7) == REC_INDIRECT_D;
    if (indirect) {
        uint32_t cur = r0.x;
        bool more;
        do {
            const uint4 r = __ldg(&rec[cur]);
#pragma unroll
            for (int j = 0; j < 7; j++) accumulate_slot(c.S, rec_slot(r, j), ge1, ge2, sat);
            const uint32_t m7 = rec_slot(r, 7);
            more = m7 == REC_CONT_D;
            if (!more) accumulate_slot(c.S, m7, ge1, ge2, sat);
            cur++;
        } while (more);
    } else {
        // literals fill the slots from the front; a word of two pad slots ends the clause
        accumulate_slot(c.S, r0.x & 0xFFFFu, ge1, ge2, sat);
        accumulate_slot(c.S, r0.x >> 16, ge1, ge2, sat);
        if (r0.y != pad2) {
            accumulate_slot(c.S, r0.y & 0xFFFFu, ge1, ge2, sat);
            accumulate_slot(c.S, r0.y >> 16, ge1, ge2, sat);
            if (r0.z != pad2) {
                accumulate_slot(c.S, r0.z & 0xFFFFu, ge1, ge2, sat);
                accumulate_slot(c.S, r0.z >> 16, ge1, ge2, sat);
                if (r0.w != pad2) {
                    accumulate_slot(c.S, r0.w & 0xFFFFu, ge1, ge2, sat);
                    accumulate_slot(c.S, r0.w >> 16, ge1, ge2, sat);
                }
            }
        }
    }
    const PT confl = ~ge1 & live;
    const PT unit = ge1 & ~ge2 & ~sat & live;
    if (confl) Plane<PT>::or_shared(c.dead, confl);
    if (unit) {
        if (indirect) {
            uint32_t cur = r0.x;
            bool more;
            do {
                const uint4 r = __ldg(&rec[cur]);
#pragma unroll
                for (int j = 0; j < 7; j++) imply_slot(c, rec_slot(r, j), unit);
                const uint32_t m7 = rec_slot(r, 7);
                more = m7 == REC_CONT_D;
                if (!more) imply_slot(c, m7, unit);
                cur++;
            } while (more);
        } else {
            imply_slot(c, r0.x & 0xFFFFu, unit);
            imply_slot(c, r0.x >> 16, unit);
            if (r0.y != pad2) {
                imply_slot(c, r0.y & 0xFFFFu, unit);
                imply_slot(c, r0.y >> 16, unit);
                if (r0.z != pad2) {
                    imply_slot(c, r0.z & 0xFFFFu, unit);
                    imply_slot(c, r0.z >> 16, unit);
                    if (r0.w != pad2) {
                        imply_slot(c, r0.w & 0xFFFFu, unit);
                        imply_slot(c, r0.w >> 16, unit);
                    }
                }
            }
        }
    }
}

// Gate in closed form on the samples of a plane.  XOR: x_0 ^ ... ^ x_6 = p.  ANDXOR:
// x_0 ^ ... ^ x_4 ^ q = p with q = Lb & Lc (q is known 1 where both literals are true, known 0
// where one is false).  Members = the xor variables (and q).  Conflict where no member is unknown
// and the parity is wrong; where exactly one member is unknown it takes the value
// p ^ xor(known values).  For the product: q must be 1 -> Lb and Lc become true; q must be 0 and
// one literal is true -> the other becomes false.  Equal to unit propagation over the gate's
// clauses (tests/test_gates.py).  The two kinds share one instruction stream (selects on the kind
// mask), so a warp that mixes them does not diverge until something is actually implied.
template <typename PT>
__device__ __forceinline__ void imply_var(const WarpCtx<PT>& c, uint32_t tcode, PT imp, PT par) {
    if (imp) set_pair(c, tcode, (PT)(imp & par), (PT)(imp & ~par));
}

template <typename PT>
__device__ __forceinline__ void eval_gate(const uint4& r, const WarpCtx<PT>& c, PT live) {
    const uint32_t tag = r.w >> 16;
    const PT am = (tag & 2u) ? ~(PT)0 : (PT)0;  // ANDXOR
    PT par = (tag & 1u) ? ~(PT)0 : (PT)0;       // p ^ xor(values): must end up 0
    PT any = 0, two = 0;
    PT u0, u1, u2, u3, u4;
    PT t, f;
#define PDSAT_XVAR(U, CODE)              \
    ld_pair(c.S, (CODE), t, f);          \
    U = ~(t | f);                        \
    par ^= t;                            \
    two |= any & U;                      \
    any |= U;
    PDSAT_XVAR(u0, r.x & 0xFFFFu)
    PDSAT_XVAR(u1, r.x >> 16)
    PDSAT_XVAR(u2, r.y & 0xFFFFu)
    PDSAT_XVAR(u3, r.y >> 16)
    PDSAT_XVAR(u4, r.z & 0xFFFFu)
#undef PDSAT_XVAR
    // slots 5, 6: two more xor variables (TRUE-plane codes, even) or the literals Lb, Lc
    const uint32_t c5 = r.z >> 16, c6 = r.w & 0xFFFFu;
    PT e, o;
    ld_pair(c.S, c5, e, o);
    const PT tb = (c5 & 1u) ? o : e, fb = (c5 & 1u) ? e : o;
    ld_pair(c.S, c6, e, o);
    const PT tc = (c6 & 1u) ? o : e, fc = (c6 & 1u) ? e : o;
    const PT q1 = tb & tc;
    const PT m5 = (~(tb | fb) & ~am) | (~(q1 | fb | fc) & am);  // x5 unknown / q unknown
    const PT m6 = ~(tc | fc) & ~am;                              // x6 unknown / -
    par ^= ((tb ^ tc) & ~am) | (q1 & am);
    two |= any & m5;
    any |= m5;
    two |= any & m6;
    any |= m6;
    const PT confl = ~any & par & live;
    const PT one = any & ~two & live;
    if (confl) Plane<PT>::or_shared(c.dead, confl);
    if (one) {
        imply_var(c, r.x & 0xFFFFu, (PT)(one & u0), par);
        imply_var(c, r.x >> 16, (PT)(one & u1), par);
        imply_var(c, r.y & 0xFFFFu, (PT)(one & u2), par);
        imply_var(c, r.y >> 16, (PT)(one & u3), par);
        imply_var(c, r.z & 0xFFFFu, (PT)(one & u4), par);
        const PT i5 = one & m5, i6 = one & m6;
        if (i5 | i6) {
            if (!am) {
                imply_var(c, c5, i5, par);
                imply_var(c, c6, i6, par);
            } else {
                const PT s1 = i5 & par, s0 = i5 & ~par;
                const PT b_true = s1 & ~tb, b_false = s0 & tc, c_true = s1 & ~tc, c_false = s0 & tb;
                if (b_true | b_false) set_pair(c, c5 & ~1u, (c5 & 1u) ? b_false : b_true, (c5 & 1u) ? b_true : b_false);
                if (c_true | c_false) set_pair(c, c6 & ~1u, (c6 & 1u) ? c_false : c_true, (c6 & 1u) ? c_true : c_false);
            }
        }
    }
}

template <typename PT>
__device__ __forceinline__ void eval_record(const uint4& r, const uint4* __restrict__ rec, const WarpCtx<PT>& c, PT live,
                                            uint32_t pad2) {
    const uint32_t tag = r.w >> 16;
    if (tag >= REC_TAG_MIN_D && tag < REC_INDIRECT_D) eval_gate(r, c, live);
    else eval_clause(r, rec, c, live, pad2);
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src) << 32) |
           __shfl_sync(0xffffffffu, (uint32_t)v, src);
}
__device__ __forceinline__ uint64_t shfl_down64(uint64_t v, int d) {
    return ((uint64_t)__shfl_down_sync(0xffffffffu, (uint32_t)(v >> 32), d) << 32) |
           __shfl_down_sync(0xffffffffu, (uint32_t)v, d);
}
__device__ __forceinline__ uint64_t shfl_plane(uint64_t v, int src) { return shfl64(v, src); }
__device__ __forceinline__ uint32_t shfl_plane(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ uint64_t shfl_down_plane(uint64_t v, int d) { return shfl_down64(v, d); }
__device__ __forceinline__ uint32_t shfl_down_plane(uint32_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += shfl_down64(v, off);
    return shfl64(v, 0);
}

static constexpr int MAX_COUNT_BITS = 15;

// Per-warp accumulators of the current point: 8 x u64 in the warp's shared memory (lane 0 adds;
// sixteen registers less in the propagation loop), flushed with <= 8 atomicAdds per point change.
#ifdef PDSAT_TIMELINE  // experiments only: per-CTA timestamps of the phases of bcp_kernel
__device__ unsigned long long g_timeline[148 * 8];
__device__ __forceinline__ void tl_mark(int i) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_timeline[blockIdx.x * 8 + i] = t;
    }
}
#define TL(i) tl_mark(i)
#else
#define TL(i)
#endif

enum { ACC_N = 0, ACC_CONFLICT, ACC_FULL, ACC_SUM, ACC_SUMSQ, ACC_IMPL, ACC_POPS, ACC_TOUCHED };

__device__ __forceinline__ void flush_acc(uint64_t* a, int point, unsigned long long* cost, int lane) {
    if (lane == 0) {
        if (point >= 0 && a[ACC_N]) {
            unsigned long long* c = cost + (size_t)point * 8;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (a[i]) atomicAdd(&c[i], (unsigned long long)a[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = 0;
    }
}

// owner of flattened item i among the lanes' lists: the smallest q with cum[q] > i (cum =
// inclusive prefix sums held one per lane)
__device__ __forceinline__ int find_owner(uint32_t cum, uint32_t i) {
    int lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t cmid = __shfl_sync(0xffffffffu, cum, lo + step - 1);
        if (cmid <= i) lo += step;
    }
    return lo;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += up;
    }
    return v;
}

// per-sample count of set planes among the literals of the bitset `set`: bit-sliced vertical
// counters of NB bits (count < 2^NB) per lane over the words the lane owns, then summed over
// the lanes; returns the counts of samples lane and (64-bit planes) lane + 32
template <int NB, typename PT>
__device__ __forceinline__ void count_planes(const PT* S, const uint32_t* set, int n_words, int lane, uint32_t& cnt0,
                                             uint32_t& cnt1) {
    PT cb[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) cb[q] = 0;
    for (int w = lane; w < n_words; w += 32) {
        uint32_t bits = set[w];
        while (bits) {
            const int bp = __ffs(bits) - 1;
            bits &= bits - 1;
            PT carry = S[32 * w + bp];
#pragma unroll
            for (int q = 0; q < NB; q++) {
                const PT t = cb[q] & carry;
                cb[q] ^= carry;
                carry = t;
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        PT carry = 0;
#pragma unroll
        for (int q = 0; q < NB; q++) {
            const PT o = shfl_down_plane(cb[q], off);
            const PT x = cb[q] ^ o;
            const PT s = x ^ carry;
            carry = (cb[q] & o) | (carry & x);
            cb[q] = s;
        }
    }
    cnt0 = cnt1 = 0;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        const PT t = shfl_plane(cb[q], 0);
        cnt0 |= (uint32_t)((t >> lane) & 1u) << q;
        if (Plane<PT>::W == 64) cnt1 |= (uint32_t)(((uint64_t)t >> (lane + 32)) & 1ull) << q;
    }
}

// Persistent kernel: grid = #SMs, each warp strides over the 64-sample groups.
//
// Per group:  (0) the group's assumption words arrive through the warp's bulk-copy ring (tiles
// of the next `stages` groups are in flight behind the propagation);  (1) write the assumption
// planes;  (2) FIRST WAVE: visit the point's candidate records - the only clauses / gates that
// can be unit or falsified when just the set variables are assigned (computed once per point on
// the host; MiniSat gets the same effect lazily from its two watched literals);  (3) visit the
// frontier level by level: every gate that contains the variable and every clause that contains
// the complement of a literal that gained samples (Solver.cc:611-655), until nothing changes;
// (4) per-sample results and integer cost accumulation;  (5) undo: clear only the planes written.
// With 32-bit planes (1) - (5) run twice per group, once per half.
// A point without candidates and without level-0-fixed set variables cannot start BCP at all:
// its groups skip (1), (2), (3) and (5) - the words still stream through the ring.
// ---- cost exchange over NVLink peer memory (PeerDev).  Block-wide helpers.
// gather: wait for the packets of every rank for (parity, tag) in this rank's own buffer, add
// them and (publish) write the totals into b.cost
__device__ __forceinline__ void peer_gather(const BatchDev& b, int parity, uint32_t tag, bool publish) {
    const int nw = b.n_points * 8, nr = b.peer.n_ranks;
    const volatile unsigned long long* mine = b.peer.shared[b.peer.me] + (size_t)parity * MAX_PEERS * 2 * nw;
    unsigned long long t_start;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    for (int i = threadIdx.x; i < nw; i += blockDim.x) {
        // all sources' packets are requested before the first tag is looked at: in the steady
        // state they arrived a launch ago and the gather costs one L2 round trip, not one per rank
        unsigned long long lo_[MAX_PEERS], hi_[MAX_PEERS];
#pragma unroll
        for (int src = 0; src < MAX_PEERS; src++) {
            if (src < nr) {
                const volatile unsigned long long* q = mine + (size_t)src * 2 * nw + 2 * i;
                lo_[src] = q[0];
                hi_[src] = q[1];
            }
        }
        unsigned long long sum = 0;
#pragma unroll
        for (int src = 0; src < MAX_PEERS; src++) {
            if (src >= nr) break;
            const volatile unsigned long long* q = mine + (size_t)src * 2 * nw + 2 * i;
            unsigned long long lo = lo_[src], hi = hi_[src];
            uint32_t spins = 0;
            while ((uint32_t)(lo >> 32) != tag || (uint32_t)(hi >> 32) != tag) {
                if ((++spins & 0xfffu) == 0) {
                    unsigned long long t_now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_now));
                    if (t_now - t_start > PEER_TIMEOUT_NS) {
                        *b.peer.timed_out = 1u;
                        break;
                    }
                }
                lo = q[0];
                hi = q[1];
            }
            sum += (hi << 32) | (uint32_t)lo;
        }
        if (publish) b.cost[i] = sum;
    }
}

// push: this rank's sums of the launch before (acc_prev) as tagged packets into its slot of every
// rank's buffer, then clear them; every rank starts with a different target so that the ranks
// do not all hit the same peer at once
__device__ __forceinline__ void peer_push(const BatchDev& b) {
    const int nw = b.n_points * 8, nr = b.peer.n_ranks, me = b.peer.me;
    const uint32_t tag = b.peer.tag;
    const size_t slot = ((size_t)b.peer.parity * MAX_PEERS + me) * 2 * nw;
    for (int j = threadIdx.x; j < 2 * nw * nr; j += blockDim.x) {
        const int e = j % (2 * nw);
        const int r = (j / (2 * nw) + me) % nr;
        const unsigned long long v = __ldcg(&b.peer.acc_prev[e >> 1]);
        const unsigned long long pkt = ((unsigned long long)tag << 32) | (uint32_t)((e & 1) ? (v >> 32) : v);
        *((volatile unsigned long long*)&b.peer.shared[r][slot + e]) = pkt;
    }
    __syncthreads();  // every packet is on its way before the scratch is cleared
    for (int i = threadIdx.x; i < nw; i += blockDim.x) b.peer.acc_prev[i] = 0;
}

// what CTA 0 of a launch does first (see PeerDev)
__device__ __forceinline__ void peer_catch_up(const BatchDev& b) {
    if (b.peer.gather_prev) {
        peer_gather(b, b.peer.g_parity, b.peer.g_tag, true);
        __syncthreads();
    }
    if (b.peer.push_prev) peer_push(b);
}

// the host wants the totals of the last launch: catch up, then gather that step as well
__global__ void __launch_bounds__(512) peer_finish_kernel(BatchDev b) {
    peer_catch_up(b);
    __syncthreads();
    peer_gather(b, b.peer.parity, b.peer.tag, true);
}

template <typename PT>
__device__ __forceinline__ void bcp_body(const DbDev& db, const BatchDev& b, int warp_bytes, unsigned char* smem,
                                         uint64_t* blob_bar) {
    constexpr int W = Plane<PT>::W;
    constexpr int HALVES = 64 / W;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    TL(0);
    // the next kernel in the stream (another propagate launched the same way) may start its own
    // prologue as soon as SMs free up; it waits at its griddepcontrol.wait for this grid to finish
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- index blob (occurrence offsets, gate lists, gate records): one bulk copy per CTA into
    // shared memory when it fits, otherwise read from global / L2
    const unsigned char* blob = db.blob;
    const int blob_smem_bytes = db.blob_in_smem ? (int)db.blob_bytes : 0;
    if (db.blob_in_smem) {
        if (threadIdx.x == 0) {
            mbar_init(blob_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(blob_bar, db.blob_bytes);
            for (uint32_t off = 0; off < db.blob_bytes; off += 32768u) {
                const uint32_t n = min(32768u, db.blob_bytes - off);
                bulk_g2s(smem + off, db.blob + off, n, blob_bar);
            }
        }
        blob = smem;
    }
    const uint32_t* occ_off = (const uint32_t*)blob;
    const uint32_t* gocc_off = (const uint32_t*)(blob + db.off_gocc_off);
    const uint16_t* gocc = (const uint16_t*)(blob + db.off_gocc);
    const uint4* grec = (const uint4*)(blob + db.off_grec);
    const uint32_t* locc_off = (const uint32_t*)(blob + db.off_locc_off);
    const uint16_t* locc = (const uint16_t*)(blob + db.off_locc);
    const uint32_t* lrec = (const uint32_t*)(blob + db.off_lrec);

    unsigned char* base = smem + blob_smem_bytes + (size_t)warp * warp_bytes;
    WarpCtx<PT> c;
    c.S = (PT*)base;
    const int n_planes = db.n_lit + 4;
    const int inq_words = (db.n_lit + 31) >> 5;
    const int iw2 = (inq_words + 1) & ~1;
    c.fill = (uint32_t*)(c.S + n_planes);
    c.drain = c.fill + iw2;
    c.tset = c.drain + iw2;
    // marks of the long clauses whose literals were falsified since they were last evaluated
    uint32_t* lmark = c.tset + iw2;
    const int lm_words = (db.n_long + 31) >> 5;
    const int lm2 = (lm_words + 1) & ~1;
    c.dead = (PT*)(lmark + lm2);
    // 16-byte aligned: the point accumulators, mbarriers, then the tiles the bulk copies write
    uint64_t* accs = (uint64_t*)(base + ((((unsigned char*)(c.dead + 2) - base) + 15) & ~(size_t)15));
    uint64_t* ring_bar = accs + 8;
    const int stages = b.ring_stages;
    unsigned char* ring = (unsigned char*)(ring_bar + ((stages + 1) & ~1));

    // the planes start clean and every group cleans up after itself; pads: plane n_lit = "always
    // false" literal (FALSE plane all ones, TRUE plane 0), n_lit + 2 / + 3 = the pad variable of
    // gate records (TRUE plane 0, FALSE plane all ones: known, value 0)
    for (int i = lane; i < n_planes; i += 32) c.S[i] = (i == db.n_lit || i == db.n_lit + 3) ? ~(PT)0 : (PT)0;
    for (int i = lane; i < 3 * iw2 + lm2; i += 32) c.fill[i] = 0;
    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&ring_bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const uint32_t pad2 = (uint32_t)db.n_lit | ((uint32_t)db.n_lit << 16);  // two pad slots
    if (lane < 8) accs[lane] = 0;
    __syncwarp();

    // The unit of work of a warp is a RUN: tile_groups consecutive groups of one point, whose
    // words are contiguous in HBM and arrive with ONE bulk copy (runs never straddle points).
    const uint64_t stride = (uint64_t)gridDim.x * wpc;
    const uint64_t r_first = (uint64_t)blockIdx.x * wpc + warp;

    // ---- producer side of the words ring (lane 0): run pf_r goes to stage pf_stage
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(ring_bar);
    const uint32_t stage_bytes = (uint32_t)b.ring_stage_bytes;
    uint64_t pf_r = r_first;
    uint32_t pf_stage = 0;
    uint64_t pf_run_end = 0;
    const uint64_t* pf_ptr = nullptr;  // source of the next run of this warp inside the current point
    uint64_t pf_step = 0;              // words between two runs of this warp
    int64_t pf_left = 0, pf_gstep = 0; // groups from the next run's first group to the end of the point
    uint32_t pf_T = 1, pf_gbytes = 0;
    auto prefetch_one = [&]() {
        if (pf_r < b.n_runs) {
            if (pf_r >= pf_run_end) {
                const int pp = find_point_run(b.points, b.n_points, pf_r);
                const PointDev* pt = &b.points[pp];
                pf_run_end = pp + 1 < b.n_points ? b.points[pp + 1].run_start : b.n_runs;
                uint64_t pt_group_start, pt_run_start, pt_word_off;
                uint32_t pt_kp;
                if (pp == 0) {  // from the kernel parameters
                    pf_T = b.p0.tile, pt_group_start = b.p0.group_start, pt_run_start = b.p0.run_start;
                    pt_word_off = b.p0.word_off, pt_kp = b.p0.kp;
                } else {
                    pf_T = pt->tile, pt_group_start = pt->group_start, pt_run_start = pt->run_start;
                    pt_word_off = pt->word_off, pt_kp = pt->kp;
                }
                const uint64_t groups = (pp + 1 < b.n_points ? b.points[pp + 1].group_start : b.n_groups) - pt_group_start;
                const uint64_t g_rel = (pf_r - pt_run_start) * pf_T;  // first group of the run inside its point
                pf_ptr = b.words + pt_word_off + g_rel * pt_kp;
                pf_left = (int64_t)(groups - g_rel);
                pf_gstep = (int64_t)(stride * pf_T);
                pf_step = stride * pf_T * pt_kp;
                pf_gbytes = pt_kp * 8u;
            }
            const uint32_t bytes = (pf_left >= (int64_t)pf_T ? pf_T : (uint32_t)pf_left) * pf_gbytes;
            mbar_expect_tx_s(bar_s + pf_stage * 8u, bytes);
            bulk_g2s_s(ring_s + pf_stage * stage_bytes, pf_ptr, bytes, bar_s + pf_stage * 8u);
            pf_ptr += pf_step;
            pf_left -= pf_gstep;
        }
        pf_r += stride;
        pf_stage = pf_stage + 1 == (uint32_t)stages ? 0 : pf_stage + 1;
    };
    // Programmatic dependent launch: the host launches this kernel with programmatic stream
    // serialisation, so its CTAs may start while the kernel before it in the stream is draining.
    // Everything above touches only this CTA's shared memory and data that is constant since the
    // upload; from here on the kernel reads what the fill stage wrote and writes accumulators the
    // launch before may still be adding to: wait for the prerequisite grid to complete and flush.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0)
        for (int s = 0; s < stages; s++) prefetch_one();

    // the index blob is waited for by the groups that propagate (streaming points never look at
    // it) and, at the latest, before the CTA leaves
    // CTA accumulator of the final flush: the first warp to finish names the point
    __shared__ unsigned long long s_acc[8];
    __shared__ int s_point;
    if (threadIdx.x < 8) s_acc[threadIdx.x] = 0;
    if (threadIdx.x == 8) s_point = -1;
    if (b.zero_next != nullptr && blockIdx.x == 0)  // direct path: clear the accumulators of the next launch
        for (int i = threadIdx.x; i < b.n_points * 8; i += blockDim.x) b.zero_next[i] = 0;
    __syncthreads();  // also: the blob barrier init is visible to everybody
    if (b.peer.n_ranks > 0 && blockIdx.x == 0) peer_catch_up(b);  // while the other CTAs already work
    TL(1);

    // current point, cached in registers (a warp usually stays inside one point for many groups)
    int p = -1, p_prev = -1;
    uint64_t P_start = 0, P_end = 0, P_n_samples = 0, P_run_start = 0, P_run_end = 0;
    uint32_t k = 0, P_var_off = 0, P_n_assumed = 0, P_cand_off = 0, P_cand_cnt = 0, P_code0 = 0, P_flags = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (uint64_t run = r_first; run < b.n_runs; run += stride) {
        if (p < 0 || run >= P_run_end) {
            p = find_point_run(b.points, b.n_points, run);
            P_end = p + 1 < b.n_points ? b.points[p + 1].group_start : b.n_groups;
            P_run_end = p + 1 < b.n_points ? b.points[p + 1].run_start : b.n_runs;
            if (p == 0) {  // from the kernel parameters
                P_start = b.p0.group_start, P_run_start = b.p0.run_start, P_n_samples = b.p0.n_samples;
                k = b.p0.k, P_var_off = b.p0.var_off, P_n_assumed = b.p0.n_assumed_live;
                P_cand_off = b.p0.cand_off, P_cand_cnt = b.p0.cand_cnt, P_flags = b.p0.flags;
            } else {
                const PointDev* pt = &b.points[p];
                P_start = pt->group_start, P_run_start = pt->run_start, P_n_samples = pt->n_samples;
                k = pt->k, P_var_off = pt->var_off, P_n_assumed = pt->n_assumed_live;
                P_cand_off = pt->cand_off, P_cand_cnt = pt->cand_cnt, P_flags = pt->flags;
            }
            P_code0 = lane < (int)k ? b.setcode[P_var_off + lane] : 0u;
            flush_acc(accs, p_prev, b.acc, lane);
            p_prev = p;
        }
        // BCP can only start from a candidate record or a contradicted level-0 value (a set that
        // covers every live variable also goes the long way: its samples are models); the host
        // marks the points where it cannot
        if (P_flags & POINT_STREAMING) {
            // ---- streaming path: nothing can be implied or falsified in the groups of this
            // point.  Each tile is released as soon as it has landed; every sample has the trail
            // level-0 + assumptions (or is a conflict if the database itself is refuted).  The
            // loop runs over all runs of this warp inside the point.
            const uint32_t tr = db.n_base_assigned + P_n_assumed;
            const uint64_t g_last = P_end - 1;  // the only group of the point that can be ragged
            const uint32_t n_tail = (uint32_t)(P_n_samples - (g_last - P_start) * 64);
            const bool want_out = b.conflict_bits != nullptr || b.trail != nullptr;
            const uint32_t T = p == 0 ? b.p0.tile : b.points[p].tile;
            uint32_t cnt = 0;
            while (true) {
                mbar_wait_s(bar_s + (uint32_t)stage * 8u, phase);  // the tile has landed (and is not needed)
                __syncwarp();
#ifdef PDSAT_TIMELINE
                if (cnt == 0) TL(2);
#endif
                if (lane == 0) prefetch_one();
                stage = stage + 1 == stages ? 0 : stage + 1;
                phase ^= (stage == 0);
                const uint64_t g0 = P_start + (run - P_run_start) * T;
                const uint32_t ng = (uint32_t)min((uint64_t)T, P_end - g0);
                cnt += g0 + ng - 1 == g_last ? (ng - 1) * 64u + n_tail : ng * 64u;
                if (want_out) {
                    for (uint32_t gg = 0; gg < ng; gg++) {
                        const uint64_t g = g0 + gg;
                        const uint32_t nv = g == g_last ? n_tail : 64u;
                        const uint64_t vm = nv >= 64 ? ~0ull : ((1ull << nv) - 1);
                        if (b.conflict_bits && lane == 0) {
                            b.conflict_bits[g] = db.unsat ? vm : 0ull;
                            b.full_bits[g] = 0ull;
                        }
                        if (b.trail) {
                            b.trail[g * 64 + lane] = (!db.unsat && ((vm >> lane) & 1ull)) ? tr : 0u;
                            b.trail[g * 64 + 32 + lane] = (!db.unsat && ((vm >> (lane + 32)) & 1ull)) ? tr : 0u;
                        }
                    }
                }
                if (run + stride >= P_run_end) break;
                run += stride;
            }
            if (lane == 0) {
                accs[ACC_N] += cnt;
                if (db.unsat) {
                    accs[ACC_CONFLICT] += cnt;
                } else {
                    accs[ACC_SUM] += (uint64_t)cnt * tr;
                    accs[ACC_SUMSQ] += (uint64_t)cnt * tr * tr;
                }
            }
            continue;
        }

        // a point in which BCP can start has runs of one group
        if (db.blob_in_smem) mbar_wait(blob_bar, 0);  // returns at once after the first time
        const uint64_t gi = run - P_run_start;
        const uint64_t g = P_start + gi;
        const uint64_t s0 = gi * 64;
        const uint64_t left = P_n_samples - s0;
        const uint64_t valid64 = left >= 64 ? ~0ull : ((1ull << left) - 1);
        const uint32_t* code = b.setcode + P_var_off;
        const uint64_t* w = (const uint64_t*)(ring + (size_t)stage * b.ring_stage_bytes);
        mbar_wait_s(bar_s + (uint32_t)stage * 8u, phase);  // the group's words have landed
        uint64_t dead64 = 0, full64 = 0;
#pragma unroll 1
        for (int half = 0; half < HALVES; half++) {
            const PT valid = (PT)(valid64 >> (W * half));
            if (HALVES > 1 && !valid) break;  // ragged last group: the upper half is empty
            if (lane == 0) *c.dead = 0;
            __syncwarp();

            // ---- (1) assumptions (src_mpi/mpi_predicter.cpp:3236-3242: true bit => positive literal)
            for (uint32_t j = lane; j < k; j += 32) {
                const uint32_t cd = j < 32 ? P_code0 : code[j];
                const PT bits = (PT)(w[j] >> (W * half));
                if (cd & SETCODE_FIXED) {
                    // variable already assigned at level 0: a contradicting value is a
                    // conflict (addClause_ drops the false literal: empty clause)
                    const PT bad = ((cd & 1u) ? bits : (PT)~bits) & valid;
                    if (bad) Plane<PT>::or_shared(c.dead, bad);
                } else {
                    typename Plane<PT>::Pair tf;
                    tf.x = bits & valid;
                    tf.y = (PT)~bits & valid;
                    *reinterpret_cast<typename Plane<PT>::Pair*>(c.S + cd) = tf;  // cd is even: TRUE / FALSE pair
                }
            }
            __syncwarp();
            if (half == HALVES - 1 || !(PT)(valid64 >> (W * (HALVES - 1)))) {
                // the stage is free again: request the tile `stages` runs ahead
                if (lane == 0) prefetch_one();
                stage = stage + 1 == stages ? 0 : stage + 1;
                phase ^= (stage == 0);
            }

            uint32_t npops = 0;
            // ---- (2) first wave over the point's candidate records
            if (P_cand_cnt) {
                const PT live = valid & ~*(volatile PT*)c.dead;
                const uint4* list = b.cand + P_cand_off;
                uint4 r = __ldg(&list[min((uint32_t)lane, P_cand_cnt - 1)]);
                for (uint32_t i = lane; i < P_cand_cnt; i += 32) {
                    const uint4 cur = r;
                    if (i + 32 < P_cand_cnt) r = __ldg(&list[i + 32]);
                    eval_record(cur, db.rec, c, live, pad2);
                }
            }

            // ---- (3) level-synchronous propagation to the fixpoint.  The frontier of a level is
            // a bitset; every lane owns the words lane, lane + 32, ... of it and hands out one
            // literal per round.  The gate lists of the round's variables, then the clause lists
            // of the literals' complements, are visited as one flattened list each, one record
            // per lane per trip.
            bool stop = false;
            while (true) {  // levels until the frontier is empty, then the marked long clauses, then again
            while (!stop) {
                __syncwarp();
                {
                    uint32_t* t_ = c.fill;
                    c.fill = c.drain;
                    c.drain = t_;
                }
                uint32_t my_w = lane, my_bits = 0;
                auto fetch = [&]() {
                    while (my_bits == 0 && my_w < (uint32_t)inq_words) {
                        const uint32_t v = c.drain[my_w];
                        if (v) {
                            c.drain[my_w] = 0;
                            c.tset[my_w] |= v;
                            my_bits = v;
                        } else
                            my_w += 32;
                    }
                };
                fetch();
                if (!__any_sync(0xffffffffu, my_bits != 0)) break;
                do {
                    const PT deadv = *(volatile PT*)c.dead;
                    if ((deadv & valid) == valid) {
                        stop = true;
                        break;
                    }
                    const PT live = valid & ~deadv;
                    const bool have = my_bits != 0;
                    uint32_t myl = 0;
                    if (have) {
                        const int bp = __ffs(my_bits) - 1;
                        my_bits &= my_bits - 1;
                        myl = 32u * my_w + bp;
                    }
                    fetch();
                    npops += __popc(__ballot_sync(0xffffffffu, have));
                    if (db.n_gates) {
                        uint32_t gbeg = 0, glen = 0;
                        if (have) {
                            const uint32_t v = myl >> 1;
                            gbeg = gocc_off[v];
                            glen = gocc_off[v + 1] - gbeg;
                        }
                        const uint32_t cum = warp_incl_scan(glen, lane);
                        const uint32_t total = __shfl_sync(0xffffffffu, cum, 31);
                        const uint32_t gbase = gbeg - (cum - glen);  // list index = gbase + i
                        for (uint32_t i0 = 0; i0 < total; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            const int q = find_owner(cum, i);
                            const uint32_t bq = __shfl_sync(0xffffffffu, gbase, q & 31);
                            if (i < total) {
                                const uint4 r = grec[gocc[bq + i]];
                                eval_gate(r, c, live);
                            }
                        }
                    }
                    if (db.n_rec) {
                        uint32_t cbeg = 0, clen = 0;
                        if (have) {
                            const uint32_t nl = myl ^ 1;  // this literal just became false somewhere
                            cbeg = occ_off[nl];
                            clen = occ_off[nl + 1] - cbeg;
                        }
                        const uint32_t cum = warp_incl_scan(clen, lane);
                        const uint32_t total = __shfl_sync(0xffffffffu, cum, 31);
                        const uint32_t cbase = cbeg - (cum - clen);
                        // the record of the NEXT trip is requested before this trip's clause is
                        // evaluated: the L2 round trip of the occurrence lists hides behind the work
                        uint4 r_next = make_uint4(0, 0, 0, 0);
                        {
                            const int q = find_owner(cum, (uint32_t)lane);
                            const uint32_t bq = __shfl_sync(0xffffffffu, cbase, q & 31);
                            if ((uint32_t)lane < total) r_next = __ldg(&db.occrec[bq + lane]);
                        }
                        for (uint32_t i0 = 0; i0 < total; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            const uint4 r = r_next;
                            if (i0 + 32 < total) {
                                const uint32_t in = i + 32;
                                const int q = find_owner(cum, in);
                                const uint32_t bq = __shfl_sync(0xffffffffu, cbase, q & 31);
                                if (in < total) r_next = __ldg(&db.occrec[bq + in]);
                            }
                            if (i < total) eval_clause(r, db.rec, c, live, pad2);
                        }
                    }
                    if (db.n_long && have) {
                        // long clauses that contain the falsified literal: only marked here
                        const uint32_t nl = myl ^ 1;
                        for (uint32_t o = locc_off[nl]; o < locc_off[nl + 1]; o++) {
                            const uint32_t id = locc[o];
                            atomicOr(&lmark[id >> 5], 1u << (id & 31));
                        }
                    }
                    __syncwarp();
                } while (__any_sync(0xffffffffu, my_bits != 0));
            }
                if (stop || !db.n_long) break;
                // the frontier is empty: every marked long clause is evaluated once (lane L takes the
                // ids congruent to L); whatever they imply starts the levels again
                __syncwarp();
                bool any_marked = false;
                const PT live_l = valid & ~*(volatile PT*)c.dead;
                for (int w0 = 0; w0 < lm_words; w0++) {
                    const uint32_t word = lmark[w0];
                    if (!word) continue;
                    any_marked = true;
                    if ((word >> lane) & 1u) {
                        const uint4 r = make_uint4(lrec[32 * w0 + lane], 0u, 0u, REC_INDIRECT_D << 16);
                        eval_clause(r, db.rec, c, live_l, pad2);
                    }
                }
                __syncwarp();
                for (int w0 = lane; w0 < lm_words; w0 += 32) lmark[w0] = 0;
                if (!any_marked) break;
            }
            if (lane == 0) accs[ACC_POPS] += npops;
            __syncwarp();

            // ---- (4) per-sample results
            const PT dead = *c.dead & valid;
            const PT okm = valid & ~dead;
            uint32_t n_touched = 0;
            {
                // everything implied in this pass: visited literals + what an early stop left behind
                uint32_t lcnt = 0;
                for (int wi = lane; wi < inq_words; wi += 32) {
                    const uint32_t u = c.tset[wi] | c.fill[wi] | c.drain[wi];
                    c.tset[wi] = u;
                    lcnt += __popc(u);
                }
                n_touched = __reduce_add_sync(0xffffffffu, lcnt);
            }
            if (lane == 0) accs[ACC_TOUCHED] += n_touched;
            uint32_t cnt0, cnt1 = 0;  // assigned live variables of samples lane, lane+32
            if (n_touched == 0 || okm == 0) {  // counts only matter for conflict-free samples
                cnt0 = cnt1 = P_n_assumed;
            } else {
                // bit-sliced population count over the implied planes: per sample, how many
                // literals of the undo set are true (each implied variable has one polarity per
                // conflict-free sample)
                if (n_touched < 256) count_planes<8, PT>(c.S, c.tset, inq_words, lane, cnt0, cnt1);
                else if (n_touched < 2048) count_planes<11, PT>(c.S, c.tset, inq_words, lane, cnt0, cnt1);
                else count_planes<MAX_COUNT_BITS, PT>(c.S, c.tset, inq_words, lane, cnt0, cnt1);
                cnt0 += P_n_assumed;
                cnt1 += P_n_assumed;
            }
            const bool ok0 = (okm >> lane) & 1u;
            const bool ok1 = W == 64 ? (bool)(((uint64_t)okm >> (lane + 32)) & 1ull) : false;
            const uint32_t full_lo = __ballot_sync(0xffffffffu, ok0 && cnt0 == (uint32_t)db.n_live);
            const uint32_t full_hi = W == 64 ? __ballot_sync(0xffffffffu, ok1 && cnt1 == (uint32_t)db.n_live) : 0u;
            const uint64_t fullm = ((uint64_t)full_hi << 32) | full_lo;  // bits of this pass
            const uint32_t t0 = db.n_base_assigned + cnt0, t1 = db.n_base_assigned + cnt1;
            if (n_touched == 0 || okm == 0) {
                const uint64_t nok = Plane<PT>::popc(okm);
                if (lane == 0) {
                    accs[ACC_SUM] += nok * t0;
                    accs[ACC_SUMSQ] += nok * (uint64_t)t0 * t0;
                }
            } else {
                // 64-bit throughout: trail can reach n_vars, its square summed over 64 samples
                // does not fit 32 bits from trail = 8192 on
                const uint64_t lsum = (ok0 ? (uint64_t)t0 : 0ull) + (ok1 ? (uint64_t)t1 : 0ull);
                const uint64_t lsq = (ok0 ? (uint64_t)t0 * t0 : 0ull) + (ok1 ? (uint64_t)t1 * t1 : 0ull);
                const uint32_t limp = (ok0 && cnt0 > P_n_assumed ? 1 : 0) + (ok1 && cnt1 > P_n_assumed ? 1 : 0);
                const uint64_t ws = warp_sum64(lsum), wq = warp_sum64(lsq);
                const uint32_t wi_ = __reduce_add_sync(0xffffffffu, limp);
                if (lane == 0) {
                    accs[ACC_SUM] += ws;
                    accs[ACC_SUMSQ] += wq;
                    accs[ACC_IMPL] += wi_;
                }
            }
            if (lane == 0) {
                accs[ACC_N] += Plane<PT>::popc(valid);
                accs[ACC_CONFLICT] += Plane<PT>::popc(dead);
                accs[ACC_FULL] += __popcll(fullm);
            }
            dead64 |= (uint64_t)dead << (W * half);
            full64 |= fullm << (W * half);

            if (b.trail) {
                b.trail[g * 64 + W * half + lane] = ok0 ? t0 : 0u;
                if (W == 64) b.trail[g * 64 + 32 + lane] = ok1 ? t1 : 0u;
            }
            if (b.assign) {
                PT* dst = (PT*)(b.assign + g * (uint64_t)db.n_lit);  // 64-bit planes per literal; a half fills its word half
                for (int i = lane; i < db.n_lit; i += 32) dst[(size_t)i * HALVES + half] = c.S[i];
            }
            if (fullm && b.max_models > 0) {
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
                            const uint32_t bit = v < db.n_live ? (uint32_t)((c.S[2 * v] >> s) & 1u) : 0u;
                            const uint32_t word = __ballot_sync(0xffffffffu, bit);
                            if (lane == 0) b.model_bits[idx * b.model_words + (v0 >> 5)] = word;
                        }
                        if (lane == 0) {
                            b.model_point[idx] = (uint32_t)p;
                            b.model_sample[idx] = s0 + (uint64_t)(W * half) + s;
                        }
                    }
                }
            }
            __syncwarp();

            // ---- (5) undo: clear the assumption planes, the implied planes and the frontier bitsets
            for (uint32_t j = lane; j < k; j += 32) {
                const uint32_t cd = j < 32 ? P_code0 : code[j];
                if (!(cd & SETCODE_FIXED)) {
                    typename Plane<PT>::Pair z;
                    z.x = z.y = 0;
                    *reinterpret_cast<typename Plane<PT>::Pair*>(c.S + cd) = z;
                }
            }
            if (db.n_long)
                for (int w0 = lane; w0 < lm_words; w0 += 32) lmark[w0] = 0;
            if (n_touched) {
                for (int wi = lane; wi < inq_words; wi += 32) {
                    uint32_t bits = c.tset[wi];
                    while (bits) {
                        const int bp = __ffs(bits) - 1;
                        bits &= bits - 1;
                        c.S[32 * wi + bp] = 0;
                    }
                    c.tset[wi] = 0;
                    c.fill[wi] = 0;
                    c.drain[wi] = 0;
                }
            }
            __syncwarp();
        }
        if (b.conflict_bits && lane == 0) {
            b.conflict_bits[g] = dead64;
            b.full_bits[g] = full64;
        }
    }
    TL(3);
    if (db.blob_in_smem) mbar_wait(blob_bar, 0);  // no bulk copy may be in flight when the CTA exits
    TL(4);

    // ---- final flush: the warps that end in the same point as the first warp to finish (usually
    // all of them) add into the CTA accumulator - 8 global atomics per CTA instead of 8 per warp on
    // the same two cache lines (a serial tail of several microseconds on a 20 us launch)
    __shared__ unsigned int s_ticket;
    {
        int cta_point = -1;
        if (p_prev >= 0) {
            if (lane == 0) {
                const int old = atomicCAS(&s_point, -1, p_prev);
                cta_point = old < 0 ? p_prev : old;
            }
            cta_point = __shfl_sync(0xffffffffu, cta_point, 0);
            if (cta_point == p_prev) {
                if (lane < 8) {
                    const unsigned long long v = accs[lane];
                    if (v) atomicAdd(&s_acc[lane], v);
                }
            } else {
                flush_acc(accs, p_prev, b.acc, lane);
                __threadfence();
            }
        }
        __syncthreads();
        TL(5);
        if (warp != 0) {
            if (b.no_ticket) return;  // direct path / peers: nothing to publish here
        } else {
            if (lane < 8 && s_point >= 0 && s_acc[lane]) atomicAdd(&b.acc[(size_t)s_point * 8 + lane], s_acc[lane]);
            if (b.no_ticket) return;
            // ---- ticket path (caller-owned cost buffer, no peers): the last CTA publishes the
            // scratch accumulators and clears them
            __threadfence();
            __syncwarp();
            if (lane == 0) s_ticket = atomicAdd(b.cta_done, 1u);
        }
        __syncthreads();
        TL(6);
        if (s_ticket != gridDim.x - 1) return;
        __threadfence();
        const int nw = b.n_points * 8;
        {
            for (int i = threadIdx.x; i < nw; i += blockDim.x) {
                b.cost[i] = __ldcg(&b.acc[i]);
                b.acc[i] = 0;
            }
        }
        if (threadIdx.x == 0) *b.cta_done = 0;
        TL(7);
    }
}

__global__ void __launch_bounds__(512, 1) bcp_kernel(DbDev db, BatchDev b, int warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t blob_bar;
    bcp_body<uint64_t>(db, b, warp_bytes, smem, &blob_bar);
}

// 32-bit planes: half the shared memory per warp.  Chosen only where fewer than 4 warps of full
// planes fit (launch_shape), i.e. for large instances with a handful of warps per CTA: 8 warps at
// most, which leaves the compiler 255 registers (at 768 threads this kernel spilled)
__global__ void __launch_bounds__(256, 1) bcp_kernel_h(DbDev db, BatchDev b, int warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t blob_bar;
    bcp_body<uint32_t>(db, b, warp_bytes, smem, &blob_bar);
}

}  // namespace pdsat
