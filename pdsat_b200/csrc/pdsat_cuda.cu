// pdsat_cuda.cu — implementation of the C-ABI declared in include/pdsat_cuda.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/pdsat_cuda.h"
#include "handoff.hpp"
#include "host_db.hpp"
#include "kernels.cuh"
#include "mt_jump.hpp"

using namespace pdsat;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(PDSAT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct LaunchShape {
    int stages = 0, stage_bytes = 0, warp_bytes = 0, wpc = 0, smem_bytes = 0, blob_in_smem = 0;
    int tile_groups = 1;  // groups per run = per bulk copy
    int plane_bits = 64;  // 64: bcp_kernel, 32: bcp_kernel_h (two half passes per group)
};

struct pdsat_db {
    HostDb host;
    int device = -1;  // -1: host-only handle (CPU tests of the upload logic)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    // device copies
    uint4* d_rec = nullptr;
    unsigned char* d_blob = nullptr;
    uint4* d_occrec = nullptr;
    DbDev dev{};
    // launch shape (per-warp bytes without the words ring, which depends on the batch's set sizes)
    int warps_per_cta = 0, warp_bytes = 0, qcap = 0, smem_bytes = 0, grid = 0;
    // set_fixed_assumptions renumbers the live variables: batches created before it are stale
    uint64_t generation = 0;
    CdclPool cdcl;  // host solvers for the hand-off (lazy)
    int live_batches = 0;
    bool zombie = false;  // pdsat_destroy was called while batches were alive
    // mt19937 jump polynomials x^(624*rounds) mod phi resident on the device (lazy), by round
    // count: powers of two and small multiples of them (the levels of the jump tree), and the
    // composite polynomials of stream offsets that keep coming back (a rank's shard offset)
    uint32_t* d_polys = nullptr;
    std::map<uint64_t, int> poly_slots, comp_seen;
};
static constexpr int POLY_CAP = 512;            // slots of d_polys
static constexpr int POLY_CAP_COMPOSITE = 448;  // composite offsets stop here: the rest stays free for the tree

// slot of x^(624*rounds) mod phi on the device (uploaded on first use); -1 = table full
static int mt_poly_slot(pdsat_db* db, uint64_t rounds, cudaStream_t st) {
    auto it = db->poly_slots.find(rounds);
    if (it != db->poly_slots.end()) return it->second;
    if ((int)db->poly_slots.size() >= POLY_CAP) return -1;
    if (!db->d_polys && cudaMalloc(&db->d_polys, sizeof(uint32_t) * 624 * POLY_CAP) != cudaSuccess) return -1;
    const uint32_t* pl = __builtin_popcountll(rounds) == 1 ? MtJumpPolys::instance().poly(__builtin_ctzll(rounds))
                                                           : MtJumpPolys::instance().poly_for_rounds(rounds);
    const int slot = (int)db->poly_slots.size();
    if (cudaMemcpyAsync(db->d_polys + (size_t)slot * 624, pl, sizeof(uint32_t) * 624, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return -1;
    db->poly_slots.emplace(rounds, slot);
    return slot;
}

struct HostPoint {
    std::vector<int32_t> set_vars;
    std::vector<uint32_t> setcode;
    int32_t mode = 0;
    uint64_t seed = 0, first_sample = 0, n_samples = 0, first_draw = 0;
    uint32_t kp = 0, flags = 0, block_bits = 0, tile = 1;
    uint64_t block_off = 0, block_cap = 0;  // ENUM_BLOCKS: offset / capacity in the prefix buffer
    uint64_t n_groups = 0, group_start = 0, word_off = 0, stream_off = 0, stream_skip = 0, stream_words = 0;
    uint64_t mt_first_round = 0, mt_rounds = 0;
    uint32_t n_assumed_live = 0;
    std::vector<uint32_t> cand;   // first-wave candidate clauses (first-record ids)
    std::vector<uint32_t> gcand;  // first-wave candidate gates
    uint32_t cand_off = 0;
    // mt19937 stream segments of 2^mt_m rounds each
    uint32_t mt_index = 0;      // index among the batch's mt19937 points (= state slot of segment 0)
    uint32_t n_segs = 0;
    uint32_t extra_base = 0;    // state slot of segment 1
    uint64_t seg_base = 0;      // global index of segment 0
};

struct pdsat_batch {
    pdsat_db* db = nullptr;
    std::vector<HostPoint> pts;
    uint32_t out_flags = 0;
    int32_t max_models = 0;
    int32_t model_words = 0;
    uint64_t n_groups = 0, n_words = 0;
    // device
    PointDev* d_points = nullptr;
    uint32_t* d_setcode = nullptr;
    uint4* d_cand = nullptr;
    uint64_t* d_words = nullptr;
    uint32_t* d_stream = nullptr;
    unsigned long long* d_cost = nullptr;   // published accumulators (ticket path), or buffer 0 of the direct path
    unsigned long long* d_cost2 = nullptr;  // buffer 1 of the direct path
    unsigned long long* d_cost_ext = nullptr;
    // Direct path (no peers, no caller-owned cost buffer): the CTAs add straight into
    // d_cost / d_cost2 alternately and CTA 0 clears the other one for the next launch - no ticket,
    // no publishing pass at the end of the kernel.
    bool timing = true, fill_timed = false, propagate_timed = false;  // pdsat_batch_set_timing
    int cost_cur = 0;          // buffer the last direct launch wrote
    bool last_direct = false;  // the last launch took the direct path
    int last_path = 1;         // 0 direct, 1 ticket, 2 peers (buffers are zeroed when it changes)
    bool no_direct = false;    // the caller holds a device pointer to the costs: keep it stable
    unsigned long long* cost_ptr() const {
        if (last_direct) return cost_cur ? d_cost2 : d_cost;
        return d_cost_ext ? d_cost_ext : d_cost;
    }
    unsigned long long* d_acc = nullptr;  // scratch accumulators, zero between launches
    uint64_t* d_conflict = nullptr;
    uint64_t* d_full = nullptr;
    uint32_t* d_trail = nullptr;
    uint64_t* d_assign = nullptr;
    unsigned long long* d_model_count = nullptr;
    uint32_t* d_model_bits = nullptr;
    uint32_t* d_model_point = nullptr;
    uint64_t* d_model_sample = nullptr;
    // mt19937 segments
    uint32_t* d_mt_states = nullptr;
    MtSegment* d_mt_segs = nullptr;
    uint32_t* h_mt_states = nullptr;  // pinned: seed states of the mt19937 points
    MtSegment* h_mt_segs = nullptr;   // pinned
    MtJumpTask* d_mt_tasks = nullptr;
    MtJumpTask* h_mt_tasks = nullptr; // pinned
    uint32_t* d_mt_srcs = nullptr;    // source slots of the sequences, level by level
    uint32_t* h_mt_srcs = nullptr;    // pinned
    size_t mt_tasks_cap = 0;
    int n_mt_points = 0;
    int n_mt_segs = 0;
    int mt_m = 1;                     // log2(rounds per segment)
    size_t n_mt_states = 0;
    // explicit values, packed on the host (pinned)
    uint32_t* h_stream = nullptr;
    bool has_explicit = false;
    // ENUM_BLOCKS prefixes (pinned staging + device)
    uint64_t* h_blocks = nullptr;
    uint64_t* d_blocks = nullptr;
    uint64_t n_blocks_total = 0;
    PointDev* h_points = nullptr;  // pinned mirror
    // fused peer reduce (multi-GPU)
    unsigned long long* d_peer_own = nullptr;          // this rank's IPC-exported buffer
    unsigned long long* peer_ptr[MAX_PEERS] = {};      // every rank's buffer as mapped here
    unsigned int* d_cta_done = nullptr;
    bool counted_mt = false;       // counted in g_mt_batches
    uint32_t* d_mt_seq = nullptr;  // sequences of one level of jump-aheads (mt19937_seq_kernel)
    size_t mt_seq_cap = 0;
    int peer_ranks = 0, peer_me = 0;
    uint64_t peer_step = 0;
    // deferred exchange (PeerDev): after launch t the sums of step t sit in d_acc / d_acc2 unsent
    // (peer_unsent) and, if step t-1 was not finished by the host, its packets are still to be
    // gathered (peer_ungathered); the next launch's CTA 0 or peer_finalize catches up
    bool peer_unsent = false, peer_ungathered = false;
    unsigned long long* d_acc2 = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d = nullptr;  // the last fill's copies out of the pinned staging buffers
    bool h2d_pending = false;
    uint64_t generation = 0;       // of the database at creation
    LaunchShape shape;
    bool filled = false, propagated = false;
    int64_t n_kernels = 0;
    int64_t h2d_bytes_last_fill = 0;
};

extern "C" {

const char* pdsat_last_error(void) { return g_err.c_str(); }

int pdsat_device_count(int* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(PDSAT_ERR_CUDA, cudaGetErrorString(e));
    }
    *n = c;
    return PDSAT_OK;
}

static void free_device_db(pdsat_db* db) {
    if (db->device < 0) return;
    cudaFree(db->d_rec);
    cudaFree(db->d_blob);
    cudaFree(db->d_occrec);
    db->d_rec = nullptr;
    db->d_blob = nullptr;
    db->d_occrec = nullptr;
}

static int count_bits_for(int n) {
    int b = 1;
    while ((1 << b) <= n) b++;
    return b;
}

static constexpr int MAX_SMEM = 227 * 1024 - 128;  // the kernel also has a few bytes of static shared memory

// launch shape of the BCP kernel for a batch whose largest set has kp_max (even) words per group
// groups per run (= per bulk copy) of a streaming point with kp words per group: ~2 KB per copy
// (the producer lane's issue work is per copy)
static uint32_t tile_groups_for(uint32_t kp) {
    if (const char* e = getenv("PDSAT_TILE_GROUPS")) return atoi(e) > 0 ? (uint32_t)atoi(e) : 1u;
    const uint32_t tg = 2048u / (kp * 8u ? kp * 8u : 16u);
    return tg < 1 ? 1 : (tg > 8 ? 8 : tg);
}

// stage_bytes = the largest tile of the batch (a point in which BCP can start has tiles of one
// group, a streaming point of tile_groups_for(kp) groups)
static LaunchShape launch_shape_bits(const pdsat_db* db, uint32_t stage_bytes, int plane_bits) {
    LaunchShape L;
    L.plane_bits = plane_bits;
    // 3 KB of tiles in flight per warp (x 12..16 warps x 148 SMs covers the HBM latency x
    // bandwidth product); at least double buffering for the large sets
    L.stage_bytes = (int)stage_bytes;
    int ring_budget = 3072;
    if (const char* e = getenv("PDSAT_RING_BYTES")) ring_budget = atoi(e) > 0 ? atoi(e) : ring_budget;
    int st = ring_budget / (L.stage_bytes > 0 ? L.stage_bytes : 16);
    L.stages = st < 2 ? 2 : (st > 8 ? 8 : st);
    const int ring = ((L.stages + 1) & ~1) * 8 + L.stages * L.stage_bytes;
    const int pb = plane_bits / 8;
    const int n_lit = 2 * db->host.n_live, inq_words = (n_lit + 31) / 32;
    // planes + three bitsets (two frontiers, undo set) + the conflict mask (+ alignment slack)
    const int lm_words = ((int)db->host.long_rec.size() + 31) / 32;
    // planes, 3 frontier / undo bitsets, long-clause marks, conflict mask, alignment, 8 point accumulators
    const int state = (n_lit + 4) * pb + 3 * ((inq_words + 1) & ~1) * 4 + ((lm_words + 1) & ~1) * 4 + 2 * pb + 16 + 64;
    L.warp_bytes = (state + ring + 15) & ~15;
    const int blob = (int)db->host.blob.size();
    // the index blob is staged in shared memory only where that costs no resident warp (measured:
    // ODLS 4 warps + staged blob 4.36e8 samples/s, 6 warps + blob through L1/L2 4.80e8; Bivium has
    // room for both)
    const int cap = plane_bits == 64 ? 16 : 8;  // __launch_bounds__(512) / (256)
    const int wpc_without = std::min(cap, MAX_SMEM / L.warp_bytes);
    const int wpc_with = blob + L.warp_bytes <= MAX_SMEM ? std::min(cap, (MAX_SMEM - blob) / L.warp_bytes) : 0;
    L.blob_in_smem = (blob <= 64 * 1024 && wpc_with >= 1 && wpc_with == wpc_without) ? 1 : 0;
    if (const char* e = getenv("PDSAT_BLOB_SMEM")) L.blob_in_smem = (atoi(e) != 0 && wpc_with >= 1) ? 1 : 0;
    int wpc = L.blob_in_smem ? wpc_with : wpc_without;
    L.wpc = wpc;
    L.smem_bytes = wpc * L.warp_bytes + (L.blob_in_smem ? blob : 0);
    return L;
}

// 64-bit planes (one pass per 64-sample group) unless the per-warp state is so large that fewer
// than 4 warps fit an SM: then half-size planes (two passes per group) multiply the resident warps.
// Measured (B200, profiles/r02_plane_width.md): A5/2 (7 087 live variables, 1 -> 3 warps) 1.46x
// faster with half planes; ODLS (5 -> 11 warps) and Bivium (16 -> 24) slower, because a literal
// implied in both halves is visited twice (ODLS: 1.97x the literal visits).
// PDSAT_PLANE_BITS=32|64 overrides (experiments, before/after profiles).
static LaunchShape launch_shape(const pdsat_db* db, uint32_t stage_bytes) {
    const char* e = getenv("PDSAT_PLANE_BITS");
    const int forced = e ? atoi(e) : 0;
    if (forced == 32 || forced == 64) return launch_shape_bits(db, stage_bytes, forced);
    const LaunchShape L64 = launch_shape_bits(db, stage_bytes, 64);
    if (L64.wpc >= 4) return L64;
    const LaunchShape L32 = launch_shape_bits(db, stage_bytes, 32);
    return L32.wpc > L64.wpc ? L32 : L64;
}

// (re)build the device copy + launch shape from db->host
static int sync_device_db(pdsat_db* db) {
    HostDb& h = db->host;
    db->generation++;
    if (2 * (int64_t)h.n_live + 4 >= REC_TAG_MIN)
        return fail(PDSAT_ERR_LIMIT, "more than 32757 live variables: 16-bit literal codes exhausted");
    if (count_bits_for(h.n_live) > MAX_COUNT_BITS) return fail(PDSAT_ERR_LIMIT, "too many live variables for the counters");
    int n_lit = 2 * h.n_live;
    int n_planes = n_lit + 4;
    (void)n_planes;
    db->qcap = 0;
    db->warp_bytes = 0;
    if (db->device < 0) {
        db->warps_per_cta = 0;
        return PDSAT_OK;
    }
    CU(cudaSetDevice(db->device));
    free_device_db(db);
    // shape for a small set, reported by pdsat_db_get_info; wpc == 0: the planes of one
    // 64-sample group do not fit in shared memory with the current level-0 assignment;
    // evaluation is refused (PDSAT_ERR_LIMIT) until fixed assumptions shrink the live set
    const LaunchShape L = launch_shape(db, 32 * 8);
    db->warps_per_cta = L.wpc;
    db->smem_bytes = L.smem_bytes;
    db->grid = db->sm_count;
    size_t nrec = (size_t)h.n_rec;
    CU(cudaMalloc(&db->d_rec, (nrec ? nrec : 1) * sizeof(uint4)));
    CU(cudaMalloc(&db->d_blob, h.blob.size()));
    const size_t n_docc = h.occrec.size() / REC_SLOTS;
    CU(cudaMalloc(&db->d_occrec, (n_docc ? n_docc : 1) * sizeof(uint4)));
    if (nrec) CU(cudaMemcpy(db->d_rec, h.rec.data(), nrec * sizeof(uint4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db->d_blob, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice));
    if (n_docc) CU(cudaMemcpy(db->d_occrec, h.occrec.data(), n_docc * sizeof(uint4), cudaMemcpyHostToDevice));
    db->dev.rec = db->d_rec;
    db->dev.occrec = db->d_occrec;
    db->dev.blob = db->d_blob;
    db->dev.blob_bytes = (uint32_t)h.blob.size();
    db->dev.off_gocc_off = h.blob_gocc_off;
    db->dev.off_gocc = h.blob_gocc;
    db->dev.off_grec = h.blob_grec;
    db->dev.off_locc_off = h.blob_locc_off;
    db->dev.off_locc = h.blob_locc;
    db->dev.off_lrec = h.blob_lrec;
    db->dev.blob_in_smem = L.blob_in_smem;
    db->dev.n_live = h.n_live;
    db->dev.n_lit = n_lit;
    db->dev.n_base_assigned = h.n_base_assigned;
    db->dev.unsat = h.base_ok ? 0 : 1;
    db->dev.n_gates = (int32_t)h.gates.size();
    db->dev.n_rec = (int32_t)n_docc;
    db->dev.n_long = (int32_t)h.long_rec.size();
    CU(cudaFuncSetAttribute(bcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    CU(cudaFuncSetAttribute(bcp_kernel_h, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    return PDSAT_OK;
}

int pdsat_upload_cnf(int device, int32_t n_vars, int64_t n_clauses, const uint32_t* offsets, const int32_t* lits,
                     pdsat_db** out) {
    if (!out || n_vars < 0 || n_clauses < 0 || (n_clauses > 0 && (!offsets || !lits)))
        return fail(PDSAT_ERR_ARG, "pdsat_upload_cnf: bad argument");
    *out = nullptr;
    pdsat_db* db = new pdsat_db();
    db->device = device;
    if (device >= 0) {
        int cnt = 0;
        cudaError_t e = cudaGetDeviceCount(&cnt);
        if (e != cudaSuccess || device >= cnt) {
            delete db;
            return fail(PDSAT_ERR_CUDA, e != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(e)
                                                         : std::string("device index out of range"));
        }
        cudaSetDevice(device);
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, device);
        db->sm_count = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete db;
            return fail(PDSAT_ERR_CUDA, "cudaStreamCreate failed");
        }
        db->own_stream = true;
    }
    static const uint32_t zero_off[1] = {0};
    if (!db->host.load(n_vars, n_clauses, n_clauses ? offsets : zero_off, lits)) {
        std::string m = db->host.error;
        pdsat_destroy(db);
        return fail(PDSAT_ERR_ARG, m);
    }
    int rc = sync_device_db(db);
    if (rc != PDSAT_OK) {
        pdsat_destroy(db);
        return rc;
    }
    *out = db;
    return PDSAT_OK;
}

int pdsat_set_fixed_assumptions(pdsat_db* db, const int32_t* lits, int32_t n) {
    if (!db || n < 0 || (n > 0 && !lits)) return fail(PDSAT_ERR_ARG, "pdsat_set_fixed_assumptions: bad argument");
    if (!db->host.set_fixed(lits, n)) return fail(PDSAT_ERR_ARG, db->host.error);
    return sync_device_db(db);
}

int pdsat_db_get_info(const pdsat_db* db, pdsat_db_info* info) {
    if (!db || !info) return fail(PDSAT_ERR_ARG, "pdsat_db_get_info: bad argument");
    const HostDb& h = db->host;
    info->n_vars = h.n_vars;
    info->n_live_vars = h.n_live;
    info->n_clauses_in = h.n_clauses_in;
    info->n_clauses_live = h.n_clauses_live;
    info->n_lits_live = h.n_lits_live;
    info->n_base_assigned = h.n_base_assigned;
    info->max_clause_len = h.max_clause_len;
    info->status = h.base_ok ? PDSAT_OK : PDSAT_ERR_UNSAT;
    info->warps_per_cta = db->warps_per_cta;
    info->smem_bytes = db->smem_bytes;
    info->grid = db->grid;
    info->db_in_smem = db->dev.blob_in_smem;
    info->sm_count = db->sm_count;
    info->n_gates = (int32_t)h.gates.size();
    info->blob_bytes = (int32_t)h.blob.size();
    info->n_gate_clauses = h.n_gate_clauses;
    return PDSAT_OK;
}

int pdsat_db_base_assignment(const pdsat_db* db, uint8_t* assigns) {
    if (!db || !assigns) return fail(PDSAT_ERR_ARG, "pdsat_db_base_assignment: bad argument");
    memcpy(assigns, db->host.base_assign.data(), (size_t)db->host.n_vars);
    return PDSAT_OK;
}

int pdsat_db_gates(const pdsat_db* db, int32_t cap, pdsat_gate* out, int32_t* n_gates) {
    if (!db || !n_gates) return fail(PDSAT_ERR_ARG, "pdsat_db_gates: bad argument");
    const HostDb& h = db->host;
    *n_gates = (int32_t)h.gates.size();
    for (int32_t g = 0; out && g < cap && g < (int32_t)h.gates.size(); g++) {
        const HostGate& hg = h.gates[g];
        pdsat_gate& o = out[g];
        memset(&o, 0, sizeof(o));
        o.type = hg.type;
        o.parity = hg.parity;
        o.m = hg.m;
        o.n_clauses = (int32_t)hg.n_clauses;
        for (int i = 0; i < hg.m; i++) o.x[i] = h.var_of_live[hg.x[i]] + 1;
        o.lb = o.lc = -1;
        if (hg.type == 1) {
            o.lb = 2 * h.var_of_live[hg.lb >> 1] + (hg.lb & 1);
            o.lc = 2 * h.var_of_live[hg.lc >> 1] + (hg.lc & 1);
        }
    }
    return PDSAT_OK;
}

int pdsat_set_stream(pdsat_db* db, void* cuda_stream) {
    if (!db) return fail(PDSAT_ERR_ARG, "pdsat_set_stream: null handle");
    if (db->device < 0) return fail(PDSAT_ERR_CUDA, "host-only handle");
    if (db->own_stream && db->stream) {
        cudaStreamSynchronize(db->stream);
        cudaStreamDestroy(db->stream);
    }
    if (cuda_stream) {
        db->stream = (cudaStream_t)cuda_stream;
        db->own_stream = false;
    } else {
        CU(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
        db->own_stream = true;
    }
    return PDSAT_OK;
}

void pdsat_destroy(pdsat_db* db) {
    if (!db) return;
    if (db->live_batches > 0) {  // released by the last pdsat_batch_destroy
        db->zombie = true;
        return;
    }
    if (db->device >= 0) {
        cudaSetDevice(db->device);
        free_device_db(db);
        cudaFree(db->d_polys);
        if (db->own_stream && db->stream) cudaStreamDestroy(db->stream);
    }
    delete db;
}

// ------------------------------------------------------------------------- batches
static int need_device(const pdsat_db* db) {
    if (!db) return fail(PDSAT_ERR_ARG, "null database handle");
    if (db->device < 0) return fail(PDSAT_ERR_CUDA, "host-only handle: no CUDA device bound (there is no CPU fallback)");
    return PDSAT_OK;
}

// position of a point's sample window inside its mt19937 stream
static void mt_plan(HostPoint& hp) {
    const uint64_t k = hp.set_vars.size();
    const uint64_t d0 = hp.first_draw + hp.first_sample * k;
    hp.mt_first_round = d0 / 624;
    hp.stream_skip = d0 - hp.mt_first_round * 624;
    // rounds to allocate: worst-case skip (623), so that a reseed never needs more
    hp.mt_rounds = (623 + hp.n_samples * k + 623) / 624;
}

// Segments (= generator warps) a batch cuts its mt19937 streams into.  More segments keep more
// schedulers busy in mt19937_bits_kernel but every segment costs one jump-ahead: a batch that is
// alone on its device gets ~4 per SM; when four or more batches with mt19937 points are alive on
// the device (a pipeline of handles, several search workers) their generator kernels overlap on
// the SMs and half as many segments per batch is the better trade (measured: profiles/README.md).
static std::atomic<int> g_mt_batches[64];
static int mt_target_segments(int device) {
    const char* e = getenv("PDSAT_MT_SEGMENTS");
    int t = e ? atoi(e) : 0;
    if (t > 0) return t;
    const int alive = device >= 0 && device < 64 ? g_mt_batches[device].load() : 1;
    return alive >= 4 ? 296 : 592;
}

int pdsat_batch_create(pdsat_db* db, int32_t n_points, const pdsat_point* points, uint32_t out_flags,
                       int32_t max_models, pdsat_batch** out) {
    int rc = need_device(db);
    if (rc) return rc;
    if (!out || n_points <= 0 || !points) return fail(PDSAT_ERR_ARG, "pdsat_batch_create: bad argument");
    if (db->warps_per_cta < 1)
        return fail(PDSAT_ERR_LIMIT, "assignment state of one sample group exceeds shared memory (" +
                                         std::to_string(db->host.n_live) + " live variables)");
    *out = nullptr;
    const HostDb& h = db->host;
    pdsat_batch* b = new pdsat_batch();
    b->db = db;
    b->generation = db->generation;
    db->live_batches++;
    b->out_flags = out_flags;
    b->max_models = max_models < 0 ? 0 : max_models;
    b->model_words = (h.n_live + 31) / 32;
    if (b->model_words == 0) b->model_words = 1;
    b->pts.resize(n_points);
    uint64_t groups = 0, words = 0, stream_words = 0;
    size_t n_codes = 0, n_cand = 0;
    std::vector<uint8_t> seen((size_t)h.n_vars);
    std::vector<uint16_t> in_set_count((size_t)h.n_rec + 1, 0);
    std::vector<uint32_t> touched_rec;
    std::vector<uint8_t> live_in_set((size_t)h.n_live + 1, 0), gate_seen(h.gates.size() + 1, 0);
    uint32_t kp_max = 2;
    for (int p = 0; p < n_points; p++) {
        const pdsat_point& in = points[p];
        HostPoint& hp = b->pts[p];
        if (in.k <= 0 || !in.set_vars || in.n_samples == 0 || in.mode < 0 || in.mode > 4 ||
            (in.mode == PDSAT_MODE_EXPLICIT && !in.explicit_values) ||
            (in.mode == PDSAT_MODE_ENUM_BLOCKS && (!in.block_prefix || in.block_bits < 6 || in.block_bits > 62 ||
                                                  (in.n_samples & ((1ull << in.block_bits) - 1)))) ||
            (in.mode == PDSAT_MODE_RANDOM_COUNTER && (in.first_sample & 63))) {
            pdsat_batch_destroy(b);
            return fail(PDSAT_ERR_ARG, "pdsat_batch_create: bad point " + std::to_string(p));
        }
        std::fill(seen.begin(), seen.end(), 0);
        hp.set_vars.assign(in.set_vars, in.set_vars + in.k);
        hp.setcode.resize(in.k);
        hp.n_assumed_live = 0;
        for (int j = 0; j < in.k; j++) {
            int32_t v = in.set_vars[j] - 1;
            if (v < 0 || v >= h.n_vars || seen[v]) {
                pdsat_batch_destroy(b);
                return fail(PDSAT_ERR_ARG, "pdsat_batch_create: set variable out of range or repeated");
            }
            seen[v] = 1;
            if (h.live_of_var[v] < 0) {
                hp.setcode[j] = SETCODE_FIXED | (uint32_t)(h.base_assign[v] & 1);
                hp.flags |= POINT_HAS_FIXED;
            } else {
                hp.setcode[j] = (uint32_t)(2 * h.live_of_var[v]);
                hp.n_assumed_live++;
            }
        }
        // first-wave candidates: clauses with at most one literal outside the set
        if (h.base_ok) {
            touched_rec.clear();
            for (int j = 0; j < in.k; j++) {
                const int32_t lv = h.live_of_var[in.set_vars[j] - 1];
                if (lv < 0) continue;
                for (int sgn = 0; sgn < 2; sgn++) {
                    const uint32_t l = (uint32_t)(2 * lv + sgn);
                    for (uint32_t o = h.occ_off[l]; o < h.occ_off[l + 1]; o++) {
                        const uint32_t r = h.occ[o];
                        if (in_set_count[r]++ == 0) touched_rec.push_back(r);
                    }
                }
            }
            for (uint32_t r : touched_rec) {
                if ((int)in_set_count[r] + 1 >= (int)h.rec_len[r]) hp.cand.push_back(r);
                in_set_count[r] = 0;
            }
            std::sort(hp.cand.begin(), hp.cand.end());
            // gates: one that has two unknown members cannot imply or conflict.  Members are the
            // xor variables and, for ANDXOR, the product Lb & Lc (possibly known as soon as one
            // of its literals is in the set)
            if (!h.gates.empty()) {
                for (int j = 0; j < in.k; j++) {
                    const int32_t lv = h.live_of_var[in.set_vars[j] - 1];
                    if (lv >= 0) live_in_set[lv] = 1;
                }
                for (int j = 0; j < in.k; j++) {
                    const int32_t lv = h.live_of_var[in.set_vars[j] - 1];
                    if (lv < 0) continue;
                    for (uint32_t o = h.gocc_off[lv]; o < h.gocc_off[lv + 1]; o++) {
                        const uint32_t g = h.gocc[o];
                        if (gate_seen[g]) continue;
                        gate_seen[g] = 1;
                        const HostGate& hg = h.gates[g];
                        int outside = 0;
                        for (int i = 0; i < hg.m; i++) outside += live_in_set[hg.x[i]] ? 0 : 1;
                        if (hg.type == 1 && !live_in_set[hg.lb >> 1] && !live_in_set[hg.lc >> 1]) outside++;
                        if (outside <= 1) hp.gcand.push_back(g);
                    }
                }
                for (int j = 0; j < in.k; j++) {
                    const int32_t lv = h.live_of_var[in.set_vars[j] - 1];
                    if (lv < 0) continue;
                    live_in_set[lv] = 0;
                    for (uint32_t o = h.gocc_off[lv]; o < h.gocc_off[lv + 1]; o++) gate_seen[h.gocc[o]] = 0;
                }
                std::sort(hp.gcand.begin(), hp.gcand.end());
            }
        }
        hp.cand_off = (uint32_t)n_cand;
        n_cand += hp.cand.size() + hp.gcand.size();
        hp.mode = in.mode;
        hp.seed = in.seed;
        hp.first_sample = in.first_sample;
        hp.first_draw = in.mode == PDSAT_MODE_RANDOM_MT19937 ? in.first_draw : 0;
        hp.n_samples = in.n_samples;
        hp.n_groups = (in.n_samples + 63) / 64;
        hp.group_start = groups;
        hp.word_off = words;
        hp.kp = ((uint32_t)in.k + 1u) & ~1u;  // 16-byte tiles for the bulk copies
        if (hp.kp > kp_max) kp_max = hp.kp;
        groups += hp.n_groups;
        words += hp.n_groups * (uint64_t)hp.kp;
        if (in.mode == PDSAT_MODE_RANDOM_MT19937) {
            mt_plan(hp);
            hp.mt_index = (uint32_t)b->n_mt_points++;
        } else if (in.mode == PDSAT_MODE_EXPLICIT) {
            hp.stream_skip = 0;
            hp.stream_words = (in.n_samples * (uint64_t)in.k + 31) / 32;
            b->has_explicit = true;
        } else if (in.mode == PDSAT_MODE_ENUM_BLOCKS) {
            hp.block_bits = (uint32_t)in.block_bits;
            hp.block_cap = in.n_samples >> in.block_bits;
            hp.block_off = b->n_blocks_total;
            b->n_blocks_total += hp.block_cap;
        }
        n_codes += (size_t)in.k;
    }
    // mt19937 streams are cut into segments of 2^m rounds (m >= 1) generated by one CTA each
    if (b->n_mt_points) {
        uint64_t total_rounds = 0;
        for (auto& hp : b->pts)
            if (hp.mode == PDSAT_MODE_RANDOM_MT19937) total_rounds += hp.mt_rounds;
        if (db->device >= 0 && db->device < 64) {
            g_mt_batches[db->device]++;
            b->counted_mt = true;
        }
        const int target = mt_target_segments(db->device);
        const uint64_t per_seg = (total_rounds + target - 1) / target;
        int m = 1;
        while ((1ull << m) < per_seg) m++;
        // A batch alone on its device with a long stream (a neighbourhood of points): the bit
        // generator dominates the jumps, and more warps per scheduler pay.  Cost model from the
        // measurements in profiles/r02_plane_width.md: a warp alone on its scheduler needs 746
        // cycles per round (303 instructions at IPC 0.41); w warps reach IPC 0.58 / 0.70 / 0.78 /
        // 0.84 / 0.90 for w = 2 / 3 / 4 / 5-7 / >= 8 (the ALU pipe caps it at 0.91); a jump costs
        // 0.43 us of the whole GPU.  Pick the power of two that minimises generator + jumps.
        if (target == 592 && !getenv("PDSAT_MT_SEGMENTS")) {
            double best = 0;
            int best_m = m;
            for (int mm = 1; mm <= m; mm++) {
                uint64_t segs = 0;
                for (auto& hp : b->pts)
                    if (hp.mode == PDSAT_MODE_RANDOM_MT19937) segs += (hp.mt_rounds + (1ull << mm) - 1) >> mm;
                const uint64_t w = (segs + 591) / 592;
                const double ipc = w <= 1 ? 0.41 : w == 2 ? 0.58 : w == 3 ? 0.70 : w == 4 ? 0.78 : w < 8 ? 0.84 : 0.90;
                const double gen_us = (double)(1ull << mm) * 303.0 * (double)w / ipc / 1900.0;
                const double jump_us = 0.43 * (double)(segs - (uint64_t)b->n_mt_points);
                const double t = gen_us + jump_us;
                if (best == 0 || t < best) best = t, best_m = mm;
            }
            m = best_m;
        }
        b->mt_m = m;
        uint32_t extra = 2u * (uint32_t)b->n_mt_points;  // slots 2p, 2p+1: ping-pong pair of point p
        uint64_t seg = 0;
        for (auto& hp : b->pts) {
            if (hp.mode != PDSAT_MODE_RANDOM_MT19937) continue;
            hp.n_segs = (uint32_t)((hp.mt_rounds + (1ull << m) - 1) >> m);
            hp.extra_base = extra;
            extra += hp.n_segs - 1;
            hp.seg_base = seg;
            seg += hp.n_segs;
            hp.stream_words = (uint64_t)hp.n_segs * ((1ull << m) / 2) * 39;
        }
        b->n_mt_segs = (int)seg;
        b->n_mt_states = extra;
    }
    for (auto& hp : b->pts) {
        hp.stream_off = stream_words;
        stream_words += hp.stream_words;
    }
    b->n_groups = groups;
    b->n_words = words;
    // points in which BCP cannot start are streamed through in runs of several groups
    uint32_t stage_bytes = 16;
    for (auto& hp : b->pts) {
        const bool can_start = h.base_ok && (!hp.cand.empty() || !hp.gcand.empty() || (hp.flags & POINT_HAS_FIXED) ||
                                             (out_flags & PDSAT_OUT_ASSIGN) || hp.n_assumed_live == (uint32_t)h.n_live);
        if (!can_start) hp.flags |= POINT_STREAMING;
        hp.tile = can_start ? 1u : tile_groups_for(hp.kp);
        stage_bytes = std::max(stage_bytes, hp.tile * hp.kp * 8u);
    }
    (void)kp_max;
    b->shape = launch_shape(db, stage_bytes);
    if (b->shape.wpc < 1) {
        pdsat_batch_destroy(b);
        return fail(PDSAT_ERR_LIMIT, "assignment state + word tiles of one sample group exceed shared memory");
    }

    auto bail = [&](int code, const std::string& m) {
        pdsat_batch_destroy(b);
        return fail(code, m);
    };
#define CUB(call)                                                                        \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return bail(PDSAT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    CUB(cudaSetDevice(db->device));
    CUB(cudaMalloc(&b->d_points, sizeof(PointDev) * n_points));
    CUB(cudaMallocHost(&b->h_points, sizeof(PointDev) * n_points));
    CUB(cudaMalloc(&b->d_setcode, sizeof(uint32_t) * n_codes));
    CUB(cudaMalloc(&b->d_cand, sizeof(uint4) * (n_cand ? n_cand : 1)));
    CUB(cudaMalloc(&b->d_words, sizeof(uint64_t) * words));
    CUB(cudaMalloc(&b->d_stream, sizeof(uint32_t) * (stream_words ? stream_words : 1)));
    CUB(cudaMalloc(&b->d_cost, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMalloc(&b->d_acc, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMemset(b->d_acc, 0, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMalloc(&b->d_acc2, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMemset(b->d_acc2, 0, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMemset(b->d_cost, 0, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMalloc(&b->d_cost2, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMemset(b->d_cost2, 0, sizeof(unsigned long long) * PDSAT_COST_WORDS * n_points));
    CUB(cudaMalloc(&b->d_cta_done, 4 * sizeof(unsigned int)));  // [0] ticket, [1] peer time-out flag, [2] gather done
    CUB(cudaMemset(b->d_cta_done, 0, 4 * sizeof(unsigned int)));
    CUB(cudaMalloc(&b->d_model_count, sizeof(unsigned long long)));
    if (b->max_models > 0) {
        CUB(cudaMalloc(&b->d_model_bits, sizeof(uint32_t) * (size_t)b->max_models * b->model_words));
        CUB(cudaMalloc(&b->d_model_point, sizeof(uint32_t) * b->max_models));
        CUB(cudaMalloc(&b->d_model_sample, sizeof(uint64_t) * b->max_models));
    }
    if (out_flags & PDSAT_OUT_FLAGS) {
        CUB(cudaMalloc(&b->d_conflict, sizeof(uint64_t) * groups));
        CUB(cudaMalloc(&b->d_full, sizeof(uint64_t) * groups));
    }
    if (out_flags & PDSAT_OUT_TRAIL) CUB(cudaMalloc(&b->d_trail, sizeof(uint32_t) * 64 * groups));
    if (out_flags & PDSAT_OUT_ASSIGN)
        CUB(cudaMalloc(&b->d_assign, sizeof(uint64_t) * groups * (size_t)(h.n_live > 0 ? 2 * h.n_live : 1)));
    if (b->n_mt_segs) {
        CUB(cudaMalloc(&b->d_mt_states, sizeof(uint32_t) * 624 * MT_PARTS * b->n_mt_states));
        CUB(cudaMalloc(&b->d_mt_segs, sizeof(MtSegment) * b->n_mt_segs));
        CUB(cudaMallocHost(&b->h_mt_states, sizeof(uint32_t) * 624 * b->n_mt_points));
        CUB(cudaMallocHost(&b->h_mt_segs, sizeof(MtSegment) * b->n_mt_segs));
        b->mt_tasks_cap = (size_t)b->n_mt_segs + 64 * (size_t)b->n_mt_points;
        CUB(cudaMalloc(&b->d_mt_tasks, sizeof(MtJumpTask) * b->mt_tasks_cap));
        CUB(cudaMallocHost(&b->h_mt_tasks, sizeof(MtJumpTask) * b->mt_tasks_cap));
        CUB(cudaMalloc(&b->d_mt_srcs, sizeof(uint32_t) * b->mt_tasks_cap));
        CUB(cudaMallocHost(&b->h_mt_srcs, sizeof(uint32_t) * b->mt_tasks_cap));
        for (auto& hp : b->pts) {
            if (hp.mode != PDSAT_MODE_RANDOM_MT19937) continue;
            for (uint32_t j = 0; j < hp.n_segs; j++) {
                MtSegment& sg = b->h_mt_segs[hp.seg_base + j];
                sg.out_off = hp.stream_off + (uint64_t)j * ((1ull << b->mt_m) / 2) * 39;
                sg.n_rounds = 1ull << b->mt_m;
                sg.state_idx = j == 0 ? 2 * hp.mt_index : hp.extra_base + j - 1;  // j == 0 fixed up per fill
                sg.pad = 0;
            }
        }
        CUB(cudaMemcpy(b->d_mt_segs, b->h_mt_segs, sizeof(MtSegment) * b->n_mt_segs, cudaMemcpyHostToDevice));
        // scratch for the sequences of one level of jumps (at most one jump per segment pair / point)
        size_t max_jumps = (size_t)(b->n_mt_segs + 1) / 2;
        if ((size_t)b->n_mt_points > max_jumps) max_jumps = (size_t)b->n_mt_points;
        CUB(cudaMalloc(&b->d_mt_seq, sizeof(uint32_t) * MT_SEQ_WORDS * max_jumps));
        CUB(cudaFuncSetAttribute(mt19937_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SEQ_SMEM));
        b->mt_seq_cap = max_jumps;
    }
    if (b->n_blocks_total) {
        CUB(cudaMalloc(&b->d_blocks, sizeof(uint64_t) * b->n_blocks_total));
        CUB(cudaMallocHost(&b->h_blocks, sizeof(uint64_t) * b->n_blocks_total));
        for (int p = 0; p < n_points; p++)
            if (b->pts[p].mode == PDSAT_MODE_ENUM_BLOCKS)
                memcpy(b->h_blocks + b->pts[p].block_off, points[p].block_prefix, sizeof(uint64_t) * b->pts[p].block_cap);
    }
    if (b->has_explicit) {
        CUB(cudaMallocHost(&b->h_stream, sizeof(uint32_t) * stream_words));
        memset(b->h_stream, 0, sizeof(uint32_t) * stream_words);
        for (int p = 0; p < n_points; p++) {
            HostPoint& hp = b->pts[p];
            if (hp.mode != PDSAT_MODE_EXPLICIT) continue;
            uint32_t* dst = b->h_stream + hp.stream_off;
            const uint8_t* src = points[p].explicit_values;
            const uint64_t nb = hp.n_samples * (uint64_t)hp.set_vars.size();
            for (uint64_t d = 0; d < nb; d++)
                if (src[d]) dst[d >> 5] |= 1u << (d & 31);
        }
    }
    {
        std::vector<uint32_t> codes;
        codes.reserve(n_codes);
        uint32_t off = 0;
        for (int p = 0; p < n_points; p++) {
            HostPoint& hp = b->pts[p];
            PointDev& pd = b->h_points[p];
            pd.n_samples = hp.n_samples;
            pd.group_start = hp.group_start;
            pd.word_off = hp.word_off;
            pd.seed = hp.seed;
            pd.first_sample = hp.first_sample;
            pd.stream_off = hp.stream_off;
            pd.stream_skip = hp.stream_skip;
            pd.k = (uint32_t)hp.set_vars.size();
            pd.var_off = off;
            pd.mode = (uint32_t)hp.mode;
            pd.n_assumed_live = hp.n_assumed_live;
            pd.cand_off = hp.cand_off;
            pd.cand_cnt = (uint32_t)(hp.cand.size() + hp.gcand.size());
            pd.kp = hp.kp;
            pd.flags = hp.flags;
            pd.block_off = hp.block_off;
            pd.block_bits = hp.block_bits;
            pd.pad_ = 0;
            pd.tile = hp.tile;
            pd.pad2_ = 0;
            off += pd.k;
            codes.insert(codes.end(), hp.setcode.begin(), hp.setcode.end());
            if (pd.cand_cnt) {  // gates first, then clauses: neighbouring lanes take the same path
                std::vector<uint16_t> recs((size_t)pd.cand_cnt * REC_SLOTS);
                size_t i = 0;
                for (uint32_t g : hp.gcand) memcpy(&recs[i++ * REC_SLOTS], &h.gate_rec[(size_t)g * REC_SLOTS], sizeof(uint16_t) * REC_SLOTS);
                for (uint32_t r : hp.cand) h.inline_record(r, &recs[i++ * REC_SLOTS]);
                CUB(cudaMemcpy(b->d_cand + hp.cand_off, recs.data(), sizeof(uint4) * pd.cand_cnt, cudaMemcpyHostToDevice));
            }
        }
        CUB(cudaMemcpy(b->d_setcode, codes.data(), sizeof(uint32_t) * n_codes, cudaMemcpyHostToDevice));
        CUB(cudaMemcpy(b->d_points, b->h_points, sizeof(PointDev) * n_points, cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 4; i++) CUB(cudaEventCreate(&b->ev[i]));
    CUB(cudaEventCreateWithFlags(&b->ev_h2d, cudaEventDisableTiming));
#undef CUB
    *out = b;
    return PDSAT_OK;
}

// the pinned staging buffers (h_points, h_mt_*) may still be read by the previous fill's
// asynchronous copies: wait for them before the host rewrites the buffers
static void staging_wait(pdsat_batch* b) {
    if (b->h2d_pending) {
        cudaEventSynchronize(b->ev_h2d);
        b->h2d_pending = false;
    }
}

int pdsat_batch_reseed(pdsat_batch* b, int32_t point, uint64_t seed, uint64_t first_sample) {
    if (!b || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_reseed: bad argument");
    HostPoint& hp = b->pts[point];
    if (hp.mode == PDSAT_MODE_EXPLICIT) return fail(PDSAT_ERR_ARG, "pdsat_batch_reseed: explicit point");
    staging_wait(b);
    if (hp.mode == PDSAT_MODE_RANDOM_COUNTER && (first_sample & 63))
        return fail(PDSAT_ERR_ARG, "pdsat_batch_reseed: first_sample must be a multiple of 64");
    hp.seed = seed;
    if (hp.mode == PDSAT_MODE_RANDOM_MT19937) {
        hp.first_sample = first_sample;
        mt_plan(hp);  // the allocation already covers the worst-case skip
    } else {
        hp.first_sample = first_sample;
    }
    PointDev& pd = b->h_points[point];
    pd.seed = hp.seed;
    pd.first_sample = hp.first_sample;
    pd.stream_skip = hp.stream_skip;
    b->filled = false;
    return PDSAT_OK;
}

int pdsat_batch_set_first_draw(pdsat_batch* b, int32_t point, uint64_t first_draw) {
    if (!b || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_set_first_draw: bad argument");
    HostPoint& hp = b->pts[point];
    if (hp.mode != PDSAT_MODE_RANDOM_MT19937) return fail(PDSAT_ERR_ARG, "pdsat_batch_set_first_draw: not a mt19937 point");
    staging_wait(b);
    hp.first_draw = first_draw;
    mt_plan(hp);
    b->h_points[point].stream_skip = hp.stream_skip;
    b->filled = false;
    return PDSAT_OK;
}

static int stale(const pdsat_batch* b) {
    if (b->generation != b->db->generation)
        return fail(PDSAT_ERR_ARG, "stale batch: pdsat_set_fixed_assumptions rebuilt the database after this batch was created");
    return PDSAT_OK;
}

int pdsat_batch_fill(pdsat_batch* b) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_fill: null batch");
    if (stale(b)) return PDSAT_ERR_ARG;
    pdsat_db* db = b->db;
    CU(cudaSetDevice(db->device));
    staging_wait(b);  // the previous fill may still be reading the pinned staging buffers
    cudaStream_t st = db->stream;
    if (b->timing) CU(cudaEventRecord(b->ev[0], st));
    {
        uint64_t runs = 0;  // first run of every point (a run = tile_groups consecutive groups, one bulk copy)
        for (size_t p = 0; p < b->pts.size(); p++) {
            b->h_points[p].run_start = runs;
            runs += (b->pts[p].n_groups + b->pts[p].tile - 1) / b->pts[p].tile;
        }
    }
    CU(cudaMemcpyAsync(b->d_points, b->h_points, sizeof(PointDev) * b->pts.size(), cudaMemcpyHostToDevice, st));
    int64_t h2d = (int64_t)(sizeof(PointDev) * b->pts.size());
    if (b->n_mt_segs) {
        // seed states -> part 0 of slot 2p (other parts zero); jump-ahead (mt_jump.hpp) places
        // every other segment
        for (auto& hp : b->pts) {
            if (hp.mode != PDSAT_MODE_RANDOM_MT19937) continue;
            mt19937_seed((uint32_t)hp.seed, b->h_mt_states + (size_t)hp.mt_index * 624);
        }
        CU(cudaMemsetAsync(b->d_mt_states, 0, sizeof(uint32_t) * 624 * MT_PARTS * 2 * b->n_mt_points, st));
        CU(cudaMemcpy2DAsync(b->d_mt_states, sizeof(uint32_t) * 624 * MT_PARTS * 2, b->h_mt_states, sizeof(uint32_t) * 624,
                             sizeof(uint32_t) * 624, b->n_mt_points, cudaMemcpyHostToDevice, st));
        h2d += (int64_t)(sizeof(uint32_t) * 624 * b->n_mt_points + sizeof(MtSegment) * b->n_mt_segs);
        // launch plan: (a) the jump to the first round of every point's window: per set bit t
        // of the round number a jump by 2^t rounds between the point's two slots - or ONE jump
        // with the composite polynomial (built on the host once, ~0.1 s) for an offset seen for
        // the second time; (b) a radix-8 tree over the segments, top level first: the up to 7
        // jumps that leave a segment share its sequence
        struct Launch { size_t off, cnt, src_off, src_cnt; };
        std::vector<Launch> launches;
        size_t nt = 0, ns = 0;
        std::vector<uint32_t> cur_slot(b->pts.size());
        for (size_t p = 0; p < b->pts.size(); p++) cur_slot[p] = 2 * b->pts[p].mt_index;
        bool table_full = false;
        auto own_seq_task = [&](size_t level_src_off, uint32_t src, uint32_t dst, int poly) {
            b->h_mt_srcs[ns] = src;
            b->h_mt_tasks[nt++] = MtJumpTask{(uint32_t)(ns - level_src_off), dst, (uint32_t)poly, 0};
            ns++;
        };
        std::vector<char> use_comp(b->pts.size(), 0);
        {
            const size_t start = nt, sstart = ns;
            for (size_t p = 0; p < b->pts.size(); p++) {
                HostPoint& hp = b->pts[p];
                const uint64_t r0 = hp.mt_first_round;
                if (hp.mode != PDSAT_MODE_RANDOM_MT19937 || __builtin_popcountll(r0) < 2) continue;
                if (!db->poly_slots.count(r0) && (++db->comp_seen[r0] < 2 || (int)db->poly_slots.size() >= POLY_CAP_COMPOSITE))
                    continue;
                const int slot = mt_poly_slot(db, r0, st);
                if (slot < 0) continue;
                own_seq_task(sstart, cur_slot[p], cur_slot[p] ^ 1u, slot);
                cur_slot[p] ^= 1u;
                use_comp[p] = 1;
            }
            if (nt > start) launches.push_back({start, nt - start, sstart, ns - sstart});
        }
        for (int t = 0; t < 64; t++) {
            const size_t start = nt, sstart = ns;
            for (size_t p = 0; p < b->pts.size(); p++) {
                HostPoint& hp = b->pts[p];
                if (hp.mode != PDSAT_MODE_RANDOM_MT19937 || use_comp[p] || !((hp.mt_first_round >> t) & 1ull)) continue;
                const int slot = mt_poly_slot(db, 1ull << t, st);
                if (slot < 0) { table_full = true; continue; }
                own_seq_task(sstart, cur_slot[p], cur_slot[p] ^ 1u, slot);
                cur_slot[p] ^= 1u;
            }
            if (nt > start) launches.push_back({start, nt - start, sstart, ns - sstart});
        }
        for (size_t p = 0; p < b->pts.size(); p++)
            if (b->pts[p].mode == PDSAT_MODE_RANDOM_MT19937) b->h_mt_segs[b->pts[p].seg_base].state_idx = cur_slot[p];
        CU(cudaMemcpyAsync(b->d_mt_segs, b->h_mt_segs, sizeof(MtSegment) * b->n_mt_segs, cudaMemcpyHostToDevice, st));
        uint32_t max_segs = 0;
        for (auto& hp : b->pts)
            if (hp.mode == PDSAT_MODE_RANDOM_MT19937 && hp.n_segs > max_segs) max_segs = hp.n_segs;
        int top = 0;
        while ((8ull << (3 * top)) < max_segs) top++;
        for (int lev = top; lev >= 0 && max_segs > 1; lev--) {
            const size_t start = nt, sstart = ns;
            const uint64_t stride = 1ull << (3 * lev);
            if (b->mt_m + 3 * lev > 58) return fail(PDSAT_ERR_LIMIT, "mt19937 stream too long");
            for (size_t p = 0; p < b->pts.size(); p++) {
                HostPoint& hp = b->pts[p];
                if (hp.mode != PDSAT_MODE_RANDOM_MT19937) continue;
                for (uint64_t j = 0; j + stride < hp.n_segs; j += 8 * stride) {
                    b->h_mt_srcs[ns] = j == 0 ? cur_slot[p] : hp.extra_base + (uint32_t)j - 1;
                    for (uint64_t d = 1; d < 8 && j + d * stride < hp.n_segs; d++) {
                        const int slot = mt_poly_slot(db, (d * stride) << b->mt_m, st);
                        if (slot < 0) { table_full = true; continue; }
                        b->h_mt_tasks[nt++] = MtJumpTask{(uint32_t)(ns - sstart), hp.extra_base + (uint32_t)(j + d * stride) - 1,
                                                         (uint32_t)slot, 0};
                    }
                    ns++;
                }
            }
            if (nt > start) launches.push_back({start, nt - start, sstart, ns - sstart});
        }
        if (table_full) return fail(PDSAT_ERR_LIMIT, "mt19937 jump polynomial table full");
        if (!launches.empty()) {
            CU(cudaMemcpyAsync(b->d_mt_tasks, b->h_mt_tasks, sizeof(MtJumpTask) * nt, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(b->d_mt_srcs, b->h_mt_srcs, sizeof(uint32_t) * ns, cudaMemcpyHostToDevice, st));
            h2d += (int64_t)(sizeof(MtJumpTask) * nt + sizeof(uint32_t) * ns);
            for (auto& L : launches) {
                if (L.src_cnt > b->mt_seq_cap) return fail(PDSAT_ERR_LIMIT, "mt19937 jump scratch too small");
                mt19937_launch_jumps(db->d_polys, b->d_mt_states, b->d_mt_srcs + L.src_off, L.src_cnt,
                                     b->d_mt_tasks + L.off, L.cnt, b->d_mt_seq, st);
                b->n_kernels += 2;
            }
        }
        mt19937_bits_kernel<<<(b->n_mt_segs + MT_BITS_WARPS - 1) / MT_BITS_WARPS, MT_BITS_WARPS * 32, 0, st>>>(
            b->d_mt_states, b->d_mt_segs, b->n_mt_segs, 1u << b->mt_m, b->d_stream);
        b->n_kernels++;
    }
    if (b->has_explicit) {
        for (auto& hp : b->pts) {
            if (hp.mode != PDSAT_MODE_EXPLICIT) continue;
            CU(cudaMemcpyAsync(b->d_stream + hp.stream_off, b->h_stream + hp.stream_off,
                               sizeof(uint32_t) * hp.stream_words, cudaMemcpyHostToDevice, st));
            h2d += (int64_t)(sizeof(uint32_t) * hp.stream_words);
        }
    }
    if (b->n_blocks_total) {
        CU(cudaMemcpyAsync(b->d_blocks, b->h_blocks, sizeof(uint64_t) * b->n_blocks_total, cudaMemcpyHostToDevice, st));
        h2d += (int64_t)(sizeof(uint64_t) * b->n_blocks_total);
    }
    CU(cudaEventRecord(b->ev_h2d, st));
    b->h2d_pending = true;
    BatchDev bd{};
    bd.points = b->d_points;
    bd.n_points = (int32_t)b->pts.size();
    bd.n_groups = b->n_groups;
    bd.n_words = b->n_words;
    bd.words = b->d_words;
    bd.stream = b->d_stream;
    bd.block_prefix = b->d_blocks;
    bool any_stream = false, any_other = false;
    for (auto& hp : b->pts) {
        if (hp.mode == PDSAT_MODE_RANDOM_MT19937 || hp.mode == PDSAT_MODE_EXPLICIT) any_stream = true;
        else any_other = true;
    }
    if (any_other) {  // one thread per word (the words of stream points are left to fill_stream_kernel)
        const int threads = 256;
        const uint64_t blocks = (b->n_words + threads - 1) / threads;
        fill_words_kernel<<<(unsigned)blocks, threads, 0, st>>>(bd);
        b->n_kernels++;
    }
    if (any_stream) {  // one warp per sample group
        const uint64_t blocks = (b->n_groups + FILL_STREAM_WARPS - 1) / FILL_STREAM_WARPS;
        fill_stream_kernel<<<(unsigned)blocks, FILL_STREAM_WARPS * 32, 0, st>>>(bd);
        b->n_kernels++;
    }
    CU(cudaGetLastError());
    if (b->timing) CU(cudaEventRecord(b->ev[1], st));
    b->fill_timed = b->timing;
    b->filled = true;
    b->h2d_bytes_last_fill = h2d;
    return PDSAT_OK;
}

static void peer_fill(pdsat_batch* b, BatchDev& bd);

int pdsat_batch_propagate(pdsat_batch* b) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_propagate: null batch");
    if (!b->filled) return fail(PDSAT_ERR_ARG, "pdsat_batch_propagate: batch not filled");
    if (stale(b)) return PDSAT_ERR_ARG;
    pdsat_db* db = b->db;
    CU(cudaSetDevice(db->device));
    cudaStream_t st = db->stream;
    const bool direct = b->peer_ranks == 0 && !b->d_cost_ext && !b->no_direct;
    const size_t cost_bytes = sizeof(unsigned long long) * PDSAT_COST_WORDS * b->pts.size();
    const int path = direct ? 0 : (b->peer_ranks > 0 ? 2 : 1);  // direct / ticket / peers
    if (path != b->last_path) {  // path change (rare): every convention starts from zeroed buffers
        CU(cudaMemsetAsync(b->d_cost, 0, cost_bytes, st));
        CU(cudaMemsetAsync(b->d_cost2, 0, cost_bytes, st));
        CU(cudaMemsetAsync(b->d_acc, 0, cost_bytes, st));
        CU(cudaMemsetAsync(b->d_acc2, 0, cost_bytes, st));
        b->last_path = path;
    }
    if (direct) b->cost_cur ^= 1;
    b->last_direct = direct;
    unsigned long long* cost = b->cost_ptr();
    if (b->timing) CU(cudaEventRecord(b->ev[2], st));
    // no memset of the accumulators: the kernel's last CTA publishes the scratch sums and clears them
    if (b->max_models > 0) CU(cudaMemsetAsync(b->d_model_count, 0, sizeof(unsigned long long), st));
    BatchDev bd{};
    bd.points = b->d_points;
    bd.n_points = (int32_t)b->pts.size();
    bd.n_groups = b->n_groups;
    bd.n_words = b->n_words;
    bd.setcode = b->d_setcode;
    bd.cand = b->d_cand;
    bd.words = b->d_words;
    bd.stream = b->d_stream;
    bd.cost = cost;
    const int peer_par = (int)(b->peer_step & 1);
    bd.acc = direct ? cost : (b->peer_ranks > 0 && peer_par ? b->d_acc2 : b->d_acc);
    bd.zero_next = direct ? (b->cost_cur ? b->d_cost : b->d_cost2) : nullptr;
    bd.no_ticket = (direct || b->peer_ranks > 0) ? 1 : 0;
    bd.cta_done = b->d_cta_done;
    bd.p0 = b->h_points[0];
    bd.conflict_bits = b->d_conflict;
    bd.full_bits = b->d_full;
    bd.trail = b->d_trail;
    bd.assign = b->d_assign;
    bd.model_count = b->d_model_count;
    bd.model_bits = b->d_model_bits;
    bd.model_point = b->d_model_point;
    bd.model_sample = b->d_model_sample;
    bd.max_models = b->max_models;
    bd.model_words = b->model_words;
    const LaunchShape& L = b->shape;
    bd.ring_stages = L.stages;
    bd.ring_stage_bytes = L.stage_bytes;
    bd.tile_groups = 0;  // per point (PointDev::tile)
    {   // runs: tile_groups consecutive groups of a point (run_start of every point went up with the fill)
        uint64_t runs = 0;
        for (size_t p = 0; p < b->pts.size(); p++) runs += (b->pts[p].n_groups + b->pts[p].tile - 1) / b->pts[p].tile;
        bd.n_runs = runs;
    }
    DbDev dev = db->dev;
    dev.blob_in_smem = L.blob_in_smem;
    uint64_t ctas_needed = (bd.n_runs + L.wpc - 1) / L.wpc;
    int grid = (int)(ctas_needed < (uint64_t)db->grid ? ctas_needed : (uint64_t)db->grid);
    if (grid < 1) grid = 1;
    const int n_cost_words = PDSAT_COST_WORDS * (int)b->pts.size();
    if (b->peer_ranks > 0) peer_fill(b, bd);  // what CTA 0 catches up on first (PeerDev)
    {
        // programmatic dependent launch (kernels.cuh: griddepcontrol): the prologue of this launch
        // overlaps the tail of the kernel before it in the stream
        static const bool use_pdl = getenv("PDSAT_NO_PDL") == nullptr;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)(L.wpc * 32));
        cfg.dynamicSmemBytes = (size_t)L.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = use_pdl ? 1 : 0;
        const int wb = L.warp_bytes;
        if (L.plane_bits == 64) CU(cudaLaunchKernelEx(&cfg, bcp_kernel, dev, bd, wb));
        else CU(cudaLaunchKernelEx(&cfg, bcp_kernel_h, dev, bd, wb));
    }
    b->n_kernels++;
    if (b->peer_ranks > 0) {  // the sums of this step stay in the scratch until the next launch / peer_finalize
        b->peer_ungathered = b->peer_unsent;  // the launch pushes the step before (if it was still unsent)
        b->peer_unsent = true;
        b->peer_step++;
    }
    CU(cudaGetLastError());
    if (b->timing) CU(cudaEventRecord(b->ev[3], st));
    b->propagate_timed = b->timing;
    b->propagated = true;
    return PDSAT_OK;
}

// peers: the PeerDev fields of a launch (or of peer_finish_kernel) that follows `peer_step` steps.
// Step numbers: the last launched step is t = peer_step - 1; tag = step + 1 (never 0).
static void peer_fill(pdsat_batch* b, BatchDev& bd) {
    for (int r = 0; r < b->peer_ranks; r++) bd.peer.shared[r] = b->peer_ptr[r];
    bd.peer.n_ranks = b->peer_ranks;
    bd.peer.me = b->peer_me;
    bd.peer.n_words = PDSAT_COST_WORDS * (int)b->pts.size();
    bd.peer.timed_out = b->d_cta_done + 1;
    const uint64_t t = b->peer_step - 1;  // only read when something is pending (peer_step >= 1 then)
    bd.peer.push_prev = b->peer_unsent ? 1 : 0;
    bd.peer.parity = (int)(t & 1);
    bd.peer.tag = (uint32_t)(t + 1);
    bd.peer.acc_prev = (t & 1) ? b->d_acc2 : b->d_acc;
    bd.peer.gather_prev = b->peer_ungathered ? 1 : 0;
    bd.peer.g_parity = (int)((t - 1) & 1);
    bd.peer.g_tag = (uint32_t)t;
}

// peers: the host wants the totals of the last launch now
static int peer_finalize(pdsat_batch* b) {
    if (b->peer_ranks == 0 || !b->peer_unsent) return PDSAT_OK;
    CU(cudaSetDevice(b->db->device));
    BatchDev bd{};
    bd.n_points = (int32_t)b->pts.size();
    bd.cost = b->cost_ptr();
    peer_fill(b, bd);
    peer_finish_kernel<<<1, 512, 0, b->db->stream>>>(bd);
    b->n_kernels++;
    CU(cudaGetLastError());
    b->peer_unsent = b->peer_ungathered = false;
    return PDSAT_OK;
}

int pdsat_batch_sync(pdsat_batch* b) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_sync: null batch");
    {
        const int rc = peer_finalize(b);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(b->db->stream));
    return PDSAT_OK;
}

int pdsat_batch_last_ms(pdsat_batch* b, float* fill_ms, float* propagate_ms) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_last_ms: null batch");
    CU(cudaStreamSynchronize(b->db->stream));
    if (fill_ms) {
        *fill_ms = 0;
        if (b->filled && b->fill_timed) CU(cudaEventElapsedTime(fill_ms, b->ev[0], b->ev[1]));
    }
    if (propagate_ms) {
        *propagate_ms = 0;
        if (b->propagated && b->propagate_timed) CU(cudaEventElapsedTime(propagate_ms, b->ev[2], b->ev[3]));
    }
    return PDSAT_OK;
}

int pdsat_batch_set_timing(pdsat_batch* b, int32_t on) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_set_timing: null batch");
    b->timing = on != 0;
    return PDSAT_OK;
}

int pdsat_batch_launch_count(pdsat_batch* b, int64_t* n_kernels) {
    if (!b || !n_kernels) return fail(PDSAT_ERR_ARG, "pdsat_batch_launch_count: bad argument");
    *n_kernels = b->n_kernels;
    return PDSAT_OK;
}

int pdsat_batch_h2d_bytes(pdsat_batch* b, int64_t* bytes) {
    if (!b || !bytes) return fail(PDSAT_ERR_ARG, "pdsat_batch_h2d_bytes: bad argument");
    *bytes = b->h2d_bytes_last_fill;
    return PDSAT_OK;
}

int pdsat_batch_cost_device_ptr(pdsat_batch* b, void** dptr) {
    if (!b || !dptr) return fail(PDSAT_ERR_ARG, "pdsat_batch_cost_device_ptr: bad argument");
    b->no_direct = true;  // the pointer stays valid for every later launch
    *dptr = b->d_cost_ext ? b->d_cost_ext : b->d_cost;
    return PDSAT_OK;
}

int pdsat_batch_set_cost_buffer(pdsat_batch* b, void* dptr) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_set_cost_buffer: null batch");
    {  // sums still to be gathered go where the caller expected them when it launched
        const int rc = peer_finalize(b);
        if (rc) return rc;
    }
    b->d_cost_ext = (unsigned long long*)dptr;
    return PDSAT_OK;
}

int pdsat_batch_costs(pdsat_batch* b, pdsat_cost* out) {
    if (!b || !out) return fail(PDSAT_ERR_ARG, "pdsat_batch_costs: bad argument");
    if (!b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_costs: batch not propagated");
    static_assert(sizeof(pdsat_cost) == 8 * PDSAT_COST_WORDS, "pdsat_cost layout");
    {
        const int rc = peer_finalize(b);
        if (rc) return rc;
    }
    unsigned long long* cost = b->cost_ptr();
    CU(cudaMemcpyAsync(out, cost, sizeof(pdsat_cost) * b->pts.size(), cudaMemcpyDeviceToHost, b->db->stream));
    unsigned int timed_out = 0;
    if (b->peer_ranks > 0)
        CU(cudaMemcpyAsync(&timed_out, b->d_cta_done + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, b->db->stream));
    CU(cudaStreamSynchronize(b->db->stream));
    if (timed_out) return fail(PDSAT_ERR_CUDA, "peer exchange timed out: a rank did not launch its step");
    return PDSAT_OK;
}

int pdsat_batch_flags(pdsat_batch* b, int32_t point, uint64_t* conflict_bits, uint64_t* full_bits) {
    if (!b || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_flags: bad argument");
    if (!b->d_conflict || !b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_flags: PDSAT_OUT_FLAGS not requested / not run");
    const HostPoint& hp = b->pts[point];
    CU(cudaStreamSynchronize(b->db->stream));
    if (conflict_bits)
        CU(cudaMemcpy(conflict_bits, b->d_conflict + hp.group_start, sizeof(uint64_t) * hp.n_groups, cudaMemcpyDeviceToHost));
    if (full_bits)
        CU(cudaMemcpy(full_bits, b->d_full + hp.group_start, sizeof(uint64_t) * hp.n_groups, cudaMemcpyDeviceToHost));
    return PDSAT_OK;
}

int pdsat_batch_trail(pdsat_batch* b, int32_t point, uint32_t* trail) {
    if (!b || !trail || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_trail: bad argument");
    if (!b->d_trail || !b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_trail: PDSAT_OUT_TRAIL not requested / not run");
    const HostPoint& hp = b->pts[point];
    CU(cudaStreamSynchronize(b->db->stream));
    CU(cudaMemcpy(trail, b->d_trail + hp.group_start * 64, sizeof(uint32_t) * hp.n_samples, cudaMemcpyDeviceToHost));
    return PDSAT_OK;
}

int pdsat_batch_assign(pdsat_batch* b, int32_t point, uint64_t first, uint64_t n, uint8_t* assigns) {
    if (!b || !assigns || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_assign: bad argument");
    if (!b->d_assign || !b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_assign: PDSAT_OUT_ASSIGN not requested / not run");
    const HostPoint& hp = b->pts[point];
    if (first + n > hp.n_samples) return fail(PDSAT_ERR_ARG, "pdsat_batch_assign: range");
    const HostDb& h = b->db->host;
    const size_t n_lit = (size_t)2 * h.n_live;
    CU(cudaStreamSynchronize(b->db->stream));
    std::vector<uint64_t> planes(n_lit ? n_lit : 1);
    uint64_t cur_group = ~0ull;
    for (uint64_t s = first; s < first + n; s++) {
        uint64_t g = s / 64;
        if (g != cur_group) {
            if (n_lit)
                CU(cudaMemcpy(planes.data(), b->d_assign + (hp.group_start + g) * n_lit, sizeof(uint64_t) * n_lit,
                              cudaMemcpyDeviceToHost));
            cur_group = g;
        }
        uint8_t* dst = assigns + (s - first) * (size_t)h.n_vars;
        memcpy(dst, h.base_assign.data(), (size_t)h.n_vars);
        const int bit = (int)(s & 63);
        for (int lv = 0; lv < h.n_live; lv++) {
            const bool t = (planes[2 * lv] >> bit) & 1, f = (planes[2 * lv + 1] >> bit) & 1;
            dst[h.var_of_live[lv]] = t ? LB_TRUE : (f ? LB_FALSE : LB_UNDEF);
        }
    }
    return PDSAT_OK;
}

int pdsat_batch_models(pdsat_batch* b, int64_t* n_found, int32_t* point_of, uint64_t* sample_of, uint8_t* models) {
    if (!b || !n_found) return fail(PDSAT_ERR_ARG, "pdsat_batch_models: bad argument");
    if (!b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_models: batch not propagated");
    CU(cudaStreamSynchronize(b->db->stream));
    unsigned long long cnt = 0;
    if (b->max_models > 0) {
        CU(cudaMemcpy(&cnt, b->d_model_count, sizeof(cnt), cudaMemcpyDeviceToHost));
    } else {  // nothing is stored: the count is the sum of the points' n_full
        {
            const int rc = peer_finalize(b);
            if (rc) return rc;
        }
        std::vector<pdsat_cost> c(b->pts.size());
        unsigned long long* cost = b->cost_ptr();
        CU(cudaMemcpy(c.data(), cost, sizeof(pdsat_cost) * c.size(), cudaMemcpyDeviceToHost));
        for (auto& x : c) cnt += x.n_full;
    }
    *n_found = (int64_t)cnt;
    const int64_t stored = (int64_t)cnt < b->max_models ? (int64_t)cnt : b->max_models;
    if (stored == 0) return PDSAT_OK;
    const HostDb& h = b->db->host;
    std::vector<uint32_t> bits((size_t)stored * b->model_words), pts((size_t)stored);
    std::vector<uint64_t> smp((size_t)stored);
    CU(cudaMemcpy(bits.data(), b->d_model_bits, sizeof(uint32_t) * bits.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pts.data(), b->d_model_point, sizeof(uint32_t) * stored, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(smp.data(), b->d_model_sample, sizeof(uint64_t) * stored, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < stored; i++) {
        if (point_of) point_of[i] = (int32_t)pts[i];
        if (sample_of) sample_of[i] = smp[i];
        if (models) {
            uint8_t* dst = models + (size_t)i * h.n_vars;
            for (int v = 0; v < h.n_vars; v++) dst[v] = h.base_assign[v] == LB_TRUE ? 1 : 0;
            for (int lv = 0; lv < h.n_live; lv++)
                dst[h.var_of_live[lv]] = (bits[(size_t)i * b->model_words + (lv >> 5)] >> (lv & 31)) & 1u;
        }
    }
    return PDSAT_OK;
}

int pdsat_batch_peer_export(pdsat_batch* b, void* handle_out) {
    if (!b || !handle_out) return fail(PDSAT_ERR_ARG, "pdsat_batch_peer_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == PDSAT_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(b->db->device));
    // receive buffer: [parity][source rank][2 packets per accumulator] (kernels.cuh PeerDev)
    const size_t words = 2 * (size_t)MAX_PEERS * 2 * (size_t)PDSAT_COST_WORDS * b->pts.size();
    if (!b->d_peer_own) {
        CU(cudaMalloc(&b->d_peer_own, sizeof(unsigned long long) * words));
        CU(cudaMemset(b->d_peer_own, 0, sizeof(unsigned long long) * words));
    }
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, b->d_peer_own));
    memcpy(handle_out, &hd, sizeof(hd));
    return PDSAT_OK;
}

int pdsat_batch_peer_connect(pdsat_batch* b, int32_t n_ranks, int32_t my_rank, const void* handles) {
    if (!b || !handles || n_ranks < 1 || n_ranks > MAX_PEERS || my_rank < 0 || my_rank >= n_ranks)
        return fail(PDSAT_ERR_ARG, "pdsat_batch_peer_connect: bad argument");
    if (!b->d_peer_own) return fail(PDSAT_ERR_ARG, "pdsat_batch_peer_connect: call pdsat_batch_peer_export first");
    CU(cudaSetDevice(b->db->device));
    for (int r = 0; r < n_ranks; r++) {
        if (r == my_rank) {
            b->peer_ptr[r] = b->d_peer_own;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles + (size_t)r * PDSAT_IPC_HANDLE_BYTES, sizeof(hd));
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        b->peer_ptr[r] = (unsigned long long*)p;
    }
    // the step tags restart at 1: packets of an earlier connection must not match them.  The
    // caller synchronises the ranks (a barrier) between connect and the first propagate.
    const size_t words = 2 * (size_t)MAX_PEERS * 2 * (size_t)PDSAT_COST_WORDS * b->pts.size();
    CU(cudaMemset(b->d_peer_own, 0, sizeof(unsigned long long) * words));
    b->peer_ranks = n_ranks;
    b->peer_me = my_rank;
    b->peer_step = 0;
    b->peer_unsent = b->peer_ungathered = false;
    CU(cudaMemset(b->d_cta_done, 0, 4 * sizeof(unsigned int)));
    return PDSAT_OK;
}

int pdsat_batch_peer_disconnect(pdsat_batch* b) {
    if (!b) return fail(PDSAT_ERR_ARG, "pdsat_batch_peer_disconnect: null batch");
    if (b->db && b->db->device >= 0 && b->peer_ranks > 0) peer_finalize(b);  // nothing of this rank stays unsent
    b->peer_unsent = b->peer_ungathered = false;
    if (b->db && b->db->device >= 0) {
        cudaSetDevice(b->db->device);
        if (b->db->stream) cudaStreamSynchronize(b->db->stream);
    }
    for (int r = 0; r < b->peer_ranks; r++)
        if (r != b->peer_me && b->peer_ptr[r]) cudaIpcCloseMemHandle(b->peer_ptr[r]);
    for (int r = 0; r < MAX_PEERS; r++) b->peer_ptr[r] = nullptr;
    b->peer_ranks = 0;
    return PDSAT_OK;
}

void pdsat_batch_destroy(pdsat_batch* b) {
    if (!b) return;
    if (b->counted_mt) g_mt_batches[b->db->device]--;
    if (b->peer_ranks) pdsat_batch_peer_disconnect(b);
    cudaFree(b->d_peer_own);
    cudaFree(b->d_cta_done);
    cudaFree(b->d_mt_seq);
    if (b->db && b->db->device >= 0) {
        cudaSetDevice(b->db->device);
        if (b->db->stream) cudaStreamSynchronize(b->db->stream);
    }
    cudaFree(b->d_points);
    cudaFree(b->d_setcode);
    cudaFree(b->d_cand);
    cudaFree(b->d_words);
    cudaFree(b->d_stream);
    cudaFree(b->d_cost);
    cudaFree(b->d_cost2);
    cudaFree(b->d_acc);
    cudaFree(b->d_acc2);
    cudaFree(b->d_conflict);
    cudaFree(b->d_full);
    cudaFree(b->d_trail);
    cudaFree(b->d_assign);
    cudaFree(b->d_model_count);
    cudaFree(b->d_model_bits);
    cudaFree(b->d_model_point);
    cudaFree(b->d_model_sample);
    cudaFree(b->d_mt_states);
    cudaFree(b->d_mt_segs);
    cudaFree(b->d_mt_tasks);
    cudaFree(b->d_mt_srcs);
    if (b->h_mt_srcs) cudaFreeHost(b->h_mt_srcs);
    cudaFree(b->d_blocks);
    if (b->h_blocks) cudaFreeHost(b->h_blocks);
    if (b->h_mt_tasks) cudaFreeHost(b->h_mt_tasks);
    if (b->h_mt_states) cudaFreeHost(b->h_mt_states);
    if (b->h_mt_segs) cudaFreeHost(b->h_mt_segs);
    if (b->h_stream) cudaFreeHost(b->h_stream);
    if (b->h_points) cudaFreeHost(b->h_points);
    for (int i = 0; i < 4; i++)
        if (b->ev[i]) cudaEventDestroy(b->ev[i]);
    if (b->ev_h2d) cudaEventDestroy(b->ev_h2d);
    pdsat_db* db = b->db;
    delete b;
    if (db && --db->live_batches == 0 && db->zombie) {
        db->zombie = false;
        pdsat_destroy(db);
    }
}

int pdsat_batch_set_blocks(pdsat_batch* b, uint64_t n_blocks, const uint64_t* block_prefix) {
    if (!b || !block_prefix || n_blocks == 0 || b->pts.size() != 1 || b->pts[0].mode != PDSAT_MODE_ENUM_BLOCKS ||
        n_blocks > b->pts[0].block_cap)
        return fail(PDSAT_ERR_ARG, "pdsat_batch_set_blocks: needs a single-point ENUM_BLOCKS batch and n_blocks within its capacity");
    staging_wait(b);
    HostPoint& hp = b->pts[0];
    memcpy(b->h_blocks, block_prefix, sizeof(uint64_t) * n_blocks);
    hp.n_samples = n_blocks << hp.block_bits;
    hp.n_groups = hp.n_samples / 64;
    b->h_points[0].n_samples = hp.n_samples;
    b->n_groups = hp.n_groups;
    b->n_words = hp.n_groups * (uint64_t)hp.kp;
    b->filled = false;
    return PDSAT_OK;
}

// ---- BCP-pruned enumeration (see pdsat_cuda.h).  Depth-first over chunks of surviving prefixes;
// one reusable single-point ENUM_BLOCKS batch per level.
namespace {
struct EnumLevel {
    int depth_before = 0;  // bits of lint fixed by the prefixes fed to this level
    int w = 0;             // bits this level enumerates
    uint64_t cap_blocks = 0;
    pdsat_batch* b = nullptr;
    std::vector<uint64_t> conflict, full;  // flag words of the last run
};
}  // namespace

int pdsat_enumerate(pdsat_db* db, const int32_t* set_vars, int32_t k, int32_t n_fixed_top, uint64_t fixed_top_value,
                    int32_t stop_at_first_model, uint64_t max_survivors, uint64_t* survivors, uint8_t* is_model,
                    uint64_t* n_survivors, pdsat_enum_stats* stats) {
    int rc = need_device(db);
    if (rc) return rc;
    if (!set_vars || k <= 0 || k > 62 || n_fixed_top < 0 || n_fixed_top >= k || !stats ||
        (n_fixed_top < 64 && n_fixed_top > 0 && (fixed_top_value >> n_fixed_top)) || (n_fixed_top == 0 && fixed_top_value))
        return fail(PDSAT_ERR_ARG, "pdsat_enumerate: bad argument");
    for (int i = 1; i < k; i++)
        if (set_vars[i] <= set_vars[i - 1]) return fail(PDSAT_ERR_ARG, "pdsat_enumerate: set_vars must be ascending");
    memset(stats, 0, sizeof(*stats));
    if (n_survivors) *n_survivors = 0;
    const int kf = k - n_fixed_top;
    stats->n_assignments = 1ull << kf;
    // level widths: the first level takes up to 16 free bits at once, the following ones 8 (never
    // fewer than 6: a block is at least one 64-sample group), the last one what is left
    std::vector<EnumLevel> lv;
    {
        int left = kf, depth = n_fixed_top;
        while (left > 0) {
            int w = lv.empty() ? 16 : 8;
            if (w > left) w = left;
            if (left - w > 0 && left - w < 6) w = left;  // do not leave a sliver for the last level
            if (w < 6) {  // tiny cubes: pad the enumeration with variables from the fixed top (w >= 6 required)
                if (lv.empty()) break;
                lv.back().w += w;
                left -= w;
                continue;
            }
            EnumLevel L;
            L.depth_before = depth;
            L.w = w;
            lv.push_back(L);
            depth += w;
            left -= w;
        }
    }
    auto cleanup = [&]() {
        for (auto& L : lv)
            if (L.b) pdsat_batch_destroy(L.b);
    };
    if (lv.empty()) {
        // fewer than 6 free bits: one plain ENUM batch over the whole (tiny) cube
        const uint64_t n = 1ull << kf;
        pdsat_point pt{};
        pt.set_vars = set_vars;
        pt.k = k;
        pt.mode = PDSAT_MODE_ENUM;
        pt.first_sample = fixed_top_value << kf;
        pt.n_samples = n;
        pdsat_batch* b = nullptr;
        if ((rc = pdsat_batch_create(db, 1, &pt, PDSAT_OUT_FLAGS, 0, &b))) return rc;
        uint64_t cw = 0, fw = 0;
        pdsat_cost c;
        rc = pdsat_batch_fill(b);
        if (!rc) rc = pdsat_batch_propagate(b);
        if (!rc) rc = pdsat_batch_costs(b, &c);
        if (!rc) rc = pdsat_batch_flags(b, 0, &cw, &fw);
        pdsat_batch_destroy(b);
        if (rc) return rc;
        stats->n_propagated = n;
        stats->n_batches = 1;
        uint64_t ns = 0;
        for (uint64_t s = 0; s < n; s++) {
            if ((cw >> s) & 1) continue;
            const bool model = (fw >> s) & 1;
            if (model) stats->n_models++;
            else stats->n_undecided++;
            if (ns < max_survivors && survivors) {
                survivors[ns] = (fixed_top_value << kf) | s;
                if (is_model) is_model[ns] = model;
            }
            ns++;
        }
        stats->n_conflict = n - stats->n_models - stats->n_undecided;
        if (n_survivors) *n_survivors = ns;
        return PDSAT_OK;
    }
    const uint64_t cap_samples = 1ull << 24;
    for (auto& L : lv) {
        L.cap_blocks = cap_samples >> L.w;
        if (L.cap_blocks < 1) L.cap_blocks = 1;
        // the level's set: the top depth_before + w variables (ascending = low bits first)
        const int d = L.depth_before + L.w;
        std::vector<uint64_t> zero(L.cap_blocks, 0);
        pdsat_point pt{};
        pt.set_vars = set_vars + (k - d);
        pt.k = d;
        pt.mode = PDSAT_MODE_ENUM_BLOCKS;
        pt.n_samples = L.cap_blocks << L.w;
        pt.block_prefix = zero.data();
        pt.block_bits = L.w;
        if ((rc = pdsat_batch_create(db, 1, &pt, PDSAT_OUT_FLAGS, 0, &L.b))) {
            cleanup();
            return rc;
        }
        L.conflict.resize((size_t)(pt.n_samples / 64));
        L.full.resize((size_t)(pt.n_samples / 64));
    }
    // explicit stack of (level, prefixes); chunks are pushed in reverse so that leaves come out
    // in ascending order
    struct Item {
        int level;
        std::vector<uint64_t> prefixes;
    };
    std::vector<Item> stack;
    stack.push_back(Item{0, {fixed_top_value}});
    uint64_t ns = 0;
    bool stop = false;
    const int last = (int)lv.size() - 1;
    while (!stack.empty() && !stop) {
        Item it = std::move(stack.back());
        stack.pop_back();
        EnumLevel& L = lv[it.level];
        const uint64_t nb = it.prefixes.size();
        if ((rc = pdsat_batch_set_blocks(L.b, nb, it.prefixes.data())) || (rc = pdsat_batch_fill(L.b)) ||
            (rc = pdsat_batch_propagate(L.b)) || (rc = pdsat_batch_flags(L.b, 0, L.conflict.data(), L.full.data()))) {
            cleanup();
            return rc;
        }
        stats->n_batches++;
        stats->n_propagated += nb << L.w;
        const uint64_t per_block_words = (1ull << L.w) / 64;
        std::vector<uint64_t> next;
        for (uint64_t blk = 0; blk < nb && !stop; blk++) {
            const uint64_t base = it.prefixes[blk] << L.w;
            for (uint64_t wi = 0; wi < per_block_words && !stop; wi++) {
                const uint64_t cw = L.conflict[blk * per_block_words + wi], fw = L.full[blk * per_block_words + wi];
                uint64_t alive = ~cw;
                while (alive) {
                    const int bit = __builtin_ctzll(alive);
                    alive &= alive - 1;
                    const uint64_t v = base | (wi * 64 + (uint64_t)bit);
                    if (it.level < last) {
                        next.push_back(v);
                    } else {
                        const bool model = (fw >> bit) & 1;
                        if (model) stats->n_models++;
                        else stats->n_undecided++;
                        if (ns < max_survivors && survivors) {
                            survivors[ns] = v;
                            if (is_model) is_model[ns] = model;
                        }
                        ns++;
                        if (model && stop_at_first_model) {
                            stop = true;
                            break;
                        }
                    }
                }
            }
        }
        if (it.level < last && !next.empty()) {
            const uint64_t cap = lv[it.level + 1].cap_blocks;
            const uint64_t n_chunks = (next.size() + cap - 1) / cap;
            for (uint64_t ch = n_chunks; ch-- > 0;) {
                const uint64_t lo = ch * cap, hi = std::min<uint64_t>(next.size(), lo + cap);
                stack.push_back(Item{it.level + 1, std::vector<uint64_t>(next.begin() + lo, next.begin() + hi)});
            }
        }
    }
    cleanup();
    if (!stop) stats->n_conflict = stats->n_assignments - stats->n_models - stats->n_undecided;
    if (n_survivors) *n_survivors = ns;
    return PDSAT_OK;
}

static pdsat_host::Cdcl::Limits cdcl_limits(const pdsat_cdcl_limits* lim, int* n_threads) {
    pdsat_host::Cdcl::Limits L;
    *n_threads = 0;
    if (lim) {
        L.max_conflicts = lim->max_conflicts;
        L.max_restarts = lim->max_restarts;
        L.max_seconds = lim->max_seconds;
        *n_threads = lim->n_threads;
    }
    return L;
}

static void cdcl_stats_out(const SolveManyOut& o, uint64_t n, pdsat_cdcl_stats* st) {
    if (!st) return;
    st->n_problems = n;
    st->n_sat = o.n_sat;
    st->n_unsat = o.n_unsat;
    st->n_unknown = o.n_unknown;
    st->decisions = o.decisions;
    st->propagations = o.propagations;
    st->conflicts = o.conflicts;
    st->seconds = o.seconds;
    st->threads = o.threads;
}

int pdsat_cdcl_solve_many(pdsat_db* db, int64_t n, int32_t k, const int32_t* assumps, const pdsat_cdcl_limits* lim,
                          uint8_t* status, uint64_t* props, int32_t max_models, int64_t* model_problem, uint8_t* models,
                          int64_t* n_models, pdsat_cdcl_stats* stats) {
    if (!db || n < 0 || k < 0 || (n > 0 && k > 0 && !assumps) || (n > 0 && !status))
        return fail(PDSAT_ERR_ARG, "pdsat_cdcl_solve_many: bad argument");
    const HostDb& h = db->host;
    for (int64_t i = 0; i < n * (int64_t)k; i++)
        if (assumps[i] < 0 || (assumps[i] >> 1) >= h.n_vars) return fail(PDSAT_ERR_ARG, "pdsat_cdcl_solve_many: literal out of range");
    int nt = 0;
    const pdsat_host::Cdcl::Limits L = cdcl_limits(lim, &nt);
    SolveManyOut o;
    if (n_models) *n_models = 0;
    if (n > 0) solve_many(db->cdcl, h, db->generation, n, k, assumps, L, nt, status, props, max_models, model_problem, models, n_models, o);
    cdcl_stats_out(o, (uint64_t)n, stats);
    return PDSAT_OK;
}

int pdsat_batch_sample_values(pdsat_batch* b, int32_t point, uint64_t first, uint64_t n, uint8_t* values) {
    if (!b || !values || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_sample_values: bad argument");
    if (!b->filled) return fail(PDSAT_ERR_ARG, "pdsat_batch_sample_values: batch not filled");
    const HostPoint& hp = b->pts[point];
    if (first + n > hp.n_samples) return fail(PDSAT_ERR_ARG, "pdsat_batch_sample_values: range");
    if (n == 0) return PDSAT_OK;
    CU(cudaSetDevice(b->db->device));
    CU(cudaStreamSynchronize(b->db->stream));
    const size_t k = hp.set_vars.size();
    const uint64_t g0 = first / 64, g1 = (first + n - 1) / 64;
    std::vector<uint64_t> w((size_t)(g1 - g0 + 1) * hp.kp);
    CU(cudaMemcpy(w.data(), b->d_words + hp.word_off + g0 * hp.kp, sizeof(uint64_t) * w.size(), cudaMemcpyDeviceToHost));
    for (uint64_t s = first; s < first + n; s++) {
        const uint64_t* gw = &w[(size_t)(s / 64 - g0) * hp.kp];
        const int bit = (int)(s & 63);
        uint8_t* dst = values + (s - first) * k;
        for (size_t j = 0; j < k; j++) dst[j] = (uint8_t)((gw[j] >> bit) & 1ull);
    }
    return PDSAT_OK;
}

int pdsat_batch_handoff(pdsat_batch* b, int32_t point, const pdsat_cdcl_limits* lim, uint8_t* status, uint64_t* cost,
                        int32_t max_models, uint64_t* model_sample, uint8_t* models, int64_t* n_models,
                        pdsat_cdcl_stats* stats) {
    if (!b || !status || point < 0 || point >= (int)b->pts.size()) return fail(PDSAT_ERR_ARG, "pdsat_batch_handoff: bad argument");
    if (!b->d_conflict || !b->propagated) return fail(PDSAT_ERR_ARG, "pdsat_batch_handoff: needs PDSAT_OUT_FLAGS and a propagated batch");
    if (cost && !b->d_trail) return fail(PDSAT_ERR_ARG, "pdsat_batch_handoff: cost needs PDSAT_OUT_TRAIL");
    pdsat_db* db = b->db;
    const HostDb& h = db->host;
    const HostPoint& hp = b->pts[point];
    const uint64_t n = hp.n_samples;
    std::vector<uint64_t> cw((size_t)hp.n_groups), fw((size_t)hp.n_groups);
    int rc = pdsat_batch_flags(b, point, cw.data(), fw.data());
    if (rc) return rc;
    std::vector<uint32_t> trail;
    if (cost) {
        trail.resize((size_t)n);
        if ((rc = pdsat_batch_trail(b, point, trail.data()))) return rc;
    }
    std::vector<uint64_t> open_idx;
    for (uint64_t s = 0; s < n; s++) {
        const bool cf = (cw[s >> 6] >> (s & 63)) & 1, fu = (fw[s >> 6] >> (s & 63)) & 1;
        status[s] = cf ? PDSAT_UNSAT : (fu ? PDSAT_SAT : PDSAT_UNKNOWN);
        if (cost) cost[s] = cf ? (uint64_t)h.n_base_assigned + hp.set_vars.size() : trail[s];
        if (!cf && !fu) open_idx.push_back(s);
    }
    if (n_models) *n_models = 0;
    SolveManyOut o;
    const size_t k = hp.set_vars.size();
    if (!open_idx.empty()) {
        // the assumptions of the open samples, as the device used them (from the bit-sliced words)
        CU(cudaSetDevice(db->device));
        std::vector<uint64_t> w((size_t)hp.n_groups * hp.kp);
        CU(cudaMemcpy(w.data(), b->d_words + hp.word_off, sizeof(uint64_t) * w.size(), cudaMemcpyDeviceToHost));
        std::vector<int32_t> assumps(open_idx.size() * k);
        for (size_t i = 0; i < open_idx.size(); i++) {
            const uint64_t s = open_idx[i];
            const uint64_t* gw = &w[(size_t)(s / 64) * hp.kp];
            for (size_t j = 0; j < k; j++)
                assumps[i * k + j] = 2 * (hp.set_vars[j] - 1) + (((gw[j] >> (s & 63)) & 1ull) ? 0 : 1);
        }
        std::vector<uint8_t> st(open_idx.size());
        std::vector<uint64_t> pr(open_idx.size());
        std::vector<int64_t> mp((size_t)(max_models > 0 ? max_models : 0));
        int nt = 0;
        const pdsat_host::Cdcl::Limits L = cdcl_limits(lim, &nt);
        int64_t nm = 0;
        solve_many(db->cdcl, h, db->generation, (int64_t)open_idx.size(), (int32_t)k, assumps.data(), L, nt, st.data(), pr.data(),
                   max_models, mp.data(), models, &nm, o);
        for (size_t i = 0; i < open_idx.size(); i++) {
            status[open_idx[i]] = st[i];
            if (cost) cost[open_idx[i]] = (uint64_t)h.n_base_assigned + pr[i];
        }
        const int64_t stored = nm < max_models ? nm : max_models;
        for (int64_t j = 0; j < stored && model_sample; j++) model_sample[j] = open_idx[(size_t)mp[(size_t)j]];
        if (n_models) *n_models = nm;
    }
    cdcl_stats_out(o, open_idx.size(), stats);
    return PDSAT_OK;
}

int pdsat_eval_batch(pdsat_db* db, int32_t n_points, const pdsat_point* points, pdsat_cost* costs) {
    if (!costs) return fail(PDSAT_ERR_ARG, "pdsat_eval_batch: null costs");
    pdsat_batch* b = nullptr;
    int rc = pdsat_batch_create(db, n_points, points, 0, 0, &b);
    if (rc) return rc;
    rc = pdsat_batch_fill(b);
    if (!rc) rc = pdsat_batch_propagate(b);
    if (!rc) rc = pdsat_batch_costs(b, costs);
    std::string keep = g_err;
    pdsat_batch_destroy(b);
    if (rc) g_err = keep;
    return rc;
}

int pdsat_mt19937_bool_rand(int device, uint32_t seed, uint64_t first, uint64_t n, uint8_t* out) {
    if (!out) return fail(PDSAT_ERR_ARG, "pdsat_mt19937_bool_rand: null out");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt)
        return fail(PDSAT_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    // same machinery as a batch: seed state, jump to round r0 (one jump per set bit between
    // the two slots), split the window into 4 segments by three jumps (L, 2L, 3L rounds) that
    // share the sequence of the first segment, emit the bits
    const uint64_t r0 = first / 624, skip = first - r0 * 624;
    const uint64_t need = (skip + n + 623) / 624;
    int m = 1;
    while ((4ull << m) < need) m++;
    if (m > 58) return fail(PDSAT_ERR_LIMIT, "mt19937 stream too long");
    const uint64_t L = 1ull << m;
    const int n_segs = 4;
    const uint64_t words = n_segs * (L / 2) * 39;
    std::vector<uint32_t> state(624), hbits(words);
    mt19937_seed(seed, state.data());
    std::vector<MtJumpTask> tasks;
    std::vector<uint32_t> srcs;
    std::vector<uint64_t> poly_rounds;  // device slot i holds x^(624 * poly_rounds[i])
    struct Launch { size_t off, cnt, src_off, src_cnt; };
    std::vector<Launch> launches;
    uint32_t cur = 0;  // slots 0/1: ping-pong pair, slots 2..4: segments 1..3
    for (int t = 0; t < 64; t++)
        if ((r0 >> t) & 1ull) {
            launches.push_back({tasks.size(), 1, srcs.size(), 1});
            tasks.push_back(MtJumpTask{0, cur ^ 1u, (uint32_t)poly_rounds.size(), 0});
            srcs.push_back(cur);
            poly_rounds.push_back(1ull << t);
            cur ^= 1u;
        }
    launches.push_back({tasks.size(), 3, srcs.size(), 1});
    srcs.push_back(cur);
    for (uint32_t d = 1; d <= 3; d++) {
        tasks.push_back(MtJumpTask{0, d + 1, (uint32_t)poly_rounds.size(), 0});
        poly_rounds.push_back(d * L);
    }
    std::vector<MtSegment> segs(n_segs);
    for (int j = 0; j < n_segs; j++) segs[j] = MtSegment{(uint64_t)j * (L / 2) * 39, L, j == 0 ? cur : (uint32_t)(j + 1), 0};
    uint32_t *d_state = nullptr, *d_out = nullptr, *d_polys = nullptr, *d_srcs = nullptr;
    MtSegment* d_seg = nullptr;
    MtJumpTask* d_tasks = nullptr;
    CU(cudaMalloc(&d_state, 624 * 4 * MT_PARTS * 5));
    CU(cudaMemset(d_state, 0, 624 * 4 * MT_PARTS * 5));
    CU(cudaMalloc(&d_seg, sizeof(MtSegment) * n_segs));
    CU(cudaMalloc(&d_out, words * 4));
    CU(cudaMalloc(&d_polys, poly_rounds.size() * 624 * 4));
    CU(cudaMalloc(&d_tasks, sizeof(MtJumpTask) * tasks.size()));
    CU(cudaMalloc(&d_srcs, sizeof(uint32_t) * srcs.size()));
    for (size_t i = 0; i < poly_rounds.size(); i++) {
        const uint64_t r = poly_rounds[i];
        const uint32_t* pl = __builtin_popcountll(r) == 1 ? MtJumpPolys::instance().poly(__builtin_ctzll(r))
                                                          : MtJumpPolys::instance().poly_for_rounds(r);
        CU(cudaMemcpy(d_polys + i * 624, pl, 624 * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(d_state, state.data(), 624 * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_seg, segs.data(), sizeof(MtSegment) * n_segs, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_tasks, tasks.data(), sizeof(MtJumpTask) * tasks.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_srcs, srcs.data(), sizeof(uint32_t) * srcs.size(), cudaMemcpyHostToDevice));
    uint32_t* d_seq = nullptr;
    CU(cudaMalloc(&d_seq, sizeof(uint32_t) * MT_SEQ_WORDS));
    CU(cudaFuncSetAttribute(mt19937_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SEQ_SMEM));
    for (auto& l : launches)
        mt19937_launch_jumps(d_polys, d_state, d_srcs + l.src_off, l.src_cnt, d_tasks + l.off, l.cnt, d_seq, 0);
    mt19937_bits_kernel<<<(n_segs + MT_BITS_WARPS - 1) / MT_BITS_WARPS, MT_BITS_WARPS * 32>>>(d_state, d_seg, n_segs, (uint32_t)L, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpy(hbits.data(), d_out, words * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_state);
    cudaFree(d_seq);
    cudaFree(d_seg);
    cudaFree(d_out);
    cudaFree(d_polys);
    cudaFree(d_tasks);
    cudaFree(d_srcs);
    for (uint64_t i = 0; i < n; i++) {
        uint64_t d = skip + i;
        out[i] = (hbits[d >> 5] >> (d & 31)) & 1u;
    }
    return PDSAT_OK;
}

#ifdef PDSAT_TIMELINE
int pdsat_debug_timeline(unsigned long long* out) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 148 * 8));
    return PDSAT_OK;
}
#endif

}  // extern "C"
