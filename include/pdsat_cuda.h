/* pdsat_cuda.h — C-ABI boundary of the B200-native PDSAT hot path.
 *
 * The reference (Nauchnik/pdsat) has no plugin / FFI interface; the seam is cut where
 * SURVEY.md §8(b) puts it.  Each entry point names the reference code it replaces
 * (paths relative to the reference root).  The library is libpdsat_b200.so, built from
 * pdsat_b200/csrc/ for sm_100a; there is NO CPU fallback behind these calls: without a
 * CUDA device every compute entry point fails with PDSAT_ERR_CUDA.
 *
 * Data contract:
 *   - literals are MiniSat `Lit.x` = 2*var + sign, var 0-based, sign 1 = negated
 *     (src_common/minisat/core/SolverTypes.h:46-62);
 *   - decomposition-set variables are 1-based DIMACS indices, as in
 *     decomp_set::var_choose_order (src_mpi/mpi_predicter.h:58-64);
 *   - lbool bytes: 0 true, 1 false, 2 undefined (SolverTypes.h:84-115).
 *
 * Threading: one pdsat_db per GPU; calls on one handle (and its batches) are serialised
 * by the caller (the reference is one single-threaded process per MPI rank).
 * Errors: int status, message via pdsat_last_error(); the library never exits.
 */
#ifndef PDSAT_CUDA_H
#define PDSAT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdsat_db pdsat_db;       /* clause database resident in one GPU's HBM */
typedef struct pdsat_batch pdsat_batch; /* device buffers of one batch of samples      */

enum {
    PDSAT_OK = 0,
    PDSAT_ERR_ARG = 1,   /* bad argument                                              */
    PDSAT_ERR_CUDA = 2,  /* CUDA runtime error / no device                            */
    PDSAT_ERR_UNSAT = 3, /* instance refuted at level 0 (Solver::ok == false)         */
    PDSAT_ERR_LIMIT = 4  /* instance exceeds a structural limit of the kernels        */
};

/* How the values of the decomposition-set variables of sample s are obtained. */
enum {
    /* bit j of sample s = bool_rand of draw first_draw + (first_sample+s)*k + j of
     * mt19937(seed): the reference's 1-worker order (src_mpi/mpi_predicter.cpp:3236-3242,
     * src_common/addit_func.cpp:193-196; bool_rand == ((draw >> 31) == 0)).  first_draw is
     * the number of draws the worker's generator has already made (it is seeded once,
     * mpi_predicter.cpp:87, and runs on across points of different k). */
    PDSAT_MODE_RANDOM_MT19937 = 0,
    /* declared counter-based generator (NOT the reference stream): the 64 values of set
     * position j for samples [64g, 64g+64) are the bits of
     * splitmix64(seed + 0x9E3779B97F4A7C15 * (g * 65536 + j + 1)), g counted from
     * first_sample / 64 (first_sample must be a multiple of 64). */
    PDSAT_MODE_RANDOM_COUNTER = 1,
    /* solve-mode enumeration: sample s is assignment index lint = first_sample + s;
     * bit r of lint is the value of set_vars[r]
     * (src_mpi/mpi_base.cpp:148-169 MakeAssignsFromMasks; pass set_vars ascending). */
    PDSAT_MODE_ENUM = 2,
    /* values supplied by the caller: explicit_values[s*k + j] in {0,1}
     * (te > 0 mode: known states + keystream, src_mpi/mpi_predicter.cpp:3188-3235). */
    PDSAT_MODE_EXPLICIT = 3,
    /* enumeration of sub-cubes: the samples come in blocks of 2^block_bits (block_bits >= 6);
     * sample s is assignment index lint = (block_prefix[s >> block_bits] << block_bits)
     * | (s mod 2^block_bits), bit r of lint = value of set_vars[r] as in PDSAT_MODE_ENUM.
     * The prefixes of a refuted partial assignment's sub-cube simply do not appear
     * (BCP-pruned enumeration, src_common/minisat/core/Solver.cc:1380-1423). */
    PDSAT_MODE_ENUM_BLOCKS = 4
};

/* One point = one decomposition set and its sample range.  A true value makes the
 * POSITIVE literal of the variable the assumption (mpi_predicter.cpp:3241). */
typedef struct {
    const int32_t* set_vars;         /* k variables, 1-based                           */
    int32_t k;
    int32_t mode;                    /* PDSAT_MODE_*                                   */
    uint64_t seed;                   /* RANDOM modes (reference: the worker's rank)    */
    uint64_t first_sample;
    uint64_t n_samples;
    const uint8_t* explicit_values;  /* EXPLICIT mode only: n_samples * k bytes        */
    uint64_t first_draw;             /* MT19937 mode: draws already consumed (0 = fresh) */
    const uint64_t* block_prefix;    /* ENUM_BLOCKS mode: n_samples >> block_bits prefixes */
    int32_t block_bits;
} pdsat_point;

/* Integer-exact Monte Carlo accumulators of one point: what GetPredict needs
 * (src_mpi/mpi_predicter.cpp:1904-2115).  `trail` of a sample is the
 * number of assigned variables at the BCP fixpoint (level-0 units + fixed assumptions +
 * set variables + implied) = Solver::propagations of a conflict-free sample. */
#define PDSAT_COST_WORDS 8
typedef struct {
    uint64_t n;            /* samples evaluated                                        */
    uint64_t n_conflict;   /* BCP conflict  (ret == l_False, isSolvedOnPreprocessing)  */
    uint64_t n_full;       /* BCP assigned every variable: a model (ret == l_True)     */
    uint64_t sum_trail;    /* over conflict-free samples                               */
    uint64_t sumsq_trail;  /* over conflict-free samples                               */
    uint64_t n_implied_any;/* conflict-free samples where BCP implied >= 1 literal     */
    uint64_t queue_pops;   /* diagnostics: occurrence lists scanned (per 64-sample group) */
    uint64_t implied_lits; /* diagnostics: distinct implied literals (per 64-sample group)*/
} pdsat_cost;

typedef struct {
    int32_t n_vars;          /* as uploaded                                            */
    int32_t n_live_vars;     /* not assigned by level-0 units / fixed assumptions      */
    int64_t n_clauses_in;    /* as uploaded                                            */
    int64_t n_clauses_live;  /* after addClause_ normalisation + level-0 simplification*/
    int64_t n_lits_live;
    int32_t n_base_assigned; /* level-0 trail size incl. fixed assumptions             */
    int32_t max_clause_len;
    int32_t status;          /* PDSAT_OK or PDSAT_ERR_UNSAT                            */
    int32_t warps_per_cta;   /* launch shape chosen for the BCP kernel                 */
    int32_t smem_bytes;      /* dynamic shared memory per CTA                          */
    int32_t grid;            /* CTAs launched                                          */
    int32_t db_in_smem;      /* 1 = index blob (occurrence offsets, gate lists, gate records)
                                staged in shared memory by one bulk copy per CTA         */
    int32_t sm_count;
    int32_t n_gates;         /* XOR / ANDXOR gates recognised among the live clauses     */
    int32_t blob_bytes;      /* size of the index blob                                   */
    int64_t n_gate_clauses;  /* live clauses replaced by gate records                    */
} pdsat_db_info;

/* One recognised gate (inspection / tests): type 0 = XOR  x[0]^...^x[m-1] = parity,
 * type 1 = ANDXOR  x[0]^...^x[m-1] ^ (lb & lc) = parity.  x[] are 1-based variables,
 * lb / lc MiniSat literals (2*var+sign). */
typedef struct {
    int32_t type, parity, m, n_clauses;
    int32_t x[7];
    int32_t lb, lc;
} pdsat_gate;

/* output selection for pdsat_batch_create */
enum {
    PDSAT_OUT_FLAGS = 1,  /* per-sample conflict / full bitmaps                        */
    PDSAT_OUT_TRAIL = 2,  /* per-sample trail size (uint32)                            */
    PDSAT_OUT_ASSIGN = 4  /* per-sample assignment of every variable (parity tests)    */
};

const char* pdsat_last_error(void);
int pdsat_device_count(int* n);

/* ---- clause database upload ----------------------------------------------------------
 * Replaces Solver::addProblem + addClause_ + attachClause
 * (src_common/minisat/core/Solver.h:478-492, Solver.cc:279-317) executed once per sample
 * by solverProgramCalling (src_mpi/mpi_predicter.cpp:703-705).  Input is the `Problem`
 * produced by minisat22_wrapper::parse_DIMACS_to_problem (src_common/minisat22_wrapper.cpp:43-69)
 * flattened to CSR.  Applies the addClause_ rules (sort, duplicate / tautology removal),
 * folds level-0 units by one host BCP and builds the flat clause-record + occurrence-list
 * layout in HBM.  A level-0 conflict is reported through pdsat_db_get_info().status and by
 * every later evaluate call counting all samples as conflicts (Solver::ok == false).
 * Host buffers stay owned by the caller. */
int pdsat_upload_cnf(int device, int32_t n_vars, int64_t n_clauses, const uint32_t* offsets,
                     const int32_t* lits, pdsat_db** out);

/* Literals assumed in EVERY sample: the keystream bits of a cipher instance
 * (src_mpi/mpi_predicter.cpp:3229-3234) or solve mode's known_dummy
 * (src_mpi/mpi_solver.cpp:199-200).  They are folded into the level-0 assignment by host
 * BCP and the device database is rebuilt; n = 0 restores the plain template.
 * The rebuild renumbers the live variables: every pdsat_batch created before this call is
 * stale and its fill / propagate calls fail with PDSAT_ERR_ARG (destroy and re-create it). */
int pdsat_set_fixed_assumptions(pdsat_db* db, const int32_t* lits, int32_t n);
/* gates recognised in the current database: *n_gates = their number, at most cap are
 * written to out (out may be NULL) */
int pdsat_db_gates(const pdsat_db* db, int32_t cap, pdsat_gate* out, int32_t* n_gates);

int pdsat_db_get_info(const pdsat_db* db, pdsat_db_info* info);
/* level-0 assignment incl. fixed assumptions, n_vars lbool bytes */
int pdsat_db_base_assignment(const pdsat_db* db, uint8_t* assigns);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
int pdsat_set_stream(pdsat_db* db, void* cuda_stream);
/* batches of this database that are still alive keep it alive: the memory is released when
 * the last of them is destroyed */
void pdsat_destroy(pdsat_db* db);

/* ---- batch evaluate --------------------------------------------------------------------
 * Replaces, for a whole batch, solvePredictSatInstance + solverProgramCalling
 * (src_mpi/mpi_predicter.cpp:3183-3277, 691-773) with evaluation_type "prep" /
 * "propagation", and SolverRun's BCP-decided share (src_mpi/mpi_solver.cpp:167-294).
 * Several points (a tabu neighbourhood, mpi_predicter.cpp:2926-2935) go in one batch. */
int pdsat_batch_create(pdsat_db* db, int32_t n_points, const pdsat_point* points, uint32_t out_flags,
                       int32_t max_models, pdsat_batch** out);
/* change the stream position of a point without reallocating (RANDOM / ENUM modes) */
int pdsat_batch_reseed(pdsat_batch* b, int32_t point, uint64_t seed, uint64_t first_sample);
/* move a MT19937 point along its worker's stream (see pdsat_point.first_draw) */
int pdsat_batch_set_first_draw(pdsat_batch* b, int32_t point, uint64_t first_draw);
/* stage 1 (async): sample values -> bit-sliced assumption words in HBM */
int pdsat_batch_fill(pdsat_batch* b);
/* stage 2 (async): BCP of every sample + per-point cost accumulation */
int pdsat_batch_propagate(pdsat_batch* b);
int pdsat_batch_sync(pdsat_batch* b);
/* device time of the last fill / propagate on this batch, CUDA events on the launch stream */
int pdsat_batch_last_ms(pdsat_batch* b, float* fill_ms, float* propagate_ms);
/* Stage timing on (default) / off.  The four event records per step are device-side work on the
 * launch stream (about 1 us each): a caller that launches steps back to back and never asks for
 * pdsat_batch_last_ms turns them off.  With timing off pdsat_batch_last_ms reports 0. */
int pdsat_batch_set_timing(pdsat_batch* b, int32_t on);
int pdsat_batch_launch_count(pdsat_batch* b, int64_t* n_kernels);
/* host->device bytes copied by the last pdsat_batch_fill (descriptors, generator seeds, jump plan,
 * explicit values) */
int pdsat_batch_h2d_bytes(pdsat_batch* b, int64_t* bytes);

/* ---- cost reduce -------------------------------------------------------------------------
 * Replaces the per-sample MPI replies (mpi_predicter.cpp:3266-3270) and their storage
 * (:504-516).  The accumulators live in HBM as n_points * PDSAT_COST_WORDS uint64; the
 * device pointer is exposed so that ranks can ncclAllReduce(sum, uint64) it in place. */
int pdsat_batch_cost_device_ptr(pdsat_batch* b, void** dptr);
/* use a caller-owned device buffer (n_points * PDSAT_COST_WORDS uint64) instead */
int pdsat_batch_set_cost_buffer(pdsat_batch* b, void* dptr);
int pdsat_batch_costs(pdsat_batch* b, pdsat_cost* out /* n_points */);

/* Multi-GPU, one process per GPU: cost all-reduce through the BCP kernel over NVLink peer memory
 * instead of a separate collective.  Every rank exports a 64-byte CUDA IPC handle of its
 * receive buffer, the ranks exchange the handles out of band (all_gather of 64 bytes), every
 * rank connects; from then on the sequence pdsat_batch_propagate ... pdsat_batch_costs is
 * COLLECTIVE (all ranks launch the same steps and ask for the costs at the same steps) and
 * pdsat_batch_costs returns the summed accumulators of ALL ranks.
 * The exchange is deferred: a launch only adds into the rank's scratch; CTA 0 of the next launch
 * gathers the peers' tagged 8-byte packets of the step before and pushes this rank's sums into
 * every peer's buffer while the other CTAs already work (no fence, no atomics, no waiting at the
 * end of a launch); pdsat_batch_costs / pdsat_batch_sync finish the last step with a one-CTA
 * kernel.  The totals are therefore valid after those two calls, not in stream order after
 * pdsat_batch_propagate.  A wait for a rank that never launched its step gives up after 20 s
 * (PDSAT_ERR_CUDA from pdsat_batch_costs).  Replaces the ncclAllReduce of SURVEY.md 8(e);
 * results are identical (integer sums).  Use it with one rank per GPU only. */
#define PDSAT_IPC_HANDLE_BYTES 64
int pdsat_batch_peer_export(pdsat_batch* b, void* handle_out /* 64 bytes */);
int pdsat_batch_peer_connect(pdsat_batch* b, int32_t n_ranks, int32_t my_rank,
                             const void* handles /* n_ranks * 64 bytes, rank order */);
int pdsat_batch_peer_disconnect(pdsat_batch* b);

/* ---- per-sample results (optional outputs) ------------------------------------------- */
/* bit s%64 of word s/64 (s counted inside the point) */
int pdsat_batch_flags(pdsat_batch* b, int32_t point, uint64_t* conflict_bits, uint64_t* full_bits);
int pdsat_batch_trail(pdsat_batch* b, int32_t point, uint32_t* trail /* n_samples */);
/* lbool bytes [n][n_vars] for samples first..first+n-1 of the point */
int pdsat_batch_assign(pdsat_batch* b, int32_t point, uint64_t first, uint64_t n, uint8_t* assigns);
/* models found by BCP alone (b_SAT_set_array, mpi_predicter.cpp:760-762): *n_found is the
 * total, at most max_models are stored; models[i*n_vars + v] in {0,1} (1 = true) */
int pdsat_batch_models(pdsat_batch* b, int64_t* n_found, int32_t* point_of, uint64_t* sample_of,
                       uint8_t* models);
void pdsat_batch_destroy(pdsat_batch* b);

/* ---- solve mode: BCP-pruned enumeration of a cube ---------------------------------------
 * Replaces MPI_Base::MakeAssignsFromMasks (src_mpi/mpi_base.cpp:107-178) + the BCP-decided share
 * of MPI_Solver::SolverRun (src_mpi/mpi_solver.cpp:167-294) for all 2^k assignments of set_vars
 * (ascending; bit r of the assignment index lint = value of set_vars[r]), with the traversal
 * idea of Solver::gen_valid_assumptions_rc2 (Solver.cc:1293-1435): the variables are assigned
 * from the most significant end in blocks, BCP runs on the device after every block, and a
 * refuted prefix prunes its whole sub-cube.  The survivors are exactly the assignments whose
 * full BCP finds no conflict (BCP is monotone), in ascending lint order.
 * n_fixed_top / fixed_top_value pin the top bits of lint: the task split of
 * ControlProcessSolve (mpi_solver.cpp:447-473), e.g. rank g of 8 takes n_fixed_top = 3,
 * fixed_top_value = g. */
typedef struct {
    uint64_t n_assignments; /* 2^(k - n_fixed_top)                                         */
    uint64_t n_conflict;    /* refuted by BCP (pruned sub-cubes included)                  */
    uint64_t n_models;      /* BCP reached a full assignment without conflict              */
    uint64_t n_undecided;   /* conflict-free, not full: would go to CDCL                   */
    uint64_t n_propagated;  /* samples actually propagated on the device                   */
    uint64_t n_batches;     /* device batches launched                                     */
} pdsat_enum_stats;
/* survivors: at most max_survivors assignment indices (models and undecided, ascending) with
 * is_model flags; stop_at_first_model ends the search at the first model (the reference stops at
 * the first SAT unless -solve_all, mpi_solver.cpp:285-286; counts are then partial). */
int pdsat_enumerate(pdsat_db* db, const int32_t* set_vars, int32_t k, int32_t n_fixed_top, uint64_t fixed_top_value,
                    int32_t stop_at_first_model, uint64_t max_survivors, uint64_t* survivors, uint8_t* is_model,
                    uint64_t* n_survivors, pdsat_enum_stats* stats);
/* change the blocks of a single-point ENUM_BLOCKS batch (n_blocks <= the number it was created with) */
int pdsat_batch_set_blocks(pdsat_batch* b, uint64_t n_blocks, const uint64_t* block_prefix);

/* ---- hand-off of what BCP leaves open to a host CDCL ---------------------------------------
 * A sample (predict) or assignment (solve) without conflict and without a full assignment needs
 * search: in the reference the rest of solve() after the level-0 propagation
 * (src_common/minisat/core/Solver.cc:775-885, called at src_mpi/mpi_predicter.cpp:731 and
 * src_mpi/mpi_solver.cpp:218-226).  The library runs its own CDCL (pdsat_b200/host/cdcl.hpp, not
 * the reference's code) on host threads, one incremental solver per thread over the uploaded
 * clauses + fixed assumptions.  Verdicts are properties of the formula (equal to the reference's);
 * decisions / propagations / conflicts are this solver's own counts. */
typedef struct {
    int64_t max_conflicts; /* per problem, < 0: none                                          */
    int64_t max_restarts;  /* per problem, <= 0: none (max_nof_restarts)                      */
    double max_seconds;    /* per problem, < 0: none (max_solving_time)                       */
    int32_t n_threads;     /* <= 0: all host cores                                            */
} pdsat_cdcl_limits;
typedef struct {
    uint64_t n_problems;   /* handed to the host                                              */
    uint64_t n_sat, n_unsat, n_unknown;
    uint64_t decisions, propagations, conflicts;
    double seconds;        /* wall clock of the hand-off                                      */
    int32_t threads;
} pdsat_cdcl_stats;
enum { PDSAT_SAT = 0, PDSAT_UNSAT = 1, PDSAT_UNKNOWN = 2 };
/* n problems of k assumption literals each (MiniSat codes, row-major).  status[i] in
 * {PDSAT_SAT, PDSAT_UNSAT, PDSAT_UNKNOWN}; props[i] (optional) = propagations of problem i;
 * the first max_models satisfying assignments go to models[j*n_vars + v] in {0,1} with
 * model_problem[j] = i (both optional). */
int pdsat_cdcl_solve_many(pdsat_db* db, int64_t n, int32_t k, const int32_t* assumps, const pdsat_cdcl_limits* lim,
                          uint8_t* status, uint64_t* props, int32_t max_models, int64_t* model_problem, uint8_t* models,
                          int64_t* n_models, pdsat_cdcl_stats* stats);
/* One point of a propagated batch (created with PDSAT_OUT_FLAGS): every sample gets a verdict -
 * BCP conflicts are PDSAT_UNSAT, BCP models PDSAT_SAT, the undecided ones are solved on the host
 * under the sample's assumptions.  cost[s] (optional) = Solver::propagations-like cost of the
 * sample: the BCP trail for samples BCP decides without conflict (what the reference counts),
 * the host solver's propagation count + the level-0 trail for handed-off samples; for a BCP
 * conflict the level-0 trail + the k assumptions (the reference's count at a conflict depends
 * on its watch order - this is the order-independent part of it).  max_models caps the models returned
 * from the handed-off samples (BCP's own models come from pdsat_batch_models). */
int pdsat_batch_handoff(pdsat_batch* b, int32_t point, const pdsat_cdcl_limits* lim, uint8_t* status, uint64_t* cost,
                        int32_t max_models, uint64_t* model_sample, uint8_t* models, int64_t* n_models,
                        pdsat_cdcl_stats* stats);
/* the values of the set variables of samples first..first+n-1 as the device used them:
 * values[(s-first)*k + j] in {0,1} */
int pdsat_batch_sample_values(pdsat_batch* b, int32_t point, uint64_t first, uint64_t n, uint8_t* values);

/* ---- one-call form (host buffers in, host buffers out; the e2e path) ---------------- */
int pdsat_eval_batch(pdsat_db* db, int32_t n_points, const pdsat_point* points, pdsat_cost* costs);

/* mt19937 bool_rand bits produced by the device generator, for stream-parity tests:
 * out[i] = bool_rand of draw first+i */
int pdsat_mt19937_bool_rand(int device, uint32_t seed, uint64_t first, uint64_t n, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PDSAT_CUDA_H */
