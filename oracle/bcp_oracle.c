/* oracle/bcp_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference algorithm on PDSAT's hot path, used ONLY as
 * the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs.  Nothing under pdsat_b200/ links, loads or calls it.
 *
 * Parity status: PINNED.  This restatement is validated (tests/test_oracle.py, run in
 * the build container where /root/reference exists) against the UNMODIFIED reference
 * MiniSat compiled into oracle/_ref/libpdsat_ref.so: conflict flag, assignment,
 * `propagations` and `watch_scans` agree sample by sample, and the resulting vectors are
 * committed under tests/golden/.  The random bit rule (Boost uniform_int_distribution,
 * Boost is absent from the tree) is restated from its published algorithm and is the one
 * declared contract: bool_rand == ((mt19937() >> 31) == 0).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- literals
 * src_common/minisat/core/SolverTypes.h:46-62: Lit.x = 2*var + sign (sign 1 = negated)
 * SolverTypes.h:84-115: lbool 0 true, 1 false, 2 undef; value(lit) = assigns[var] ^ sign */
#define VAR(l) ((l) >> 1)
#define SIGN(l) ((l)&1)
#define NEG(l) ((l) ^ 1)
enum { LB_TRUE = 0, LB_FALSE = 1, LB_UNDEF = 2 };

typedef struct {
    int32_t cref;
    int32_t blocker;
} watcher_t;

typedef struct {
    watcher_t* w;
    int32_t n, cap;
} wlist_t;

typedef struct {
    int n_vars;
    /* input problem, kept verbatim for the fresh-solver-per-sample path */
    int64_t in_n_clauses;
    uint32_t* in_offs;
    int32_t* in_lits;
} problem_t;

typedef struct {
    int n_vars;
    int ok;
    uint8_t* assigns;   /* per var: LB_* */
    int32_t* trail;
    int32_t trail_n, qhead;
    int32_t* arena;     /* clause arena: [size][lits...] */
    int64_t arena_n, arena_cap;
    int32_t* clauses;   /* crefs in insertion order */
    int64_t clauses_n, clauses_cap;
    wlist_t* watches;   /* 2*n_vars lists, indexed by literal */
    uint64_t propagations;
    uint64_t watch_scans;
} solver_t;

typedef struct {
    problem_t P;
    solver_t* persistent;
} handle_t;

typedef struct {
    int32_t conflict;
    int32_t full;
    int32_t sat;
    int32_t prep;
    int32_t ret;
    int32_t trail_size;
    uint64_t propagations;
    uint64_t watch_scans;
    uint64_t decisions;
    uint64_t conflicts;
} orc_result;

static inline int lit_value(const solver_t* S, int32_t l) {
    uint8_t a = S->assigns[VAR(l)];
    return a == LB_UNDEF ? LB_UNDEF : (a ^ SIGN(l));
}

static void wl_push(wlist_t* L, int32_t cref, int32_t blocker) {
    if (L->n == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 4;
        L->w = (watcher_t*)realloc(L->w, sizeof(watcher_t) * (size_t)L->cap);
    }
    L->w[L->n].cref = cref;
    L->w[L->n].blocker = blocker;
    L->n++;
}

static solver_t* solver_new(int n_vars) {
    solver_t* S = (solver_t*)calloc(1, sizeof(solver_t));
    S->n_vars = n_vars;
    S->ok = 1;
    S->assigns = (uint8_t*)malloc((size_t)n_vars + 1);
    memset(S->assigns, LB_UNDEF, (size_t)n_vars + 1);
    S->trail = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n_vars + 1));
    S->watches = (wlist_t*)calloc((size_t)2 * n_vars + 2, sizeof(wlist_t));
    return S;
}

static void solver_free(solver_t* S) {
    if (!S) return;
    for (int i = 0; i < 2 * S->n_vars; i++) free(S->watches[i].w);
    free(S->watches);
    free(S->assigns);
    free(S->trail);
    free(S->arena);
    free(S->clauses);
    free(S);
}

/* Solver.cc:579-585 uncheckedEnqueue */
static inline void enqueue(solver_t* S, int32_t p) {
    S->assigns[VAR(p)] = (uint8_t)SIGN(p); /* lbool(!sign): positive literal -> true(0) */
    S->trail[S->trail_n++] = p;
}

/* Solver.cc:599-662 propagate: two-watched-literal BCP with blockers; returns the
 * conflicting cref or -1.  Counts propagations (dequeues) and pdsat's watch_scans (:654). */
static int32_t propagate(solver_t* S) {
    int32_t confl = -1;
    uint64_t num_props = 0;
    while (S->qhead < S->trail_n) {
        int32_t p = S->trail[S->qhead++];
        wlist_t* ws = &S->watches[p];
        int32_t i = 0, j = 0, end = ws->n;
        num_props++;
        while (i != end) {
            int32_t blocker = ws->w[i].blocker;
            if (lit_value(S, blocker) == LB_TRUE) {
                ws->w[j++] = ws->w[i++];
                continue;
            }
            int32_t cr = ws->w[i].cref;
            int32_t* c = S->arena + cr + 1;
            int32_t size = S->arena[cr];
            int32_t false_lit = NEG(p);
            if (c[0] == false_lit) {
                c[0] = c[1];
                c[1] = false_lit;
            }
            i++;
            int32_t first = c[0];
            watcher_t w;
            w.cref = cr;
            w.blocker = first;
            if (first != blocker && lit_value(S, first) == LB_TRUE) {
                ws->w[j++] = w;
                continue; /* note: `continue` skips the watch_scans++ at :654, as in the reference */
            }
            int found = 0;
            for (int32_t k = 2; k < size; k++) {
                if (lit_value(S, c[k]) != LB_FALSE) {
                    c[1] = c[k];
                    c[k] = false_lit;
                    wl_push(&S->watches[NEG(c[1])], w.cref, w.blocker);
                    ws = &S->watches[p]; /* realloc of another list cannot move this one, but be explicit */
                    found = 1;
                    break;
                }
            }
            if (!found) {
                ws->w[j++] = w;
                if (lit_value(S, first) == LB_FALSE) {
                    confl = cr;
                    S->qhead = S->trail_n;
                    while (i < end) ws->w[j++] = ws->w[i++];
                } else {
                    enqueue(S, first);
                }
            }
            S->watch_scans++;
        }
        ws->n -= (i - j);
    }
    S->propagations += num_props;
    return confl;
}

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return x < y ? -1 : (x > y);
}

/* Solver.cc:279-307 addClause_ (+ attachClause :311-317): sort, drop duplicate and
 * level-0-false literals, skip tautologies and level-0-satisfied clauses; unit ->
 * enqueue + propagate; empty -> ok = false. */
static int add_clause(solver_t* S, int32_t* ps, int n) {
    if (!S->ok) return 0;
    qsort(ps, (size_t)n, sizeof(int32_t), cmp_i32);
    int32_t p = -2;
    int i, j;
    for (i = j = 0; i < n; i++) {
        int v = lit_value(S, ps[i]);
        if (v == LB_TRUE || ps[i] == NEG(p)) return 1;
        if (v != LB_FALSE && ps[i] != p) ps[j++] = p = ps[i];
    }
    n = j;
    if (n == 0) return S->ok = 0;
    if (n == 1) {
        enqueue(S, ps[0]);
        return S->ok = (propagate(S) == -1);
    }
    if (S->arena_n + n + 1 > S->arena_cap) {
        S->arena_cap = S->arena_cap ? S->arena_cap * 2 : 1 << 16;
        while (S->arena_n + n + 1 > S->arena_cap) S->arena_cap *= 2;
        S->arena = (int32_t*)realloc(S->arena, sizeof(int32_t) * (size_t)S->arena_cap);
    }
    int32_t cr = (int32_t)S->arena_n;
    S->arena[cr] = n;
    memcpy(S->arena + cr + 1, ps, sizeof(int32_t) * (size_t)n);
    S->arena_n += n + 1;
    if (S->clauses_n == S->clauses_cap) {
        S->clauses_cap = S->clauses_cap ? S->clauses_cap * 2 : 1024;
        S->clauses = (int32_t*)realloc(S->clauses, sizeof(int32_t) * (size_t)S->clauses_cap);
    }
    S->clauses[S->clauses_n++] = cr;
    wl_push(&S->watches[NEG(ps[0])], cr, ps[1]);
    wl_push(&S->watches[NEG(ps[1])], cr, ps[0]);
    return 1;
}

/* Solver.h:478-492 addProblem */
static void add_problem(solver_t* S, const problem_t* P) {
    int32_t tmp[4096];
    int32_t* big = NULL;
    for (int64_t c = 0; c < P->in_n_clauses; c++) {
        int n = (int)(P->in_offs[c + 1] - P->in_offs[c]);
        int32_t* buf = tmp;
        if (n > 4096) buf = big = (int32_t*)realloc(big, sizeof(int32_t) * (size_t)n);
        memcpy(buf, P->in_lits + P->in_offs[c], sizeof(int32_t) * (size_t)n);
        if (!add_clause(S, buf, n)) break;
    }
    free(big);
}

/* ---------------------------------------------------------------- C API */
void* orc_create(int n_vars, int64_t n_clauses, const uint32_t* offsets, const int32_t* lits) {
    handle_t* h = (handle_t*)calloc(1, sizeof(handle_t));
    h->P.n_vars = n_vars;
    h->P.in_n_clauses = n_clauses;
    h->P.in_offs = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_clauses + 1));
    memcpy(h->P.in_offs, offsets, sizeof(uint32_t) * ((size_t)n_clauses + 1));
    size_t nl = offsets[n_clauses];
    h->P.in_lits = (int32_t*)malloc(sizeof(int32_t) * (nl ? nl : 1));
    memcpy(h->P.in_lits, lits, sizeof(int32_t) * nl);
    return h;
}

void orc_destroy(void* hv) {
    handle_t* h = (handle_t*)hv;
    if (!h) return;
    solver_free(h->persistent);
    free(h->P.in_offs);
    free(h->P.in_lits);
    free(h);
}

static void fill_assigns(const solver_t* S, uint8_t* assigns) {
    if (assigns) memcpy(assigns, S->assigns, (size_t)S->n_vars);
}

/* src_mpi/mpi_predicter.cpp:691-773 solverProgramCalling with evaluation_type "prep"
 * (eval_type 0), "propagation" (1) or "watch_scans" (2) up to the point where BCP alone
 * decides; the prep fast exit is Solver.cc:874-877.  For eval_type != 0 a sample BCP does
 * not decide would need CDCL (out of scope): it is reported as ret = 2 (l_Undef) too. */
int orc_predict_sample(void* hv, const int32_t* assumps, int n, int eval_type, orc_result* r,
                       uint8_t* assigns) {
    handle_t* h = (handle_t*)hv;
    (void)eval_type;
    solver_t* S = solver_new(h->P.n_vars);
    add_problem(S, &h->P);                     /* :705 */
    for (int i = 0; i < n; i++) {              /* :706-707 one unit clause per assumption */
        int32_t u = assumps[i];
        add_clause(S, &u, 1);
    }
    memset(r, 0, sizeof(*r));
    uint64_t decisions = 0;
    int fast_exit = 0;
    int ret;
    if (!S->ok) {
        ret = LB_FALSE;                        /* Solver.cc:979 */
    } else if (propagate(S) != -1) {
        ret = LB_FALSE;
    } else {
        decisions = 1;                         /* Solver.cc:866-877 first decision */
        if (S->trail_n == S->n_vars)
            ret = LB_TRUE;                     /* pickBranchLit == lit_Undef: model found */
        else {
            fast_exit = 1;                     /* isNonPrepFastExit */
            ret = LB_UNDEF;
        }
    }
    r->ret = ret;
    r->prep = (decisions <= 1 && !fast_exit) ? 1 : 0;   /* :745-746 */
    r->sat = (ret == LB_TRUE && !fast_exit) ? 1 : 0;      /* :755 */
    r->conflict = ret == LB_FALSE;
    r->trail_size = S->trail_n;
    r->full = (!r->conflict && S->trail_n == S->n_vars) ? 1 : 0;
    r->propagations = S->propagations;
    r->watch_scans = S->watch_scans;
    r->decisions = decisions;
    fill_assigns(S, assigns);
    solver_free(S);
    return 0;
}

static solver_t* persistent(handle_t* h) {
    if (!h->persistent) {
        h->persistent = solver_new(h->P.n_vars);
        add_problem(h->persistent, &h->P);
    }
    return h->persistent;
}

int orc_ok(void* hv) { return persistent((handle_t*)hv)->ok; }

/* src_mpi/mpi_solver.cpp:199-200: known_dummy added as unit clauses */
int orc_add_units(void* hv, const int32_t* lits, int n) {
    solver_t* S = persistent((handle_t*)hv);
    for (int i = 0; i < n; i++) {
        int32_t u = lits[i];
        add_clause(S, &u, 1);
    }
    return S->ok ? 0 : 1;
}

/* BCP of the assumption literals on the persistent DB, then undo (Solver.cc:599-662 for
 * the propagation; the undo mirrors cancelUntil(0), Solver.cc:559-573: assignments are
 * cleared, watch lists keep their new order). */
int orc_bcp(void* hv, const int32_t* assumps, int n, orc_result* r, uint8_t* assigns) {
    handle_t* h = (handle_t*)hv;
    solver_t* S = persistent(h);
    memset(r, 0, sizeof(*r));
    if (!S->ok) {
        r->conflict = 1;
        return 0;
    }
    int32_t base = S->trail_n;
    uint64_t p0 = S->propagations, w0 = S->watch_scans;
    int conflict = 0;
    for (int i = 0; i < n && !conflict; i++) {
        int v = lit_value(S, assumps[i]);
        if (v == LB_TRUE) continue;
        if (v == LB_FALSE) {
            conflict = 1;
            break;
        }
        enqueue(S, assumps[i]);
        if (propagate(S) != -1) conflict = 1;
    }
    r->conflict = conflict;
    r->trail_size = S->trail_n;
    r->full = (!conflict && S->trail_n == S->n_vars) ? 1 : 0;
    r->propagations = S->propagations - p0;
    r->watch_scans = S->watch_scans - w0;
    fill_assigns(S, assigns);
    for (int32_t t = S->trail_n - 1; t >= base; t--) S->assigns[VAR(S->trail[t])] = LB_UNDEF;
    S->trail_n = base;
    S->qhead = base;
    return 0;
}

/* ---------------------------------------------------------------- MT19937
 * The reference draws bits with boost::random::mt19937 seeded by gen.seed(rank)
 * (src_mpi/mpi_predicter.cpp:87) through Addit_func::bool_rand
 * (src_common/addit_func.cpp:193-196) = uniform_int_distribution<uint32_t>(0,1)(gen)==0.
 * Boost (un-vendored, unversioned; a VS project hints boost_1_53_0) maps a 32-bit engine
 * onto [0,1] by buckets of 2^31 without rejection, i.e. result = draw >> 31; so
 * bool_rand == ((draw >> 31) == 0).  mt19937 itself is the published Matsumoto-Nishimura
 * generator (identical in boost and std). */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937_t;

static void mt_seed(mt19937_t* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937_t* g) {
    if (g->idx >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
            g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* raw 32-bit draws first..first+n-1 of the stream (Addit_func::uint_rand, :188-191) */
void orc_mt19937_draws(uint32_t seed, uint64_t first, uint64_t n, uint32_t* out) {
    mt19937_t g;
    mt_seed(&g, seed);
    for (uint64_t i = 0; i < first; i++) (void)mt_next(&g);
    for (uint64_t i = 0; i < n; i++) out[i] = mt_next(&g);
}

/* bool_rand results for draws first..first+n-1, one byte each (1 = true = positive lit) */
void orc_bool_rand(uint32_t seed, uint64_t first, uint64_t n, uint8_t* out) {
    mt19937_t g;
    mt_seed(&g, seed);
    for (uint64_t i = 0; i < first; i++) (void)mt_next(&g);
    for (uint64_t i = 0; i < n; i++) out[i] = (uint8_t)((mt_next(&g) >> 31) == 0);
}

/* src_mpi/mpi_predicter.cpp:3236-3242: assumptions of sample `s` in the 1-worker order:
 * sample s uses draws s*k .. s*k+k-1; true bit => positive literal of set var. */
void orc_sample_assumptions(uint32_t seed, const int32_t* set_vars /*1-based*/, int k, uint64_t first_sample,
                            uint64_t n_samples, int32_t* out /* n_samples*k MiniSat lits */) {
    mt19937_t g;
    mt_seed(&g, seed);
    for (uint64_t i = 0; i < first_sample * (uint64_t)k; i++) (void)mt_next(&g);
    for (uint64_t s = 0; s < n_samples; s++)
        for (int j = 0; j < k; j++) {
            int b = (mt_next(&g) >> 31) == 0;
            int32_t v = set_vars[j] - 1;
            out[s * (uint64_t)k + j] = 2 * v + (b ? 0 : 1);
        }
}

/* Batch form of the predict path for the bench / full-size parity checks (loops in C instead of
 * one ctypes call per sample).  Samples first_sample .. first_sample+n-1 of the 1-worker order
 * (mpi_predicter.cpp:3236-3242) starting `first_draw` draws into mt19937(seed).
 *   mode 0: BCP only on the persistent DB (orc_bcp)
 *   mode 1: fresh solver + addProblem per sample, eval "prep" (orc_predict_sample)
 * Per sample: bit s of conflict_bits / full_bits, trail[s] (0 for conflict samples);
 * sums[0..5] += n_conflict, n_full, sum trail, sum trail^2, propagations, watch_scans
 * (trail sums over conflict-free samples).  Any output pointer may be NULL. */
void orc_batch_mt(void* hv, int mode, uint32_t seed, const int32_t* set_vars, int k, const int32_t* extra, int n_extra,
                  uint64_t first_draw, uint64_t first_sample, uint64_t n, uint64_t* conflict_bits, uint64_t* full_bits,
                  uint32_t* trail, uint64_t* sums) {
    mt19937_t g;
    mt_seed(&g, seed);
    for (uint64_t i = 0; i < first_draw + first_sample * (uint64_t)k; i++) (void)mt_next(&g);
    int32_t* a = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k + n_extra + 1));
    for (uint64_t s = 0; s < n; s++) {
        for (int j = 0; j < k; j++) {
            int b = (mt_next(&g) >> 31) == 0;
            a[j] = 2 * (set_vars[j] - 1) + (b ? 0 : 1);
        }
        for (int j = 0; j < n_extra; j++) a[k + j] = extra[j];
        orc_result r;
        if (mode == 0) orc_bcp(hv, a, k, &r, NULL);
        else orc_predict_sample(hv, a, k + n_extra, 0, &r, NULL);
        if (conflict_bits && r.conflict) conflict_bits[s >> 6] |= 1ull << (s & 63);
        if (full_bits && r.full) full_bits[s >> 6] |= 1ull << (s & 63);
        if (trail) trail[s] = r.conflict ? 0u : (uint32_t)r.trail_size;
        if (sums) {
            sums[0] += (uint64_t)r.conflict;
            sums[1] += (uint64_t)r.full;
            if (!r.conflict) {
                sums[2] += (uint64_t)r.trail_size;
                sums[3] += (uint64_t)r.trail_size * (uint64_t)r.trail_size;
            }
            sums[4] += r.propagations;
            sums[5] += r.watch_scans;
        }
    }
    free(a);
}

/* Batch form of the solve-mode enumeration (mpi_base.cpp:148-169 order): assignment indices
 * first .. first+n-1 over vars_sorted, BCP only on the persistent DB.  Same outputs as
 * orc_batch_mt. */
void orc_batch_enum(void* hv, const int32_t* vars_sorted, int k, uint64_t first, uint64_t n, uint64_t* conflict_bits,
                    uint64_t* full_bits, uint64_t* sums) {
    int32_t* a = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k + 1));
    for (uint64_t s = 0; s < n; s++) {
        const uint64_t lint = first + s;
        for (int r = 0; r < k; r++) a[r] = 2 * (vars_sorted[r] - 1) + (((lint >> r) & 1) ? 0 : 1);
        orc_result r;
        orc_bcp(hv, a, k, &r, NULL);
        if (conflict_bits && r.conflict) conflict_bits[s >> 6] |= 1ull << (s & 63);
        if (full_bits && r.full) full_bits[s >> 6] |= 1ull << (s & 63);
        if (sums) {
            sums[0] += (uint64_t)r.conflict;
            sums[1] += (uint64_t)r.full;
            sums[4] += r.propagations;
            sums[5] += r.watch_scans;
        }
    }
    free(a);
}

/* ---------------------------------------------------------------- enumeration order
 * src_mpi/mpi_base.cpp:107-178 MakeAssignsFromMasks: bit r of lint <-> r-th variable
 * (ascending var index) of the variate part; set bit => positive literal. */
void orc_enum_assumptions(const int32_t* vars_sorted /*1-based ascending*/, int k, uint64_t lint,
                          int32_t* out /* k MiniSat lits */) {
    for (int r = 0; r < k; r++) {
        int32_t v = vars_sorted[r] - 1;
        out[r] = 2 * v + (((lint >> r) & 1) ? 0 : 1);
    }
}

/* ---------------------------------------------------------------- estimator
 * src_mpi/mpi_predicter.cpp:1976-1996 (GetPredict, te == 0 and eval != prep):
 *   sum (double, ascending j) ; med = sum / n (long double) ;
 *   predict = med / proc_count * pow(2, k)  (long double).
 * Returned as long double through a pointer (ctypes c_longdouble). */
void orc_predict_mean(const double* cost, uint64_t n, int k, int proc_count, double* sum_out,
                      long double* med_out, long double* predict_out) {
    double sum = 0.0;
    for (uint64_t j = 0; j < n; j++) sum += cost[j];
    long double med = sum / (double)n;      /* med_time_arr is vector<long double> (mpi_predicter.h:237) */
    long double pred = med / (double)proc_count;
    long double p2 = 1.0L;
    for (int i = 0; i < k; i++) p2 *= 2.0L; /* pow(2.0, k) is exact */
    pred *= (double)p2;
    *sum_out = sum;
    *med_out = med;
    *predict_out = pred;
}

/* src_mpi/mpi_predicter.cpp:1846-1901 getCurPredictTime, evaluation_type "prep":
 *   solved = #samples with prep flag; skipped unless 100*solved/n > 0.5;
 *   predict = pow(2,k) * 3.0 / (solved/n)   (computed in double, stored long double).
 * Returns HUGE_DOUBLE (1e308, mpi_base.h:43) when the point is skipped. */
long double orc_predict_prep(uint64_t solved, uint64_t n, int k) {
    double percent = (double)solved * 100 / (double)n;
    if (percent <= 0.5) return 1e+308L;
    long double prob = (double)solved / (double)n; /* cur_probability is long double */
    double p2 = 1.0;
    for (int i = 0; i < k; i++) p2 *= 2.0;
    return p2 * 3.0 / prob;
}

/* src_mpi/mpi_predicter.cpp:2219-2224 sample variance (two pass, divisor solved-1) */
double orc_sample_variance(const double* cost, uint64_t n) {
    double med = 0.0;
    for (uint64_t j = 0; j < n; j++) med += cost[j];
    med /= (double)n;
    double var = 0.0;
    for (uint64_t j = 0; j < n; j++) var += (cost[j] - med) * (cost[j] - med);
    return var / (double)(n - 1);
}
