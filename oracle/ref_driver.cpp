// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI driver around the UNMODIFIED reference MiniSat (compiled from the
// sources where they lie under /root/reference by oracle/Makefile; output goes
// to oracle/_ref/libpdsat_ref.so).  It replays, sample by sample, what the
// reference does on the hot path:
//
//   orc_predict_sample  = MPI_Predicter::solverProgramCalling
//                         (src_mpi/mpi_predicter.cpp:691-773) with
//                         evaluation_type "prep" / "propagation" / "watch_scans"
//   orc_bcp             = the BCP part alone on a persistent clause DB
//                         (Solver::propagate, src_common/minisat/core/Solver.cc:599-662)
//   orc_solve_assumps   = MPI_Solver::SolverRun's inner call
//                         S->solve(dummy_vec[i]) (src_mpi/mpi_solver.cpp:218-245)
//   orc_rc2             = Solver::gen_valid_assumptions_rc2 (Solver.cc:1293-1435)
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  Nothing under pdsat_b200/ does.
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "core/Solver.h"

using namespace Minisat;

namespace {

// Protected members (trail, ok, propagate, ...) are reachable from a subclass.
struct PS : public Solver {
    bool okay_() const { return ok; }
    int trail_size() const { return trail.size(); }
    Lit trail_at(int i) const { return trail[i]; }
    // BCP of a set of assumption literals at decision level 1 on a persistent
    // DB; returns 1 on conflict.  Order of enqueue = order given, with a full
    // propagate() between literals, like repeated addClause(Lit) would do
    // (mpi_predicter.cpp:706-707) but undoable.
    int bcp(const int32_t* lits, int n, uint64_t* props, uint64_t* scans) {
        uint64_t p0 = propagations;
        long long w0 = watch_scans;
        int conflict = 0;
        newDecisionLevel();
        for (int i = 0; i < n && !conflict; i++) {
            Lit p;
            p.x = lits[i];
            if (value(p) == l_True) continue;
            if (value(p) == l_False) {
                conflict = 1;
                break;
            }
            uncheckedEnqueue(p);
            if (propagate() != CRef_Undef) conflict = 1;
        }
        *props = propagations - p0;
        *scans = (uint64_t)(watch_scans - w0);
        return conflict;
    }
    void undo() { cancelUntil(0); }
};

struct Handle {
    Problem cnf;
    int n_vars;
    PS* persistent;  // for orc_bcp / orc_solve_assumps
    Handle() : n_vars(0), persistent(0) {}
};

void fill_assigns(const Solver& S, int n_vars, uint8_t* assigns) {
    if (!assigns) return;
    for (int v = 0; v < n_vars; v++) {
        lbool val = v < S.nVars() ? S.value((Var)v) : l_Undef;
        assigns[v] = val == l_True ? 0 : (val == l_False ? 1 : 2);
    }
}

}  // namespace

extern "C" {

struct orc_result {
    int32_t conflict;      // 1 = BCP (or level-0 load) hit a conflict
    int32_t full;          // 1 = every variable assigned, no conflict
    int32_t sat;           // process_sat_count contribution (ret==l_True && !fastExit)
    int32_t prep;          // isSolvedOnPreprocessing
    int32_t ret;           // 0 l_True, 1 l_False, 2 l_Undef
    int32_t trail_size;    // literals on the trail at level 0 after the call
    uint64_t propagations; // S->propagations
    uint64_t watch_scans;  // S->watch_scans
    uint64_t decisions;
    uint64_t conflicts;
};

void* orc_create(int n_vars, int64_t n_clauses, const uint32_t* offsets, const int32_t* lits) {
    Handle* h = new Handle();
    h->n_vars = n_vars;
    for (int64_t c = 0; c < n_clauses; c++) {
        Disjunct* d = new Disjunct;
        for (uint32_t i = offsets[c]; i < offsets[c + 1]; i++) {
            Lit p;
            p.x = lits[i];
            d->push(p);
        }
        h->cnf.push_back(d);
    }
    return h;
}

void orc_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    if (!h) return;
    for (size_t i = 0; i < h->cnf.size(); i++) delete h->cnf[i];
    delete h->persistent;
    delete h;
}

// mpi_predicter.cpp:691-773, eval_type: 0 prep, 1 propagation, 2 watch_scans
int orc_predict_sample(void* hv, const int32_t* assumps, int n, int eval_type, orc_result* r,
                       uint8_t* assigns) {
    Handle* h = (Handle*)hv;
    PS* S = new PS();
    S->addProblem(h->cnf);
    while (S->nVars() < h->n_vars) S->newVar();
    for (int i = 0; i < n; i++) {
        Lit p;
        p.x = assumps[i];
        S->addClause(p);
    }
    S->pdsat_verbosity = 0;
    S->isPredict = true;
    S->rank = 1;
    S->core_len = 0;
    S->start_activity = 0;
    S->evaluation_type = eval_type == 0 ? "prep" : (eval_type == 1 ? "propagation" : "watch_scans");
    uint64_t prev_decisions = S->decisions;
    S->resetVarActivity();
    bool ok_before = S->okay_();
    lbool ret = S->solve();
    memset(r, 0, sizeof(*r));
    r->ret = ret == l_True ? 0 : (ret == l_False ? 1 : 2);
    r->prep = ((S->decisions <= prev_decisions + 1) && !S->isNonPrepFastExit) ? 1 : 0;
    r->sat = (ret == l_True && !S->isNonPrepFastExit) ? 1 : 0;
    r->conflict = ok_before ? (ret == l_False ? 1 : 0) : 1;
    r->trail_size = S->trail_size();
    r->full = (!r->conflict && S->trail_size() == S->nVars()) ? 1 : 0;
    r->propagations = S->propagations;
    r->watch_scans = (uint64_t)S->watch_scans;
    r->decisions = S->decisions;
    r->conflicts = S->conflicts;
    if (ret == l_True && assigns) {
        for (int v = 0; v < h->n_vars; v++)
            assigns[v] = S->model[v] == l_True ? 0 : (S->model[v] == l_False ? 1 : 2);
    } else {
        fill_assigns(*S, h->n_vars, assigns);
    }
    delete S;
    return 0;
}

static PS* persistent(Handle* h) {
    if (!h->persistent) {
        h->persistent = new PS();
        h->persistent->addProblem(h->cnf);
        while (h->persistent->nVars() < h->n_vars) h->persistent->newVar();
        h->persistent->pdsat_verbosity = 0;
        h->persistent->isPredict = false;
        h->persistent->rank = 1;
        h->persistent->core_len = 0;
        h->persistent->start_activity = 0;
    }
    return h->persistent;
}

int orc_ok(void* hv) { return persistent((Handle*)hv)->okay_() ? 1 : 0; }

// add unit clauses permanently to the persistent solver (mpi_solver.cpp:199-200)
int orc_add_units(void* hv, const int32_t* lits, int n) {
    PS* S = persistent((Handle*)hv);
    for (int i = 0; i < n; i++) {
        Lit p;
        p.x = lits[i];
        S->addClause(p);
    }
    return S->okay_() ? 0 : 1;
}

// BCP only, persistent DB.  assigns = state at the BCP fixpoint (or at the
// moment of conflict), before undoing.
int orc_bcp(void* hv, const int32_t* assumps, int n, orc_result* r, uint8_t* assigns) {
    Handle* h = (Handle*)hv;
    PS* S = persistent(h);
    memset(r, 0, sizeof(*r));
    if (!S->okay_()) {
        r->conflict = 1;
        return 0;
    }
    uint64_t props = 0, scans = 0;
    r->conflict = S->bcp(assumps, n, &props, &scans);
    r->trail_size = S->trail_size();
    r->full = (!r->conflict && S->trail_size() == S->nVars()) ? 1 : 0;
    r->propagations = props;
    r->watch_scans = scans;
    fill_assigns(*S, h->n_vars, assigns);
    S->undo();
    return 0;
}

// Batch form of the predict path (bench / full-size parity): samples in the 1-worker order
// (mpi_predicter.cpp:3236-3242) of std::mt19937(seed) - the same engine as boost::mt19937 -
// with bool_rand == ((draw >> 31) == 0) (addit_func.cpp:193-196, Boost's bucket rule).
//   mode 0: BCP only on the persistent DB;  mode 1: solverProgramCalling, eval "prep"
// (fresh Solver + addProblem per sample; `extra` = literals assumed in every sample).
void orc_batch_mt(void* hv, int mode, uint32_t seed, const int32_t* set_vars, int k, const int32_t* extra, int n_extra,
                  uint64_t first_draw, uint64_t first_sample, uint64_t n, uint64_t* conflict_bits, uint64_t* full_bits,
                  uint32_t* trail, uint64_t* sums) {
    std::mt19937 gen(seed);
    gen.discard(first_draw + first_sample * (uint64_t)k);
    std::vector<int32_t> a((size_t)(k + n_extra));
    for (uint64_t s = 0; s < n; s++) {
        for (int j = 0; j < k; j++) {
            bool b = (gen() >> 31) == 0;
            a[j] = 2 * (set_vars[j] - 1) + (b ? 0 : 1);
        }
        for (int j = 0; j < n_extra; j++) a[k + j] = extra[j];
        orc_result r;
        if (mode == 0) orc_bcp(hv, a.data(), k, &r, nullptr);
        else orc_predict_sample(hv, a.data(), k + n_extra, 0, &r, nullptr);
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
}

// Batch form of the solve-mode enumeration (mpi_base.cpp:148-169 order), BCP only.
void orc_batch_enum(void* hv, const int32_t* vars_sorted, int k, uint64_t first, uint64_t n, uint64_t* conflict_bits,
                    uint64_t* full_bits, uint64_t* sums) {
    std::vector<int32_t> a((size_t)k);
    for (uint64_t s = 0; s < n; s++) {
        const uint64_t lint = first + s;
        for (int r = 0; r < k; r++) a[r] = 2 * (vars_sorted[r] - 1) + (((lint >> r) & 1) ? 0 : 1);
        orc_result r;
        orc_bcp(hv, a.data(), k, &r, nullptr);
        if (conflict_bits && r.conflict) conflict_bits[s >> 6] |= 1ull << (s & 63);
        if (full_bits && r.full) full_bits[s >> 6] |= 1ull << (s & 63);
        if (sums) {
            sums[0] += (uint64_t)r.conflict;
            sums[1] += (uint64_t)r.full;
            sums[4] += r.propagations;
            sums[5] += r.watch_scans;
        }
    }
}

// mpi_solver.cpp:218-245: incremental solve under assumptions + classification
// state: 0 Solved, 1 SolvedOnPreprocessing, 2 Interrupted
int orc_solve_assumps(void* hv, const int32_t* assumps, int n, int max_nof_restarts, orc_result* r,
                      int* state, uint8_t* model) {
    Handle* h = (Handle*)hv;
    PS* S = persistent(h);
    memset(r, 0, sizeof(*r));
    vec<Lit> dummy;
    for (int i = 0; i < n; i++) {
        Lit p;
        p.x = assumps[i];
        dummy.push(p);
    }
    S->max_nof_restarts = max_nof_restarts;
    uint64_t prev_starts = S->starts, prev_conflicts = S->conflicts;
    lbool ret = S->solve(dummy);
    S->watch_scans = 0;
    r->ret = ret == l_True ? 0 : (ret == l_False ? 1 : 2);
    if (ret == l_Undef)
        *state = 2;
    else if ((S->starts - prev_starts <= 1) && (S->conflicts == prev_conflicts))
        *state = 1;
    else
        *state = 0;
    r->sat = ret == l_True;
    r->conflict = ret == l_False;
    r->decisions = S->decisions;
    r->conflicts = S->conflicts;
    r->propagations = S->propagations;
    if (ret == l_True && model)
        for (int v = 0; v < h->n_vars; v++) model[v] = S->model[v] == l_True ? 0 : 1;
    return 0;
}

// Solver.cc:1293-1435.  d_set 1-based vars; start[i] in {0,1}.  Survivors are
// written row-major into out (cap rows of k ints); returns the number of rows
// written; *valid = interval_nonprepr_number.
int64_t orc_rc2(void* hv, const int32_t* d_set, const int32_t* start, int k, uint64_t diapason_size,
                uint64_t cap, uint64_t* valid, int32_t* out) {
    Handle* h = (Handle*)hv;
    PS* S = persistent(h);
    std::vector<int> ds(d_set, d_set + k), st(start, start + k);
    std::vector<std::vector<int> > voa;
    unsigned long long nonprep = 0;
    S->gen_valid_assumptions_rc2(ds, st, diapason_size, cap, nonprep, voa);
    S->undo();
    *valid = nonprep;
    for (size_t i = 0; i < voa.size(); i++)
        for (int j = 0; j < k; j++) out[i * k + j] = voa[i][j];
    return (int64_t)voa.size();
}

}  // extern "C"
