// pd_sat_main.cpp — C++ host driver over the C ABI (libpdsat_b200.so).
//
// Keeps the reference's command surface for the hot path (src_mpi/pd-sat.cpp:14-47, 279-478):
//   pd-sat-b200 <input_cnf> [split_var_count] [-name=value ...]
// predict mode is implied by -deep_predict>0 or -cnf_in_set>0 (pd-sat.cpp:464-470), otherwise
// solve mode.  Flags honoured: -cnf_in_set[_count]= -deep_predict= -proc[_count]= -core[_len]=
// -schema= -eval={prep,propagation} -ts_strategy= -algorithm_type= -beta_value=
// -max_L2_hamming_distance= -interval_predict -max_solving_time= -max_nof_restarts= -solve_all
// -koef= -verb= ; accepted and ignored: -te= -start_activity= -no_increm -skip_tasks=
// -no_first_stage -solver= -predict_to= (MPI dispatch / CDCL tuning of the reference).
// -eval=time and -eval=watch_scans are REFUSED: wall-clock time is not reproducible and
// watch_scans is order dependent.  Default -eval=prep (exact parity with the
// reference's prep estimate); -eval=propagation hands BCP-undecided samples to the host CDCL
// (bounded by -max_conflicts / -max_solving_time / -max_nof_restarts; an interrupted sample marks
// its set STOPPED, as a stopped task does in the reference); -eval=bcp_trail is the BCP-only
// trail-size cost (a deviation from the reference, named as such).
// Additions: -gpus=N (one database per device, sample ranges / index sub-cubes split across them,
// integer sums added on the host), -device=G, -seed=N (the worker generator: seeded ONCE, its
// stream runs on across points as in src_mpi/mpi_predicter.cpp:87, 3236-3242), -max_steps=N,
// -max_models=N, -max_conflicts=N, -threads=N, -interval_size=N, -interval_required=N,
// -exhaustive_small_sets (enumerate all 2^k assignments of a set when the sample count reaches
// 2^k instead of drawing them; the reference always draws).
// Files written (formats of src_mpi/mpi_predicter.cpp:2118-2281, 1772-1807, 1398-1404):
// `predict`, `predict_<n>`, `deep_predict`, `graph_file`, `L2_list`; `known_point` is read if
// present (mpi_base.cpp:285-297); solve mode appends models to `satisfying_assignments`.
//
// Self-tests that need no GPU:  --estimate n n_conflict n_full sum k proc eval ; --parse-only <cnf>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/pdsat_cuda.h"
#include "dimacs.hpp"
#include "predict.hpp"
#include "search.hpp"

using namespace pdsat_host;

struct Flags {
    std::string cnf, schema, eval = "prep";
    uint64_t cnf_in_set = 0;
    int deep_predict = 0, proc = 1, core_len = 0, verb = 0, device = 0, gpus = 1, koef = 4, max_steps = 100;
    int split_var_count = 0, max_models = 64, ts_strategy = 0, algorithm_type = 0, beta_value = 2, threads = 0;
    unsigned max_L2 = 2;
    int64_t max_conflicts = -1, max_restarts = -1;
    double max_solving_time = -1;
    uint64_t seed = 1, interval_size = 100000, interval_required = 100;
    bool solve_all = false, interval_predict = false, exhaustive_small = false, no_files = false;
};

static bool has_prefix(const std::string& s, const char* p, std::string& val) {
    size_t n = strlen(p);
    if (s.compare(0, n, p) != 0) return false;
    val = s.substr(n);
    return true;
}

static bool parse_flags(int argc, char** argv, Flags& f) {
    int positional = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], v;
        if (a[0] != '-') {
            if (positional == 0) f.cnf = a;
            else if (positional == 1) f.split_var_count = atoi(a.c_str());
            positional++;
        } else if (has_prefix(a, "-cnf_in_set_count=", v) || has_prefix(a, "-cnf_in_set=", v)) f.cnf_in_set = strtoull(v.c_str(), 0, 10);
        else if (has_prefix(a, "-deep_predict=", v)) f.deep_predict = atoi(v.c_str());
        else if (has_prefix(a, "-proc_count=", v) || has_prefix(a, "-proc=", v)) f.proc = atoi(v.c_str());
        else if (has_prefix(a, "-core_len=", v) || has_prefix(a, "-core=", v)) f.core_len = atoi(v.c_str());
        else if (has_prefix(a, "-schema=", v)) f.schema = v;
        else if (has_prefix(a, "-eval=", v)) f.eval = v;
        else if (has_prefix(a, "-verb=", v)) f.verb = atoi(v.c_str());
        else if (has_prefix(a, "-koef=", v)) f.koef = atoi(v.c_str());
        else if (has_prefix(a, "-device=", v)) f.device = atoi(v.c_str());
        else if (has_prefix(a, "-gpus=", v)) f.gpus = atoi(v.c_str());
        else if (has_prefix(a, "-seed=", v)) f.seed = strtoull(v.c_str(), 0, 10);
        else if (has_prefix(a, "-max_steps=", v)) f.max_steps = atoi(v.c_str());
        else if (has_prefix(a, "-max_models=", v)) f.max_models = atoi(v.c_str());
        else if (has_prefix(a, "-ts_strategy=", v)) f.ts_strategy = atoi(v.c_str());
        else if (has_prefix(a, "-algorithm_type=", v)) f.algorithm_type = atoi(v.c_str());
        else if (has_prefix(a, "-beta_value=", v)) f.beta_value = atoi(v.c_str());
        else if (has_prefix(a, "-max_L2_hamming_distance=", v)) f.max_L2 = (unsigned)atoi(v.c_str());
        else if (has_prefix(a, "-max_solving_time=", v)) f.max_solving_time = atof(v.c_str());
        else if (has_prefix(a, "-max_nof_restarts=", v)) f.max_restarts = atoll(v.c_str());
        else if (has_prefix(a, "-max_conflicts=", v)) f.max_conflicts = atoll(v.c_str());
        else if (has_prefix(a, "-threads=", v)) f.threads = atoi(v.c_str());
        else if (has_prefix(a, "-interval_size=", v)) f.interval_size = strtoull(v.c_str(), 0, 10);
        else if (has_prefix(a, "-interval_required=", v)) f.interval_required = strtoull(v.c_str(), 0, 10);
        else if (a == "-interval_predict") f.interval_predict = true;
        else if (a == "-exhaustive_small_sets") f.exhaustive_small = true;
        else if (a == "-no_files") f.no_files = true;
        else if (a == "-solve_all") f.solve_all = true;
        else if (has_prefix(a, "-te=", v) || has_prefix(a, "-start_activity=", v) || a == "-no_increm" || has_prefix(a, "-skip_tasks=", v) ||
                 a == "-no_first_stage" || has_prefix(a, "-solver=", v) || has_prefix(a, "-s=", v) || has_prefix(a, "-predict_to=", v) ||
                 has_prefix(a, "-integer_variables=", v) || a == "-collect_interrupted" || a == "-conseq") {
            if (f.verb) std::cerr << "note: " << a << " concerns the reference's MPI dispatch / CDCL tuning, ignored" << std::endl;
        } else
            std::cerr << "unknown parameter " << a << std::endl;  // pd-sat.cpp: warning only
    }
    return !f.cnf.empty();
}

static std::string set_to_string(const std::vector<int32_t>& s) {
    std::ostringstream o;
    for (size_t i = 0; i < s.size(); i++) o << (i ? " " : "") << s[i];
    return o.str();
}

static int fail_pdsat(const char* what) {
    std::cerr << what << ": " << pdsat_last_error() << std::endl;
    return 1;
}

static bool initial_set(const Flags& f, const Cnf& cnf, int core_len, std::vector<int32_t>& set) {
    std::ifstream kp("known_point");  // mpi_base.cpp:285-297
    if (kp.is_open()) {
        int v;
        while (kp >> v) set.push_back(v);
        std::sort(set.begin(), set.end());
        return !set.empty();
    }
    if (!cnf.var_set.empty() && f.schema.empty() && f.split_var_count == 0) {
        set = cnf.var_set;
        std::sort(set.begin(), set.end());
        return true;
    }
    unsigned count = f.split_var_count > 0 ? (unsigned)f.split_var_count : (unsigned)core_len;
    return schema_vars(f.schema, count, set) && !set.empty();
}

static pdsat_cdcl_limits limits_of(const Flags& f) {
    pdsat_cdcl_limits L;
    L.max_conflicts = f.max_conflicts;
    L.max_restarts = f.max_restarts;
    L.max_seconds = f.max_solving_time;
    L.n_threads = f.threads;
    return L;
}

// One neighbourhood on all devices: every point's sample range is cut into one slice per device
// (same mt19937 stream, reached by jump-ahead), the launches of all devices are in flight
// together, the integer accumulators are added on the host (the cost "all-reduce" of a
// single-process host).
struct Evaluator {
    const Flags& f;
    std::vector<pdsat_db*> dbs;
    uint64_t draws = 0;  // draws the worker generator has made so far (its stream runs on across points)

    bool operator()(std::vector<PointEval>& pts) {
        const size_t np = pts.size(), nd = dbs.size();
        const bool handoff = f.eval == "propagation";
        std::vector<uint64_t> n_of(np), first_draw(np);
        std::vector<int> mode(np);
        for (size_t i = 0; i < np; i++) {
            const size_t k = pts[i].set.size();
            n_of[i] = samples_for_set(k, f.cnf_in_set);
            mode[i] = (f.exhaustive_small && k <= MAX_STEP_CNF_IN_SET && n_of[i] == (1ull << k)) ? PDSAT_MODE_ENUM : PDSAT_MODE_RANDOM_MT19937;
            first_draw[i] = draws;
            if (mode[i] == PDSAT_MODE_RANDOM_MT19937) draws += n_of[i] * k;
            memset(&pts[i].cost, 0, sizeof(pdsat_cost));
            pts[i].n_interrupted = 0;
        }
        std::vector<pdsat_batch*> batch(nd, nullptr);
        std::vector<std::vector<pdsat_point>> dp(nd);
        std::vector<std::vector<size_t>> owner(nd);
        for (size_t d = 0; d < nd; d++) {
            for (size_t i = 0; i < np; i++) {
                const uint64_t lo = n_of[i] * d / nd, hi = n_of[i] * (d + 1) / nd;
                if (hi <= lo) continue;
                pdsat_point p{};
                p.set_vars = pts[i].set.data();
                p.k = (int32_t)pts[i].set.size();
                p.mode = mode[i];
                p.seed = f.seed;
                p.first_sample = lo;
                p.n_samples = hi - lo;
                p.first_draw = first_draw[i];
                dp[d].push_back(p);
                owner[d].push_back(i);
            }
            if (dp[d].empty()) continue;
            const uint32_t out = handoff ? (PDSAT_OUT_FLAGS | PDSAT_OUT_TRAIL) : 0;
            if (pdsat_batch_create(dbs[d], (int32_t)dp[d].size(), dp[d].data(), out, 0, &batch[d]) != PDSAT_OK) return false;
            if (pdsat_batch_fill(batch[d]) || pdsat_batch_propagate(batch[d])) return false;  // asynchronous
        }
        std::vector<double> sum(np, 0.0), sumsq(np, 0.0);
        pdsat_db_info dbi;
        pdsat_db_get_info(dbs[0], &dbi);
        const int info_base = dbi.n_base_assigned;
        for (size_t d = 0; d < nd; d++) {
            if (!batch[d]) continue;
            std::vector<pdsat_cost> c(dp[d].size());
            if (pdsat_batch_costs(batch[d], c.data()) != PDSAT_OK) return false;
            for (size_t j = 0; j < c.size(); j++) {
                pdsat_cost& t = pts[owner[d][j]].cost;
                const uint64_t* s = (const uint64_t*)&c[j];
                uint64_t* a = (uint64_t*)&t;
                for (int w = 0; w < PDSAT_COST_WORDS; w++) a[w] += s[w];
                if (handoff && c[j].n > c[j].n_conflict + c[j].n_full) {
                    // samples BCP leaves open: the host CDCL decides them; their cost replaces the trail
                    std::vector<uint8_t> status((size_t)c[j].n);
                    std::vector<uint64_t> cost((size_t)c[j].n);
                    pdsat_cdcl_limits L = limits_of(f);
                    pdsat_cdcl_stats st;
                    if (pdsat_batch_handoff(batch[d], (int32_t)j, &L, status.data(), cost.data(), 0, nullptr, nullptr, nullptr, &st) != PDSAT_OK) return false;
                    pts[owner[d][j]].n_interrupted += st.n_unknown;
                    for (size_t s2 = 0; s2 < status.size(); s2++) {
                        sum[owner[d][j]] += (double)cost[s2];
                        sumsq[owner[d][j]] += (double)cost[s2] * (double)cost[s2];
                    }
                } else if (handoff) {
                    // nothing open: BCP trails + (level-0 trail + k) for every BCP conflict
                    const double cc = (double)info_base + (double)dp[d][j].k;
                    sum[owner[d][j]] += (double)c[j].sum_trail + (double)c[j].n_conflict * cc;
                    sumsq[owner[d][j]] += (double)c[j].sumsq_trail + (double)c[j].n_conflict * cc * cc;
                }
            }
            pdsat_batch_destroy(batch[d]);
        }
        for (size_t i = 0; i < np; i++) {
            pts[i].est = estimate(pts[i].cost, pts[i].set.size(), f.eval, f.proc);
            if (handoff) {  // GetPredict on the per-sample costs (BCP trail or host solver's count)
                pts[i].est.sum = sum[i];
                pts[i].est.med = sum[i] / (double)pts[i].cost.n;
                pts[i].est.predict = pts[i].est.med / (double)f.proc;
                pts[i].est.predict *= pow(2.0, (double)pts[i].set.size());
                const double n = (double)pts[i].cost.n;
                pts[i].s2 = n > 1 ? (sumsq[i] - sum[i] * sum[i] / n) / (n - 1) : -1.0;
            } else
                pts[i].s2 = sample_variance(pts[i].cost);
        }
        return true;
    }
};

static int run_predict(const Flags& f, const Cnf& cnf, std::vector<pdsat_db*>& dbs, int core_len) {
    std::vector<int32_t> centre;
    if (!initial_set(f, cnf, core_len, centre)) {
        std::cerr << "cannot build the initial decomposition set" << std::endl;
        return 1;
    }
    SearchParams P;
    if (!cnf.var_set.empty()) P.full_vars = cnf.var_set;  // c var_set: the candidate variables (mpi_base.cpp:546-599)
    else
        for (int v = 1; v <= core_len; v++) P.full_vars.push_back(v);
    std::sort(P.full_vars.begin(), P.full_vars.end());
    P.deep_predict = f.deep_predict;
    P.ts_strategy = f.ts_strategy;
    P.algorithm_type = f.algorithm_type;
    P.beta_value = f.beta_value;
    P.max_L2_hamming_distance = f.max_L2;
    P.seed = (uint32_t)f.seed;
    P.max_steps = f.deep_predict ? f.max_steps : 0;
    P.cnf_in_set_count = f.cnf_in_set;
    P.proc_count = f.proc;
    P.eval = f.eval;
    P.schema = f.schema;
    P.verbosity = f.verb > 1 ? 1 : 0;
    P.write_files = !f.no_files;
    Evaluator ev{f, dbs};
    SearchDriver S(P, [&](std::vector<PointEval>& pts) { return ev(pts); });
    if (!S.run(centre)) {
        std::cerr << "search failed: " << S.error << " " << pdsat_last_error() << std::endl;
        return 1;
    }
    char buf[128];
    snprintf(buf, sizeof buf, "%.21Lg", S.best_predict);
    std::cout << "best_predict " << buf << std::endl;
    std::cout << "best_set_size " << S.best_set.size() << std::endl;
    std::cout << "best_set " << set_to_string(S.best_set) << std::endl;
    std::cout << "points_evaluated " << S.points_evaluated << std::endl;
    std::cout << "steps " << S.steps_done << std::endl;
    std::cout << "records " << S.record_count << std::endl;
    for (size_t i = 0; i < S.centre_history.size(); i++) std::cout << "centre " << i << " : " << set_to_string(S.centre_history[i]) << std::endl;
    return 0;
}

// Interval estimation (src_mpi/mpi_predicter.cpp:3058-3181, estim_type rc2): `size` consecutive
// assignments of the set in binary order with d_set[0] the MOST significant bit (Solver.cc:1380-1423),
// starting at a random corner; BCP on the device separates the assignments solved on
// preprocessing from the survivors; up to `required` survivors are solved by the host CDCL.
static int run_interval(const Flags& f, pdsat_db* db, const std::vector<int32_t>& set) {
    const size_t k = set.size();
    if (k > 62) return 1;
    std::vector<int32_t> rev(set.rbegin(), set.rend());  // bit 0 of the index = last variable of the set
    const uint64_t cube = 1ull << k;
    uint64_t size = f.interval_size < cube ? f.interval_size : cube;
    size &= ~63ull;
    if (size == 0) size = cube;
    // random corner: k bool_rand-like draws of the worker generator (seeded once)
    std::mt19937 gen((uint32_t)f.seed);
    uint64_t start = 0;
    for (size_t i = 0; i < k; i++) start = (start << 1) | (uint64_t)((gen() >> 31) == 0);
    if (start + size > cube) start = cube - size;
    pdsat_point pt{};
    pt.set_vars = rev.data();
    pt.k = (int32_t)k;
    pt.mode = PDSAT_MODE_ENUM;
    pt.first_sample = start;
    pt.n_samples = size;
    pdsat_batch* b = nullptr;
    if (pdsat_batch_create(db, 1, &pt, PDSAT_OUT_FLAGS | PDSAT_OUT_TRAIL, 0, &b) != PDSAT_OK) return fail_pdsat("pdsat_batch_create");
    pdsat_cost c;
    if (pdsat_batch_fill(b) || pdsat_batch_propagate(b) || pdsat_batch_costs(b, &c)) return fail_pdsat("evaluate");
    std::vector<uint64_t> cw((size_t)((size + 63) / 64)), fw(cw.size());
    if (pdsat_batch_flags(b, 0, cw.data(), fw.data())) return fail_pdsat("pdsat_batch_flags");
    pdsat_batch_destroy(b);
    std::vector<int32_t> assumps;
    uint64_t n_surv = 0, n_taken = 0;
    for (uint64_t s = 0; s < size; s++) {
        if ((cw[s >> 6] >> (s & 63)) & 1) continue;
        if ((fw[s >> 6] >> (s & 63)) & 1) continue;
        n_surv++;
        if (n_taken < f.interval_required) {
            const uint64_t lint = start + s;
            for (size_t r = 0; r < k; r++) assumps.push_back(2 * (rev[r] - 1) + (((lint >> r) & 1) ? 0 : 1));
            n_taken++;
        }
    }
    pdsat_cdcl_stats st{};
    std::vector<uint8_t> status((size_t)n_taken);
    std::vector<uint64_t> props((size_t)n_taken);
    pdsat_cdcl_limits L = limits_of(f);
    int64_t nm = 0;
    if (n_taken && pdsat_cdcl_solve_many(db, (int64_t)n_taken, (int32_t)k, assumps.data(), &L, status.data(), props.data(), 0, nullptr, nullptr, &nm, &st))
        return fail_pdsat("pdsat_cdcl_solve_many");
    const uint64_t n_prepr = size - n_surv;
    // mean cost per assignment of the interval, in propagations: BCP's literals for the part solved on
    // preprocessing, the host solver's for the sampled survivors scaled to all survivors
    double surv_mean = n_taken ? (double)st.propagations / (double)n_taken : 0.0;
    pdsat_db_info info;
    pdsat_db_get_info(db, &info);
    const double prepr_cost = (double)(c.sum_trail - (c.n - c.n_conflict) * (uint64_t)info.n_base_assigned) + (double)c.n_conflict * (double)k;
    const double med = (prepr_cost + surv_mean * (double)n_surv) / (double)size;
    std::cout << "interval_start " << start << std::endl << "interval_size " << size << std::endl;
    std::cout << "interval_prepr_number " << n_prepr << std::endl << "interval_nonprepr_number " << n_surv << std::endl;
    std::cout << "survivors_solved " << n_taken << " sat " << st.n_sat << " unsat " << st.n_unsat << " interrupted " << st.n_unknown << std::endl;
    printf("med_sample_cost %.17g\n", med);
    printf("interval_predict %.17g\n", med * pow(2.0, (double)k) / (double)f.proc);
    return 0;
}

static int run_solve(const Flags& f, const Cnf& cnf, std::vector<pdsat_db*>& dbs, int core_len) {
    std::vector<int32_t> set;
    if (!initial_set(f, cnf, core_len, set)) {
        std::cerr << "cannot build the decomposition set" << std::endl;
        return 1;
    }
    const size_t k = set.size();
    if (k > 62) {
        std::cerr << "solve mode: decomposition set too large (" << k << " > 62)" << std::endl;
        return 1;
    }
    if (f.interval_predict) return run_interval(f, dbs[0], set);
    // task split (mpi_solver.cpp:447-473): the top bits of the assignment index pick the device;
    // inside a share the enumeration is BCP-pruned on the device (pdsat_enumerate)
    int top = 0;
    while ((1u << (top + 1)) <= dbs.size() && (size_t)(top + 1) < k) top++;
    const uint64_t shares = 1ull << top;
    uint64_t n_conflict = 0, n_models = 0, n_undecided = 0, done = 0, n_propagated = 0;
    uint64_t solved_cdcl = 0, sat_cdcl = 0, interrupted = 0;
    std::ofstream sat("satisfying_assignments", std::ios::app);
    const uint64_t cap = 1ull << 20;
    std::vector<uint64_t> surv((size_t)cap);
    std::vector<uint8_t> ism((size_t)cap);
    bool found = false;
    for (uint64_t g = 0; g < shares && !(found && !f.solve_all); g++) {
        pdsat_db* db = dbs[(size_t)(g % dbs.size())];
        uint64_t ns = 0;
        pdsat_enum_stats st;
        if (pdsat_enumerate(db, set.data(), (int32_t)k, top, g, f.solve_all ? 0 : 1, cap, surv.data(), ism.data(), &ns, &st) != PDSAT_OK)
            return fail_pdsat("pdsat_enumerate");
        n_conflict += st.n_conflict;
        n_models += st.n_models;
        n_undecided += st.n_undecided;
        n_propagated += st.n_propagated;
        done += st.n_assignments;
        const uint64_t stored = ns < cap ? ns : cap;
        std::vector<int32_t> assumps;
        std::vector<uint64_t> open_idx;
        for (uint64_t i = 0; i < stored; i++) {
            std::vector<int32_t> a(k);
            for (size_t r = 0; r < k; r++) a[r] = 2 * (set[r] - 1) + (((surv[i] >> r) & 1) ? 0 : 1);
            if (ism[i]) {
                // BCP alone completed the assignment: read the model back through a one-sample batch
                std::vector<uint8_t> vals(k);
                for (size_t r = 0; r < k; r++) vals[r] = (uint8_t)((surv[i] >> r) & 1);
                pdsat_point pt{};
                pt.set_vars = set.data();
                pt.k = (int32_t)k;
                pt.mode = PDSAT_MODE_EXPLICIT;
                pt.n_samples = 1;
                pt.explicit_values = vals.data();
                pdsat_batch* b = nullptr;
                std::vector<uint8_t> model((size_t)cnf.n_vars);
                int64_t nf = 0;
                int32_t pof = 0;
                uint64_t sof = 0;
                pdsat_cost c;
                if (pdsat_batch_create(db, 1, &pt, 0, 1, &b) || pdsat_batch_fill(b) || pdsat_batch_propagate(b) || pdsat_batch_costs(b, &c) ||
                    pdsat_batch_models(b, &nf, &pof, &sof, model.data()))
                    return fail_pdsat("model read-back");
                pdsat_batch_destroy(b);
                std::cout << "SAT assignment_index " << surv[i] << std::endl;
                for (int v = 0; v < cnf.n_vars; v++) sat << (int)model[(size_t)v];
                sat << std::endl;
                found = true;
                if (!f.solve_all) break;
            } else {
                open_idx.push_back(surv[i]);
                assumps.insert(assumps.end(), a.begin(), a.end());
            }
        }
        if (!open_idx.empty() && !(found && !f.solve_all)) {
            // MPI_Solver::SolverRun's solve(dummy_vec[i]) for what BCP leaves open (mpi_solver.cpp:218-245)
            std::vector<uint8_t> status(open_idx.size());
            const int mm = f.max_models;
            std::vector<int64_t> mp((size_t)mm);
            std::vector<uint8_t> models((size_t)mm * cnf.n_vars);
            int64_t nm = 0;
            pdsat_cdcl_limits L = limits_of(f);
            pdsat_cdcl_stats cs;
            if (pdsat_cdcl_solve_many(db, (int64_t)open_idx.size(), (int32_t)k, assumps.data(), &L, status.data(), nullptr, mm, mp.data(), models.data(), &nm, &cs))
                return fail_pdsat("pdsat_cdcl_solve_many");
            solved_cdcl += cs.n_sat + cs.n_unsat;
            sat_cdcl += cs.n_sat;
            interrupted += cs.n_unknown;
            const int64_t stored_m = nm < mm ? nm : mm;
            for (int64_t j = 0; j < stored_m; j++) {
                std::cout << "SAT assignment_index " << open_idx[(size_t)mp[(size_t)j]] << std::endl;
                for (int v = 0; v < cnf.n_vars; v++) sat << (int)models[(size_t)j * cnf.n_vars + v];
                sat << std::endl;
                found = true;
            }
        }
    }
    std::cout << "assignments " << done << std::endl;
    std::cout << "solved_on_preprocessing " << (n_conflict + n_models) << std::endl;
    std::cout << "conflicts " << n_conflict << std::endl;
    std::cout << "models " << (n_models + sat_cdcl) << std::endl;
    std::cout << "undecided " << n_undecided << std::endl;
    std::cout << "solved " << solved_cdcl << std::endl;
    std::cout << "interrupted " << interrupted << std::endl;
    std::cout << "samples_propagated " << n_propagated << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 9 && std::string(argv[1]) == "--estimate") {
        pdsat_cost c{};
        c.n = strtoull(argv[2], 0, 10);
        c.n_conflict = strtoull(argv[3], 0, 10);
        c.n_full = strtoull(argv[4], 0, 10);
        c.sum_trail = strtoull(argv[5], 0, 10);
        Estimate e = estimate(c, (size_t)atoi(argv[6]), argv[8], atoi(argv[7]));
        printf("%.17g %.21Lg %.21Lg %" PRIu64 "\n", e.sum, e.med, e.predict, e.solved);
        return 0;
    }
    if (argc >= 3 && std::string(argv[1]) == "--parse-only") {
        Cnf cnf;
        std::string err;
        if (!read_dimacs(argv[2], cnf, err)) {
            std::cerr << err << std::endl;
            return 1;
        }
        printf("%d %" PRId64 " %zu %d %d %d %zu\n", cnf.n_vars, cnf.n_clauses(), cnf.lits.size(), cnf.input_variables,
               cnf.output_variables, cnf.known_bits, cnf.var_set.size());
        return 0;
    }
    if (argc >= 8 && std::string(argv[1]) == "--search-selftest") {
        // the search driver on a synthetic objective (no GPU): pd-sat-b200 --search-selftest n_vars deep ts seed steps init_size
        // predict(set) = 1000 + sum over v in set of ((37 v) mod 11 - 4) + 3 * (|set| mod 5)  (integers: exact everywhere)
        SearchParams P;
        const int n = atoi(argv[2]);
        for (int v = 1; v <= n; v++) P.full_vars.push_back(v);
        P.deep_predict = atoi(argv[3]);
        P.ts_strategy = atoi(argv[4]);
        P.seed = (uint32_t)strtoul(argv[5], 0, 10);
        P.max_steps = atoi(argv[6]);
        P.write_files = false;
        P.min_temperature = 0.5;
        SearchDriver S(P, [&](std::vector<PointEval>& pts) {
            for (auto& q : pts) {
                long long p = 1000 + 3 * (long long)(q.set.size() % 5);
                for (int32_t v : q.set) p += (37 * v) % 11 - 4;
                q.est.predict = (long double)p;
                q.est.sum = (double)p;
                q.cost.n = 1;
            }
            return true;
        });
        std::vector<int32_t> init;
        for (int v = 1; v <= atoi(argv[7]); v++) init.push_back(v);
        if (!S.run(init)) return 1;
        for (size_t i = 0; i < S.centre_history.size(); i++) std::cout << "centre " << i << " : " << set_to_string(S.centre_history[i]) << std::endl;
        std::cout << "best_set " << set_to_string(S.best_set) << std::endl;
        std::cout << "best_predict " << (double)S.best_predict << std::endl;
        std::cout << "records " << S.record_count << std::endl;
        return 0;
    }
    Flags f;
    if (!parse_flags(argc, argv, f)) {
        std::cerr << "usage: pd-sat-b200 <input_cnf> [split_var_count] [-cnf_in_set=N] [-deep_predict=D] [-proc=P] [-core=N]\n"
                     "       [-schema=S] [-eval=prep|propagation] [-ts_strategy=T] [-algorithm_type=A] [-solve_all] [-interval_predict]\n"
                     "       [-gpus=N] [-device=G] [-seed=S] [-verb=V]"
                  << std::endl;
        return 1;
    }
    if (f.eval != "prep" && f.eval != "propagation" && f.eval != "bcp_trail") {
        std::cerr << "-eval=" << f.eval << " is not supported: 'time' is not reproducible and 'watch_scans' depends on the order in which the\n"
                  << "reference visits its watch lists; use -eval=prep (exact), -eval=propagation (BCP trail + host CDCL propagations; give\n"
                  << "-max_conflicts / -max_solving_time / -max_nof_restarts) or -eval=bcp_trail (BCP-only trail sums)" << std::endl;
        return 1;
    }
    Cnf cnf;
    std::string err;
    if (!read_dimacs(f.cnf, cnf, err)) {
        std::cerr << err << std::endl;
        return 1;
    }
    int core_len = f.core_len > 0 ? f.core_len : (cnf.input_variables > 0 ? cnf.input_variables : cnf.n_vars);
    int n_dev = 0;
    pdsat_device_count(&n_dev);
    if (f.gpus < 1) f.gpus = 1;
    if (f.device + f.gpus > n_dev && n_dev > 0) f.gpus = n_dev - f.device > 0 ? n_dev - f.device : 1;
    std::vector<pdsat_db*> dbs;
    for (int d = 0; d < f.gpus; d++) {
        pdsat_db* db = nullptr;
        if (pdsat_upload_cnf(f.device + d, cnf.n_vars, cnf.n_clauses(), cnf.offsets.data(), cnf.lits.data(), &db) != PDSAT_OK)
            return fail_pdsat("pdsat_upload_cnf");
        dbs.push_back(db);
    }
    pdsat_db_info info;
    pdsat_db_get_info(dbs[0], &info);
    if (f.verb)
        std::cout << "vars " << info.n_vars << " live " << info.n_live_vars << " clauses " << info.n_clauses_live << " gates " << info.n_gates << " core_len "
                  << core_len << " gpus " << dbs.size() << (info.status == PDSAT_ERR_UNSAT ? " UNSAT at level 0" : "") << std::endl;
    const bool predict = f.deep_predict > 0 || f.cnf_in_set > 0;  // pd-sat.cpp:464-470
    int rc = predict ? run_predict(f, cnf, dbs, core_len) : run_solve(f, cnf, dbs, core_len);
    for (auto* db : dbs) pdsat_destroy(db);
    return rc;
}
