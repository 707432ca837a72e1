// pd_sat_main.cpp — C++ host driver over the C ABI (libpdsat_b200.so).
//
// Keeps the reference's command surface for the hot path (src_mpi/pd-sat.cpp:14-47, 279-478):
//   pd-sat-b200 <input_cnf> [split_var_count] [-name=value ...]
// predict mode is implied by -deep_predict>0 or -cnf_in_set>0 (pd-sat.cpp:464-470), otherwise
// solve mode.  Flags honoured: -cnf_in_set[_count]= -deep_predict= -proc[_count]= -core[_len]=
// -schema= -eval={prep,propagation} -solve_all -koef= -verb= ; accepted and ignored (CDCL-only):
// -te= -max_solving_time= -max_nof_restarts= -start_activity= -ts_strategy= -no_increm ...
// Additions: -device=N, -seed=N, -max_steps=N, -max_models=N.
// Files written (formats after src_mpi/mpi_predicter.cpp:2118-2281, 1772-1807): `predict`
// (one row per evaluated point), `deep_predict` (search log), `known_point` is read if present
// (mpi_base.cpp:285-297); solve mode appends models to `satisfying_assignments`.
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

using namespace pdsat_host;

struct Flags {
    std::string cnf, schema, eval = "propagation";
    uint64_t cnf_in_set = 0;
    int deep_predict = 0, proc = 1, core_len = 0, verb = 0, device = 0, koef = 4, max_steps = 100;
    int split_var_count = 0, max_models = 64;
    uint64_t seed = 1;
    bool solve_all = false;
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
        else if (has_prefix(a, "-seed=", v)) f.seed = strtoull(v.c_str(), 0, 10);
        else if (has_prefix(a, "-max_steps=", v)) f.max_steps = atoi(v.c_str());
        else if (has_prefix(a, "-max_models=", v)) f.max_models = atoi(v.c_str());
        else if (a == "-solve_all") f.solve_all = true;
        else if (has_prefix(a, "-te=", v) || has_prefix(a, "-max_solving_time=", v) || has_prefix(a, "-max_nof_restarts=", v) ||
                 has_prefix(a, "-start_activity=", v) || has_prefix(a, "-ts_strategy=", v) || a == "-no_increm" ||
                 has_prefix(a, "-algorithm_type=", v) || has_prefix(a, "-skip_tasks=", v) || a == "-no_first_stage") {
            if (f.verb) std::cerr << "note: " << a << " concerns the CDCL path, ignored" << std::endl;
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

static int run_predict(const Flags& f, const Cnf& cnf, pdsat_db* db, int core_len) {
    std::vector<int32_t> centre;
    if (!initial_set(f, cnf, core_len, centre)) {
        std::cerr << "cannot build the initial decomposition set" << std::endl;
        return 1;
    }
    std::ofstream deep("deep_predict", std::ios::app), table("predict");
    table << "size status sum pred med solved conflicts models samples set" << std::endl;
    std::set<std::vector<int32_t>> checked;  // tabu list: never evaluate a point twice
    long double best = HUGE_DOUBLE;
    std::vector<int32_t> best_set = centre;
    uint64_t points_total = 0;
    for (int step = 0; step <= (f.deep_predict ? f.max_steps : 0); step++) {
        std::vector<std::vector<int32_t>> sets;
        if (step == 0) sets.push_back(centre);
        else
            for (auto& s : hamming1(centre, core_len))
                if (!checked.count(s)) sets.push_back(s);
        if (sets.empty()) break;
        std::vector<pdsat_point> pts(sets.size());
        for (size_t i = 0; i < sets.size(); i++) {
            uint64_t n = samples_for_set(sets[i].size(), f.cnf_in_set);
            pts[i].set_vars = sets[i].data();
            pts[i].k = (int32_t)sets[i].size();
            bool all = sets[i].size() <= MAX_STEP_CNF_IN_SET && n == (1ull << sets[i].size());
            pts[i].mode = all ? PDSAT_MODE_ENUM : PDSAT_MODE_RANDOM_MT19937;
            pts[i].seed = f.seed + points_total + i;
            pts[i].first_sample = 0;
            pts[i].n_samples = n;
            pts[i].explicit_values = nullptr;
        }
        std::vector<pdsat_cost> cost(sets.size());
        if (pdsat_eval_batch(db, (int32_t)pts.size(), pts.data(), cost.data()) != PDSAT_OK) return fail_pdsat("pdsat_eval_batch");
        points_total += sets.size();
        long double step_best = HUGE_DOUBLE;
        size_t step_arg = 0;
        for (size_t i = 0; i < sets.size(); i++) {
            checked.insert(sets[i]);
            Estimate e = estimate(cost[i], sets[i].size(), f.eval, f.proc);
            char buf[512];
            snprintf(buf, sizeof buf, "%zu 4 %.17g %.21Lg %.21Lg %" PRIu64 " %" PRIu64 " %" PRIu64 " %" PRIu64, sets[i].size(), e.sum,
                     e.predict, e.med, e.solved, (uint64_t)cost[i].n_conflict, (uint64_t)cost[i].n_full, (uint64_t)cost[i].n);
            table << buf << " " << set_to_string(sets[i]) << std::endl;
            if (e.predict > 0 && e.predict < step_best) {
                step_best = e.predict;
                step_arg = i;
            }
        }
        if (step_best < best) {  // new record point (GetPredict :2004-2030, NewRecordPoint :1687-1808)
            best = step_best;
            best_set = sets[step_arg];
            centre = best_set;
            char buf[128];
            snprintf(buf, sizeof buf, "%.21Lg", best);
            deep << "step " << step << " points " << sets.size() << " record " << buf << " size " << best_set.size() << " set "
                 << set_to_string(best_set) << std::endl;
            if (f.verb) std::cout << "step " << step << " new record " << buf << " |set| " << best_set.size() << std::endl;
        } else {
            deep << "step " << step << " points " << sets.size() << " no improvement" << std::endl;
            if (step > 0) break;  // local minimum of the Hamming-1 neighbourhood
        }
    }
    char buf[128];
    snprintf(buf, sizeof buf, "%.21Lg", best);
    std::cout << "best_predict " << buf << std::endl;
    std::cout << "best_set_size " << best_set.size() << std::endl;
    std::cout << "best_set " << set_to_string(best_set) << std::endl;
    std::cout << "points_evaluated " << points_total << std::endl;
    return 0;
}

static int run_solve(const Flags& f, const Cnf& cnf, pdsat_db* db, int core_len) {
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
    // tasks of at most 2^26 assignments each (the reference splits with part_mask_var_count,
    // mpi_solver.cpp:447-473; here a task is just an index range)
    const uint64_t total = 1ull << k, chunk = 1ull << 26;
    uint64_t n_conflict = 0, n_models = 0, done = 0;
    std::ofstream sat("satisfying_assignments", std::ios::app);
    std::vector<uint8_t> models((size_t)f.max_models * cnf.n_vars);
    std::vector<int32_t> point_of(f.max_models);
    std::vector<uint64_t> sample_of(f.max_models);
    for (uint64_t first = 0; first < total; first += chunk) {
        pdsat_point pt{};
        pt.set_vars = set.data();
        pt.k = (int32_t)k;
        pt.mode = PDSAT_MODE_ENUM;
        pt.first_sample = first;
        pt.n_samples = std::min(chunk, total - first);
        pdsat_batch* b = nullptr;
        if (pdsat_batch_create(db, 1, &pt, 0, f.max_models, &b) != PDSAT_OK) return fail_pdsat("pdsat_batch_create");
        pdsat_cost c;
        if (pdsat_batch_fill(b) || pdsat_batch_propagate(b) || pdsat_batch_costs(b, &c)) return fail_pdsat("evaluate");
        int64_t found = 0;
        if (pdsat_batch_models(b, &found, point_of.data(), sample_of.data(), models.data())) return fail_pdsat("pdsat_batch_models");
        pdsat_batch_destroy(b);
        n_conflict += c.n_conflict;
        n_models += c.n_full;
        done += c.n;
        int64_t stored = found < f.max_models ? found : f.max_models;
        for (int64_t i = 0; i < stored; i++) {
            std::cout << "SAT assignment_index " << (first + sample_of[i]) << std::endl;
            for (int v = 0; v < cnf.n_vars; v++) sat << (int)models[(size_t)i * cnf.n_vars + v];
            sat << std::endl;
        }
        if (found > 0 && !f.solve_all) break;  // stop unless -solve_all (mpi_solver.cpp:285-286)
    }
    std::cout << "assignments " << done << std::endl;
    std::cout << "solved_on_preprocessing " << (n_conflict + n_models) << std::endl;
    std::cout << "conflicts " << n_conflict << std::endl;
    std::cout << "models " << n_models << std::endl;
    std::cout << "undecided " << (done - n_conflict - n_models) << std::endl;
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
    Flags f;
    if (!parse_flags(argc, argv, f)) {
        std::cerr << "usage: pd-sat-b200 <input_cnf> [split_var_count] [-cnf_in_set=N] [-deep_predict=D] [-proc=P] [-core=N]\n"
                     "       [-schema=S] [-eval=prep|propagation] [-solve_all] [-device=G] [-seed=S] [-verb=V]"
                  << std::endl;
        return 1;
    }
    Cnf cnf;
    std::string err;
    if (!read_dimacs(f.cnf, cnf, err)) {
        std::cerr << err << std::endl;
        return 1;
    }
    int core_len = f.core_len > 0 ? f.core_len : (cnf.input_variables > 0 ? cnf.input_variables : cnf.n_vars);
    pdsat_db* db = nullptr;
    if (pdsat_upload_cnf(f.device, cnf.n_vars, cnf.n_clauses(), cnf.offsets.data(), cnf.lits.data(), &db) != PDSAT_OK)
        return fail_pdsat("pdsat_upload_cnf");
    pdsat_db_info info;
    pdsat_db_get_info(db, &info);
    if (f.verb)
        std::cout << "vars " << info.n_vars << " live " << info.n_live_vars << " clauses " << info.n_clauses_live << " core_len "
                  << core_len << (info.status == PDSAT_ERR_UNSAT ? " UNSAT at level 0" : "") << std::endl;
    const bool predict = f.deep_predict > 0 || f.cnf_in_set > 0;  // pd-sat.cpp:464-470
    int rc = predict ? run_predict(f, cnf, db, core_len) : run_solve(f, cnf, db, core_len);
    pdsat_destroy(db);
    return rc;
}
