// handoff.hpp — host threads running the library's own CDCL (../host/cdcl.hpp) on the problems
// BCP leaves open.  One incremental solver per thread, built once per database generation from
// the normalised clauses + the level-0 assignment (fixed assumptions included).
#pragma once
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#include <vector>

#include "../host/cdcl.hpp"
#include "host_db.hpp"

namespace pdsat {

struct CdclPool {
    uint64_t generation = ~0ull;
    std::vector<std::unique_ptr<pdsat_host::Cdcl>> solvers;

    // (re)build `n` solvers for the current database
    void prepare(const HostDb& h, uint64_t gen, int n) {
        if (gen != generation) {
            solvers.clear();
            generation = gen;
        }
        while ((int)solvers.size() < n) {
            std::unique_ptr<pdsat_host::Cdcl> S(new pdsat_host::Cdcl());
            const int64_t ncl = (int64_t)h.norm_off.size() - 1;
            S->load(h.n_vars, ncl, h.norm_off.data(), h.norm_lits.data());
            if (h.base_ok)
                for (int32_t v = 0; v < h.n_vars && S->ok(); v++)
                    if (h.base_assign[v] != LB_UNDEF) S->add_unit(2 * v + (h.base_assign[v] == LB_TRUE ? 0 : 1));
            solvers.push_back(std::move(S));
        }
    }
};

struct SolveManyOut {
    uint64_t n_sat = 0, n_unsat = 0, n_unknown = 0, decisions = 0, propagations = 0, conflicts = 0;
    double seconds = 0;
    int threads = 0;
};

// problems i = 0..n-1: assumptions assumps[i*k .. i*k+k-1]; dynamic distribution over the threads
inline void solve_many(CdclPool& pool, const HostDb& h, uint64_t gen, int64_t n, int32_t k, const int32_t* assumps,
                       const pdsat_host::Cdcl::Limits& lim, int n_threads, uint8_t* status, uint64_t* props, int32_t max_models,
                       int64_t* model_problem, uint8_t* models, int64_t* n_models, SolveManyOut& out) {
    const auto t0 = std::chrono::steady_clock::now();
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((int64_t)T > n) T = (int)(n > 0 ? n : 1);
    pool.prepare(h, gen, T);
    std::atomic<int64_t> next(0), n_mod(0);
    std::vector<SolveManyOut> part((size_t)T);
    auto work = [&](int t) {
        pdsat_host::Cdcl& S = *pool.solvers[(size_t)t];
        SolveManyOut& o = part[(size_t)t];
        const int64_t chunk = 4;
        while (true) {
            const int64_t lo = next.fetch_add(chunk);
            if (lo >= n) break;
            const int64_t hi = lo + chunk < n ? lo + chunk : n;
            for (int64_t i = lo; i < hi; i++) {
                const pdsat_host::Cdcl::Stats before = S.stats;
                const pdsat_host::Cdcl::Result r = h.base_ok ? S.solve(assumps + i * (int64_t)k, k, lim) : pdsat_host::Cdcl::UNSAT;
                status[i] = (uint8_t)r;
                const uint64_t dp = S.stats.propagations - before.propagations;
                if (props) props[i] = dp;
                o.decisions += S.stats.decisions - before.decisions;
                o.propagations += dp;
                o.conflicts += S.stats.conflicts - before.conflicts;
                if (r == pdsat_host::Cdcl::SAT) {
                    o.n_sat++;
                    if (max_models > 0) {
                        const int64_t slot = n_mod.fetch_add(1);
                        if (slot < max_models) {
                            if (model_problem) model_problem[slot] = i;
                            if (models) memcpy(models + (size_t)slot * h.n_vars, S.model().data(), (size_t)h.n_vars);
                        }
                    }
                } else if (r == pdsat_host::Cdcl::UNSAT)
                    o.n_unsat++;
                else
                    o.n_unknown++;
            }
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (auto& o : part) {
        out.n_sat += o.n_sat;
        out.n_unsat += o.n_unsat;
        out.n_unknown += o.n_unknown;
        out.decisions += o.decisions;
        out.propagations += o.propagations;
        out.conflicts += o.conflicts;
    }
    if (n_models) {
        const int64_t m = n_mod.load();
        *n_models = max_models > 0 ? m : (int64_t)out.n_sat;
    }
    out.threads = T;
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace pdsat
