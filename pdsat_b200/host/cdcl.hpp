// cdcl.hpp — a compact conflict-driven clause-learning SAT solver for the host side of the hand-off.
//
// The device decides what unit propagation alone can decide; a sample (predict mode) or an
// assignment (solve mode) that BCP leaves open - no conflict, not every variable assigned - needs
// search.  In the reference that is the rest of MiniSat's solve() after the level-0 propagation
// (src_common/minisat/core/Solver.cc:775-885 search, :961-1069 solve_; called from
// src_mpi/mpi_predicter.cpp:731 and src_mpi/mpi_solver.cpp:218-226).  This file is an
// independent implementation of the same textbook algorithm (two watched literals with blockers,
// first-UIP learning with local minimisation, VSIDS + phase saving, Luby restarts, activity-based
// learnt-clause reduction, solving under assumptions); it shares no code with the reference and
// is never used as a checker.  Verdicts (SAT / UNSAT) are properties of the formula, so they are
// compared with the reference's in tests/test_cdcl.py; search statistics (decisions,
// propagations, conflicts) depend on the heuristics of the solver at hand and are reported as
// this solver's own.
//
// Literals are MiniSat codes 2*var + sign (sign 1 = negated), as everywhere across the C ABI.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <vector>

namespace pdsat_host {

class Cdcl {
   public:
    enum Result { SAT = 0, UNSAT = 1, UNKNOWN = 2 };
    struct Limits {
        int64_t max_conflicts = -1;  // < 0: none
        int64_t max_restarts = -1;   // <= 0: none; the reference's max_nof_restarts (Solver.cc:1010-1018)
        double max_seconds = -1.0;   // the reference's max_solving_time
    };
    struct Stats {
        uint64_t decisions = 0, propagations = 0, conflicts = 0, restarts = 0, learnt = 0, watch_scans = 0;
    };

    Stats stats;

    bool load(int32_t n_vars, int64_t n_clauses, const uint32_t* offsets, const int32_t* lits) {
        nv_ = n_vars;
        val_.assign((size_t)2 * nv_, 0);
        level_.assign((size_t)nv_, 0);
        reason_.assign((size_t)nv_, NO_REASON);
        activity_.assign((size_t)nv_, 0.0);
        phase_.assign((size_t)nv_, 1);  // first try "false", like MiniSat's default polarity
        seen_.assign((size_t)nv_, 0);
        heap_pos_.assign((size_t)nv_, -1);
        watches_.assign((size_t)2 * nv_, {});
        arena_.clear();
        clauses_.clear();
        learnts_.clear();
        trail_.clear();
        lim_.clear();
        qhead_ = 0;
        ok_ = true;
        heap_.clear();
        for (int v = 0; v < nv_; v++) heap_insert(v);
        std::vector<int32_t> tmp;
        for (int64_t c = 0; c < n_clauses && ok_; c++) {
            tmp.assign(lits + offsets[c], lits + offsets[c + 1]);
            add_clause(tmp);
        }
        n_original_ = (int64_t)clauses_.size();
        return ok_;
    }

    bool ok() const { return ok_; }
    int32_t n_vars() const { return nv_; }

    // permanent unit (solve mode's known_dummy, the keystream of a cipher instance)
    bool add_unit(int32_t lit) {
        cancel_until(0);
        std::vector<int32_t> u{lit};
        add_clause(u);
        return ok_;
    }

    // model[v] in {0, 1} (1 = true) after SAT
    const std::vector<uint8_t>& model() const { return model_; }

    Result solve(const int32_t* assumps, int n, const Limits& lim) {
        cancel_until(0);
        if (!ok_) return UNSAT;
        assumps_.assign(assumps, assumps + n);
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t conflicts0 = stats.conflicts;
        max_learnts_ = std::max<double>((double)n_original_ / 3.0, 2000.0);
        int64_t restarts = 0;
        Result res = UNKNOWN;
        while (res == UNKNOWN) {
            if (lim.max_restarts > 0 && restarts >= lim.max_restarts) break;
            const int64_t budget = (int64_t)(luby(2.0, (int)restarts) * 100.0);
            res = search(budget, lim, conflicts0, t0);
            restarts++;
            stats.restarts++;
            if (res == UNKNOWN && limit_hit_) break;
        }
        if (res == SAT) {
            model_.resize((size_t)nv_);
            for (int v = 0; v < nv_; v++) model_[v] = val_[2 * v] == 1 ? 1 : 0;
        }
        cancel_until(0);
        return res;
    }

   private:
    static constexpr int32_t NO_REASON = -1;
    struct Watcher {
        int32_t cref, blocker;
    };
    // clause layout in the arena: [size][learnt flag][activity as float bits][lits...]
    static constexpr int HDR = 3;

    int32_t nv_ = 0;
    std::vector<uint8_t> val_;  // per literal: 0 undef, 1 true, 2 false
    std::vector<int32_t> level_, reason_;
    std::vector<double> activity_;
    std::vector<uint8_t> phase_, seen_;
    std::vector<int32_t> heap_, heap_pos_;
    std::vector<std::vector<Watcher>> watches_;  // watches_[l]: clauses watching literal l
    std::vector<int32_t> arena_;
    std::vector<int32_t> clauses_, learnts_;
    std::vector<int32_t> trail_, lim_;
    std::vector<int32_t> assumps_;
    std::vector<uint8_t> model_;
    size_t qhead_ = 0;
    bool ok_ = true, limit_hit_ = false;
    double var_inc_ = 1.0, cla_inc_ = 1.0, max_learnts_ = 0;
    int64_t n_original_ = 0;

    int decision_level() const { return (int)lim_.size(); }
    int32_t* cl_lits(int32_t cr) { return &arena_[(size_t)cr + HDR]; }
    int32_t cl_size(int32_t cr) const { return arena_[(size_t)cr]; }
    float& cl_act(int32_t cr) { return *reinterpret_cast<float*>(&arena_[(size_t)cr + 2]); }

    static double luby(double y, int x) {
        int size, seq;
        for (size = 1, seq = 0; size < x + 1; seq++, size = 2 * size + 1) {
        }
        while (size - 1 != x) {
            size = (size - 1) >> 1;
            seq--;
            x = x % size;
        }
        double r = 1.0;
        for (int i = 0; i < seq; i++) r *= y;
        return r;
    }

    // ---- variable order: binary max-heap on activity
    bool heap_less(int a, int b) const { return activity_[a] > activity_[b]; }
    void heap_up(int i) {
        int x = heap_[i];
        while (i > 0) {
            int p = (i - 1) >> 1;
            if (!heap_less(x, heap_[p])) break;
            heap_[i] = heap_[p];
            heap_pos_[heap_[i]] = i;
            i = p;
        }
        heap_[i] = x;
        heap_pos_[x] = i;
    }
    void heap_down(int i) {
        int x = heap_[i];
        const int n = (int)heap_.size();
        while (true) {
            int l = 2 * i + 1, r = l + 1;
            if (l >= n) break;
            int c = (r < n && heap_less(heap_[r], heap_[l])) ? r : l;
            if (!heap_less(heap_[c], x)) break;
            heap_[i] = heap_[c];
            heap_pos_[heap_[i]] = i;
            i = c;
        }
        heap_[i] = x;
        heap_pos_[x] = i;
    }
    void heap_insert(int v) {
        if (heap_pos_[v] >= 0) return;
        heap_.push_back(v);
        heap_pos_[v] = (int)heap_.size() - 1;
        heap_up((int)heap_.size() - 1);
    }
    int heap_pop() {
        int x = heap_[0];
        heap_pos_[x] = -1;
        int last = heap_.back();
        heap_.pop_back();
        if (!heap_.empty()) {
            heap_[0] = last;
            heap_pos_[last] = 0;
            heap_down(0);
        }
        return x;
    }
    void bump_var(int v) {
        if ((activity_[v] += var_inc_) > 1e100) {
            for (auto& a : activity_) a *= 1e-100;
            var_inc_ *= 1e-100;
        }
        if (heap_pos_[v] >= 0) heap_up(heap_pos_[v]);
    }
    void bump_clause(int32_t cr) {
        if ((cl_act(cr) += (float)cla_inc_) > 1e20f) {
            for (int32_t l : learnts_) cl_act(l) *= 1e-20f;
            cla_inc_ *= 1e-20;
        }
    }

    // ---- assignment
    void enqueue(int32_t l, int32_t reason) {
        val_[l] = 1;
        val_[l ^ 1] = 2;
        const int v = l >> 1;
        level_[v] = decision_level();
        reason_[v] = reason;
        trail_.push_back(l);
    }
    void cancel_until(int lvl) {
        if (decision_level() <= lvl) return;
        for (size_t i = trail_.size(); i-- > (size_t)lim_[lvl];) {
            const int32_t l = trail_[i];
            const int v = l >> 1;
            val_[l] = val_[l ^ 1] = 0;
            phase_[v] = (uint8_t)(l & 1);
            heap_insert(v);
        }
        trail_.resize((size_t)lim_[lvl]);
        qhead_ = trail_.size();
        lim_.resize((size_t)lvl);
    }

    // ---- clauses
    int32_t alloc_clause(const std::vector<int32_t>& ls, bool learnt) {
        const int32_t cr = (int32_t)arena_.size();
        arena_.push_back((int32_t)ls.size());
        arena_.push_back(learnt ? 1 : 0);
        arena_.push_back(0);
        arena_.insert(arena_.end(), ls.begin(), ls.end());
        return cr;
    }
    void attach(int32_t cr) {
        int32_t* c = cl_lits(cr);
        watches_[c[0]].push_back(Watcher{cr, c[1]});
        watches_[c[1]].push_back(Watcher{cr, c[0]});
    }
    void detach(int32_t cr) {
        int32_t* c = cl_lits(cr);
        for (int k = 0; k < 2; k++) {
            auto& ws = watches_[c[k]];
            for (size_t i = 0; i < ws.size(); i++)
                if (ws[i].cref == cr) {
                    ws[i] = ws.back();
                    ws.pop_back();
                    break;
                }
        }
    }
    // original clause at level 0: sort, drop duplicates / false literals, skip tautologies and
    // satisfied clauses; units are propagated at once
    void add_clause(std::vector<int32_t>& ls) {
        if (!ok_) return;
        std::sort(ls.begin(), ls.end());
        size_t j = 0;
        int32_t prev = -2;
        for (size_t i = 0; i < ls.size(); i++) {
            const int32_t l = ls[i];
            if (val_[l] == 1 || l == (prev ^ 1)) return;  // satisfied / tautology (sorted: l, l^1 adjacent)
            if (val_[l] == 2 || l == prev) continue;
            ls[j++] = prev = l;
        }
        ls.resize(j);
        if (ls.empty()) {
            ok_ = false;
            return;
        }
        if (ls.size() == 1) {
            enqueue(ls[0], NO_REASON);
            if (propagate() != NO_REASON) ok_ = false;
            return;
        }
        const int32_t cr = alloc_clause(ls, false);
        clauses_.push_back(cr);
        attach(cr);
    }

    // ---- unit propagation: returns the conflicting clause or NO_REASON
    int32_t propagate() {
        int32_t confl = NO_REASON;
        while (qhead_ < trail_.size()) {
            const int32_t p = trail_[qhead_++];
            const int32_t fl = p ^ 1;  // this literal just became false
            stats.propagations++;
            auto& ws = watches_[fl];
            size_t i = 0, j = 0;
            const size_t n = ws.size();
            while (i < n) {
                const Watcher w = ws[i++];
                if (val_[w.blocker] == 1) {
                    ws[j++] = w;
                    continue;
                }
                stats.watch_scans++;
                int32_t* c = cl_lits(w.cref);
                if (c[0] == fl) std::swap(c[0], c[1]);  // the false literal sits at c[1]
                const int32_t first = c[0];
                if (first != w.blocker && val_[first] == 1) {
                    ws[j++] = Watcher{w.cref, first};
                    continue;
                }
                const int sz = cl_size(w.cref);
                bool moved = false;
                for (int k = 2; k < sz; k++)
                    if (val_[c[k]] != 2) {
                        c[1] = c[k];
                        c[k] = fl;
                        watches_[c[1]].push_back(Watcher{w.cref, first});
                        moved = true;
                        break;
                    }
                if (moved) continue;
                ws[j++] = Watcher{w.cref, first};
                if (val_[first] == 2) {
                    confl = w.cref;
                    qhead_ = trail_.size();
                    while (i < n) ws[j++] = ws[i++];
                } else {
                    enqueue(first, w.cref);
                }
            }
            ws.resize(j);
        }
        return confl;
    }

    // ---- first-UIP conflict analysis
    void analyze(int32_t confl, std::vector<int32_t>& out, int& bt_level) {
        out.clear();
        out.push_back(-1);
        int path = 0;
        int32_t p = -1;
        size_t idx = trail_.size();
        std::vector<int32_t>& touched = analyze_touched_;
        touched.clear();
        do {
            if (arena_[(size_t)confl + 1]) bump_clause(confl);
            int32_t* c = cl_lits(confl);
            const int sz = cl_size(confl);
            for (int k = (p == -1 ? 0 : 1); k < sz; k++) {
                const int32_t q = c[k];
                const int v = q >> 1;
                if (!seen_[v] && level_[v] > 0) {
                    seen_[v] = 1;
                    touched.push_back(v);
                    bump_var(v);
                    if (level_[v] >= decision_level()) path++;
                    else out.push_back(q);
                }
            }
            while (!seen_[trail_[--idx] >> 1]) {
            }
            p = trail_[idx];
            confl = reason_[p >> 1];
            seen_[p >> 1] = 0;
            path--;
            if (confl != NO_REASON && cl_lits(confl)[0] != p) {
                // keep the implied literal in front (it is, by construction of enqueue from propagate)
                int32_t* c2 = cl_lits(confl);
                const int sz2 = cl_size(confl);
                for (int k = 1; k < sz2; k++)
                    if (c2[k] == p) {
                        std::swap(c2[0], c2[k]);
                        break;
                    }
            }
        } while (path > 0);
        out[0] = p ^ 1;
        // local minimisation: drop a literal whose reason's other literals are all in the clause
        size_t j = 1;
        for (size_t i = 1; i < out.size(); i++) {
            const int v = out[i] >> 1;
            const int32_t r = reason_[v];
            bool redundant = r != NO_REASON;
            if (redundant) {
                int32_t* c = cl_lits(r);
                const int sz = cl_size(r);
                for (int k = 0; k < sz; k++) {
                    const int u = c[k] >> 1;
                    if (u != v && !seen_[u] && level_[u] > 0) {
                        redundant = false;
                        break;
                    }
                }
            }
            if (!redundant) out[j++] = out[i];
        }
        out.resize(j);
        for (int v : touched) seen_[v] = 0;
        bt_level = 0;
        if (out.size() > 1) {
            size_t m = 1;
            for (size_t i = 2; i < out.size(); i++)
                if (level_[out[i] >> 1] > level_[out[m] >> 1]) m = i;
            std::swap(out[1], out[m]);
            bt_level = level_[out[1] >> 1];
        }
    }
    std::vector<int32_t> analyze_touched_;

    bool locked(int32_t cr) {
        const int32_t l0 = cl_lits(cr)[0];
        return val_[l0] == 1 && reason_[l0 >> 1] == cr;
    }
    void reduce_db() {
        std::sort(learnts_.begin(), learnts_.end(), [&](int32_t a, int32_t b) {
            const bool sa = cl_size(a) > 2, sb = cl_size(b) > 2;
            if (sa != sb) return sa;  // binary learnts last (kept)
            return cl_act(a) < cl_act(b);
        });
        size_t j = 0;
        const size_t half = learnts_.size() / 2;
        for (size_t i = 0; i < learnts_.size(); i++) {
            const int32_t cr = learnts_[i];
            if (i < half && cl_size(cr) > 2 && !locked(cr)) detach(cr);  // arena space is not reclaimed
            else learnts_[j++] = cr;
        }
        learnts_.resize(j);
    }

    int32_t pick_branch() {
        while (!heap_.empty()) {
            const int v = heap_pop();
            if (val_[2 * v] == 0) return 2 * v + (int32_t)phase_[v];
        }
        return -1;
    }

    Result search(int64_t budget, const Limits& lim, uint64_t conflicts0, std::chrono::steady_clock::time_point t0) {
        int64_t n_confl = 0;
        std::vector<int32_t> learnt;
        limit_hit_ = false;
        while (true) {
            const int32_t confl = propagate();
            if (confl != NO_REASON) {
                stats.conflicts++;
                n_confl++;
                if (decision_level() == 0) {
                    ok_ = false;
                    return UNSAT;
                }
                if (decision_level() <= (int)assumps_.size()) {
                    // conflict among the assumption levels: does it depend on decisions at all?
                    // analyse normally; a learnt clause that backjumps below an assumption will be
                    // re-propagated and the assumption found false at its level
                }
                int bt;
                analyze(confl, learnt, bt);
                cancel_until(bt);
                if (learnt.size() == 1) {
                    enqueue(learnt[0], NO_REASON);
                } else {
                    const int32_t cr = alloc_clause(learnt, true);
                    learnts_.push_back(cr);
                    attach(cr);
                    bump_clause(cr);
                    enqueue(learnt[0], cr);
                    stats.learnt++;
                }
                var_inc_ *= 1.0 / 0.95;
                cla_inc_ *= 1.0 / 0.999;
                if ((stats.conflicts & 255) == 0 && lim.max_seconds >= 0) {
                    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    if (el > lim.max_seconds) {
                        limit_hit_ = true;
                        cancel_until(0);
                        return UNKNOWN;
                    }
                }
                if (lim.max_conflicts >= 0 && (int64_t)(stats.conflicts - conflicts0) >= lim.max_conflicts) {
                    limit_hit_ = true;
                    cancel_until(0);
                    return UNKNOWN;
                }
            } else {
                if (n_confl >= budget) {
                    cancel_until(0);
                    return UNKNOWN;  // restart
                }
                if ((double)learnts_.size() - (double)trail_.size() >= max_learnts_) {
                    reduce_db();
                    max_learnts_ *= 1.1;
                }
                int32_t next = -1;
                while (decision_level() < (int)assumps_.size()) {
                    const int32_t a = assumps_[(size_t)decision_level()];
                    if (val_[a] == 1) {
                        lim_.push_back((int32_t)trail_.size());  // dummy level
                    } else if (val_[a] == 2) {
                        cancel_until(0);
                        return UNSAT;  // the assumptions are inconsistent with the formula
                    } else {
                        next = a;
                        break;
                    }
                }
                if (next == -1) {
                    stats.decisions++;
                    next = pick_branch();
                    if (next == -1) return SAT;
                }
                lim_.push_back((int32_t)trail_.size());
                enqueue(next, NO_REASON);
            }
        }
    }
};

}  // namespace pdsat_host
