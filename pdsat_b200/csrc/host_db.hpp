// host_db.hpp — host side of the clause-database upload.
//
// Turns the caller's CSR `Problem` into the flat layout the BCP kernel reads:
//   1. addClause_ normalisation (src_common/minisat/core/Solver.cc:279-307): sort
//      literals, drop duplicates, skip tautologies;
//   2. level-0 unit propagation to a fixpoint (what the unit branch of addClause_,
//      Solver.cc:297-299, does incrementally) + the caller's fixed assumptions;
//   3. simplification against that assignment (satisfied clauses dropped, false literals
//      removed) — fixpoint-equivalent to the reference's insertion-time filtering
//      (Solver.cc:288-293): the stored clauses differ, every later BCP result does not;
//   4. compaction of the unassigned ("live") variables to 0..n_live-1 and emission of
//      16-byte clause records + per-literal occurrence lists.
//
// Pure C++ (no CUDA): also used by the CPU-only tests through a host-only handle.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace pdsat {

enum : uint8_t { LB_TRUE = 0, LB_FALSE = 1, LB_UNDEF = 2 };

// unused literal slots hold the code of the pad plane (= 2*n_live, "always false")
static constexpr uint16_t REC_CONT = 0xFFFE;      // slot 7: clause continues in the next record
static constexpr uint16_t REC_INDIRECT = 0xFFFD;  // slot 7 of an inline record: slots 0-1 = index of the chain in rec
static constexpr int REC_SLOTS = 8;           // uint16 slots per 16-byte record

struct HostDb {
    int32_t n_vars = 0;
    int64_t n_clauses_in = 0;
    bool ok = true;                       // false: refuted at level 0
    std::string error;

    // normalised input clauses (MiniSat literals), CSR
    std::vector<uint32_t> norm_off;
    std::vector<int32_t> norm_lits;

    // level-0 state
    std::vector<uint8_t> template_assign; // after the template's own units
    std::vector<uint8_t> base_assign;     // + fixed assumptions
    int32_t n_base_assigned = 0;

    // compact live-variable numbering
    std::vector<int32_t> live_of_var;     // orig var -> live index or -1
    std::vector<int32_t> var_of_live;     // live index -> orig var
    int32_t n_live = 0;

    // device layout
    std::vector<uint16_t> rec;            // n_rec * 8 slots; slot = code of the literal's
                                          // FALSE plane (compact literal ^ 1)
    std::vector<uint32_t> occ_off;        // 2*n_live + 1
    std::vector<uint32_t> occ;            // record index of each clause containing the literal
    std::vector<uint16_t> rec_len;        // clause length, indexed by its first record
    std::vector<uint16_t> occrec;         // per literal (occ_off): the records themselves, inline
    int64_t n_rec = 0;
    int64_t n_clauses_live = 0;
    int64_t n_lits_live = 0;
    int32_t max_clause_len = 0;

    static inline int32_t neg(int32_t l) { return l ^ 1; }
    static inline uint8_t lit_val(const std::vector<uint8_t>& a, int32_t l) {
        uint8_t v = a[l >> 1];
        return v == LB_UNDEF ? LB_UNDEF : (uint8_t)(v ^ (l & 1));
    }

    bool load(int32_t nv, int64_t nc, const uint32_t* offs, const int32_t* lits) {
        n_vars = nv;
        n_clauses_in = nc;
        ok = true;
        norm_off.assign(1, 0);
        norm_lits.clear();
        std::vector<int32_t> tmp;
        for (int64_t c = 0; c < nc; c++) {
            tmp.assign(lits + offs[c], lits + offs[c + 1]);
            for (int32_t l : tmp)
                if (l < 0 || (l >> 1) >= nv) {
                    error = "literal out of range in clause " + std::to_string(c);
                    return false;
                }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            bool taut = false;
            for (size_t i = 0; i + 1 < tmp.size(); i++)
                if (tmp[i + 1] == neg(tmp[i]) && (tmp[i] & 1) == 0) taut = true;
            if (taut) continue;
            if (tmp.empty()) ok = false;  // empty clause (Solver.cc:294-295)
            norm_lits.insert(norm_lits.end(), tmp.begin(), tmp.end());
            norm_off.push_back((uint32_t)norm_lits.size());
        }
        template_assign.assign((size_t)nv, LB_UNDEF);
        if (ok) ok = bcp_level0(template_assign, nullptr, 0);
        return set_fixed(nullptr, 0);
    }

    // level-0 BCP over the normalised clauses from `assign` plus `extra` literals
    bool bcp_level0(std::vector<uint8_t>& assign, const int32_t* extra, int n_extra) {
        const size_t ncl = norm_off.size() - 1;
        // occurrence lists over normalised clauses
        std::vector<uint32_t> ooff((size_t)2 * n_vars + 1, 0);
        for (int32_t l : norm_lits) ooff[(size_t)l + 1]++;
        for (size_t i = 0; i < (size_t)2 * n_vars; i++) ooff[i + 1] += ooff[i];
        std::vector<uint32_t> olist(norm_lits.size());
        {
            std::vector<uint32_t> pos(ooff.begin(), ooff.end() - 1);
            for (size_t c = 0; c < ncl; c++)
                for (uint32_t i = norm_off[c]; i < norm_off[c + 1]; i++) olist[pos[norm_lits[i]]++] = (uint32_t)c;
        }
        std::vector<int32_t> queue;
        auto enqueue = [&](int32_t l) -> bool {
            uint8_t v = lit_val(assign, l);
            if (v == LB_TRUE) return true;
            if (v == LB_FALSE) return false;
            assign[l >> 1] = (uint8_t)(l & 1);
            queue.push_back(l);
            return true;
        };
        // examine one clause; false on conflict
        auto visit = [&](size_t c) -> bool {
            int32_t unit = -1;
            int undef = 0;
            for (uint32_t i = norm_off[c]; i < norm_off[c + 1]; i++) {
                uint8_t v = lit_val(assign, norm_lits[i]);
                if (v == LB_TRUE) return true;
                if (v == LB_UNDEF) {
                    undef++;
                    unit = norm_lits[i];
                }
            }
            if (undef == 0) return false;
            if (undef == 1) return enqueue(unit);
            return true;
        };
        size_t qh = 0;
        // one sweep finds the unit clauses (and anything unit under the incoming assignment)
        for (size_t c = 0; c < ncl; c++)
            if (!visit(c)) return false;
        for (int i = 0; i < n_extra; i++)
            if (!enqueue(extra[i])) return false;
        while (qh < queue.size()) {
            int32_t p = queue[qh++];
            int32_t f = neg(p);
            for (uint32_t i = ooff[f]; i < ooff[(size_t)f + 1]; i++)
                if (!visit(olist[i])) return false;
        }
        return true;
    }

    // fold fixed assumptions, rebuild the compact device layout
    bool set_fixed(const int32_t* lits, int n) {
        for (int i = 0; i < n; i++)
            if (lits[i] < 0 || (lits[i] >> 1) >= n_vars) {
                error = "fixed assumption out of range";
                return false;
            }
        base_assign = template_assign;
        base_ok = ok;
        if (base_ok) base_ok = bcp_level0(base_assign, lits, n);
        build_layout();
        return true;
    }

    bool base_ok = true;  // ok after fixed assumptions

    void build_layout() {
        live_of_var.assign((size_t)n_vars, -1);
        var_of_live.clear();
        n_base_assigned = 0;
        for (int32_t v = 0; v < n_vars; v++) {
            if (base_assign[v] == LB_UNDEF) {
                live_of_var[v] = (int32_t)var_of_live.size();
                var_of_live.push_back(v);
            } else
                n_base_assigned++;
        }
        n_live = (int32_t)var_of_live.size();
        rec.clear();
        rec_len.clear();
        occ_off.assign((size_t)2 * n_live + 1, 0);
        occ.clear();
        occrec.clear();
        n_rec = n_clauses_live = n_lits_live = 0;
        max_clause_len = 0;
        if (!base_ok) return;
        const size_t ncl = norm_off.size() - 1;
        std::vector<std::pair<uint32_t, uint32_t>> occ_pairs;  // (compact literal, record)
        std::vector<int32_t> cl;
        for (size_t c = 0; c < ncl; c++) {
            cl.clear();
            bool sat = false;
            for (uint32_t i = norm_off[c]; i < norm_off[c + 1]; i++) {
                int32_t l = norm_lits[i];
                uint8_t v = lit_val(base_assign, l);
                if (v == LB_TRUE) {
                    sat = true;
                    break;
                }
                if (v == LB_UNDEF) cl.push_back(2 * live_of_var[l >> 1] + (l & 1));
            }
            if (sat) continue;
            // at a conflict-free fixpoint every remaining clause has >= 2 literals
            uint32_t first_rec = (uint32_t)n_rec;
            size_t pos = 0, len = cl.size();
            max_clause_len = std::max<int32_t>(max_clause_len, (int32_t)len);
            while (true) {
                size_t remaining = len - pos;
                uint16_t slots[REC_SLOTS];
                for (int s = 0; s < REC_SLOTS; s++) slots[s] = (uint16_t)(2 * n_live);
                if (remaining <= REC_SLOTS) {
                    for (size_t s = 0; s < remaining; s++) slots[s] = (uint16_t)(cl[pos + s] ^ 1);
                    pos = len;
                } else {
                    for (int s = 0; s < REC_SLOTS - 1; s++) slots[s] = (uint16_t)(cl[pos + s] ^ 1);
                    slots[REC_SLOTS - 1] = REC_CONT;
                    pos += REC_SLOTS - 1;
                }
                rec.insert(rec.end(), slots, slots + REC_SLOTS);
                rec_len.push_back((uint16_t)(n_rec == (int64_t)first_rec ? len : 0));
                n_rec++;
                if (pos == len) break;
            }
            for (int32_t l : cl) occ_pairs.emplace_back((uint32_t)l, first_rec);
            n_clauses_live++;
            n_lits_live += (int64_t)len;
        }
        for (auto& pr : occ_pairs) occ_off[(size_t)pr.first + 1]++;
        for (size_t i = 0; i < (size_t)2 * n_live; i++) occ_off[i + 1] += occ_off[i];
        occ.resize(occ_pairs.size());
        std::vector<uint32_t> pos(occ_off.begin(), occ_off.end() - 1);
        for (auto& pr : occ_pairs) occ[pos[pr.first]++] = pr.second;
        occrec.resize(occ.size() * REC_SLOTS);
        for (size_t i = 0; i < occ.size(); i++) inline_record(occ[i], &occrec[i * REC_SLOTS]);
    }

    // the 16-byte record a list holds for the clause whose chain starts at rec[r]
    void inline_record(uint32_t r, uint16_t* out) const {
        if (rec_len[r] <= REC_SLOTS) {
            memcpy(out, &rec[(size_t)r * REC_SLOTS], sizeof(uint16_t) * REC_SLOTS);
        } else {
            for (int s = 0; s < REC_SLOTS; s++) out[s] = 0;
            out[0] = (uint16_t)(r & 0xFFFFu);
            out[1] = (uint16_t)(r >> 16);
            out[REC_SLOTS - 1] = REC_INDIRECT;
        }
    }
};

}  // namespace pdsat
