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
//   4. compaction of the unassigned ("live") variables to 0..n_live-1;
//   5. GATE EXTRACTION: groups of clauses that are exactly the canonical CNF of
//        XOR-m      x_1 ^ ... ^ x_m = p                      (2^(m-1) clauses of m literals)
//        ANDXOR-m   x_1 ^ ... ^ x_m ^ (Lb & Lc) = p          (2 * 2^(m-1) clauses of m+1 literals
//                                                             + 2^(m-1) clauses of m+2 literals)
//      are replaced by one 16-byte gate record each.  The device evaluates a gate in closed
//      form on the 64 samples of a group; the closed forms reach the same fixpoint and the same
//      conflicts as unit propagation over the gate's clauses (exhaustively checked over all
//      3^n partial assignments in tests/test_gates.py), so every BCP result is unchanged.
//      Bivium: 12 800 clauses = 600 gates, no plain clause left.  PDSAT_NO_GATES=1 disables it;
//   6. emission of 16-byte clause records + per-literal occurrence lists for the remaining
//      clauses, per-variable gate lists, and the "index blob" (occurrence offsets, gate lists,
//      gate records) that the kernel stages into shared memory with one bulk copy.
//
// Pure C++ (no CUDA): also used by the CPU-only tests through a host-only handle.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <string>
#include <vector>

namespace pdsat {

enum : uint8_t { LB_TRUE = 0, LB_FALSE = 1, LB_UNDEF = 2 };

// unused literal slots hold the code of the pad plane (= 2*n_live, "always false")
static constexpr uint16_t REC_CONT = 0xFFFE;      // slot 7: clause continues in the next record
static constexpr uint16_t REC_INDIRECT = 0xFFFD;  // slot 7 of an inline record: slots 0-1 = index of the chain in rec
static constexpr int REC_SLOTS = 8;           // uint16 slots per 16-byte record
// gate records: slot 7 = tag | parity; slots 0.. = TRUE-plane codes (2*live) of the xor
// variables (unused slots = code of the pad variable: known, value 0); ANDXOR keeps the codes of
// the literals Lb, Lc in slots 5, 6
static constexpr uint16_t REC_GATE_XOR = 0xFFF0;     // | parity
static constexpr uint16_t REC_GATE_ANDXOR = 0xFFF2;  // | parity
static constexpr uint16_t REC_TAG_MIN = 0xFFF0;      // plane codes stay below this
static constexpr int GATE_XOR_MAX = 7;               // xor variables of an XOR gate
static constexpr int GATE_ANDXOR_MAX = 5;            // xor variables of an ANDXOR gate

struct HostGate {
    uint8_t type = 0;            // 0 XOR, 1 ANDXOR
    uint8_t parity = 0;          // p
    uint8_t m = 0;               // xor variables
    int32_t x[GATE_XOR_MAX];     // live indices of the xor variables
    int32_t lb = -1, lc = -1;    // compact literal codes (ANDXOR)
    uint32_t n_clauses = 0;      // clauses it replaces
};

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
    std::vector<uint16_t> occrec;         // per literal (docc_off): the records of its SHORT clauses, inline
    std::vector<uint32_t> docc_off;       // 2*n_live + 1: device occurrence offsets (short clauses only)
    // long clauses (> 8 literals) are visited lazily: per literal the ids of the long clauses that
    // contain it; the device marks them when the literal is falsified and evaluates every marked
    // clause ONCE when the frontier is empty (instead of at every falsification)
    std::vector<uint32_t> long_rec;       // id -> first record of the chain in rec
    std::vector<uint32_t> locc_off;       // 2*n_live + 1
    std::vector<uint16_t> locc;           // long clause ids
    bool lazy_long = false;
    // gates
    std::vector<HostGate> gates;
    std::vector<uint16_t> gate_rec;       // n_gates * 8 slots
    std::vector<uint32_t> gocc_off;       // n_live + 1: per VARIABLE, the gates that contain it
    std::vector<uint16_t> gocc;           // gate ids
    int64_t n_gate_clauses = 0;           // clauses replaced by gates
    // index blob staged into shared memory by the kernel (16-byte aligned sections)
    std::vector<uint8_t> blob;
    uint32_t blob_occ_off = 0, blob_gocc_off = 0, blob_gocc = 0, blob_grec = 0, blob_locc_off = 0, blob_locc = 0, blob_lrec = 0;
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

    // ---- gate extraction over the live clauses (compact literals, sorted) ---------------
    struct ClauseKey {
        uint64_t a = 0, b = 0;  // up to 8 literals of 16 bits
        bool operator==(const ClauseKey& o) const { return a == o.a && b == o.b; }
    };
    struct ClauseKeyHash {
        size_t operator()(const ClauseKey& k) const {
            uint64_t h = k.a * 0x9E3779B97F4A7C15ull ^ (k.b + 0x7F4A7C159E3779B9ull) * 0xBF58476D1CE4E5B9ull;
            return (size_t)(h ^ (h >> 29));
        }
    };
    static ClauseKey key_of(const int32_t* l, size_t n) {
        ClauseKey k;
        for (size_t i = 0; i < n; i++) {
            const uint64_t v = (uint64_t)(uint16_t)(l[i] + 1);  // +1: an absent slot is 0
            if (i < 4) k.a |= v << (16 * i);
            else k.b |= v << (16 * (i - 4));
        }
        return k;
    }

    void extract_gates(const std::vector<std::vector<int32_t>>& cls, std::vector<char>& covered) {
        gates.clear();
        covered.assign(cls.size(), 0);
        if (getenv("PDSAT_NO_GATES")) return;
        std::unordered_map<ClauseKey, uint32_t, ClauseKeyHash> index;  // clause -> id (first of duplicates)
        index.reserve(cls.size() * 2);
        for (size_t c = 0; c < cls.size(); c++)
            if (cls[c].size() >= 2 && cls[c].size() <= 8) index.emplace(key_of(cls[c].data(), cls[c].size()), (uint32_t)c);
        // the canonical clause over variables vs[0..n) (ascending) with sign bits `signs`
        int32_t tmp[8];
        auto find = [&](const int32_t* vs, int n, uint32_t signs) -> int64_t {
            for (int i = 0; i < n; i++) tmp[i] = 2 * vs[i] + (int32_t)((signs >> i) & 1u);
            auto it = index.find(key_of(tmp, (size_t)n));
            if (it == index.end() || covered[it->second]) return -1;
            return (int64_t)it->second;
        };
        std::vector<uint32_t> found;
        // all 2^(n-1) clauses over vs whose sign parity is `par`, with the sign of the positions in
        // `fixed_mask` pinned to `fixed_signs`; appends their ids to `found`, false if one is missing
        auto collect = [&](const int32_t* vs, int n, uint32_t fixed_mask, uint32_t fixed_signs, int par_free) -> bool {
            const uint32_t all = 1u << n;
            for (uint32_t sg = 0; sg < all; sg++) {
                if ((sg & fixed_mask) != fixed_signs) continue;
                if ((__builtin_popcount(sg & ~fixed_mask) & 1) != par_free) continue;
                const int64_t id = find(vs, n, sg);
                if (id < 0) return false;
                found.push_back((uint32_t)id);
            }
            return true;
        };
        auto mark = [&]() {
            for (uint32_t id : found) covered[id] = 1;
        };
        // pass 1: ANDXOR (its widest clauses identify it); pass 2: XOR over what is left.
        // ANDXOR first, because the b- and c-halves of a gate with one xor variable fixed could
        // otherwise be claimed by nothing, and an XOR never shares clauses with an ANDXOR.
        for (size_t c = 0; c < cls.size() && gates.size() < 0xFFF0; c++) {
            const std::vector<int32_t>& cl = cls[c];
            const int n = (int)cl.size();
            if (covered[c] || n < 3 || n > GATE_ANDXOR_MAX + 2) continue;
            int32_t vs[8];
            uint32_t sg0 = 0;
            for (int i = 0; i < n; i++) {
                vs[i] = cl[i] >> 1;
                sg0 |= (uint32_t)(cl[i] & 1) << i;
            }
            bool done = false;
            // this clause as a member of G_bc: positions pb < pc hold ~Lb, ~Lc
            for (int pb = 0; pb < n && !done; pb++)
                for (int pc = pb + 1; pc < n && !done; pc++) {
                    const uint32_t fm = (1u << pb) | (1u << pc);
                    const uint32_t fs = sg0 & fm;
                    const int par1 = __builtin_popcount(sg0 & ~fm) & 1;  // forbidden-point parity of G_bc
                    found.clear();
                    if (!collect(vs, n, fm, fs, par1)) continue;
                    // G_b over X + {b}: literal Lb (the opposite sign), xor parity par1 ^ 1
                    int32_t vb[8], vc[8];
                    int nb = 0, nc = 0, posb = 0, posc = 0;
                    for (int i = 0; i < n; i++) {
                        if (i != pc) {
                            if (i == pb) posb = nb;
                            vb[nb++] = vs[i];
                        }
                        if (i != pb) {
                            if (i == pc) posc = nc;
                            vc[nc++] = vs[i];
                        }
                    }
                    const uint32_t sb = ((sg0 >> pb) & 1u) ^ 1u, sc = ((sg0 >> pc) & 1u) ^ 1u;
                    if (!collect(vb, nb, 1u << posb, sb << posb, par1 ^ 1)) continue;
                    if (!collect(vc, nc, 1u << posc, sc << posc, par1 ^ 1)) continue;
                    HostGate g;
                    g.type = 1;
                    g.m = (uint8_t)(n - 2);
                    int xi = 0;
                    for (int i = 0; i < n; i++)
                        if (i != pb && i != pc) g.x[xi++] = vs[i];
                    g.lb = 2 * vs[pb] + (int32_t)sb;  // Lb: the literal whose falsity forces the xor
                    g.lc = 2 * vs[pc] + (int32_t)sc;
                    // Lb false: xor(X) = par0 ^ 1 with par0 = par1 ^ 1  =>  xor(X) ^ (Lb & Lc) = par1
                    g.parity = (uint8_t)par1;
                    g.n_clauses = (uint32_t)found.size();
                    mark();
                    gates.push_back(g);
                    done = true;
                }
        }
        for (size_t c = 0; c < cls.size() && gates.size() < 0xFFF0; c++) {
            const std::vector<int32_t>& cl = cls[c];
            const int n = (int)cl.size();
            if (covered[c] || n < 2 || n > GATE_XOR_MAX) continue;
            int32_t vs[8];
            uint32_t sg0 = 0;
            for (int i = 0; i < n; i++) {
                vs[i] = cl[i] >> 1;
                sg0 |= (uint32_t)(cl[i] & 1) << i;
            }
            const int par = __builtin_popcount(sg0) & 1;  // parity of the forbidden points
            found.clear();
            if (!collect(vs, n, 0, 0, par)) continue;
            HostGate g;
            g.type = 0;
            g.m = (uint8_t)n;
            for (int i = 0; i < n; i++) g.x[i] = vs[i];
            g.parity = (uint8_t)(par ^ 1);  // xor(x) != par
            g.n_clauses = (uint32_t)found.size();
            mark();
            gates.push_back(g);
        }
    }

    void gate_record(const HostGate& g, uint16_t* out) const {
        const uint16_t pad_var = (uint16_t)(2 * n_live + 2);  // TRUE plane 0, FALSE plane all ones
        for (int s = 0; s < REC_SLOTS - 1; s++) out[s] = pad_var;
        for (int i = 0; i < g.m; i++) out[i] = (uint16_t)(2 * g.x[i]);
        if (g.type == 1) {
            out[5] = (uint16_t)g.lb;
            out[6] = (uint16_t)g.lc;
            out[7] = (uint16_t)(REC_GATE_ANDXOR | g.parity);
        } else
            out[7] = (uint16_t)(REC_GATE_XOR | g.parity);
    }

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
        docc_off.assign((size_t)2 * n_live + 1, 0);
        locc_off.assign((size_t)2 * n_live + 1, 0);
        locc.clear();
        long_rec.clear();
        lazy_long = false;
        gates.clear();
        gate_rec.clear();
        gocc_off.assign((size_t)n_live + 1, 0);
        gocc.clear();
        n_rec = n_clauses_live = n_lits_live = n_gate_clauses = 0;
        max_clause_len = 0;
        if (base_ok) {
            const size_t ncl = norm_off.size() - 1;
            // live clauses in compact literals (sorted: the compaction is monotone)
            std::vector<std::vector<int32_t>> cls;
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
                cls.push_back(cl);
                n_clauses_live++;
                n_lits_live += (int64_t)cl.size();
                max_clause_len = std::max<int32_t>(max_clause_len, (int32_t)cl.size());
            }
            std::vector<char> covered;
            extract_gates(cls, covered);
            std::vector<std::pair<uint32_t, uint32_t>> occ_pairs;  // (compact literal, record)
            for (size_t c = 0; c < cls.size(); c++) {
                if (covered[c]) {
                    n_gate_clauses++;
                    continue;
                }
                const std::vector<int32_t>& k = cls[c];
                uint32_t first_rec = (uint32_t)n_rec;
                size_t pos = 0, len = k.size();
                while (true) {
                    size_t remaining = len - pos;
                    uint16_t slots[REC_SLOTS];
                    for (int s = 0; s < REC_SLOTS; s++) slots[s] = (uint16_t)(2 * n_live);
                    if (remaining <= REC_SLOTS) {
                        for (size_t s = 0; s < remaining; s++) slots[s] = (uint16_t)(k[pos + s] ^ 1);
                        pos = len;
                    } else {
                        for (int s = 0; s < REC_SLOTS - 1; s++) slots[s] = (uint16_t)(k[pos + s] ^ 1);
                        slots[REC_SLOTS - 1] = REC_CONT;
                        pos += REC_SLOTS - 1;
                    }
                    rec.insert(rec.end(), slots, slots + REC_SLOTS);
                    rec_len.push_back((uint16_t)(n_rec == (int64_t)first_rec ? len : 0));
                    n_rec++;
                    if (pos == len) break;
                }
                for (int32_t l : k) occ_pairs.emplace_back((uint32_t)l, first_rec);
            }
            for (auto& pr : occ_pairs) occ_off[(size_t)pr.first + 1]++;
            for (size_t i = 0; i < (size_t)2 * n_live; i++) occ_off[i + 1] += occ_off[i];
            occ.resize(occ_pairs.size());
            std::vector<uint32_t> pos(occ_off.begin(), occ_off.end() - 1);
            for (auto& pr : occ_pairs) occ[pos[pr.first]++] = pr.second;
            // every list: short clauses first, so that the lanes of a trip agree on how many of the
            // 8 record slots are in use (the device skips pad slot pairs)
            for (size_t l = 0; l < (size_t)2 * n_live; l++)
                std::stable_sort(occ.begin() + occ_off[l], occ.begin() + occ_off[l + 1],
                                 [&](uint32_t x, uint32_t y) { return rec_len[x] < rec_len[y]; });
            // device lists: short clauses inline; long clauses by id (lazy) when their ids fit 16 bits
            size_t n_long = 0;
            for (size_t r = 0; r < rec_len.size(); r++)
                if (rec_len[r] > REC_SLOTS) n_long++;
            lazy_long = n_long > 0 && n_long < 0xFFFF && !getenv("PDSAT_NO_LAZY_LONG");
            std::vector<uint32_t> long_id(rec_len.size(), 0);
            if (lazy_long)
                for (size_t r = 0; r < rec_len.size(); r++)
                    if (rec_len[r] > REC_SLOTS) {
                        long_id[r] = (uint32_t)long_rec.size();
                        long_rec.push_back((uint32_t)r);
                    }
            docc_off.assign((size_t)2 * n_live + 1, 0);
            locc_off.assign((size_t)2 * n_live + 1, 0);
            std::vector<uint32_t> docc;
            for (size_t l = 0; l < (size_t)2 * n_live; l++) {
                for (uint32_t o = occ_off[l]; o < occ_off[l + 1]; o++) {
                    const uint32_t r = occ[o];
                    if (lazy_long && rec_len[r] > REC_SLOTS) locc.push_back((uint16_t)long_id[r]);
                    else docc.push_back(r);
                }
                docc_off[l + 1] = (uint32_t)docc.size();
                locc_off[l + 1] = (uint32_t)locc.size();
            }
            occrec.resize(docc.size() * REC_SLOTS);
            for (size_t i = 0; i < docc.size(); i++) inline_record(docc[i], &occrec[i * REC_SLOTS]);
            // gates: records + per-variable lists
            gate_rec.resize(gates.size() * REC_SLOTS);
            std::vector<std::pair<uint32_t, uint32_t>> gp;  // (live var, gate)
            for (size_t g = 0; g < gates.size(); g++) {
                gate_record(gates[g], &gate_rec[g * REC_SLOTS]);
                for (int i = 0; i < gates[g].m; i++) gp.emplace_back((uint32_t)gates[g].x[i], (uint32_t)g);
                if (gates[g].type == 1) {
                    gp.emplace_back((uint32_t)(gates[g].lb >> 1), (uint32_t)g);
                    gp.emplace_back((uint32_t)(gates[g].lc >> 1), (uint32_t)g);
                }
            }
            for (auto& pr : gp) gocc_off[(size_t)pr.first + 1]++;
            for (size_t i = 0; i < (size_t)n_live; i++) gocc_off[i + 1] += gocc_off[i];
            gocc.resize(gp.size());
            std::vector<uint32_t> gpos(gocc_off.begin(), gocc_off.end() - 1);
            for (auto& pr : gp) gocc[gpos[pr.first]++] = (uint16_t)pr.second;
        }
        // index blob: [occ offsets (short clauses)][gocc_off][gocc][gate records][locc_off][locc]
        // [long_rec], sections 16-byte aligned; the long-clause sections only when there are any
        auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        blob_occ_off = 0;
        blob_gocc_off = (uint32_t)align16(docc_off.size() * 4);
        blob_gocc = (uint32_t)align16(blob_gocc_off + gocc_off.size() * 4);
        blob_grec = (uint32_t)align16(blob_gocc + gocc.size() * 2);
        blob_locc_off = (uint32_t)align16(blob_grec + gate_rec.size() * 2);
        blob_locc = (uint32_t)align16(blob_locc_off + (lazy_long ? locc_off.size() * 4 : 0));
        blob_lrec = (uint32_t)align16(blob_locc + locc.size() * 2);
        blob.assign(align16(blob_lrec + long_rec.size() * 4), 0);
        memcpy(blob.data() + blob_occ_off, docc_off.data(), docc_off.size() * 4);
        memcpy(blob.data() + blob_gocc_off, gocc_off.data(), gocc_off.size() * 4);
        if (!gocc.empty()) memcpy(blob.data() + blob_gocc, gocc.data(), gocc.size() * 2);
        if (!gate_rec.empty()) memcpy(blob.data() + blob_grec, gate_rec.data(), gate_rec.size() * 2);
        if (lazy_long) {
            memcpy(blob.data() + blob_locc_off, locc_off.data(), locc_off.size() * 4);
            memcpy(blob.data() + blob_locc, locc.data(), locc.size() * 2);
            memcpy(blob.data() + blob_lrec, long_rec.data(), long_rec.size() * 4);
        }
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
