// search.hpp — the decomposition-set search driver re-hosted on batched neighbourhood evaluation.
//
// What the reference's control process does around GetPredict (src_mpi/mpi_predicter.cpp):
//   DeepPredictMain                  :1410-1571  the outer loop, stop conditions (SA temperature)
//   GetDeepPredictTasks              :2697-3056  points of the next neighbourhood (Hamming-1 of the
//                                                current area's centre minus its checked points;
//                                                ts_strategy 4: swap one variable; ts_strategy 2:
//                                                at most TS2_POINTS_COUNT points; FEA / GA points)
//   GetPredict (record part)         :2001-2030  a fully evaluated point that beats the record - or
//                                                is granted by simulated annealing - is the new record
//   NewRecordPoint                   :1687-1808  record bookkeeping, temperature, predict_<n>, graph_file
//   checkSimulatedGranted            :1810-1828  exp(-delta / T) acceptance against a draw in {0.1..0.9}
//   AddNewUncheckedArea              :2405-2510  tabu lists: L2 = areas (centre + which of its Hamming-1
//                                                neighbours are checked), L1 = fully checked centres
//   DeepPredictFindNewUncheckedArea  :1123-1408  where to go next: the record's own area, or - when
//                                                the record was not improved - a random L2 area at
//                                                minimal Hamming distance from the record
//   WritePredictToFile               :2118-2281  the `predict` table
// Here a step evaluates the WHOLE neighbourhood with one batched device call (EvalFn), so the
// per-sample dispatch, stop messages and restart flags of the MPI control loop have no counterpart.
// Random choices use one std::mt19937 (the reference's uint_rand(gen) = raw 32-bit draws of the
// rank-0 generator); points of a step are taken in (size, index) order, where the reference
// shuffles with the unseeded C rand() - a run is reproducible from -seed.
#pragma once
#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <list>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "predict.hpp"

namespace pdsat_host {

static const int MAX_DISTANCE_TO_RECORD = 20;  // mpi_predicter.h:25
static const unsigned TS2_POINTS_COUNT = 100;  // mpi_predicter.h:27
static const int L_POINTS_RADIUS = 1;          // mpi_predicter.h:33

typedef std::vector<uint8_t> Bits;  // one byte per candidate variable (position in full_vars)

inline unsigned bits_count(const Bits& b) {
    unsigned c = 0;
    for (uint8_t x : b) c += x;
    return c;
}
inline unsigned bits_distance(const Bits& a, const Bits& b, int* first_diff = nullptr) {
    unsigned d = 0;
    for (size_t i = 0; i < a.size(); i++)
        if (a[i] != b[i]) {
            if (!d && first_diff) *first_diff = (int)i;
            d++;
        }
    return d;
}
inline std::string bits_string(const Bits& b) {
    std::string s;
    for (uint8_t x : b) s += x ? '1' : '0';
    return s;
}

struct PointEval {
    std::vector<int32_t> set;      // 1-based variables, ascending
    pdsat_cost cost{};
    Estimate est;
    uint64_t n_interrupted = 0;    // handed-off samples the host solver gave up on (set status STOPPED)
    int status = 0;                // set_status_arr: 1 stopped, 2 record, 4 solved
    double s2 = -1.0;              // sample variance of the cost (-1: not available)
    int cur_var_changing = 1;
    bool added_to_L2 = false;
};

struct SearchParams {
    std::vector<int32_t> full_vars;  // candidate variables (full_var_choose_order), ascending
    int deep_predict = 6;            // 5 simulated annealing, >= 6 tabu search, 3/4 plain descent
    int ts_strategy = 0;
    int algorithm_type = 0;          // 0 tabu / SA, 1 FEA, 2 GA
    int beta_value = 2;
    unsigned max_L2_hamming_distance = 2;
    double min_temperature = 20, temperature_multiply_koef = 0.9, start_temperature_koef = 0.1;
    int ga_N = 10, ga_L = 2, ga_H = 4, ga_G = 4;
    uint32_t seed = 0;
    int max_steps = 1000;
    uint64_t cnf_in_set_count = 0;
    int proc_count = 1;
    std::string eval = "prep", schema = "", solver_name = "pdsat-b200 bcp_kernel";
    int verbosity = 0;
    bool write_files = true;
};

class SearchDriver {
   public:
    typedef std::function<bool(std::vector<PointEval>&)> EvalFn;  // fills cost / est / n_interrupted of every point

    SearchParams P;
    EvalFn evaluate;
    // results
    long double best_predict = 0;
    std::vector<int32_t> best_set;
    double best_sum = 0;
    uint64_t best_solved = 0, best_n = 0;
    int record_count = 0, steps_done = 0;
    uint64_t points_evaluated = 0, points_skipped = 0;
    double cur_temperature = 0;
    std::vector<std::vector<int32_t>> centre_history;  // centre of every step's neighbourhood
    std::vector<std::vector<int32_t>> record_history;

    SearchDriver(const SearchParams& p, EvalFn fn) : P(p), evaluate(fn), gen_(p.seed), ga_gen_(p.seed + 7919u) {}

    bool run(const std::vector<int32_t>& init_set) {
        n_ = P.full_vars.size();
        for (size_t i = 0; i < n_; i++) pos_of_var_[P.full_vars[i]] = (int)i;
        for (int32_t v : init_set)
            if (!pos_of_var_.count(v)) {
                error = "initial point holds a variable outside the candidate variables";
                return false;
            }
        centre_set_ = init_set;
        std::sort(centre_set_.begin(), centre_set_.end());
        first_point_ = true;
        std::ostringstream log;
        while (steps_done <= P.max_steps) {
            if (P.deep_predict == 5 && !(cur_temperature == 0 || cur_temperature > P.min_temperature)) break;
            std::vector<PointEval> pts;
            make_tasks(pts, log);
            if (pts.empty() && !first_point_) {
                log << "no unchecked point in this neighbourhood" << std::endl;
                record_updated_ = false;
            } else {
                if (!evaluate(pts)) {
                    error = "evaluation failed";
                    flush_log(log);
                    return false;
                }
                points_evaluated += pts.size();
                centre_history.push_back(area_.centre.empty() ? centre_set_ : to_set(area_.centre));
                record_updated_ = false;
                take_results(pts, log);
                if (P.write_files) write_predict_file("predict", pts);
            }
            if (first_point_) {
                if (L2_.empty()) add_area(to_bits(centre_set_), log);
                area_ = L2_.front();
                log << "First point" << std::endl << "L2.size() " << L2_.size() << std::endl;
                first_point_ = false;
            }
            steps_done++;
            if (!find_new_area(log)) break;
            if (centre_set_.empty()) break;
            flush_log(log);
        }
        log << std::endl << "best_predict_time " << (double)best_predict << " s" << std::endl << "best_var_num " << best_set.size() << std::endl
            << "best_var_choose_order" << std::endl;
        for (int32_t v : best_set) log << v << " ";
        log << std::endl << "points count " << points_evaluated << std::endl << "temperature " << cur_temperature << std::endl;
        flush_log(log);
        return true;
    }

    std::string error;

   private:
    struct Area {
        Bits centre, checked;  // checked[i] = the neighbour that differs in position i was evaluated
        bool partly = false;
    };
    size_t n_ = 0;
    std::map<int32_t, int> pos_of_var_;
    std::list<Area> L2_;
    std::list<Bits> L1_;
    Area area_;
    std::vector<int32_t> centre_set_;  // var_choose_order
    bool first_point_ = true, record_updated_ = false, sa_granted_ = false;
    unsigned unupdated_count_ = 0, stagnation_ = 0, max_stagnation_ = 100;
    double sa_delta_ = 0, sa_exp_ = 0;
    std::mt19937 gen_;     // uint_rand(gen)
    std::mt19937 ga_gen_;  // FEA / GA mutation draws (the reference uses std::random_device there)
    std::vector<std::vector<int32_t>> fea_records_;
    std::vector<std::pair<Bits, double>> ga_records_, ga_population_;
    std::vector<PointEval> last_pts_;

    Bits to_bits(const std::vector<int32_t>& set) const {
        Bits b(n_, 0);
        for (int32_t v : set) b[(size_t)pos_of_var_.at(v)] = 1;
        return b;
    }
    std::vector<int32_t> to_set(const Bits& b) const {
        std::vector<int32_t> s;
        for (size_t i = 0; i < n_; i++)
            if (b[i]) s.push_back(P.full_vars[i]);
        return s;
    }
    void flush_log(std::ostringstream& log) {
        if (P.write_files) {
            std::ofstream f("deep_predict", std::ios::app);
            f << log.str();
        }
        if (P.verbosity > 0) fputs(log.str().c_str(), stdout);
        log.str("");
        log.clear();
    }

    // ---- FEA / GA point generation (mpi_predicter.cpp:2513-2694) ------------------------------
    size_t doerr_alpha(size_t core, int beta) {  // heavy-tailed mutation strength
        if (core < 4) return 1;
        const size_t m = core / 2;
        std::uniform_real_distribution<double> U(0.0, 1.0);
        if (beta == 1) return U(ga_gen_) < 1.0 / (double)core ? m : 1;
        double c = 0;
        for (size_t i = 1; i <= m; i++) c += 1.0 / std::pow((double)i, beta);
        const double p = U(ga_gen_);
        double acc = 0;
        for (size_t i = 1; i <= m; i++) {
            acc += 1.0 / (std::pow((double)i, beta) * c);
            if (p < acc) return i;
        }
        return m;
    }
    Bits mutate(const Bits& c, double p) {
        Bits b = c;
        std::uniform_real_distribution<double> U(0.0, 1.0);
        int changed = 0;
        for (auto& x : b)
            if (U(ga_gen_) < p) {
                x ^= 1;
                changed++;
            }
        if (!changed) b[ga_gen_() % b.size()] ^= 1;
        return b;
    }
    size_t pick_by_fitness(const std::vector<std::pair<Bits, double>>& pop) {
        std::vector<double> w;
        for (auto& q : pop) w.push_back(1.0 / q.second);
        std::discrete_distribution<size_t> d(w.begin(), w.end());
        return d(ga_gen_);
    }
    bool in_L2(const Bits& b) const {
        for (auto& a : L2_)
            if (a.centre == b) return true;
        return false;
    }

    // ---- GetDeepPredictTasks -------------------------------------------------------------------
    void make_tasks(std::vector<PointEval>& pts, std::ostringstream& log) {
        log << std::endl << "*** GetDeepPredictTasks" << std::endl;
        if (P.algorithm_type == 0) log << "ts_strategy " << P.ts_strategy << std::endl;
        std::vector<Bits> cand;
        int var_changing = 1;
        uint64_t skipped = 0;
        if (first_point_) {
            cand.push_back(to_bits(centre_set_));
        } else if (P.algorithm_type == 1 || (P.algorithm_type == 2 && ga_records_.size() < (size_t)P.ga_N)) {
            Bits b;
            int guard = 0;
            do {
                const size_t alpha = doerr_alpha(n_, P.beta_value);
                b = mutate(area_.centre, (double)alpha / (double)n_);
            } while (in_L2(b) && ++guard < 1000);
            if (bits_count(b)) cand.push_back(b);
            log << (P.algorithm_type == 1 ? "using fast evolution algoritm (FEA)" : "creating start population using FEA") << std::endl;
        } else if (P.algorithm_type == 2) {
            if (ga_population_.empty()) {
                for (int i = 0; i < P.ga_N; i++) ga_population_.push_back(ga_records_[ga_records_.size() - 1 - (size_t)i]);
            } else if (last_pts_.size() > 1) {
                std::vector<std::pair<Bits, double>> keep = ga_population_;
                std::sort(keep.begin(), keep.end(), [](const std::pair<Bits, double>& a, const std::pair<Bits, double>& b) { return a.second < b.second; });
                keep.resize(std::min<size_t>(keep.size(), (size_t)P.ga_L));
                ga_population_.clear();
                for (auto& q : last_pts_) ga_population_.push_back({to_bits(q.set), (double)q.est.predict});
                for (auto& q : keep) ga_population_.push_back(q);
            }
            for (int i = 0; i < P.ga_H + P.ga_G; i++) {
                Bits b;
                int guard = 0;
                do {
                    if (i < P.ga_G) {  // crossover of two fitness-chosen parents
                        size_t a = pick_by_fitness(ga_population_), c = pick_by_fitness(ga_population_);
                        for (int g = 0; g < 100 && c == a; g++) c = pick_by_fitness(ga_population_);
                        b = ga_population_[a].first;
                        for (size_t j = 0; j < n_; j++)
                            if (ga_gen_() & 1u) b[j] = ga_population_[c].first[j];
                    } else  // 1+1 random mutation
                        b = mutate(ga_population_[pick_by_fitness(ga_population_)].first, 1.0 / (double)n_);
                } while ((in_L2(b) || !bits_count(b)) && ++guard < 1000);
                cand.push_back(b);
            }
            log << "using genetic algorithm (GA) with parameters (N, L, H, G): " << P.ga_N << ", " << P.ga_L << ", " << P.ga_H << ", " << P.ga_G << std::endl;
        } else if (P.ts_strategy == 4 && !record_updated_) {
            // swap neighbourhood: every centre variable against every other candidate variable
            var_changing = 2;
            for (size_t i = 0; i < n_; i++)
                if (area_.centre[i])
                    for (size_t j = 0; j < n_; j++)
                        if (!area_.centre[j]) {
                            Bits b = area_.centre;
                            b[i] = 0;
                            b[j] = 1;
                            cand.push_back(b);
                        }
        } else {
            for (size_t i = 0; i < n_; i++) {
                if (area_.checked[i]) {
                    skipped++;
                    continue;  // checked already
                }
                Bits b = area_.centre;
                b[i] ^= 1;
                if (bits_count(b)) cand.push_back(b);
            }
        }
        log << "points_to_check " << (first_point_ ? 1 : (P.algorithm_type ? cand.size() : n_)) << std::endl;
        for (auto& b : cand) {
            PointEval pe;
            pe.set = to_set(b);
            pe.cur_var_changing = var_changing;
            pts.push_back(pe);
        }
        if (P.algorithm_type != 2)
            std::stable_sort(pts.begin(), pts.end(), [](const PointEval& a, const PointEval& b) { return a.set.size() < b.set.size(); });
        if (P.ts_strategy == 2 && pts.size() > TS2_POINTS_COUNT) {
            std::vector<PointEval> smaller, larger, out;
            for (auto& q : pts) (q.set.size() < centre_set_.size() ? smaller : larger).push_back(q);
            for (size_t i = 0; i < smaller.size() && i < TS2_POINTS_COUNT / 2; i++) out.push_back(smaller[i]);
            for (size_t i = 0; i < larger.size() && i < TS2_POINTS_COUNT / 2; i++) out.push_back(larger[larger.size() - 1 - i]);
            pts.swap(out);
        }
        points_skipped += skipped;
        log << "must be checked " << pts.size() << std::endl << "current_skipped " << skipped << std::endl << "L1.size() " << L1_.size() << std::endl
            << "L2.size() " << L2_.size() << std::endl << "best_var " << best_set.size() << std::endl << "best_predict_time " << (double)best_predict << std::endl;
    }

    // ---- checkSimulatedGranted -------------------------------------------------------------------
    bool simulated_granted(long double predict) {
        if (P.deep_predict != 5) return false;
        sa_delta_ = (double)(predict - best_predict);
        if (sa_delta_ < 0) return true;
        double r = 0;
        while (r == 0) r = (double)(gen_() % 10) * 0.1;  // a value of 0.1 .. 0.9
        sa_exp_ = std::exp(-sa_delta_ / cur_temperature);
        if (r < sa_exp_) {
            sa_granted_ = true;
            return true;
        }
        return false;
    }

    // ---- GetPredict (record part) + NewRecordPoint + AddNewUncheckedArea ------------------------
    void take_results(std::vector<PointEval>& pts, std::ostringstream& log) {
        for (auto& q : pts) {
            q.status = q.n_interrupted ? 1 : 4;
            const long double pr = q.est.predict;
            if (q.status == 4 && pr > 0 &&
                (best_predict == 0 || pr < best_predict || simulated_granted(pr))) {
                q.status = 2;
                record_updated_ = true;
                best_predict = pr;
                best_sum = q.est.sum;
                best_solved = q.est.solved;
                best_n = q.cost.n;
                new_record(q, log);
            }
        }
        if (P.deep_predict >= 5)
            for (auto& q : pts)
                if (!q.added_to_L2 && q.status > 0) {
                    q.added_to_L2 = true;
                    const Bits b = to_bits(q.set);
                    if (P.algorithm_type >= 1) {
                        Area a;
                        a.centre = b;
                        a.checked.assign(n_, 0);
                        L2_.push_back(a);
                    } else
                        add_area(b, log);
                }
        last_pts_ = pts;
    }

    void new_record(const PointEval& q, std::ostringstream& log) {
        best_set = q.set;
        centre_set_ = q.set;
        record_history.push_back(q.set);
        record_count++;
        log << "NewRecordPoint()" << std::endl << "best_predict_time " << (double)best_predict << " s" << std::endl;
        if (P.algorithm_type == 1) fea_records_.push_back(q.set);
        if (P.algorithm_type == 2) ga_records_.push_back({to_bits(q.set), (double)best_predict});
        if (P.deep_predict == 5) {
            if (sa_granted_) {
                log << "*** Simulated granted" << std::endl << "delta " << sa_delta_ << std::endl << "exp_value " << sa_exp_ << std::endl;
                sa_granted_ = false;
            }
            if (cur_temperature == 0.0 || cur_temperature > (double)best_predict * P.start_temperature_koef)
                cur_temperature = (double)best_predict * P.start_temperature_koef;
            else
                cur_temperature *= P.temperature_multiply_koef;
            log << "cur_temperature " << cur_temperature << std::endl;
        }
        log << "best_var_num " << best_set.size() << std::endl << "best_var_choose_order " << std::endl;
        for (int32_t v : best_set) log << v << " ";
        log << std::endl << std::endl;
        if (!P.write_files) return;
        {
            std::vector<PointEval> one{q};
            one[0].status = 2;
            write_predict_file("predict_" + std::to_string(record_count), one);
        }
        std::ofstream g;
        if (record_count == 1) {
            g.open("graph_file", std::ios::out);
            g << "# best_var_num best_predict_time best_sum_time cnf_in_set_count last_predict_record_time current_predict_time";
            if (P.deep_predict == 5) g << " cur_temperature";
            g << std::endl;
        } else
            g.open("graph_file", std::ios::app);
        g << record_count << " " << best_set.size() << " " << (double)best_predict << " " << best_sum << " " << best_n << " " << 0 << " " << 0 << " " << 0;
        if (P.deep_predict == 5) g << " " << cur_temperature;
        g << " time_limit=" << 0 << std::endl;
    }

    // L2 gets the new centre; every L1 / L2 area at Hamming distance 1 marks the shared neighbour as
    // checked on both sides; an area whose neighbours are all checked moves to L1
    void add_area(const Bits& point, std::ostringstream& log) {
        Area nu;
        nu.centre = point;
        nu.checked.assign(n_, 0);
        const int cnt = (int)bits_count(point);
        for (auto& c : L1_)
            if (std::abs((int)bits_count(c) - cnt) <= L_POINTS_RADIUS) {
                int i = 0;
                if ((int)bits_distance(nu.centre, c, &i) <= L_POINTS_RADIUS && nu.centre != c) nu.checked[(size_t)i] = 1;
            }
        std::vector<std::list<Area>::iterator> touched;
        for (auto it = L2_.begin(); it != L2_.end(); ++it)
            if (std::abs((int)bits_count(it->centre) - cnt) <= L_POINTS_RADIUS) {
                int i = 0;
                const unsigned d = bits_distance(nu.centre, it->centre, &i);
                if (d == 0) return;  // already an area (the reference leaves this unchecked: "too expensive")
                if ((int)d <= L_POINTS_RADIUS) {
                    touched.push_back(it);
                    it->checked[(size_t)i] = 1;
                    nu.checked[(size_t)i] = 1;
                }
            }
        unsigned moved = 0;
        for (auto it : touched)
            if (bits_count(it->checked) == n_) {
                L1_.push_back(it->centre);
                if (area_.centre == it->centre)
                    log << std::endl << "*** Checked area with Hamming distance 1" << std::endl << "best_predict_time " << (double)best_predict << std::endl << std::endl;
                L2_.erase(it);
                moved++;
            }
        if (moved) log << "Moved " << moved << " areas from L2 to L1" << std::endl;
        L2_.push_back(nu);
    }

    // ---- DeepPredictFindNewUncheckedArea ---------------------------------------------------------
    bool find_new_area(std::ostringstream& log) {
        if (record_updated_) {
            log << std::endl << "---Record Updated---" << std::endl;
            const Bits bs = to_bits(centre_set_);
            const unsigned cb = bits_count(bs);
            if (P.algorithm_type >= 1) {
                if (P.beta_value > 1) P.beta_value = 3;
                stagnation_ = 0;
                max_stagnation_ = 100;
                if (!L2_.empty()) area_ = L2_.back();
                else {
                    area_.centre = bs;
                    area_.checked.assign(n_, 0);
                }
            } else {
                bool found = false;
                unsigned e1 = 0, e2 = 0;
                for (auto it = L2_.begin(); it != L2_.end();) {
                    const unsigned c = bits_count(it->centre);
                    if (c > cb && c - cb > (unsigned)MAX_DISTANCE_TO_RECORD) {
                        it = L2_.erase(it);
                        e2++;
                    } else {
                        if (it->centre == bs) {
                            area_ = *it;
                            found = true;
                        }
                        ++it;
                    }
                }
                for (auto it = L1_.begin(); it != L1_.end();) {
                    const unsigned c = bits_count(*it);
                    if (c > cb && c - cb > (unsigned)MAX_DISTANCE_TO_RECORD) {
                        it = L1_.erase(it);
                        e1++;
                    } else
                        ++it;
                }
                if (!found) {  // deep_predict < 5 keeps no lists: the record's own fresh area
                    area_.centre = bs;
                    area_.checked.assign(n_, 0);
                    area_.partly = false;
                    if (P.deep_predict >= 5) log << "***Error. !IsCorrectAddingArea" << std::endl;
                }
                (void)e1;
                (void)e2;
            }
            log << "current_unchecked_area center " << std::endl << bits_string(area_.centre) << std::endl << "current_unchecked_area checked_points " << std::endl
                << bits_string(area_.checked) << std::endl;
            unupdated_count_ = 0;
        } else if (P.algorithm_type >= 1) {
            log << std::endl << "---Record not updated---" << std::endl;
            if (stagnation_ < max_stagnation_) {
                log << "stagnation_count " << ++stagnation_ << std::endl;
            } else if (P.algorithm_type == 1) {
                if (P.beta_value == 1) {
                    if (fea_records_.size() > 1) {
                        centre_set_ = fea_records_[fea_records_.size() - 2];
                        area_.centre = to_bits(centre_set_);
                        area_.checked.assign(n_, 0);
                        max_stagnation_ += 50;
                    }
                } else {
                    P.beta_value = 2;
                    max_stagnation_ += 50;
                }
                log << "stagnation_count " << ++stagnation_ << std::endl;
            }
            return true;
        } else if (P.ts_strategy == 4) {
            log << std::endl << "---Record not updated---" << std::endl << "ts_strategy == 4" << std::endl << "same center, other neighbourhood" << std::endl;
            if (swap_done_) return false;  // the swap neighbourhood did not help either: local minimum
            swap_done_ = true;
            return true;
        } else {
            log << std::endl << "---Record not updated---" << std::endl;
            unupdated_count_++;
            if (P.ts_strategy == 2)
                for (auto& a : L2_)
                    if (a.centre == area_.centre) a.partly = true;
            log << "unupdated_count " << unupdated_count_ << std::endl;
            if (P.deep_predict < 5) return false;  // plain descent: local minimum of the neighbourhood
            const Bits bs = to_bits(best_set);
            log << "finding new unchecked_area" << std::endl << "ts_strategy " << P.ts_strategy << std::endl;
            unsigned min_d = (unsigned)n_;
            for (auto& a : L2_) {
                if (a.partly) continue;
                const unsigned d = bits_distance(a.centre, bs);
                if (d == 0) continue;
                if (d < min_d) min_d = d;
            }
            log << "min hamming distance from L2 " << min_d << std::endl;
            if (min_d > P.max_L2_hamming_distance) {
                log << "min_hamming_distance > max_L2_hamming_distance: " << min_d << " > " << P.max_L2_hamming_distance << std::endl;
                return false;
            }
            std::vector<const Area*> matches;
            for (auto& a : L2_)
                if (!a.partly && bits_distance(a.centre, bs) == min_d) matches.push_back(&a);
            if (matches.empty()) return false;
            std::vector<unsigned> powers;
            for (auto* a : matches) {
                const unsigned c = bits_count(a->centre);
                if (std::find(powers.begin(), powers.end(), c) == powers.end()) powers.push_back(c);
            }
            const unsigned start = gen_() % matches.size();
            const unsigned power = powers[gen_() % powers.size()];
            const Area* pick = nullptr;
            for (size_t i = start; i < matches.size() && !pick; i++)
                if (bits_count(matches[i]->centre) == power) pick = matches[i];
            for (size_t i = 0; i < start && !pick; i++)
                if (bits_count(matches[i]->centre) == power) pick = matches[i];
            area_ = *pick;
            centre_set_ = to_set(area_.centre);
            log << "L2_matches.size() " << matches.size() << std::endl << "rand_power_value " << power << std::endl << "rand_L2_start_search_index " << start << std::endl
                << "current_unchecked_area center " << std::endl << bits_string(area_.centre) << std::endl << "var_choose_order " << std::endl;
            for (int32_t v : centre_set_) log << v << " ";
            log << std::endl;
        }
        swap_done_ = false;
        if (P.write_files) {
            std::map<unsigned, unsigned> by_size;
            for (auto& a : L2_) by_size[bits_count(a.centre)]++;
            std::ofstream f("L2_list");
            f << "L2 point # : set_count : size" << std::endl;
            unsigned k = 1;
            for (auto it = by_size.rbegin(); it != by_size.rend(); ++it) f << k++ << " : " << it->first << " : " << it->second << std::endl;
        }
        return true;
    }
    bool swap_done_ = false;

    // ---- WritePredictToFile: same header, same per-point columns ----------------------------------
    void write_predict_file(const std::string& name, const std::vector<PointEval>& pts) {
        std::ofstream f(name.c_str());
        uint64_t all_tasks = 0;
        for (auto& q : pts) all_tasks += q.cost.n;
        f << "Processor count " << P.proc_count << std::endl << "Count of CNF in set " << P.cnf_in_set_count << std::endl << "Count of CNF in all sets " << all_tasks
          << std::endl << "Count of core-variables " << n_ << std::endl << "Solver name " << P.solver_name << std::endl << "Schema type " << P.schema << std::endl
          << "start activity " << 0 << std::endl << "deep_predict " << P.deep_predict << std::endl << "decomp_set_arr.size() " << pts.size();
        char buf[1024];
        for (auto& q : pts) {
            const uint64_t solved = q.cost.n - q.n_interrupted;
            const uint64_t prepr = q.cost.n_conflict + q.cost.n_full;
            // med / s^2 over the samples' costs: eval=prep charges MIN_SOLVE_TIME (1e-7) per sample
            double med = -1.0, s2 = -1.0, mn = -1.0, mx = -1.0;
            if (q.n_interrupted == 0 && solved > 0) {
                if (P.eval == "prep") {
                    med = mn = mx = 1e-7;
                    s2 = 0.0;
                } else {
                    med = q.est.sum / (double)solved;
                    s2 = q.s2;
                }
            }
            snprintf(buf, sizeof buf, "\n %zu %d sum %-12.4f pred %-10.2e min %-15.10f max %-15.6f med %-15.10f s^2 %-15.10f stopped %d skipped %d solved %" PRIu64
                                      " prepr: %" PRIu64 " (0, 0.001): 0 (0.001, 0.01): 0 (0.01, 0.1): 0 (0.1, 1): 0 (1, 10): 0 (10, 100): 0 (100, 1000): 0 (1000, .): %" PRIu64,
                     q.set.size(), q.status, q.est.sum, (double)q.est.predict, mn, mx, med, s2, (int)q.n_interrupted, 0, solved, prepr,
                     (P.eval == "prep") ? (uint64_t)0 : solved - std::min(solved, prepr));
            f << buf;
        }
        snprintf(buf, sizeof buf, "\n\nAll skipped count %d\nCurrent solved tasks count %" PRIu64 "\nBest var num %zu\nBest predict time %e\nPredict took time %d\n", 0,
                 all_tasks, best_set.size(), (double)best_predict, 0);
        f << buf;
    }
};

}  // namespace pdsat_host
