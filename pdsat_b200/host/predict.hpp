// predict.hpp — host half of the predict path: sample counts, schemas, neighbourhoods and the
// estimator arithmetic, with the reference's types (long double for med / predict,
// src_mpi/mpi_predicter.h:237-239).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <set>
#include <string>
#include <vector>

#include "../../include/pdsat_cuda.h"

namespace pdsat_host {

static const unsigned MAX_STEP_CNF_IN_SET = 30;         // mpi_predicter.h:17
static const double MIN_PERCENT_SOLVED_IN_TIME = 0.5;   // mpi_predicter.h:29
static const long double HUGE_DOUBLE = 1e+308L;         // mpi_base.h:43

// MPI_Predicter::AllocatePredictArrays (mpi_predicter.cpp:2346-2367)
inline uint64_t samples_for_set(size_t k, uint64_t cnf_in_set_count) {
    if (k > MAX_STEP_CNF_IN_SET) return cnf_in_set_count;
    uint64_t all = 1ull << k;
    return cnf_in_set_count < all ? cnf_in_set_count : all;
}

// MPI_Base::MakeVarChoose (mpi_base.cpp:275-347); result sorted ascending like the reference
inline bool schema_vars(const std::string& schema, unsigned count, std::vector<int32_t>& out) {
    out.clear();
    if (schema.empty() || schema == "0" || schema == "bivium_Beginning1")
        for (unsigned i = 0; i < count; i++) out.push_back(i + 1);
    else if (schema == "bivium_Beginning2")
        for (unsigned i = 0; i < count; i++) out.push_back(i + 84 + 1);
    else if (schema == "bivium_Beginning_halved") {
        for (unsigned i = 0; i < count / 2; i++) out.push_back(i + 1);
        for (unsigned i = 0; i < count / 2; i++) out.push_back(i + 84 + 1);
    } else if (schema == "bivium_Ending1")
        for (unsigned i = 0; i < count; i++) out.push_back(84 - i);
    else if (schema == "bivium_Ending2")
        for (unsigned i = 0; i < count; i++) out.push_back(177 - i);
    else if (schema == "bivium_Ending_halved") {
        for (unsigned i = 0; i < count / 2; i++) out.push_back(84 - i);
        for (unsigned i = 0; i < count / 2; i++) out.push_back(177 - i);
    } else if (schema == "a5") {
        for (unsigned i = 0; i < 9; i++) out.push_back(i + 1);
        for (unsigned i = 0; i < 11; i++) out.push_back(i + 19 + 1);
        for (unsigned i = 0; i < 11; i++) out.push_back(i + 41 + 1);
    } else
        return false;
    std::sort(out.begin(), out.end());
    return true;
}

// GetDeepPredictTasks default strategy (mpi_predicter.cpp:2926-2935): flip each core variable
inline std::vector<std::vector<int32_t>> hamming1(const std::vector<int32_t>& centre, int core_len) {
    std::set<int32_t> c(centre.begin(), centre.end());
    std::vector<std::vector<int32_t>> out;
    for (int v = 1; v <= core_len; v++) {
        std::set<int32_t> s = c;
        if (s.count(v)) s.erase(v);
        else s.insert(v);
        if (!s.empty()) out.emplace_back(s.begin(), s.end());
    }
    return out;
}

struct Estimate {
    double sum = 0;            // sum_time_arr
    long double med = 0;       // med_time_arr
    long double predict = 0;   // predict_time_arr
    uint64_t solved = 0;       // solved on preprocessing (conflict or model)
};

// GetPredict (mpi_predicter.cpp:1976-1996) on the device's integer accumulators:
//   "prep"        getCurPredictTime (:1846-1901) with solved = samples BCP decides (conflict or
//                 model) - exactly the reference's prep estimate;
//   "bcp_trail"   mean BCP trail of the conflict-free samples (the size of the propagated
//                 assignment; conflict samples count 0) x 2^k / proc - a BCP-only cost, NOT the
//                 reference's "propagation" (which counts the propagations of the full CDCL run);
//   "propagation" is computed by the caller from per-sample costs after the CDCL hand-off
//                 (pd_sat_main.cpp Evaluator); this function then only supplies the counts.
inline Estimate estimate(const pdsat_cost& c, size_t k, const std::string& eval, int proc_count) {
    Estimate e;
    e.solved = c.n_conflict + c.n_full;
    if (eval == "prep") {
        e.sum = 1e-7 * (double)c.n;  // MIN_SOLVE_TIME per sample (mpi_base.h:35)
        double percent = (double)e.solved * 100 / (double)c.n;
        if (percent <= MIN_PERCENT_SOLVED_IN_TIME) {
            e.predict = HUGE_DOUBLE;
            return e;
        }
        long double prob = (double)e.solved / (double)c.n;
        e.predict = pow(2.0, (double)k) * 3.0 / prob;
        return e;
    }
    e.sum = (double)c.sum_trail;  // exact: an integer below 2^53
    e.med = e.sum / (double)c.n;
    e.predict = e.med / (double)proc_count;
    e.predict *= pow(2.0, (double)k);
    return e;
}

// sample variance of the conflict-free samples' cost (WritePredictToFile,
// mpi_predicter.cpp:2219-2224) from the integer sums; numerator exact in 128-bit integers
inline double sample_variance(const pdsat_cost& c) {
    const uint64_t n = c.n - c.n_conflict;
    if (n < 2) return -1.0;
    const unsigned __int128 num = (unsigned __int128)n * c.sumsq_trail - (unsigned __int128)c.sum_trail * c.sum_trail;
    return (double)num / ((double)n * (double)(n - 1));
}

}  // namespace pdsat_host
