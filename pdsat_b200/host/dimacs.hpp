// dimacs.hpp — host parser feeding pdsat_upload_cnf.
// Mirrors minisat22_wrapper::parse_DIMACS_to_problem (src_common/minisat22_wrapper.cpp:43-69)
// for the clause stream and MPI_Base::ReadIntCNF's comment protocol
// (src_mpi/mpi_base.cpp:546-599): `c input variables N`, `c output variables N`,
// `c known_bits N`, `c var_set v1 v2 ...`.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pdsat_host {

struct Cnf {
    int32_t n_vars = 0;
    std::vector<uint32_t> offsets{0};
    std::vector<int32_t> lits;  // MiniSat encoding 2*var+sign (SolverTypes.h:46-62)
    int32_t input_variables = 0;   // core_len
    int32_t output_variables = 0;  // output_len
    int32_t known_bits = 0;
    std::vector<int32_t> var_set;  // 1-based
    int64_t n_clauses() const { return (int64_t)offsets.size() - 1; }
};

inline bool read_dimacs(const std::string& path, Cnf& cnf, std::string& err) {
    std::ifstream in(path);
    if (!in.is_open()) {
        err = "cannot open " + path;
        return false;
    }
    std::string line;
    int32_t max_var = 0, header_vars = 0;
    while (std::getline(in, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos) continue;
        if (line[p] == 'c') {
            std::istringstream ss(line.substr(p + 1));
            std::string w1, w2;
            ss >> w1;
            if (w1 == "input" && (ss >> w2) && w2 == "variables") ss >> cnf.input_variables;
            else if (w1 == "output" && (ss >> w2) && w2 == "variables") ss >> cnf.output_variables;
            else if (w1 == "known_bits") ss >> cnf.known_bits;
            else if (w1 == "var_set") {
                int v;
                while (ss >> v) cnf.var_set.push_back(v);
            }
            continue;
        }
        if (line[p] == 'p') {
            std::istringstream ss(line.substr(p));
            std::string a, b;
            long nc = 0;
            ss >> a >> b >> header_vars >> nc;
            continue;
        }
        std::istringstream ss(line.substr(p));
        long v;
        while (ss >> v) {
            if (v == 0) {
                cnf.offsets.push_back((uint32_t)cnf.lits.size());
            } else {
                int32_t var = (int32_t)std::labs(v) - 1;
                if (var + 1 > max_var) max_var = var + 1;
                cnf.lits.push_back(2 * var + (v < 0 ? 1 : 0));
            }
        }
    }
    if (cnf.offsets.back() != cnf.lits.size()) cnf.offsets.push_back((uint32_t)cnf.lits.size());
    cnf.n_vars = max_var > header_vars ? max_var : header_vars;
    return true;
}

}  // namespace pdsat_host
