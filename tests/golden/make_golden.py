"""Generate tests/golden/*.npz from the UNMODIFIED reference MiniSat (oracle/_ref).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py
The vectors pin, per sample, what MPI_Predicter::solverProgramCalling
(src_mpi/mpi_predicter.cpp:691-773, evaluation_type "prep") observes on the reference
solver: conflict, full assignment, process_sat_count, isSolvedOnPreprocessing, trail size,
Solver::propagations, Solver::watch_scans and the level-0 assignment.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from pdsat_b200 import fixtures as fx  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bivium_cases():
    """(name, set_vars, with_keystream, n_samples, seed)"""
    return [
        ("ending2_k30", [177 - i for i in range(30)], True, 256, 1),
        ("beginning1_k30", list(range(1, 31)), True, 256, 1),
        ("ending2_k85", [177 - i for i in range(85)], True, 128, 1),
        ("ending2_k86", [177 - i for i in range(86)], True, 256, 1),
        ("ending2_k90", [177 - i for i in range(90)], True, 128, 1),
        ("ending2_k100", [177 - i for i in range(100)], True, 128, 1),
        ("ending2_k120", [177 - i for i in range(120)], True, 128, 2),
        ("ending2_k177", [177 - i for i in range(177)], True, 64, 3),
        ("nostream_k177", list(range(1, 178)), False, 96, 1),
        ("nostream_k150", list(range(1, 151)), False, 96, 5),
    ]


def keystream_lits(secret_seed: int = 0):
    """Secret state = first 177 bool_rand bits of mt19937(secret_seed); 200 keystream bits
    as unit assumptions on vars 578..777."""
    secret = orc.bool_rand(secret_seed, 0, 177)
    ks = fx.BiviumModel(secret).stream(200)
    return secret, [2 * (577 + i) + (0 if ks[i] else 1) for i in range(200)]


def main():
    assert orc.have_ref(), "oracle/_ref/libpdsat_ref.so missing: run `make -C oracle ref`"
    nv, cl, _ = fx.bivium_template()
    offs, lits = fx.to_csr(cl)
    R = orc.Oracle(nv, offs, lits, "ref")
    secret, stream = keystream_lits(0)
    out = {"secret": np.asarray(secret, dtype=np.uint8), "stream_lits": np.asarray(stream, dtype=np.int32)}
    for name, sv, with_ks, n, seed in bivium_cases():
        A = orc.sample_assumptions(seed, sv, 0, n)
        rows = np.zeros((n, 7), dtype=np.int64)
        asg = np.zeros((n, nv), dtype=np.uint8)
        for s in range(n):
            a = list(A[s]) + (stream if with_ks else [])
            r, assigns = R.predict_sample(a, 0)
            rows[s] = (r.conflict, r.full, r.sat, r.prep, r.trail_size, r.propagations, r.watch_scans)
            asg[s] = assigns
        out[name + "_set"] = np.asarray(sv, dtype=np.int32)
        out[name + "_seed"] = np.asarray([seed, int(with_ks)], dtype=np.int64)
        out[name + "_rows"] = rows
        out[name + "_assign"] = np.packbits(asg == 0, axis=1)     # true plane
        out[name + "_assign_f"] = np.packbits(asg == 1, axis=1)   # false plane
        print(name, "conflicts", int(rows[:, 0].sum()), "full", int(rows[:, 1].sum()),
              "mean trail %.1f" % rows[:, 4].mean())
    # te > 0 mode (src_mpi/mpi_predicter.cpp:3188-3235): per-sample KNOWN state + its own
    # keystream; assumptions = set-variable values of that state, then all 200 stream bits.
    # The reference seeds the state generator from time(NULL) (src_mpi/mpi_base.cpp:946); a
    # fixed seed (7) is the declared deviation.
    n_te = 128
    states = orc.bool_rand(7, 0, 177 * n_te).reshape(n_te, 177)
    streams = np.asarray([fx.BiviumModel(states[s]).stream(200) for s in range(n_te)], dtype=np.uint8)
    out["te_states"] = np.packbits(states, axis=1)
    out["te_streams"] = np.packbits(streams, axis=1)
    for k in (80, 86, 88, 100):
        sv = [177 - i for i in range(k)]
        rows = np.zeros((n_te, 7), dtype=np.int64)
        asg = np.zeros((n_te, nv), dtype=np.uint8)
        for s in range(n_te):
            a = [2 * (v - 1) + (0 if states[s][v - 1] else 1) for v in sv] + \
                [2 * (577 + i) + (0 if streams[s][i] else 1) for i in range(200)]
            r, assigns = R.predict_sample(a, 0)
            rows[s] = (r.conflict, r.full, r.sat, r.prep, r.trail_size, r.propagations, r.watch_scans)
            asg[s] = assigns
        name = "te_k%d" % k
        out[name + "_set"] = np.asarray(sv, dtype=np.int32)
        out[name + "_rows"] = rows
        out[name + "_assign"] = np.packbits(asg == 0, axis=1)
        out[name + "_assign_f"] = np.packbits(asg == 1, axis=1)
        print(name, "conflicts", int(rows[:, 0].sum()), "full", int(rows[:, 1].sum()),
              "mean trail %.1f" % rows[:, 4].mean())
    np.savez_compressed(os.path.join(HERE, "bivium_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "bivium_golden.npz"))


if __name__ == "__main__":
    main()


def odls_sample():
    """Data fixture: the first 4 ODLS pairs of the reference's ODLS_10_pairs.txt."""
    pairs = fx.parse_odls_pairs(open("/root/reference/src_common/ODLS_10_pairs.txt").read())
    with open(os.path.join(HERE, "odls_pairs_sample.txt"), "w") as f:
        f.write("# first 4 pairs of orthogonal diagonal Latin squares of order 10 from the reference's\n"
                "# src_common/ODLS_10_pairs.txt (data fixture; written by tests/golden/make_golden.py)\n")
        for a, b in pairs[:4]:
            for ra, rb in zip(a, b):
                f.write("".join(map(str, ra)) + " " + "".join(map(str, rb)) + "\n")
            f.write("\n")


def latin_conseq_rows():
    """Data fixture: the 11 assumption rows (26 positive literals each: first row of square A and
    8 cells of its second and third rows) of the reference's pdsat_conseq sample input
    VS_latin_conseq/in:2-12 - the problems latin_square::SolveOneProblem
    (src_common/latin_squares.cpp:133-233) is fed."""
    rows = []
    for line in open("/root/reference/VS_latin_conseq/in"):
        t = line.split()
        if len(t) > 2 and t[0] == "c" and t[1].isdigit():
            rows.append([int(x) for x in t[1:]])
    assert len(rows) == 11 and all(len(r) == 26 for r in rows)
    with open(os.path.join(HERE, "latin_conseq_rows.txt"), "w") as f:
        f.write("# assumption rows of the reference's VS_latin_conseq/in:2-12 (positive literals, 1-based\n"
                "# variables p*1000 + i*100 + j*10 + z + 1); written by tests/golden/make_golden.py\n")
        for r in rows:
            f.write(" ".join(map(str, r)) + "\n")


def rc2_golden():
    """Solver::gen_valid_assumptions_rc2 (Solver.cc:1293-1435) run by the reference itself on
    the weakened-Bivium cubes of SURVEY.md 8c-2b: template + keystream + the first N state bits
    as units, enumerate the last k state variables (d_set ascending, d_set[0] most significant,
    start = all negative)."""
    nv, cl, _ = fx.bivium_template()
    offs, lits = fx.to_csr(cl)
    secret, stream = keystream_lits(0)
    out = {}
    for n_fixed, k in ((161, 16), (100, 14), (60, 12), (150, 10), (85, 12), (90, 12)):
        R = orc.Oracle(nv, offs, lits, "ref")
        R.add_units(stream + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)])
        d_set = list(range(178 - k, 178))
        valid, rows = R.rc2(d_set, [0] * k, 1 << k, 1 << k)
        out["n%d_k%d_valid" % (n_fixed, k)] = np.asarray([valid], dtype=np.int64)
        out["n%d_k%d_rows" % (n_fixed, k)] = np.packbits(rows.astype(np.uint8), axis=1)
        # the same cube assignment by assignment through the reference's propagate()
        # (index bit j <-> d_set[j]); rc2 itself over-reports on some cubes, see DESIGN.md
        R2 = orc.Oracle(nv, offs, lits, "ref")
        R2.add_units(stream + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)])
        surv = []
        for idx in range(1 << k):
            a = [2 * (v - 1) + (0 if (idx >> j) & 1 else 1) for j, v in enumerate(d_set)]
            r, _ = R2.bcp(a, want_assigns=False)
            if not r.conflict:
                surv.append(idx)
        out["n%d_k%d_bcp_survivors" % (n_fixed, k)] = np.asarray(surv, dtype=np.int64)
        print("rc2 N=%d k=%d: rc2 reports %d, propagate() leaves %d" % (n_fixed, k, valid, len(surv)))
    np.savez_compressed(os.path.join(HERE, "rc2_golden.npz"), **out)


if __name__ == "__main__":
    odls_sample()
    latin_conseq_rows()
    rc2_golden()
