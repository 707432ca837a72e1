"""Synthetic instance generators for the PDSAT hot path (tests + bench fixtures).

Nothing here is copied from the reference tree: the CNFs are *generated* from the
cipher / Latin-square definitions and, where the reference ships the same instance,
checked against it in tests (`tests/test_host.py::test_bivium_template_equals_reference`, only when
/root/reference exists).

* `bivium_template()`     reproduces src_common/bivium_template.cnf line for line
                          (777 vars, 12 800 clauses).
* `BiviumModel`           independent bit-level cipher model (follows the register
                          update of src_common/Bivium.cpp:69-89) -> known-answer tests.
* `odls_pair_cnf(10)`     naive order-n pair-of-diagonal-Latin-squares encoding with
                          var(p,i,j,z) = p*n^3 + i*n^2 + j*n + z + 1
                          (src_common/latin_squares.cpp:462); for n=10 it has exactly
                          2 000 vars / 434 440 clauses = header of VS_latin_conseq/in:14.
* `a5_2_cnf(...)`         Tseitin encoding of A5/2 (no A5/2 CNF ships in the reference).

Literal conventions: DIMACS ints in clause lists; `to_minisat()` converts to the
MiniSat encoding 2*var+sign (src_common/minisat/core/SolverTypes.h:46-62) used at the
C-ABI boundary.
"""
from __future__ import annotations

import itertools
from typing import Iterable, List, Sequence, Tuple

import numpy as np

Clause = List[int]


# --------------------------------------------------------------------------- DIMACS
def write_dimacs(path: str, n_vars: int, clauses: Sequence[Clause], comments: Sequence[str] = ()) -> None:
    with open(path, "w") as f:
        for c in comments:
            f.write("c " + c + "\n")
        f.write("p cnf %d %d\n" % (n_vars, len(clauses)))
        for cl in clauses:
            f.write(" ".join(map(str, cl)) + " 0\n")


def read_dimacs(path: str) -> Tuple[int, List[Clause], dict]:
    """Parse DIMACS + PDSAT's comment protocol (src_mpi/mpi_base.cpp:546-599):
    `c input variables N`, `c output variables N`, `c known_bits N`, `c var_set v1 v2 ...`."""
    n_vars = 0
    clauses: List[Clause] = []
    meta: dict = {}
    cur: Clause = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s:
                continue
            if s[0] == "c":
                w = s.split()
                if len(w) >= 4 and w[1] == "input" and w[2] == "variables":
                    meta["input_variables"] = int(w[3])
                elif len(w) >= 4 and w[1] == "output" and w[2] == "variables":
                    meta["output_variables"] = int(w[3])
                elif len(w) >= 3 and w[1] == "known_bits":
                    meta["known_bits"] = int(w[2])
                elif len(w) >= 2 and w[1] == "var_set":
                    meta["var_set"] = [int(x) for x in w[2:]]
                continue
            if s[0] == "p":
                n_vars = int(s.split()[2])
                continue
            for tok in s.split():
                v = int(tok)
                if v == 0:
                    clauses.append(cur)
                    cur = []
                else:
                    cur.append(v)
    for cl in clauses:
        for l in cl:
            n_vars = max(n_vars, abs(l))
    return n_vars, clauses, meta


def to_minisat(lit: int) -> int:
    """DIMACS literal -> MiniSat Lit.x (SolverTypes.h:46-62): 2*var + sign, var 0-based."""
    return 2 * (abs(lit) - 1) + (1 if lit < 0 else 0)


def to_csr(clauses: Sequence[Clause]) -> Tuple[np.ndarray, np.ndarray]:
    """Flat CSR (offsets uint32[n+1], lits int32 in MiniSat encoding) = the upload format."""
    offs = np.zeros(len(clauses) + 1, dtype=np.uint32)
    total = 0
    for i, cl in enumerate(clauses):
        total += len(cl)
        offs[i + 1] = total
    lits = np.empty(total, dtype=np.int32)
    p = 0
    for cl in clauses:
        for l in cl:
            lits[p] = to_minisat(l)
            p += 1
    return offs, lits


# --------------------------------------------------------------------------- Bivium
BIVIUM_LEN_A = 93
BIVIUM_LEN_B = 84
BIVIUM_STATE = 177
BIVIUM_STREAM = 200


def _xor_gate(y: int, xs: Sequence[int]) -> List[Clause]:
    """Clauses of y = xor(xs) with the last input's sign fixed by parity; sign order
    (+ before -) over (y, xs[:-1]) as in the reference template's output blocks."""
    out = []
    head = [y] + list(xs[:-1])
    for signs in itertools.product((1, -1), repeat=len(head)):
        vals = [0 if s > 0 else 1 for s in signs]  # falsifying value of each var
        vlast = 1
        for v in vals:
            vlast ^= v
        out.append([s * v for s, v in zip(signs, head)] + [xs[-1] if vlast == 0 else -xs[-1]])
    return out


def _and_xor_gate(y: int, a: int, b: int, c: int, d: int, e: int) -> List[Clause]:
    """Clauses of y = a ^ (b & c) ^ d ^ e, 24 clauses (16 five-literal, 8 six-literal)."""
    out = []
    for sy in (1, -1):
        for sa in (1, -1):
            for part in (0, 1, 2):
                for sd in (1, -1):
                    vy, va, vd = (0 if sy > 0 else 1), (0 if sa > 0 else 1), (0 if sd > 0 else 1)
                    bc = 1 if part == 2 else 0
                    ve = 1 ^ vy ^ va ^ bc ^ vd
                    mid = [b] if part == 0 else ([c] if part == 1 else [-b, -c])
                    out.append([sy * y, sa * a] + mid + [sd * d, e if ve == 0 else -e])
    return out


def bivium_template(stream_len: int = BIVIUM_STREAM) -> Tuple[int, List[Clause], List[str]]:
    """Bivium-B template: vars 1..177 state, 178..177+2*L register-feedback bits,
    then L keystream vars.  For L=200 this equals the reference's template."""
    reg = list(range(1, BIVIUM_STATE + 1))
    first_out = BIVIUM_STATE + 2 * stream_len + 1
    gates: List[Clause] = []
    outs: List[Clause] = []
    for i in range(stream_len):
        outs += _xor_gate(first_out + i, [reg[65], reg[92], reg[161], reg[176]])
        t2 = BIVIUM_STATE + 1 + 2 * i
        t1 = t2 + 1
        gates += _and_xor_gate(t2, reg[161], reg[174], reg[175], reg[176], reg[68])
        gates += _and_xor_gate(t1, reg[65], reg[90], reg[91], reg[92], reg[170])
        reg = [t2] + reg[0:92] + [t1] + reg[93:176]
    n_vars = BIVIUM_STATE + 3 * stream_len
    comments = ["input variables %d" % BIVIUM_STATE, "output variables %d" % stream_len]
    return n_vars, gates + outs, comments


class BiviumModel:
    """Bit-level Bivium-B: 177-bit register, update as src_common/Bivium.cpp:69-89."""

    def __init__(self, state: Sequence[int]):
        assert len(state) == BIVIUM_STATE
        self.reg = [int(b) & 1 for b in state]

    def _shift(self) -> None:
        r = self.reg
        t1 = r[65] ^ (r[90] & r[91]) ^ r[92] ^ r[170]
        t2 = r[161] ^ (r[174] & r[175]) ^ r[176] ^ r[68]
        self.reg = [t2] + r[0:92] + [t1] + r[93:176]

    def next_bit(self) -> int:
        r = self.reg
        z = r[65] ^ r[92] ^ r[161] ^ r[176]
        self._shift()
        return z

    def stream(self, n: int) -> List[int]:
        return [self.next_bit() for _ in range(n)]

    def trace(self, n: int) -> List[int]:
        """Values of all 177+3n template variables for this start state."""
        state0 = list(self.reg)
        m = BiviumModel(state0)
        fb = []
        ks = []
        for _ in range(n):
            r = m.reg
            ks.append(r[65] ^ r[92] ^ r[161] ^ r[176])
            t1 = r[65] ^ (r[90] & r[91]) ^ r[92] ^ r[170]
            t2 = r[161] ^ (r[174] & r[175]) ^ r[176] ^ r[68]
            fb += [t2, t1]
            m._shift()
        return state0 + fb + ks


# --------------------------------------------------------------------------- ODLS
def odls_var(n: int, p: int, i: int, j: int, z: int) -> int:
    return p * n ** 3 + i * n ** 2 + j * n + z + 1


def odls_pair_cnf(n: int = 10) -> Tuple[int, List[Clause]]:
    """Pair of orthogonal diagonal Latin squares of order n, naive encoding.
    n=10 -> 2 000 vars, 640 ALO + 28 800 AMO + 405 000 orthogonality = 434 440 clauses."""
    alo: List[Clause] = []
    amo: List[Clause] = []
    for p in range(2):
        groups = []
        for i in range(n):
            for j in range(n):
                groups.append([odls_var(n, p, i, j, z) for z in range(n)])  # cell
        for i in range(n):
            for z in range(n):
                groups.append([odls_var(n, p, i, j, z) for j in range(n)])  # row
        for j in range(n):
            for z in range(n):
                groups.append([odls_var(n, p, i, j, z) for i in range(n)])  # column
        for z in range(n):
            groups.append([odls_var(n, p, i, i, z) for i in range(n)])  # main diagonal
        for z in range(n):
            groups.append([odls_var(n, p, i, n - 1 - i, z) for i in range(n)])  # anti-diagonal
        for g in groups:
            alo.append(list(g))
            for a, b in itertools.combinations(g, 2):
                amo.append([-a, -b])
    orth: List[Clause] = []
    cells = [(i, j) for i in range(n) for j in range(n)]
    for a in range(n):
        for b in range(n):
            for c1, c2 in itertools.combinations(cells, 2):
                if c1[0] == c2[0] or c1[1] == c2[1]:
                    continue
                orth.append([-odls_var(n, 0, c1[0], c1[1], a), -odls_var(n, 1, c1[0], c1[1], b),
                             -odls_var(n, 0, c2[0], c2[1], a), -odls_var(n, 1, c2[0], c2[1], b)])
    return 2 * n ** 3, alo + amo + orth


def odls_assumptions_from_squares(n: int, sq_a: Sequence[Sequence[int]], sq_b: Sequence[Sequence[int]]) -> List[int]:
    """Positive cell literals of a pair (src_common/latin_squares.cpp:145-150: assumptions
    are always positive literals mkLit(v-1))."""
    out = []
    for p, sq in enumerate((sq_a, sq_b)):
        for i in range(n):
            for j in range(n):
                out.append(odls_var(n, p, i, j, int(sq[i][j])))
    return out


def parse_odls_pairs(text: str, n: int = 10) -> List[Tuple[List[List[int]], List[List[int]]]]:
    """Parse the text-grid pair format of src_common/ODLS_10_pairs.txt: blocks of n rows,
    each row = n digits of square A, whitespace, n digits of square B."""
    rows = []
    for line in text.splitlines():
        toks = line.split()
        digs = "".join(toks)
        if len(digs) == 2 * n and digs.isdigit():
            rows.append(([int(c) for c in digs[:n]], [int(c) for c in digs[n:]]))
    pairs = []
    for k in range(0, len(rows) - n + 1, n):
        a = [r[0] for r in rows[k:k + n]]
        b = [r[1] for r in rows[k:k + n]]
        pairs.append((a, b))
    return pairs


# --------------------------------------------------------------------------- A5/2
class _Tseitin:
    def __init__(self, n_vars: int):
        self.n = n_vars
        self.clauses: List[Clause] = []

    def new(self) -> int:
        self.n += 1
        return self.n

    def xor(self, xs: Sequence[int]) -> int:
        y = self.new()
        self.clauses += _xor_gate(y, list(xs))
        return y

    def and2(self, a: int, b: int) -> int:
        y = self.new()
        self.clauses += [[-y, a], [-y, b], [y, -a, -b]]
        return y

    def maj(self, a: int, b: int, c: int) -> int:
        y = self.new()
        self.clauses += [[-y, a, b], [-y, a, c], [-y, b, c], [y, -a, -b], [y, -a, -c], [y, -b, -c]]
        return y

    def mux(self, s: int, a: int, b: int) -> int:
        """y = s ? a : b"""
        y = self.new()
        self.clauses += [[-s, -a, y], [-s, a, -y], [s, -b, y], [s, b, -y]]
        return y


A52_LEN = (19, 22, 23, 17)
# feedback taps (bit indices inside each register; register sizes as src_common/a5_2.h:37-71,
# with R4 actually clocked and indexed within its 17 bits)
A52_TAPS = ((18, 17, 16, 13), (21, 20), (22, 21, 20, 7), (16, 11))
A52_CLK = (10, 3, 7)          # R4 bits controlling R1, R2, R3 (majority rule)
A52_OUT = ((12, 14, 15), (9, 13, 16), (13, 16, 18))  # majority inputs (second of each is complemented)
A52_COMPL = (14, 16, 13)


def _a52_step_bits(regs):
    r1, r2, r3, r4 = regs
    m = (r4[A52_CLK[0]] + r4[A52_CLK[1]] + r4[A52_CLK[2]]) >= 2
    new = []
    for idx, (r, taps) in enumerate(((r1, A52_TAPS[0]), (r2, A52_TAPS[1]), (r3, A52_TAPS[2]))):
        if r4[A52_CLK[idx]] == m:
            fb = 0
            for t in taps:
                fb ^= r[t]
            r = [fb] + r[:-1]
        new.append(r)
    fb = 0
    for t in A52_TAPS[3]:
        fb ^= r4[t]
    new.append([fb] + r4[:-1])
    return new


class A52Model:
    """Bit-level A5/2 keystream generator from an 81-bit post-initialisation state
    (R1 19, R2 22, R3 23, R4 17)."""

    def __init__(self, state: Sequence[int]):
        assert len(state) == sum(A52_LEN)
        s = [int(b) & 1 for b in state]
        self.regs = [s[0:19], s[19:41], s[41:64], s[64:81]]

    def stream(self, n: int) -> List[int]:
        out = []
        for _ in range(n):
            self.regs = _a52_step_bits(self.regs)
            z = 0
            for idx, top in enumerate((18, 21, 22)):
                r = self.regs[idx]
                a, b, c = A52_OUT[idx]
                bits = [r[a], r[b], r[c]]
                bits[1] ^= 1
                z ^= (1 if sum(bits) >= 2 else 0) ^ r[top]
            out.append(z)
        return out


def a5_2_cnf(stream_len: int = 114) -> Tuple[int, List[Clause], List[str], List[int]]:
    """Tseitin CNF of A5/2: vars 1..64 = R1|R2|R3, 65..81 = R4, then gate vars; returns
    (n_vars, clauses, comments, keystream_vars)."""
    t = _Tseitin(81)
    r1 = list(range(1, 20))
    r2 = list(range(20, 42))
    r3 = list(range(42, 65))
    r4 = list(range(65, 82))
    ks = []
    for _ in range(stream_len):
        m = t.maj(r4[A52_CLK[0]], r4[A52_CLK[1]], r4[A52_CLK[2]])
        regs = [r1, r2, r3]
        new = []
        for idx, r in enumerate(regs):
            # clocked iff control bit == majority
            neq = t.xor([r4[A52_CLK[idx]], m])
            fb = t.xor([r[x] for x in A52_TAPS[idx]])
            shifted = [fb] + r[:-1]
            new.append([t.mux(neq, r[i], shifted[i]) for i in range(len(r))])
        r1, r2, r3 = new
        fb4 = t.xor([r4[x] for x in A52_TAPS[3]])
        r4 = [fb4] + r4[:-1]
        parts = []
        for idx, (r, top) in enumerate(((r1, 18), (r2, 21), (r3, 22))):
            a, b, c = A52_OUT[idx]
            mj = t.maj(r[a], -r[b], r[c])
            parts += [mj, r[top]]
        ks.append(t.xor(parts))
    comments = ["input variables 81", "output variables %d" % stream_len]
    return t.n, t.clauses, comments, ks


def a5_2_weakened_instance(n_known: int = 32, seed: int = 9, stream_len: int = 114):
    """BASELINE config 4: A5/2 CNF + keystream of a secret state + R4 + n_known of the 64
    R1|R2|R3 bits as fixed assumptions; the other 64 - n_known state variables form the
    decomposition set.  Returns (n_vars, clauses, fixed MiniSat literals, set variables (1-based,
    ascending), secret state bits)."""
    nv, cl, _, ks = a5_2_cnf(stream_len)
    rng = np.random.default_rng(seed)
    secret = rng.integers(0, 2, 81)
    stream = A52Model(secret).stream(stream_len)
    fixed = [2 * (v - 1) + (0 if bit else 1) for v, bit in zip(ks, stream)]
    fixed += [2 * v + (0 if secret[v] else 1) for v in range(64, 81)]
    known = sorted(int(v) for v in rng.choice(np.arange(64), size=n_known, replace=False))
    fixed += [2 * v + (0 if secret[v] else 1) for v in known]
    dset = [v + 1 for v in range(64) if v not in known]
    return nv, cl, fixed, dset, secret
