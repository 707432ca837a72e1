"""Python mirror of the tabu / simulated-annealing part of pdsat_b200/host/search.hpp, used by the
tests to check the C++ driver step for step.  Same decisions, same random draws (raw mt19937
outputs, the reference's uint_rand(gen)); the evaluation callback is supplied by the test (CPU
oracle).  Reference: src_mpi/mpi_predicter.cpp:1123-1408, 1687-1828, 2405-2510, 2697-3056."""
import math

MAX_DISTANCE_TO_RECORD = 20


class Mirror:
    def __init__(self, full_vars, evaluate, deep_predict=6, ts_strategy=0, max_L2=2, seed=0, max_steps=10, draws=None,
                 min_temperature=20.0, temperature_multiply_koef=0.9, start_temperature_koef=0.1):
        self.full = sorted(full_vars)
        self.n = len(self.full)
        self.pos = {v: i for i, v in enumerate(self.full)}
        self.evaluate = evaluate          # list of sets -> list of (predict, interrupted)
        self.deep, self.ts, self.max_L2, self.max_steps = deep_predict, ts_strategy, max_L2, max_steps
        self.draws = iter(draws(seed))    # generator of raw 32-bit mt19937 draws
        self.min_t, self.t_mul, self.t_start = min_temperature, temperature_multiply_koef, start_temperature_koef
        self.L1, self.L2 = [], []         # L2 entries: [centre tuple, checked list, partly]
        self.area = None
        self.best, self.best_set = 0.0, []
        self.centres, self.records = [], []
        self.temperature = 0.0

    def bits(self, s):
        b = [0] * self.n
        for v in s:
            b[self.pos[v]] = 1
        return tuple(b)

    def to_set(self, b):
        return [self.full[i] for i in range(self.n) if b[i]]

    @staticmethod
    def dist(a, b):
        d, first = 0, -1
        for i, (x, y) in enumerate(zip(a, b)):
            if x != y:
                if d == 0:
                    first = i
                d += 1
        return d, first

    def add_area(self, point):
        nu = [point, [0] * self.n, False]
        cnt = sum(point)
        for c in self.L1:
            if abs(sum(c) - cnt) <= 1:
                d, i = self.dist(point, c)
                if d <= 1 and point != c:
                    nu[1][i] = 1
        touched = []
        for a in self.L2:
            if abs(sum(a[0]) - cnt) <= 1:
                d, i = self.dist(point, a[0])
                if d == 0:
                    return
                if d <= 1:
                    touched.append(a)
                    a[1][i] = 1
                    nu[1][i] = 1
        for a in touched:
            if sum(a[1]) == self.n:
                self.L1.append(a[0])
                self.L2.remove(a)
        self.L2.append(nu)

    def granted(self, predict):
        if self.deep != 5:
            return False
        delta = predict - self.best
        if delta < 0:
            return True
        r = 0.0
        while r == 0:
            r = (next(self.draws) % 10) * 0.1
        return r < math.exp(-delta / self.temperature)

    def run(self, init):
        centre_set = sorted(init)
        first = True
        steps = 0
        updated = False
        swap_done = False
        while steps <= self.max_steps:
            if self.deep == 5 and not (self.temperature == 0 or self.temperature > self.min_t):
                break
            # --- tasks
            if first:
                cand = [self.bits(centre_set)]
            elif self.ts == 4 and not updated:
                c = self.area[0]
                cand = []
                for i in range(self.n):
                    if c[i]:
                        for j in range(self.n):
                            if not c[j]:
                                b = list(c)
                                b[i], b[j] = 0, 1
                                cand.append(tuple(b))
            else:
                cand = []
                for i in range(self.n):
                    if self.area[1][i]:
                        continue
                    b = list(self.area[0])
                    b[i] ^= 1
                    if sum(b):
                        cand.append(tuple(b))
            sets = sorted((self.to_set(b) for b in cand), key=len)  # stable: (size, index)
            if sets or first:
                res = self.evaluate(sets)
                self.centres.append(centre_set if self.area is None else self.to_set(self.area[0]))
                updated = False
                status = []
                for s, (pred, interrupted) in zip(sets, res):
                    st = 1 if interrupted else 4
                    if st == 4 and pred > 0 and (self.best == 0 or pred < self.best or self.granted(pred)):
                        st = 2
                        updated = True
                        self.best, self.best_set, centre_set = pred, s, s
                        self.records.append(s)
                        if self.deep == 5:
                            if self.temperature == 0.0 or self.temperature > self.best * self.t_start:
                                self.temperature = self.best * self.t_start
                            else:
                                self.temperature *= self.t_mul
                    status.append(st)
                if self.deep >= 5:
                    for s in sets:
                        self.add_area(self.bits(s))
            else:
                updated = False
            if first:
                if not self.L2:
                    self.add_area(self.bits(centre_set))
                self.area = [self.L2[0][0], list(self.L2[0][1]), self.L2[0][2]]
                first = False
            steps += 1
            # --- where next
            if updated:
                bs = self.bits(centre_set)
                cb = sum(bs)
                found = None
                self.L2 = [a for a in self.L2 if not (sum(a[0]) > cb and sum(a[0]) - cb > MAX_DISTANCE_TO_RECORD)]
                self.L1 = [c for c in self.L1 if not (sum(c) > cb and sum(c) - cb > MAX_DISTANCE_TO_RECORD)]
                for a in self.L2:
                    if a[0] == bs:
                        found = a
                self.area = [found[0], list(found[1]), found[2]] if found else [bs, [0] * self.n, False]
            elif self.ts == 4:
                if swap_done:
                    break
                swap_done = True
                continue
            else:
                if self.ts == 2:
                    for a in self.L2:
                        if a[0] == self.area[0]:
                            a[2] = True
                if self.deep < 5:
                    break
                bs = self.bits(self.best_set)
                min_d = self.n
                for a in self.L2:
                    if a[2]:
                        continue
                    d, _ = self.dist(a[0], bs)
                    if d and d < min_d:
                        min_d = d
                if min_d > self.max_L2:
                    break
                matches = [a for a in self.L2 if not a[2] and self.dist(a[0], bs)[0] == min_d]
                if not matches:
                    break
                powers = []
                for a in matches:
                    if sum(a[0]) not in powers:
                        powers.append(sum(a[0]))
                start = next(self.draws) % len(matches)
                power = powers[next(self.draws) % len(powers)]
                pick = None
                for a in matches[start:] + matches[:start]:
                    if sum(a[0]) == power:
                        pick = a
                        break
                self.area = [pick[0], list(pick[1]), pick[2]]
                centre_set = self.to_set(pick[0])
            swap_done = False
            if not centre_set:
                break
        return self
