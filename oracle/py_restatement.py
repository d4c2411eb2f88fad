"""Independent pure-Python restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/cv_oracle.h).  This file
restates the same reference lines as cv_oracle.c but was written separately,
object by object as the Rust is structured (HMM accessors -> MetaElements ->
CPSolver), so that a misreading in one restatement shows up as a mismatch
between the two.  Pure Python floats are IEEE binary64; every `x + y` below is
exactly one rounded add, as in the Rust.  Use on tiny instances only.
"""
from __future__ import annotations

import itertools
import math

NEG_INF = -math.inf


def argmax(v):
    """ndarray-stats 0.5 QuantileExt::argmax: first strictly-greater wins; NaN -> error."""
    if len(v) == 0:
        raise ValueError("EmptyInput")
    cur, idx = v[0], 0
    for j, x in enumerate(v):
        if math.isnan(x) or math.isnan(cur):
            raise ValueError("UndefinedOrder")
        if x > cur:
            cur, idx = x, j
    return idx


class HMM:
    """hmm.rs:10-18, accessors hmm.rs:207-234. a[from][to], b[state][obs], pi[state]."""

    def __init__(self, a, b, pi):
        self.a = [list(map(float, r)) for r in a]
        self.b = [list(map(float, r)) for r in b]
        self.pi = list(map(float, pi))

    def nstates(self):
        return len(self.a)

    def init_prob(self, state, obs):           # hmm.rs:211-213
        return self.pi[state] + self.b[state][obs]

    def init_probs(self, obs):                 # hmm.rs:215-218
        return [self.pi[s] + self.b[s][obs] for s in range(self.nstates())]

    def transition_prob(self, f, t, obs):      # hmm.rs:220-222
        return self.a[f][t] + self.b[t][obs]

    def transitions_to(self, t):               # hmm.rs:224-226
        return [self.a[f][t] for f in range(self.nstates())]

    def emit_prob(self, state, obs):           # hmm.rs:228-230
        return self.b[state][obs]


def decode(sequence, hmm):
    """viterbi.rs:5-32. Returns (path, delta_rows, bt_rows)."""
    T, K = len(sequence), hmm.nstates()
    if T == 0:
        raise IndexError("empty sequence")
    array = [[0.0] * K for _ in range(T)]
    bt = [[0] * K for _ in range(T)]
    for t in range(1, T):
        for to in range(K):
            e = hmm.emit_prob(to, sequence[t])
            if e > NEG_INF:
                prev = array[t - 1]
                tr = hmm.transitions_to(to)
                probs = [prev[j] + tr[j] for j in range(K)]
                sf = argmax(probs)
                array[t][to] = probs[sf] + e
                bt[t][to] = sf
            else:
                array[t][to] = NEG_INF
    end = argmax(array[T - 1])
    pred = [0] * T
    pred[T - 1] = end
    for t in range(T - 2, -1, -1):
        end = bt[t + 1][end]
        pred[t] = end
    return pred, array, bt


class Element:
    """MetaElements, viterbi_solver/utils.rs:8-47 (only the fields the solver reads)."""

    def __init__(self, t, value, comp, active):
        self.t, self.value, self.constraint_component, self.active = t, value, comp, active

    def arc_p(self, hmm, f, s):                # utils.rs:24-30
        if self.t == 0:
            return hmm.init_prob(s, self.value)
        return hmm.transition_prob(f, s, self.value)

    def transitions(self, hmm, s):             # utils.rs:32-38
        if self.t == 0:
            return [hmm.pi[s]] * hmm.nstates()
        return [hmm.a[j][s] for j in range(hmm.nstates())]

    def is_constrained(self):                  # utils.rs:44-46
        return self.active


class CPSolver:
    """cp.rs:8-152."""

    def __init__(self, hmm, elements, ncomp, max_nodes=0):
        self.hmm, self.seq = hmm, elements
        self.constraints = [[] for _ in range(ncomp)]          # cp.rs:21
        for t, el in enumerate(elements):                       # cp.rs:22-27
            if el.is_constrained():
                self.constraints[el.constraint_component].append(t)
        self.choices = [None] * ncomp                           # cp.rs:28
        self.best_obj = NEG_INF
        self.best_sol = [0] * len(elements)
        self.explored = 0
        self.max_nodes = max_nodes
        self.steps = 0
        self.ub_log = []
        self.state_log = []
        self.keep_states = False

    def _fixed(self, t):
        el = self.seq[t]
        return el.is_constrained() and self.choices[el.constraint_component] is not None

    def _step(self, array, bt, t):
        K = self.hmm.nstates()
        for s in range(K):
            prev = array[t - 1]
            tr = self.seq[t].transitions(self.hmm, s)
            probs = [prev[j] + tr[j] for j in range(K)]
            sf = argmax(probs)
            arc = self.seq[t].arc_p(self.hmm, sf, s)
            array[t][s] = prev[sf] + arc
            bt[t][s] = sf
        self.steps += 1

    def viterbi_from(self, array, bt, frm, node):               # cp.rs:32-61
        K = self.hmm.nstates()
        array[frm] = [NEG_INF] * K
        array[frm][node] = 0.0
        if frm != 0:
            prev = array[frm - 1]
            tr = self.seq[frm].transitions(self.hmm, node)
            bt[frm][node] = argmax([prev[j] + tr[j] for j in range(K)])
        if frm + 1 < len(self.seq) and self._fixed(frm + 1):
            bt[frm + 1][self.choices[self.seq[frm + 1].constraint_component]] = node
        t = frm + 1
        while t < len(self.seq) and not self._fixed(t):
            self._step(array, bt, t)
            t += 1

    def init_viterbi(self, array, bt):                          # cp.rs:63-83
        t = 0
        while t < len(self.seq) and not self.seq[t].is_constrained():
            if t == 0:
                array[0] = self.hmm.init_probs(self.seq[0].value)
            else:
                self._step(array, bt, t)
            t += 1

    def backtrack(self, array, bt, obj):                        # cp.rs:85-93
        cur = argmax(array[len(self.seq) - 1])
        assert obj > self.best_obj
        self.best_obj = obj
        for t in range(len(self.seq) - 1, -1, -1):
            self.best_sol[t] = cur
            cur = bt[t][cur]

    def solve_r(self, array, bt, comp):                         # cp.rs:95-126
        for state in range(self.hmm.nstates()):
            if self.max_nodes and self.explored >= self.max_nodes:
                break
            self.explored += 1
            self.choices[comp] = state
            for pos in self.constraints[comp]:
                self.viterbi_from(array, bt, pos, state)
            ub = 0.0
            for cid in range(comp + 1):
                st = self.choices[cid]
                for t in self.constraints[cid]:
                    if t == 0:
                        ub += self.seq[0].arc_p(self.hmm, 0, st)
                    else:
                        sf = bt[t][st]
                        arc = self.seq[t].arc_p(self.hmm, sf, st)
                        ub += array[t - 1][sf] + arc
            self.ub_log.append(ub)
            if self.keep_states:
                self.state_log.append(([r[:] for r in array], [r[:] for r in bt]))
            if ub > self.best_obj:
                if comp + 1 < len(self.constraints):
                    self.solve_r(array, bt, comp + 1)
                else:
                    self.backtrack(array, bt, ub)
        self.choices[comp] = None

    def solve(self):                                            # cp.rs:133-143
        N, K = len(self.seq), self.hmm.nstates()
        array = [[0.0] * K for _ in range(N)]
        bt = [[0] * K for _ in range(N)]
        self.init_viterbi(array, bt)
        if len(self.constraints) > 0:
            self.solve_r(array, bt, 0)
        else:
            obj = max(array[N - 1])
            self.backtrack(array, bt, obj)
        self.final_state = (array, bt)


# --------------------------------------------------------------------------
# brute-force known-answer helpers (enumerate all K^T state paths)
# --------------------------------------------------------------------------

def r1_path_score(path, sequence, hmm):
    """Score of one state path under viterbi.rs's association: ((d + a) + b), d0 = 0.0."""
    d = 0.0
    for t in range(1, len(sequence)):
        e = hmm.emit_prob(path[t], sequence[t])
        if not (e > NEG_INF):
            return NEG_INF
        d = (d + hmm.a[path[t - 1]][path[t]]) + e
    return d


def r1_bruteforce_best(sequence, hmm):
    """max over all K^T paths of r1_path_score. fl(x + c) is monotone in x, so the
    DP value delta[T-1][end] must equal this bit-for-bit."""
    K, T = hmm.nstates(), len(sequence)
    best = NEG_INF
    for p in itertools.product(range(K), repeat=T):
        s = r1_path_score(p, sequence, hmm)
        if s > best:
            best = s
    return best


def r2_chain_score(path, elements, hmm):
    """Score of a state path along the R2 chain: d0 = pi + b, then d + (a + b)
    (or d + (pi + b) at sequence starts)."""
    d = hmm.init_prob(path[0], elements[0].value)
    for t in range(1, len(elements)):
        d = d + elements[t].arc_p(hmm, path[t - 1], path[t])
    return d
