"""TEST INFRASTRUCTURE ONLY -- a pure-Python model of the SHARDED constrained solve (csrc/cp_dist.cuh).

It restates the ownership rules of the multi-GPU path on top of the literal restatement of cp.rs
(py_restatement.CPSolver): every rank keeps delta rows [lo, hi) and psi rows (lo, hi]; rows it does not own
are POISONED (NaN / None), so any read across a cut blows up instead of silently working.  The ranks exchange
only (1) the bound terms of a node, (2) one K-entry backtrack map per rank, (3) their rows of the solution --
through `exchange(obj) -> list of every rank's obj` (torch.distributed.all_gather_object in the gloo tests).
If the sharded run reproduces the sequential solver's ub trace, node count, objective and solution, the
partition argument (no sweep crosses a position of component 0, cp.rs:48) holds on that instance.
"""
from __future__ import annotations

from .py_restatement import NEG_INF, CPSolver, argmax

NAN = float("nan")


def plan_cuts(comp, N, R):
    """Same rule as cp_plan_cuts (csrc/cp_host.inl); returns None when the problem cannot be cut."""
    cuts = [0] * (R + 1)
    cuts[R] = N
    if R == 1:
        return cuts
    c0 = [t for t in range(1, N) if comp[t] == 0]
    import bisect

    for r in range(1, R):
        want = N // R * r + (N % R) * r // R
        i = bisect.bisect_left(c0, want)
        best = -1
        if i < len(c0):
            best = c0[i]
        if i > 0 and (best < 0 or want - c0[i - 1] <= best - want):
            best = c0[i - 1]
        if best <= cuts[r - 1]:
            j = bisect.bisect_right(c0, cuts[r - 1])
            if j == len(c0):
                return None
            best = c0[j]
        cuts[r] = best
    for r in range(1, R + 1):
        if cuts[r] <= cuts[r - 1]:
            return None
    return cuts


class ShardedCPSolver(CPSolver):
    def __init__(self, hmm, elements, ncomp, rank, R, cuts, exchange, max_nodes=0):
        super().__init__(hmm, elements, ncomp, max_nodes)
        self.rank, self.R, self.exchange = rank, R, exchange
        self.lo, self.hi = cuts[rank], cuts[rank + 1]
        self.N = len(elements)

    # -- one node = phases A+B, C1, C2 on the rows this rank owns (SURVEY Q9 + cp_dist.cuh) --
    def _node(self, array, bt, comp, state):
        K = self.hmm.nstates()
        lo, hi, N = self.lo, self.hi, self.N
        pos_all = self.constraints[comp]
        for pos in pos_all:                                     # A + B: sweeps that start in [lo, hi)
            if not (lo <= pos < hi):
                continue
            array[pos] = [NEG_INF] * K
            array[pos][state] = 0.0
            t = pos + 1
            while t < N and not self._fixed(t):
                assert t < hi, "a sweep crossed a cut"
                self._step(array, bt, t)
                t += 1
        for pos in pos_all:                                     # C1, owner of pos
            if lo <= pos < hi and pos + 1 < N and self._fixed(pos + 1):
                c1 = self.seq[pos + 1].constraint_component
                if c1 != comp:                                  # same component: rewritten by its own C2
                    bt[pos + 1][self.choices[c1]] = state
        for pos in pos_all:                                     # C2, owner of row pos-1
            if pos != 0 and lo < pos <= hi:
                prev = array[pos - 1]
                tr = self.seq[pos].transitions(self.hmm, state)
                bt[pos][state] = argmax([prev[j] + tr[j] for j in range(K)])

    def _local_terms(self, array, bt, comp):
        out = []
        g = 0
        for cid in range(comp + 1):
            st = self.choices[cid]
            for t in self.constraints[cid]:
                if t == 0 and self.lo == 0:
                    out.append((g, self.seq[0].arc_p(self.hmm, 0, st)))
                elif t != 0 and self.lo < t <= self.hi:
                    sf = bt[t][st]
                    out.append((g, array[t - 1][sf] + self.seq[t].arc_p(self.hmm, sf, st)))
                g += 1
        return out, g

    def solve_r(self, array, bt, comp):
        for state in range(self.hmm.nstates()):
            if self.max_nodes and self.explored >= self.max_nodes:
                break
            self.explored += 1
            self.choices[comp] = state
            self._node(array, bt, comp, state)
            mine, nterms = self._local_terms(array, bt, comp)
            terms = [None] * nterms
            for part in self.exchange(mine):                   # every rank's (index, term) pairs
                for g, v in part:
                    assert terms[g] is None
                    terms[g] = v
            ub = 0.0
            for v in terms:                                     # the reference's order (cp.rs:103-116)
                ub += v
            self.ub_log.append(ub)
            if ub > self.best_obj:
                if comp + 1 < len(self.constraints):
                    self.solve_r(array, bt, comp + 1)
                else:
                    self.backtrack(array, bt, ub)
        self.choices[comp] = None

    def backtrack(self, array, bt, obj):
        assert obj > self.best_obj
        self.best_obj = obj
        K, last = self.hmm.nstates(), self.rank == self.R - 1
        rtop = self.N - 1 if last else self.hi

        def walk(cur, write):
            for t in range(rtop, self.lo, -1):
                if write and t < self.hi:
                    self.best_sol[t] = cur
                cur = bt[t][cur]
            if write:
                self.best_sol[self.lo] = cur
            return cur

        mymap = [walk(e, False) for e in range(K)]
        end = argmax(array[self.N - 1]) if last else None
        allmaps = self.exchange((mymap, end))
        cur = allmaps[self.R - 1][1]
        for q in range(self.R - 1, self.rank, -1):
            cur = allmaps[q][0][cur]
        walk(cur, True)

    def solve(self):
        N, K = self.N, self.hmm.nstates()
        lo, hi = self.lo, self.hi
        array = [[0.0] * K if lo <= t < hi else [NAN] * K for t in range(N)]
        bt = [[0] * K if lo < t <= hi or (t == 0 and lo == 0) else [None] * K for t in range(N)]
        if lo == 0:
            self.init_viterbi(array, bt)
        assert len(self.constraints) > 0
        self.solve_r(array, bt, 0)
        parts = self.exchange(self.best_sol[lo:hi])
        self.best_sol = [s for part in parts for s in part]
        self.final_state = (array, bt)
