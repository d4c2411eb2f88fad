"""Host-side problem assembly, mirroring the reference so that the GPU solver sees exactly the
inputs CPSolver::new sees:

  load_sequences / load_tags   src/utils.rs:7-60
  Constraints.from_tags        src/viterbi_solver/constraints.rs:40-69
  SuperSequence                src/viterbi_solver/utils.rs:49-211 (from, recompute_constraints,
                               reorder, parse_solution, number_constraints, Index)

recompute_constraints(prop) draws `rng.gen::<f64>() <= prop` from rand 0.8 StdRng seeded with 3019
(utils.rs:101,170).  For prop in {0, 1} the outcome does not depend on the stream (gen::<f64>() is in
[0,1)); for 0 < prop < 1 the stream comes from `StdRng` below, a restatement of ChaCha12 + the PCG32 seed expansion.
No Rust toolchain exists here, but the restatement is pinned by the value-stability vectors the crates
themselves publish in their test suites (tests/test_cli.py::test_stdrng_matches_published_rand_vectors):
rand 0.8 `rngs::std::test_stdrng_construction`, rand_chacha 0.3 `test_chacha_construction`, rand_pcg 0.3
`test_lcg64xsh32_construction` (seed_from_u64) and rand 0.8 `distributions::float::value_stability` (f64).
"""
from __future__ import annotations

import numpy as np

from .hmm import HMM


class StdRng:
    """rand 0.8 `StdRng::seed_from_u64` + `gen::<f64>()` (ChaCha12, key expanded from the u64 with PCG32), the
    same restatement as host/rng_chacha12.h; it only matters for 0 < prop < 1.  `ROUNDS` = 12 (StdRng of
    rand 0.8 = ChaCha12Rng); the value-stability test also runs the 20-round variant the crate publishes."""

    ROUNDS = 12

    @staticmethod
    def seed_words_from_u64(state: int, nwords: int = 8):
        """rand_core 0.6 `SeedableRng::seed_from_u64`: a PCG32 (XSH-RR) stream fills the seed, 4 bytes at a time."""
        M64 = (1 << 64) - 1
        state, words = state & M64, []
        for _ in range(nwords):
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            words.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
        return words

    def __init__(self, seed: int):
        self.key, self.counter, self.buf, self.idx = self.seed_words_from_u64(seed), 0, [], 16

    @classmethod
    def from_seed(cls, seed: bytes):
        """`SeedableRng::from_seed` with the 32-byte ChaCha key (little-endian words)."""
        assert len(seed) == 32
        r = cls(0)
        r.key = [int.from_bytes(seed[4 * i: 4 * i + 4], "little") for i in range(8)]
        return r

    @classmethod
    def from_rng(cls, other: "StdRng"):
        """`SeedableRng::from_rng`: the new key is the next 32 bytes of `other` (fill_bytes = consecutive words)."""
        words = []
        for _ in range(4):
            v = other.next_u64()
            words += [v & 0xFFFFFFFF, v >> 32]
        return cls.from_seed(b"".join(w.to_bytes(4, "little") for w in words))

    def _block(self):
        M = 0xFFFFFFFF
        inp = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self.key + [self.counter & M, (self.counter >> 32) & M, 0, 0]
        x = list(inp)

        def rotl(v, n):
            return ((v << n) | (v >> (32 - n))) & M

        def qr(a, b, c, d):
            x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 16)
            x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 12)
            x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 8)
            x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 7)

        for _ in range(self.ROUNDS // 2):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.buf = [(x[i] + inp[i]) & M for i in range(16)]
        self.counter += 1
        self.idx = 0

    def next_u64(self):
        if self.idx >= 16:
            self._block()
        lo, hi = self.buf[self.idx], self.buf[self.idx + 1]
        self.idx += 2
        return (hi << 32) | lo

    @staticmethod
    def f64_from_u64(v: int) -> float:
        """rand 0.8 `Standard` for f64: 53 random bits scaled into [0, 1)."""
        return float(v >> 11) * (1.0 / 9007199254740992.0)

    def gen_f64(self):
        return self.f64_from_u64(self.next_u64())


def load_sequences(path, D=2):
    """utils.rs:7-34: lines `seq_id f1 .. fD` split on single spaces; a new sequence starts whenever
    seq_id changes; missing features stay 0."""
    ret, cur, last = [], [], None
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            s = [int(x) for x in line.split(" ")]
            if last is not None and s[0] != last:
                ret.append(cur)
                cur = []
            last = s[0]
            el = [0] * D
            for i in range(1, len(s)):
                el[i - 1] = s[i]          # IndexError if more than D features, like the Rust array index
            cur.append(el)
    ret.append(cur)
    return ret


def load_tags(path):
    """utils.rs:36-60: lines `seq_id tag`, `-1` -> None."""
    ret, cur, last = [], [], None
    with open(path) as f:
        for line in f:
            s = line.rstrip("\n").split(" ")
            tid = int(s[0])
            if last is not None and tid != last:
                ret.append(cur)
                cur = []
            cur.append(None if s[1] == "-1" else int(s[1]))
            last = tid
    ret.append(cur)
    return ret


class Constraints:
    """constraints.rs:6-70. components: list of sets of (seq_id, t)."""

    def __init__(self, components):
        self.components = components

    @classmethod
    def from_tags(cls, truth):
        comp_value, components = [], []
        for seq_id, tags in enumerate(truth):
            for t, tag in enumerate(tags):
                if tag is None:
                    continue
                if tag in comp_value:
                    components[comp_value.index(tag)].add((seq_id, t))
                else:
                    comp_value.append(tag)
                    components.append({(seq_id, t)})
        return cls(components)


class SuperSequence:
    """utils.rs:49-211.  Element arrays (numpy): seq, t, value[D], comp (constraint_component), active."""

    def __init__(self, sequences, constraints: Constraints, hmm: HMM):
        self.sequences, self.constraints, self.hmm = sequences, constraints, hmm
        sizes = [len(s) for s in sequences]
        N = int(sum(sizes))
        self.orig_seq_sizes = sizes
        self.super_seq_start = list(np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(int)) if sizes else []
        D = len(hmm.bdims)
        self.seq = np.zeros(N, dtype=np.int64)
        self.t = np.zeros(N, dtype=np.int64)
        self.value = np.zeros((N, D), dtype=np.int64)
        self.comp = np.full(N, -1, dtype=np.int32)
        lookup = {}
        for cid, members in enumerate(constraints.components):     # first component containing (seq,t) wins
            for m in members:
                lookup.setdefault(m, cid)
        i = 0
        for sid, s in enumerate(sequences):
            for tt, v in enumerate(s):
                self.seq[i], self.t[i] = sid, tt
                self.value[i, : len(v)] = v[:D]
                self.comp[i] = lookup.get((sid, tt), -1)
                i += 1
        self.active = self.comp >= 0                                # utils.rs:76
        self.rng = StdRng(3019)                                     # utils.rs:101
        self._count_active()

    def _count_active(self):                                        # utils.rs:88-99 / 152-163
        self.nb_active_cstr = int(len(set(int(c) for c in self.comp[self.active])))

    def __len__(self):
        return len(self.seq)

    def number_constraints(self):                                   # utils.rs:200-202
        return self.nb_active_cstr

    def _ordering(self):                                            # utils.rs:105-136
        K = self.hmm.nstates()
        flat = self.hmm.flatten_obs(self.value) if len(self) else np.zeros(0, dtype=np.uint32)
        emit_ok = (self.hmm.b.reshape(K, -1)[:, flat] > -np.inf).sum(axis=0).astype(np.float64)
        keys = []
        for sid in range(len(self.super_seq_start)):
            start, size = self.super_seq_start[sid], self.orig_seq_sizes[sid]
            possible = 0.0
            for v in emit_ok[start:start + size]:                   # f64 accumulation in element order
                possible += float(v)
            last_active = bool(self.active[start + size - 1]) if size else False   # flag of the LAST element (Q8)
            keys.append((1 if last_active else 0, possible / size if size else float("nan"), sid))
        keys.sort()                                                 # stable; tuples compare like partial_cmp
        return [k[2] for k in keys]

    def reorder(self):                                              # utils.rs:138-165
        order = self._ordering()
        idx = []
        new_start = list(self.super_seq_start)
        pos = 0
        for sid in order:
            start, size = self.super_seq_start[sid], self.orig_seq_sizes[sid]
            new_start[sid] = pos
            idx.extend(range(start, start + size))
            pos += size
        idx = np.array(idx, dtype=np.int64)
        for name in ("seq", "t", "value", "comp", "active"):
            setattr(self, name, getattr(self, name)[idx])
        self.super_seq_start = new_start
        self._count_active()

    def recompute_constraints(self, proportion, uniform=None):      # utils.rs:168-177
        if uniform is None:
            uniform = self.rng.gen_f64                              # StdRng(3019) stream, state carries over calls
        act = np.zeros(len(self), dtype=bool)
        for i in range(len(self)):
            if self.comp[i] != -1:                                  # RNG consumed only for constrained elements
                act[i] = uniform() <= proportion
        self.active = act
        self.reorder()

    def parse_solution(self, solution):                             # utils.rs:183-190
        sol = [np.zeros(n, dtype=np.uint64) for n in self.orig_seq_sizes]
        for i in range(len(self)):
            sol[self.seq[i]][self.t[i]] = solution[i]
        return sol

    # ---- the arrays the C ABI takes ----
    def solver_inputs(self):
        obs = self.hmm.flatten_obs(self.value) if len(self) else np.zeros(0, dtype=np.uint32)
        start = (self.t == 0).astype(np.uint8)
        comp = np.where(self.active, self.comp, -1).astype(np.int32)
        return obs, start, comp, self.number_constraints()
