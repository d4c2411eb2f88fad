"""Host driver (host/cv_viterbi, C++ mirror of main.rs): file formats, problem assembly and output file.
CPU: the assembled solver inputs equal the Python SuperSequence mirror's (dry run, no GPU call).
GPU: the written result file equals what the oracle gives on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import consistent_viterbi_b200 as cv
from oracle import pyoracle as po
from util import random_hmm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "cv_viterbi")


def _dataset(tmp_path, seed=5, nseq=9):
    rng = np.random.default_rng(seed)
    K, bd = 5, (4, 3)
    A, B, pi = random_hmm(rng, K, bd[0] * bd[1], zero_frac=0.1)
    hmm = cv.HMM(A, B.reshape(K, *bd), pi)
    seqs = [[[int(rng.integers(0, bd[0])), int(rng.integers(0, bd[1]))] for _ in range(int(rng.integers(2, 15)))] for _ in range(nseq)]
    ctl = [[(int(rng.integers(0, 3)) if rng.random() < 0.2 else None) for _ in s] for s in seqs]
    d = tmp_path / "data"
    d.mkdir()
    with open(d / "sequences", "w") as f, open(d / "tags", "w") as g, open(d / "test_tags", "w") as t:
        for i, s in enumerate(seqs):
            for k, v in enumerate(s):
                f.write(f"{i + 3} {v[0]} {v[1]}\n")          # sequence ids need not start at 0
                g.write(f"{i + 3} 0\n")
                t.write(f"{i + 3} {-1 if ctl[i][k] is None else ctl[i][k]}\n")
    hmm.write(d / "hmm.json")
    return d, hmm, seqs, ctl, (A, B, pi)


def _run(d, out, prop, env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([EXE, "-i", str(d), "-o", str(out), "-n", "5", "-b", "4", "3", "-p", prop], env=env,
                          capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(not os.path.exists(EXE), reason="host driver not built")
@pytest.mark.parametrize("prop", ["1", "0", "0.5"])
def test_cli_assembly_matches_python_mirror(tmp_path, prop):
    d, hmm, seqs, ctl, _ = _dataset(tmp_path)
    dump = tmp_path / "inputs.txt"
    r = _run(d, tmp_path, prop, {"CV_DRY_RUN": "1", "CV_DUMP_INPUTS": str(dump)})
    assert r.returncode == 0, r.stderr
    ss = cv.SuperSequence(seqs, cv.Constraints.from_tags(ctl), hmm)
    ss.recompute_constraints(float(prop))
    if prop not in ("0", "1"):
        ss.recompute_constraints(float(prop))       # main.rs:113-115 draws again inside the run loop
    obs, start, comp, ncomp = ss.solver_inputs()
    lines = dump.read_text().split("\n")
    n, nc = map(int, lines[0].split())
    assert n == len(obs) and nc == ncomp
    got = np.array([[int(x) for x in ln.split()] for ln in lines[1:n + 1]])
    assert (got[:, 0] == ss.seq).all() and (got[:, 1] == obs).all() and (got[:, 2] == start).all() and (got[:, 3] == comp).all()


@pytest.mark.skipif(not os.path.exists(EXE), reason="host driver not built")
def test_cli_rejects_training_and_bad_args(tmp_path):
    d, *_ = _dataset(tmp_path)
    r = subprocess.run([EXE, "-i", str(d), "-n", "5", "-b", "4", "3", "-p", "1", "-t"], capture_output=True, text=True)
    assert r.returncode != 0 and "outside the GPU hot path" in r.stderr
    r = subprocess.run([EXE, "-i", str(d), "-n", "5", "-p", "1"], capture_output=True, text=True)
    assert r.returncode != 0


def test_stdrng_matches_published_rand_vectors():
    """Pins the RNG stream of `SuperSequence::recompute_constraints` (viterbi_solver/utils.rs:101,168-177:
    `StdRng::seed_from_u64(3019)`, `rng.gen::<f64>() <= prop`) to value-stability vectors that the rand crates
    publish in their own test suites (restated here, the crates are not in /root/reference):

      rand 0.8        rngs/std.rs            test_stdrng_construction   StdRng = ChaCha12, from_seed / from_rng, next_u64
      rand_chacha 0.3 src/chacha.rs          test_chacha_construction   same block function with 20 rounds
      rand_pcg 0.3    tests/lcg64xsh32.rs    test_lcg64xsh32_construction   seed_from_u64(0): the PCG32 seed expansion
      rand 0.8        distributions/float.rs value_stability            Standard f64 = (u64 >> 11) * 2^-53
    """
    from consistent_viterbi_b200.superseq import StdRng
    M64 = (1 << 64) - 1
    # -- StdRng::from_seed(seed).next_u64(), StdRng::from_rng(rng0).next_u64()  (ChaCha12)
    seed = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)
    rng0 = StdRng.from_seed(seed)
    x0 = rng0.next_u64()
    x1 = StdRng.from_rng(rng0).next_u64()
    assert [x0, x1] == [10719222850664546238, 14064965282130556830]

    # -- ChaChaRng (20 rounds): from_seed(..).next_u32() == 137206642, from_rng(..).next_u32() == 1325750369
    class ChaCha20Rng(StdRng):
        ROUNDS = 20
    r1 = ChaCha20Rng.from_seed(bytes([0] * 8 + [1] + [0] * 7 + [2] + [0] * 7 + [3] + [0] * 7))
    r1._block()
    assert r1.buf[0] == 137206642
    r2 = ChaCha20Rng.from_seed(b"".join(w.to_bytes(4, "little") for w in r1.buf[1:9]))
    r2._block()
    assert r2.buf[0] == 1325750369
    # and the RFC 7539 keystream of the all-zero key
    z = ChaCha20Rng.from_seed(bytes(32))
    z._block()
    assert b"".join(w.to_bytes(4, "little") for w in z.buf)[:16].hex() == "76b8e0ada0f13d90405d6ae55386bd28"

    # -- a PCG32 (Lcg64Xsh32) built from the published recipe, to check the two pieces StdRng shares with it
    class Lcg64Xsh32:
        def __init__(self, seed16):
            self.inc = int.from_bytes(seed16[8:16], "little") | 1
            self.state = (int.from_bytes(seed16[:8], "little") + self.inc) & M64
            self.step()

        def step(self):
            self.state = (self.state * 6364136223846793005 + self.inc) & M64

        def next_u32(self):
            st = self.state
            self.step()
            xs, rot = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF, st >> 59
            return ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF

        def next_u64(self):
            lo = self.next_u32()
            return (self.next_u32() << 32) | lo
    assert Lcg64Xsh32(bytes(range(1, 17))).next_u64() == 1204678643940597513          # the helper itself is right
    # seed_from_u64(0): the seed bytes come from StdRng.seed_words_from_u64 (the code recompute_constraints uses)
    seed16 = b"".join(w.to_bytes(4, "little") for w in StdRng.seed_words_from_u64(0, 4))
    assert Lcg64Xsh32(seed16).next_u64() == 18195738587432868099
    # Standard f64 from rand::test::rng(0x6f44f5646c2a7334) = Pcg32::new(seed, 11634580027462260723)
    p = Lcg64Xsh32((0x6f44f5646c2a7334).to_bytes(8, "little") + (11634580027462260723 << 1 & M64).to_bytes(8, "little"))
    assert [StdRng.f64_from_u64(p.next_u64()) for _ in range(3)] == [0.7346051961657583, 0.20298547462974248, 0.8166436635290655]

    # -- the stream the reference consumes, frozen (tests/golden/stdrng_3019.json, written by tools/make_golden.py)
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "stdrng_3019.json")))
    r = StdRng(3019)
    assert [r.next_u64() for _ in range(len(g["next_u64"]))] == g["next_u64"]
    r = StdRng(3019)
    assert [r.gen_f64().hex() for _ in range(len(g["gen_f64_hex"]))] == g["gen_f64_hex"]
    a, b = StdRng(3019), StdRng(3019)
    xs = [a.gen_f64() for _ in range(100)]
    assert xs == [b.gen_f64() for _ in range(100)] and all(0.0 <= x < 1.0 for x in xs) and len(set(xs)) == 100


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(EXE), reason="host driver not built")
def test_cli_end_to_end_gpu(tmp_path):
    d, hmm, seqs, ctl, (A, B, pi) = _dataset(tmp_path, seed=8, nseq=12)
    r = _run(d, tmp_path, "1", {})
    assert r.returncode == 0, r.stderr
    ss = cv.SuperSequence(seqs, cv.Constraints.from_tags(ctl), hmm)
    ss.recompute_constraints(1.0)
    obs, start, comp, ncomp = ss.solver_inputs()
    ref = po.cp_solve(A, B, pi, obs, start, comp, ncomp)
    lines = (tmp_path / "1_0").read_text().split("\n")
    head = lines[0].split()
    assert int(head[1]) == ref["explored"]
    assert float(head[0]) == ref["obj"] or (head[0] == "-inf" and ref["obj"] == -np.inf)
    assert int(lines[1]) >= 0
    body = np.array([[int(x) for x in ln.split()] for ln in lines[2:2 + len(obs)]])
    assert (body[:, 0] == ss.seq).all()                  # main.rs:132 prints MetaElements.seq (index of the sequence)
    assert (body[:, 1] == ref["sol"]).all()
