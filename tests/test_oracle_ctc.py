"""The oracle's CTC decoders against brute force.  rten-text (the reference's decoder crate,
call sites recognition.rs:509-515) is not vendored, so the definition is the check: with a beam
wide enough to keep every prefix, prefix beam search must return the labelling with the largest
total probability over all alignments."""
import itertools
import math

import numpy as np
import pytest

from oracle.recognition import ctc_decode_beam, ctc_decode_greedy


def _collapse(path):
    out, last = [], 0
    for c in path:
        if c != last and c != 0:
            out.append(c)
        last = c
    return tuple(out)


def _brute_force(logp):
    T, C = logp.shape
    mass = {}
    for path in itertools.product(range(C), repeat=T):
        lab = _collapse(path)
        mass[lab] = mass.get(lab, 0.0) + math.exp(sum(logp[t, c] for t, c in enumerate(path)))
    return mass


@pytest.mark.parametrize("seed", range(8))
def test_beam_equals_exhaustive_search(seed):
    rng = np.random.default_rng(seed)
    T, C = 5, 4
    x = rng.normal(0, 1.5, (T, C))
    logp = (x - np.log(np.exp(x).sum(-1, keepdims=True))).astype(np.float32)
    mass = _brute_force(logp.astype(np.float64))
    best = max(mass, key=mass.get)
    steps, score = ctc_decode_beam(logp, width=10_000)
    assert tuple(s.label for s in steps) == best
    assert math.isclose(math.exp(score), mass[best], rel_tol=1e-9)
    # positions are increasing timesteps at which the label can be emitted
    pos = [s.pos for s in steps]
    assert pos == sorted(set(pos)) and all(0 <= p < T for p in pos)


def test_beam_width_one_on_peaked_input_equals_greedy():
    rng = np.random.default_rng(3)
    T, C = 30, 8
    labels = rng.integers(0, C, T)
    x = rng.uniform(0, 0.1, (T, C))
    x[np.arange(T), labels] = 12.0
    logp = (x - np.log(np.exp(x).sum(-1, keepdims=True))).astype(np.float32)
    g, _ = ctc_decode_greedy(logp)
    b, _ = ctc_decode_beam(logp, width=1)
    assert [(s.label, s.pos) for s in g] == [(s.label, s.pos) for s in b]
