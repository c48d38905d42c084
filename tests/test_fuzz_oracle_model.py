"""CPU fuzz: frames with a VALID header followed by RANDOM payload bits, through the C oracle and
the independent Python model.  Random bits reach states an encoder never produces (escapes inside
zero runs, histories at the clamp, zero-run k = 16 from the clz(0) quirk, runs past the frame end).
Wherever the oracle reports OK the two restatements must agree byte for byte; wherever the model
refuses (`Unmodelled`) the oracle must report a non-OK status."""
import numpy as np
import pytest

from pymodel import alac_model as M


class BW:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, val, bits):
        self.v = (self.v << bits) | (val & ((1 << bits) - 1))
        self.n += bits

    def bytes(self):
        pad = (-self.n) % 8
        return ((self.v << pad).to_bytes((self.n + pad) // 8, "big")) if self.n else b""


def random_frame(rng, sample_size, max_n):
    stereo = bool(rng.integers(0, 2))
    n = int(rng.integers(1, max_n + 1))
    hassize = n != max_n or bool(rng.integers(0, 2))
    ub = int(rng.choice([0, 0, 0, 1, 2])) if sample_size == 24 else int(rng.choice([0, 0, 0, 1]))
    escape = int(rng.random() < 0.1)
    w = BW()
    w.put(1 if stereo else 0, 3); w.put(0, 16); w.put(int(hassize), 1); w.put(ub, 2); w.put(escape, 1)
    if hassize:
        w.put(n, 32)
    if not escape:
        w.put(int(rng.integers(0, 6)), 8)            # mix shift
        w.put(int(rng.integers(0, 40)), 8)           # mix weight (may exceed 2^shift: decoder just wraps)
        for _ in range(2 if stereo else 1):
            order = int(rng.choice([0, 1, 2, 4, 8, 15, 30, 31]))
            w.put(0, 4); w.put(int(rng.integers(0, 16)), 4); w.put(int(rng.integers(0, 8)), 3); w.put(order, 5)
            for _ in range(order):
                w.put(int(rng.integers(-2000, 2000)) & 0xFFFF, 16)
    # payload: random bits with a bias towards zeros so Rice prefixes stay short and zero runs happen
    nbits = int(rng.integers(200, 6000))
    p = float(rng.choice([0.15, 0.3, 0.5]))
    bits = (rng.random(nbits) < p).astype(np.uint8)
    for b in bits.tolist():
        w.put(b, 1)
    return w.bytes(), n


@pytest.mark.parametrize("seed", range(6))
def test_random_payloads_agree(seed, oracle):
    rng = np.random.default_rng(9000 + seed)
    agree = refused = 0
    for _ in range(120):
        ss = int(rng.choice([16, 24]))
        cch = int(rng.choice([1, 2]))
        max_n = int(rng.choice([16, 64, 200]))
        kw = dict(sample_size=ss, num_channels=cch, max_samples_per_frame=max_n,
                  rice_history_mult=int(rng.choice([40, 40, 4, 255])), rice_initial_history=int(rng.choice([10, 10, 0, 255])),
                  rice_kmodifier=int(rng.choice([14, 14, 6, 20])))
        frame, n = random_frame(rng, ss, max_n)
        pcm, st = oracle.read_frame(oracle.make_cfg(**kw), frame)
        ck = M.Cookie(ss, cch, max_n, kw["rice_history_mult"], kw["rice_initial_history"], kw["rice_kmodifier"])
        try:
            model = M.read_frame(ck, frame)
        except M.Unmodelled:
            assert st != 0, "the model refuses a frame the oracle calls OK"
            refused += 1
            continue
        if st == 0:
            assert model == pcm
            agree += 1
    assert agree >= 20, (agree, refused)       # the fuzz must actually reach OK frames
