"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference mask generators.

Restates, for checking the CUDA mask kernel bit-for-bit:
  * compute_mask_indices, live path only        (reference wavjepa/audio_masking.py:46-47,57-70,82-93,141-145,151-181,194)
  * TimeInverseBlockMasker.forward              (reference wavjepa/masking.py:66-128)
  * SpeechMasker.filter_small_clusters/forward  (reference wavjepa/masking.py:150-207)
and the third-party RNG chain the reference gets from numpy (pinned numpy 2.2.6, requirements.txt:38):
  SeedSequence -> PCG64 (XSL-RR 128/64) -> Generator.random / Generator.choice(replace=False) (Floyd + Lemire).

Two restatements are provided: `*_np` drives numpy's own Generator (numpy IS the reference's dependency), and the
`Pcg64`/`span_mask_py` pure-Python chain restates numpy's published algorithm integer by integer; the tests pin
both against the executed reference (tests/golden/masks_*.npz) and against each other.

Seed contract (the reference defines none, production masks use OS entropy -- SURVEY.md 8a-M1): call number
`call_idx` of attempt `attempt` for global row `row` uses default_rng([base_seed, row, attempt*8 + call_idx]),
call_idx 0 = context mask, 1..4 = target groups.

Pinning status: pinned against the reference executed in the build container (see tests/golden/make_golden.py).
Nothing in the product package imports this file.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
M128 = (1 << 128) - 1


# ------------------------------------------------------------------------------------------- numpy SeedSequence
def seed_sequence_state(words):
    """numpy.random.SeedSequence(entropy=words).generate_state(4, uint64), for <= 4 uint32 entropy words."""
    INIT_A, MULT_A = 0x43B0D7E5, 0x931E8875
    INIT_B, MULT_B = 0x8B51F9DD, 0x58F38DED
    MIX_L, MIX_R = 0xCA01F9DD, 0x4973F715
    hc = [INIT_A]

    def hashmix(v):
        v = (v ^ hc[0]) & M32
        hc[0] = (hc[0] * MULT_A) & M32
        v = (v * hc[0]) & M32
        v ^= v >> 16
        return v

    def mix(x, y):
        r = ((MIX_L * x) & M32) - ((MIX_R * y) & M32)
        r &= M32
        r ^= r >> 16
        return r

    assert len(words) <= 4
    pool = [hashmix(words[i] if i < len(words) else 0) for i in range(4)]
    for src in range(4):
        for dst in range(4):
            if src != dst:
                pool[dst] = mix(pool[dst], hashmix(pool[src]))
    hc2 = INIT_B
    out32 = []
    for i in range(8):
        d = pool[i % 4] ^ hc2
        hc2 = (hc2 * MULT_B) & M32
        d = (d * hc2) & M32
        d ^= d >> 16
        out32.append(d)
    return [out32[2 * i] | (out32[2 * i + 1] << 32) for i in range(4)]


class Pcg64:
    """numpy's PCG64 bit generator (pcg64.h: pcg_setseq_128_xsl_rr_64) seeded like default_rng(words)."""

    MULT = 0x2360ED051FC65DA44385DF649FCCF645

    def __init__(self, words):
        s = seed_sequence_state(list(words))
        initstate = (s[0] << 64) | s[1]
        initseq = (s[2] << 64) | s[3]
        self.inc = ((initseq << 1) | 1) & M128
        self.state = 0
        self._step()
        self.state = (self.state + initstate) & M128
        self._step()
        self.has32 = False
        self.buf32 = 0

    def _step(self):
        self.state = (self.state * self.MULT + self.inc) & M128

    def next64(self):
        self._step()
        hi, lo = self.state >> 64, self.state & M64
        x = hi ^ lo
        rot = self.state >> 122
        return ((x >> rot) | (x << ((-rot) & 63))) & M64

    def next32(self):
        if self.has32:
            self.has32 = False
            return self.buf32
        n = self.next64()
        self.has32 = True
        self.buf32 = n >> 32
        return n & M32

    def random(self):
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    def bounded(self, r):
        """random_bounded_uint64(off=0, rng=r) for r < 2**32-1 : Lemire, 32-bit."""
        if r == 0:
            return 0
        excl = r + 1
        m = self.next32() * excl
        left = m & M32
        if left < excl:
            thr = (M32 - r) % excl
            while left < thr:
                m = self.next32() * excl
                left = m & M32
        return m >> 32

    def choice_no_replace(self, n, k):
        """Generator.choice(n, k, replace=False) (Floyd branch; trailing shuffle omitted: it does not change the set)."""
        assert not (n > 10000 and k > n // 50)
        chosen = set()
        out = []
        for j in range(n - k, n):
            v = self.bounded(j)
            if v in chosen:
                v = j
            chosen.add(v)
            out.append(v)
        return out


def num_spans(p: float, sz: int, length: int, u: float) -> int:
    # audio_masking.py:82-87: int(mask_prob * sz / float(mask_length) + rng.random()), IEEE double, no FMA
    return int(p * sz / float(length) + u)


def _paint(sz, starts, length):
    m = np.zeros(sz, dtype=bool)
    for s in starts:
        m[s:min(s + length, sz)] = True
    return m


def span_mask_py(words, sz: int, p: float, length: int) -> np.ndarray:
    """compute_mask_indices(shape=(1,sz), None, p, length) with the pure-Python RNG chain."""
    rng = Pcg64(words)
    num = num_spans(p, sz, length, rng.random())
    if num == 0:
        raise ValueError("this should never happens")  # audio_masking.py:101-103
    min_len = length
    if sz - min_len <= num:
        min_len = sz - num - 1
    starts = rng.choice_no_replace(sz - min_len, num)
    return _paint(sz, starts, length)


def span_mask_np(words, sz: int, p: float, length: int) -> np.ndarray:
    """Same, driving numpy's own Generator."""
    rng = np.random.default_rng(list(words))
    num = num_spans(p, sz, length, rng.random())
    if num == 0:
        raise ValueError("this should never happens")
    min_len = length
    if sz - min_len <= num:
        min_len = sz - num - 1
    starts = rng.choice(sz - min_len, num, replace=False)
    return _paint(sz, starts, length)


def _expand_channels(ctx_hidden, tgt, vis, C):
    # masking.py:120-126: repeat over channels and flatten "(S C)" (time-major interleave)
    if C > 1:
        ctx_hidden = np.repeat(ctx_hidden, C, axis=-1)
        tgt = np.repeat(tgt, C, axis=-1)
        vis = np.repeat(vis, C, axis=-1)
    return ctx_hidden, tgt, vis


def time_inverse_masks(base_seed: int, row0: int, batch: int, n_times: int, in_channels: int = 1, n_targets: int = 4,
                       ctx_prob: float = 0.65, ctx_len: int = 10, tgt_prob: float = 0.25, tgt_len: int = 10,
                       cutoff: float = 0.1, channel_based: bool = False, span=span_mask_np):
    """TimeInverseBlockMasker.forward (masking.py:66-128) under the seed contract. Returns three bool arrays
    (ctx hidden [B,T], targets [B,G,T], ctx-and-target hidden [B,G,T]) plus attempts per row."""
    T = n_times // in_channels
    ctx_hidden = np.zeros((batch, T), dtype=bool)
    tgt = np.zeros((batch, n_targets, T), dtype=bool)
    attempts = np.zeros(batch, dtype=np.int32)
    for i in range(batch):
        row = row0 + i
        attempt = 0
        while True:
            ctx = ~span([base_seed, row, attempt * 8 + 0], T, ctx_prob, ctx_len)
            tg = np.stack([span([base_seed, row, attempt * 8 + 1 + g], T, tgt_prob, tgt_len) for g in range(n_targets)])
            ctx = ctx & ~tg.any(axis=0)
            ratio = np.float32(ctx.sum()) / np.float32(T)  # torch: int64 sum / int -> float32
            attempt += 1
            if ratio >= np.float32(cutoff):
                break
        ctx_hidden[i] = ~ctx
        tgt[i] = tg
        attempts[i] = attempt
    vis = np.logical_xor(ctx_hidden[:, None, :], tgt)
    if channel_based:
        ctx_hidden, tgt, vis = _expand_channels(ctx_hidden, tgt, vis, in_channels)
    return ctx_hidden, tgt, vis, attempts


def filter_small_clusters(mask: np.ndarray, min_len: int) -> np.ndarray:
    """SpeechMasker.filter_small_clusters (masking.py:150-165): True-runs shorter than min_len become False."""
    out = mask.copy()
    n = len(mask)
    i = 0
    while i < n:
        j = i
        while j < n and mask[j] == mask[i]:
            j += 1
        if mask[i] and (j - i) < min_len:
            out[i:j] = False
        i = j
    return out


def speech_masks(base_seed: int, row0: int, batch: int, n_times: int, in_channels: int = 1, n_targets: int = 4,
                 tgt_prob: float = 0.25, tgt_len: int = 5, cutoff: float = 0.3, min_context_len: int = 5,
                 channel_based: bool = False, span=span_mask_np):
    """SpeechMasker.forward (masking.py:167-207) under the seed contract (call_idx 1..4)."""
    T = n_times // in_channels
    ctx_hidden = np.zeros((batch, T), dtype=bool)
    tgt = np.zeros((batch, n_targets, T), dtype=bool)
    attempts = np.zeros(batch, dtype=np.int32)
    for i in range(batch):
        row = row0 + i
        attempt = 0
        while True:
            tg = np.stack([span([base_seed, row, attempt * 8 + 1 + g], T, tgt_prob, tgt_len) for g in range(n_targets)])
            ctx = filter_small_clusters(~tg.any(axis=0), min_context_len)
            ratio = np.float32(ctx.sum()) / np.float32(T)
            attempt += 1
            if ratio >= np.float32(cutoff):
                break
        ctx_hidden[i] = ~ctx
        tgt[i] = tg
        attempts[i] = attempt
    vis = np.logical_xor(ctx_hidden[:, None, :], tgt)
    if channel_based:
        ctx_hidden, tgt, vis = _expand_channels(ctx_hidden, tgt, vis, in_channels)
    return ctx_hidden, tgt, vis, attempts
