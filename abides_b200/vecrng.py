"""n independent numpy-legacy RandomState streams (MT19937) advanced together with numpy array operations.

A seed sweep executes the same config body once per seed (reference: config/parallel.py:9-22 spawns one process per
seed); every config-time draw -- agent seeds, noise wake times, latency coordinates, the oracle's first megashock --
comes from that instance's own `np.random.seed(seed)` stream.  `VecRandomState` holds all n generators as one
[n, 624] array and reproduces the legacy transforms word for word, so n instances are built in the time
of a few thousand array operations instead of n Python executions of the config.

Only the draws the supported config families make are implemented: random_sample / rand, uniform, randint (32-bit
masked rejection), standard gauss / normal (polar method with numpy's cached second value), exponential.
"""
import math

import numpy as np

N, M = 624, 397
_U, _L, _MAG = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)


def init_genrand(seeds):
    """mt19937_seed (numpy legacy seeding of an integer seed): [n] uint32 seeds -> [n, 624] states."""
    seeds = np.asarray(seeds, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    key = np.empty((len(seeds), N), dtype=np.uint32)
    x = seeds.copy()
    key[:, 0] = x.astype(np.uint32)
    m = np.uint64(1812433253)
    for i in range(1, N):
        x = (m * (x ^ (x >> np.uint64(30))) + np.uint64(i)) & np.uint64(0xFFFFFFFF)
        key[:, i] = x.astype(np.uint32)
    return key


def _mix(a, b):
    y = (a & _U) | (b & _L)
    return (y >> np.uint32(1)) ^ ((y & np.uint32(1)) * _MAG)


def twist(key):
    """One generation step of every row of key [m, 624], in place.  Word k needs the NEW word k - 227 for k >= 227, so
    the block is advanced in chunks of 227."""
    key[:, 0:227] = key[:, 397:624] ^ _mix(key[:, 0:227], key[:, 1:228])
    key[:, 227:454] = key[:, 0:227] ^ _mix(key[:, 227:454], key[:, 228:455])
    key[:, 454:623] = key[:, 227:396] ^ _mix(key[:, 454:623], key[:, 455:624])
    key[:, 623] = key[:, 396] ^ _mix(key[:, 623], key[:, 0])


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


class VecRandomState:
    """`n` RandomState(seed=seeds[i]) generators.  Every draw method returns one value per stream; `rows` restricts a
    draw to a subset (used by the rejection loops, where streams consume different numbers of words)."""

    def __init__(self, seeds):
        self.seeds = np.asarray(seeds, dtype=np.uint64).astype(np.uint32)
        self.n = len(self.seeds)
        self.key = init_genrand(self.seeds)
        self.pos = np.full(self.n, N, dtype=np.int64)
        self.consumed = np.zeros(self.n, dtype=np.int64)       # 32-bit words drawn since seeding
        self.has_gauss = np.zeros(self.n, dtype=np.int32)
        self.gauss = np.zeros(self.n, dtype=np.float64)
        self._all = np.arange(self.n)
        self._lockstep = N                                     # common position of all streams, -1 once they diverge

    # ---- words ----
    def u32(self, rows=None):
        if rows is None and self._lockstep >= 0:               # every stream at the same position: plain column reads
            p = self._lockstep
            if p == N:
                twist(self.key)
                p = 0
            y = self.key[:, p]
            self._lockstep = p + 1
            self.pos[:] = p + 1
            self.consumed += 1
            return temper(y)
        self._lockstep = -1
        rows = self._all if rows is None else rows
        p = self.pos[rows]
        need = p == N
        if need.any():
            r = rows[need]
            k = self.key[r]
            twist(k)
            self.key[r] = k
            self.pos[r] = 0
            p = self.pos[rows]
        y = self.key[rows, p]
        self.pos[rows] = p + 1
        self.consumed[rows] += 1
        return temper(y)

    def skip_words(self, n_words):
        """Advance every stream by `n_words` words without producing them (e.g. the N x N latency draw of which the
        sparse_zi configs only read one row): bookkeeping only -- the device fast-forwards seeded streams itself."""
        self.consumed += int(n_words)
        self._lockstep = -1
        self.pos[:] = -1                                       # the host copy is no longer usable

    # ---- legacy transforms ----
    def random_sample(self, rows=None):
        a = (self.u32(rows) >> np.uint32(5)).astype(np.float64)
        b = (self.u32(rows) >> np.uint32(6)).astype(np.float64)
        return (a * 67108864.0 + b) / 9007199254740992.0

    rand = random_sample

    def uniform(self, low, high, rows=None):
        return low + (high - low) * self.random_sample(rows)

    def randint(self, low, high):
        """randint(low, high) with one (scalar) range for all streams: masked rejection over 32-bit words."""
        rng = int(high) - int(low) - 1
        if rng < 0:
            raise ValueError("low >= high")
        if rng == 0:
            return np.full(self.n, int(low), dtype=np.int64)
        if rng > 0xFFFFFFFF:
            raise NotImplementedError("randint range beyond 32 bits")
        if rng == 0xFFFFFFFF:
            return int(low) + self.u32().astype(np.int64)
        mask = rng
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        out = np.empty(self.n, dtype=np.int64)
        pending = self._all
        while len(pending):
            v = self.u32(pending) & np.uint32(mask)
            ok = v <= rng
            out[pending[ok]] = v[ok].astype(np.int64)
            pending = pending[~ok]
        return int(low) + out

    def standard_gauss(self):
        """legacy_gauss: polar Box-Muller; the second value of a pair is cached per stream."""
        out = np.empty(self.n, dtype=np.float64)
        cached = self.has_gauss.astype(bool)
        out[cached] = self.gauss[cached]
        self.gauss[cached] = 0.0
        self.has_gauss[cached] = 0
        pending = self._all[~cached]
        while len(pending):
            x1 = 2.0 * self.random_sample(pending) - 1.0
            x2 = 2.0 * self.random_sample(pending) - 1.0
            r2 = x1 * x1 + x2 * x2
            ok = ~((r2 >= 1.0) | (r2 == 0.0))
            rows = pending[ok]
            # scalar libm log / sqrt, as numpy's C code calls them (array ufuncs may use SIMD variants)
            f = np.array([math.sqrt(-2.0 * math.log(v) / v) for v in r2[ok].tolist()], dtype=np.float64)
            self.gauss[rows] = f * x1[ok]
            self.has_gauss[rows] = 1
            out[rows] = f * x2[ok]
            pending = pending[~ok]
        return out

    def normal(self, loc, scale):
        return loc + scale * self.standard_gauss()

    def exponential(self, scale):
        u = self.random_sample()
        return scale * np.array([-math.log(1.0 - v) for v in u.tolist()], dtype=np.float64)

    # ---- comparison with numpy (tests) ----
    def numpy_state(self, i):
        """The legacy state tuple of stream i (only valid while the host copy is usable)."""
        if self.pos[i] < 0:
            raise ValueError("stream was fast-forwarded")
        return ("MT19937", self.key[i].copy(), int(self.pos[i]), int(self.has_gauss[i]), float(self.gauss[i]))
