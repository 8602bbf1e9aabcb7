"""ORACLE tooling (test infrastructure): seeded synthetic average hashes for the 'hash' dedup branch."""
import numpy as np


def make_hashes(n, seed):
    """Synthetic 64-bit average hashes (dupes.py:11-15 yields `1 * diff.flatten()`, an int64 0/1 vector): random
    rows, every fourth row a copy of an earlier one with 0..15 flipped bits so distances straddle hash_thr = 8."""
    rng = np.random.default_rng(seed)
    H = rng.integers(0, 2, size=(n, 64)).astype(np.int64)
    for i in range(4, n, 4):
        j = int(rng.integers(0, i))
        H[i] = H[j]
        flip = rng.choice(64, size=int(rng.integers(0, 16)), replace=False)
        H[i, flip] ^= 1
    return H
