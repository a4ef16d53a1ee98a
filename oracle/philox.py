"""TEST INFRASTRUCTURE: numpy restatement of the dropout RNG contract of include/mtn_b200.h (MtnLinearArgs).

Philox-4x32-10 (Salmon et al., SC'11; the published constants), counter = (e >> 3 low word, e >> 3 high word, site,
0x6d746e62), key = (seed low word, seed high word); the four output words give eight 16-bit values, element
8*(e >> 3) + 2j is the low half of word j, + 2j + 1 its high half; an element is KEPT iff its value >= thresh.
The reference uses torch's nn.Dropout, whose random stream cannot be reproduced by any other implementation; what
can be pinned is the distribution (keep probability, scaling by 1/(1-p)) and that forward and backward agree."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + np.uint64(W0)) & MASK, (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def keep_mask(seed, site, thresh, n):
    """bool[n]: keep decision of elements 0..n-1 of dropout site `site` under `seed` (n padded up to 8 internally)."""
    g = (n + 7) // 8
    idx8 = np.arange(g, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(idx8 & MASK, idx8 >> np.uint64(32), np.full(g, site, np.uint64), np.full(g, 0x6d746e62, np.uint64),
                      seed & 0xFFFFFFFF, seed >> 32)
    vals = np.empty((g, 8), dtype=np.uint64)
    for j in range(4):
        vals[:, 2 * j] = r[j] & np.uint64(0xFFFF)
        vals[:, 2 * j + 1] = r[j] >> np.uint64(16)
    return (vals >= np.uint64(thresh)).reshape(-1)[:n]
