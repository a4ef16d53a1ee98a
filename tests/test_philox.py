"""CPU tier: the numpy restatement of the dropout RNG contract (oracle/philox.py) against the published
Random123 known-answer vectors of Philox-4x32-10, and the keep probability of the 16-bit threshold rule."""
import numpy as np

import philox


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        r = philox.philox4x32_10(*[[c] for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in r) == out


def test_keep_probability_and_site_independence():
    th = int(round(0.1 * 65536))
    a = philox.keep_mask(7, 1, th, 400000)
    b = philox.keep_mask(7, 2, th, 400000)
    c = philox.keep_mask(8, 1, th, 400000)
    for m in (a, b, c):
        assert abs(m.mean() - (1 - th / 65536.0)) < 3e-3
    assert 0.78 < (a == b).mean() < 0.86 and 0.78 < (a == c).mean() < 0.86      # independent: 0.9^2 + 0.1^2 = 0.82
    assert np.array_equal(a, philox.keep_mask(7, 1, th, 400000))
