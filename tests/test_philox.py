"""simdata/philox.py (host statement of the device simulator's noise function): Random123's published known-answer vectors for
Philox4x32-10, and first / second moments of the Box-Muller draws."""
import numpy as np

from simdata import philox


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want


def test_normal_pairs_are_standard_normal_and_keyed():
    z0, z1 = philox.normal_pair(12345, philox.STREAM_VISION, np.arange(4000)[:, None], np.arange(32)[None, :])
    z = np.concatenate([z0.ravel(), z1.ravel()])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and abs(np.corrcoef(z0.ravel(), z1.ravel())[0, 1]) < 0.01
    a, _ = philox.normal_pair(1, philox.STREAM_IMU, 5, 0)
    b, _ = philox.normal_pair(2, philox.STREAM_IMU, 5, 0)
    c, _ = philox.normal_pair(1, philox.STREAM_VISION, 5, 0)
    assert a != b and a != c
    a2, _ = philox.normal_pair(1, philox.STREAM_IMU, 5, 0)
    assert a == a2
