"""Pins oracle/philox_np.py (the checker of csrc/rng.cu) against the published known-answer vectors of
Philox4x32-10 (Random123 `kat_vectors`, Salmon et al. SC'11), and checks the layout contract."""
import numpy as np

from oracle import philox_np as P


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = tuple(int(x) for x in P.philox4x32_10(ctr, key))
        assert got == want


def test_draws_are_keyed_by_global_ray_index():
    full = P.keyed_uniform(1234, 0, 40, 13, P.STREAM_T_RAND)
    assert full.dtype == np.float32 and full.min() >= 0.0 and full.max() < 1.0
    a = P.keyed_uniform(1234, 0, 17, 13, P.STREAM_T_RAND)
    b = P.keyed_uniform(1234, 17, 23, 13, P.STREAM_T_RAND)
    assert np.array_equal(np.concatenate([a, b]), full)
    assert not np.array_equal(full, P.keyed_uniform(1235, 0, 40, 13, P.STREAM_T_RAND))
    assert not np.array_equal(full, P.keyed_uniform(1234, 0, 40, 13, P.STREAM_U))
    n = P.keyed_normal(7, 1 << 33, 4096, 64, P.STREAM_NOISE_COARSE)
    assert abs(float(n.mean())) < 0.01 and abs(float(n.std()) - 1.0) < 0.01
    assert np.array_equal(P.keyed_normal(7, 1 << 33, 4096, 64, P.STREAM_NOISE_COARSE, scale=2.0), n * np.float32(2.0))
