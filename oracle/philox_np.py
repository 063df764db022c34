"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the counter-based generator behind
`nerf_projects_b200.keyed_rng` (csrc/rng.cu).

The reference draws `t_rand`, `u` and the density noise from torch's global generator
(ipynb:428-444, ipynb:312-325, nerf_helpers.py:404), so a run's draws depend on how the rays are
chunked and sharded.  SURVEY.md 8(e) asks for streams keyed by (seed, global ray index) instead; the
reference holds no algorithm for that, so the generator is Philox4x32-10 (Salmon, Moraes, Dror, Shaw:
"Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123 library), restated here from the paper
and pinned against Random123's published known-answer vectors (`tests/test_philox_cpu.py`).

Layout of one draw (the contract of csrc/rng.cu):
    counter = (ray_lo, ray_hi, sample // 4, stream_id),  key = (seed_lo, seed_hi)
    word j = sample % 4 of the 4x32 output
    uniform  = (word >> 8) * 2^-24                       in [0, 1)
    normal   = Box-Muller on word pairs (0,1) and (2,3): u1 = ((w_a >> 8) + 1) * 2^-24, u2 = (w_b >> 8) * 2^-24,
               r = sqrt(-2 ln u1); (r cos(2 pi u2), r sin(2 pi u2))
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_T_RAND, STREAM_U, STREAM_NOISE_COARSE, STREAM_NOISE_FINE = 0, 1, 2, 3


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 python ints -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def _words(seed, ray_offset, R, S, stream_id):
    """[R, 4*ceil(S/4)] uint32 words of the rays ray_offset .. ray_offset+R-1."""
    G = (S + 3) // 4
    ray = (np.arange(R, dtype=np.uint64) + np.uint64(ray_offset))[:, None]
    g = np.arange(G, dtype=np.uint64)[None, :]
    w = philox4x32_10((ray & MASK, ray >> np.uint64(32), g, np.uint64(stream_id)),
                      (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return np.stack(w, axis=-1).reshape(R, 4 * G)


def keyed_uniform(seed, ray_offset, R, S, stream_id):
    w = _words(seed, ray_offset, R, S, stream_id)[:, :S]
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def keyed_normal(seed, ray_offset, R, S, stream_id, scale=1.0):
    w = _words(seed, ray_offset, R, S, stream_id)
    a, b = w[:, 0::2], w[:, 1::2]
    u1 = ((a >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (b >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u1))
    out = np.empty(w.shape, np.float64)
    out[:, 0::2] = r * np.cos(2 * np.pi * u2)
    out[:, 1::2] = r * np.sin(2 * np.pi * u2)
    return (out[:, :S] * scale).astype(np.float32)
