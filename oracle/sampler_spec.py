"""CPU ORACLE (test infrastructure) for the library's OWN device-RNG sampler.

The reference samples with numpy's global RNG on the host (lib/region.py:43-57); that
stream cannot be reproduced on the device, so the CUDA sampler has its own
counter-based specification (DESIGN.md "Samplers").  This file restates that
specification in numpy so tests can check the kernel bit-for-bit; parity with the
reference itself is checked separately through the rng='numpy' mode and through the
distribution-free properties (counts, subset-of-candidates, positives keep their GT).
"""
import numpy as np

M64 = (1 << 64) - 1


def mix_key(seed, i):
    z = (seed + 0x9E3779B97F4A7C15 * (i + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return (z >> 32) & 0xFFFFFFFF


M32 = (1 << 32) - 1


def mix32(key, v):
    """32-bit keyed round function of the Feistel permutation (murmur3-style finaliser)."""
    h = (v * 0x9E3779B1 + key) & M32
    h ^= h >> 15
    h = (h * 0x85EBCA77) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE3D) & M32
    h ^= h >> 16
    return h


def feistel(x, half_bits, seed):
    mask = (1 << half_bits) - 1
    l, r = x >> half_bits, x & mask
    for rnd in range(4):
        ks = (seed + 0x1000003 * (rnd + 1)) & M64
        f = mix32((ks & M32) ^ (ks >> 32), r) & mask
        l, r = r, l ^ f
    return (l << half_bits) | r


def sample(labels, max_num, pos_num, seed, image_index=0):
    """labels: int64[n] (-1/0/>0).  Returns the ascending chosen index list."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    sd = (seed + 0x632BE59BD9B4E019 * (image_index + 1)) & M64
    pos = np.nonzero(labels > 0)[0]
    if pos.size > pos_num:
        comps = sorted(((mix_key(sd, int(i)) << 32) | int(i)) for i in pos)
        pos = np.array([c & 0xFFFFFFFF for c in comps[:pos_num]], dtype=np.int64)
    keep_pos = pos.size
    nneg = int((labels == 0).sum())
    want = min(max(max_num - keep_pos, 0), nneg)
    neg = []
    if want > 0:
        bits = 2
        while (1 << bits) < n:
            bits += 1
        if bits & 1:
            bits += 1
        half = bits >> 1
        s2 = sd ^ 0xA5A5A5A5DEADBEEF
        for t in range(1 << bits):
            y = feistel(t, half, s2)
            if y < n and labels[y] == 0:
                neg.append(y)
                if len(neg) >= want:
                    break
    return np.sort(np.concatenate([pos.astype(np.int64), np.asarray(neg, dtype=np.int64)]))
