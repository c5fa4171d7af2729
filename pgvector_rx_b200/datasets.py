"""Seeded synthetic workloads.  numpy's PCG64 streams are reproducible on every
box that runs this image, so the CPU oracle and the GPU path always see identical inputs."""
import math

import numpy as np


def gaussian_f32(n, dim, seed):
    """n x dim fp32 N(0,1) rows (configs C1/C2/C4)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    step = max(1, (1 << 24) // dim)
    for s in range(0, n, step):
        out[s:s + step] = rng.standard_normal((min(step, n - s), dim), dtype=np.float32)
    return out


def random_bits(n, nbits, seed):
    """n x ceil(nbits/8) packed iid Bernoulli(0.5) rows (config C5)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, (nbits + 7) // 8), dtype=np.uint8)


def clustered_bits(n, nbits, seed, centres=1000, flip=0.10):
    """C5 clustered variant: random centres with `flip` of the bits flipped per row."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 2, size=(centres, nbits), dtype=np.uint8)
    which = rng.integers(0, centres, size=n)
    rows = c[which] ^ (rng.random((n, nbits)) < flip).astype(np.uint8)
    return np.packbits(rows, axis=1)


def hnsw_levels(n, m, seed):
    """Element levels per the reference's formula floor(-ln(max(r, MIN_POSITIVE)) / ln m), capped
    (src/index/build.rs:373-377, src/types/hnsw.rs:337-349), with r from a seeded stream instead
    of the reference's unseeded rand::random."""
    rng = np.random.default_rng(seed)
    r = np.maximum(rng.random(n), np.finfo(np.float64).tiny)
    ml = 1.0 / math.log(m)
    cap = min((8192 - 24 - 8 - 4 - 4) // 6 // m - 2, 255)
    return np.minimum(np.floor(-np.log(r) * ml), cap).astype(np.int32)
