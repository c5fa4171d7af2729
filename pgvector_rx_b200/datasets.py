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


def clustered_bits(n, nbits, seed, centres=1000, flip=0.10, centre_seed=9):
    """C5 clustered variant: `centres` random centres, every row is one of them with `flip` of its bits
    flipped.  The centres come from `centre_seed`, so base rows and queries drawn with different `seed`s share them
    (iid bits make every distance ~ nbits/2 and recall meaningless).  Generated in row chunks: 5M x 1024 bits never
    needs more than ~0.3 GB of temporaries."""
    c = np.random.default_rng(centre_seed).integers(0, 2, size=(centres, nbits), dtype=np.uint8)
    rng = np.random.default_rng(seed)
    which = rng.integers(0, centres, size=n)
    out = np.empty((n, (nbits + 7) // 8), np.uint8)
    step = max(1, (1 << 26) // nbits)
    for s in range(0, n, step):
        e = min(n, s + step)
        flips = rng.random((e - s, nbits), dtype=np.float32) < np.float32(flip)
        out[s:e] = np.packbits(c[which[s:e]] ^ flips.astype(np.uint8), axis=1)
    return out


def low_dim_f32(n, dim, seed, latent=16, noise=0.05, basis_seed=10):
    """n x dim fp32 rows of low intrinsic dimension: a `latent`-dimensional N(0,1) point mapped through a fixed random
    dim-dimensional basis (from `basis_seed`, shared by base rows and queries) plus `noise` * N(0,1) per coordinate.
    iid N(0,1) data in 128 dimensions has no neighbourhood structure (recall@10 ~ 0.17 at ef_search = 40); on data
    like this an ANN index reaches the recall it is deployed at."""
    basis = np.random.default_rng(basis_seed).standard_normal((latent, dim)).astype(np.float32) / np.float32(math.sqrt(latent))
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    step = max(1, (1 << 22) // dim)
    for s in range(0, n, step):
        e = min(n, s + step)
        z = rng.standard_normal((e - s, latent), dtype=np.float32)
        out[s:e] = z @ basis + np.float32(noise) * rng.standard_normal((e - s, dim), dtype=np.float32)
    return out


def hnsw_levels(n, m, seed):
    """Element levels per the reference's formula floor(-ln(max(r, MIN_POSITIVE)) / ln m), capped
    (src/index/build.rs:373-377, src/types/hnsw.rs:337-349), with r from a seeded stream instead
    of the reference's unseeded rand::random."""
    rng = np.random.default_rng(seed)
    r = np.maximum(rng.random(n), np.finfo(np.float64).tiny)
    ml = 1.0 / math.log(m)
    cap = min((8192 - 24 - 8 - 4 - 4) // 6 // m - 2, 255)
    return np.minimum(np.floor(-np.log(r) * ml), cap).astype(np.int32)
