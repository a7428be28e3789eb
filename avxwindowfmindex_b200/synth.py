"""Deterministic synthetic texts and query sets: splitmix64 streams with fixed seeds.

Element i of a stream with seed s is mix(s + (i+1) * GAMMA); the same formula is used by the CUDA generators in
csrc/awfm_build.cu so host- and device-generated data are identical.
  nucleotide letter = "ACGT"[z >> 62]          amino letter = AMINO[((z >> 32) * 20) >> 32]
"""
import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
NUC = np.frombuffer(b"ACGT", dtype=np.uint8)
AMINO = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)  # index order of src/AwFmLetter.c:59-61
TEXT_SEED = 0xA5F00001
QUERY_SEED = 0xC0FFEE00


def splitmix64(seed, start, count):
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + 1 + count, dtype=np.uint64)
        z = np.uint64(seed) + i * GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def letters(seed, count, amino=False, start=0):
    z = splitmix64(seed, start, count)
    if amino:
        return AMINO[(((z >> np.uint64(32)) * np.uint64(20)) >> np.uint64(32)).astype(np.int64)]
    return NUC[(z >> np.uint64(62)).astype(np.int64)]


def random_text(n, amino=False, seed=TEXT_SEED):
    return letters(seed, n, amino)


def random_queries(num, length, amino=False, seed=QUERY_SEED):
    """`num` iid-uniform queries of `length` letters, packed (fixed length): uint8[num*length]"""
    return letters(seed, num * length, amino)


def sampled_queries(text, num, length, seed=QUERY_SEED):
    """queries cut from the text at pseudo-random offsets (every query has >= 1 hit)"""
    z = splitmix64(seed, 0, num)
    starts = (z % np.uint64(len(text) - length + 1)).astype(np.int64)
    idx = starts[:, None] + np.arange(length)[None, :]
    return text[idx].reshape(-1).copy(), starts
