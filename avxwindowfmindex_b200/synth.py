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


def multi_fasta_lengths(num_records, min_len, max_len, seed=TEXT_SEED):
    """record lengths, uniform in [min_len, max_len] (BASELINE cfg 5: 10 000 contigs of 50 k-150 k)"""
    z = splitmix64(seed ^ 0x5EC0, 0, num_records)
    return (np.uint64(min_len) + z % np.uint64(max_len - min_len + 1)).astype(np.int64)


def multi_fasta_text(lengths, amino=False, seed=TEXT_SEED):
    """The text awFmCreateIndexFromFasta indexes (lib/FastaVector/src/FastaVector.c:54-170): records concatenated,
    each followed by one NUL separator; plus the reference's record table (headerEnd, sequenceEnd) for headers
    `contig{i}`.  Returns (text uint8, metadata uint64 (n, 2), header bytes)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    ends = np.cumsum(lengths + 1)
    text = letters(seed, int(ends[-1]), amino)
    text[ends - 1] = 0
    header = b"".join(b"contig%d\0" % i for i in range(len(lengths)))
    header_ends = np.cumsum([len(b"contig%d" % i) + 1 for i in range(len(lengths))])
    meta = np.stack([header_ends.astype(np.uint64), ends.astype(np.uint64)], axis=1)
    return text, meta, header


def write_fasta(path, text, metadata, line=60):
    """FASTA file (60-column lines, headers >contig{i}) that fastaVectorReadFasta turns back into `text`."""
    ends = metadata[:, 1].astype(np.int64)
    with open(path, "wb") as f:
        start = 0
        for i, e in enumerate(ends):
            f.write(b">contig%d\n" % i)
            rec = text[start:e - 1].tobytes()
            for o in range(0, len(rec), line):
                f.write(rec[o:o + line] + b"\n")
            start = e


def sampled_record_queries(text, metadata, num, length, seed=QUERY_SEED):
    """`num` queries of `length` letters cut from inside random records (never across a separator), with the
    by-construction answer: (record index, offset in record, global position)."""
    ends = metadata[:, 1].astype(np.int64)
    starts = np.concatenate([[0], ends[:-1]])
    lens = ends - starts - 1
    ok = np.nonzero(lens >= length)[0]
    z = splitmix64(seed, 0, 2 * num)
    rec = ok[(z[:num] % np.uint64(len(ok))).astype(np.int64)]
    off = (z[num:] % (lens[rec] - length + 1).astype(np.uint64)).astype(np.int64)
    g = starts[rec] + off
    idx = g[:, None] + np.arange(length)[None, :]
    return text[idx].reshape(-1).copy(), rec, off, g
