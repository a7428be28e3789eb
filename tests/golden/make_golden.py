"""Generates the committed golden fixtures by running the UNMODIFIED reference (oracle/_ref/libawfm_ref.so).

    python tests/golden/make_golden.py        (only where /root/reference exists; outputs are committed)

Each case = an `.awfmi` file written by the reference's awFmCreateIndex / awFmCreateIndexFromFasta, plus an `.npz`
with the packed queries and the reference's own awFmParallelSearchCount / awFmParallelSearchLocate outputs
(counts, CSR hit offsets, positions in SA order) and, for the FASTA case, the (sequence, local offset) of every
hit from awFmGetLocalSequencePositionFromIndexPosition.  The known-answer FASTA is the 4-record amino example the
reference's test/multiSequenceIndexTest uses (records acdef / g / hikl / m), re-typed here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from avxwindowfmindex_b200 import abi  # noqa: E402
from avxwindowfmindex_b200.search import pack_queries  # noqa: E402
from oracle import harness  # noqa: E402
from conftest import make_queries, make_text  # noqa: E402


def csr(pos_lists):
    off = np.zeros(len(pos_lists) + 1, np.uint64)
    off[1:] = np.cumsum([len(p) for p in pos_lists])
    flat = np.concatenate(pos_lists) if len(pos_lists) else np.zeros(0, np.uint64)
    return off, flat.astype(np.uint64)


def emit(ref, name, index_ptr, queries, text=None):
    letters, offsets = pack_queries(queries)
    counts = ref.count(index_ptr, letters, offsets, threads=2)
    rc, counts2, pos = ref.locate(index_ptr, letters, offsets, threads=2)
    assert rc == abi.AwFmSuccess and np.array_equal(counts, counts2)
    hit_offsets, positions = csr(pos)
    extra = {}
    s = ref.struct(index_ptr)
    if s.fastaVector:
        seq_loc = np.array([ref.contig_of(index_ptr, int(p))[1:] for p in positions], dtype=np.uint64).reshape(-1, 2)
        extra["contig_of_hit"] = seq_loc
    np.savez_compressed(os.path.join(HERE, name + ".npz"), letters=letters, offsets=offsets, counts=counts,
                        hit_offsets=hit_offsets, positions=positions,
                        text=np.zeros(0, np.uint8) if text is None else text, **extra)
    print(f"{name}: {len(queries)} queries, {int(hit_offsets[-1])} hits, "
          f"{os.path.getsize(os.path.join(HERE, name + '.awfmi'))} B index")


def main():
    harness.build(ref=True)
    ref = harness.Reference()
    # 1. nucleotide, ambiguity letters + mixed case in text and queries, seed k=4, SA ratio 4
    text = make_text(4099, False, seed=1, ambiguity_every=300, mixed_case=True)
    ix = ref.create_index(text.tobytes(), os.path.join(HERE, "nuc_k4_r4.awfmi"), abi.AwFmAlphabetDna, 4, 4)
    emit(ref, "nuc_k4_r4", ix, make_queries(text, False, 2, 300, 1, 14, 4), text)
    # 2. amino, seed k=2, SA ratio 3 (not a power of two)
    text = make_text(3001, True, seed=3, ambiguity_every=250)
    ix = ref.create_index(text.tobytes(), os.path.join(HERE, "amino_k2_r3.awfmi"), abi.AwFmAlphabetAmino, 2, 3)
    emit(ref, "amino_k2_r3", ix, make_queries(text, True, 4, 300, 1, 9, 2), text)
    # 3. nucleotide, SA ratio 1 (every position sampled) and a long-backtrace ratio 251
    text = make_text(2500, False, seed=5)
    ix = ref.create_index(text.tobytes(), os.path.join(HERE, "nuc_k3_r1.awfmi"), abi.AwFmAlphabetDna, 3, 1)
    emit(ref, "nuc_k3_r1", ix, make_queries(text, False, 6, 150, 1, 12, 3), text)
    ix = ref.create_index(text.tobytes(), os.path.join(HERE, "nuc_k3_r251.awfmi"), abi.AwFmAlphabetDna, 3, 251)
    emit(ref, "nuc_k3_r251", ix, make_queries(text, False, 7, 150, 1, 12, 3), text)
    # 4. the reference's known-answer multi-sequence case (test/multiSequenceIndexTest/AwFmMultiSequenceTest.c:627-754)
    fasta = os.path.join(HERE, "four_records.fa")
    with open(fasta, "w") as f:
        f.write(">t\nacdef\n>v\ng\n>w\nhikl\n>y\nm\n")
    ix = ref.create_index_from_fasta(fasta, os.path.join(HERE, "four_records.awfmi"), abi.AwFmAlphabetAmino, 2, 2)
    qs = [b"acdef", b"g", b"hikl", b"m", b"fg", b"gh", b"lm", b"t", b"v", b"w", b"y", b"cd", b"ik", b"kl", b"de"]
    emit(ref, "four_records", ix, qs)


if __name__ == "__main__":
    main()
