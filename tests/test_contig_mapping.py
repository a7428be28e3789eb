"""SURVEY.md §8 row f2 — multi-sequence (FASTA) indexes: hits never span records, and global positions map to
(record, offset) exactly as awFmGetLocalSequencePositionFromIndexPosition (src/AwFmSearch.c:284-301,
lib/FastaVector/src/FastaVector.c:338-381) does, including its behaviour at and beyond the last record's end."""
import ctypes as C
import os

import numpy as np
import pytest

from avxwindowfmindex_b200 import abi, capi, read_awfmi, synth
from avxwindowfmindex_b200.search import KmerSearchList, pack_queries, parallel_search_locate
from oracle import harness

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_map(oracle, ends, positions):
    seq = np.zeros(len(positions), np.uint64)
    loc = np.zeros(len(positions), np.uint64)
    bad = 0
    for i, p in enumerate(positions):
        s, l = C.c_uint64(), C.c_uint64()
        if oracle.lib.awfm_oracle_contig_of(ends.ctypes.data, len(ends), int(p), C.byref(s), C.byref(l)) == 0:
            seq[i], loc[i] = s.value, l.value
        else:
            seq[i] = loc[i] = np.uint64(2**64 - 1)
            bad += 1
    return seq, loc, bad


@pytest.fixture(scope="module")
def fasta_index(reference, tmp_path_factory):
    """37 nucleotide records of 1..400 letters, index built by the reference's awFmCreateIndexFromFasta."""
    tmp = str(tmp_path_factory.mktemp("fasta"))
    lengths = synth.multi_fasta_lengths(37, 1, 400, seed=11)
    text, meta, header = synth.multi_fasta_text(lengths, seed=12)
    fasta = os.path.join(tmp, "records.fa")
    synth.write_fasta(fasta, text, meta)
    ptr = reference.create_index_from_fasta(fasta, os.path.join(tmp, "records.awfmi"), abi.AwFmAlphabetDna, 4, 5)
    return {"ptr": ptr, "arrays": reference.arrays(ptr), "text": text, "meta": meta, "header": header,
            "path": os.path.join(tmp, "records.awfmi")}


def test_generator_round_trips_through_the_reference_reader(fasta_index):
    """the synthetic FASTA, read by fastaVectorReadFasta, gives the record table the generator predicted"""
    a = fasta_index["arrays"]
    assert a.bwt_length == len(fasta_index["text"]) + 1
    assert np.array_equal(a.fasta_metadata, fasta_index["meta"])
    assert a.fasta_header == fasta_index["header"]
    on_disk = read_awfmi(fasta_index["path"])
    assert np.array_equal(on_disk.fasta_metadata, fasta_index["meta"]) and on_disk.feature_flags & 1


def test_oracle_mapping_matches_reference_everywhere(fasta_index, reference):
    a = fasta_index["arrays"]
    ends = np.ascontiguousarray(a.fasta_metadata[:, 1])
    oracle = harness.Oracle(a)
    last = int(ends[-1])
    positions = np.arange(0, last + 40, dtype=np.uint64)
    seq, loc, bad = oracle_map(oracle, ends, positions)
    assert bad == 39  # everything strictly beyond the last end
    for p in positions:
        rc, s, l = reference.contig_of(fasta_index["ptr"], int(p))
        if int(p) > last:
            assert rc == abi.AwFmIllegalPositionError and seq[int(p)] == np.uint64(2**64 - 1)
        else:
            assert rc == abi.AwFmSuccess and (s, l) == (int(seq[int(p)]), int(loc[int(p)])), int(p)


def test_oracle_hits_never_span_records(fasta_index, reference):
    """sampled in-record queries are found where they were cut; queries glued across a separator are not found"""
    a, text, meta = fasta_index["arrays"], fasta_index["text"], fasta_index["meta"]
    letters, rec, off, g = synth.sampled_record_queries(text, meta, 200, 9, seed=5)
    oracle = harness.Oracle(a)
    hit, pos, _ = oracle.locate(letters, fixed_len=9)
    ends = np.ascontiguousarray(meta[:, 1])
    for i in range(200):
        mine = pos[int(hit[i]):int(hit[i + 1])]
        assert g[i] in mine
        s, l, bad = oracle_map(oracle, ends, mine)
        assert bad == 0 and (rec[i], off[i]) in set(zip(s.tolist(), l.tolist()))
        assert np.all(l + np.uint64(9) <= (ends[s.astype(np.int64)] - np.where(s > 0, ends[s.astype(np.int64) - 1], 0) - 1))
    rc, counts, r_pos = reference.locate(fasta_index["ptr"], letters, fixed_len=9, threads=2)
    assert rc == abi.AwFmSuccess and np.array_equal(np.concatenate(r_pos), pos)
    # across a separator: last 4 letters of record r + first 5 of record r+1
    e = meta[:, 1].astype(np.int64)
    glued = []
    for r in range(len(e) - 1):
        a0, b0 = text[max(0, e[r] - 5):e[r] - 1], text[e[r]:e[r] + 5]
        if len(a0) == 4 and len(b0) == 5 and 0 not in b0 and 0 not in a0:
            glued.append(bytes(a0) + bytes(b0))
    gl, go = pack_queries(glued)
    counts, _, _ = oracle.count(gl, go)
    brute = [sum(1 for i in range(len(text) - 8) if bytes(text[i:i + 9]) == q) for q in glued]
    assert counts.tolist() == brute


@pytest.mark.gpu
def test_cuda_mapping_matches_oracle(fasta_index):
    from avxwindowfmindex_b200 import GpuIndex
    a = fasta_index["arrays"]
    gpu = GpuIndex(a)  # the record table rides along with the arrays
    ends = np.ascontiguousarray(a.fasta_metadata[:, 1])
    rng = np.random.default_rng(3)
    positions = np.concatenate([np.arange(0, int(ends[-1]) + 40, dtype=np.uint64),
                                rng.integers(0, int(ends[-1]) + 1, 100000).astype(np.uint64),
                                np.array([2**63, 2**64 - 1], dtype=np.uint64)])
    seq, loc, bad = gpu.map_positions(positions)
    o_seq, o_loc, o_bad = oracle_map(harness.Oracle(a), ends, positions)
    assert bad == o_bad and np.array_equal(seq, o_seq) and np.array_equal(loc, o_loc)
    # empty batch, and a context without a record table
    s0, l0, b0 = gpu.map_positions(np.zeros(0, np.uint64))
    assert len(s0) == 0 and b0 == 0
    gpu.close()
    plain = GpuIndex(read_awfmi(os.path.join(GOLDEN, "nuc_k4_r4.awfmi")))
    with pytest.raises(capi.AwfmGpuError):
        plain.map_positions(np.zeros(4, np.uint64))
    plain.close()


@pytest.mark.gpu
@pytest.mark.parametrize("num_records", [1, 255, 256, 257, 5000])
def test_cuda_mapping_table_sizes(num_records):
    """record tables around the 256-entry shared-memory sample, single-letter records included"""
    from avxwindowfmindex_b200 import GpuIndex
    arrays = read_awfmi(os.path.join(GOLDEN, "nuc_k4_r4.awfmi"))
    gpu = GpuIndex(arrays)
    lengths = synth.multi_fasta_lengths(num_records, 0, 7, seed=num_records)  # zero-length records too
    ends = np.cumsum(lengths + 1).astype(np.uint64)
    meta = np.stack([np.zeros_like(ends), ends], axis=1)
    gpu.set_sequences(meta)
    positions = np.arange(0, int(ends[-1]) + 3, dtype=np.uint64)
    seq, loc, bad = gpu.map_positions(positions)
    o_seq, o_loc, o_bad = oracle_map(harness.Oracle(arrays), np.ascontiguousarray(ends), positions)
    assert bad == o_bad == 2 and np.array_equal(seq, o_seq) and np.array_equal(loc, o_loc)
    gpu.close()


@pytest.mark.gpu
def test_drop_in_locate_and_map_on_reference_owned_fasta_index(fasta_index, reference):
    """the reference's own struct AwFmIndex (with its FastaVector) handed to the drop-in: locate, then the batched
    awFmGpuGetLocalSequencePositions against per-hit awFmGetLocalSequencePositionFromIndexPosition"""
    lib = capi.load()
    ip = fasta_index["ptr"]
    letters, rec, off, g = synth.sampled_record_queries(fasta_index["text"], fasta_index["meta"], 500, 7, seed=9)
    sl = KmerSearchList(lib, 500).fill(letters, fixed_len=7)
    assert parallel_search_locate(lib, ip, sl, 2) == abi.AwFmSuccess
    rc, r_counts, r_pos = reference.locate(ip, letters, fixed_len=7, threads=2)
    mine = sl.positions()
    assert all(np.array_equal(p, q) for p, q in zip(r_pos, mine))
    flat = np.concatenate(mine).astype(np.uint64)
    seq = np.zeros(len(flat), np.uint64)
    loc = np.zeros(len(flat), np.uint64)
    assert lib.awFmGpuGetLocalSequencePositions(ip, flat.ctypes.data, len(flat), seq.ctypes.data, loc.ctypes.data) == abi.AwFmSuccess
    for p, s, l in zip(flat, seq, loc):
        assert reference.contig_of(ip, int(p)) == (abi.AwFmSuccess, int(s), int(l))
    # one illegal position flips the return code and marks only that entry
    flat2 = np.append(flat[:5], np.uint64(int(fasta_index["meta"][-1, 1]) + 1))
    seq2, loc2 = np.zeros(6, np.uint64), np.zeros(6, np.uint64)
    assert lib.awFmGpuGetLocalSequencePositions(ip, flat2.ctypes.data, 6, seq2.ctypes.data, loc2.ctypes.data) == abi.AwFmIllegalPositionError
    assert np.array_equal(seq2[:5], seq[:5]) and seq2[5] == np.uint64(2**64 - 1)
    sl.close()
    lib.awFmGpuReleaseIndex(ip)
    # an index without FastaVector: the reference's AwFmUnsupportedVersionError
    plain = read_awfmi(os.path.join(GOLDEN, "nuc_k4_r4.awfmi")).as_awfm_index()
    assert lib.awFmGpuGetLocalSequencePositions(C.addressof(plain), flat.ctypes.data, 1, seq.ctypes.data, loc.ctypes.data) == abi.AwFmUnsupportedVersionError
