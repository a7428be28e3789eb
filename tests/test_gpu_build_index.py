"""Device index construction (row f4) must be byte-identical to the reference's awFmCreateIndex on the same text:
blocks, prefix sums, seed table and the bit-packed sampled SA including its 8 trailing bytes."""
import numpy as np
import pytest

from avxwindowfmindex_b200 import DeviceBuiltIndex, abi, synth, write_awfmi
from conftest import make_text

pytestmark = pytest.mark.gpu


def assert_same(built, ref_arrays, tag):
    mine = built.to_host()
    assert mine.bwt_length == ref_arrays.bwt_length
    assert np.array_equal(mine.prefix_sums, ref_arrays.prefix_sums), tag
    assert np.array_equal(mine.blocks, ref_arrays.blocks), tag
    assert np.array_equal(mine.seed_table, ref_arrays.seed_table), tag
    assert np.array_equal(mine.sa_bytes, ref_arrays.sa_bytes), tag
    return mine


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r1", "nuc_r3", "nuc_r16", "nuc_r200", "nuc_r255", "amino_r8",
                                  "amino_r2", "amino_r1"])
def test_matches_reference_built_indexes(small_indexes, name):
    b = small_indexes[name]
    built = DeviceBuiltIndex.from_host_text(b.text, b.arrays.alphabet, b.arrays.seed_k, b.arrays.sa_ratio)
    assert_same(built, b.arrays, name)
    built.close()


@pytest.mark.parametrize("n", [1, 2, 5, 255, 256, 257, 511, 513, 1000])
def test_tiny_texts(reference, tmp_path, n):
    text = make_text(n, False, seed=n)
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "t.awfmi"), abi.AwFmAlphabetDna, 2, 3)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetDna, 2, 3)
    assert_same(built, reference.arrays(ptr), n)
    built.close()
    reference.dealloc_index(ptr)


def test_long_repeats_are_finished_by_prefix_doubling(reference, tmp_path):
    """What real genomes hold and iid text does not: a tandem array (satellite DNA), a segment duplicated far away, a
    homopolymer run and a 2-periodic run — suffixes that agree for up to 600 k symbols.  The device finishes them by
    prefix doubling (log2 of the longest repeat rounds over the tied suffixes only); the result is byte-identical to
    what the reference's libdivsufsort-based awFmCreateIndex builds."""
    rng = np.random.default_rng(4)
    unit = make_text(171, False, seed=9)
    segment = make_text(300_000, False, seed=10)
    text = np.concatenate([make_text(50_000, False, seed=11), np.tile(unit, 3500), make_text(70_000, False, seed=12),
                           segment, np.full(40_000, ord("A"), np.uint8), make_text(20_000, False, seed=13), segment,
                           np.frombuffer(b"AC" * 30_000, np.uint8), np.tile(unit, 200), make_text(10_000, False, seed=14)]).copy()
    text[rng.integers(0, len(text), 40)] = ord("N")
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "long.awfmi"), abi.AwFmAlphabetDna, 6, 4)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetDna, 6, 4)
    assert built.tie_suffixes > 900_000 and built.tie_rounds >= 12, (built.tie_suffixes, built.tie_rounds)
    assert_same(built, reference.arrays(ptr), "long repeats")
    built.close()
    reference.dealloc_index(ptr)
    # amino: a repeated domain and a low-complexity run
    domain = make_text(400, True, seed=15)
    atext = np.concatenate([make_text(20_000, True, seed=16), np.tile(domain, 150), np.full(5_000, ord("Q"), np.uint8),
                            make_text(9_000, True, seed=17), np.tile(domain, 40)]).copy()
    ptr = reference.create_index(atext.tobytes(), str(tmp_path / "along.awfmi"), abi.AwFmAlphabetAmino, 3, 3)
    built = DeviceBuiltIndex.from_host_text(atext, abi.AwFmAlphabetAmino, 3, 3)
    assert built.tie_suffixes > 60_000 and built.tie_rounds >= 8, (built.tie_suffixes, built.tie_rounds)
    assert_same(built, reference.arrays(ptr), "amino repeats")
    built.close()
    reference.dealloc_index(ptr)


def test_duplicated_megabase_segments(reference, tmp_path):
    """12 Mbp in which every suffix has three copies 3 Mbp apart (a segmental duplication at genome scale): all 12 M
    suffixes tie after the radix pass and stay tied for ~17 doubling rounds."""
    segment = synth.random_text(3_000_000)
    text = np.concatenate([segment, segment, segment, segment])
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "dup.awfmi"), abi.AwFmAlphabetDna, 8, 16)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetDna, 8, 16)
    assert built.tie_suffixes > 11_900_000 and built.tie_rounds >= 17, (built.tie_suffixes, built.tie_rounds)
    assert_same(built, reference.arrays(ptr), "duplicated segments")
    built.close()
    reference.dealloc_index(ptr)


def test_repetitive_text_goes_through_tie_resolution(reference, tmp_path):
    """Repeats defeat the 22-symbol radix pass; the tied groups are finished by prefix doubling on the device."""
    rng = np.random.default_rng(0)
    unit = make_text(300, False, seed=1)
    text = np.concatenate([unit] * 20 + [make_text(500, False, seed=2)] + [np.frombuffer(b"ACGT" * 200, np.uint8)])
    text = text.copy()
    text[rng.integers(0, len(text), 5)] = ord("N")
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "rep.awfmi"), abi.AwFmAlphabetDna, 4, 5)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetDna, 4, 5)
    assert built.tie_suffixes > 1000
    assert_same(built, reference.arrays(ptr), "repetitive")
    built.close()
    reference.dealloc_index(ptr)


def test_device_generated_text_and_file_round_trip(reference, tmp_path):
    """Text generated on the device (same splitmix64 stream as synth.py), built on the device, written as .awfmi,
    read back by the reference's awFmReadIndexFromFile, and compared with the reference's own build."""
    import torch
    from avxwindowfmindex_b200 import capi
    n = 1_000_003
    d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
    capi.check(capi.load().awfm_gpu_synth_letters(0, d_text.data_ptr(), n, synth.TEXT_SEED, 0, 0))
    text = synth.random_text(n)
    assert np.array_equal(d_text.cpu().numpy(), text)
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), n, abi.AwFmAlphabetDna, 8, 8)
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "ref.awfmi"), abi.AwFmAlphabetDna, 8, 8)
    mine = assert_same(built, reference.arrays(ptr), "1Mbp")
    path = str(tmp_path / "mine.awfmi")
    write_awfmi(mine, path)
    assert open(path, "rb").read() == open(str(tmp_path / "ref.awfmi"), "rb").read()
    ptr2 = reference.read_index(path)
    q = synth.random_queries(20000, 12)
    assert np.array_equal(reference.count(ptr2, q, fixed_len=12, threads=4), reference.count(ptr, q, fixed_len=12, threads=4))
    # searchable straight from device memory, no host round trip
    gpu = built.gpu_index()
    counts = np.zeros(20000, np.uint32)
    d_q = torch.from_numpy(q).cuda()
    d_c = torch.zeros(20000, dtype=torch.int32, device="cuda")
    gpu.count_device(d_q.data_ptr(), None, 12, 20000, d_c.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_c.cpu().numpy().astype(np.uint32), reference.count(ptr, q, fixed_len=12, threads=4))
    gpu.close()
    built.close()
    reference.dealloc_index(ptr)
    reference.dealloc_index(ptr2)


def test_amino_device_text(reference, tmp_path):
    n = 300_007
    text = synth.random_text(n, amino=True)
    text[::5003] = ord("X")
    ptr = reference.create_index(text.tobytes(), str(tmp_path / "a.awfmi"), abi.AwFmAlphabetAmino, 3, 7)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetAmino, 3, 7)
    assert_same(built, reference.arrays(ptr), "amino")
    built.close()
    reference.dealloc_index(ptr)


def test_multi_fasta_text_matches_create_index_from_fasta(reference, tmp_path):
    """The text awFmCreateIndexFromFasta indexes is the records joined by NUL separators, one after the last record
    too (lib/FastaVector/src/FastaVector.c:54-170); the separators are sanitised to the ambiguity letter
    (src/AwFmLetter.c:24-42).  Built on the device from that text, the arrays equal the reference's."""
    lengths = synth.multi_fasta_lengths(300, 0, 900, seed=21)
    text, meta, header = synth.multi_fasta_text(lengths, seed=22)
    fasta = str(tmp_path / "m.fa")
    synth.write_fasta(fasta, text, meta)
    ptr = reference.create_index_from_fasta(fasta, str(tmp_path / "m.awfmi"), abi.AwFmAlphabetDna, 5, 8)
    ref_arrays = reference.arrays(ptr)
    assert np.array_equal(ref_arrays.fasta_metadata, meta)
    built = DeviceBuiltIndex.from_host_text(text, abi.AwFmAlphabetDna, 5, 8)
    assert_same(built, ref_arrays, "multi-fasta")
    built.close()
