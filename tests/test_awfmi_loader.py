"""SURVEY.md §8 row f3 — awfm_gpu_ctx_create_from_file: the unchanged `.awfmi` version-8 file (src/AwFmFile.c:20-193)
goes from the page cache to HBM.  Format errors must surface before any CUDA call (so they are checked on the CPU
box too); on the GPU the loaded index must answer exactly like the one uploaded from host arrays."""
import glob
import os
import struct

import numpy as np
import pytest

from avxwindowfmindex_b200 import GpuIndex, capi, read_awfmi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
ERR_ARG, ERR_NO_DEVICE, ERR_NO_SA = -1, -2, -5


def test_format_errors_are_reported_without_a_device(tmp_path):
    good = open(os.path.join(GOLDEN, "nuc_k4_r4.awfmi"), "rb").read()
    cases = {
        "magic": b"AwFmIndeX\n" + good[10:],
        "version": good[:10] + struct.pack("<I", 7) + good[14:],
        "alphabet": good[:20] + b"\x09" + good[21:],
        "truncated": good[: len(good) // 2],
        "short": good[:12],
        "ratio0": good[:18] + b"\x00" + good[19:],
    }
    for name, blob in cases.items():
        path = tmp_path / (name + ".awfmi")
        path.write_bytes(blob)
        with pytest.raises(capi.AwfmGpuError) as e:
            GpuIndex.from_file(path)
        assert e.value.code == ERR_ARG, name
    with pytest.raises(capi.AwfmGpuError) as e:
        GpuIndex.from_file(tmp_path / "does_not_exist.awfmi")
    assert e.value.code == ERR_ARG
    # a FASTA index whose record table is cut off
    fasta = open(os.path.join(GOLDEN, "four_records.awfmi"), "rb").read()
    (tmp_path / "cut.awfmi").write_bytes(fasta[:-8])
    with pytest.raises(capi.AwfmGpuError) as e:
        GpuIndex.from_file(tmp_path / "cut.awfmi")
    assert e.value.code == ERR_ARG


def test_valid_file_without_a_device_fails_loudly():
    if capi.load().awfm_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.AwfmGpuError) as e:
        GpuIndex.from_file(os.path.join(GOLDEN, "nuc_k4_r4.awfmi"))
    assert e.value.code == ERR_NO_DEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_loaded_index_answers_like_the_golden(case):
    path = os.path.join(GOLDEN, case + ".awfmi")
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    arrays = read_awfmi(path)
    gpu = GpuIndex.from_file(path)
    i = gpu.info
    assert (i.bwtLength, i.alphabetType, i.kmerLengthInSeedTable, i.suffixArrayCompressionRatio, i.versionNumber) == \
        (arrays.bwt_length, arrays.alphabet, arrays.seed_k, arrays.sa_ratio, 8)
    assert i.suffixArrayByteLength == len(arrays.sa_bytes) and i.featureFlags == arrays.feature_flags
    counts, ranges = gpu.count(g["letters"], g["offsets"], want_ranges=True)
    assert np.array_equal(counts, g["counts"])
    hit, pos = gpu.locate(g["letters"], g["offsets"])
    assert np.array_equal(hit, g["hit_offsets"]) and np.array_equal(pos, g["positions"])
    uploaded = GpuIndex(arrays)
    c2, r2 = uploaded.count(g["letters"], g["offsets"], want_ranges=True)
    assert np.array_equal(ranges, r2) and gpu.device_bytes() == uploaded.device_bytes()
    uploaded.close()
    if "contig_of_hit" in g.files:  # the record table in the file is installed by the loader
        assert i.numSequences == len(arrays.fasta_metadata)
        seq, loc, bad = gpu.map_positions(pos)
        assert bad == 0 and np.array_equal(np.stack([seq, loc], axis=1), g["contig_of_hit"])
    gpu.close()


@pytest.mark.gpu
def test_count_only_load_refuses_locate():
    g = np.load(os.path.join(GOLDEN, "nuc_k4_r4.npz"))
    gpu = GpuIndex.from_file(os.path.join(GOLDEN, "nuc_k4_r4.awfmi"), want_suffix_array=False)
    assert np.array_equal(gpu.count(g["letters"], g["offsets"]), g["counts"])
    with pytest.raises(capi.AwfmGpuError) as e:
        gpu.locate(g["letters"], g["offsets"])
    assert e.value.code == ERR_NO_SA
    gpu.close()
