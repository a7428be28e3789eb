import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from avxwindowfmindex_b200 import abi  # noqa: E402
from oracle import harness  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are SKIPPED (not failed) on a box without a usable CUDA device, so a plain `pytest` is green on
    the CPU box and the GPU box runs everything.  The product itself still fails loudly without a device
    (tests/test_library_symbols.py::test_no_device_means_loud_failure)."""
    try:
        from avxwindowfmindex_b200 import capi
        have_gpu = capi.load().awfm_gpu_device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    harness.build(ref=True)


@pytest.fixture(scope="session")
def reference():
    if not harness.have_reference():
        pytest.skip("oracle/_ref/libawfm_ref.so not built (needs /root/reference at build time)")
    return harness.Reference()


def make_text(n, amino, seed, ambiguity_every=0, mixed_case=False):
    """Random text; optionally sprinkles ambiguity letters and lower-case (nucleotide only: the amino sanitizer does
    not fold case, src/AwFmLetter.c:69-79)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY" if amino else b"ACGT", dtype=np.uint8)
    text = alphabet[rng.integers(0, len(alphabet), n)]
    if ambiguity_every:
        pos = rng.integers(0, n, max(1, n // ambiguity_every))
        text[pos] = ord("X") if amino else ord("N")
    if mixed_case and not amino:
        lower = rng.random(n) < 0.3
        text = np.where(lower, text | 0x20, text).astype(np.uint8)
    return text


def make_queries(text, amino, seed, num, min_len, max_len, seed_k, frac_sampled=0.6, ambiguity=True):
    """Mixed query set: substrings of the text (hits), random strings (mostly misses), queries shorter than the
    seed length, queries with ambiguity letters inside / outside the seed window, lower-case letters."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY" if amino else b"ACGT", dtype=np.uint8)
    out = []
    for i in range(num):
        length = int(rng.integers(min_len, max_len + 1))
        if rng.random() < frac_sampled and length <= len(text):
            start = int(rng.integers(0, len(text) - length + 1))
            q = text[start:start + length].copy()
        else:
            q = alphabet[rng.integers(0, len(alphabet), length)]
        r = rng.random()
        if ambiguity and r < 0.08:
            q[int(rng.integers(0, length))] = ord("x") if amino else ord("n")
        elif ambiguity and r < 0.12 and not amino:
            q = q | 0x20  # lower case
        out.append(q.astype(np.uint8).tobytes())
    # a few deterministic edge cases
    out.append(bytes(alphabet[:1]))                      # single letter
    out.append(bytes(text[-min(len(text), max_len):]))  # suffix of the text (touches the sentinel neighbourhood)
    out.append(bytes(text[:min(len(text), max_len)]))   # prefix of the text (backtrace reaches the sentinel)
    if seed_k > 1:
        out.append(bytes(text[5:5 + seed_k - 1]))        # shorter than the seed length
    out.append(bytes(text[7:7 + seed_k]))                # exactly the seed length
    return out


class BuiltIndex:
    def __init__(self, reference, tmpdir, name, text, amino, seed_k, sa_ratio):
        self.reference = reference
        self.path = os.path.join(tmpdir, name + ".awfmi")
        self.text = text
        self.amino = amino
        alphabet = abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna
        self.ptr = reference.create_index(text.tobytes(), self.path, alphabet, seed_k, sa_ratio)
        self.arrays = reference.arrays(self.ptr)


@pytest.fixture(scope="session")
def small_indexes(reference, tmp_path_factory):
    """A spread of reference-built indexes: both alphabets, SA ratios 1/2/3/8/16/200/255, ambiguity letters."""
    tmp = str(tmp_path_factory.mktemp("awfm"))
    specs = [
        ("nuc_r8", False, 30011, 5, 8, 97, True),
        ("nuc_r1", False, 5003, 4, 1, 0, False),
        ("nuc_r3", False, 8009, 6, 3, 211, True),
        ("nuc_r16", False, 70001, 7, 16, 0, False),
        ("nuc_r200", False, 3001, 3, 200, 50, False),
        ("nuc_r255", False, 2000, 2, 255, 0, False),
        ("amino_r8", True, 20011, 3, 8, 101, False),
        ("amino_r2", True, 4001, 2, 2, 0, False),
        ("amino_r1", True, 1500, 1, 1, 37, False),
    ]
    out = {}
    for name, amino, n, k, ratio, amb, mixed in specs:
        text = make_text(n, amino, seed=zlib.crc32(name.encode()) % 10007, ambiguity_every=amb, mixed_case=mixed)  # stable across runs
        out[name] = BuiltIndex(reference, tmp, name, text, amino, k, ratio)
    return out
