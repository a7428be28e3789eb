"""The drop-in claim with a real C program (INTEGRATION.md sections 1 and 2): tests/c_consumer/consumer.c is a plain user
of the reference's public API, compiled against the reference's OWN header (/root/reference/src/AwFmIndex.h) by
oracle/Makefile where that tree exists.  The same program must print byte-identical results
  * linked against the reference only,
  * linked with libawfm_b200.so ahead of the reference on the link line (the four entry points bind to ours),
  * linked against the reference only and run under LD_PRELOAD=libawfm_b200.so,
and the two drop-in runs must really have been answered by the CUDA engine (the program reports on stderr whether the
additive awFmGpu* symbols are present in the process)."""
import os
import subprocess

import pytest

from avxwindowfmindex_b200 import capi
from oracle import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CONSUMER_REF = os.path.join(REF_DIR, "consumer_ref")
CONSUMER_LINKED = os.path.join(REF_DIR, "consumer_linked")


def run(binary, alphabet, path, preload=None):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    return subprocess.run([binary, alphabet, path], env=env, capture_output=True, text=True, timeout=600)


def test_consumer_binaries_bind_as_documented():
    """CPU box: the link-order binary resolves the four entry points to our library, the reference-only one does not
    even load it (no compute: only the dynamic symbol tables are inspected)."""
    if not os.path.exists(CONSUMER_LINKED):
        pytest.skip("oracle/_ref/consumer_* not built (needs /root/reference at build time)")
    needed = subprocess.run(["readelf", "-d", CONSUMER_LINKED], capture_output=True, text=True).stdout
    order = [line.split("[")[1].rstrip("]") for line in needed.splitlines() if "NEEDED" in line and "awfm" in line]
    assert order == ["libawfm_b200.so", "libawfm_ref.so"], order  # ours first: its symbols win the lookup
    needed = subprocess.run(["readelf", "-d", CONSUMER_REF], capture_output=True, text=True).stdout
    assert "libawfm_b200.so" not in needed and "libawfm_ref.so" in needed


@pytest.mark.gpu
@pytest.mark.parametrize("alphabet", ["dna", "amino"])
def test_c_program_gets_identical_results_through_the_drop_in(alphabet, tmp_path):
    if not (os.path.exists(CONSUMER_REF) and os.path.exists(CONSUMER_LINKED) and harness.have_reference()):
        pytest.skip("oracle/_ref/consumer_* not built (needs /root/reference at build time)")
    expected = run(CONSUMER_REF, alphabet, str(tmp_path / "ref.awfmi"))
    assert expected.returncode == 0 and "engine: reference" in expected.stderr, expected.stderr[-500:]
    linked = run(CONSUMER_LINKED, alphabet, str(tmp_path / "linked.awfmi"))
    assert linked.returncode == 0, linked.stderr[-2000:]
    assert "engine: b200 drop-in devices=1" in linked.stderr, linked.stderr[-500:]
    assert linked.stdout == expected.stdout
    preloaded = run(CONSUMER_REF, alphabet, str(tmp_path / "preload.awfmi"), preload=capi.LIB_PATH)
    assert preloaded.returncode == 0, preloaded.stderr[-2000:]
    assert "engine: b200 drop-in devices=1" in preloaded.stderr, preloaded.stderr[-500:]
    assert preloaded.stdout == expected.stdout
    assert "locate: rc=1" in expected.stdout and "hits=" in expected.stdout
