#!/usr/bin/env python
"""Builds bench.py's index with the UNMODIFIED reference (awFmCreateIndex + libdivsufsort, oracle/_ref) on the CPU and
records SHA-256 digests of every section, so a run on the GPU box (where /root/reference does not exist and a 20-minute
CPU build does not fit) can prove that the device-built index it searches is byte-identical to the reference's.

    python tools/ref_index_hashes.py --bp 3100000000 --seed-k 12 --sa-ratio 8 --out tests/golden/cfg2_index_sha256.json

Needs ~35 GB of host memory at 3.1 Gbp (text + sanitized copy + 8-byte suffix array).  TEST INFRASTRUCTURE: only
bench.py's checker and tests read the JSON it writes.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from avxwindowfmindex_b200 import abi, synth  # noqa: E402
from avxwindowfmindex_b200.index import section_digests  # noqa: E402
from oracle import harness  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=3_100_000_000)
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--sa-ratio", type=int, default=8)
    ap.add_argument("--amino", action="store_true")
    ap.add_argument("--text-seed", type=lambda s: int(s, 0), default=synth.TEXT_SEED + 2)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "cfg2_index_sha256.json"))
    ap.add_argument("--awfmi", default="/tmp/ref_index_hashes.awfmi")
    args = ap.parse_args()

    t0 = time.time()
    text = np.empty(args.bp, dtype=np.uint8)
    step = 1 << 26
    for s in range(0, args.bp, step):  # same stream as awfm_gpu_synth_letters(seed, start=0)
        n = min(step, args.bp - s)
        text[s:s + n] = synth.letters(args.text_seed, n, args.amino, start=s)
    print(f"text generated in {time.time() - t0:.0f} s", flush=True)

    ref = harness.Reference()
    alphabet = abi.AwFmAlphabetAmino if args.amino else abi.AwFmAlphabetDna
    cfg = abi.AwFmIndexConfiguration(args.sa_ratio, args.seed_k, alphabet, True, False)
    out = C.c_void_p()
    if os.path.exists(args.awfmi):
        os.remove(args.awfmi)
    t1 = time.time()
    rc = ref.lib.awFmCreateIndex(C.byref(out), C.byref(cfg), text.ctypes.data, len(text), args.awfmi.encode())
    if rc < 0:
        raise SystemExit(f"awFmCreateIndex failed: {rc}")
    build_s = time.time() - t1
    print(f"awFmCreateIndex: {build_s:.0f} s", flush=True)
    arrays = ref.arrays(out, copy=False)
    digests = section_digests(arrays)
    record = {
        "what": "SHA-256 of every section of the index the reference's awFmCreateIndex builds for bench.py's text",
        "built_by": "unmodified reference (oracle/_ref/libawfm_ref.so: awFmCreateIndex + divsufsort64), CPU, tools/ref_index_hashes.py",
        "text": {"generator": "splitmix64 (avxwindowfmindex_b200/synth.py)", "seed": args.text_seed, "length": args.bp,
                 "alphabet": "amino" if args.amino else "nucleotide"},
        "seed_k": args.seed_k, "sa_ratio": args.sa_ratio, "bwt_length": int(arrays.bwt_length),
        "build_seconds": round(build_s, 1),
        "sections": digests,
    }
    with open(args.out, "w") as f:
        json.dump(record, f, indent=1)
        f.write("\n")
    print(json.dumps(record["sections"], indent=1)[:2000])
    ref.dealloc_index(out)


if __name__ == "__main__":
    main()
