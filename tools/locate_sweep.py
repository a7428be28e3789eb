"""Backtrace-kernel sweep on sparse hits (BASELINE cfg 3 / cfg 4 shape): locate_variant x locate_lpq, device-resident,
CUDA-event timed, bit-exact check of every variant against the first.  Appends JSON lines to
gpurun_out/locate_sweep.jsonl.  Measurement tool, not product."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from avxwindowfmindex_b200 import DeviceBuiltIndex, abi, capi, synth  # noqa: E402


def sweep(name, amino, bp, seed_k, ratio, nq, L, reps=5):
    lib = capi.load()
    dev = torch.device("cuda:0")
    alphabet = abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna
    d_text = torch.empty(bp, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), bp, synth.TEXT_SEED + 3, 0, int(amino)))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), bp, alphabet, seed_k, ratio)
    del d_text
    gpu = built.gpu_index()
    built.close()
    torch.cuda.empty_cache()
    stream = torch.cuda.current_stream().cuda_stream
    d_q = torch.empty(nq * L + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), nq * L, synth.QUERY_SEED + 3, 0, int(amino)))
    d_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((nq, 2), dtype=torch.int64, device=dev)
    d_hit = torch.zeros(nq + 1, dtype=torch.int64, device=dev)
    count_rows = []
    first_counts = None
    for variant in (0, 1):
        for lpq in (1, 2, 4):
            gpu.set_tuning(count_variant=variant, count_lpq=lpq)
            best = 1e30
            for _ in range(reps + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                gpu.count_device(d_q.data_ptr(), None, L, nq, d_counts.data_ptr(), None, stream)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            if first_counts is None:
                first_counts = d_counts.clone()
            count_rows.append({"variant": variant, "lpq": lpq, "ms": round(best, 4), "Gq_per_s": round(nq / best / 1e6, 3),
                               "same_as_first": bool(torch.equal(d_counts, first_counts))})
    gpu.set_tuning(count_variant=1, count_lpq=2)
    gpu.count_device(d_q.data_ptr(), None, L, nq, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), nq, d_hit.data_ptr(), stream)
    total = int(d_hit[-1].item())
    d_pos = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
    first = None
    rows = []
    for variant in (0, 1):
        for lpq in (1, 2, 4):
            gpu.set_tuning(locate_variant=variant, locate_lpq=lpq)
            best = 1e30
            for _ in range(reps + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), nq, 0, total, d_pos.data_ptr(), stream)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            got = d_pos.clone()
            if first is None:
                first = got
            rows.append({"variant": variant, "lpq": lpq, "ms": round(best, 4), "Ghits_per_s": round(total / best / 1e6, 3),
                         "same_as_first": bool(torch.equal(got, first))})
    out = {"config": name, "hits": total, "queries": nq, "ratio": ratio, "count_rows": count_rows, "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "locate_sweep.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    print(json.dumps(out), flush=True)
    gpu.close()
    del d_q, d_counts, d_ranges, d_hit, d_pos
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="nuc8,nuc16,amino")
    a = ap.parse_args()
    if "nuc8" in a.which:
        sweep("3.1 Gbp, 10 M 16-mers, ratio 8", False, 3_100_000_000, 12, 8, 10_000_000, 16)
    if "nuc16" in a.which:
        sweep("3.1 Gbp, 10 M 16-mers, ratio 16", False, 3_100_000_000, 12, 16, 10_000_000, 16)
    if "dense" in a.which:
        sweep("3.1 Gbp, 40 M 14-mers, ratio 8", False, 3_100_000_000, 12, 8, 40_000_000, 14)
    if "count20" in a.which:
        sweep("3.1 Gbp, 100 M 20-mers, ratio 8", False, 3_100_000_000, 12, 8, 100_000_000, 20)
    if "amino" in a.which:
        sweep("1 G residues, 50 M 8-mers, ratio 8", True, 1_000_000_000, 5, 8, 50_000_000, 8)


if __name__ == "__main__":
    main()
