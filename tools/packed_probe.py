#!/usr/bin/env python
"""End-to-end probe of the packed-batch engine (awfm_gpu_group_count) at BASELINE cfg 2 size: host buffers in, host
counts out, H2D/D2H inside the timed region.  Sweeps the chunk size and the query format.  One JSON line per config."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pack_2bit_torch(d_letters, n, L):
    """2-bit packing on the GPU (setup only): same format as avxwindowfmindex_b200.search.pack_queries_bits."""
    import torch
    out = torch.empty((n, (L + 3) // 4), dtype=torch.uint8, device=d_letters.device)
    step = 1 << 24
    shifts = (2 * torch.arange(L, device=d_letters.device, dtype=torch.int64))
    for a in range(0, n, step):
        m = min(step, n - a)
        w = d_letters[a * L:(a + m) * L].view(m, L).to(torch.int64)
        code = ((w >> 1) ^ (w >> 2)) & 3
        acc = (code << shifts).sum(dim=1)  # L <= 31
        by = torch.stack([(acc >> (8 * i)) & 0xFF for i in range((L + 3) // 4)], dim=1).to(torch.uint8)
        out[a:a + m] = by
    return out.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=3_100_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--kmer", type=int, default=20)
    ap.add_argument("--chunks", default="8388608,16777216,33554432,50000128")
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    from avxwindowfmindex_b200 import DeviceBuiltIndex, GpuGroup, PinnedArray, abi, capi, synth
    from avxwindowfmindex_b200.search import QUERY_2BIT, QUERY_ASCII
    lib = capi.load()
    dev = torch.device("cuda:0")
    d_text = torch.empty(args.bp, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), args.bp, synth.TEXT_SEED + 2, 0, 0))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), args.bp, abi.AwFmAlphabetDna, 12, 8, device=0)
    del d_text
    gpu = built.gpu_index()
    built.close()
    torch.cuda.empty_cache()
    n, L = args.queries, args.kmer
    d_letters = torch.empty(n * L, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_letters.data_ptr(), n * L, synth.QUERY_SEED + 2, 0, 0))
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = d_counts.cpu().numpy().astype(np.uint32)
    h_ascii = PinnedArray(n * L, np.uint8)
    h_ascii.array[:] = d_letters.cpu().numpy()
    packed = pack_2bit_torch(d_letters, n, L)
    h_bits = PinnedArray(packed.numel(), np.uint8)
    h_bits.array[:] = packed.cpu().numpy()
    del packed, d_letters
    h_counts = PinnedArray(n, np.uint32)
    group = GpuGroup(indexes=[gpu])
    for fmt, name, buf in ((QUERY_2BIT, "2bit", h_bits), (QUERY_ASCII, "ascii", h_ascii)):
        for chunk in [int(c) for c in args.chunks.split(",")]:
            group.set_tuning(packed_chunk_queries=chunk)
            h_counts.array[:] = 0xFFFFFFFF
            group.count(buf.array, fmt, fixed_len=L, out=h_counts.array)
            ok = bool(np.array_equal(h_counts.array, want))
            times = []
            for _ in range(args.reps):
                t0 = time.perf_counter()
                group.count(buf.array, fmt, fixed_len=L, out=h_counts.array)
                times.append(time.perf_counter() - t0)
            best = min(times)
            print(json.dumps({"format": name, "chunk_queries": chunk, "queries": n, "best_ms": 1e3 * best,
                              "mean_ms": 1e3 * sum(times) / len(times), "queries_per_s": n / best,
                              "bit_exact_vs_device_path": ok, "launches": group.stats()["launches"]}), flush=True)
    # pageable buffers (staged through page-locked memory by the library)
    pageable = np.array(h_bits.array)
    out = np.zeros(n, np.uint32)
    group.set_tuning(packed_chunk_queries=16777216)
    group.count(pageable, QUERY_2BIT, fixed_len=L, out=out)
    t0 = time.perf_counter()
    group.count(pageable, QUERY_2BIT, fixed_len=L, out=out)
    dt = time.perf_counter() - t0
    print(json.dumps({"format": "2bit, pageable in/out", "queries": n, "best_ms": 1e3 * dt, "queries_per_s": n / dt,
                      "bit_exact_vs_device_path": bool(np.array_equal(out, want))}), flush=True)


if __name__ == "__main__":
    main()
