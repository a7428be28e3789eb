"""Development probe (run under gpurun): random-gather roofline numbers and count/locate kernel timings for every
kernel variant on a reference-built index.  Writes gpurun_out/probe_*.json.  Not part of the product."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from avxwindowfmindex_b200 import GpuIndex, abi, capi, synth  # noqa: E402


def gather_table(lib, out):
    gb = C.c_double()
    rows = []
    for array_mb, rec, lanes, reads in [(2048, 256, 8, 1 << 26), (2048, 128, 8, 1 << 27), (2048, 128, 4, 1 << 27),
                                        (2048, 64, 4, 1 << 27), (2048, 64, 2, 1 << 27), (2048, 64, 1, 1 << 27),
                                        (2048, 32, 2, 1 << 27), (2048, 32, 1, 1 << 27), (256, 16, 1, 1 << 27),
                                        (2048, 16, 1, 1 << 27), (16384, 64, 4, 1 << 27)]:
        capi.check(lib.awfm_gpu_gather_bandwidth(0, array_mb << 20, rec, reads, lanes, C.byref(gb)))
        rows.append({"array_MB": array_mb, "record_B": rec, "lanes": lanes, "GBps": round(gb.value, 1),
                     "Greads_per_s": round(gb.value / rec, 3)})
        print(rows[-1], flush=True)
    out["gather"] = rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=50_000_000)
    ap.add_argument("--queries", type=int, default=20_000_000)
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--skip-gather", action="store_true")
    ap.add_argument("--only-gather", action="store_true")
    args = ap.parse_args()
    import torch
    from oracle import harness
    lib = capi.load()
    out = {"gpu": torch.cuda.get_device_name(0), "bp": args.bp}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not args.skip_gather:
        gather_table(lib, out)
    if args.only_gather:
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_gather.json"), "w"), indent=1)
        return
    from avxwindowfmindex_b200 import DeviceBuiltIndex
    t0 = time.time()
    d_text = torch.empty(args.bp, dtype=torch.uint8, device="cuda")
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), args.bp, synth.TEXT_SEED, 0, 0))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), args.bp, abi.AwFmAlphabetDna, args.seed_k, 8)
    del d_text
    torch.cuda.synchronize()
    out["build_s"] = round(time.time() - t0, 1)
    out["build_gpu_ms"] = built.build_ms
    out["tie_suffixes"] = built.tie_suffixes
    print("index built", out["build_s"], "s; gpu ms", built.build_ms, "ties", built.tie_suffixes, flush=True)
    t0 = time.time()
    arrays = built.to_host()
    print("download", round(time.time() - t0, 1), "s", flush=True)
    gpu = built.gpu_index()
    built.close()
    out["device_bytes"] = gpu.device_bytes()
    n, L = args.queries, 20
    q = torch.from_numpy(synth.random_queries(n, L)).cuda()
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    # exact algorithmic bytes from the oracle on a 1M-query sample
    sample = min(n, 1_000_000)
    oc, _, work = harness.Oracle(arrays).count(q[: sample * L].cpu().numpy(), fixed_len=L, threads=os.cpu_count())
    out["oracle_work_per_query"] = {k: v / sample for k, v in work.items()}
    rows = []
    for variant in (0, 1):
        for lpq in (8, 4, 2, 1):
            gpu.set_tuning(count_variant=variant, count_lpq=lpq)
            for _ in range(2):
                gpu.count_device(q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            reps = 3
            for _ in range(reps):
                gpu.count_device(q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            ok = bool(np.array_equal(d_counts[:sample].cpu().numpy().astype(np.uint32), oc))
            rows.append({"variant": variant, "lpq": lpq, "ms": round(ms, 3), "Mq_per_s": round(n / ms / 1e3, 1),
                         "alg_GBps": round(work["countBytes"] / sample * n / ms / 1e6, 1), "parity": ok})
            print(rows[-1], flush=True)
    out["count"] = rows
    # locate: random 12-mers on this index have ~bp/4^12 hits each
    n2, L2 = 2_000_000, 12
    q2 = torch.from_numpy(synth.random_queries(n2, L2, seed=7)).cuda()
    d_hit = torch.zeros(n2 + 1, dtype=torch.int64, device="cuda")
    gpu.set_tuning(count_variant=1, count_lpq=8)
    gpu.count_device(q2.data_ptr(), None, L2, n2, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), n2, d_hit.data_ptr(), stream)
    total = int(d_hit[-1].item())
    d_pos = torch.zeros(total, dtype=torch.int64, device="cuda")
    sample2 = 100_000
    oh, op, work2 = harness.Oracle(arrays).locate(q2[: sample2 * L2].cpu().numpy(), fixed_len=L2, threads=os.cpu_count())
    rows = []
    for lpq in (8, 4, 2, 1):
        gpu.set_tuning(locate_lpq=lpq)
        gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n2, 0, total, d_pos.data_ptr(), stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n2, 0, total, d_pos.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        ok = bool(np.array_equal(d_pos[: int(oh[-1])].cpu().numpy().astype(np.uint64), op))
        rows.append({"lpq": lpq, "hits": total, "ms": round(ms, 3), "Mhits_per_s": round(total / ms / 1e3, 1),
                     "alg_GBps": round(work2["locateBytes"] / max(1, work2["hits"]) * total / ms / 1e6, 1), "parity": ok})
        print(rows[-1], flush=True)
    out["locate"] = rows
    out["locate_work_per_hit"] = {k: v / max(1, work2["hits"]) for k, v in work2.items()}
    with open(os.path.join(ROOT, "gpurun_out", f"probe_{args.bp}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
