"""Development probe (run under gpurun): sweep count path vs the tile kernel on the bench workload — total time,
per-stage device time, sort-bits and batch-size sweeps; every variant's counts compared with the tile kernel's over
the whole batch.  Writes gpurun_out/sweep_probe.jsonl.  Not part of the product."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from avxwindowfmindex_b200 import DeviceBuiltIndex, abi, capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=3_100_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--kmer", type=int, default=20)
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--bits", type=str, default="16,24,8,12,20")
    ap.add_argument("--sizes", type=str, default="50000000,25000000,12500000,4194304")
    ap.add_argument("--deep", type=int, default=0)
    ap.add_argument("--items", type=str, default="4")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--local", type=str, default="8")
    ap.add_argument("--first-items", type=int, default=4)
    ap.add_argument("--no-tile", action="store_true")
    ap.add_argument("--amino", action="store_true", help="cfg 4 shape: pass --bp 1000000000 --queries 50000000 --kmer 8 --seed-k 5")
    ap.add_argument("--own-sort", type=str, default="1,0", help="ordering step: 1 = csrc/awfm_sort.cuh, 0 = CUB")
    ap.add_argument("--variable", type=str, default="", help="lo,hi: also time a variable-length batch of --queries queries with lengths uniform in lo..hi (tile kernel vs sweep)")
    ap.add_argument("--wide", action="store_true", help="also time the 64-bit-position passes (sweep_wide=1) on the same batch")
    ap.add_argument("--nvtx", action="store_true", help="wrap one extra sweep call in the NVTX range 'sweepcall' (for ncu --nvtx)")
    args = ap.parse_args()
    lib = capi.load()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "sweep_probe.jsonl"), "a")

    def emit(row):
        print(json.dumps(row), flush=True)
        out.write(json.dumps(row) + "\n")
        out.flush()

    d_text = torch.empty(args.bp, dtype=torch.uint8, device="cuda")
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), args.bp, synth.TEXT_SEED + 2, 0, int(args.amino)))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), args.bp,
                                              abi.AwFmAlphabetAmino if args.amino else abi.AwFmAlphabetDna, args.seed_k, 8)
    del d_text
    gpu = built.gpu_index()
    built.close()
    torch.cuda.empty_cache()
    n, L = args.queries, args.kmer
    d_letters = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
    capi.check(lib.awfm_gpu_synth_letters(0, d_letters.data_ptr(), n * L, synth.QUERY_SEED + 2, 0, int(args.amino)))
    d_ref = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(nq, dst, reps=args.reps):
        best = 1e30
        for _ in range(reps + 1):
            a.record(stream)
            gpu.count_device(d_letters.data_ptr(), None, L, nq, dst.data_ptr(), None, stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    gpu.set_tuning(sweep_min_queries=-1)
    ms = 0.0 if args.no_tile else timed(n, d_ref)
    emit({"variant": "tile", "queries": n, "ms": ms, "Gq_per_s": n / max(ms, 1e-9) / 1e6, "hits": int(d_ref.sum(dtype=torch.int64))})
    gpu.set_tuning(sweep_min_queries=1, sweep_profile=1)
    for own, items, bits, local in [(int(o), int(i), int(x), int(l)) for o in args.own_sort.split(",")
                                    for i in args.items.split(",") for x in args.bits.split(",") if x
                                    for l in args.local.split(",")]:
        gpu.set_tuning(sweep_sort_bits=bits, sweep_items=items, sweep_local_bits=local, sweep_first_items=args.first_items,
                       sweep_own_sort=own)
        d_counts.fill_(-1)
        ms = timed(n, d_counts)
        emit({"variant": "sweep", "own_sort": own, "amino": args.amino, "sort_bits": bits, "local_bits": local, "items": items, "first_items": args.first_items, "queries": n, "ms": ms, "Gq_per_s": n / ms / 1e6,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()],
              "equal_to_tile": None if args.no_tile else bool(torch.equal(d_counts, d_ref)),
              "device_bytes": gpu.device_bytes()})
    for compact in (1, 0):  # 8-byte words through the bucket passes (csrc/awfm_sort.cuh) vs 4 + 8 bytes
        gpu.set_tuning(sweep_sort_bits=32, sweep_items=4, sweep_local_bits=-1, sweep_own_sort=1, sweep_compact_pairs=compact)
        d_counts.fill_(-1)
        ms = timed(n, d_counts)
        emit({"variant": "compact_pairs", "on": compact, "queries": n, "ms": ms, "Gq_per_s": n / ms / 1e6,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()],
              "equal_to_tile": None if args.no_tile else bool(torch.equal(d_counts, d_ref))})
    gpu.set_tuning(sweep_compact_pairs=1)
    if args.wide:
        gpu.set_tuning(sweep_sort_bits=32, sweep_items=4, sweep_local_bits=-1, sweep_own_sort=1, sweep_wide=1)
        d_counts.fill_(-1)
        ms = timed(n, d_counts)
        emit({"variant": "sweep_wide", "queries": n, "ms": ms, "Gq_per_s": n / ms / 1e6,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()],
              "equal_to_tile": None if args.no_tile else bool(torch.equal(d_counts, d_ref))})
        gpu.set_tuning(sweep_wide=0)
    if args.nvtx:
        gpu.set_tuning(sweep_sort_bits=32, sweep_items=4, sweep_profile=0)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("sweepcall")
        gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    gpu.set_tuning(sweep_sort_bits=32, sweep_own_sort=1)
    for nq in [int(x) for x in args.sizes.split(",") if x]:
        gpu.set_tuning(sweep_min_queries=-1)
        t_tile = timed(nq, d_ref, reps=2)
        gpu.set_tuning(sweep_min_queries=1)
        d_counts.fill_(-1)
        t_sweep = timed(nq, d_counts, reps=2)
        emit({"variant": "size", "queries": nq, "tile_ms": t_tile, "sweep_ms": t_sweep,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()],
              "equal_to_tile": bool(torch.equal(d_counts[:nq], d_ref[:nq]))})
    if args.variable:
        lo, hi = (int(x) for x in args.variable.split(","))
        g = torch.Generator(device="cuda").manual_seed(7)
        lengths = torch.randint(lo, hi + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
        d_offsets = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        torch.cumsum(lengths, 0, out=d_offsets[1:])
        total = int(d_offsets[-1])
        del lengths
        d_var = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
        capi.check(lib.awfm_gpu_synth_letters(0, d_var.data_ptr(), total, synth.QUERY_SEED + 5, 0, int(args.amino)))

        def timed_var(dst, reps=3):
            best = 1e30
            for _ in range(reps + 1):
                a.record(stream)
                gpu.count_device(d_var.data_ptr(), d_offsets.data_ptr(), 0, n, dst.data_ptr(), None, stream.cuda_stream)
                b.record(stream)
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            return best

        gpu.set_tuning(sweep_min_queries=-1)
        t_tile = timed_var(d_ref)
        gpu.set_tuning(sweep_min_queries=1, sweep_profile=1, sweep_items=4, sweep_local_bits=-1)
        d_counts.fill_(-1)
        t_sweep = timed_var(d_counts)
        emit({"variant": "variable", "lengths": [lo, hi], "queries": n, "letters": total, "tile_ms": t_tile, "sweep_ms": t_sweep,
              "tile_Gq_per_s": n / t_tile / 1e6, "sweep_Gq_per_s": n / t_sweep / 1e6,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()], "equal_to_tile": bool(torch.equal(d_counts, d_ref)),
              "hits": int(d_ref.sum(dtype=torch.int64))})
        del d_var, d_offsets
    if args.deep:
        gpu.extend_seed_table(args.deep)
        gpu.set_tuning(sweep_min_queries=-1)
        t_tile = timed(n, d_ref, reps=2)
        gpu.set_tuning(sweep_min_queries=1)
        d_counts.fill_(-1)
        t_sweep = timed(n, d_counts, reps=2)
        emit({"variant": "deep", "depth": args.deep, "queries": n, "tile_ms": t_tile, "sweep_ms": t_sweep,
              "stage_ms": [round(x, 3) for x in gpu.sweep_stage_ms()], "equal_to_tile": bool(torch.equal(d_counts, d_ref))})
    gpu.close()


if __name__ == "__main__":
    main()
