"""BASELINE.json configs[4]: multi-sequence synthetic FASTA (10 k contigs, ~1 Gbp), locate 32-mers with the sampled SA
in HBM, then map every hit to (contig, offset) on the device, query-sharded over N GPUs with
the index replicated and the per-rank CSR results gathered onto rank 0 over NCCL (the path's only collective).

    python tools/cfg5_multiseq.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/cfg5_multiseq.py                                      N GPUs, weak scaling (queries per GPU fixed)

Queries are substrings cut from inside random contigs, so every query has >= 1 hit and a by-construction answer
(contig, offset) that is checked for ALL queries on the device; a bounded sample is also searched and mapped by the
unmodified reference on the host cores (oracle/_ref) and compared element-wise.  One JSON line, appended to
gpurun_out/cfg5.jsonl.  Measurement tool, not product (the product is the C-ABI it calls)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from avxwindowfmindex_b200 import DeviceBuiltIndex, KmerSearchList, abi, capi, sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=10_000)
    ap.add_argument("--min-len", type=int, default=50_000)
    ap.add_argument("--max-len", type=int, default=150_000)
    ap.add_argument("--queries", type=int, default=10_000_000, help="queries per GPU")
    ap.add_argument("--kmer", type=int, default=32)
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--sa-ratio", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()

    # ---- the FASTA text (records + NUL separators), generated on the device; index built there too ----
    lengths = synth.multi_fasta_lengths(a.records, a.min_len, a.max_len, seed=synth.TEXT_SEED + 5)
    ends = np.cumsum(lengths + 1)
    total = int(ends[-1])
    header_ends = np.cumsum([len(b"contig%d" % i) + 1 for i in range(a.records)])
    meta = np.stack([header_ends.astype(np.uint64), ends.astype(np.uint64)], axis=1)
    t0 = time.time()
    d_text = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(local, d_text.data_ptr(), total, synth.TEXT_SEED + 5, 0, 0))
    d_text[torch.from_numpy(ends - 1).to(dev)] = 0
    torch.cuda.synchronize()
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), total, abi.AwFmAlphabetDna, a.seed_k, a.sa_ratio,
                                              device=local)
    gpu = built.gpu_index()
    gpu.set_sequences(meta)
    arrays = built.to_host() if rank == 0 else None
    build_ms, ties = built.build_ms, built.tie_suffixes
    built.close()
    setup_s = time.time() - t0

    # ---- this rank's shard of the global query stream: cut from inside random contigs ----
    n, L = a.queries, a.kmer
    starts = np.concatenate([[0], ends[:-1]])
    z = synth.splitmix64(synth.QUERY_SEED + 5, rank * 2 * n, 2 * n)
    rec = (z[:n] % np.uint64(a.records)).astype(np.int64)
    off = (z[n:] % (lengths[rec] - L + 1).astype(np.uint64)).astype(np.int64)
    g = starts[rec] + off
    d_g = torch.from_numpy(g).to(dev)
    d_rec, d_off = torch.from_numpy(rec).to(dev), torch.from_numpy(off).to(dev)
    d_q = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    step = 1 << 20
    ar = torch.arange(L, device=dev)
    for s in range(0, n, step):
        e = min(n, s + step)
        d_q[s * L:e * L] = d_text[(d_g[s:e, None] + ar[None, :]).reshape(-1)]
    del d_text
    torch.cuda.empty_cache()

    stream = torch.cuda.current_stream().cuda_stream
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    d_hit = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
    hits = int(d_hit[-1].item())
    d_pos = torch.zeros(hits, dtype=torch.int64, device=dev)
    d_seq = torch.zeros(hits, dtype=torch.int64, device=dev)
    d_loc = torch.zeros(hits, dtype=torch.int64, device=dev)

    def search():
        gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
        gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
        gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, hits, d_pos.data_ptr(), stream)
        gpu.map_positions_device(d_pos.data_ptr(), hits, d_seq.data_ptr(), d_loc.data_ptr(), stream)

    gathered = {}

    def step_fn():
        search()
        if world > 1:  # rank 0 ends up with the global CSR of (position, contig, offset)
            h, p = sharding.gather_hits(d_hit, torch.stack([d_pos, d_seq, d_loc], dim=1), n * world)
            gathered.update(hit=h, payload=p)

    for _ in range(a.warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        step_fn()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1) / a.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # kernel-only breakdown on this rank
    def timed(fn):
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            a0.record()
            fn()
            b0.record()
            torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(b0))
        return best
    ms_count = timed(lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream))
    ms_bt = timed(lambda: gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, hits, d_pos.data_ptr(), stream))
    ms_map = timed(lambda: gpu.map_positions_device(d_pos.data_ptr(), hits, d_seq.data_ptr(), d_loc.data_ptr(), stream))

    # ---- by-construction check of EVERY query of this rank: (contig, offset) it was cut from is among its hits ----
    cnt = (d_hit[1:] - d_hit[:-1])
    q_of_hit = torch.repeat_interleave(torch.arange(n, device=dev), cnt)
    match = (d_seq == d_rec[q_of_hit]) & (d_loc == d_off[q_of_hit]) & (d_pos == d_g[q_of_hit])
    found = torch.zeros(n, dtype=torch.int32, device=dev).index_add_(0, q_of_hit, match.to(torch.int32))
    all_found = bool((found >= 1).all().item()) and bool((cnt >= 1).all().item())
    inside = bool(((d_loc + L) <= torch.from_numpy(lengths).to(dev)[d_seq]).all().item())
    ok = torch.tensor([int(all_found and inside)], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:  # the gathered CSR equals rank 0's own shard at its head
            same = torch.equal(gathered["payload"][:hits], torch.stack([d_pos, d_seq, d_loc], dim=1))
            ok[0] = min(int(ok.item()), int(same))
    total_hits = torch.tensor([hits], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_hits)
    if rank != 0:
        dist.destroy_process_group()
        return

    out = {"config": "cfg5 multi-sequence FASTA locate + contig mapping", "n_gpus": world, "records": a.records,
           "text": total, "seed_k": a.seed_k, "sa_ratio": a.sa_ratio, "kmer": L, "queries_per_gpu": n,
           "hits_total": int(total_hits.item()), "ms_per_step": ms,
           "locate_queries_per_s": world * n / ms * 1e3, "located_and_mapped_hits_per_s": int(total_hits.item()) / ms * 1e3,
           "kernel_ms_rank0": {"count_with_ranges": ms_count, "expand+backtrace": ms_bt, "contig_map": ms_map},
           "contig_map_hits_per_s": hits / ms_map * 1e3,
           "every_query_found_at_its_origin": bool(ok.item()), "index_build_gpu_ms": build_ms, "tie_suffixes": ties,
           "setup_s": round(setup_s, 1), "device_bytes": gpu.device_bytes(), "scaling": "weak",
           "gather": "NCCL gather of padded per-rank CSR segments onto rank 0, inside the timed step" if world > 1 else None}

    # ---- the unmodified reference on the host cores: bounded sample, compared element-wise ----
    from oracle import harness
    ns = min(a.cpu_sample, n)
    hq = d_q[: ns * L].cpu().numpy()
    h_hit = d_hit[: ns + 1].cpu().numpy().astype(np.uint64)
    nh = int(h_hit[-1])
    h_pos, h_seq, h_loc = (x[:nh].cpu().numpy().astype(np.uint64) for x in (d_pos, d_seq, d_loc))
    if harness.have_reference():
        ref = harness.Reference()
        ix = arrays.as_awfm_index()
        fv = abi.FastaVector()
        fv.metadata.data = meta.ctypes.data
        fv.metadata.count = fv.metadata.capacity = len(meta)
        ix.fastaVector = C.addressof(fv)
        ix.featureFlags = 1
        ip = C.addressof(ix)
        threads = os.cpu_count()
        sl = KmerSearchList(ref.lib, ns).fill(hq, fixed_len=L)
        ref.lib.awFmParallelSearchLocate(ip, sl.ptr, threads)
        t1 = time.perf_counter()
        rc = ref.lib.awFmParallelSearchLocate(ip, sl.ptr, threads)
        t_loc = time.perf_counter() - t1
        r_pos = np.concatenate(sl.positions())
        same_pos = rc == abi.AwFmSuccess and np.array_equal(r_pos, h_pos)
        t1 = time.perf_counter()
        m = min(nh, 20_000)
        same_map = all(ref.contig_of(ip, int(h_pos[i])) == (abi.AwFmSuccess, int(h_seq[i]), int(h_loc[i])) for i in range(m))
        out["cpu_reference"] = {"queries": ns, "cores": threads, "locate_queries_per_s": ns / t_loc,
                                "located_hits_per_s": nh / t_loc, "positions_bit_exact": bool(same_pos),
                                "contig_mapping_checked_hits": m, "contig_mapping_identical": bool(same_map)}
        sl.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "cfg5.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    print(json.dumps(out), flush=True)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
