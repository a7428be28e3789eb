"""Development sweep of the drop-in locate path (awFmParallelSearchLocate on a host AwFmKmerSearchList): chunk size x
host threads x pinned/pageable query strings, with the engine's own phase timings (AWFM_GPU_VERBOSE=1) on stderr.
cfg 3 shape: 3.1 Gbp index (seed k=12, SA ratio RATIO), NQ random 16-mers.  One RESULT line per setting."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avxwindowfmindex_b200 import DeviceBuiltIndex, KmerSearchList, abi, capi, synth
lib = capi.load()
bp, n, L = int(os.environ.get("BP", 3_100_000_000)), int(os.environ.get("NQ", 10_000_000)), int(os.environ.get("KMER", 16))
ratio = int(os.environ.get("RATIO", 8))
d_text = torch.empty(bp, dtype=torch.uint8, device="cuda")
capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), bp, synth.TEXT_SEED + 2, 0, 0))
built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), bp, abi.AwFmAlphabetDna, 12, ratio)
del d_text
arrays = built.to_host(); built.close()
ix = arrays.as_awfm_index(); ip = C.addressof(ix)
d_q = torch.empty(n * L, dtype=torch.uint8, device="cuda")
capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), n * L, synth.QUERY_SEED + 3, 0, 0))
pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory(); pinned.copy_(d_q)
pageable = pinned.numpy().copy()
del d_q
sl = KmerSearchList(lib, n)
os.environ["AWFM_GPU_VERBOSE"] = "1"
cores = os.cpu_count()
chunks = [int(x) for x in os.environ.get("CHUNKS", "65536,131072,262144,524288,1048576").split(",")]
thread_list = [int(x) for x in os.environ.get("THREADS", f"{cores},{max(1, cores // 2)}").split(",")]
for src_name, src in (("pinned", pinned.numpy()), ("pageable", pageable)):
    sl.fill(src, fixed_len=L)
    for chunk in chunks:
        for threads in thread_list:
            os.environ["AWFM_GPU_LOCATE_CHUNK_QUERIES"] = str(chunk)
            lib.awFmGpuReleaseIndex(ip)
            assert lib.awFmParallelSearchLocate(ip, sl.ptr, threads) == abi.AwFmSuccess
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                rc = lib.awFmParallelSearchLocate(ip, sl.ptr, threads)
                best = min(best, time.perf_counter() - t0)
            hits = int(sl.entries()["count"][:n].sum(dtype=np.uint64))
            print("RESULT " + json.dumps({"source": src_name, "chunk": chunk, "threads": threads, "ms": round(best * 1e3, 2),
                                          "Mq_per_s": round(n / best / 1e6, 1), "Mhits_per_s": round(hits / best / 1e6, 1),
                                          "hits": hits, "sa_ratio": ratio}), flush=True)
