"""Development sweep of the drop-in count path (host marshalling): chunk size x pinned/pageable source."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avxwindowfmindex_b200 import DeviceBuiltIndex, KmerSearchList, abi, capi, synth
lib = capi.load()
bp, n, L = int(os.environ.get("BP", 3_100_000_000)), int(os.environ.get("NQ", 100_000_000)), 20
d_text = torch.empty(bp, dtype=torch.uint8, device="cuda")
capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), bp, synth.TEXT_SEED + 2, 0, 0))
built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), bp, abi.AwFmAlphabetDna, 12, 8)
del d_text
arrays = built.to_host(); built.close()
ix = arrays.as_awfm_index(); ip = C.addressof(ix)
d_q = torch.empty(n * L, dtype=torch.uint8, device="cuda")
capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), n * L, synth.QUERY_SEED + 2, 0, 0))
pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory(); pinned.copy_(d_q)
pageable = pinned.numpy().copy()
del d_q
sl = KmerSearchList(lib, n)
os.environ["AWFM_GPU_VERBOSE"] = "1"
threads = os.cpu_count()
chunks = [int(x) for x in os.environ.get("CHUNKS", "32768,65536,131072,262144,1048576").split(",")]
sources = (("pinned", pinned.numpy()), ("pageable", pageable)) if not os.environ.get("PINNED_ONLY") else (("pinned", pinned.numpy()),)
for src_name, src in sources:
    sl.fill(src, fixed_len=L)
    counts = sl.entries()["count"]
    for chunk in chunks:
        os.environ["AWFM_GPU_CHUNK_QUERIES"] = str(chunk)
        lib.awFmGpuReleaseIndex(ip)
        lib.awFmParallelSearchCount(ip, sl.ptr, threads)
        # fresh: count == 0 everywhere before the call (awFmCreateKmerSearchList's state); stale: every count has to be rewritten
        for state, value in (("fresh", 0), ("stale", 0xFFFFFFFF)):
            best = 1e9
            for _ in range(3):
                counts[:n] = value
                t0 = time.perf_counter(); lib.awFmParallelSearchCount(ip, sl.ptr, threads); best = min(best, time.perf_counter() - t0)
            print(f"RESULT {src_name} chunk={chunk} list={state} best={best*1e3:.1f} ms -> {n/best/1e6:.0f} Mq/s", flush=True)
