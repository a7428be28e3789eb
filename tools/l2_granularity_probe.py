import ctypes as C, sys
sys.path.insert(0,'.')
from avxwindowfmindex_b200 import capi
lib=capi.load()
act=C.c_int(); gb=C.c_double()
lib.awfm_gpu_set_l2_fetch_granularity(0,0,C.byref(act)); print("default granularity", act.value)
for g in (128,64,32):
    capi.check(lib.awfm_gpu_set_l2_fetch_granularity(0,g,C.byref(act)))
    for rec,lanes in ((128,4),(64,4),(32,2),(16,1)):
        capi.check(lib.awfm_gpu_gather_bandwidth(0, 2048<<20, rec, 1<<27, lanes, C.byref(gb)))
        print("gran",act.value,"rec",rec,"lanes",lanes,"GB/s",round(gb.value,1),"Greads/s",round(gb.value/rec,2), flush=True)
