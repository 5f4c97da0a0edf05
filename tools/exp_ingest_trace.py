import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os, time, numpy as np, torch, json
import yams_b200 as Y
assert Y.plugin_init()==0
n = 8<<30
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
Y.synth_bytes_device(12345, 0, n, buf.data_ptr())
for seg in ("1024","4096","256"):
    os.environ["YAMS_B200_SEGMENT_MIB"]=seg
    for i in range(3):
        t0=time.perf_counter(); ch = Y.chunk_and_hash_device(buf.data_ptr(), n, Y.default_config()); dt=time.perf_counter()-t0
    print("seg",seg,"wall %.1f ms"%(dt*1e3), json.dumps(Y.ingest_last_timings()))
