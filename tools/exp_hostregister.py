"""Pageable-source upload: cudaHostRegister + direct DMA vs the staged copy (host threads -> pinned ring -> H2D)."""
import ctypes as C
import time

import numpy as np
import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

torch.cuda.init()
for gib in (1, 4):
    n = gib << 30
    src = np.ones(n, dtype=np.uint8)      # pageable, touched
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for slice_mib in (64, 256, n >> 20):
        sl = slice_mib << 20
        t0 = time.perf_counter()
        t_reg = 0.0
        for lo in range(0, n, sl):
            a = src.ctypes.data + lo
            t1 = time.perf_counter()
            rc = rt.cudaHostRegister(a, min(sl, n - lo), 0)
            t_reg += time.perf_counter() - t1
            assert rc == 0, rc
            rt.cudaMemcpyAsync(dst.data_ptr() + lo, a, min(sl, n - lo), 1, None)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for lo in range(0, n, sl):
            rt.cudaHostUnregister(src.ctypes.data + lo)
        t3 = time.perf_counter()
        print(f"{gib} GiB slices of {slice_mib} MiB: register+copy {n / (t2 - t0) / 1e9:.1f} GB/s (register alone {t_reg * 1e3:.0f} ms), "
              f"with unregister {n / (t3 - t0) / 1e9:.1f} GB/s", flush=True)
    t0 = time.perf_counter()
    dst.copy_(torch.from_numpy(src))
    torch.cuda.synchronize()
    print(f"{gib} GiB plain pageable cudaMemcpy {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
