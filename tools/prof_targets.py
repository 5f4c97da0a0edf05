"""Small targets for ncu captures (tools/profile_r2.sh): `sha` = SHA-256 over the chunk table of a 4 GiB resident segment,
`pq` = the SimeonPqAdc scan over 4 M x 768 rows, `knn` = two C2 batches through search_device."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import yams_b200 as Y  # noqa: E402

what = sys.argv[1]
assert Y.plugin_init() == 0, Y.health()
if what == "sha":
    n = 4 << 30
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    Y.synth_bytes_device(12345, 0, n, buf.data_ptr())
    for _ in range(2):
        ch = Y.chunk_and_hash_device(buf.data_ptr(), n, Y.default_config())
    print("chunks", len(ch), Y.ingest_last_timings())
elif what == "pq":
    n, d = 4_000_000, 768
    c = Y.Corpus(d, Y.F16, Y.COSINE, capacity_hint=n)
    for r0 in range(0, n, 1_000_000):
        c.append_synthetic(42, r0, 1_000_000)
    cb = (np.random.default_rng(1).normal(size=(32, 256, d // 32)) / np.sqrt(d)).astype(np.float32)
    pq = Y.PqIndex(c, 32, 256, cb)
    q = np.random.default_rng(2).normal(size=(4, d)).astype(np.float32)
    for _ in range(2):
        pq.search(q, 10)
else:
    n, d, nq = 10_000_000, 768, 1024
    c = Y.Corpus(d, Y.F16, Y.COSINE, capacity_hint=n)
    for r0 in range(0, n, 1_000_000):
        c.append_synthetic(42, r0, 1_000_000)
    q = torch.empty((nq, d), dtype=torch.float32, device="cuda")
    Y.synth_rows_device(43, 0, nq, d, q.data_ptr())
    out = torch.empty(nq * 10 * 12, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        c.search_device(q.data_ptr(), nq, 10, -1.0, out.data_ptr(), out.data_ptr() + nq * 80)
    c.sync()
    print(c.last_timings())
