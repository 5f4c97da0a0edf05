"""Seeded byte streams shared by the oracle pin tests and the GPU parity tests."""
import numpy as np


def pattern_bytes(n):
    # /root/reference/tests/unit/chunking/chunking_test.cpp:55-62
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(1315423911) + np.uint64(0x9E3779B9)) & np.uint64(0xFF)).astype(np.uint8)


def make_stream(name, O):
    if name == "splitmix_12345_8MiB":
        return O.gen_bytes(12345, 0, 8 << 20)
    if name == "pattern_4MiB":
        return pattern_bytes(4 << 20)
    if name == "const42_3MiB_plus7":
        return np.full((3 << 20) + 7, 0x42, dtype=np.uint8)
    if name == "splitmix_7_100000":
        return O.gen_bytes(7, 0, 100000)
    if name == "empty":
        return np.zeros(0, dtype=np.uint8)
    if name == "one_byte":
        return np.array([0xC5], dtype=np.uint8)
    raise KeyError(name)


def cfg_from_dict(O, d):
    return O.CdcConfig(d["window_size"], d["min_chunk"], d["max_chunk"], d["polynomial"], d["mask"],
                       d["variant"])
