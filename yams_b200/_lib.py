"""ctypes binding of include/yams_b200.h (every exported symbol is bound here; tests check that)."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libyams_b200.so")

STREAMING, RABIN = 0, 1
F32, F16 = 0, 1
COSINE, L2 = 0, 1
FLAG_TIE_AT_K, FLAG_FALLBACK_PATH = 1, 2

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)


class YamsB200Error(RuntimeError):
    def __init__(self, status: int, where: str, text: str):
        super().__init__(f"{where}: status {status}: {text}")
        self.status = status


class CdcConfig(C.Structure):
    """yams_cdc_config == ChunkingConfig (/root/reference/include/yams/chunking/chunker.h:44-51)."""
    _fields_ = [("window_size", C.c_uint64), ("min_chunk", C.c_uint64), ("max_chunk", C.c_uint64),
                ("polynomial", C.c_uint64), ("mask", C.c_uint64), ("variant", C.c_int32),
                ("reserved", C.c_int32)]


class ChunkDesc(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("size", C.c_uint64), ("digest", C.c_uint8 * 32)]


class ManifestSummary(C.Structure):
    _fields_ = [("checksum", C.c_uint32), ("valid", C.c_uint32), ("offsets_sequential", C.c_uint32), ("sizes_valid", C.c_uint32),
                ("chunk_count", C.c_uint64), ("total_size", C.c_uint64), ("checksum_text_bytes", C.c_uint64)]


class SimeonConfig(C.Structure):
    _fields_ = [("ngram_min", C.c_uint32), ("ngram_max", C.c_uint32), ("sketch_dim", C.c_uint32), ("output_dim", C.c_uint32),
                ("hash_seed", C.c_uint64), ("projection_seed", C.c_uint64), ("l2_normalize", C.c_int32), ("flags", C.c_int32)]


SIMEON_WORD_TOKENS = 1        # NGramMode::CharAndWord
SIMEON_PROJECTION_FWHT = 2    # ProjectionMode::Fwht (default projection: AchlioptasSparse)

CHUNK_REF_DTYPE = np.dtype([("hash", "S64"), ("offset", "<u8"), ("size", "<u4"), ("flags", "<u4")])
assert CHUNK_REF_DTYPE.itemsize == 80
CHUNK_DTYPE = np.dtype([("offset", "<u8"), ("size", "<u8"), ("digest", "u1", (32,))])
assert CHUNK_DTYPE.itemsize == C.sizeof(ChunkDesc) == 48

# every symbol include/yams_b200.h declares: name -> (restype, argtypes)
_descpp = C.POINTER(C.POINTER(ChunkDesc))
_szp = C.POINTER(C.c_size_t)
SYMBOLS = {
    "yams_plugin_get_abi_version": (C.c_int, []),
    "yams_plugin_get_name": (C.c_char_p, []),
    "yams_plugin_get_version": (C.c_char_p, []),
    "yams_plugin_get_manifest_json": (C.c_char_p, []),
    "yams_plugin_init": (C.c_int, [C.c_char_p, C.c_void_p]),
    "yams_plugin_shutdown": (None, []),
    "yams_plugin_get_interface": (C.c_int, [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "yams_plugin_get_health_json": (C.c_int, [C.POINTER(C.c_void_p)]),
    "yams_b200_cdc_default_config": (None, [C.POINTER(CdcConfig)]),
    "yams_b200_chunk_and_hash": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(CdcConfig), _descpp, _szp]),
    "yams_b200_free_chunks": (None, [C.c_void_p, C.POINTER(ChunkDesc), C.c_size_t]),
    "yams_b200_chunk_and_hash_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(CdcConfig), _descpp, _szp]),
    "yams_b200_ingest_open": (C.c_int, [C.c_void_p, C.POINTER(CdcConfig), C.POINTER(C.c_void_p)]),
    "yams_b200_ingest_feed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, _descpp, _szp]),
    "yams_b200_ingest_finish": (C.c_int, [C.c_void_p, _descpp, _szp]),
    "yams_b200_ingest_close": (None, [C.c_void_p]),
    "yams_b200_sha256_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, u64p, u64p, C.c_size_t, u8p]),
    "yams_b200_sha256_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, u64p, u64p, C.c_size_t, u8p]),
    "yams_b200_chunk_boundaries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(CdcConfig), _descpp, _szp]),
    "yams_b200_chunk_and_hash_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t,
                                                 C.POINTER(CdcConfig), C.POINTER(C.POINTER(ChunkDesc)), C.POINTER(C.c_size_t), u64p]),
    "yams_b200_sha256_many": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p]),
    "yams_b200_digest_set_create": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "yams_b200_digest_set_insert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, u64p]),
    "yams_b200_digest_set_contains": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "yams_b200_digest_set_size": (C.c_int, [C.c_void_p, u64p]),
    "yams_b200_digest_set_last_ms": (C.c_int, [C.c_void_p, f32p]),
    "yams_b200_digest_set_destroy": (None, [C.c_void_p]),
    "yams_b200_manifest_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(ManifestSummary)]),
    "yams_b200_ingest_last_timings": (C.c_int, [C.c_void_p, f32p]),
    "yams_b200_dedup_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "yams_b200_corpus_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "yams_b200_corpus_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, i64p]),
    "yams_b200_corpus_append_f32_as_f16": (C.c_int, [C.c_void_p, f32p, C.c_uint64, i64p]),
    "yams_b200_corpus_append_synthetic": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "yams_b200_corpus_remove": (C.c_int, [C.c_void_p, i64p, C.c_uint64, u64p]),
    "yams_b200_corpus_clear": (C.c_int, [C.c_void_p]),
    "yams_b200_batch_distance": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_uint32, f32p, C.c_uint64, C.c_int, C.c_uint64,
                                           C.c_float, u64p, f32p, u64p]),
    "yams_b200_compute_cosine_similarity": (C.c_int, [C.c_void_p, f32p, C.c_size_t, f32p, C.c_size_t,
                                                      C.POINTER(C.c_double)]),
    "yams_b200_compute_cosine_similarity_many": (C.c_int, [C.c_void_p, f32p, f32p, C.c_size_t, C.c_size_t, C.POINTER(C.c_double)]),
    "yams_b200_corpus_size": (C.c_int, [C.c_void_p, u64p]),
    "yams_b200_corpus_destroy": (None, [C.c_void_p]),
    "yams_b200_search": (C.c_int, [C.c_void_p, f32p, C.c_uint32, C.c_uint32, C.c_float, i64p, u64p, i64p, f32p, u32p, u64p]),
    "yams_b200_search_all_matching": (C.c_int, [C.c_void_p, f32p, C.c_float, i64p, C.c_uint64, i64p, f32p, u64p]),
    "yams_b200_search_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]),
    "yams_b200_search_device_finish": (C.c_int, [C.c_void_p, u32p]),
    "yams_b200_search_exhaustive": (C.c_int, [C.c_void_p, f32p, C.c_uint32, C.c_uint32, C.c_float, i64p, f32p, u32p, u64p]),
    "yams_b200_merge_packed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "yams_b200_merge_partials_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "yams_b200_corpus_sync": (C.c_int, [C.c_void_p]),
    "yams_b200_corpus_stream": (C.c_void_p, [C.c_void_p]),
    "yams_b200_vec0_exact": (C.c_int, [C.c_void_p, f32p, C.c_uint32, f32p, i64p, C.c_uint64, C.c_uint64, C.c_int,
                                       C.c_int64, C.c_int64, i64p, f32p, u64p]),
    "yams_b200_search_last_timings": (C.c_int, [C.c_void_p, f32p]),
    "sqlite3_vec_distance_l2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, f32p]),
    "sqlite3_vec_distance_cosine": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, f32p]),
    "yams_b200_vec_distance_l1": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, f32p]),
    "yams_b200_synth_bytes_device": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "yams_b200_debug_stage1_scores": (C.c_int, [C.c_void_p, f32p, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, f32p]),
    "yams_b200_synth_rows_device": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "yams_b200_simeon_default_config": (None, [C.POINTER(SimeonConfig)]),
    "yams_b200_simeon_yams_config": (None, [C.POINTER(SimeonConfig), C.c_uint32]),
    "yams_b200_simeon_create": (C.c_int, [C.c_void_p, C.POINTER(SimeonConfig), C.POINTER(C.c_void_p)]),
    "yams_b200_simeon_encode": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, f32p]),
    "yams_b200_simeon_destroy": (None, [C.c_void_p]),
    "yams_b200_pq_build": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, f32p, u64p, C.POINTER(C.c_void_p)]),
    "yams_b200_pq_search": (C.c_int, [C.c_void_p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, i64p, f32p, u32p, u64p]),
    "yams_b200_pq_codes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u64p]),
    "yams_b200_pq_destroy": (None, [C.c_void_p]),
    "yams_b200_debug_last_eps": (C.c_int, [C.c_void_p, C.c_uint32, f32p]),
    "yams_b200_device_count": (C.c_int, []),
    "yams_b200_last_error": (C.c_char_p, []),
}

_lib = None


def lib_path() -> str:
    return _SO


def lib():
    """The loaded shared library. Fails loudly when it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(
                f"{_SO} is missing: build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
                "(nvcc, sm_100a). yams_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(_SO)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int, where: str):
    if rc != 0:
        raise YamsB200Error(rc, where, (lib().yams_b200_last_error() or b"").decode(errors="replace"))


def device_count() -> int:
    return lib().yams_b200_device_count()


def plugin_init(config: Optional[dict] = None) -> int:
    return lib().yams_plugin_init(json.dumps(config or {}).encode(), None)


def health() -> dict:
    p = C.c_void_p()
    lib().yams_plugin_get_health_json(C.byref(p))
    s = C.cast(p, C.c_char_p).value.decode()
    C.CDLL(None).free(p)
    return json.loads(s)


def default_config(variant: int = STREAMING, **kw) -> CdcConfig:
    cfg = CdcConfig()
    lib().yams_b200_cdc_default_config(C.byref(cfg))
    cfg.variant = variant
    for k, v in kw.items():
        if k not in {f[0] for f in CdcConfig._fields_}:
            raise TypeError(f"unknown CdcConfig field {k!r}")
        setattr(cfg, k, v)
    return cfg


def lib_chunk_dtype():
    """numpy dtype of yams_chunk_desc (offset, size, digest[32])."""
    return CHUNK_DTYPE


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _take(out_p, out_n) -> np.ndarray:
    n = out_n.value
    if n == 0:
        return np.zeros(0, dtype=CHUNK_DTYPE)
    buf = (ChunkDesc * n).from_address(C.addressof(out_p.contents))
    arr = np.frombuffer(buf, dtype=CHUNK_DTYPE).copy()
    lib().yams_b200_free_chunks(None, out_p, n)
    return arr


def _chunks(fn_name: str, ptr, length: int, cfg: CdcConfig) -> np.ndarray:
    out_p = C.POINTER(ChunkDesc)()
    out_n = C.c_size_t(0)
    rc = getattr(lib(), fn_name)(None, ptr, length, C.byref(cfg), C.byref(out_p), C.byref(out_n))
    _check(rc, fn_name)
    return _take(out_p, out_n)


def chunk_and_hash(data, cfg: Optional[CdcConfig] = None) -> np.ndarray:
    """IChunker::chunkDataLazy over a host buffer -> structured array (offset, size, digest[32])."""
    a = _as_u8(data)
    return _chunks("yams_b200_chunk_and_hash", a.ctypes.data if a.size else None, a.size, cfg or default_config())


def chunk_and_hash_batch(files, cfg: Optional[CdcConfig] = None):
    """Many files in one call (`yams add -r`): -> list of per-file chunk tables (offsets relative to each file)."""
    arrs = [_as_u8(f) for f in files]
    n = len(arrs)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (C.c_size_t * max(n, 1))(*[a.size for a in arrs])
    first = np.zeros(n + 1, dtype=np.uint64)
    out_p = C.POINTER(ChunkDesc)()
    out_n = C.c_size_t(0)
    c = cfg or default_config()
    rc = lib().yams_b200_chunk_and_hash_batch(None, ptrs, lens, n, C.byref(c), C.byref(out_p), C.byref(out_n),
                                              first.ctypes.data_as(u64p))
    _check(rc, "chunk_and_hash_batch")
    table = _take(out_p, out_n)
    return [table[int(first[i]):int(first[i + 1])] for i in range(n)]


def sha256_many(messages) -> np.ndarray:
    """SHA-256 of separate host buffers in one device pass -> uint8[n, 32]."""
    arrs = [_as_u8(m) for m in messages]
    n = len(arrs)
    out = np.zeros((n, 32), dtype=np.uint8)
    if n:
        ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
        lens = (C.c_size_t * n)(*[a.size for a in arrs])
        _check(lib().yams_b200_sha256_many(None, ptrs, lens, n, out.ctypes.data), "sha256_many")
    return out


def chunk_boundaries(data, cfg: Optional[CdcConfig] = None) -> np.ndarray:
    a = _as_u8(data)
    return _chunks("yams_b200_chunk_boundaries", a.ctypes.data if a.size else None, a.size, cfg or default_config())


def chunk_and_hash_device(dev_ptr: int, length: int, cfg: Optional[CdcConfig] = None) -> np.ndarray:
    """Input already resident in HBM (e.g. ``tensor.data_ptr()``)."""
    return _chunks("yams_b200_chunk_and_hash_device", dev_ptr, length, cfg or default_config())


def dedup_stats(chunks: np.ndarray) -> dict:
    """calculateDeduplication over a chunk table (structured array from chunk_and_hash)."""
    arr = np.ascontiguousarray(chunks, dtype=CHUNK_DTYPE)
    out = (C.c_uint64 * 4)()
    _check(lib().yams_b200_dedup_stats(None, arr.ctypes.data if len(arr) else None, len(arr), out), "dedup_stats")
    total, unique, count, ucount = (int(x) for x in out)
    return {"totalSize": total, "uniqueSize": unique, "chunkCount": count, "uniqueChunks": ucount,
            "ratio": (1.0 - unique / total) if total else 0.0}


def manifest_build(chunks: np.ndarray, file_digest, file_size: int, want_refs: bool = True):
    """ManifestManager::createManifest over a chunk table: -> (ChunkRef table or None, summary dict incl. the CRC checksum)."""
    arr = np.ascontiguousarray(chunks, dtype=CHUNK_DTYPE)
    dg = np.ascontiguousarray(np.frombuffer(bytes(file_digest), dtype=np.uint8))
    assert dg.size == 32
    refs = np.zeros(max(len(arr), 1), dtype=CHUNK_REF_DTYPE) if want_refs else None
    out = ManifestSummary()
    _check(lib().yams_b200_manifest_build(None, arr.ctypes.data if len(arr) else None, len(arr), dg.ctypes.data, file_size,
                                          refs.ctypes.data if want_refs else None, C.byref(out)), "manifest_build")
    summary = {f[0]: int(getattr(out, f[0])) for f in ManifestSummary._fields_}
    return (refs[:len(arr)] if want_refs else None), summary


def ingest_last_timings() -> dict:
    ms = (C.c_float * 8)()
    lib().yams_b200_ingest_last_timings(None, ms)
    return {"scan_ms": ms[0], "select_ms": ms[1], "sha256_ms": ms[2], "total_ms": ms[3],
            "host_sync1_ms": ms[4], "host_sync2_ms": ms[5], "host_alloc_ms": ms[6], "host_process_ms": ms[7]}


class IngestSession:
    """StreamingChunker::processStream shaped session (open / feed... / finish)."""

    def __init__(self, cfg: Optional[CdcConfig] = None):
        self._h = C.c_void_p()
        self._cfg = cfg or default_config()
        _check(lib().yams_b200_ingest_open(None, C.byref(self._cfg), C.byref(self._h)), "ingest_open")

    def feed(self, data) -> np.ndarray:
        a = _as_u8(data)
        out_p = C.POINTER(ChunkDesc)()
        out_n = C.c_size_t(0)
        _check(lib().yams_b200_ingest_feed(self._h, a.ctypes.data if a.size else None, a.size, C.byref(out_p),
                                           C.byref(out_n)), "ingest_feed")
        return _take(out_p, out_n)

    def finish(self) -> np.ndarray:
        out_p = C.POINTER(ChunkDesc)()
        out_n = C.c_size_t(0)
        _check(lib().yams_b200_ingest_finish(self._h, C.byref(out_p), C.byref(out_n)), "ingest_finish")
        return _take(out_p, out_n)

    def close(self):
        if self._h:
            lib().yams_b200_ingest_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sha256_batch(base, offsets, sizes) -> np.ndarray:
    a = _as_u8(base)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    szs = np.ascontiguousarray(sizes, dtype=np.uint64)
    out = np.zeros((len(offs), 32), dtype=np.uint8)
    _check(lib().yams_b200_sha256_batch(None, a.ctypes.data if a.size else None, a.size, offs.ctypes.data_as(u64p),
                                        szs.ctypes.data_as(u64p), len(offs), out.ctypes.data_as(u8p)), "sha256_batch")
    return out


def sha256_batch_device(dev_ptr: int, base_len: int, offsets, sizes) -> np.ndarray:
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    szs = np.ascontiguousarray(sizes, dtype=np.uint64)
    out = np.zeros((len(offs), 32), dtype=np.uint8)
    _check(lib().yams_b200_sha256_batch_device(None, dev_ptr, base_len, offs.ctypes.data_as(u64p), szs.ctypes.data_as(u64p),
                                               len(offs), out.ctypes.data_as(u8p)), "sha256_batch_device")
    return out


def synth_bytes_device(seed: int, start: int, n: int, dev_ptr: int):
    _check(lib().yams_b200_synth_bytes_device(seed, start, n, dev_ptr), "synth_bytes_device")


def synth_rows_device(seed: int, first_row: int, n: int, dim: int, dev_ptr: int):
    """n x dim fp32 rows of the synthetic vector generator into a device buffer."""
    _check(lib().yams_b200_synth_rows_device(seed, first_row, n, dim, dev_ptr), "synth_rows_device")


class Corpus:
    """Device-resident mirror of the `vectors` table (row-major [n, dim], fp32 or fp16)."""

    def __init__(self, dim: int, dtype: int = F16, metric: int = COSINE, capacity_hint: int = 0):
        self._h = C.c_void_p()
        self.dim, self.dtype, self.metric = dim, dtype, metric
        _check(lib().yams_b200_corpus_create(None, dim, dtype, metric, capacity_hint, C.byref(self._h)), "corpus_create")

    @property
    def handle(self):
        return self._h

    def append(self, rows: np.ndarray, rowids=None):
        rows = np.ascontiguousarray(rows)
        n = rows.shape[0]
        rid = np.ascontiguousarray(rowids, dtype=np.int64) if rowids is not None else None
        ridp = rid.ctypes.data_as(i64p) if rid is not None else None
        if self.dtype == F16 and rows.dtype == np.float32:
            _check(lib().yams_b200_corpus_append_f32_as_f16(self._h, rows.ctypes.data_as(f32p), n, ridp), "corpus_append_f32_as_f16")
        else:
            want = np.uint16 if self.dtype == F16 else np.float32
            if self.dtype == F16 and rows.dtype == np.float16:
                rows = rows.view(np.uint16)
            assert rows.dtype == want, (rows.dtype, want)
            _check(lib().yams_b200_corpus_append(self._h, rows.ctypes.data, n, ridp), "corpus_append")

    def append_synthetic(self, seed: int, first_row: int, n: int):
        _check(lib().yams_b200_corpus_append_synthetic(self._h, seed, first_row, n), "corpus_append_synthetic")

    def clear(self):
        _check(lib().yams_b200_corpus_clear(self._h), "corpus_clear")

    def remove(self, rowids) -> int:
        """deleteVector / deleteVectorsByDocument mirror: drop rows by rowid, keep the rest in rowid order."""
        r = np.ascontiguousarray(rowids, dtype=np.int64).reshape(-1)
        out = C.c_uint64(0)
        _check(lib().yams_b200_corpus_remove(self._h, r.ctypes.data_as(i64p) if len(r) else None, len(r), C.byref(out)),
               "corpus_remove")
        return out.value

    def __len__(self):
        n = C.c_uint64(0)
        _check(lib().yams_b200_corpus_size(self._h, C.byref(n)), "corpus_size")
        return n.value

    def search(self, queries: np.ndarray, k: int, threshold: float = -1.0, allowed=None):
        """-> (rowids i64[Q,k], scores f32[Q,k], counts u32[Q], flags u64[Q]).
        allowed: optional list (len Q) of ascending rowid arrays (CandidateFilterMode::Exact)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        out_r = np.full((nq, max(k, 1)), -1, dtype=np.int64)
        out_s = np.zeros((nq, max(k, 1)), dtype=np.float32)
        out_c = np.zeros(nq, dtype=np.uint32)
        out_f = np.zeros(nq, dtype=np.uint64)
        al_p = of_p = None
        if allowed is not None:
            offs = np.zeros(nq + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(a) for a in allowed])
            al = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int64) for a in allowed])
                                      if nq and int(offs[-1]) else np.zeros(1, dtype=np.int64), dtype=np.int64)
            al_p, of_p = al.ctypes.data_as(i64p), offs.ctypes.data_as(u64p)
        rc = lib().yams_b200_search(self._h, q.ctypes.data_as(f32p), nq, k, threshold, al_p, of_p,
                                    out_r.ctypes.data_as(i64p), out_s.ctypes.data_as(f32p), out_c.ctypes.data_as(u32p),
                                    out_f.ctypes.data_as(u64p))
        _check(rc, "search")
        return out_r[:, :k], out_s[:, :k], out_c, out_f

    def search_all_matching(self, query, threshold: float = -1.0, allowed=None):
        """ExactRowSelection::AllMatching: every passing row of the candidate set, sorted (sim desc, rowid asc)."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        al = np.ascontiguousarray(allowed, dtype=np.int64) if allowed is not None else None
        m = len(al) if al is not None else len(self)
        out_r = np.empty(max(m, 1), dtype=np.int64)
        out_s = np.empty(max(m, 1), dtype=np.float32)
        cnt = C.c_uint64(0)
        rc = lib().yams_b200_search_all_matching(self._h, q.ctypes.data_as(f32p), threshold,
                                                 al.ctypes.data_as(i64p) if al is not None and len(al) else None,
                                                 len(al) if al is not None else 0, out_r.ctypes.data_as(i64p),
                                                 out_s.ctypes.data_as(f32p), C.byref(cnt))
        _check(rc, "search_all_matching")
        return out_r[:cnt.value].copy(), out_s[:cnt.value].copy()

    def search_device(self, d_queries: int, nq: int, k: int, threshold: float, d_out_rowids: int, d_out_scores: int):
        _check(lib().yams_b200_search_device(self._h, d_queries, nq, k, threshold, d_out_rowids, d_out_scores), "search_device")

    def search_device_finish(self) -> int:
        """Blocks until the last search_device is done; raises on invalid queries; -> queries resolved exhaustively."""
        n = C.c_uint32(0)
        _check(lib().yams_b200_search_device_finish(self._h, C.byref(n)), "search_device_finish")
        return n.value

    def search_exhaustive(self, queries: np.ndarray, k: int, threshold: float = -1.0):
        """The library's own exhaustive reference (exact score of every row + global sort); same outputs as search."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        out_r = np.full((nq, k), -1, dtype=np.int64)
        out_s = np.zeros((nq, k), dtype=np.float32)
        out_c = np.zeros(nq, dtype=np.uint32)
        out_f = np.zeros(nq, dtype=np.uint64)
        _check(lib().yams_b200_search_exhaustive(self._h, q.ctypes.data_as(f32p), nq, k, threshold, out_r.ctypes.data_as(i64p),
                                                 out_s.ctypes.data_as(f32p), out_c.ctypes.data_as(u32p), out_f.ctypes.data_as(u64p)),
               "search_exhaustive")
        return out_r, out_s, out_c, out_f

    def merge_packed_device(self, d_packed: int, nranks: int, nq: int, k: int, d_out_rowids: int, d_out_scores: int,
                            d_out_counts: int = 0, stream: int = 0):
        _check(lib().yams_b200_merge_packed_device(self._h, d_packed, nranks, nq, k, d_out_rowids, d_out_scores,
                                                   d_out_counts or None, stream or None), "merge_packed_device")

    def merge_partials_device(self, d_rowids: int, d_scores: int, nranks: int, nq: int, k: int, d_out_rowids: int,
                              d_out_scores: int, d_out_counts: int = 0):
        _check(lib().yams_b200_merge_partials_device(self._h, d_rowids, d_scores, nranks, nq, k, d_out_rowids, d_out_scores,
                                                     d_out_counts or None), "merge_partials_device")

    def debug_stage1_scores(self, queries, engine: int, row_start: int = 0, row_stride: int = 1, nrows: int = 0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nrows = nrows or len(self)
        out = np.empty((q.shape[0], nrows), dtype=np.float32)
        _check(lib().yams_b200_debug_stage1_scores(self._h, q.ctypes.data_as(f32p), q.shape[0], engine, row_start, row_stride,
                                                   nrows, out.ctypes.data_as(f32p)), "debug_stage1_scores")
        return out

    def debug_last_eps(self, nq: int) -> np.ndarray:
        out = np.empty(nq, dtype=np.float32)
        _check(lib().yams_b200_debug_last_eps(self._h, nq, out.ctypes.data_as(f32p)), "debug_last_eps")
        return out

    def sync(self):
        _check(lib().yams_b200_corpus_sync(self._h), "corpus_sync")

    @property
    def stream(self) -> int:
        return lib().yams_b200_corpus_stream(self._h) or 0

    def last_timings(self) -> dict:
        ms = (C.c_float * 8)()
        lib().yams_b200_search_last_timings(self._h, ms)
        return {"stage1_ms": ms[0], "stage2_ms": ms[1], "total_ms": ms[2], "scan_kernel_ms": ms[5],
                "engine": "tcgen05" if ms[4] else "cuda-core", "resolved_exhaustively": int(ms[6])}

    def close(self):
        if self._h:
            lib().yams_b200_corpus_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SimeonEncoder:
    """The Simeon text encoder: `profile="simeon-v1-384"` (byte n-grams -> count sketch -> Achlioptas projection -> L2) or
    `profile="yams-default"` (what an unconfigured YAMS runs: byte n-grams + word tokens -> count sketch -> FWHT -> L2,
    output_dim = embedding_dim)."""

    def __init__(self, profile: str = "simeon-v1-384", embedding_dim: int = 0, **overrides):
        cfg = SimeonConfig()
        if profile == "yams-default":
            lib().yams_b200_simeon_yams_config(C.byref(cfg), embedding_dim)
        else:
            assert profile == "simeon-v1-384", profile
            lib().yams_b200_simeon_default_config(C.byref(cfg))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self._h = C.c_void_p()
        _check(lib().yams_b200_simeon_create(None, C.byref(cfg), C.byref(self._h)), "simeon_create")

    def encode(self, texts) -> np.ndarray:
        raw = [t.encode("utf-8") if isinstance(t, str) else bytes(t) for t in texts]
        n = len(raw)
        out = np.zeros((n, self.cfg.output_dim), dtype=np.float32)
        if n:
            ptrs = (C.c_char_p * n)(*raw)
            lens = (C.c_size_t * n)(*[len(r) for r in raw])
            _check(lib().yams_b200_simeon_encode(self._h, ptrs, lens, n, out.ctypes.data_as(f32p)), "simeon_encode")
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().yams_b200_simeon_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PqIndex:
    """SimeonPqAdc index over a Corpus (codes in HBM beside the rows); search = ADC scan + exact rerank."""

    def __init__(self, corpus: Corpus, m: int, k: int, codebooks: np.ndarray, tie_break_keys=None):
        cb = np.ascontiguousarray(codebooks, dtype=np.float32)
        assert cb.size == m * k * (corpus.dim // m)
        tk = np.ascontiguousarray(tie_break_keys, dtype=np.uint64) if tie_break_keys is not None else None
        self._h = C.c_void_p()
        self.m, self.k, self._corpus = m, k, corpus
        _check(lib().yams_b200_pq_build(corpus.handle, m, k, cb.ctypes.data_as(f32p), tk.ctypes.data_as(u64p) if tk is not None else None,
                                        C.byref(self._h)), "pq_build")

    def codes(self):
        n = C.c_uint64(0)
        _check(lib().yams_b200_pq_codes(self._h, None, None, C.byref(n)), "pq_codes")
        codes = np.zeros((max(n.value, 1), self.m), dtype=np.uint8)
        rowids = np.zeros(max(n.value, 1), dtype=np.int64)
        _check(lib().yams_b200_pq_codes(self._h, codes.ctypes.data, rowids.ctypes.data, C.byref(n)), "pq_codes")
        return codes[:n.value], rowids[:n.value]

    def search(self, queries, k: int, rerank_factor: int = 2, threshold: float = 0.0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        out_r = np.full((nq, max(k, 1)), -1, dtype=np.int64)
        out_s = np.zeros((nq, max(k, 1)), dtype=np.float32)
        out_c = np.zeros(nq, dtype=np.uint32)
        out_f = np.zeros(nq, dtype=np.uint64)
        _check(lib().yams_b200_pq_search(self._h, q.ctypes.data_as(f32p), nq, k, rerank_factor, threshold, out_r.ctypes.data_as(i64p),
                                         out_s.ctypes.data_as(f32p), out_c.ctypes.data_as(u32p), out_f.ctypes.data_as(u64p)), "pq_search")
        return out_r[:, :k], out_s[:, :k], out_c, out_f

    def close(self):
        if getattr(self, "_h", None):
            lib().yams_b200_pq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def vec0_exact(query, rows, k: int = 0, rowids=None, rowid_range=None):
    """vec0_run_exact_query: (rowids, distances) ascending by L2 distance."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(query, dtype=np.float32)
    n, d = rows.shape if rows.ndim == 2 else (0, q.size)
    out_r = np.zeros(max(n, 1), dtype=np.int64)
    out_d = np.zeros(max(n, 1), dtype=np.float32)
    cnt = C.c_uint64(0)
    rid = np.ascontiguousarray(rowids, dtype=np.int64) if rowids is not None else None
    lo, hi = rowid_range if rowid_range is not None else (0, 0)
    rc = lib().yams_b200_vec0_exact(None, q.ctypes.data_as(f32p), d, rows.ctypes.data_as(f32p),
                                    rid.ctypes.data_as(i64p) if rid is not None else None, n, k,
                                    int(rowid_range is not None), lo, hi, out_r.ctypes.data_as(i64p),
                                    out_d.ctypes.data_as(f32p), C.byref(cnt))
    _check(rc, "vec0_exact")
    return out_r[:cnt.value].copy(), out_d[:cnt.value].copy()


class DigestSet:
    """Device-resident set of chunk digests: the batched `storage_->exists` / `store` loop
    (content_store_impl.cpp:245-288)."""

    def __init__(self, capacity_hint: int = 0):
        h = C.c_void_p()
        _check(lib().yams_b200_digest_set_create(None, capacity_hint, C.byref(h)), "digest_set_create")
        self._h = h

    @staticmethod
    def _digests(x):
        """chunk table (structured, field 'digest') or [n,32] uint8 -> (base pointer, stride, n, keepalive)"""
        a = np.asarray(x)
        if a.dtype.names and "digest" in a.dtype.names:
            a = np.ascontiguousarray(a)
            off = a.dtype.fields["digest"][1]
            return a.ctypes.data + off, a.dtype.itemsize, len(a), a
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        return a.ctypes.data, 32, len(a), a

    def insert(self, digests):
        """-> (existed uint8[n], n_new)"""
        ptr, stride, n, keep = self._digests(digests)
        existed = np.zeros(max(n, 1), dtype=np.uint8)
        new = C.c_uint64(0)
        _check(lib().yams_b200_digest_set_insert(self._h, ptr if n else None, stride, n, existed.ctypes.data, C.byref(new)),
               "digest_set_insert")
        return existed[:n], new.value

    def contains(self, digests) -> np.ndarray:
        ptr, stride, n, keep = self._digests(digests)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        _check(lib().yams_b200_digest_set_contains(self._h, ptr if n else None, stride, n, out.ctypes.data), "digest_set_contains")
        return out[:n]

    def __len__(self):
        n = C.c_uint64(0)
        _check(lib().yams_b200_digest_set_size(self._h, C.byref(n)), "digest_set_size")
        return n.value

    def last_ms(self) -> float:
        v = C.c_float(0)
        _check(lib().yams_b200_digest_set_last_ms(self._h, C.byref(v)), "digest_set_last_ms")
        return v.value

    def close(self):
        if getattr(self, "_h", None):
            lib().yams_b200_digest_set_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


BATCH_ALL, BATCH_TOP_K, BATCH_FILTERED = 0, 1, 2


def batch_distance(query, database, metric: int = COSINE, mode: int = BATCH_ALL, k: int = 0, threshold: float = 0.0):
    """distances/batch.hpp: ALL -> dist[n]; TOP_K -> (idx[k], dist[k]); FILTERED -> (idx[m], dist[m])."""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
    db = np.ascontiguousarray(database, dtype=np.float32)
    n = db.shape[0] if db.ndim == 2 else 0
    out_i = np.zeros(max(n, 1), dtype=np.uint64)
    out_d = np.zeros(max(n, 1), dtype=np.float32)
    cnt = C.c_uint64(0)
    rc = lib().yams_b200_batch_distance(None, metric, q.ctypes.data_as(f32p), q.size, db.ctypes.data_as(f32p) if n else None,
                                        n, mode, k, threshold, out_i.ctypes.data_as(u64p), out_d.ctypes.data_as(f32p),
                                        C.byref(cnt))
    _check(rc, "batch_distance")
    if mode == BATCH_ALL:
        return out_d[:cnt.value].copy()
    return out_i[:cnt.value].copy(), out_d[:cnt.value].copy()


def compute_cosine_similarity(a, b) -> float:
    """VectorDatabase::computeCosineSimilarity (double)."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    out = C.c_double(0)
    _check(lib().yams_b200_compute_cosine_similarity(None, a.ctypes.data_as(f32p), a.size, b.ctypes.data_as(f32p), b.size,
                                                     C.byref(out)), "compute_cosine_similarity")
    return out.value


def compute_cosine_similarity_many(a, b) -> np.ndarray:
    """n pairs in one device pass: a[n, d], b[n, d] -> float64[n]."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 2
    out = np.zeros(a.shape[0], dtype=np.float64)
    _check(lib().yams_b200_compute_cosine_similarity_many(None, a.ctypes.data_as(f32p), b.ctypes.data_as(f32p), a.shape[0], a.shape[1],
                                                          out.ctypes.data_as(C.POINTER(C.c_double))), "compute_cosine_similarity_many")
    return out


def _pair(fn, a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = C.c_float(0)
    rc = fn(a.ctypes.data, a.nbytes, b.ctypes.data, b.nbytes, C.byref(out))
    return rc, out.value


def vec_distance_l2(a, b):
    return _pair(lib().sqlite3_vec_distance_l2, a, b)


def vec_distance_cosine(a, b):
    return _pair(lib().sqlite3_vec_distance_cosine, a, b)


def vec_distance_l1(a, b):
    return _pair(lib().yams_b200_vec_distance_l1, a, b)
