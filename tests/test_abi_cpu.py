"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/yams_b200.h declares, follows the plugin-envelope conventions, and fails LOUDLY (no CPU fallback)
when asked to compute without a device."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Y():
    so = os.path.join(ROOT, "yams_b200", "libyams_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "yams_b200", "csrc")])
    import yams_b200
    return yams_b200


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yams_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define\s+YAMS_B200_API.*", "", text)
    names = set(re.findall(r"YAMS_B200_API\s+[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text))
    return names


def test_header_symbols_all_exported_and_bound(Y):
    from yams_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 35
    L = C.CDLL(Y.lib_path())
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/yams_b200.h but not exported"
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)


def test_plugin_envelope(Y):
    L = Y.lib()
    assert L.yams_plugin_get_abi_version() == 1            # abi.h:19
    assert L.yams_plugin_get_name() == b"yams_b200"
    man = json.loads(L.yams_plugin_get_manifest_json())
    ids = {(i["id"], i["version"]) for i in man["interfaces"]}
    assert ids == {("vector_scan_v1", 1), ("content_ingest_v1", 1)}
    # the reference loader regex-parses "name"/"version"/"interfaces" (abi_plugin_loader.cpp:54-79)
    assert man["name"] == "yams_b200" and re.match(r"\d+\.\d+\.\d+", man["version"])
    iface = C.c_void_p()
    assert L.yams_plugin_get_interface(None, 1, C.byref(iface)) == -4      # YAMS_PLUGIN_ERR_INVALID


def test_struct_layouts(Y):
    from yams_b200 import _lib
    assert C.sizeof(_lib.ChunkDesc) == 48
    assert C.sizeof(_lib.CdcConfig) == 48
    cfg = Y.default_config()
    # ChunkingConfig defaults: chunker.h:44-51, core/types.h:280-285
    assert (cfg.window_size, cfg.min_chunk, cfg.max_chunk, cfg.polynomial, cfg.mask, cfg.variant) == \
        (48, 16384, 1048576, 0x3DA3358B4DC173, 0x1FFF, 0)


def test_no_cpu_fallback_without_gpu(Y):
    if Y.device_count() > 0:
        pytest.skip("a GPU is present")
    assert Y.plugin_init() == -3                            # YAMS_PLUGIN_ERR_INIT_FAILED
    with pytest.raises(Y.YamsB200Error) as e:
        Y.chunk_and_hash(b"abc")
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)
    with pytest.raises(Y.YamsB200Error):
        Y.Corpus(8)
    rc, _ = Y.vec_distance_l2([1.0, 2.0], [1.0, 2.0])
    assert rc != 0
    # every other compute entry fails loudly as well -- nothing is answered on the CPU
    import numpy as np
    for call in (lambda: Y.chunk_and_hash_batch([b"abc", b"defg"]),
                 lambda: Y.chunk_boundaries(b"abc"),
                 lambda: Y.sha256_many([b"abc"]),
                 lambda: Y.sha256_batch(np.frombuffer(b"abcdef", dtype=np.uint8), [0], [3]),
                 lambda: Y.dedup_stats(np.zeros(2, dtype=Y.lib_chunk_dtype())),
                 lambda: Y.DigestSet(),
                 lambda: Y.IngestSession().feed(b"abc"),
                 lambda: Y.batch_distance(np.ones(4, np.float32), np.ones((3, 4), np.float32)),
                 lambda: Y.compute_cosine_similarity([1.0, 2.0], [2.0, 1.0]),
                 lambda: Y.vec0_exact(np.ones(4, np.float32), np.ones((3, 4), np.float32))):
        with pytest.raises(Y.YamsB200Error):
            call()
    assert "no CUDA device" in Y.health()["last_error"]


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under yams_b200/ or include/ may reference it."""
    bad = []
    for base in ("yams_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|yams_oracle\.h|libyams_oracle|oracle/_ref|libyams_ref", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
