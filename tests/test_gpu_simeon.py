"""GPU parity of the Simeon text encoder's default profile against simeon's own Encoder compiled in place
(third_party/simeon/src/{simeon,projection,hasher,tokenizer,...}.cpp): bit-identical embeddings."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Y():
    import yams_b200
    assert yams_b200.device_count() > 0
    assert yams_b200.plugin_init() == 0, yams_b200.health()
    return yams_b200


def corpus_of_texts(rng, n):
    words = ["yams", "content", "addressed", "storage", "chunk", "hash", "vector", "search", "naïve", "日本語", "emoji🙂", "x", "_", "42"]
    texts = [b"", b"a", b"ab", b"abc", b"abcd", "The quick brown fox jumps over the lazy dog".encode()]
    for _ in range(n):
        kind = rng.integers(0, 3)
        if kind == 0:
            L = int(rng.integers(1, 60))
            texts.append(" ".join(words[int(i)] for i in rng.integers(0, len(words), size=L)).encode("utf-8"))
        elif kind == 1:
            texts.append(bytes(rng.integers(0, 256, size=int(rng.integers(1, 3000)), dtype=np.uint8)))   # arbitrary bytes, NULs included
        else:
            texts.append(bytes(rng.integers(97, 123, size=int(rng.integers(5, 20000)), dtype=np.uint8)))
    return texts


@pytest.mark.parametrize("cfg", [
    dict(),                                                         # simeon-v1-384: what YAMS runs
    dict(sketch_dim=8192, output_dim=768),                          # the 768-d variant of BASELINE config C2
    dict(ngram_min=1, ngram_max=8),                                 # 8-byte grams take the whole-word hash path
    dict(ngram_min=7, ngram_max=12, sketch_dim=5000, output_dim=100),   # > 8 bytes: word + tail; non-power-of-two sketch (modulo bucket)
    dict(l2_normalize=0, output_dim=33),                            # raw projection, odd dimension (scalar tail of the normaliser unused)
    dict(output_dim=40),                                            # 40 = 2 x 16 + 8: the normaliser's fused tail
])
def test_simeon_encoder_is_bit_identical(Y, oracle, cfg):
    O = oracle
    if not O.ref_available():
        pytest.skip("needs simeon compiled in place")
    rng = np.random.default_rng(len(cfg) + 1)
    texts = corpus_of_texts(rng, 300)
    enc = Y.SimeonEncoder(**cfg)
    got = enc.encode(texts)
    want = O.simeon_encode_ref(texts, **cfg)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if cfg.get("l2_normalize", 1):
        norms = np.linalg.norm(got.astype(np.float64), axis=1)
        assert np.all((np.abs(norms - 1) < 1e-5) | (norms == 0))    # texts shorter than ngram_min embed to zero
    # one text at a time == the batch (generateEmbedding vs generateEmbeddings, simeon_embedding_backend.cpp:183-209)
    for i in (0, 5, 17):
        assert np.array_equal(enc.encode([texts[i]])[0], got[i])
    enc.close()


def test_simeon_embeddings_feed_the_scan(Y, oracle):
    """End of the pipeline the encoder exists for: encode -> corpus -> exact search finds the text itself first."""
    rng = np.random.default_rng(9)
    texts = corpus_of_texts(rng, 2000)[6:]
    enc = Y.SimeonEncoder()
    emb = enc.encode(texts)
    c = Y.Corpus(384, Y.F32, Y.COSINE)
    c.append(emb)
    rid, sc, cnt, _ = c.search(emb[:50], 3, threshold=-1.0)
    assert all(rid[i, 0] == i or sc[i, 0] == sc[i, 1] for i in range(50)) and np.all(sc[:, 0] > 0.9999)
    c.close()
    enc.close()
