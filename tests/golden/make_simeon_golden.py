"""Generate tests/golden/simeon_golden.json by EXECUTING simeon's own Encoder (oracle/_ref, built from /root/reference/third_party/simeon
by oracle/Makefile).  Run in the build container only:  python tests/golden/make_simeon_golden.py

The fixture pins the device encoder on machines where the reference cannot be built: embeddings (as uint32 bit patterns) of a
few fixed texts under the two profiles YAMS runs -- `simeon-v1-384` and the configurable default (CharAndWord + Fwht, 4096 -> 1024).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

TEXTS = ["", "a", "abc", "The quick brown fox jumps over the lazy dog", "content addressed storage: chunk, hash, dedup_42",
         "naïve 日本語 emoji🙂 _under_score_ x", "word " * 200]


def main():
    assert O.ref_available()
    out = {"texts": TEXTS,
           "simeon_v1_384": O.simeon_encode_ref(TEXTS).view(np.uint32).tolist(),
           "yams_default_1024": O.simeon_encode_modes_ref(TEXTS, ngram_mode="CharAndWord", projection="Fwht", output_dim=1024).view(np.uint32).tolist(),
           "yams_default_384": O.simeon_encode_modes_ref(TEXTS, ngram_mode="CharAndWord", projection="Fwht", output_dim=384).view(np.uint32).tolist()}
    with open(os.path.join(HERE, "simeon_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "simeon_golden.json"))


if __name__ == "__main__":
    main()
