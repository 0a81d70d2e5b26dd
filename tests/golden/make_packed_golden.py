#!/usr/bin/env python
"""Generates tests/golden/packed_golden.npz: posting lists and their packed byte streams produced by the REFERENCE's own encoder
(IdRelType::packWithoutArrayIdxs through oracle/_ref, see oracle/ref_ft_facade.cc::ref_ft_pack_list).  Run in the authoring
container (needs /root/reference to have built oracle/_ref); the fixture travels to the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ft_oracle as F  # noqa: E402


def make_list(rng, ndocs, max_doc, nfields, max_pos, max_word):
    docs = np.sort(rng.choice(np.arange(1, max_doc), size=ndocs, replace=False)).astype(np.uint32)
    npos = rng.integers(1, max_pos + 1, size=ndocs)
    begin = np.concatenate([[0], np.cumsum(npos)]).astype(np.uint32)
    positions = np.zeros(begin[-1], np.uint32)
    for i in range(ndocs):
        f = np.sort(rng.integers(0, nfields, size=npos[i]))          # positions ascending by (field, pos) like the indexer emits them
        w = rng.integers(0, max_word, size=npos[i])
        order = np.lexsort((w, f))
        positions[begin[i]:begin[i + 1]] = (w[order] | (f[order] << 24)).astype(np.uint32)
    return docs, begin, positions


def main():
    rng = np.random.default_rng(20260923)
    out = {}
    cases = [("tiny", 1, 10, 1, 1, 5), ("single_field", 300, 5000, 1, 4, 200), ("multi_field", 500, 100000, 5, 6, 3000),
             ("long_words", 200, 1 << 22, 3, 3, 1 << 23), ("dense", 4000, 4100, 2, 2, 100), ("wide_fields", 100, 1000, 64, 5, 50)]
    names = []
    for name, ndocs, max_doc, nfields, max_pos, max_word in cases:
        d, b, p = make_list(rng, ndocs, max_doc, nfields, max_pos, max_word)
        out[f"{name}/doc_ids"], out[f"{name}/pos_begin"], out[f"{name}/positions"] = d, b, p
        out[f"{name}/packed"] = F.ref_pack_list(d, b, p)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "packed_golden.npz"), **out)
    print({n: (len(out[f"{n}/doc_ids"]), len(out[f"{n}/packed"])) for n in names})


if __name__ == "__main__":
    main()
