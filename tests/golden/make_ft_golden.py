"""Generates tests/golden/ft_golden.npz from the REFERENCE's own merger (oracle/_ref/liboracle_ref_ft.so).
Run in the authoring container only:  python tests/golden/make_ft_golden.py
Stores, per case, the full input (posting arrays, doc statistics, query, config) next to ft::Merger::Merge's output, plus the
calcTermRank values that the reference's own test FTGenericApi.DebugInfo pins (gtests/tests/unit/ft/ft_generic.cc:297-445)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ft_helpers import random_problem  # noqa: E402

from oracle import ft_oracle as F  # noqa: E402

assert F.ref_available()
out = {}
CASES = [  # name, kwargs
    ("simple_1field", dict(seed=11, total_docs=300, nfields=1, nterms=1, max_sub=3)),
    ("simple_3fields", dict(seed=12, total_docs=300, nfields=3, nterms=1, max_sub=3, field_boost_zero=True)),
    ("or3_1field", dict(seed=13, total_docs=500, nfields=1, nterms=3, ops=[F.OP_OR] * 3)),
    ("and_not_2fields", dict(seed=14, total_docs=500, nfields=2, nterms=4, ops=[F.OP_OR, F.OP_AND, F.OP_NOT, F.OP_OR], removed_frac=0.05,
                             excluded_frac=0.05)),
    ("preselect_small_limit", dict(seed=15, total_docs=600, nfields=2, nterms=3, ops=[F.OP_OR] * 3, merge_limit=40, density=0.5)),
    ("preselect_removed", dict(seed=16, total_docs=400, nfields=1, nterms=2, ops=[F.OP_OR] * 2, merge_limit=25, density=0.6,
                               removed_frac=0.1)),
]
names = []
for name, kw in CASES:
    seed = kw.pop("seed")
    p = random_problem(seed, **kw)
    names.append(name)
    out[f"{name}/words"], out[f"{name}/avg"] = p.words, p.avg
    out[f"{name}/removed"] = p.removed if p.removed is not None else np.zeros(0, np.uint8)
    out[f"{name}/excluded"] = p.excluded if p.excluded is not None else np.zeros(0, np.uint8)
    out[f"{name}/nlists"] = np.int64(len(p.lists))
    for i, (d, b, pos) in enumerate(p.lists):
        out[f"{name}/list{i}/docs"], out[f"{name}/list{i}/begin"], out[f"{name}/list{i}/pos"] = d, b, pos
    out[f"{name}/nterms"] = np.int64(len(p.terms))
    for i, t in enumerate(p.terms):
        out[f"{name}/term{i}/scalars"] = np.array([t["op"], t["boost"], t["term_len_boost"]], np.float64)
        out[f"{name}/term{i}/field_boosts"], out[f"{name}/term{i}/postings"], out[f"{name}/term{i}/procs"] = (
            t["field_boosts"], t["postings"], t["procs"])
    out[f"{name}/merge_limit"] = np.int64(p.cfg["merge_limit"])
    for rst in (F.RANK_AND_ID, F.RANK_ONLY):
        res, _ = F.ref_merge(p, rst)
        out[f"{name}/result{rst}"] = res
out["names"] = np.array(names)

# KAT of FTGenericApi.DebugInfo: 5 docs of 8,3,6,12,3 words; "маша" hits doc 1 at pos 0 (proc 100) and its stem variant at pos 6 (proc 80)
p = F.FtProblem(6, np.array([0, 8, 3, 6, 12, 3], np.uint32))
l0, l1 = p.add_list([1], [[(0, 0)]]), p.add_list([1], [[(6, 0)]])
p.add_term([(l0, 100.0), (l1, 80.0)])
a = F.ref_calc_term_rank(p, 0, 100.0, 5, 1, [0], [8])
b = F.ref_calc_term_rank(p, 0, 80.0, 5, 1, [6], [8])
assert f"{a['term_rank']:.4f}" == "97.9844" and f"{a['bm25_norm']:.6f}" == "0.979844", a  # ft_generic.cc:326
assert f"{b['term_rank']:.5f}" == "77.91719" and f"{b['position_rank']:.3f}" == "0.994", b  # ft_generic.cc:327
out["kat/term_rank"] = np.array([a["term_rank"], b["term_rank"]], np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ft_golden.npz"), **out)
print("wrote", len(out), "arrays")
