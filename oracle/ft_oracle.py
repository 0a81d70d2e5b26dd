"""ORACLE / TEST INFRASTRUCTURE ONLY: ctypes access to the full-text merge checkers.

  ref  -- oracle/_ref/liboracle_ref_ft.so: the reference's own ft::Merger compiled in place (oracle/ref_ft_facade.cc)
  port -- oracle/liboracle_port.so: the plain-C restatement (oracle/ft_port.c)
The flat problem structs (oracle/ft_problem.h) have the same layout as the product ABI's rxgpu_ft_* structs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OP_OR, OP_AND, OP_NOT = 1, 2, 3
RANK_ONLY, RANK_AND_ID, EXTERNAL_EXPR, ID_ONLY, ID_AND_POSITIONS = 0, 1, 2, 3, 4  # reindexer::RankSortType (core/enums.h:103)

_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


class Postings(C.Structure):
    _fields_ = [("ndocs", C.c_uint32), ("doc_ids", _u32p), ("pos_begin", _u32p), ("positions", _u32p)]


class FieldConfig(C.Structure):
    _fields_ = [("bm25_boost", C.c_double), ("bm25_weight", C.c_double), ("term_len_boost", C.c_double), ("term_len_weight", C.c_double),
                ("position_boost", C.c_double), ("position_weight", C.c_double)]


class Config(C.Structure):
    _fields_ = [("merge_limit", C.c_uint32), ("min_rank", C.c_int32), ("bm25_k1", C.c_double), ("bm25_b", C.c_double),
                ("bm25_type", C.c_int32), ("distance_boost", C.c_double), ("distance_weight", C.c_double),
                ("full_match_boost", C.c_double), ("nfields", C.c_uint32), ("fields", C.POINTER(FieldConfig)),
                ("summation_ranks_by_fields_ratio", C.c_double)]


class Term(C.Structure):
    _fields_ = [("op", C.c_int32), ("boost", C.c_float), ("term_len_boost", C.c_float), ("field_boosts", _f32p), ("nsubterms", C.c_uint32),
                ("postings", _u32p), ("procs", _f32p), ("need_sum_rank", _u8p), ("suppressed", _u8p), ("nsynonyms", C.c_uint32),
                ("synonym_ids", _u32p), ("phrase_num", C.c_int32), ("distance", C.c_int32)]


class Synonym(C.Structure):
    _fields_ = [("nterms", C.c_uint32), ("terms", C.POINTER(Term))]


class MergeInfo(C.Structure):
    _fields_ = [("id", C.c_int32), ("proc", C.c_float), ("field", C.c_uint8), ("normalized_proc", C.c_uint8)]


MERGE_INFO_DTYPE = np.dtype([("id", np.int32), ("proc", np.float32), ("field", np.uint8), ("normalized_proc", np.uint8)], align=True)
assert MERGE_INFO_DTYPE.itemsize == C.sizeof(MergeInfo) == 12


def _p(a, t):
    return a.ctypes.data_as(t)


class FtProblem:
    """Owns the numpy arrays of one merge problem and exposes the ctypes views (kept alive by this object)."""

    def __init__(self, total_docs, words, avg=None, removed=None, excluded=None):
        words = np.ascontiguousarray(words, np.uint32)
        if words.ndim == 1:
            words = words.reshape(-1, 1)
        assert words.shape[0] == total_docs
        self.total_docs, self.nfields = total_docs, words.shape[1]
        self.words = words
        if avg is None:  # AvgWordsCount over the real documents (vdoc 0 is a dummy)
            avg = words[1:].mean(axis=0) if total_docs > 1 else np.ones(self.nfields)
        self.avg = np.ascontiguousarray(avg, np.float32)
        self.removed = None if removed is None else np.ascontiguousarray(removed, np.uint8)
        self.excluded = None if excluded is None else np.ascontiguousarray(excluded, np.uint8)
        self.lists = []  # (doc_ids, pos_begin, positions)
        self.terms = []  # dict(op, boost, term_len_boost, field_boosts, postings, procs)
        self.synonyms = []  # multi-word synonyms: lists of term dicts (ft::Synonym)
        self.cfg = dict(merge_limit=20000, min_rank=5, bm25_k1=2.0, bm25_b=0.75, bm25_type=0, distance_boost=1.0, distance_weight=0.5,
                        full_match_boost=1.1, summation_ranks_by_fields_ratio=0.0)
        self.field_cfg = [dict(bm25_boost=1.0, bm25_weight=0.1, term_len_boost=1.0, term_len_weight=0.3, position_boost=1.0,
                               position_weight=0.1) for _ in range(self.nfields)]

    def add_list(self, doc_ids, positions_per_doc):
        """positions_per_doc: list (per doc) of iterables of (pos, field); stored sorted by (field, pos)."""
        doc_ids = np.ascontiguousarray(doc_ids, np.uint32)
        begin = np.zeros(len(doc_ids) + 1, np.uint32)
        flat = []
        for i, pp in enumerate(positions_per_doc):
            packed = sorted({(int(f) << 24) | int(p) for p, f in pp})
            flat.extend(packed)
            begin[i + 1] = len(flat)
        self.lists.append((doc_ids, begin, np.ascontiguousarray(flat, np.uint32)))
        return len(self.lists) - 1

    def add_list_arrays(self, doc_ids, pos_begin, positions):
        self.lists.append((np.ascontiguousarray(doc_ids, np.uint32), np.ascontiguousarray(pos_begin, np.uint32),
                           np.ascontiguousarray(positions, np.uint32)))
        return len(self.lists) - 1

    def _term(self, subterms, op, boost, term_len_boost, field_boosts, need_sum_rank, synonym_ids=(), phrase_num=0, distance=0):
        fb = np.ones(self.nfields, np.float32) if field_boosts is None else np.ascontiguousarray(field_boosts, np.float32)
        ns = None if need_sum_rank is None else np.ascontiguousarray(need_sum_rank, np.uint8)
        sup = np.ascontiguousarray([1 if len(s) > 2 and s[2] else 0 for s in subterms], np.uint8)
        return dict(op=op, boost=boost, term_len_boost=term_len_boost, field_boosts=fb, need_sum_rank=ns,
                    postings=np.ascontiguousarray([s[0] for s in subterms], np.uint32),
                    procs=np.ascontiguousarray([s[1] for s in subterms], np.float32), suppressed=sup if sup.any() else None,
                    synonym_ids=np.ascontiguousarray(list(synonym_ids), np.uint32), phrase_num=int(phrase_num), distance=int(distance))

    def add_term(self, subterms, op=OP_OR, boost=1.0, term_len_boost=1.0, field_boosts=None, need_sum_rank=None, synonym_ids=(), phrase_num=0,
                 distance=0):
        """subterms: list of (list id, proc); synonym_ids: indexes of multi-word synonyms of this query part (add_synonym);
        phrase_num: consecutive terms with the same non-zero number form a phrase, distance = FtDslOpts::distance of the term"""
        self.terms.append(self._term(subterms, op, boost, term_len_boost, field_boosts, need_sum_rank, synonym_ids, phrase_num, distance))

    def add_synonym(self, terms):
        """terms: list of dict(subterms=[(list id, proc[, suppressed])], op=..., boost=..., term_len_boost=..., field_boosts=...);
        returns the synonym's index"""
        self.synonyms.append([self._term(t["subterms"], t.get("op", OP_OR), t.get("boost", 1.0), t.get("term_len_boost", 1.0),
                                         t.get("field_boosts"), t.get("need_sum_rank")) for t in terms])
        return len(self.synonyms) - 1

    # ctypes views --------------------------------------------------------------------------------------------------------
    def c_lists(self):
        arr = (Postings * max(len(self.lists), 1))()
        for i, (d, b, p) in enumerate(self.lists):
            arr[i] = Postings(len(d), _p(d, _u32p), _p(b, _u32p), _p(p, _u32p))
        return arr

    def c_config(self):
        self._fc = (FieldConfig * self.nfields)(*[FieldConfig(**f) for f in self.field_cfg])
        return Config(self.cfg["merge_limit"], self.cfg["min_rank"], self.cfg["bm25_k1"], self.cfg["bm25_b"], self.cfg["bm25_type"],
                      self.cfg["distance_boost"], self.cfg["distance_weight"], self.cfg["full_match_boost"], self.nfields, self._fc,
                      self.cfg.get("summation_ranks_by_fields_ratio", 0.0))

    @staticmethod
    def _c_term(t):
        return Term(t["op"], t["boost"], t["term_len_boost"], _p(t["field_boosts"], _f32p), len(t["postings"]), _p(t["postings"], _u32p),
                    _p(t["procs"], _f32p), None if t.get("need_sum_rank") is None else _p(t["need_sum_rank"], _u8p),
                    None if t.get("suppressed") is None else _p(t["suppressed"], _u8p), len(t.get("synonym_ids", ())),
                    _p(t["synonym_ids"], _u32p) if len(t.get("synonym_ids", ())) else None, t.get("phrase_num", 0), t.get("distance", 0))

    def c_terms(self):
        arr = (Term * max(len(self.terms), 1))()
        for i, t in enumerate(self.terms):
            arr[i] = self._c_term(t)
        return arr

    def c_synonyms(self):
        arr = (Synonym * max(len(self.synonyms), 1))()
        self._syn_terms = []
        for i, syn in enumerate(self.synonyms):
            ta = (Term * len(syn))(*[self._c_term(t) for t in syn])
            self._syn_terms.append(ta)
            arr[i] = Synonym(len(syn), ta)
        return arr


_ref = None
_port_ready = False


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "liboracle_ref_ft.so"))


def _merge_argtypes(fn, with_packed):
    args = [C.c_uint32, C.c_uint32, _u32p, _f32p, _u8p, _u8p, C.c_uint32, C.POINTER(Postings), C.POINTER(Config), C.c_uint32,
            C.POINTER(Term), C.c_int]
    if with_packed:
        args.append(C.c_int)
    args += [C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    fn.argtypes = args
    fn.restype = C.c_int


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(HERE, "_ref", "liboracle_ref_ft.so"))
        lib.ref_ft_last_error.restype = C.c_char_p
        _merge_argtypes(lib.ref_ft_merge, True)
        lib.ref_ft_merge_query.restype = C.c_int
        lib.ref_ft_merge_query.argtypes = [C.c_uint32, C.c_uint32, _u32p, _f32p, _u8p, _u8p, C.c_uint32, C.POINTER(Postings), C.POINTER(Config),
                                           C.c_uint32, C.POINTER(Term), C.c_uint32, C.POINTER(Synonym), C.c_int, C.c_int, C.c_uint64,
                                           C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        lib.ref_ft_calc_term_rank.restype = C.c_int
        lib.ref_ft_calc_term_rank.argtypes = [C.c_uint32, C.POINTER(Config), C.POINTER(Term), C.c_float, C.c_double, C.c_double, C.c_uint32,
                                              _u32p, _u32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_int)]
        lib.ref_ft_pack_list.restype = C.c_int64
        lib.ref_ft_pack_list.argtypes = [C.POINTER(Postings), _u8p, C.c_uint64]
        _ref = lib
    return _ref


def ref_pack_list(doc_ids, pos_begin, positions):
    """the reference's packed posting stream (PackedIdRelVec::data_) of one list, produced by the reference's own encoder"""
    d = np.ascontiguousarray(doc_ids, np.uint32)
    b = np.ascontiguousarray(pos_begin, np.uint32)
    q = np.ascontiguousarray(positions, np.uint32)
    pl = Postings(len(d), d.ctypes.data_as(_u32p), b.ctypes.data_as(_u32p), q.ctypes.data_as(_u32p))
    cap = 16 + len(d) * 16 + len(q) * 12
    out = np.zeros(cap, np.uint8)
    n = ref_lib().ref_ft_pack_list(C.byref(pl), out.ctypes.data_as(_u8p), cap)
    assert n >= 0, n
    return out[:n].copy()


def port_lib():
    global _port_ready
    from . import oracle as O

    lib = O.port_lib()
    if not _port_ready:
        _merge_argtypes(lib.port_ft_merge, False)
        _port_ready = True
    return lib


def _run(fn, prob: FtProblem, rank_sort_type, packed=None, max_out=None):
    lists, cfg, terms = prob.c_lists(), prob.c_config(), prob.c_terms()
    max_out = prob.total_docs if max_out is None else max_out
    out = np.zeros(max(max_out, 1), MERGE_INFO_DTYPE)
    n, ns = C.c_uint64(0), C.c_int64(0)
    args = [prob.total_docs, prob.nfields, _p(prob.words, _u32p), _p(prob.avg, _f32p),
            None if prob.removed is None else _p(prob.removed, _u8p), None if prob.excluded is None else _p(prob.excluded, _u8p),
            len(prob.lists), lists, C.byref(cfg), len(prob.terms), terms, rank_sort_type]
    if packed is not None:
        args.append(int(packed))
    args += [max_out, out.ctypes.data, C.byref(n), C.byref(ns)]
    rc = fn(*args)
    return rc, out[:min(n.value, max_out)].copy(), n.value, ns.value


def ref_merge(prob, rank_sort_type=RANK_AND_ID, packed=False):
    if prob.synonyms or any(t.get("phrase_num", 0) for t in prob.terms):
        lists, cfg, terms, syns = prob.c_lists(), prob.c_config(), prob.c_terms(), prob.c_synonyms()
        out = np.zeros(max(prob.total_docs, 1), MERGE_INFO_DTYPE)
        n, ns = C.c_uint64(0), C.c_int64(0)
        rc = ref_lib().ref_ft_merge_query(prob.total_docs, prob.nfields, _p(prob.words, _u32p), _p(prob.avg, _f32p),
                                          None if prob.removed is None else _p(prob.removed, _u8p),
                                          None if prob.excluded is None else _p(prob.excluded, _u8p), len(prob.lists), lists, C.byref(cfg),
                                          len(prob.terms), terms, len(prob.synonyms), syns, rank_sort_type, int(packed), prob.total_docs,
                                          out.ctypes.data, C.byref(n), C.byref(ns))
        assert rc == 0, ref_lib().ref_ft_last_error().decode()
        return out[:n.value].copy(), ns.value
    rc, out, n, ns = _run(ref_lib().ref_ft_merge, prob, rank_sort_type, packed)
    assert rc == 0, ref_lib().ref_ft_last_error().decode()
    return out, ns


def port_merge(prob, rank_sort_type=RANK_AND_ID):
    assert not prob.synonyms and not any(t.get("phrase_num", 0) for t in prob.terms), \
        "the C port does not restate multi-word synonyms and phrases; use the reference facade (oracle/_ref)"
    rc, out, n, ns = _run(port_lib().port_ft_merge, prob, rank_sort_type)
    assert rc == 0, "port_ft_merge failed"
    return out, ns


def best_merge(prob, rank_sort_type=RANK_AND_ID):
    return ref_merge(prob, rank_sort_type) if ref_available() else port_merge(prob, rank_sort_type)


def after_select_order(res):
    """IndexText::sortAfterSelect for RankAndID (cpp_src/core/index/indextext/indextext.cc:487-498): rank desc, id asc."""
    order = np.lexsort((res["id"], -res["normalized_proc"].astype(np.int32)))
    return res[order]


def ref_calc_term_rank(prob: FtProblem, term_idx, subterm_proc, total_docs, matched_docs, positions, words_in_fields):
    lib = ref_lib()
    cfg, terms = prob.c_config(), prob.c_terms()
    pos = np.ascontiguousarray(positions, np.uint32)
    w = np.ascontiguousarray(words_in_fields, np.uint32)
    tr, bn, pr, tl, fld = C.c_float(), C.c_float(), C.c_float(), C.c_float(), C.c_int()
    rc = lib.ref_ft_calc_term_rank(prob.nfields, C.byref(cfg), C.byref(terms[term_idx]), subterm_proc, total_docs, matched_docs, len(pos),
                                   _p(pos, _u32p), _p(w, _u32p), _p(prob.avg, _f32p), C.byref(tr), C.byref(bn), C.byref(pr),
                                   C.byref(tl), C.byref(fld))
    assert rc == 0
    return dict(term_rank=tr.value, bm25_norm=bn.value, position_rank=pr.value, term_len_boost=tl.value, field=fld.value)
