// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product (librxgpu.so).
//
// extern "C" facade over the *unmodified* reference full-text merger, compiled in place from /root/reference/cpp_src:
//   ft::Merger<IdCont, ft::MergeData, OffsetT>::Merge<Bm25Rx|Bm25Classic|TermCount>   core/ft/ft_fast/mergerimpl.h:466-566
//   calcTermRank                                                                        core/ft/ft_fast/phrasemergerimpl.h:13-91
//   IdRelVec / PackedIdRelVec posting containers                                        core/ft/idrelset.{h,cc}
// The container/offset-type choice follows Selector::Process (ft_fast/selecterimpl.h:629-645).
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

// clang-format off
#include "tools/float_comparison.h"
#include "core/ft/ft_fast/mergerimpl.h"
// clang-format on
#include "core/ft/idrelset.h"
#include "core/rdxcontext.h"
#include "ft_problem.h"

namespace {

thread_local std::string g_err;

struct Stats {
	const uint32_t* words;
	const float* avg;
	const uint8_t* removed;
	uint32_t nfields;
	bool DocRemoved(uint32_t vdoc) const noexcept { return removed && removed[vdoc]; }
	size_t NumWordsInField(uint32_t vdoc, uint32_t f) const noexcept { return words[size_t(vdoc) * nfields + f]; }
	float AvgWordsCount(uint32_t f) const noexcept { return avg[f]; }
};

void fillConfig(reindexer::FTConfig& c, const ft_config* in) {
	c.mergeLimit = in->merge_limit;
	c.minRank = in->min_rank;
	c.bm25Config.bm25k1 = in->bm25_k1;
	c.bm25Config.bm25b = in->bm25_b;
	c.distanceBoost = in->distance_boost;
	c.distanceWeight = in->distance_weight;
	c.fullMatchBoost = in->full_match_boost;
	c.summationRanksByFieldsRatio = in->summation_ranks_by_fields_ratio;
	for (uint32_t f = 0; f < in->nfields; ++f) {
		auto& fc = c.fieldsCfg[f];
		fc.bm25Boost = in->fields[f].bm25_boost;
		fc.bm25Weight = in->fields[f].bm25_weight;
		fc.termLenBoost = in->fields[f].term_len_boost;
		fc.termLenWeight = in->fields[f].term_len_weight;
		fc.positionBoost = in->fields[f].position_boost;
		fc.positionWeight = in->fields[f].position_weight;
	}
}

reindexer::IdRelType makeRel(const ft_postings& l, uint32_t i) {
	reindexer::IdRelType r(l.doc_ids[i]);
	for (uint32_t p = l.pos_begin[i]; p < l.pos_begin[i + 1]; ++p) {
		r.Add(l.positions[p] & 0xFFFFFF, l.positions[p] >> 24, 0);
	}
	return r;
}

template <typename IdCont>
void buildList(const ft_postings& l, IdCont& out);
template <>
void buildList(const ft_postings& l, reindexer::IdRelVec& out) {
	out.reserve(l.ndocs);
	for (uint32_t i = 0; i < l.ndocs; ++i) {
		out.emplace_back(makeRel(l, i));
	}
}
template <>
void buildList(const ft_postings& l, reindexer::PackedIdRelVec& out) {
	std::vector<reindexer::IdRelType> tmp;
	tmp.reserve(l.ndocs);
	for (uint32_t i = 0; i < l.ndocs; ++i) {
		tmp.emplace_back(makeRel(l, i));
	}
	out.insert_back(tmp.begin(), tmp.end());
}

template <typename IdCont>
reindexer::ft::TermResults<IdCont> makeTerm(const ft_term& t, uint32_t nfields, std::vector<IdCont>& lists) {
	reindexer::FtDSLEntry e;
	e.Opts().op = OpType(t.op);
	e.Opts().boost = t.boost;
	e.Opts().termLenBoost = t.term_len_boost;
	e.Opts().phraseNum = t.phrase_num;
	e.Opts().distance = t.distance;
	e.Opts().fieldsOpts.resize(nfields);
	for (uint32_t f = 0; f < nfields; ++f) {
		e.Opts().fieldsOpts[f].boost = t.field_boosts[f];
		e.Opts().fieldsOpts[f].needSumRank = t.need_sum_rank && t.need_sum_rank[f];
	}
	reindexer::ft::TermResults<IdCont> tr(std::move(e));
	for (uint32_t s = 0; s < t.nsubterms; ++s) {
		tr.AddSubterm(lists[t.postings[s]], "w", reindexer::WordIdType{}, t.procs[s]);
		if (t.suppressed && t.suppressed[s]) {
			tr.Subterm(s).SetSuppressed(true);  // what QueryMergeData::SupressDuplicatesInSynonyms decides from the pattern ids
		}
	}
	return tr;
}

template <typename IdCont, typename OffsetT>
int64_t runMerge(uint32_t totalDocs, const Stats& stats, const uint8_t* excluded, std::vector<IdCont>& lists, reindexer::FTConfig& cfg,
				 uint32_t nterms, const ft_term* terms, uint32_t nsyn, const ft_synonym* syns, int rankSortType,
				 std::vector<reindexer::ft::MergeInfo>& out) {
	reindexer::ft::QueryMergeData<IdCont> q;
	for (uint32_t y = 0; y < nsyn; ++y) {  // the selecter appends synonyms while it walks the terms (selecterimpl.h:440-466,580-603)
		reindexer::ft::Synonym<IdCont> syn;
		for (uint32_t t = 0; t < syns[y].nterms; ++t) {
			auto tr = makeTerm(syns[y].terms[t], stats.nfields, lists);
			q.totalORVids += tr.MaxVDocs();
			syn.AddTerm(std::move(tr));
		}
		q.synonyms.emplace_back(std::move(syn));
	}
	reindexer::ft::PhraseResults<IdCont> nextPhrase;  // grouped like the selecter does (selecterimpl.h:546-566)
	int curPhraseNum = 0;
	for (uint32_t t = 0; t < nterms; ++t) {
		auto tr = makeTerm(terms[t], stats.nfields, lists);
		q.totalORVids += tr.MaxVDocs();
		if (terms[t].phrase_num != 0) {
			if (nextPhrase.NumTerms() && curPhraseNum != terms[t].phrase_num) {
				q.queryParts.emplace_back(std::move(nextPhrase));
				nextPhrase.clear();
			}
			curPhraseNum = terms[t].phrase_num;
			nextPhrase.Add(std::move(tr));
			continue;
		}
		if (nextPhrase.NumTerms()) {
			q.queryParts.emplace_back(std::move(nextPhrase));
			nextPhrase.clear();
		}
		q.queryParts.emplace_back(std::move(tr));
		for (uint32_t y = 0; y < terms[t].nsynonyms; ++y) {
			q.queryParts.back().AddSynonymId(terms[t].synonym_ids[y]);
		}
	}
	if (nextPhrase.NumTerms()) {
		q.queryParts.emplace_back(std::move(nextPhrase));
		nextPhrase.clear();
	}
	reindexer::FtMergeStatuses::Statuses docsExcluded(totalDocs, false);
	if (excluded) {
		for (uint32_t i = 0; i < totalDocs; ++i) {
			if (excluded[i]) {
				docsExcluded.set(i);
			}
		}
	}
	reindexer::RdxContext ctx;
	reindexer::ft::Merger<IdCont, reindexer::ft::MergeData, OffsetT> m(totalDocs, &cfg, docsExcluded, stats.nfields, 5, false, ctx);
	const auto t0 = std::chrono::steady_clock::now();
	switch (cfg.bm25Config.bm25Type) {
		case reindexer::FTConfig::Bm25Config::Bm25Type::classic:
			{
				auto r = m.template Merge<reindexer::Bm25Classic>(q, reindexer::RankSortType(rankSortType), stats);
				out.swap(r);
			}
			break;
		case reindexer::FTConfig::Bm25Config::Bm25Type::wordCount:
			{
				auto r = m.template Merge<reindexer::TermCount>(q, reindexer::RankSortType(rankSortType), stats);
				out.swap(r);
			}
			break;
		default:
			{
				auto r = m.template Merge<reindexer::Bm25Rx>(q, reindexer::RankSortType(rankSortType), stats);
				out.swap(r);
			}
	}
	return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
}

template <typename IdCont>
int64_t dispatch(uint32_t totalDocs, const Stats& stats, const uint8_t* excluded, uint32_t nlists, const ft_postings* lists,
				 reindexer::FTConfig& cfg, uint32_t nterms, const ft_term* terms, uint32_t nsyn, const ft_synonym* syns, int rankSortType,
				 std::vector<reindexer::ft::MergeInfo>& out) {
	std::vector<IdCont> conts(nlists);
	for (uint32_t i = 0; i < nlists; ++i) {
		buildList(lists[i], conts[i]);
	}
	uint64_t totalOR = 0;
	auto count = [&](const ft_term& t) {
		for (uint32_t s = 0; s < t.nsubterms; ++s) {
			totalOR += lists[t.postings[s]].ndocs;
		}
	};
	for (uint32_t t = 0; t < nterms; ++t) {
		count(terms[t]);
	}
	for (uint32_t y = 0; y < nsyn; ++y) {
		for (uint32_t t = 0; t < syns[y].nterms; ++t) {
			count(syns[y].terms[t]);
		}
	}
	const uint64_t maxMerged = std::min<uint64_t>(cfg.mergeLimit, totalOR);  // selecterimpl.h:637-644
	if (maxMerged < 0xFFFF) {
		return runMerge<IdCont, uint16_t>(totalDocs, stats, excluded, conts, cfg, nterms, terms, nsyn, syns, rankSortType, out);
	}
	return runMerge<IdCont, uint32_t>(totalDocs, stats, excluded, conts, cfg, nterms, terms, nsyn, syns, rankSortType, out);
}

}  // namespace

extern "C" {

const char* ref_ft_last_error() { return g_err.c_str(); }

// ft::Merger::Merge on one problem.  packed != 0 uses PackedIdRelVec (the default Optimization::Memory container).
// *merge_ns = wall time of the Merge call alone (container construction excluded).  Returns 0 on success.
int ref_ft_merge_query(uint32_t total_docs, uint32_t nfields, const uint32_t* words, const float* avg, const uint8_t* removed,
					   const uint8_t* excluded, uint32_t nlists, const ft_postings* lists, const ft_config* cfg, uint32_t nterms,
					   const ft_term* terms, uint32_t nsyn, const ft_synonym* syns, int rank_sort_type, int packed, uint64_t max_out,
					   ft_merge_info* out, uint64_t* out_n, int64_t* merge_ns) {
	try {
		reindexer::FTConfig c(nfields);
		fillConfig(c, cfg);
		c.bm25Config.bm25Type = reindexer::FTConfig::Bm25Config::Bm25Type(cfg->bm25_type == 1	? 0
																		   : cfg->bm25_type == 2 ? 2
																								 : 1);
		Stats stats{words, avg, removed, nfields};
		std::vector<reindexer::ft::MergeInfo> res;
		const int64_t ns = packed ? dispatch<reindexer::PackedIdRelVec>(total_docs, stats, excluded, nlists, lists, c, nterms, terms, nsyn,
																		syns, rank_sort_type, res)
								  : dispatch<reindexer::IdRelVec>(total_docs, stats, excluded, nlists, lists, c, nterms, terms, nsyn, syns,
																  rank_sort_type, res);
		if (merge_ns) {
			*merge_ns = ns;
		}
		*out_n = res.size();
		for (size_t i = 0; i < res.size() && i < max_out; ++i) {
			out[i].id = res[i].id.ToNumber();
			out[i].proc = res[i].proc;
			out[i].field = res[i].field;
			out[i].normalized_proc = res[i].normalizedProc;
		}
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}
int ref_ft_merge(uint32_t total_docs, uint32_t nfields, const uint32_t* words, const float* avg, const uint8_t* removed,
				 const uint8_t* excluded, uint32_t nlists, const ft_postings* lists, const ft_config* cfg, uint32_t nterms,
				 const ft_term* terms, int rank_sort_type, int packed, uint64_t max_out, ft_merge_info* out, uint64_t* out_n,
				 int64_t* merge_ns) {
	return ref_ft_merge_query(total_docs, nfields, words, avg, removed, excluded, nlists, lists, cfg, nterms, terms, 0, nullptr,
							  rank_sort_type, packed, max_out, out, out_n, merge_ns);
}

// calcTermRank on one posting: the numbers FTGenericApi.DebugInfo pins (gtests/tests/unit/ft/ft_generic.cc:297-445)
int ref_ft_calc_term_rank(uint32_t nfields, const ft_config* cfg, const ft_term* term, float subterm_proc, double total_docs,
						  double matched_docs, uint32_t npos, const uint32_t* positions, const uint32_t* words_in_fields,
						  const float* avg, float* term_rank, float* bm25_norm, float* position_rank, float* term_len_boost,
						  int* field) {
	try {
		reindexer::FTConfig c(nfields);
		fillConfig(c, cfg);
		reindexer::FtDslOpts opts;
		opts.boost = term->boost;
		opts.termLenBoost = term->term_len_boost;
		opts.fieldsOpts.resize(nfields);
		for (uint32_t f = 0; f < nfields; ++f) {
			opts.fieldsOpts[f].boost = term->field_boosts[f];
			opts.fieldsOpts[f].needSumRank = term->need_sum_rank && term->need_sum_rank[f];
		}
		reindexer::IdRelType rel(1);
		for (uint32_t p = 0; p < npos; ++p) {
			rel.Add(positions[p] & 0xFFFFFF, positions[p] >> 24, 0);
		}
		Stats stats{words_in_fields - nfields, avg, nullptr, nfields};  // vdoc 1 -> words_in_fields[0..nfields)
		reindexer::ft::TermRankInfo inf;
		inf.proc = subterm_proc;
		reindexer::Bm25Calculator<reindexer::Bm25Rx> bm(total_docs, matched_docs, c.bm25Config.bm25k1, c.bm25Config.bm25b);
		auto [rank, fld] = reindexer::ft::calcTermRank(opts, bm, rel, inf, &c, stats);
		*term_rank = rank;
		*bm25_norm = inf.bm25Norm;
		*position_rank = inf.positionRank;
		*term_len_boost = inf.termLenBoost;
		*field = fld;
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// The reference's packed posting stream for one list: IdRelType::packWithoutArrayIdxs (core/ft/idrelset.cc:139-183) applied record
// by record with the state chain of PackedIdRelVec::insert_back (core/ft/idrelset.h:226-260; data_ is private, so the loop is
// repeated here around the reference's own encoder).  Returns the number of bytes, or -1 when `cap` is too small.
int64_t ref_ft_pack_list(const ft_postings* list, uint8_t* out, uint64_t cap) {
	try {
		uint64_t p = 0;
		uint32_t lastId = 0, lastField = 0;
		for (uint32_t i = 0; i < list->ndocs; ++i) {
			const reindexer::IdRelType r = makeRel(*list, i);
			if (p + r.maxpackedsize() > cap) {
				return -1;
			}
			p += r.packWithoutArrayIdxs(out + p, lastId, lastField);
			lastId = r.Id();
			lastField = r.Pos()[0].field();
		}
		return int64_t(p);
	} catch (const std::exception& e) {
		g_err = e.what();
		return -2;
	}
}

}  // extern "C"
