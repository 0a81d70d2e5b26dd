// TEST INFRASTRUCTURE.  Compiles the product's ft_fast adapter (reindexer_b200/host/gpu_ft_merge.h) against the reference's own
// headers and runs it and the reference's ft::Merger<IdCont, ft::MergeData, OffsetT>::Merge side by side on the SAME
// ft::QueryMergeData built from the reference's own containers (IdRelVec and PackedIdRelVec), exactly as Selector::Process /
// mergeResults would (cpp_src/core/ft/ft_fast/selecterimpl.h:609-645) -- then diffs ft::MergeData entry by entry (ids, order, uint8
// ranks, field).  Built by tests/cpp/Makefile only where /root/reference exists; the binary travels to the GPU box.
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

#include "gpu_ft_merge.h"
#include "core/rdxcontext.h"

namespace {
int g_nonEmpty = 0, g_runs = 0;

struct Stats {
	std::vector<uint32_t> words;  // [docs][fields]
	std::vector<float> avg;
	std::vector<uint8_t> removed;
	uint32_t nfields;
	bool DocRemoved(uint32_t vdoc) const noexcept { return removed[vdoc]; }
	size_t NumWordsInField(uint32_t vdoc, uint32_t f) const noexcept { return words[size_t(vdoc) * nfields + f]; }
	float AvgWordsCount(uint32_t f) const noexcept { return avg[f]; }
};

struct Problem {
	uint32_t totalDocs, nfields;
	Stats stats;
	std::vector<std::vector<reindexer::IdRelType>> lists;
	struct Term {
		OpType op;
		float boost, termLenBoost;
		std::vector<float> fieldBoosts;
		std::vector<bool> needSum;
		std::vector<std::pair<size_t, float>> subterms;  // (list, proc)
		std::vector<size_t> synonymIds;
		int phraseNum = 0, distance = 0;
	};
	std::vector<Term> terms;
	std::vector<std::vector<Term>> synonyms;  // multi-word synonyms (ft::Synonym)
	reindexer::FTConfig cfg;
	std::vector<uint8_t> excluded;
	explicit Problem(uint32_t nf) : cfg(nf) {}
};

Problem makeProblem(uint32_t seed, uint32_t totalDocs, uint32_t nfields, uint32_t nterms, uint32_t mergeLimit, double sumRatio) {
	std::mt19937 rng(seed);
	auto uni = [&](uint32_t lo, uint32_t hi) { return lo + rng() % (hi - lo + 1); };
	Problem p(nfields);
	p.totalDocs = totalDocs;
	p.nfields = nfields;
	p.stats.nfields = nfields;
	p.stats.words.assign(size_t(totalDocs) * nfields, 0);
	p.stats.removed.assign(totalDocs, 0);
	p.stats.avg.assign(nfields, 0.f);
	for (uint32_t d = 1; d < totalDocs; ++d) {
		p.stats.removed[d] = (rng() % 23) == 0;
		for (uint32_t f = 0; f < nfields; ++f) {
			p.stats.words[size_t(d) * nfields + f] = uni(3, 40);
			p.stats.avg[f] += float(p.stats.words[size_t(d) * nfields + f]) / float(totalDocs - 1);
		}
	}
	p.excluded.assign(totalDocs, 0);
	for (uint32_t d = 1; d < totalDocs; ++d) {
		p.excluded[d] = (rng() % 31) == 0;
	}
	p.cfg.mergeLimit = mergeLimit;
	p.cfg.summationRanksByFieldsRatio = sumRatio;
	p.cfg.bm25Config.bm25Type = reindexer::FTConfig::Bm25Config::Bm25Type(seed % 3);
	const float procsPool[] = {100.f, 90.f, 85.f, 80.f, 72.f, 65.f, 57.f, 50.f};
	for (uint32_t t = 0; t < nterms; ++t) {
		Problem::Term term;
		term.op = t == 0 ? OpOr : (rng() % 4 == 0 ? OpAnd : (rng() % 5 == 0 ? OpNot : OpOr));
		term.boost = (rng() % 2) ? 1.f : 0.7f;
		term.termLenBoost = (rng() % 2) ? 1.f : 0.8f;
		for (uint32_t f = 0; f < nfields; ++f) {
			term.fieldBoosts.push_back((rng() % 3) ? 1.f : 1.5f);
			term.needSum.push_back(sumRatio > 0 && (rng() % 3) != 0);
		}
		const uint32_t nsub = uni(1, 3);
		for (uint32_t s = 0; s < nsub; ++s) {
			std::vector<reindexer::IdRelType> list;
			const uint32_t step = uni(2, 9) * (s + 1);
			for (uint32_t d = 1 + rng() % step; d < totalDocs; d += 1 + rng() % step) {
				reindexer::IdRelType r(d);
				std::vector<std::pair<uint32_t, uint32_t>> pos;  // (field, pos) ascending like the index builds them
				const uint32_t np = uni(1, 4);
				for (uint32_t i = 0; i < np; ++i) {
					const uint32_t f = rng() % nfields;
					pos.emplace_back(f, rng() % std::max<uint32_t>(1, p.stats.words[size_t(d) * nfields + f]));
				}
				std::sort(pos.begin(), pos.end());
				pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
				for (auto [f, w] : pos) {
					r.Add(w, f, 0);
				}
				list.emplace_back(std::move(r));
			}
			p.lists.emplace_back(std::move(list));
			term.subterms.emplace_back(p.lists.size() - 1, procsPool[(s * 3 + t) % 8]);  // distinct procs per term
		}
		p.terms.emplace_back(std::move(term));
	}
	// every third problem: the first two terms form a phrase (dense enough lists make adjacent positions common)
	if (seed % 3 == 0 && p.terms.size() >= 2) {
		for (int k = 0; k < 2; ++k) {
			p.terms[k].phraseNum = 1;
			p.terms[k].distance = k ? int(uni(1, 6)) : 0;
			p.terms[k].op = p.terms[0].op;
		}
	}
	// multi-word synonyms on every other problem: 2-3 terms each, attached to one or two non-NOT query parts; one of their subterms may
	// repeat a word (= posting list) of the query, which QueryMergeData::SupressDuplicatesInSynonyms then suppresses
	for (uint32_t y = 0; y < (seed % 2 ? 1 + seed % 3 % 2 : 0); ++y) {
		std::vector<Problem::Term> syn;
		const uint32_t nst = uni(2, 3);
		for (uint32_t t = 0; t < nst; ++t) {
			Problem::Term term;
			term.op = OpOr;
			term.boost = 1.f;
			term.termLenBoost = 1.f;
			term.fieldBoosts.assign(nfields, 1.f);
			term.needSum.assign(nfields, false);
			std::vector<reindexer::IdRelType> list;
			const uint32_t step = uni(2, 5);
			for (uint32_t d = 1 + rng() % step; d < totalDocs; d += 1 + rng() % step) {
				reindexer::IdRelType r(d);
				const uint32_t f = rng() % nfields;
				r.Add(rng() % std::max<uint32_t>(1, p.stats.words[size_t(d) * nfields + f]), f, 0);
				list.emplace_back(std::move(r));
			}
			p.lists.emplace_back(std::move(list));
			term.subterms.emplace_back(p.lists.size() - 1, 45.f - 5.f * float(t));
			if (rng() % 3 == 0) {
				term.subterms.emplace_back(p.terms[rng() % p.terms.size()].subterms[0].first, 20.f);
			}
			syn.emplace_back(std::move(term));
		}
		p.synonyms.emplace_back(std::move(syn));
		for (uint32_t n = 0, tries = 0; n < 1 + rng() % 2 && tries < 8; ++tries) {
			auto& host = p.terms[rng() % p.terms.size()];
			if (host.op != OpNot && host.phraseNum == 0 && std::find(host.synonymIds.begin(), host.synonymIds.end(), p.synonyms.size() - 1) == host.synonymIds.end()) {
				host.synonymIds.push_back(p.synonyms.size() - 1);
				++n;
			}
		}
	}
	return p;
}

template <typename IdCont>
void buildCont(const std::vector<reindexer::IdRelType>& src, IdCont& out);
template <>
void buildCont(const std::vector<reindexer::IdRelType>& src, reindexer::IdRelVec& out) {
	for (const auto& r : src) {
		out.emplace_back(r);
	}
}
template <>
void buildCont(const std::vector<reindexer::IdRelType>& src, reindexer::PackedIdRelVec& out) {
	std::vector<reindexer::IdRelType> tmp(src);  // insert_back wants mutable records
	out.insert_back(tmp.begin(), tmp.end());
}

template <typename IdCont>
reindexer::ft::QueryMergeData<IdCont> buildQuery(const Problem& p, const std::vector<IdCont>& conts) {
	reindexer::ft::QueryMergeData<IdCont> q;
	auto makeTerm = [&](const Problem::Term& t) {
		reindexer::FtDSLEntry e;
		e.Opts().op = t.op;
		e.Opts().boost = t.boost;
		e.Opts().termLenBoost = t.termLenBoost;
		e.Opts().phraseNum = t.phraseNum;
		e.Opts().distance = t.distance;
		e.Opts().fieldsOpts.resize(p.nfields);
		for (uint32_t f = 0; f < p.nfields; ++f) {
			e.Opts().fieldsOpts[f].boost = t.fieldBoosts[f];
			e.Opts().fieldsOpts[f].needSumRank = t.needSum[f];
		}
		reindexer::ft::TermResults<IdCont> tr(std::move(e));
		for (const auto& [li, proc] : t.subterms) {
			reindexer::WordIdType wid;
			wid.SetID(int32_t(li));  // one word per posting list
			tr.AddSubterm(conts[li], "w", wid, proc);
		}
		q.totalORVids += tr.MaxVDocs();
		return tr;
	};
	for (const auto& syn : p.synonyms) {
		reindexer::ft::Synonym<IdCont> s;
		for (const auto& t : syn) {
			s.AddTerm(makeTerm(t));
		}
		q.synonyms.emplace_back(std::move(s));
	}
	reindexer::ft::PhraseResults<IdCont> nextPhrase;  // grouped like selecterimpl.h:546-566
	int curPhrase = 0;
	for (const auto& t : p.terms) {
		if (t.phraseNum) {
			if (nextPhrase.NumTerms() && curPhrase != t.phraseNum) {
				q.queryParts.emplace_back(std::move(nextPhrase));
				nextPhrase.clear();
			}
			curPhrase = t.phraseNum;
			nextPhrase.Add(makeTerm(t));
			continue;
		}
		if (nextPhrase.NumTerms()) {
			q.queryParts.emplace_back(std::move(nextPhrase));
			nextPhrase.clear();
		}
		q.queryParts.emplace_back(makeTerm(t));
		for (const size_t id : t.synonymIds) {
			q.queryParts.back().AddSynonymId(id);
		}
	}
	if (nextPhrase.NumTerms()) {
		q.queryParts.emplace_back(std::move(nextPhrase));
		nextPhrase.clear();
	}
	q.SupressDuplicatesInSynonyms();  // selecterimpl.h:606
	return q;
}

template <typename IdCont>
bool runCase(const Problem& p, const char* name) {
	std::vector<IdCont> conts(p.lists.size());
	for (size_t i = 0; i < p.lists.size(); ++i) {
		buildCont(p.lists[i], conts[i]);
	}
	reindexer::FtMergeStatuses::Statuses excluded(p.totalDocs, false);
	for (uint32_t d = 0; d < p.totalDocs; ++d) {
		if (p.excluded[d]) {
			excluded.set(d);
		}
	}
	reindexer::ft::GpuFtMerger<IdCont> gpu(p.totalDocs, p.nfields, p.stats);
	bool ok = true;
	for (auto rst : {reindexer::RankSortType::RankAndID, reindexer::RankSortType::IDOnly}) {
		auto qRef = buildQuery(p, conts);
		auto qGpu = buildQuery(p, conts);
		reindexer::FTConfig cfg = p.cfg;
		reindexer::RdxContext ctx;
		auto excludedRef = excluded;  // Merge consumes it (swap into the restricting mask, mergerimpl.h:328)
		const auto maxMerged = std::min<uint64_t>(cfg.mergeLimit, qRef.totalORVids);
		std::vector<reindexer::ft::MergeInfo> ref;
		auto call = [&](auto& merger) {
			switch (cfg.bm25Config.bm25Type) {
				case reindexer::FTConfig::Bm25Config::Bm25Type::classic:
					return merger.template Merge<reindexer::Bm25Classic>(qRef, rst, p.stats);
				case reindexer::FTConfig::Bm25Config::Bm25Type::wordCount:
					return merger.template Merge<reindexer::TermCount>(qRef, rst, p.stats);
				default:
					return merger.template Merge<reindexer::Bm25Rx>(qRef, rst, p.stats);
			}
		};
		if (maxMerged < 0xFFFF) {  // Selector::Process, selecterimpl.h:637-644
			reindexer::ft::Merger<IdCont, reindexer::ft::MergeData, uint16_t> m(p.totalDocs, &cfg, excludedRef, p.nfields, 5, false, ctx);
			auto r = call(m);
			ref.swap(r);
		} else {
			reindexer::ft::Merger<IdCont, reindexer::ft::MergeData, uint32_t> m(p.totalDocs, &cfg, excludedRef, p.nfields, 5, false, ctx);
			auto r = call(m);
			ref.swap(r);
		}
		if (!reindexer::ft::GpuFtMerger<IdCont>::Mergeable(qGpu)) {
			std::printf("%s: not mergeable?\n", name);
			return false;
		}
		const auto res = gpu.Merge(qGpu, rst, excluded, p.cfg);
		bool same = ref.size() == res.size();
		for (size_t i = 0; same && i < ref.size(); ++i) {
			same = ref[i].id.ToNumber() == res[i].id.ToNumber() && ref[i].normalizedProc == res[i].normalizedProc && ref[i].field == res[i].field &&
				   ref[i].proc == res[i].proc;
		}
		if (!same) {
			std::printf("%s rst %d: reference %zu docs, device %zu docs -> MISMATCH\n", name, int(rst), ref.size(), res.size());
		}
		ok = ok && same;
		g_nonEmpty += !ref.empty();
		++g_runs;
	}
	return ok;
}

}  // namespace

int main() {
	int bad = 0, cases = 0;
	for (uint32_t seed = 0; seed < 24; ++seed) {
		const uint32_t nfields = 1 + seed % 3, nterms = 1 + seed % 4;
		const Problem p = makeProblem(seed, 300 + 97 * seed, nfields, nterms, seed % 4 == 1 ? 25 : 20000, seed % 3 == 2 ? 0.5 : 0.0);
		const bool a = runCase<reindexer::IdRelVec>(p, "IdRelVec");
		const bool b = runCase<reindexer::PackedIdRelVec>(p, "PackedIdRelVec");
		bad += !a + !b;
		cases += 2;
	}
	bad += g_nonEmpty * 10 < g_runs * 8;  // most results must be non-empty for the comparison to mean something
	std::printf("ft merge adapter: %d of %d merges non-empty; ", g_nonEmpty, g_runs);
	std::printf("ft merge adapter: %d cases (IdRelVec and PackedIdRelVec, AND/OR/NOT, preselect cut, all bm25 variants, summation of field ranks, multi-word synonyms with suppressed subterms, phrases): "
				"%s\n",
				cases, bad ? "MISMATCH" : "MATCH MATCH MATCH MATCH");
	return bad;
}
