/* ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle_port.h.  Never part of the product path.
 *
 * Plain-C restatement of the reference's ft_fast merge for term-only queries (no phrases, no multi-word synonyms -- those
 * query parts are built by the CPU-side DSL/variant code that is out of scope):
 *   Bm25Rx / Bm25Classic / TermCount          cpp_src/core/ft/bm25.h:8-68
 *   FTFieldConfig::bound / pos2rank            cpp_src/core/ft/config/ftconfig.h:127-146
 *   calcTermRank                               cpp_src/core/ft/ft_fast/phrasemergerimpl.h:13-91
 *   PositionsDistance                          cpp_src/core/ft/ft_fast/mergerimpl.h:20-37
 *   Merger::Merge                              mergerimpl.h:466-566   (init merger.h:64-90)
 *   mergeSimple :194-250, mergeTerm :107-192, switchToNextWord merger.h:220-228
 *   buildRestrictingBitmask :326-384 (calcTermBitmask :252-274, excludeTermFromBitmask :276-287)
 *   estimateNumDocsInMerge merger.h:239-267, preselectMostRelevantDocs :386-464 (calcTermScores :289-324)
 *   addFullMatchBoost merger.h:100-109, postProcessResults merger.h:111-155
 * Arithmetic follows the reference expression by expression (which operands are float, which double, where a double is
 * rounded to float); built with -ffp-contract=off.  Pinned by tests/test_ft_oracle_pin.py against oracle/_ref (the reference's
 * own merger) and the committed fixtures, including the values FTGenericApi.DebugInfo pins. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ft_problem.h"

typedef struct {
	double k1, b, idf;
	int type;
} bm25_t;

static bm25_t bm25_make(int type, double totalDocCount, double matchedDocCount, double k1, double b) {
	bm25_t r = {k1, b, 0.0, type};
	if (type == 0) { /* Bm25Rx::IDF bm25.h:21-27 */
		double f = log((totalDocCount - matchedDocCount + 1) / matchedDocCount) / log(1 + totalDocCount);
		if (f < 0.2) {
			f = 0.2;
		}
		r.idf = f;
	} else if (type == 1) { /* Bm25Classic::IDF bm25.h:51-53 */
		r.idf = log(totalDocCount / (matchedDocCount + 1)) + 1;
	}
	return r;
}
static double bm25_get(const bm25_t* c, double termCountInDoc, double wordsInDoc, double avgDocLen) {
	if (c->type == 2) {
		return termCountInDoc; /* TermCount bm25.h:60-68 */
	}
	const double termFreq = c->type == 0 ? termCountInDoc : termCountInDoc / wordsInDoc;
	return c->idf * termFreq * (c->k1 + 1.0) / (termFreq + c->k1 * (1.0 - c->b + c->b * wordsInDoc / avgDocLen));
}

/* ftconfig.h:146: float bound(float k, float weight, float boost) { return (1.0 - weight) + k * boost * weight; } */
static float bound_f(float k, float weight, float boost) { return (float)((1.0 - weight) + k * boost * weight); }
static float pos2rank(unsigned pos) { /* ftconfig.h:127-144 */
	if (pos <= 10) {
		return (float)(1.0 - (pos / 100.0));
	}
	if (pos <= 100) {
		return (float)(0.9 - (pos / 1000.0));
	}
	if (pos <= 1000) {
		return (float)(0.8 - (pos / 10000.0));
	}
	if (pos <= 10000) {
		return (float)(0.7 - (pos / 100000.0));
	}
	if (pos <= 100000) {
		return (float)(0.6 - (pos / 1000000.0));
	}
	return 0.5f;
}

typedef struct {
	uint32_t total_docs, nfields;
	const uint32_t* words;
	const float* avg;
	const uint8_t* removed;
} stats_t;

/* phrasemergerimpl.h:13-91, incl. the summation of the other fields' ranks (summationRanksByFieldsRatio > 0, needSumRank fields) */
static float calc_term_rank(const ft_term* term, const bm25_t* bm, uint32_t doc, const uint32_t* pos, uint32_t npos, float proc,
							const ft_config* cfg, const stats_t* st, uint8_t* fieldOut) {
	uint8_t fieldWithMaxRank = 0;
	float termRank = 0.f;
	float ranksInFields[64];
	uint32_t nranks = 0;
	int needToSumWinner = 0;
	const int needSumRanks = cfg->summation_ranks_by_fields_ratio > 0.0;
	for (uint32_t idx = 0; idx < npos;) {
		const unsigned f = pos[idx] >> 24;
		const uint32_t fieldBegin = idx;
		++idx;
		while (idx < npos && (pos[idx] >> 24) == f) {
			++idx;
		}
		if (term->field_boosts[f] == 0.f) {
			continue;
		}
		const ft_field_config* fc = &cfg->fields[f];
		const uint32_t wordsInField = idx - fieldBegin;
		const float bm25 = (float)bm25_get(bm, (double)wordsInField, (double)st->words[(size_t)doc * st->nfields + f], (double)st->avg[f]);
		const float normBm25 = bound_f(bm25, (float)fc->bm25_weight, (float)fc->bm25_boost);
		const float positionRank = bound_f(pos2rank(pos[fieldBegin] & 0xFFFFFF), (float)fc->position_weight, (float)fc->position_boost);
		const float termLenBoost = bound_f(term->term_len_boost, (float)fc->term_len_weight, (float)fc->term_len_boost);
		const float termRankTmp = term->field_boosts[f] * normBm25 * termLenBoost * positionRank;
		const int needSum = term->need_sum_rank && term->need_sum_rank[f];
		if (termRankTmp > termRank) {
			fieldWithMaxRank = (uint8_t)f;
			termRank = termRankTmp;
			needToSumWinner = needSum;
		}
		if (needSum) {
			ranksInFields[nranks++] = termRankTmp;
		}
	}
	if (termRank > 0.0 && needSumRanks) { /* :70-78: ranks descending, geometric weights k, k^2, ...; the winner is not added twice */
		for (uint32_t i = 1; i < nranks; ++i) {
			const float v = ranksInFields[i];
			uint32_t j = i;
			for (; j > 0 && ranksInFields[j - 1] < v; --j) {
				ranksInFields[j] = ranksInFields[j - 1];
			}
			ranksInFields[j] = v;
		}
		float k = (float)cfg->summation_ranks_by_fields_ratio;
		for (uint32_t i = needToSumWinner ? 1 : 0; i < nranks; ++i) {
			termRank += (k * ranksInFields[i]);
			k = (float)(k * cfg->summation_ranks_by_fields_ratio);
		}
	}
	*fieldOut = fieldWithMaxRank;
	return term->boost * proc * termRank;
}

/* mergerimpl.h:20-37.  PosType::fullPos() returns the LOW 32 bits of pos | arrayIdx << 28 | field << 56 (idrelset.h:24),
 * i.e. the word position only (arrayIdx is 0 here): the two-pointer walk advances by position even across fields, while
 * fullField() (= arrayIdx | field << 28, truncated) decides whether a pair counts. */
static unsigned positions_distance(const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb) {
	unsigned res = 0xFFFFFFFFu;
	uint32_t i = 0, j = 0;
	while (i < na && j < nb) {
		const uint32_t pa = a[i] & 0xFFFFFF, pb = b[j] & 0xFFFFFF;
		const int sign = pa > pb;
		if ((a[i] >> 24) == (b[j] >> 24)) {
			const unsigned dst = sign ? pa - pb : pb - pa;
			if (dst < res) {
				res = dst;
				if (res <= 1) {
					break;
				}
			}
		}
		if (sign) {
			j++;
		} else {
			i++;
		}
	}
	return res == 0xFFFFFFFFu ? 0 : res;
}

typedef struct {
	const uint32_t* last_pos;
	uint32_t last_n;
	const uint32_t* next_pos;
	uint32_t next_n;
	float rank;
	uint16_t lastTermCounted, termsCounter;
} md_ext;

typedef struct {
	uint32_t list;
	float proc;
} sub_t;
static int sub_cmp(const void* a, const void* b) {
	const float x = ((const sub_t*)a)->proc, y = ((const sub_t*)b)->proc;
	return (x < y) - (x > y); /* proc descending (SortSubterms, querymergedata.h); ties keep input order below */
}
static void sort_subterms(sub_t* s, uint32_t n) { /* stable insertion sort: tiny n */
	for (uint32_t i = 1; i < n; ++i) {
		sub_t v = s[i];
		uint32_t j = i;
		while (j > 0 && sub_cmp(&s[j - 1], &v) > 0) {
			s[j] = s[j - 1];
			--j;
		}
		s[j] = v;
	}
}

static void post_process(ft_merge_info* md, size_t* n_io, const ft_config* cfg, int rank_sort_type) { /* merger.h:111-155 */
	size_t n = *n_io;
	float maxProc = 0.f;
	for (size_t i = 0; i < n; ++i) {
		maxProc = md[i].proc > maxProc ? md[i].proc : maxProc;
	}
	const float scalingFactor = maxProc > 255 ? (float)(255.0 / maxProc) : 1.0f;
	const float minProc = (float)cfg->min_rank;
	size_t passed = n;
	while (passed > 0 && md[passed - 1].proc < minProc) {
		passed--;
	}
	for (size_t i = 0; i + 1 < passed; i++) {
		if (md[i].proc < minProc) {
			md[i] = md[passed - 1];
			passed--;
			while (passed > i && md[passed - 1].proc < minProc) {
				passed--;
			}
		}
	}
	n = passed;
	for (size_t i = 0; i < n; ++i) {
		md[i].normalized_proc = (uint8_t)(md[i].proc * scalingFactor);
		md[i].proc = md[i].normalized_proc;
	}
	if (rank_sort_type == 0 || rank_sort_type == 4) { /* RankOnly / IDAndPositions: sort by normalizedProc desc (pdqsort, unstable) */
		for (size_t i = 1; i < n; ++i) {              /* the port uses a stable sort; callers compare as multisets per rank */
			ft_merge_info v = md[i];
			size_t j = i;
			while (j > 0 && md[j - 1].normalized_proc < v.normalized_proc) {
				md[j] = md[j - 1];
				--j;
			}
			md[j] = v;
		}
	}
	*n_io = n;
}

#define BIT_GET(m, i) (((m)[(i) >> 6] >> ((i) & 63)) & 1ull)
#define BIT_SET(m, i) ((m)[(i) >> 6] |= 1ull << ((i) & 63))
#define BIT_CLR(m, i) ((m)[(i) >> 6] &= ~(1ull << ((i) & 63)))

int port_ft_merge(uint32_t total_docs, uint32_t nfields, const uint32_t* words, const float* avg, const uint8_t* removed,
				  const uint8_t* excluded, uint32_t nlists, const ft_postings* lists, const ft_config* cfg, uint32_t nterms,
				  const ft_term* terms, int rank_sort_type, uint64_t max_out, ft_merge_info* out, uint64_t* out_n, int64_t* merge_ns) {
	(void)nlists;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	*out_n = 0;
	if (nterms == 0 || (nterms == 1 && terms[0].op == 3) || total_docs == 0) { /* Empty(), mergerimpl.h:472 */
		return 0;
	}
	const stats_t st = {total_docs, nfields, words, avg, removed};
	uint64_t totalORVids = 0;
	for (uint32_t t = 0; t < nterms; ++t) {
		for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
			totalORVids += lists[terms[t].postings[s]].ndocs;
		}
	}
	const uint32_t maxMerged = (uint32_t)(cfg->merge_limit < totalORVids ? cfg->merge_limit : totalORVids); /* init(), merger.h:66-67 */
	const int simple = nterms == 1 && terms[0].op != 3;
	const int trivial = simple && terms[0].nsubterms == 1;

	ft_merge_info* md = (ft_merge_info*)malloc(((size_t)maxMerged + 1) * sizeof(ft_merge_info));
	md_ext* ext = simple ? NULL : (md_ext*)calloc((size_t)maxMerged + 1, sizeof(md_ext));
	uint32_t* idoffsets = NULL;
	if (!trivial) {
		idoffsets = (uint32_t*)malloc((size_t)total_docs * sizeof(uint32_t));
		for (uint32_t i = 0; i < total_docs; ++i) {
			idoffsets[i] = maxMerged;
		}
	}
	size_t n = 0;

	/* SortSubterms */
	sub_t** subs = (sub_t**)malloc(nterms * sizeof(sub_t*));
	for (uint32_t t = 0; t < nterms; ++t) {
		subs[t] = (sub_t*)malloc((terms[t].nsubterms + 1) * sizeof(sub_t));
		for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
			subs[t][s].list = terms[t].postings[s];
			subs[t][s].proc = terms[t].procs[s];
		}
		sort_subterms(subs[t], terms[t].nsubterms);
	}

	if (simple) { /* mergeSimple, mergerimpl.h:194-250 */
		const ft_term* term = &terms[0];
		for (uint32_t s = 0; s < term->nsubterms; ++s) {
			const ft_postings* l = &lists[subs[0][s].list];
			const bm25_t bm = bm25_make(cfg->bm25_type, (double)(total_docs - 1), (double)l->ndocs, cfg->bm25_k1, cfg->bm25_b);
			for (uint32_t i = 0; i < l->ndocs; ++i) {
				const uint32_t doc = l->doc_ids[i];
				if ((excluded && excluded[doc]) || (removed && removed[doc])) {
					continue;
				}
				const int added = idoffsets && idoffsets[doc] != maxMerged;
				if (!added && n >= maxMerged) {
					continue;
				}
				uint8_t field;
				const float rank = calc_term_rank(term, &bm, doc, l->positions + l->pos_begin[i], l->pos_begin[i + 1] - l->pos_begin[i],
												  subs[0][s].proc, cfg, &st, &field);
				if (rank == 0.f) {
					continue;
				}
				if (!added) {
					md[n].id = (int32_t)doc;
					md[n].proc = rank;
					md[n].field = field;
					md[n].normalized_proc = 0;
					if (idoffsets) {
						idoffsets[doc] = (uint32_t)n;
					}
					n++;
				} else {
					ft_merge_info* m = &md[idoffsets[doc]];
					if (m->proc < rank) {
						m->proc = rank;
						m->field = field;
					}
				}
			}
		}
		for (size_t i = 0; i < n; ++i) { /* addFullMatchBoost(1) */
			if (words[(size_t)md[i].id * nfields + md[i].field] == 1) {
				md[i].proc = (float)(md[i].proc * cfg->full_match_boost);
			}
		}
	} else {
		const size_t mwords = ((size_t)total_docs + 63) / 64;
		uint64_t* mask = (uint64_t*)calloc(mwords, 8);
		uint64_t* tmask = (uint64_t*)calloc(mwords, 8);
		/* buildRestrictingBitmask :326-384: mask = ~docsExcluded */
		for (uint32_t i = 0; i < total_docs; ++i) {
			if (!(excluded && excluded[i])) {
				BIT_SET(mask, i);
			}
		}
		for (uint32_t t = 0; t < nterms; ++t) {
			if (terms[t].op != 2) {
				continue;
			}
			memset(tmask, 0, mwords * 8);
			int allPositive = 1;
			for (uint32_t f = 0; f < nfields; ++f) {
				allPositive &= terms[t].field_boosts[f] != 0.f;
			}
			for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
				const ft_postings* l = &lists[subs[t][s].list];
				for (uint32_t i = 0; i < l->ndocs; ++i) {
					const uint32_t doc = l->doc_ids[i];
					if (BIT_GET(tmask, doc)) {
						continue;
					}
					int relevant = allPositive;
					for (uint32_t p = l->pos_begin[i]; !relevant && p < l->pos_begin[i + 1]; ++p) {
						relevant = terms[t].field_boosts[l->positions[p] >> 24] != 0.f; /* checkFieldsRelevance */
					}
					if (relevant) {
						BIT_SET(tmask, doc);
					}
				}
			}
			for (size_t w = 0; w < mwords; ++w) {
				mask[w] &= tmask[w];
			}
		}
		for (uint32_t t = 0; t < nterms; ++t) {
			if (terms[t].op != 3) {
				continue;
			}
			for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
				const ft_postings* l = &lists[subs[t][s].list];
				for (uint32_t i = 0; i < l->ndocs; ++i) {
					BIT_CLR(mask, l->doc_ids[i]);
				}
			}
		}
		/* estimateNumDocsInMerge merger.h:239-267 */
		uint64_t estOr = 0, estAnd = UINT64_MAX;
		for (uint32_t t = 0; t < nterms; ++t) {
			if (terms[t].op == 3) {
				continue;
			}
			uint64_t nd = 0;
			for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
				nd += lists[subs[t][s].list].ndocs;
			}
			if (terms[t].op == 2) {
				estAnd = nd < estAnd ? nd : estAnd;
			} else {
				estOr += nd;
			}
		}
		uint64_t est = estOr < estAnd ? estOr : estAnd;
		est = est < total_docs ? est : total_docs;
		uint64_t pop = 0;
		for (size_t w = 0; w < mwords; ++w) {
			pop += (uint64_t)__builtin_popcountll(mask[w]);
		}
		int needCheckRemoved = 1;
		if (est > cfg->merge_limit && total_docs > cfg->merge_limit && pop > cfg->merge_limit) {
			/* preselectMostRelevantDocs :386-464 */
			uint16_t* score = (uint16_t*)calloc(total_docs, sizeof(uint16_t));
			for (uint32_t t = 0; t < nterms; ++t) {
				if (terms[t].op == 3) {
					continue;
				}
				memset(tmask, 0, mwords * 8); /* calcTermScores :289-324 */
				const float fieldsBoost = terms[t].field_boosts[0];
				int allSame = 1;
				for (uint32_t f = 0; f < nfields; ++f) {
					allSame &= terms[t].field_boosts[f] == fieldsBoost;
				}
				for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
					const ft_postings* l = &lists[subs[t][s].list];
					for (uint32_t i = 0; i < l->ndocs; ++i) {
						const uint32_t doc = l->doc_ids[i];
						if (!BIT_GET(mask, doc)) {
							continue;
						}
						float maxBoost = fieldsBoost;
						if (!allSame) {
							maxBoost = 0.f;
							for (uint32_t p = l->pos_begin[i]; p < l->pos_begin[i + 1]; ++p) {
								const float b = terms[t].field_boosts[l->positions[p] >> 24];
								maxBoost = b > maxBoost ? b : maxBoost;
							}
						}
						if (maxBoost > 0.0f && !BIT_GET(tmask, doc)) {
							const float proc = subs[t][s].proc * maxBoost * terms[t].boost;
							uint16_t proc16 = (uint16_t)proc;
							proc16 = proc16 < 65535 / 4 ? proc16 : 65535 / 4;
							const uint16_t room = (uint16_t)(65535 - score[doc]);
							proc16 = proc16 < room ? proc16 : room;
							score[doc] = (uint16_t)(score[doc] + proc16);
							BIT_SET(tmask, doc);
						}
					}
				}
			}
			size_t* hist = (size_t*)calloc(65536, sizeof(size_t));
			for (uint32_t i = 0; i < total_docs; ++i) {
				if (!BIT_GET(mask, i) || (removed && removed[i])) {
					score[i] = 0;
				}
				hist[score[i]]++;
			}
			size_t minScore = 65535, minScoreDocs = 0, docsTaken = 0;
			for (size_t sc = 65535; sc > 0; sc--) {
				if (docsTaken >= maxMerged) {
					break;
				}
				minScore = sc;
				minScoreDocs = maxMerged - docsTaken;
				docsTaken += hist[sc];
			}
			size_t minScoreDocsTaken = 0;
			for (uint32_t i = 0; i < total_docs; ++i) {
				if (!BIT_GET(mask, i)) {
					continue;
				}
				if (score[i] > minScore) {
					continue;
				} else if (score[i] == minScore && minScoreDocsTaken < minScoreDocs) {
					++minScoreDocsTaken;
					continue;
				}
				BIT_CLR(mask, i);
			}
			needCheckRemoved = 0;
			free(hist);
			free(score);
		}
		/* mergeTerm per query part :107-192 */
		uint16_t qpIdx = 0;
		for (uint32_t t = 0; t < nterms; ++t) {
			if (terms[t].op == 3) {
				continue;
			}
			++qpIdx;
			for (size_t i = 0; i < n; ++i) { /* switchToNextWord merger.h:220-228 */
				if (ext[i].next_n) {
					ext[i].last_pos = ext[i].next_pos;
					ext[i].last_n = ext[i].next_n;
					ext[i].next_n = 0;
					ext[i].rank = 0;
				}
			}
			for (uint32_t s = 0; s < terms[t].nsubterms; ++s) {
				const ft_postings* l = &lists[subs[t][s].list];
				const bm25_t bm = bm25_make(cfg->bm25_type, (double)(total_docs - 1), (double)l->ndocs, cfg->bm25_k1, cfg->bm25_b);
				for (uint32_t i = 0; i < l->ndocs; ++i) {
					const uint32_t doc = l->doc_ids[i];
					if (!BIT_GET(mask, doc)) {
						continue;
					}
					const int added = idoffsets[doc] != maxMerged;
					if (!added && n >= maxMerged) {
						continue;
					}
					if (needCheckRemoved && removed && removed[doc]) {
						continue;
					}
					const uint32_t* pos = l->positions + l->pos_begin[i];
					const uint32_t npos = l->pos_begin[i + 1] - l->pos_begin[i];
					uint8_t field;
					const float rank = calc_term_rank(&terms[t], &bm, doc, pos, npos, subs[t][s].proc, cfg, &st, &field);
					if (rank == 0.f) {
						continue;
					}
					if (!added) {
						md[n].id = (int32_t)doc;
						md[n].proc = rank;
						md[n].field = field;
						md[n].normalized_proc = 0;
						memset(&ext[n], 0, sizeof(md_ext));
						ext[n].next_pos = pos;
						ext[n].next_n = npos;
						ext[n].rank = rank;
						idoffsets[doc] = (uint32_t)n;
						n++;
					}
					md_ext* e = &ext[idoffsets[doc]];
					if (e->lastTermCounted < qpIdx) { /* InreaseTermsCounter merger.h:31-36 */
						++e->termsCounter;
						e->lastTermCounted = qpIdx;
					}
					if (added) {
						ft_merge_info* m = &md[idoffsets[doc]];
						unsigned distance = positions_distance(e->last_pos, e->last_n, pos, npos);
						if (distance < 1) {
							distance = 1;
						}
						const float normDist = bound_f((float)(1.0 / (float)distance), (float)cfg->distance_weight, (float)cfg->distance_boost);
						const float finalRank = normDist * rank;
						if (finalRank > e->rank) {
							m->proc -= e->rank;
							m->proc += finalRank;
							e->next_pos = pos;
							e->next_n = npos;
							e->rank = finalRank;
						}
					}
				}
			}
		}
		/* canBeBoostedByFullMatch :533-537 + addFullMatchBoost(QueryLength) */
		for (size_t i = 0; i < n; ++i) {
			if (ext[i].termsCounter == nterms && words[(size_t)md[i].id * nfields + md[i].field] == nterms) {
				md[i].proc = (float)(md[i].proc * cfg->full_match_boost);
			}
		}
		free(mask);
		free(tmask);
	}
	post_process(md, &n, cfg, rank_sort_type);
	*out_n = n;
	for (size_t i = 0; i < n && i < max_out; ++i) {
		out[i] = md[i];
	}
	for (uint32_t t = 0; t < nterms; ++t) {
		free(subs[t]);
	}
	free(subs);
	free(md);
	free(ext);
	free(idoffsets);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (merge_ns) {
		*merge_ns = (int64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec);
	}
	return 0;
}
