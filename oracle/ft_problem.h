/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Flat description of one ft_fast merge problem, shared by the reference facade (ref_ft_facade.cc) and the C port (ft_port.c).
 * The structs mirror, field for field, the ones of the product ABI (include/rxgpu.h: rxgpu_ft_*), so one ctypes definition
 * serves the oracle and the product in the tests. */
#ifndef ORACLE_FT_PROBLEM_H
#define ORACLE_FT_PROBLEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	uint32_t ndocs;
	const uint32_t* doc_ids;   /* ascending vdoc ids, >= 1 (vdoc 0 is the reference's dummy, mergerimpl.h:122) */
	const uint32_t* pos_begin; /* ndocs + 1 offsets into positions */
	const uint32_t* positions; /* pos | field << 24, ascending within a doc (= PosType::fullPos order with arrayIdx 0) */
} ft_postings;

typedef struct { /* FTFieldConfig, cpp_src/core/ft/config/ftconfig.h:118-146 */
	double bm25_boost, bm25_weight, term_len_boost, term_len_weight, position_boost, position_weight;
} ft_field_config;

typedef struct { /* the FTConfig members the merger reads, ftconfig.h:151-236 */
	uint32_t merge_limit;
	int32_t min_rank;
	double bm25_k1, bm25_b;
	int32_t bm25_type; /* 0 rx, 1 classic, 2 wordCount */
	double distance_boost, distance_weight, full_match_boost;
	uint32_t nfields;
	const ft_field_config* fields;
	double summation_ranks_by_fields_ratio; /* ftconfig.h:210 */
} ft_config;

typedef struct { /* one TermResults: FtDslOpts (ftdsl.h:18-30) + its SubtermResults (querymergedata.h) */
	int32_t op; /* OpType: 1 = OpOr, 2 = OpAnd, 3 = OpNot (core/type_consts.h:187) */
	float boost;
	float term_len_boost;
	const float* field_boosts; /* nfields */
	uint32_t nsubterms;
	const uint32_t* postings; /* indexes into the lists array */
	const float* procs;
	const uint8_t* need_sum_rank; /* nfields flags (FtDslFieldOpts::needSumRank) or NULL */
	const uint8_t* suppressed;    /* nsubterms flags (SubtermResults::Suppressed) or NULL */
	uint32_t nsynonyms;           /* PhraseOrTerm::SynonymsIds */
	const uint32_t* synonym_ids;
	int32_t phrase_num;           /* FtDslOpts::phraseNum: consecutive terms with the same non-zero number form a phrase */
	int32_t distance;             /* FtDslOpts::distance */
} ft_term;

typedef struct { /* ft::Synonym, querymergedata.h:168-188 */
	uint32_t nterms;
	const ft_term* terms;
} ft_synonym;

typedef struct { /* ft::MergeInfo, ft_fast/phrasemerger.h:57-62 */
	int32_t id;
	float proc;
	uint8_t field;
	uint8_t normalized_proc;
} ft_merge_info;

#ifdef __cplusplus
}
#endif
#endif
