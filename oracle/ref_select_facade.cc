// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product (librxgpu.so).
//
// Pins the product's restatement of HnswIndexBase<Map>::select / selectRaw post-processing (reindexer_b200/host/knn_select.h:
// selectPostprocess) to the REFERENCE'S OWN CODE.  hnsw_index.cc cannot be compiled standalone (it pulls the whole Index / payload /
// namespace stack), so oracle/Makefile extracts, at build time and into oracle/_ref/gen/ (never committed, never copied into the repo),
// the text of three reference functions where it lies:
//   select_drain.inc    HnswIndexBase<Map>::select, the block between `auto knnRes = search(key, params);` and the IdSetPlain
//                       construction                                   cpp_src/core/index/float_vector/hnsw_index.cc:234-282
//   remove_over_k.inc   HnswIndexBase<Map>::removeOverK body             cpp_src/core/index/float_vector/hnsw_index.cc:195-202
//   remove_dup.inc      FloatVectorIndex::removeDuplicateRowId body      cpp_src/core/index/float_vector/float_vector_index.h:142-159
// and this file compiles them verbatim behind duck-typed stand-ins for the few names they use (KnnCtx::NeedSort, Opts(), params.K() /
// Radius() / BruteForce() / Hnsw(), metric_).  All containers and value types are the reference's own (h_vector, base_idset, RankT,
// FloatVectorId, SearchResultQueue, fast_hash_set).
#include <cstdint>
#include <optional>
#include <type_traits>
#include <vector>

#include "core/enums.h"
#include "core/idset/idset.h"
#include "core/index/float_vector/float_vector_id.h"
#include "core/index/float_vector/hnswlib/hnsw_interface.h"
#include "core/rank_t.h"
#include "estl/fast_hash_set.h"
#include "estl/h_vector.h"
#include "tools/assertrx.h"

namespace hnswlib {
class BruteforceSearch;  // only named in a std::is_same_v of the extracted text
}
namespace faiss {
using idx_t = int64_t;  // removeDuplicateRowId has an overload for FAISS ids
}

namespace reindexer {
namespace {

struct Params {
	std::optional<size_t> k;
	std::optional<float> radius;
	std::optional<size_t> K() const noexcept { return k; }
	std::optional<float> Radius() const noexcept { return radius; }
	const Params& BruteForce() const noexcept { return *this; }
	const Params& Hnsw() const noexcept { return *this; }
};
struct Ctx {
	bool needSort;
	bool NeedSort() const noexcept { return needSort; }
};
struct FvOpts {
	std::optional<float> radius;
	std::optional<float> Radius() const noexcept { return radius; }
};
struct IndexOpts {
	bool isArray;
	FvOpts fv;
	bool IsArray() const noexcept { return isArray; }
	const FvOpts& FloatVector() const noexcept { return fv; }
};

template <typename Map>
struct Harness {
	VectorMetric metric_;
	IndexOpts opts_;
	const IndexOpts& Opts() const noexcept { return opts_; }

	template <typename Ids, typename Ranks>
	static void removeDuplicateRowId(Ids& ids, Ranks& ranks, size_t count) {
#include "_ref/gen/remove_dup.inc"
	}
	void removeOverK(auto& ids, h_vector<RankT, 128>& dists, const auto& params) const {
#include "_ref/gen/remove_over_k.inc"
	}
	void select(hnswlib::SearchResultQueue& knnRes, const Params& params, const Ctx& ctx, std::vector<int32_t>& outIds, std::vector<float>& outRanks) const {
#include "_ref/gen/select_drain.inc"
		outIds.clear();
		outRanks.clear();
		for (size_t i = 0; i < idset.size(); ++i) {
			outIds.push_back(idset[i].ToNumber());
			outRanks.push_back(dists[i].Value());
		}
	}
};

}  // namespace
}  // namespace reindexer

extern "C" {

// results: n (dist, label) pairs in ANY order (they are pushed into the reference's max-heap exactly like a Map returns them);
// metric 0 L2 / 1 IP / 2 Cosine; has_k + has_radius select removeOverK's trimming; index_radius = Opts().FloatVector().Radius().
// Returns the number of rows written (<= n).
int64_t ref_select_postprocess(int metric, int is_bruteforce, int need_sort, int is_array, int has_k, uint64_t k, int has_radius, float radius,
							   int has_index_radius, float index_radius, uint64_t n, const float* dists, const uint64_t* labels, int32_t* out_ids,
							   float* out_ranks) {
	using namespace reindexer;
	try {
		using pair_t = std::pair<float, hnswlib::labeltype>;
		std::vector<pair_t> cont;
		cont.reserve(n);
		for (uint64_t i = 0; i < n; ++i) {
			cont.emplace_back(dists[i], labels[i]);
		}
		hnswlib::SearchResultQueue q(std::less<pair_t>(), std::move(cont));
		Params p;
		if (has_k) {
			p.k = size_t(k);
		}
		if (has_radius) {
			p.radius = radius;
		}
		const Ctx ctx{need_sort != 0};
		IndexOpts o{is_array != 0, FvOpts{has_index_radius ? std::optional<float>(index_radius) : std::nullopt}};
		const VectorMetric m = metric == 0 ? VectorMetric::L2 : (metric == 1 ? VectorMetric::InnerProduct : VectorMetric::Cosine);
		std::vector<int32_t> ids;
		std::vector<float> ranks;
		if (is_bruteforce) {
			Harness<hnswlib::BruteforceSearch>{m, o}.select(q, p, ctx, ids, ranks);
		} else {
			Harness<int>{m, o}.select(q, p, ctx, ids, ranks);
		}
		for (size_t i = 0; i < ids.size(); ++i) {
			out_ids[i] = ids[i];
			out_ranks[i] = ranks[i];
		}
		return int64_t(ids.size());
	} catch (...) {
		return -1;
	}
}

}  // extern "C"
