// Drop-in `Map` for HnswIndexBase<Map> (cpp_src/core/index/float_vector/hnsw_index.h:13-57): same member surface as
// hnswlib::BruteforceSearch (cpp_src/core/index/float_vector/hnswlib/bruteforce.h:14-68), backed by librxgpu.so
// (include/rxgpu.h) instead of the host-memory scan.  This header is meant to be dropped into
// cpp_src/core/index/float_vector/hnswlib/ of the reference tree; it includes the reference's own headers and is therefore
// compiled only where that tree is available (tests/cpp/dropin_check.cc does so in the authoring container).
// INTEGRATION.md shows the 3-line patch of hnsw_index.cc that instantiates HnswIndexBase<hnswlib::GpuBruteforceSearch>.
#pragma once

#include <cstdlib>
#include <optional>
#include <stdexcept>
#include <utility>
#include <vector>

#include "core/index/float_vector/float_vector_id.h"
#include "core/index/float_vector/hnswlib/hnsw_interface.h"
#include "core/keyvalue/float_vector.h"
#include "rxgpu.h"

namespace hnswlib {

class [[nodiscard]] GpuBruteforceSearch {
public:
	GpuBruteforceSearch(reindexer::VectorMetric metric, size_t dim, size_t maxElements) : dim_(dim) {
		// the namespace needs host pointers to stored vectors (FloatPtrByExternalLabel, hnsw_index.cc:368): keep the mirror
		check(rxgpu_index_create(&h_, toMetric(metric), uint32_t(dim), maxElements, deviceFromEnv(), RXGPU_FLAG_HOST_MIRROR));
	}
	GpuBruteforceSearch(const GpuBruteforceSearch& other, size_t newMaxElements) : dim_(other.dim_) {
		check(rxgpu_index_clone(&h_, other.h_, newMaxElements));
	}
	GpuBruteforceSearch& operator=(const GpuBruteforceSearch&) = delete;
	~GpuBruteforceSearch() { rxgpu_index_destroy(h_); }

	size_t MaxElements() const noexcept { return rxgpu_index_capacity(h_); }
	size_t CurrentElementCount() const noexcept { return rxgpu_index_size(h_); }
	size_t ElementSize() const noexcept { return rxgpu_index_element_size(h_); }  // dim*4 + 8, as the memstat test expects

	const float* FloatPtrByExternalLabel(labeltype label) const {
		const float* row = nullptr;
		check(rxgpu_index_get(h_, label, &row));  // "Label not found"
		return row;
	}

	void AddPointNoLock(reindexer::ConstFloatVectorView vect, reindexer::FloatVectorId id) {
		check(rxgpu_index_upsert(h_, id.AsNumber(), vect.Data()));  // "The number of elements exceeds the specified limit"
	}
	[[noreturn]] void AddPointConcurrent(reindexer::ConstFloatVectorView, reindexer::FloatVectorId) {
		throw std::logic_error("This brute force index does not support concurrent insertions");
	}
	void RemovePoint(labeltype label) { check(rxgpu_index_remove(h_, label)); }
	void ResizeIndex(size_t newMaxElements) { check(rxgpu_index_resize(h_, newMaxElements)); }

	// returns the reference's max-heap (worst on top) so that HnswIndexBase::select drains it unchanged
	SearchResultQueue SearchKnn(const float* queryData, std::optional<float> /*queryDataNorm*/, size_t k, size_t /*ef*/ = 0) const {
		using pair_t = std::pair<float, labeltype>;
		const size_t n = std::min<size_t>(k, CurrentElementCount());
		if (n == 0) {
			return SearchResultQueue();
		}
		std::vector<float> dists(n);
		std::vector<uint64_t> labels(n);
		uint32_t count = 0;
		check(rxgpu_search_knn(h_, 1, queryData, uint32_t(n), dists.data(), labels.data(), &count));
		std::vector<pair_t> container;
		container.reserve(count);
		for (uint32_t i = 0; i < count; ++i) {
			container.emplace_back(dists[i], labels[i]);
		}
		return SearchResultQueue(std::less<pair_t>(), std::move(container));
	}
	SearchResultQueue SearchRange(const float* queryData, std::optional<float> /*queryDataNorm*/, float radius, size_t /*ef*/) const {
		using pair_t = std::pair<float, labeltype>;
		uint64_t total = 0;
		std::vector<float> dists(64);
		std::vector<uint64_t> labels(64);
		check(rxgpu_search_range(h_, queryData, radius, dists.size(), dists.data(), labels.data(), &total));
		if (total > dists.size()) {  // the library retained the whole result of this thread's scan: fetch the tail, no rescan
			const uint64_t have = dists.size();
			dists.resize(total);
			labels.resize(total);
			check(rxgpu_last_range_results(have, total - have, dists.data() + have, labels.data() + have));
		}
		std::vector<pair_t> container;
		container.reserve(total);
		for (uint64_t i = 0; i < total; ++i) {
			container.emplace_back(dists[i], labels[i]);
		}
		return SearchResultQueue(std::less<pair_t>(), std::move(container));
	}

	bool IsQuantized() const noexcept { return false; }
	bool QuantizationAvailable() const noexcept { return false; }
	size_t AllocatedMemSize() const noexcept {
		// host side only, like the reference's accounting of not-yet-used capacity (bruteforce.h:41-44); HBM is reported apart
		return (MaxElements() - CurrentElementCount()) * ElementSize() + sizeof(GpuBruteforceSearch);
	}
	size_t DeviceMemSize() const noexcept { return rxgpu_index_device_bytes(h_); }

private:
	static rxgpu_metric toMetric(reindexer::VectorMetric m) {
		switch (m) {
			case reindexer::VectorMetric::L2:
				return RXGPU_L2;
			case reindexer::VectorMetric::InnerProduct:
				return RXGPU_IP;
			case reindexer::VectorMetric::Cosine:
				return RXGPU_COS;
		}
		std::abort();
	}
	static int deviceFromEnv() {
		const char* e = std::getenv("RX_GPU_DEVICE");
		return e ? std::atoi(e) : 0;
	}
	static void check(int rc) {
		if (rc == RXGPU_OK) {
			return;
		}
		// the reference's BruteforceSearch throws std::runtime_error with these texts; HnswIndexBase lets them propagate
		throw std::runtime_error(rxgpu_last_error());
	}

	size_t dim_;
	rxgpu_index* h_ = nullptr;
};

}  // namespace hnswlib
