// Drop-in replacement for the `std::unique_ptr<faiss::IndexIVFFlat> map_` member of reindexer::IvfIndex
// (cpp_src/core/index/float_vector/ivf_index.h:75, used at ivf_index.cc:87-132 upsert / del, :150-300 search / range_search through the
// `const auto& map` templates, :309-316, :420-429): the CPU faiss::IndexIVFFlat stays the row store (getView, reconstruct, index cache,
// clone, RebuildCentroids read its inverted lists and direct map through operator->), the four calls on the query/update path --
//   search(1, key, k, dists, ids, &IVFSearchParameters{nprobe})      range_search(1, key, radius, &result, &params)
//   add_with_ids(1, vec[, norm], &id)                                 remove_ids(IDSelectorArray{1, &id})
// -- are served by librxgpu (include/rxgpu.h: rxgpu_ivf_create / _add / _remove / _search_knn / _search_range).  The device lists are
// filled once, from the trained index, by the first search; after that every upsert / delete patches them in place (the list number
// is read back from FAISS' direct map, so both sides agree on the assignment bit for bit).  Distances follow FAISS' conventions
// (L2: squared distance ascending; inner product / cosine: +similarity descending, labels -1 past the end).
// Meant to be dropped into cpp_src/core/index/float_vector/; compiled only where the reference tree is available
// (tests/cpp/dropin_ivf_check.cc does so in the authoring container).  INTEGRATION.md section 8 shows the patch.
#pragma once

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "faiss/IndexIVFFlat.h"
#include "faiss/impl/AuxIndexStructures.h"
#include "faiss/impl/IDSelector.h"
#include "faiss/invlists/DirectMap.h"
#include "rxgpu.h"

namespace reindexer {

class [[nodiscard]] GpuIvfMap {
public:
	GpuIvfMap() = default;
	explicit GpuIvfMap(std::unique_ptr<faiss::IndexIVFFlat> idx) : cpu_(std::move(idx)) {}
	GpuIvfMap(const GpuIvfMap&) = delete;
	GpuIvfMap& operator=(const GpuIvfMap&) = delete;
	GpuIvfMap& operator=(std::unique_ptr<faiss::IndexIVFFlat>&& idx) noexcept {  // map_ = std::move(idx), ivf_index.cc:104,609,681
		releaseDevice();
		cpu_ = std::move(idx);
		return *this;
	}
	~GpuIvfMap() { releaseDevice(); }

	explicit operator bool() const noexcept { return bool(cpu_); }
	faiss::IndexIVFFlat* operator->() const noexcept { return cpu_.get(); }
	faiss::IndexIVFFlat& operator*() const noexcept { return *cpu_; }
	faiss::IndexIVFFlat* get() const noexcept { return cpu_.get(); }
	void reset() noexcept {
		releaseDevice();
		cpu_.reset();
	}

	void add_with_ids(faiss::idx_t n, const float* x, const faiss::idx_t* ids) { add_with_ids(n, x, nullptr, ids); }
	void add_with_ids(faiss::idx_t n, const float* x, const float* norms, const faiss::idx_t* ids) {
		if (norms) {
			cpu_->add_with_ids(n, x, norms, ids);
		} else {
			cpu_->add_with_ids(n, x, ids);
		}
		std::lock_guard<std::mutex> lck(mtx_);
		if (!gpu_) {
			return;  // filled from the CPU index by the first search
		}
		std::vector<uint32_t> lists(n);
		std::vector<uint64_t> labels(n);
		for (faiss::idx_t i = 0; i < n; ++i) {
			lists[i] = listOf(ids[i]);
			labels[i] = uint64_t(ids[i]);
		}
		if (rxgpu_ivf_add(gpu_, uint64_t(n), lists.data(), labels.data(), x) != RXGPU_OK) {
			lastError_ = rxgpu_last_error();
			releaseDevice();  // rebuilt from the CPU index by the next search
		}
	}
	size_t remove_ids(const faiss::IDSelector& sel) {
		const auto* arr = dynamic_cast<const faiss::IDSelectorArray*>(&sel);
		if (!arr) {
			throw std::logic_error("GpuIvfMap: remove_ids takes an IDSelectorArray (what IvfIndex::del passes)");
		}
		std::vector<faiss::idx_t> present;
		for (size_t i = 0; i < arr->n; ++i) {
			if (cpu_->direct_map.hashtable.find(arr->ids[i]) != cpu_->direct_map.hashtable.end()) {
				present.push_back(arr->ids[i]);
			}
		}
		const size_t removed = cpu_->remove_ids(sel);
		std::lock_guard<std::mutex> lck(mtx_);
		for (const faiss::idx_t id : present) {
			if (gpu_ && rxgpu_ivf_remove(gpu_, uint64_t(id)) != RXGPU_OK) {
				lastError_ = rxgpu_last_error();
				releaseDevice();
			}
		}
		return removed;
	}

	void search(faiss::idx_t n, const float* x, faiss::idx_t k, float* distances, faiss::idx_t* labels,
				const faiss::SearchParameters* params = nullptr) const {
		ensureDevice();
		const uint32_t nprobe = nprobeOf(params);
		std::vector<uint64_t> lab(size_t(n) * k);
		std::vector<uint32_t> cnt(n);
		check(rxgpu_ivf_search_knn(gpu_, uint32_t(n), x, uint32_t(k), nprobe, distances, lab.data(), cnt.data()));
		const bool similarity = cpu_->metric_type != faiss::METRIC_L2;
		for (faiss::idx_t q = 0; q < n; ++q) {
			for (faiss::idx_t j = 0; j < k; ++j) {
				const size_t at = size_t(q) * k + j;
				if (j < cnt[q]) {
					labels[at] = faiss::idx_t(lab[at]);
					if (similarity) {
						distances[at] = -distances[at];
					}
				} else {  // faiss pads with -1 and the heap's neutral value
					labels[at] = -1;
					distances[at] = similarity ? -std::numeric_limits<float>::max() : std::numeric_limits<float>::max();
				}
			}
		}
	}
	void range_search(faiss::idx_t n, const float* x, float radius, faiss::RangeSearchResult* result,
					  const faiss::SearchParameters* params = nullptr) const {
		ensureDevice();
		const uint32_t nprobe = nprobeOf(params);
		const bool similarity = cpu_->metric_type != faiss::METRIC_L2;
		std::vector<std::vector<float>> d(n);
		std::vector<std::vector<uint64_t>> l(n);
		for (faiss::idx_t q = 0; q < n; ++q) {
			uint64_t total = 0;
			d[q].resize(256);
			l[q].resize(256);
			const float r = similarity ? -radius : radius;  // map space: dist < r
			check(rxgpu_ivf_search_range(gpu_, x + size_t(q) * cpu_->d, r, nprobe, d[q].size(), d[q].data(), l[q].data(), &total));
			if (total > d[q].size()) {
				d[q].resize(total);
				l[q].resize(total);
				check(rxgpu_ivf_search_range(gpu_, x + size_t(q) * cpu_->d, r, nprobe, d[q].size(), d[q].data(), l[q].data(), &total));
			}
			d[q].resize(total);
			l[q].resize(total);
			result->lims[q] = total;
		}
		result->do_allocation();  // turns the counts in lims into offsets and allocates labels / distances
		for (faiss::idx_t q = 0; q < n; ++q) {
			for (size_t i = 0; i < d[q].size(); ++i) {
				result->labels[result->lims[q] + i] = faiss::idx_t(l[q][i]);
				result->distances[result->lims[q] + i] = similarity ? -d[q][i] : d[q][i];
			}
		}
	}

	size_t DeviceImports() const noexcept { return imports_; }
	const std::string& LastDeviceError() const noexcept { return lastError_; }

private:
	static void check(int rc) {
		if (rc != RXGPU_OK) {
			throw std::runtime_error(rxgpu_last_error());
		}
	}
	static int deviceFromEnv() {
		const char* e = std::getenv("RX_GPU_DEVICE");
		return e ? std::atoi(e) : 0;
	}
	uint32_t nprobeOf(const faiss::SearchParameters* params) const {
		if (const auto* p = dynamic_cast<const faiss::IVFSearchParameters*>(params)) {
			return uint32_t(p->nprobe);
		}
		return uint32_t(cpu_->nprobe);
	}
	uint32_t listOf(faiss::idx_t id) const {
		const auto it = cpu_->direct_map.hashtable.find(id);
		if (it == cpu_->direct_map.hashtable.end()) {
			throw std::logic_error("GpuIvfMap: the id is missing from FAISS' direct map (set_direct_map_type(Hashtable) is required)");
		}
		return uint32_t(faiss::lo_listno(it->second));
	}
	void releaseDevice() noexcept {
		if (gpu_) {
			rxgpu_index_destroy(gpu_);
			gpu_ = nullptr;
		}
	}
	// one import of the trained index: centroids from the coarse quantiser, then every inverted list as one batch
	void ensureDevice() const {
		std::lock_guard<std::mutex> lck(mtx_);
		if (gpu_) {
			return;
		}
		const faiss::IndexIVFFlat& idx = *cpu_;
		const auto metric = idx.metric_type == faiss::METRIC_L2 ? RXGPU_L2 : idx.is_cosine ? RXGPU_COS : RXGPU_IP;
		rxgpu_index* ix = nullptr;
		check(rxgpu_index_create(&ix, metric, uint32_t(idx.d), 16, deviceFromEnv(), 0));
		std::unique_ptr<rxgpu_index, void (*)(rxgpu_index*)> guard(ix, rxgpu_index_destroy);
		std::vector<float> centroids(idx.nlist * size_t(idx.d));
		idx.quantizer->reconstruct_n(0, faiss::idx_t(idx.nlist), centroids.data());
		check(rxgpu_ivf_create(ix, uint32_t(idx.nlist), centroids.data()));
		std::vector<uint32_t> lists;
		std::vector<uint64_t> labels;
		for (size_t l = 0; l < idx.nlist; ++l) {
			const size_t sz = idx.invlists->list_size(l);
			if (!sz) {
				continue;
			}
			faiss::InvertedLists::ScopedCodes codes(idx.invlists, l);
			faiss::InvertedLists::ScopedIds ids(idx.invlists, l);
			lists.assign(sz, uint32_t(l));
			labels.assign(ids.get(), ids.get() + sz);
			check(rxgpu_ivf_add(ix, sz, lists.data(), labels.data(), reinterpret_cast<const float*>(codes.get())));
		}
		gpu_ = guard.release();
		++imports_;
	}

	std::unique_ptr<faiss::IndexIVFFlat> cpu_;
	mutable rxgpu_index* gpu_ = nullptr;
	mutable std::mutex mtx_;
	mutable size_t imports_ = 0;
	mutable std::string lastError_;
};

}  // namespace reindexer
