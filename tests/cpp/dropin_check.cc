// TEST INFRASTRUCTURE.  Compiles the product's C++ adapter (reindexer_b200/host/gpu_bruteforce.h) against the reference's own
// headers and drives it and the reference's hnswlib::BruteforceSearch through ONE template -- the same member calls
// HnswIndexBase<Map> makes (cpp_src/core/index/float_vector/hnsw_index.cc:61-70, 89-97, 121-123, 160-191, 232-288) -- then
// diffs the drained heaps.  Built by tests/cpp/Makefile only where /root/reference exists; the binary travels to the GPU box.
#include <cstdio>
#include <cstring>
#include <span>
#include <vector>

#include "core/index/float_vector/hnswlib/bruteforce.h"
#include "gpu_bruteforce.h"

extern "C" float port_synth_value(uint64_t seed, uint64_t index);  // oracle/knn_port.c

template <typename Map>
std::vector<std::pair<float, uint64_t>> drive(reindexer::VectorMetric metric, size_t dim, size_t n, size_t k, bool range) {
	Map map(metric, dim, n / 2);
	std::vector<float> v(dim);
	for (size_t i = 0; i < n; ++i) {
		if (map.CurrentElementCount() >= map.MaxElements()) {
			map.ResizeIndex(map.MaxElements() * 2);  // HnswIndexBase::upsert grows the map (hnsw_index.cc:89-92)
		}
		for (size_t c = 0; c < dim; ++c) {
			v[c] = port_synth_value(77, i * dim + c);
		}
		map.AddPointNoLock(reindexer::ConstFloatVectorView{std::span<const float>{v}},
						   reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i / 2)), uint32_t(i % 2)});
	}
	for (size_t i = 0; i < n; i += 7) {
		map.RemovePoint(reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i / 2)), uint32_t(i % 2)}.AsNumber());
	}
	Map clone(map, map.MaxElements() + 10);  // COW namespace clone (hnsw_index.cc:126-128)
	std::vector<float> q(dim);
	for (size_t c = 0; c < dim; ++c) {
		q[c] = port_synth_value(78, c);
	}
	if (std::memcmp(clone.FloatPtrByExternalLabel(reindexer::FloatVectorId{reindexer::IdType::FromNumber(1), 1}.AsNumber()),
					map.FloatPtrByExternalLabel(reindexer::FloatVectorId{reindexer::IdType::FromNumber(1), 1}.AsNumber()),
					dim * sizeof(float)) != 0) {
		throw std::runtime_error("clone row mismatch");
	}
	auto res = range ? clone.SearchRange(q.data(), std::nullopt, metric == reindexer::VectorMetric::L2 ? 45.f : -1.0f, 0)
					 : clone.SearchKnn(q.data(), std::nullopt, k);
	std::vector<std::pair<float, uint64_t>> out(res.size());
	for (auto i = res.size(); !res.empty(); res.pop()) {
		out[--i] = res.top();
	}
	if (map.ElementSize() != dim * 4 + 8) {
		throw std::runtime_error("ElementSize");
	}
	return out;
}

int main() {
	int bad = 0;
	for (auto metric : {reindexer::VectorMetric::L2, reindexer::VectorMetric::InnerProduct}) {
		for (bool range : {false, true}) {
			const auto ref = drive<hnswlib::BruteforceSearch>(metric, 96, 3000, 25, range);
			const auto gpu = drive<hnswlib::GpuBruteforceSearch>(metric, 96, 3000, 25, range);
			bool ok = ref.size() == gpu.size() && !ref.empty();
			for (size_t i = 0; ok && i < ref.size(); ++i) {
				ok = ref[i].second == gpu[i].second && std::abs(ref[i].first - gpu[i].first) <= 1e-4f * std::abs(ref[i].first) + 2e-6f;
			}
			std::printf("metric %d %s: %zu results %s\n", int(metric), range ? "range" : "knn", ref.size(), ok ? "MATCH" : "MISMATCH");
			bad += !ok;
		}
	}
	return bad;
}
