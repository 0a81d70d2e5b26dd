// TEST INFRASTRUCTURE.  Compiles the product's HNSW adapter (reindexer_b200/host/gpu_hnsw.h) against the reference's own headers and
// drives it and the reference's hnswlib::HierarchicalNSW<Synchronization::None> through ONE template -- the member calls
// HnswIndexBase<Map> makes (cpp_src/core/index/float_vector/hnsw_index.cc:46-58, 89-97, 116-118, 160-191) -- then diffs the drained
// heaps.  Both maps build the same graph (the adapter delegates insertion to the reference's inserter, level RNG seeded with 100), so
// the answers may differ only where the reference's AVX-512 sums and the device's sums differ in the last bits.
// Built by tests/cpp/Makefile only where /root/reference exists; the binary travels to the GPU box.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <span>
#include <utility>
#include <vector>

#include "core/index/float_vector/hnswlib/hnsw.h"
#include "gpu_hnsw.h"
#include "tools/normalize.h"

extern "C" float port_synth_value(uint64_t seed, uint64_t index);  // oracle/knn_port.c

using Results = std::vector<std::vector<std::pair<float, uint64_t>>>;

template <typename Map>
Results drive(reindexer::VectorMetric metric, size_t dim, size_t n, size_t k, size_t ef, size_t nq) {
	Map map(reindexer::IsArray_False, metric, dim, n / 2, 16, 200);
	std::vector<float> v(dim);
	auto add = [&](size_t i) {
		if (map.CurrentElementCount() >= map.MaxElements()) {
			map.ResizeIndex(map.MaxElements() * 2);  // HnswIndexBase::upsert grows the map (hnsw_index.cc:89-92)
		}
		for (size_t c = 0; c < dim; ++c) {
			v[c] = port_synth_value(177, i * dim + c) + 0.6f * port_synth_value(178, (i % 37) * dim + c);  // clustered
		}
		map.AddPointNoLock(reindexer::ConstFloatVectorView{std::span<const float>{v}}, reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i)), 0});
	};
	Results out;
	std::vector<float> q(dim), qn(dim);
	auto search = [&](const Map& map, bool range) {
		for (size_t qi = 0; qi < nq; ++qi) {
			for (size_t c = 0; c < dim; ++c) {
				q[c] = port_synth_value(179, qi * dim + c) + 0.6f * port_synth_value(178, (qi % 37) * dim + c);
			}
			const float* key = q.data();
			std::optional<float> norm;
			if (metric == reindexer::VectorMetric::Cosine) {  // HnswIndexBase::search (hnsw_index.cc:166-171)
				norm = 1.f / reindexer::ann::NormalizeCopyVector(q.data(), int32_t(dim), qn.data());
				key = qn.data();
			}
			auto res = range ? map.SearchRange(key, norm, metric == reindexer::VectorMetric::L2 ? 40.f : -0.55f, ef) : map.SearchKnn(key, norm, k, ef);
			std::vector<std::pair<float, uint64_t>> r(res.size());
			for (auto i = res.size(); !res.empty(); res.pop()) {
				r[--i] = res.top();
			}
			out.emplace_back(std::move(r));
		}
	};
	for (size_t i = 0; i < n; ++i) {
		add(i);
	}
	search(map, false);
	for (size_t i = 0; i < n; i += 9) {  // tombstones: applied in place on the device copy
		map.MarkDelete(reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i)), 0});
	}
	search(map, false);
	search(map, true);
	for (size_t i = n; i < n + n / 4; ++i) {  // more inserts (some reuse tombstoned slots): the device copy is rebuilt by the next search
		add(i);
	}
	search(map, false);
	const Map clone(std::as_const(map), map.MaxElements() + 10);  // COW namespace clone (hnsw_index.cc:66-68)
	search(clone, false);
	return out;
}

int main() {
	int bad = 0;
	const size_t dim = 48, n = 3000, k = 10, ef = 64, nq = 40;
	for (auto metric : {reindexer::VectorMetric::L2, reindexer::VectorMetric::Cosine, reindexer::VectorMetric::InnerProduct}) {
		const auto ref = drive<hnswlib::HierarchicalNSW<hnswlib::Synchronization::None>>(metric, dim, n, k, ef, nq);
		const auto gpu = drive<hnswlib::GpuHnsw<hnswlib::Synchronization::None>>(metric, dim, n, k, ef, nq);
		size_t same = 0, total = ref.size(), closeDist = 0, nonEmpty = 0;
		for (size_t i = 0; i < total; ++i) {
			bool ids = ref[i].size() == gpu[i].size();
			bool dists = ids;
			for (size_t j = 0; ids && j < ref[i].size(); ++j) {
				ids = ref[i][j].second == gpu[i][j].second;
			}
			for (size_t j = 0; dists && j < ref[i].size(); ++j) {
				dists = std::abs(ref[i][j].first - gpu[i][j].first) <= 1e-4f * std::abs(ref[i][j].first) + 2e-6f;
			}
			same += ids;
			closeDist += dists;
			nonEmpty += !ref[i].empty();
		}
		const bool ok = total == gpu.size() && same * 100 >= total * 95 && closeDist * 100 >= total * 95 && nonEmpty * 2 >= total;
		std::printf("metric %d: %zu searches (knn, knn with tombstones, range, knn after more inserts, knn on a clone), identical ids %zu, distances within 1e-4 %zu, "
					"non-empty %zu -> %s\n",
					int(metric), total, same, closeDist, nonEmpty, ok ? "MATCH" : "MISMATCH");
		bad += !ok;
	}
	return bad;
}
