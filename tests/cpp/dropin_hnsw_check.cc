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
#include <string>
#include <atomic>
#include <utility>
#include <vector>

#include "core/index/float_vector/hnswlib/hnsw.h"
#include "gpu_hnsw.h"
#include "tools/normalize.h"

extern "C" float port_synth_value(uint64_t seed, uint64_t index);  // oracle/knn_port.c

using Results = std::vector<std::vector<std::pair<float, uint64_t>>>;

// In-memory stand-ins for the namespace's index-cache writer / reader (hnsw_index.cc:389-507): the graph goes out WITHOUT the vectors
// (only the primary key of every live element), the reader fetches each vector back by its key -- here from a map the test keeps.
struct Token {
	uint64_t u = 0;
	int64_t i = 0;
	float f = 0.f;
	std::string s;
};
class MemWriter final : public hnswlib::IWriter {
public:
	std::vector<Token> tokens;
	void PutVarUInt(uint64_t v) override { tokens.push_back(Token{v, 0, 0.f, {}}); }
	void PutVarUInt(uint32_t v) override { tokens.push_back(Token{v, 0, 0.f, {}}); }
	void PutVarInt(int64_t v) override { tokens.push_back(Token{0, v, 0.f, {}}); }
	void PutVarInt(int32_t v) override { tokens.push_back(Token{0, v, 0.f, {}}); }
	void PutVString(std::string_view v) override { tokens.push_back(Token{0, 0, 0.f, std::string(v)}); }
	void PutFloat(float v) override { tokens.push_back(Token{0, 0, v, {}}); }
	void AppendPKByID(hnswlib::labeltype label) override { tokens.push_back(Token{label, 0, 0.f, {}}); }
};
class MemReader final : public hnswlib::IReader {
public:
	MemReader(const std::vector<Token>& t, const std::vector<std::vector<float>>& rows) : tokens_(t), rows_(rows) {}
	uint64_t GetVarUInt() override { return tokens_[pos_++].u; }
	int64_t GetVarInt() override { return tokens_[pos_++].i; }
	std::string_view GetVString() override { return tokens_[pos_++].s; }
	float GetFloat() override { return tokens_[pos_++].f; }
	hnswlib::labeltype ReadPkEncodedData(float* destBuf) override {
		const uint64_t label = tokens_[pos_++].u;
		const auto& v = rows_[size_t(label >> 32)];
		std::memcpy(destBuf, v.data(), v.size() * sizeof(float));
		return label;
	}
	bool WithQuantizer() const override { return false; }

private:
	const std::vector<Token>& tokens_;
	const std::vector<std::vector<float>>& rows_;
	size_t pos_ = 0;
};

template <typename Map>
Results drive(reindexer::VectorMetric metric, size_t dim, size_t n, size_t k, size_t ef, size_t nq) {
	Map map(reindexer::IsArray_False, metric, dim, n / 2, 16, 200);
	std::vector<float> v(dim);
	std::vector<std::vector<float>> rowsByPk;  // what the namespace holds: vectors by row id
	auto add = [&](size_t i) {
		if (map.CurrentElementCount() >= map.MaxElements()) {
			map.ResizeIndex(map.MaxElements() * 2);  // HnswIndexBase::upsert grows the map (hnsw_index.cc:89-92)
		}
		for (size_t c = 0; c < dim; ++c) {
			v[c] = port_synth_value(177, i * dim + c) + 0.6f * port_synth_value(178, (i % 37) * dim + c);  // clustered
		}
		map.AddPointNoLock(reindexer::ConstFloatVectorView{std::span<const float>{v}}, reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i)), 0});
		if (rowsByPk.size() <= i) {
			rowsByPk.resize(i + 1);
		}
		rowsByPk[i] = v;
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
	// steady state: upserts and deletes interleaved with searches.  The device copy must be PATCHED (rxgpu_hnsw_update), not re-imported.
	size_t importsBefore = 0;
	if constexpr (requires { map.DeviceImports(); }) {
		importsBefore = map.DeviceImports();
	}
	for (size_t i = n + n / 4, round = 0; round < 4; ++round) {
		for (size_t j = 0; j < 40; ++j, ++i) {
			add(i);
			if (j % 8 == 3) {  // a tombstone, reused by the insert after the next one
				map.MarkDelete(reindexer::FloatVectorId{reindexer::IdType::FromNumber(int(i - 2)), 0});
			}
			if (j % 16 == 9) {  // the same id again with another vector (updatePoint on the live node)
				add(i - 1);
			}
		}
		search(map, false);
	}
	if constexpr (requires { map.DeviceImports(); }) {
		if (map.DeviceImports() != importsBefore || map.DevicePatchedNodes() == 0) {
			std::printf("steady-state upserts re-imported the device copy (%zu -> %zu imports, %zu patched nodes): %s\n", importsBefore, map.DeviceImports(),
						map.DevicePatchedNodes(), map.LastPatchError().c_str());
			std::exit(2);
		}
		std::printf("  steady state: 160 upserts + tombstones patched %zu nodes in place, %zu imports in total\n", map.DevicePatchedNodes(), map.DeviceImports());
	}
	const Map clone(std::as_const(map), map.MaxElements() + 10);  // COW namespace clone (hnsw_index.cc:66-68)
	search(clone, false);
	// index cache round trip (WriteIndexCache / LoadIndexCache, hnsw_index.cc:389-507; hnswalg.h:1213-1263, loader :297-...)
	MemWriter w;
	const std::atomic_int32_t cancel{0};
	map.SaveIndex(w, cancel);
	Map loaded(reindexer::IsArray_False, metric, dim, map.MaxElements(), 16, 200);
	MemReader r(w.tokens, rowsByPk);
	loaded.LoadIndex(r);
	search(loaded, false);
	if constexpr (requires { map.DeviceImports(); }) {
		// the same cache restored by the library itself, with NO host graph (rxgpu_hnsw_load_index_cache): the token callbacks sit on the
		// same reader; the answers must equal those of the map restored through the reference's loader
		MemReader r2(w.tokens, rowsByPk);
		(void)r2.GetVarUInt();  // the quantisation header HierarchicalNSW::SaveIndex puts first (hnsw.cc:52-58)
		rxgpu_hnsw_cache_reader cb{};
		cb.ctx = &r2;
		cb.get_var_uint = [](void* c) { return static_cast<MemReader*>(c)->GetVarUInt(); };
		cb.get_var_int = [](void* c) { return static_cast<MemReader*>(c)->GetVarInt(); };
		cb.get_vstring = [](void* c, const char** data, uint64_t* len) {
			const std::string_view v = static_cast<MemReader*>(c)->GetVString();
			*data = v.data();
			*len = v.size();
			return 0;
		};
		cb.read_pk_encoded_data = [](void* c, float* dest) { return uint64_t(static_cast<MemReader*>(c)->ReadPkEncodedData(dest)); };
		rxgpu_index* ix = nullptr;
		const auto m = metric == reindexer::VectorMetric::L2 ? RXGPU_L2 : metric == reindexer::VectorMetric::Cosine ? RXGPU_COS : RXGPU_IP;
		if (rxgpu_index_create(&ix, m, uint32_t(dim), map.MaxElements(), 0, 0) != RXGPU_OK) {
			std::printf("rxgpu_index_create: %s\n", rxgpu_last_error());
			std::exit(3);
		}
		rxgpu_hnsw_cache_info info{};
		if (rxgpu_hnsw_load_index_cache(ix, &cb, &info) != RXGPU_OK || info.count != map.CurrentElementCount() ||
			info.deleted != map.DeletedCountUnsafe()) {
			std::printf("rxgpu_hnsw_load_index_cache: %s (count %llu of %zu, deleted %u of %zu)\n", rxgpu_last_error(),
						(unsigned long long)info.count, map.CurrentElementCount(), info.deleted, map.DeletedCountUnsafe());
			std::exit(3);
		}
		const size_t first = out.size() - nq;  // the answers of `loaded`
		size_t same = 0;
		for (size_t qi = 0; qi < nq; ++qi) {
			for (size_t c = 0; c < dim; ++c) {
				q[c] = port_synth_value(179, qi * dim + c) + 0.6f * port_synth_value(178, (qi % 37) * dim + c);
			}
			const float* key = q.data();
			if (metric == reindexer::VectorMetric::Cosine) {
				(void)reindexer::ann::NormalizeCopyVector(q.data(), int32_t(dim), qn.data());
				key = qn.data();
			}
			std::vector<float> d(k);
			std::vector<uint64_t> l(k);
			uint32_t cnt = 0;
			if (rxgpu_hnsw_search_knn(ix, 1, key, uint32_t(k), uint32_t(ef), d.data(), l.data(), &cnt, nullptr) != RXGPU_OK) {
				std::printf("search on the restored index: %s\n", rxgpu_last_error());
				std::exit(3);
			}
			bool eq = cnt == out[first + qi].size();
			for (uint32_t j = 0; eq && j < cnt; ++j) {
				eq = l[j] == out[first + qi][j].second && d[j] == out[first + qi][j].first;
			}
			same += eq;
		}
		rxgpu_index_destroy(ix);
		std::printf("  index cache restored by the library (no host graph): %zu of %zu answers bit-identical to the adapter's restored map\n", same, nq);
		if (same != nq) {
			std::exit(3);
		}
	}
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
		std::printf("metric %d: %zu searches (knn, knn with tombstones, range, knn after more inserts, knn between steady-state upserts, knn on a clone, knn on a map restored from its index cache), identical ids %zu, distances within 1e-4 %zu, "
					"non-empty %zu -> %s\n",
					int(metric), total, same, closeDist, nonEmpty, ok ? "MATCH" : "MISMATCH");
		bad += !ok;
	}
	return bad;
}
