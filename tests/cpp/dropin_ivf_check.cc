// TEST INFRASTRUCTURE.  Compiles the product's IVF adapter (reindexer_b200/host/gpu_ivf.h) against the reference's own vendored FAISS
// headers and drives it and a plain faiss::IndexIVFFlat side by side exactly like reindexer::IvfIndex drives its map_
// (cpp_src/core/index/float_vector/ivf_index.cc:87-132 upsert / del, :150-300 search / range_search): same trained centroids, same
// upserts and deletes, every search and range search compared.  Built by tests/cpp/Makefile only where /root/reference exists.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <set>
#include <vector>

#include "gpu_ivf.h"
#include "faiss/IndexFlat.h"
#include "tools/normalize.h"

namespace {

std::unique_ptr<faiss::IndexFlat> newSpace(size_t dim, int metric) {  // IvfIndex::newSpace, ivf_index.cc:686-695
	if (metric == 0) {
		return std::make_unique<faiss::IndexFlatL2>(dim);
	}
	if (metric == 1) {
		return std::make_unique<faiss::IndexFlatIP>(dim);
	}
	return std::make_unique<faiss::IndexFlatCosine>(dim);
}

struct Cpu {
	std::unique_ptr<faiss::IndexFlat> space;
	std::unique_ptr<faiss::IndexIVFFlat> map;
};

Cpu make(size_t dim, size_t nlist, int metric) {
	Cpu c;
	c.space = newSpace(dim, metric);
	c.map = std::make_unique<faiss::IndexIVFFlat>(c.space.get(), dim, nlist, metric == 0 ? faiss::METRIC_L2 : faiss::METRIC_INNER_PRODUCT, metric == 2);
	c.map->set_direct_map_type(faiss::DirectMap::Type::Hashtable);
	return c;
}

bool sameKnn(const std::vector<float>& da, const std::vector<faiss::idx_t>& ia, const std::vector<float>& db, const std::vector<faiss::idx_t>& ib) {
	for (size_t j = 0; j < ia.size(); ++j) {
		if ((ia[j] < 0) != (ib[j] < 0)) {
			return false;
		}
		if (ia[j] < 0) {
			continue;
		}
		if (std::abs(da[j] - db[j]) > 1e-4f * std::abs(da[j]) + 2e-6f) {
			return false;
		}
	}
	std::multiset<faiss::idx_t> a(ia.begin(), ia.end()), b(ib.begin(), ib.end());
	if (a == b) {
		return true;
	}
	// ids may differ only at the cut-off where the k-th and (k+1)-th distances are within fp noise
	size_t diff = 0;
	for (const auto id : a) {
		diff += b.count(id) == 0;
	}
	return diff <= 1;
}

int runMetric(int metric) {
	const size_t dim = 48, nlist = 20, n0 = 4000, extra = 1500, nq = 24;
	std::mt19937 rng(1234 + metric);
	std::normal_distribution<float> gauss(0.f, 1.f);
	std::vector<float> centers(64 * dim);
	for (auto& v : centers) {
		v = gauss(rng);
	}
	auto makeVec = [&](float* out) {
		const size_t c = rng() % 64;
		for (size_t i = 0; i < dim; ++i) {
			out[i] = centers[c * dim + i] + 0.4f * gauss(rng);
		}
	};
	std::vector<float> vecs((n0 + extra) * dim);
	for (size_t i = 0; i < n0 + extra; ++i) {
		makeVec(vecs.data() + i * dim);
	}
	std::vector<faiss::idx_t> ids(n0 + extra);
	for (size_t i = 0; i < ids.size(); ++i) {
		ids[i] = faiss::idx_t(i) << 32;  // FloatVectorId numbers: row id in the upper half
	}
	Cpu ref = make(dim, nlist, metric);
	ref.map->train(faiss::idx_t(n0), vecs.data());
	// the adapter's CPU half gets the SAME trained centroids (k-means is not re-run)
	Cpu mine = make(dim, nlist, metric);
	std::vector<float> cent(nlist * dim);
	ref.map->quantizer->reconstruct_n(0, faiss::idx_t(nlist), cent.data());
	mine.map->quantizer->add(faiss::idx_t(nlist), cent.data());
	mine.map->is_trained = true;
	ref.map->add_with_ids(faiss::idx_t(n0), vecs.data(), ids.data());
	mine.map->add_with_ids(faiss::idx_t(n0), vecs.data(), ids.data());
	reindexer::GpuIvfMap gpu(std::move(mine.map));

	std::vector<float> queries(nq * dim), qn(dim);
	for (size_t q = 0; q < nq; ++q) {
		makeVec(queries.data() + q * dim);
		if (metric == 2) {  // FloatVectorIndex normalises the key for Cosine (ivf_index.cc:307-316 via NormalizeCopyVector)
			reindexer::ann::NormalizeCopyVector(queries.data() + q * dim, int32_t(dim), qn.data());
			std::copy(qn.begin(), qn.end(), queries.begin() + q * dim);
		}
	}
	size_t searches = 0, same = 0, ranges = 0, sameRanges = 0;
	auto compare = [&]() {
		for (const size_t nprobe : {size_t(1), size_t(4), nlist}) {
			faiss::IVFSearchParameters params;
			params.nprobe = nprobe;
			for (size_t q = 0; q < nq; ++q) {
				const size_t k = q % 2 ? 10 : 25;
				std::vector<float> da(k), db(k);
				std::vector<faiss::idx_t> ia(k), ib(k);
				ref.map->search(1, queries.data() + q * dim, faiss::idx_t(k), da.data(), ia.data(), &params);
				gpu.search(1, queries.data() + q * dim, faiss::idx_t(k), db.data(), ib.data(), &params);
				same += sameKnn(da, ia, db, ib);
				++searches;
			}
			for (size_t q = 0; q < 4; ++q) {
				std::vector<float> d(20);
				std::vector<faiss::idx_t> i(20);
				ref.map->search(1, queries.data() + q * dim, 20, d.data(), i.data(), &params);
				const float radius = (d[10] + d[11]) / 2;  // between two neighbours: no boundary ambiguity
				faiss::RangeSearchResult ra(1), rb(1);
				ref.map->range_search(1, queries.data() + q * dim, radius, &ra, &params);
				gpu.range_search(1, queries.data() + q * dim, radius, &rb, &params);
				std::set<faiss::idx_t> sa(ra.labels + ra.lims[0], ra.labels + ra.lims[1]), sb(rb.labels + rb.lims[0], rb.labels + rb.lims[1]);
				sameRanges += sa == sb && !sa.empty();
				++ranges;
			}
		}
	};
	compare();
	size_t done = n0;
	std::vector<faiss::idx_t> alive(ids.begin(), ids.begin() + n0);
	for (const size_t burst : {size_t(1), size_t(60), size_t(700), size_t(739)}) {
		for (size_t i = done; i < done + burst; ++i) {  // IvfIndex::upsert: one add_with_ids per row
			ref.map->add_with_ids(1, vecs.data() + i * dim, &ids[i]);
			gpu.add_with_ids(1, vecs.data() + i * dim, &ids[i]);
			alive.push_back(ids[i]);
		}
		done += burst;
		for (size_t r = 0; r < alive.size() / 12; ++r) {  // IvfIndex::del
			const size_t at = rng() % alive.size();
			const faiss::idx_t id = alive[at];
			alive[at] = alive.back();
			alive.pop_back();
			ref.map->remove_ids(faiss::IDSelectorArray{1, &id});
			gpu.remove_ids(faiss::IDSelectorArray{1, &id});
		}
		compare();
	}
	const bool ok = same == searches && sameRanges == ranges && gpu.DeviceImports() == 1 && size_t(gpu->ntotal) == alive.size();
	std::printf("metric %d: %zu knn searches identical %zu, %zu range searches identical %zu, device imports %zu (upserts and deletes patched in "
				"place), rows %zu -> %s %s\n",
				metric, searches, same, ranges, sameRanges, gpu.DeviceImports(), alive.size(), ok ? "MATCH" : "MISMATCH", gpu.LastDeviceError().c_str());
	return ok ? 0 : 1;
}

}  // namespace

int main() {
	int bad = 0;
	for (const int metric : {0, 1, 2}) {
		bad += runMetric(metric);
	}
	return bad;
}
