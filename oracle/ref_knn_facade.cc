// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product (librxgpu.so).
//
// extern "C" facade over the *unmodified* reference classes, compiled in place from
// /root/reference/cpp_src by oracle/Makefile into oracle/_ref/liboracle_ref_knn.so:
//   hnswlib::BruteforceSearch                       cpp_src/core/index/float_vector/hnswlib/bruteforce.{h,cc}
//   hnswlib::HierarchicalNSWImpl<float, Sync>       cpp_src/core/index/float_vector/hnswlib/hnswalg.h
//   reindexer::ann::{CalculateL2Module,NormalizeCopyVector}   cpp_src/tools/normalize.{h,cc}
//   reindexer::vector_dists::{L2SqrDistance,InnerProductDistance}  cpp_src/tools/distances/{l2,ip}_dist.{h,cc}
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
//
// Result convention of every *_search_* function here: the reference returns a max-heap
// (worst on top); we drain it exactly like HnswIndexBase::select does
// (cpp_src/core/index/float_vector/hnsw_index.cc:258-276): popped elements fill slots
// n-1 ... 0, so out[0] is the best.  Distances keep the map-space sign convention
// (L2: +L2^2, IP: -IP, Cosine: -cos), i.e. "smaller is better".
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <span>
#include <string>
#include <thread>
#include <vector>

#include "core/index/float_vector/hnswlib/bruteforce.h"
#include "core/index/float_vector/hnswlib/hnswalg.h"
#include "tools/cpucheck.h"
#include "tools/distances/ip_dist.h"
#include "tools/distances/l2_dist.h"
#include "tools/normalize.h"

namespace {

thread_local std::string g_err;

reindexer::VectorMetric toMetric(int m) {
	switch (m) {
		case 0:
			return reindexer::VectorMetric::L2;
		case 1:
			return reindexer::VectorMetric::InnerProduct;
		default:
			return reindexer::VectorMetric::Cosine;
	}
}

size_t drain(hnswlib::SearchResultQueue& q, size_t maxOut, float* dists, uint64_t* labels) {
	const size_t n = q.size();
	for (size_t i = n; !q.empty(); q.pop()) {
		--i;
		if (i < maxOut) {
			dists[i] = q.top().first;
			labels[i] = q.top().second;
		}
	}
	return n;
}

template <typename Fn>
int guarded(Fn&& fn) noexcept {
	try {
		fn();
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	} catch (...) {
		g_err = "unknown exception";
		return 1;
	}
}

using HnswST = hnswlib::HierarchicalNSWImpl<float, hnswlib::Synchronization::None>;
using HnswMT = hnswlib::HierarchicalNSWImpl<float, hnswlib::Synchronization::OnInsertions>;

struct HnswHandle {
	std::unique_ptr<HnswST> st;
	std::unique_ptr<HnswMT> mt;
	size_t dim = 0;
	size_t M = 0;
};

template <typename F>
auto withHnsw(HnswHandle* h, F&& f) {
	return h->st ? f(*h->st) : f(*h->mt);
}

template <typename SearchFn>
void parallelQueries(uint32_t nq, int threads, SearchFn&& fn) {
	if (threads <= 1) {
		for (uint32_t i = 0; i < nq; ++i) {
			fn(i);
		}
		return;
	}
	std::atomic<uint32_t> next{0};
	std::vector<std::thread> pool;
	pool.reserve(threads);
	for (int t = 0; t < threads; ++t) {
		pool.emplace_back([&] {
			for (uint32_t i = next.fetch_add(1); i < nq; i = next.fetch_add(1)) {
				fn(i);
			}
		});
	}
	for (auto& th : pool) {
		th.join();
	}
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// 3 = AVX-512, 2 = AVX2, 1 = AVX, 0 = SSE/scalar: what the reference's runtime dispatch picked (tools/cpucheck.cc:199-230)
int ref_isa_level() {
	if (reindexer::IsAVX512Allowed()) {
		return 3;
	}
	if (reindexer::IsAVX2Allowed()) {
		return 2;
	}
	if (reindexer::IsAVXAllowed()) {
		return 1;
	}
	return 0;
}

float ref_l2sqr(const float* a, const float* b, size_t d) { return reindexer::vector_dists::L2SqrDistance(a, b, d); }
float ref_ip(const float* a, const float* b, size_t d) { return reindexer::vector_dists::InnerProductDistance(a, b, d); }
float ref_calc_l2_module(const float* x, int32_t d) { return reindexer::ann::CalculateL2Module(x, d); }
float ref_normalize_copy(const float* x, int32_t d, float* out) { return reindexer::ann::NormalizeCopyVector(x, d, out); }

// ---------------------------------------------------------------- brute force
void* ref_bf_create(int metric, size_t dim, size_t capacity) {
	void* res = nullptr;
	guarded([&] { res = new hnswlib::BruteforceSearch(toMetric(metric), dim, capacity); });
	return res;
}
void* ref_bf_clone(const void* h, size_t newCapacity) {
	void* res = nullptr;
	guarded([&] { res = new hnswlib::BruteforceSearch(*static_cast<const hnswlib::BruteforceSearch*>(h), newCapacity); });
	return res;
}
void ref_bf_destroy(void* h) { delete static_cast<hnswlib::BruteforceSearch*>(h); }
size_t ref_bf_size(const void* h) { return static_cast<const hnswlib::BruteforceSearch*>(h)->CurrentElementCount(); }
size_t ref_bf_capacity(const void* h) { return static_cast<const hnswlib::BruteforceSearch*>(h)->MaxElements(); }
size_t ref_bf_element_size(const void* h) { return static_cast<const hnswlib::BruteforceSearch*>(h)->ElementSize(); }

int ref_bf_add(void* h, size_t dim, const float* vec, uint64_t label) {
	return guarded([&] {
		static_cast<hnswlib::BruteforceSearch*>(h)->AddPointNoLock(reindexer::ConstFloatVectorView{std::span<const float>{vec, dim}},
																	 reindexer::FloatVectorId::FromNumber(label));
	});
}
int ref_bf_add_batch(void* h, size_t dim, size_t n, const uint64_t* labels, const float* vecs) {
	return guarded([&] {
		auto* bf = static_cast<hnswlib::BruteforceSearch*>(h);
		for (size_t i = 0; i < n; ++i) {
			bf->AddPointNoLock(reindexer::ConstFloatVectorView{std::span<const float>{vecs + i * dim, dim}},
							   reindexer::FloatVectorId::FromNumber(labels[i]));
		}
	});
}
int ref_bf_remove(void* h, uint64_t label) {
	return guarded([&] { static_cast<hnswlib::BruteforceSearch*>(h)->RemovePoint(label); });
}
int ref_bf_resize(void* h, size_t newCapacity) {
	return guarded([&] { static_cast<hnswlib::BruteforceSearch*>(h)->ResizeIndex(newCapacity); });
}
int ref_bf_get(const void* h, uint64_t label, const float** row) {
	return guarded([&] { *row = static_cast<const hnswlib::BruteforceSearch*>(h)->FloatPtrByExternalLabel(label); });
}

// query must already be L2-normalised for Cosine (as HnswIndexBase::search does, hnsw_index.cc:166-171)
int64_t ref_bf_search_knn(const void* h, const float* query, size_t k, float* dists, uint64_t* labels) {
	int64_t n = -1;
	guarded([&] {
		auto q = static_cast<const hnswlib::BruteforceSearch*>(h)->SearchKnn(query, std::nullopt, k);
		n = int64_t(drain(q, k, dists, labels));
	});
	return n;
}
int64_t ref_bf_search_range(const void* h, const float* query, float radius, size_t maxOut, float* dists, uint64_t* labels) {
	int64_t n = -1;
	guarded([&] {
		auto q = static_cast<const hnswlib::BruteforceSearch*>(h)->SearchRange(query, std::nullopt, radius, 0);
		n = int64_t(drain(q, maxOut, dists, labels));
	});
	return n;
}
// nq independent single-threaded queries issued from `threads` host threads against one shared read-only index --
// exactly the concurrency the reference permits ("Read-only concurrency expected").  counts[i] = results of query i.
int ref_bf_search_knn_batch(const void* h, size_t dim, uint32_t nq, const float* queries, size_t k, int threads, float* dists,
							uint64_t* labels, uint32_t* counts) {
	return guarded([&] {
		auto* bf = static_cast<const hnswlib::BruteforceSearch*>(h);
		parallelQueries(nq, threads, [&](uint32_t i) {
			auto q = bf->SearchKnn(queries + size_t(i) * dim, std::nullopt, k);
			counts[i] = uint32_t(drain(q, k, dists + size_t(i) * k, labels + size_t(i) * k));
		});
	});
}

// ---------------------------------------------------------------- HNSW
// sync = 0: Synchronization::None (deterministic single-thread build, level RNG seeded with `seed`; the product uses 100, hnsw.h:73)
// sync = 1: Synchronization::OnInsertions (multithreaded build through AddPointConcurrent)
void* ref_hnsw_create(int metric, size_t dim, size_t capacity, size_t M, size_t efConstruction, size_t seed, int sync) {
	HnswHandle* res = nullptr;
	guarded([&] {
		auto h = std::make_unique<HnswHandle>();
		h->dim = dim;
		h->M = M;
		if (sync) {
			h->mt = std::make_unique<HnswMT>(toMetric(metric), dim, capacity, M, efConstruction, seed, reindexer::ReplaceDeleted_True);
		} else {
			h->st = std::make_unique<HnswST>(toMetric(metric), dim, capacity, M, efConstruction, seed, reindexer::ReplaceDeleted_True);
		}
		res = h.release();
	});
	return res;
}
void ref_hnsw_destroy(void* h) { delete static_cast<HnswHandle*>(h); }
size_t ref_hnsw_size(const void* h) {
	return withHnsw(const_cast<HnswHandle*>(static_cast<const HnswHandle*>(h)), [](auto& g) { return g.CurrentElementCount(); });
}

int ref_hnsw_add_batch(void* hv, size_t n, const uint64_t* labels, const float* vecs, int threads) {
	auto* h = static_cast<HnswHandle*>(hv);
	return guarded([&] {
		if (h->st) {
			for (size_t i = 0; i < n; ++i) {
				h->st->AddPointNoLock(vecs + i * h->dim, labels[i]);
			}
			return;
		}
		std::atomic<size_t> next{0};
		std::atomic<bool> failed{false};
		std::string err;
		reindexer::mutex errMtx;
		std::vector<std::thread> pool;
		for (int t = 0; t < std::max(threads, 1); ++t) {
			pool.emplace_back([&] {
				try {
					for (size_t i = next.fetch_add(1); i < n && !failed; i = next.fetch_add(1)) {
						h->mt->AddPointConcurrent(vecs + i * h->dim, labels[i]);
					}
				} catch (const std::exception& e) {
					reindexer::lock_guard lck(errMtx);
					failed = true;
					err = e.what();
				}
			});
		}
		for (auto& th : pool) {
			th.join();
		}
		if (failed) {
			throw std::runtime_error(err);
		}
	});
}

// query pre-normalised for Cosine; qnorm = ||q|| (hnsw_index.cc:169), ignored (nullopt) when has_norm == 0
int64_t ref_hnsw_search_knn(const void* hv, const float* query, int has_norm, float qnorm, size_t k, size_t ef, float* dists,
							uint64_t* labels) {
	int64_t n = -1;
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	guarded([&] {
		withHnsw(h, [&](auto& g) {
			auto q = g.SearchKnn(query, has_norm ? std::optional<float>(qnorm) : std::nullopt, k, ef);
			n = int64_t(drain(q, k, dists, labels));
		});
	});
	return n;
}
// ---- SQ8: the reference's scalar quantisation of an HNSW graph (groundwork for SURVEY §8 f2) ------------------------------------
// HierarchicalNSW::Impl::Quantize (hnsw.cc:133-139) builds HierarchicalNSWImpl<uint8_t> from the float graph: codes
// u8 = clamp((x - minQ) / alpha, 0, 255) (scalar_quantization/quantizer.h:64-66), one corrective offset per vector (:93-125),
// distances alpha^2 * int_dist(u8, u8) + offset_l + offset_r (hnswlib.h:192-197).  Single-threaded graphs only.
using HnswQ = hnswlib::HierarchicalNSWImpl<uint8_t, hnswlib::Synchronization::None>;
struct QuantHandle {
	std::unique_ptr<HnswQ> q;
	size_t dim = 0;
};
void* ref_hnsw_quantize(const void* hv, int has_quantile, float quantile, size_t sample_size) {
	auto* h = static_cast<const HnswHandle*>(hv);
	QuantHandle* out = nullptr;
	guarded([&] {
		if (!h->st) {
			throw std::runtime_error("ref_hnsw_quantize: single-threaded graph expected");
		}
		hnswlib::QuantizationConfig cfg;
		if (has_quantile) {
			cfg.quantile = quantile;
		}
		if (sample_size) {
			cfg.sampleSize = sample_size;
		}
		auto qh = std::make_unique<QuantHandle>();
		qh->dim = h->dim;
		qh->q = std::make_unique<HnswQ>(*h->st, h->st->MaxElements(), std::optional(cfg));
		out = qh.release();
	});
	return out;
}
void ref_hnsw_q_destroy(void* qv) { delete static_cast<QuantHandle*>(qv); }
// params: [minQ, maxQ, alpha, alpha_2, delta]
int ref_hnsw_q_params(const void* qv, float* params) {
	auto* q = static_cast<const QuantHandle*>(qv);
	return guarded([&] {
		const auto& p = q->q->quantizer_->Params();
		params[0] = p.minQ;
		params[1] = p.maxQ;
		params[2] = p.alpha;
		params[3] = p.alpha_2;
		params[4] = p.delta;
	});
}
// codes [n][dim] u8 and corrective offsets [n] by internal id
int ref_hnsw_q_export(const void* qv, uint8_t* codes, float* offsets) {
	auto* q = static_cast<const QuantHandle*>(qv);
	return guarded([&] {
		const size_t n = q->q->CurrentElementCount();
		for (size_t i = 0; i < n; ++i) {
			std::memcpy(codes + i * q->dim, q->q->getDataByInternalId(hnswlib::tableint(i)), q->dim);
			offsets[i] = q->q->fstdistfunc_.Sq8CorrectiveOffsets()[i];
		}
	});
}
// the quantised query exactly as search() prepares it (prepareData, hnswalg.h:510-535): codes [dim] + its corrective offset
int ref_hnsw_q_prepare_query(const void* qv, const float* query, uint8_t* codes, float* offset) {
	auto* q = static_cast<const QuantHandle*>(qv);
	return guarded([&] {
		auto holder = q->q->prepareData(query, 1.f);
		std::memcpy(codes, holder.get(), q->dim);
		std::memcpy(offset, holder.get() + q->dim, sizeof(float));
	});
}
int64_t ref_hnsw_q_search_knn(const void* qv, const float* query, int has_norm, float qnorm, size_t k, size_t ef, float* dists,
							  uint64_t* labels) {
	auto* q = static_cast<const QuantHandle*>(qv);
	int64_t n = -1;
	guarded([&] {
		auto res = q->q->SearchKnn(query, has_norm ? std::optional<float>(qnorm) : std::nullopt, k, ef);
		n = int64_t(drain(res, k, dists, labels));
	});
	return n;
}

// Streaming search (hnswalg.h:1864-1975): BeginStreamingSearch once, then ContinueStreamingSearch(batchSize) until exhausted.
// The session borrows the graph: destroy it before the graph.
struct StreamHandle {
	hnswlib::StreamingSearchSession session;
	std::vector<float> query;  // the session keeps a pointer into the caller's query for non-quantised graphs: keep it alive
};
void* ref_hnsw_stream_begin(const void* hv, const float* query, int has_norm, float qnorm, size_t ef) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	StreamHandle* out = nullptr;
	guarded([&] {
		withHnsw(h, [&](auto& g) {
			std::vector<float> q(query, query + h->dim);
			const float* qp = q.data();  // std::vector's buffer survives the move below
			auto sess = g.BeginStreamingSearch(qp, has_norm ? std::optional<float>(qnorm) : std::nullopt, hnswlib::StreamingSearchOptions{ef});
			out = new StreamHandle{std::move(sess), std::move(q)};
		});
	});
	return out;
}
// returns the number of results of this batch (best first), -1 on error; *exhausted = 1 when the search has nothing more to give
int64_t ref_hnsw_stream_next(const void* hv, void* sv, size_t batchSize, float* dists, uint64_t* labels, int* exhausted) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	auto* s = static_cast<StreamHandle*>(sv);
	int64_t n = -1;
	guarded([&] {
		withHnsw(h, [&](auto& g) {
			auto batch = g.ContinueStreamingSearch(s->session, batchSize);
			*exhausted = batch.exhausted ? 1 : 0;
			n = int64_t(drain(batch.results, batchSize, dists, labels));
		});
	});
	return n;
}
void ref_hnsw_stream_end(void* sv) { delete static_cast<StreamHandle*>(sv); }

int ref_hnsw_mark_delete(void* hv, uint64_t label) {
	auto* h = static_cast<HnswHandle*>(hv);
	return guarded([&] { withHnsw(h, [&](auto& g) { g.MarkDelete(label); }); });
}

// HierarchicalNSWImpl::SearchRange (hnswalg.h:2015-2070): ef-search seed, then BFS over neighbours with dist < radius
int64_t ref_hnsw_search_range(const void* hv, const float* query, int has_norm, float qnorm, float radius, size_t ef, size_t maxOut,
							  float* dists, uint64_t* labels) {
	int64_t n = -1;
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	guarded([&] {
		withHnsw(h, [&](auto& g) {
			auto q = g.SearchRange(query, has_norm ? std::optional<float>(qnorm) : std::nullopt, radius, ef);
			n = int64_t(drain(q, maxOut, dists, labels));
		});
	});
	return n;
}
int ref_hnsw_search_knn_batch(const void* hv, uint32_t nq, const float* queries, const float* qnorms, size_t k, size_t ef, int threads,
							  float* dists, uint64_t* labels, uint32_t* counts) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	return guarded([&] {
		withHnsw(h, [&](auto& g) {
			parallelQueries(nq, threads, [&](uint32_t i) {
				auto q = g.SearchKnn(queries + size_t(i) * h->dim, qnorms ? std::optional<float>(qnorms[i]) : std::nullopt, k, ef);
				counts[i] = uint32_t(drain(q, k, dists + size_t(i) * k, labels + size_t(i) * k));
			});
		});
	});
}

// The reference's own work counters (hnswalg.h:250-251): level-0 search re-run with collect_metrics=true.
int ref_hnsw_search_metrics(const void* hv, const float* query, int has_norm, float qnorm, size_t ef, int64_t* distComps, int64_t* hops) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	return guarded([&] {
		withHnsw(h, [&](auto& g) {
			g.metric_distance_computations = 0;
			g.metric_hops = 0;
			(void)has_norm, (void)qnorm;  // queryNormCoef() is 1 for a non-quantised graph (hnswalg.h:1855-1863)
			const auto ep = g.getLayer0EntryPoint(query, 1.f);
			if (g.DeletedCountUnsafe() == 0) {  // the branch search() takes (hnswalg.h:1982)
				auto top = g.template searchBaseLayerST<true, true>(ep, query, 1.f, ef);
			} else {
				auto top = g.template searchBaseLayerST<false, true>(ep, query, 1.f, ef);
			}
			*distComps = g.metric_distance_computations.load();
			*hops = g.metric_hops.load();
		});
	});
}

// Graph export for the device upload (layout facts: hnswalg.h:221-228, 1034-1040, 1372):
//   header[0]=cur_element_count header[1]=maxlevel header[2]=enterpoint header[3]=M header[4]=maxM0 header[5]=total upper-level slots
int ref_hnsw_export_header(const void* hv, int64_t* header) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	return guarded([&] {
		withHnsw(h, [&](auto& g) {
			const size_t n = g.cur_element_count;
			int64_t upper = 0;
			for (size_t i = 0; i < n; ++i) {
				upper += g.element_levels_[i];
			}
			header[0] = int64_t(n);
			header[1] = g.maxlevel_;
			header[2] = int64_t(g.enterpoint_node_);
			header[3] = int64_t(g.M_);
			header[4] = int64_t(g.maxM0_);
			header[5] = upper;
		});
	});
}
// level0: n x (1 + maxM0) u32  [count | neighbours...]; levels: n i32; upperOffsets: n+1 i64 (slot index of element's level-1 list);
// upper: totalUpper x (1 + M) u32 [count | neighbours...] (levels 1..L of one element are consecutive); labels: n u64; vectors: n x dim f32
int ref_hnsw_export(const void* hv, uint32_t* level0, int32_t* levels, int64_t* upperOffsets, uint32_t* upper, uint64_t* labels,
					float* vectors) {
	auto* h = const_cast<HnswHandle*>(static_cast<const HnswHandle*>(hv));
	return guarded([&] {
		withHnsw(h, [&](auto& g) {
			const size_t n = g.cur_element_count;
			const size_t m0 = g.maxM0_, m = g.M_;
			int64_t slot = 0;
			for (size_t i = 0; i < n; ++i) {
				const auto* ll0 = g.get_linklist0(hnswlib::tableint(i));
				const unsigned cnt0 = g.getListCount(ll0);
				uint32_t* dst0 = level0 + i * (1 + m0);
				dst0[0] = cnt0;
				for (size_t j = 0; j < m0; ++j) {
					dst0[1 + j] = j < cnt0 ? hnswlib::readLinkListNeighbor(ll0, j) : 0u;
				}
				levels[i] = g.element_levels_[i];
				upperOffsets[i] = slot;
				for (int lvl = 1; lvl <= g.element_levels_[i]; ++lvl, ++slot) {
					const auto* ll = g.get_linklist(hnswlib::tableint(i), lvl);
					const unsigned cnt = g.getListCount(ll);
					uint32_t* dst = upper + slot * (1 + m);
					dst[0] = cnt;
					for (size_t j = 0; j < m; ++j) {
						dst[1 + j] = j < cnt ? hnswlib::readLinkListNeighbor(ll, j) : 0u;
					}
				}
				labels[i] = g.ExternalLabel(hnswlib::tableint(i));
				if (vectors) {
					std::memcpy(vectors + i * h->dim, g.getDataByInternalId(hnswlib::tableint(i)), h->dim * sizeof(float));
				}
			}
			upperOffsets[n] = slot;
		});
	});
}

}  // extern "C"
