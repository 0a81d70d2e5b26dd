// ORACLE / TEST INFRASTRUCTURE ONLY.
// extern "C" facade over the reference's IVF index engine: the FAISS copy vendored (and patched) by the reference under
// /root/reference/cpp_src/vendor_subdirs/faiss, compiled in place (see oracle/Makefile), driven exactly like
// reindexer::IvfIndex drives it (cpp_src/core/index/float_vector/ivf_index.cc):
//   construction        :97-99    IndexFlatL2 / IndexFlatIP quantizer + faiss::IndexIVFFlat(space, dim, nCentroids, metric, isCosine)
//   training + filling  :469-484, :101-102   set_direct_map_type(Hashtable), train(n, vecs, norms), add_with_ids(n, vecs, norms, ids)
//   search              :150-204  map->search(1, key, k, dists, ids, &IVFSearchParameters{nprobe})
// plus read access to the trained state (centroids, inverted lists) so that the device index can be filled with the same lists.
// The reference loads a BLAS with dlopen at run time (tools/blas_extension.cc); there is none in this image, so sgemm_dlwrp_ is
// provided here as a plain triple loop (used by k-means training only) and the LAPACK entry points FAISS' unused transforms
// reference abort if they are ever called.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "faiss/IndexFlat.h"
#include "faiss/IndexIVFFlat.h"
#include "faiss/impl/AuxIndexStructures.h"
#include "faiss/invlists/DirectMap.h"
#include "tools/normalize.h"

namespace {
thread_local std::string g_err;

struct IvfHandle {
	std::unique_ptr<faiss::IndexFlat> space;
	std::unique_ptr<faiss::IndexIVFFlat> map;
	size_t dim = 0;
	int metric = 0;
};

template <typename Fn>
int guarded(Fn&& fn) noexcept {
	try {
		fn();
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
	} catch (...) {
		g_err = "unknown exception";
	}
	return 1;
}
}  // namespace

extern "C" {

const char* ref_ivf_last_error() { return g_err.c_str(); }

// metric: 0 = L2, 1 = InnerProduct, 2 = Cosine (reindexer::VectorMetric)
void* ref_ivf_create(int metric, size_t dim, size_t nlist) {
	IvfHandle* h = nullptr;
	guarded([&] {
		auto hh = std::make_unique<IvfHandle>();
		hh->dim = dim;
		hh->metric = metric;
		if (metric == 0) {
			hh->space = std::make_unique<faiss::IndexFlatL2>(dim);
		} else if (metric == 1) {
			hh->space = std::make_unique<faiss::IndexFlatIP>(dim);
		} else if (metric == 2) {
			hh->space = std::make_unique<faiss::IndexFlatCosine>(dim);  // IvfIndex::newSpace, ivf_index.cc:686-695
		} else {
			throw std::runtime_error("ref_ivf_create: unknown metric");
		}
		hh->map = std::make_unique<faiss::IndexIVFFlat>(hh->space.get(), dim, nlist, metric == 0 ? faiss::METRIC_L2 : faiss::METRIC_INNER_PRODUCT,
														 metric == 2);
		h = hh.release();
	});
	return h;
}
void ref_ivf_destroy(void* h) { delete static_cast<IvfHandle*>(h); }

int ref_ivf_train_add(void* hv, size_t n, const float* vecs, const int64_t* ids) {
	auto* h = static_cast<IvfHandle*>(hv);
	return guarded([&] {
		h->map->set_direct_map_type(faiss::DirectMap::Type::Hashtable);
		std::vector<float> norms;  // Cosine: the norm coefficients IvfIndex hands over (space_->get_xb_norms(), ivf_index.cc:101-102)
		if (h->metric == 2) {
			norms.resize(n);
			for (size_t i = 0; i < n; ++i) {
				norms[i] = reindexer::ann::CalculateL2Module(vecs + i * h->dim, int32_t(h->dim));
			}
		}
		const float* np = norms.empty() ? nullptr : norms.data();
		h->map->train(faiss::idx_t(n), vecs, np);
		h->map->add_with_ids(faiss::idx_t(n), vecs, np, reinterpret_cast<const faiss::idx_t*>(ids));
	});
}

// IvfIndex::upsert on a trained index: map_->add_with_ids(1, vect, &id) (ivf_index.cc:87-91); Cosine norms like train_add
int ref_ivf_add(void* hv, size_t n, const float* vecs, const int64_t* ids) {
	auto* h = static_cast<IvfHandle*>(hv);
	return guarded([&] {
		for (size_t i = 0; i < n; ++i) {
			const faiss::idx_t id = ids[i];
			if (h->metric == 2) {
				const float norm = reindexer::ann::CalculateL2Module(vecs + i * h->dim, int32_t(h->dim));
				h->map->add_with_ids(1, vecs + i * h->dim, &norm, &id);
			} else {
				h->map->add_with_ids(1, vecs + i * h->dim, &id);
			}
		}
	});
}
// IvfIndex::del: map_->remove_ids(IDSelectorArray{1, &id}) (ivf_index.cc:120-124)
int ref_ivf_remove(void* hv, int64_t id) {
	auto* h = static_cast<IvfHandle*>(hv);
	return guarded([&] {
		const faiss::idx_t fid = id;
		h->map->remove_ids(faiss::IDSelectorArray{1, &fid});
	});
}
// the list every id lives in, from the direct map (the coarse quantiser's assignment at add time)
int ref_ivf_list_of(const void* hv, size_t n, const int64_t* ids, uint32_t* list_nos) {
	auto* h = static_cast<const IvfHandle*>(hv);
	return guarded([&] {
		for (size_t i = 0; i < n; ++i) {
			const auto it = h->map->direct_map.hashtable.find(ids[i]);
			if (it == h->map->direct_map.hashtable.end()) {
				throw std::runtime_error("ref_ivf_list_of: unknown id");
			}
			list_nos[i] = uint32_t(faiss::lo_listno(it->second));
		}
	});
}

// returns the number of results (ids >= 0), best first as FAISS returns them; distances in FAISS' convention (L2: squared distance
// ascending, IP: inner product descending)
int64_t ref_ivf_search(const void* hv, const float* query, size_t k, size_t nprobe, float* dists, int64_t* ids) {
	auto* h = static_cast<const IvfHandle*>(hv);
	int64_t n = -1;
	guarded([&] {
		faiss::IVFSearchParameters p;
		p.nprobe = nprobe;
		h->map->search(1, query, faiss::idx_t(k), dists, reinterpret_cast<faiss::idx_t*>(ids), &p);
		n = 0;
		while (size_t(n) < k && ids[n] >= 0) {
			++n;
		}
	});
	return n;
}

// nq queries in one call: FAISS parallelises over queries with OpenMP (the most favourable way to run the reference's engine)
int ref_ivf_search_batch(const void* hv, size_t nq, const float* queries, size_t k, size_t nprobe, float* dists, int64_t* ids) {
	auto* h = static_cast<const IvfHandle*>(hv);
	return guarded([&] {
		faiss::IVFSearchParameters p;
		p.nprobe = nprobe;
		h->map->search(faiss::idx_t(nq), queries, faiss::idx_t(k), dists, reinterpret_cast<faiss::idx_t*>(ids), &p);
	});
}

// map_->range_search(1, key, radius, &result, &params) (ivf_index.cc:211-212); returns the number of results, writes at most maxOut
// (unsorted, as FAISS returns them); radius in FAISS' convention (L2: dis < radius, IP / Cosine: dis > radius)
int64_t ref_ivf_range_search(const void* hv, const float* query, float radius, size_t nprobe, size_t maxOut, float* dists, int64_t* ids) {
	auto* h = static_cast<const IvfHandle*>(hv);
	int64_t n = -1;
	guarded([&] {
		faiss::IVFSearchParameters p;
		p.nprobe = nprobe;
		faiss::RangeSearchResult res(1);
		h->map->range_search(1, query, radius, &res, &p);
		n = int64_t(res.lims[1] - res.lims[0]);
		for (int64_t i = 0; i < n && size_t(i) < maxOut; ++i) {
			dists[i] = res.distances[res.lims[0] + i];
			ids[i] = res.labels[res.lims[0] + i];
		}
	});
	return n;
}

// hdr: [nlist, ntotal]
int ref_ivf_export_header(const void* hv, int64_t* hdr) {
	auto* h = static_cast<const IvfHandle*>(hv);
	return guarded([&] {
		hdr[0] = int64_t(h->map->nlist);
		hdr[1] = int64_t(h->map->ntotal);
	});
}
// centroids [nlist][dim]; list_sizes [nlist]; ids / vecs concatenated list by list, in list order
int ref_ivf_export(const void* hv, float* centroids, int64_t* list_sizes, int64_t* ids, float* vecs) {
	auto* h = static_cast<const IvfHandle*>(hv);
	return guarded([&] {
		std::memcpy(centroids, h->space->get_xb(), h->map->nlist * h->dim * sizeof(float));
		size_t at = 0;
		for (size_t l = 0; l < h->map->nlist; ++l) {
			const size_t sz = h->map->invlists->list_size(l);
			list_sizes[l] = int64_t(sz);
			faiss::InvertedLists::ScopedCodes codes(h->map->invlists, l);
			faiss::InvertedLists::ScopedIds lids(h->map->invlists, l);
			std::memcpy(vecs + at * h->dim, codes.get(), sz * h->dim * sizeof(float));
			std::memcpy(ids + at, lids.get(), sz * sizeof(int64_t));
			at += sz;
		}
	});
}

// ---- BLAS / LAPACK entry points of tools/blas_extension.cc (dlopen of a system BLAS in the reference) ---------------------------
// column-major sgemm: C = alpha * op(A) * op(B) + beta * C
int sgemm_dlwrp_(const char* transa, const char* transb, int* m, int* n, int* k, const float* alpha, const float* a, int* lda, const float* b,
				 int* ldb, float* beta, float* c, int* ldc) {
	const bool ta = *transa == 'T' || *transa == 't', tb = *transb == 'T' || *transb == 't';
#pragma omp parallel for
	for (int j = 0; j < *n; ++j) {
		for (int i = 0; i < *m; ++i) {
			double s = 0;
			for (int l = 0; l < *k; ++l) {
				const float av = ta ? a[size_t(i) * *lda + l] : a[size_t(l) * *lda + i];
				const float bv = tb ? b[size_t(l) * *ldb + j] : b[size_t(j) * *ldb + l];
				s += double(av) * bv;
			}
			float& out = c[size_t(j) * *ldc + i];
			out = float(*alpha * s + (*beta != 0.f ? *beta * out : 0.f));
		}
	}
	return 0;
}
#define RX_ORACLE_BLAS_STUB(name)                                                      \
	void name() {                                                                      \
		std::fprintf(stderr, "oracle: " #name " is not available in this build\n");   \
		std::abort();                                                                  \
	}
RX_ORACLE_BLAS_STUB(sgeqrf_dlwrp_)
RX_ORACLE_BLAS_STUB(sorgqr_dlwrp_)
RX_ORACLE_BLAS_STUB(ssyev_dlwrp_)
RX_ORACLE_BLAS_STUB(dsyev_dlwrp_)
RX_ORACLE_BLAS_STUB(sgesvd_dlwrp_)
RX_ORACLE_BLAS_STUB(dgesvd_dlwrp_)
RX_ORACLE_BLAS_STUB(sgelsd_dlwrp_)
RX_ORACLE_BLAS_STUB(dgemm_dlwrp_)
RX_ORACLE_BLAS_STUB(ssyrk_dlwrp_)
RX_ORACLE_BLAS_STUB(sgetrf_dlwrp_)
RX_ORACLE_BLAS_STUB(sgetri_dlwrp_)
RX_ORACLE_BLAS_STUB(dgetri_dlwrp_)
RX_ORACLE_BLAS_STUB(dgetrf_dlwrp_)
RX_ORACLE_BLAS_STUB(sgemv_dlwrp_)

}  // extern "C"
