// Internal declarations shared by the translation units of librxgpu (index.cu, hnsw.cu, ft_bm25.cu).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rxgpu.h"
#include "../host/flat_map.h"

namespace rxgpu {

extern thread_local std::string g_err;
extern thread_local rxgpu_search_stats g_stats;
extern thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
extern std::atomic<int> g_profile;

// sum the event pairs recorded by this thread's launches; call after the stream has been synchronised
inline void collectProfile() {
	for (auto& ev : g_prof_events) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
			g_stats.scan_kernel_ms += ms;
			g_stats.scan_launches += 1;
		}
		cudaEventDestroy(ev.first);
		cudaEventDestroy(ev.second);
	}
	g_prof_events.clear();
}

inline int fail(int code, std::string msg) {
	g_err = std::move(msg);
	return code;
}

#define RX_CUDA(expr)                                                                                        \
	do {                                                                                                     \
		cudaError_t e_ = (expr);                                                                             \
		if (e_ != cudaSuccess) {                                                                             \
			return fail(RXGPU_ERR_SYSTEM, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr); \
		}                                                                                                    \
	} while (0)

template <typename T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	~DevBuf() { release(); }
	void release() {
		if (p) {
			cudaFree(p);
			p = nullptr;
			n = 0;
		}
	}
	cudaError_t ensure(size_t want) {
		if (want <= n) {
			return cudaSuccess;
		}
		release();
		cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
		if (e == cudaSuccess) {
			n = want;
		}
		return e;
	}
};
template <typename T>
struct PinBuf {
	T* p = nullptr;
	size_t n = 0;
	~PinBuf() {
		if (p) {
			cudaFreeHost(p);
		}
	}
	cudaError_t ensure(size_t want) {
		if (want <= n) {
			return cudaSuccess;
		}
		if (p) {
			cudaFreeHost(p);
			p = nullptr;
			n = 0;
		}
		cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), want * sizeof(T));
		if (e == cudaSuccess) {
			n = want;
		}
		return e;
	}
};

// The dynamic shared-memory ceiling of a kernel is a per-device FUNCTION attribute: searches run concurrently from many threads
// (the reference's read-side concurrency), so it is raised once per (kernel, device) to the budget every caller stays within --
// never per launch, where a thread asking for less would lower it under another thread's launch (cudaErrorInvalidValue).
constexpr int kScanSmemBudget = 100 * 1024;
inline cudaError_t raiseSmemCeilingOnceImpl(const void* fn, int device, int bytes) {
	static std::mutex mtx;
	static std::vector<std::pair<const void*, int>> done;  // (kernel entry point, device); a handful of entries
	std::lock_guard<std::mutex> lck(mtx);
	for (const auto& d : done) {
		if (d.first == fn && d.second == device) {
			return cudaSuccess;
		}
	}
	const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if (e == cudaSuccess) {
		done.emplace_back(fn, device);
	}
	return e;
}
template <typename Kernel>
inline cudaError_t raiseSmemCeilingOnce(Kernel kfn, int device, int bytes) {  // keyed by the entry point, not by its type
	return raiseSmemCeilingOnceImpl(reinterpret_cast<const void*>(kfn), device, bytes);
}

// per-call scratch: searches are re-entrant, each takes one workspace from the pool
struct Workspace {
	cudaStream_t stream = nullptr;
	cudaStream_t tail_stream = nullptr;  // the filter's tail grid (CTA pairs on the SMs the cluster-of-4 grid strands) runs beside the main one
	cudaEvent_t tail_fork = nullptr, tail_join = nullptr;
	DevBuf<float> d_queries;
	DevBuf<uint64_t> d_lists;
	DevBuf<uint64_t> d_floor;  // per-query floor keys between the rounds of a k > 255 search
	DevBuf<float> d_out_dist;
	DevBuf<uint32_t> d_out_idx;
	DevBuf<uint64_t> d_out_label;
	DevBuf<uint32_t> d_out_count;
	DevBuf<uint64_t> d_range;
	DevBuf<unsigned long long> d_range_count;
	DevBuf<uint16_t> d_qbf;  // bf16 query block for the tensor-core filter
	DevBuf<float> d_qnorm;
	DevBuf<unsigned int> d_tau, d_cand_count, d_ub_lock;
	DevBuf<float> d_ub_list;
	DevBuf<uint32_t> d_cand_rows;
	PinBuf<unsigned int> h_cand_count;
	PinBuf<float> h_queries;
	PinBuf<float> h_out_dist;
	PinBuf<uint32_t> h_out_idx;
	PinBuf<uint64_t> h_out_label;
	PinBuf<uint32_t> h_out_count;
	PinBuf<uint64_t> h_range;
	// state of the last scanTopK on this workspace: when the tensor-core filter answered it, the per-query candidate lists
	// (d_cand_rows / h_cand_count) hold every row at or below each query's k1-th distance -- tieRowsAfterScan reads them
	bool tc_lists_valid = false;
	uint32_t tc_lists_nq = 0;
	uint64_t tc_lists_version = 0;
	DevBuf<uint32_t> d_sel;
	DevBuf<float> d_selbound;
	DevBuf<float> d_tie_dist;
	DevBuf<uint32_t> d_tie_idx, d_tie_count;
	DevBuf<uint64_t> d_tie_label;
	~Workspace() {
		if (stream) {
			cudaStreamDestroy(stream);
		}
		if (tail_stream) {
			cudaStreamDestroy(tail_stream);
			cudaEventDestroy(tail_fork);
			cudaEventDestroy(tail_join);
		}
	}
};

}  // namespace rxgpu

struct rxgpu_hnsw_device;  // hnsw.cu
struct rxgpu_ivf_device;   // index.cu
struct rxgpu_sq8_device;   // sq8.cu
namespace rxgpu {
void hnswRelease(rxgpu_hnsw_device*);
void ivfRelease(rxgpu_ivf_device*);
void sq8Release(rxgpu_sq8_device*);
}

struct rxgpu_index {
	int metric = 0;
	uint32_t dim = 0;
	uint32_t pitch = 0;  // floats, multiple of 4
	uint64_t capacity = 0;
	uint64_t size = 0;
	int device = 0;
	uint32_t flags = 0;
	int sm_count = 148;
	uint32_t qt_override = 0;
	uint64_t version = 0;  // bumped by every mutation of rows/labels (staleness check of attached structures)

	float* d_rows = nullptr;
	uint64_t* d_labels = nullptr;
	float* d_norms = nullptr;  // Cosine only (DistCalculator::normCoefs_, hnswlib.h:33-35)

	std::vector<uint64_t> h_labels;  // by internal index
	rxgpu::LabelMap dict;
	std::vector<float> h_rows;  // optional host mirror [capacity][dim]

	cudaStream_t stream = nullptr;  // maintenance stream
	rxgpu::DevBuf<float> st_rows;   // staging of scattered upserts (mutators run one at a time under the namespace write lock)
	rxgpu::DevBuf<uint32_t> st_dst;
	rxgpu::DevBuf<uint64_t> st_labels;
	mutable std::mutex ws_mtx;
	mutable std::vector<std::unique_ptr<rxgpu::Workspace>> ws_free;
	rxgpu_hnsw_device* hnsw = nullptr;  // graph attached by rxgpu_hnsw_import (hnsw.cu)
	rxgpu_ivf_device* ivf = nullptr;    // centroids + list boundaries attached by rxgpu_ivf_import
	rxgpu_sq8_device* sq8 = nullptr;    // SQ8 codes + corrective offsets attached by rxgpu_sq8_attach (sq8.cu)

	// tensor-core filter state, built lazily by the first large-batch search: bf16 shadow of the rows + row norms
	mutable std::mutex tc_mtx;
	mutable void* d_shadow = nullptr;  // __nv_bfloat16 [capacity][pitch_bf]
	mutable float* d_vnorm = nullptr;  // [capacity] ||row||_2
	mutable float2* d_vw = nullptr;    // [capacity] per-row (max(||row||, tiny), w) pairs the filter epilogue consumes, 512 B per 64-row tile
	mutable uint32_t pitch_bf = 0;
	mutable uint64_t shadow_version = ~0ull;
	// rows rewritten since the shadow was last brought up to date (the mutations log them beside `version`): ensureShadow converts
	// only these; the log gives up (full rebuild) beyond kShadowLogMax ranges
	static constexpr size_t kShadowLogMax = 4096;
	mutable std::vector<std::pair<uint32_t, uint32_t>> shadow_dirty;
	mutable bool shadow_dirty_all = true;
	void touchRows(uint64_t begin, uint64_t end) {
		std::lock_guard<std::mutex> lck(tc_mtx);
		if (!d_shadow || shadow_dirty_all || begin >= end) {
			return;
		}
		if (!shadow_dirty.empty() && shadow_dirty.back().second == begin) {
			shadow_dirty.back().second = uint32_t(end);
		} else if (shadow_dirty.size() < kShadowLogMax) {
			shadow_dirty.emplace_back(uint32_t(begin), uint32_t(end));
		} else {
			shadow_dirty_all = true;
			shadow_dirty.clear();
		}
	}
	uint32_t tc_mode = 0;  // 0 auto, 1 force on, 2 off
	uint32_t tc_tail = 1;         // 1 = a tail grid of 2-CTA clusters scans a slice of the rows on the SMs the main grid cannot use
	uint32_t tc_variant = 0;      // 0 = knn_tc_filter_q (query block in TMEM) when the dimension allows; 14 = knn_tc_filter_p (CTA pairs, cta_group::2); 3 / 4 = first-generation kernel (1 CTA / CTA pair)
	uint32_t tc_cluster_max = 0;  // 0 = up to 4 CTAs per cluster

	~rxgpu_index() {
		cudaSetDevice(device);
		ws_free.clear();
		if (hnsw) {
			rxgpu::hnswRelease(hnsw);
		}
		if (ivf) {
			rxgpu::ivfRelease(ivf);
		}
		if (sq8) {
			rxgpu::sq8Release(sq8);
		}
		if (d_rows) {
			cudaFree(d_rows);
		}
		if (d_labels) {
			cudaFree(d_labels);
		}
		if (d_norms) {
			cudaFree(d_norms);
		}
		if (d_shadow) {
			cudaFree(d_shadow);
		}
		if (d_vnorm) {
			cudaFree(d_vnorm);
		}
		if (d_vw) {
			cudaFree(d_vw);
		}
		if (stream) {
			cudaStreamDestroy(stream);
		}
	}
};

namespace rxgpu {

struct WsLease {
	const rxgpu_index* idx;
	std::unique_ptr<Workspace> ws;
	explicit WsLease(const rxgpu_index* i) : idx(i) {
		{
			std::lock_guard<std::mutex> lck(idx->ws_mtx);
			if (!idx->ws_free.empty()) {
				ws = std::move(idx->ws_free.back());
				idx->ws_free.pop_back();
			}
		}
		if (!ws) {
			ws = std::make_unique<Workspace>();
		}
	}
	~WsLease() {
		std::lock_guard<std::mutex> lck(idx->ws_mtx);
		idx->ws_free.emplace_back(std::move(ws));
	}
};

// index.cu -- shared with shard.cu
// Top-k1 rows per query under (dist, internal row) [kModeTopK], or the first k1 rows in internal order with dist <= bound
// [kModeTieRows, one query]; large batches go through the tensor-core filter + exact re-rank (same bits).  Device in / out.
int scanTopK(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nq, uint32_t k1, int mode, float bound,
			 float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count);
// After a scanTopK(kModeTopK) of `d_queries` on the SAME workspace: for the selected queries sel[i] the first k rows in internal order
// with dist <= dstar[i] (what the reference's tie rule needs, SURVEY.md 8a rule 2), [nsel][k] device outputs.  Served from the
// filter's candidate lists when they exist (no second pass over the rows), else by one kModeTieRows scan per selected query.
int tieRowsAfterScan(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nsel, const uint32_t* sel,
					 const float* dstar, uint32_t k, float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count);

// writes row `idx` (== size: appends) with a new vector and label, keeping the label dictionary consistent (index.cu)
int setRowAt(rxgpu_index* ix, uint32_t idx, uint64_t label, const float* vec);

inline int checkIndex(const rxgpu_index* ix) {
	if (!ix) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null index handle");
	}
	RX_CUDA(cudaSetDevice(ix->device));
	return 0;
}

// ---- collectives over an rxgpu_comm (shard.cu): NCCL between processes, a host rendezvous between the threads of one process -------
enum class CommOp { SumU64, SumU32, MaxU32 };
int commRank(const rxgpu_comm*);
int commSize(const rxgpu_comm*);
int commDevice(const rxgpu_comm*);
std::mutex& commMutex(rxgpu_comm*);
int commAllReduce(rxgpu_comm*, void* d_buf, size_t count, CommOp op, cudaStream_t st);        // in place
int commAllGather(rxgpu_comm*, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st);  // d_recv: nranks x bytes

}  // namespace rxgpu
