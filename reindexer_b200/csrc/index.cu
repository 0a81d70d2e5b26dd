// librxgpu: C ABI (include/rxgpu.h) over the sm_100a brute-force float_vector kernels.
// Host logic mirrors hnswlib::BruteforceSearch (cpp_src/core/index/float_vector/hnswlib/bruteforce.{h,cc}) and the
// search/select wrappers of HnswIndexBase<Map> (cpp_src/core/index/float_vector/hnsw_index.cc:160-288).
// There is no CPU fallback anywhere in this file: without a usable CUDA device every compute entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/rxgpu.h"
#include "internal.h"
#include "../host/knn_select.h"
#include "knn_scan.cuh"
#include "knn_tc.cuh"
#include "knn_tc_q.cuh"
#include "knn_tc_p.cuh"

using namespace rxgpu;

namespace rxgpu {
thread_local std::string g_err;
thread_local rxgpu_search_stats g_stats{};
thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
std::atomic<int> g_profile{0};
}  // namespace rxgpu

namespace {

thread_local std::vector<float> g_row_scratch;
thread_local std::vector<Hit> g_range_result;  // the last range search of this thread, retained for rxgpu_last_range_results()

int allocDevice(rxgpu_index* ix, uint64_t capacity, float** rows, uint64_t** labels, float** norms) {
	const size_t cap = capacity ? capacity : 1;
	RX_CUDA(cudaMalloc(reinterpret_cast<void**>(rows), cap * ix->pitch * sizeof(float)));
	RX_CUDA(cudaMalloc(reinterpret_cast<void**>(labels), cap * sizeof(uint64_t)));
	*norms = nullptr;
	if (ix->metric == RXGPU_COS) {
		RX_CUDA(cudaMalloc(reinterpret_cast<void**>(norms), cap * sizeof(float)));
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------- launches
template <int QT, int RW, int CG>
cudaError_t launchScanT(const rxgpu_index* ix, const ScanArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
	if (smem > size_t(kScanSmemBudget)) {
		return cudaErrorInvalidValue;
	}
	if (ix->metric == RXGPU_L2) {
		auto kfn = knn_scan_warp<QT, RW, CG, true>;
		if (const cudaError_t e = raiseSmemCeilingOnce(kfn, ix->device, kScanSmemBudget); e != cudaSuccess) {
			return e;
		}
		kfn<<<grid, kScanThreads, smem, st>>>(a);
	} else {
		auto kfn = knn_scan_warp<QT, RW, CG, false>;
		if (const cudaError_t e = raiseSmemCeilingOnce(kfn, ix->device, kScanSmemBudget); e != cudaSuccess) {
			return e;
		}
		kfn<<<grid, kScanThreads, smem, st>>>(a);
	}
	return cudaGetLastError();
}

template <int QT>
cudaError_t launchScanQ(const rxgpu_index* ix, const ScanArgs& a, uint32_t nch, unsigned* gridOut, cudaStream_t st, bool dryRun) {
	// chunk group CG divides nch; RW*CG float4 loads in flight per lane
	int cg, rw;
	if (nch % 6 == 0) {
		cg = 6, rw = 2;
	} else if (nch % 4 == 0) {
		cg = 4, rw = 2;
	} else if (nch % 3 == 0) {
		cg = 3, rw = 4;
	} else if (nch % 2 == 0) {
		cg = 2, rw = 4;
	} else {
		cg = 1, rw = 8;
	}
	const uint32_t nrows = a.row_end - a.row_begin;
	const uint32_t ngroups = (nrows + rw - 1) / rw;
	unsigned grid = std::min<unsigned>(unsigned(ix->sm_count) * 2u, std::max<unsigned>(1u, (ngroups + kScanWarps - 1) / kScanWarps));
	if (a.work != nullptr) {
		grid = a.nwork;  // one CTA per (query, inverted list)
	}
	*gridOut = grid;
	if (dryRun) {
		return cudaSuccess;
	}
	const size_t smem = scan_smem_bytes(QT, a.dim, a.k1);
	switch (cg) {
		case 6:
			return launchScanT<QT, 2, 6>(ix, a, grid, smem, st);
		case 4:
			return launchScanT<QT, 2, 4>(ix, a, grid, smem, st);
		case 3:
			return launchScanT<QT, 4, 3>(ix, a, grid, smem, st);
		case 2:
			return launchScanT<QT, 4, 2>(ix, a, grid, smem, st);
		default:
			return launchScanT<QT, 8, 1>(ix, a, grid, smem, st);
	}
}

cudaError_t launchScan(const rxgpu_index* ix, int qt, const ScanArgs& a, unsigned* gridOut, cudaStream_t st, bool dryRun = false) {
	const uint32_t nch = (a.dim + 127u) / 128u;
	switch (qt) {
		case 4:
			return launchScanQ<4>(ix, a, nch, gridOut, st, dryRun);
		case 2:
			return launchScanQ<2>(ix, a, nch, gridOut, st, dryRun);
		default:
			return launchScanQ<1>(ix, a, nch, gridOut, st, dryRun);
	}
}

int pickQueryTile(const rxgpu_index* ix, uint32_t nq) {
	uint32_t qt = ix->qt_override ? ix->qt_override : (nq >= 4 ? 4u : (nq >= 2 ? 2u : 1u));
	return qt >= 4 ? 4 : (qt >= 2 ? 2 : 1);
}

// Top-k1 rows per query under the total order (dist, internal index) -- or, in tie mode, the first k1 rows in internal
// order with dist <= bound.  Everything stays on the device; results land in d_out_* ([nq][k1]).
int scanTopKExact(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nq, uint32_t k1, int mode,
				  float bound, float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count) {
	// k1 > kMaxFusedK1: rounds of <= kMaxFusedK1 results; a round only admits keys above the previous round's last key, so the
	// rounds concatenate to the top-k1 under the same total order (one pass over the rows per round)
	const uint32_t kr = std::min<uint32_t>(k1, kMaxFusedK1);
	const uint32_t rounds = (k1 + kr - 1) / kr;
	int qt = mode == kModeTieRows ? 1 : pickQueryTile(ix, nq);
	while (qt > 1 && scan_smem_bytes(qt, ix->dim, kr) > 100 * 1024) {
		qt /= 2;
	}
	if (scan_smem_bytes(qt, ix->dim, kr) > 100 * 1024) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: dimension/k combination exceeds the fused top-k shared-memory budget");
	}
	ScanArgs a{};
	a.rows = ix->d_rows;
	a.norm_coefs = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
	a.pitch = ix->pitch;
	a.dim = ix->dim;
	a.row_begin = 0;
	a.row_end = uint32_t(ix->size);
	a.k1 = kr;
	a.mode = mode;
	a.bound = bound;
	unsigned grid = 0;
	RX_CUDA(launchScan(ix, qt, a, &grid, st, true));
	if (grid > 256u * kMergeOwn) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: scan grid exceeds the merge fan-in");
	}
	RX_CUDA(ws.d_lists.ensure(size_t(grid) * qt * kr));
	a.lists = ws.d_lists.p;
	if (rounds > 1) {
		RX_CUDA(ws.d_floor.ensure(size_t(qt)));
	}
	for (uint32_t q0 = 0; q0 < nq; q0 += qt) {
		a.queries = d_queries + size_t(q0) * ix->dim;
		a.nq = std::min<uint32_t>(qt, nq - q0);
		for (uint32_t round = 0; round < rounds; ++round) {
			a.k1 = std::min(kr, k1 - round * kr);
			a.floor_keys = round ? ws.d_floor.p : nullptr;
			cudaEvent_t e0 = nullptr, e1 = nullptr;
			if (g_profile.load(std::memory_order_relaxed)) {
				RX_CUDA(cudaEventCreate(&e0));
				RX_CUDA(cudaEventCreate(&e1));
				RX_CUDA(cudaEventRecord(e0, st));
			}
			RX_CUDA(launchScan(ix, qt, a, &grid, st));
			if (e0) {
				RX_CUDA(cudaEventRecord(e1, st));
				g_prof_events.emplace_back(e0, e1);
			}
			MergeArgs m{};
			m.lists = ws.d_lists.p;
			m.labels = ix->d_labels;
			m.out_dist = d_out_dist;
			m.out_idx = d_out_idx;
			m.out_label = d_out_label;
			m.out_count = d_out_count;
			m.floor_out = rounds > 1 ? ws.d_floor.p : nullptr;
			m.nlists = grid;
			m.qt = qt;
			m.k1 = a.k1;
			m.q_offset = q0;
			m.out_stride = k1;
			m.out_offset = round * kr;
			m.mode = mode;
			knn_merge_lists<<<a.nq, 256, 0, st>>>(m);
			RX_CUDA(cudaGetLastError());
			g_stats.launches += 2;
			g_stats.passes += 1;
		}
	}
	g_stats.query_tile = uint32_t(qt);
	const uint64_t perPass = uint64_t(ix->size) * ix->dim * 4 + (ix->metric == RXGPU_COS ? uint64_t(ix->size) * 4 : 0) +
							 uint64_t(qt) * ix->dim * 4 + uint64_t(qt) * kr * 12;
	g_stats.algorithmic_bytes += perPass * ((nq + qt - 1) / qt) * rounds;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------- tensor-core filter
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
								   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
								   CUtensorMapFloatOOBfill);
EncodeTiledFn encodeTiled() {  // libcuda is never linked: resolve the one driver entry point we need at run time
	static EncodeTiledFn fn = [] {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
			cudaGetLastError();
			return EncodeTiledFn(nullptr);
		}
		return reinterpret_cast<EncodeTiledFn>(p);
	}();
	return fn;
}
int makeBf16Map(CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint64_t pitchBytes, uint32_t boxRows) {
	const cuuint64_t gdim[2] = {cols, rows};
	const cuuint64_t gstr[1] = {pitchBytes};
	const cuuint32_t box[2] = {uint32_t(kTcChunkK), boxRows};
	const cuuint32_t estr[2] = {1, 1};
	const CUresult r = encodeTiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
									  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
	}
	return 0;
}

constexpr size_t kTcSmemLimit = 227 * 1024;
constexpr uint32_t kTcCandCap = 4096;

uint32_t tcQueryBlock(uint32_t nq, uint32_t kchunks) {
	uint32_t nqb = std::min<uint32_t>(256, (nq + 31u) & ~31u);
	while (nqb >= 32 && tc_smem_bytes(nqb, kchunks) > kTcSmemLimit) {
		nqb -= 32;
	}
	if (nqb < 32) {
		return 0;
	}
	const uint32_t blocks = (nq + nqb - 1) / nqb;
	return std::min(nqb, (((nq + blocks - 1) / blocks) + 31u) & ~31u);  // even out the blocks
}

bool tcEligible(const rxgpu_index* ix, uint32_t nq, uint32_t k1, int mode) {
	if (ix->tc_mode == 2 || mode != kModeTopK || k1 > kTcMaxK1 || !encodeTiled()) {
		return false;
	}
	if (tcQueryBlock(nq, (ix->dim + kTcChunkK - 1) / kTcChunkK) == 0) {
		return false;
	}
	// the error coefficient kTcErrCoef = 0.0042 certifies |q~.v~ - q.v| <= c ||q|| ||v|| only while
	// 2^-8 + 2^-18 + dim * 2^-23 <= 0.0042 (two bf16 roundings + fp32 accumulation): dim <= 2400; beyond 2048 dims the exact scan answers
	if (ix->dim > 2048) {
		return false;
	}
	return ix->tc_mode == 1 || (nq >= 64 && ix->size >= 100000);
}

// bf16 shadow + row norms, rebuilt when the rows changed since the last large-batch search (10M x 768: ~7 ms)
int ensureShadow(const rxgpu_index* ix, cudaStream_t st) {
	std::lock_guard<std::mutex> lck(ix->tc_mtx);
	const uint32_t pitchBf = (ix->dim + kTcChunkK - 1) / kTcChunkK * kTcChunkK;
	if (!ix->d_shadow) {
		const size_t cap = (size_t(ix->capacity ? ix->capacity : 1) + kTcTileRows - 1) / kTcTileRows * kTcTileRows;  // whole tiles
		RX_CUDA(cudaMalloc(&ix->d_shadow, cap * pitchBf * 2));
		RX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ix->d_vnorm), cap * sizeof(float)));
		RX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ix->d_vw), cap * sizeof(float2)));
		ix->pitch_bf = pitchBf;
		ix->shadow_version = ~0ull;
		ix->shadow_dirty_all = true;
		ix->shadow_dirty.clear();
	}
	if (ix->shadow_version != ix->version) {
		// only the rows the mutations since the last search rewrote (an upsert, a swap-remove, an appended run), unless the log gave up
		auto convert = [&](uint32_t b, uint32_t e) {
			const unsigned blocks = unsigned((uint64_t(e - b) * 32 + 255) / 256);
			tc_convert_rows<<<blocks, 256, 0, st>>>(ix->d_rows, ix->pitch, ix->dim, b, e, static_cast<__nv_bfloat16*>(ix->d_shadow),
													pitchBf / kTcChunkK, ix->d_vnorm);
		};
		if (ix->shadow_dirty_all) {
			if (ix->size) {
				convert(0, uint32_t(ix->size));
			}
		} else {
			for (const auto& r : ix->shadow_dirty) {
				const uint32_t e = uint32_t(std::min<uint64_t>(r.second, ix->size));
				if (r.first < e) {
					convert(r.first, e);
				}
			}
		}
		ix->shadow_dirty_all = false;
		ix->shadow_dirty.clear();
		const uint32_t padded = uint32_t((ix->size + kTcTileRows - 1) / kTcTileRows * kTcTileRows);
		tc_make_vw<<<(padded + 255) / 256, 256, 0, st>>>(ix->d_vnorm, uint32_t(ix->size), padded, ix->metric, ix->d_vw);
		RX_CUDA(cudaGetLastError());
		RX_CUDA(cudaStreamSynchronize(st));
		ix->shadow_version = ix->version;
		g_stats.launches += 1;
	}
	return 0;
}

int scanTopKExact(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nq, uint32_t k1, int mode,
				  float bound, float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count);

// Large batches: approximate bf16 tensor-core scores with a certified error bound select the candidates, the exact fp32 routine
// re-ranks them.  Output = the same top-k1 under (dist, internal index) as scanTopKExact, bit for bit.
int scanTopKTensorCore(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nq, uint32_t k1,
					   float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count) {
	if (int rc = ensureShadow(ix, st)) {
		return rc;
	}
	const uint32_t pitchBf = ix->pitch_bf, kchunks = pitchBf / kTcChunkK;
	const uint32_t nqb = tcQueryBlock(nq, kchunks);
	const uint32_t nblocks = (nq + nqb - 1) / nqb;
	const uint32_t nqPad = std::max<uint32_t>(nblocks * nqb, (nq + 511u) / 512u * 512u);  // both kernel generations index it
	RX_CUDA(ws.d_qbf.ensure(size_t(nqPad) * pitchBf));
	RX_CUDA(ws.d_qnorm.ensure(nqPad));
	RX_CUDA(ws.d_tau.ensure(nqPad));
	RX_CUDA(ws.d_ub_list.ensure(size_t(nqPad) * kTcMaxK1));
	RX_CUDA(ws.d_ub_lock.ensure(nqPad));
	RX_CUDA(ws.d_cand_count.ensure(nqPad));
	RX_CUDA(ws.d_cand_rows.ensure(size_t(nqPad) * kTcCandCap));
	RX_CUDA(ws.h_cand_count.ensure(nqPad));
	RX_CUDA(ws.d_lists.ensure(size_t(nq) * k1));
	tc_prepare_queries<<<(nqPad * 32 + 255) / 256, 256, 0, st>>>(d_queries, nq, nqPad, ix->dim, pitchBf,
																 reinterpret_cast<__nv_bfloat16*>(ws.d_qbf.p), ws.d_qnorm.p);
	RX_CUDA(raiseSmemCeilingOnce(tc_init_tau, ix->device, int(tc_init_smem_bytes(2048))));
	tc_init_tau<<<(nq + kTcInitQ - 1) / kTcInitQ, 256, tc_init_smem_bytes(ix->pitch), st>>>(
		ix->d_rows, ix->pitch, ix->dim, ix->metric == RXGPU_COS ? ix->d_norms : nullptr, uint32_t(std::min<uint64_t>(ix->size, kTcInitRows)),
		d_queries, nq, k1, ix->metric, ws.d_tau.p, ws.d_ub_list.p, ws.d_ub_lock.p);
	RX_CUDA(cudaMemsetAsync(ws.d_cand_count.p, 0, size_t(nqPad) * 4, st));
	RX_CUDA(cudaGetLastError());
	g_stats.launches += 2;
	const uint32_t ntiles = uint32_t((ix->size + kTcTileRows - 1) / kTcTileRows);
	bool launched = false;
	if ((ix->tc_variant == 0 || ix->tc_variant == 14) && kchunks <= kTqMaxKchunks) {
		// query block in tensor memory, deep TMA ring, row tiles multicast inside a cluster.  tc_variant 0: knn_tc_filter_q (every CTA
		// multiplies on its own, two accumulators of 64 rows; default); 14: knn_tc_filter_p (CTA pairs multiply as one, cta_group::2, every
		// SM stages half a 128-row tile, one accumulator).  Both run at the board's power limit and land on the same time.
		using TqKernel = void (*)(TqArgs);
		const bool pairs = ix->tc_variant == 14;
		const TqKernel kernels[4] = {pairs ? knn_tc_filter_p<2> : knn_tc_filter_q<1>, pairs ? knn_tc_filter_p<2> : knn_tc_filter_q<2>,
									 pairs ? knn_tc_filter_p<4> : knn_tc_filter_q<4>, pairs ? knn_tc_filter_p<8> : knn_tc_filter_q<8>};
		auto kernelOf = [&](int c) { return kernels[c == 8 ? 3 : (c == 4 ? 2 : (c == 2 ? 1 : 0))]; };
		const uint32_t tileRows = pairs ? kTpTileRows : kTqTileRows;
		const unsigned threads = pairs ? kTpThreads : kTqThreads;
		auto smemOf = [&](uint32_t st) { return pairs ? tp_smem_bytes(st) : tq_smem_bytes(st); };
		const uint32_t qblocks = (nq + kTqQueries - 1) / kTqQueries;
		uint32_t stages = 2;
		while (smemOf(stages + 1) <= kTcSmemLimit && stages < 64) {
			++stages;
		}
		const size_t smem = smemOf(stages);
		for (const TqKernel kfn : kernels) {
			RX_CUDA(raiseSmemCeilingOnce(kfn, ix->device, int(kTcSmemLimit)));
		}
		// a cluster of C CTAs reads every row tile from HBM once for C x 128 queries (TMA multicast)
		const uint32_t clusterMax = ix->tc_cluster_max ? ix->tc_cluster_max : 4u;  // mode 9: up to 8 (one launch serves 1024 queries)
		int cluster = qblocks >= 5 ? 8 : (qblocks >= 3 ? 4 : (qblocks == 2 ? 2 : 1));
		cluster = std::min<int>(cluster, int(clusterMax));
		if (pairs) {
			cluster = std::max(cluster, 2);  // whole CTA pairs (a second CTA without queries multiplies zeros)
		}
		const uint32_t qtiles = uint32_t((ix->size + tileRows - 1) / tileRows);
		unsigned grid = 0;
		int residentClusters = 0;
		for (;;) {  // how many clusters of this size can be resident at once (GPC boundaries strand SMs for size 4)
			cudaLaunchConfig_t cfg{};
			cfg.gridDim = dim3(unsigned(ix->sm_count) / cluster * cluster);
			cfg.blockDim = dim3(threads);
			cfg.dynamicSmemBytes = smem;
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeClusterDimension;
			attr[0].val.clusterDim.x = unsigned(cluster);
			attr[0].val.clusterDim.y = 1;
			attr[0].val.clusterDim.z = 1;
			cfg.attrs = attr;
			cfg.numAttrs = 1;
			int maxClusters = 0;
			const cudaError_t e = cudaOccupancyMaxActiveClusters(&maxClusters, kernelOf(cluster), &cfg);
			if (e == cudaSuccess && maxClusters > 0) {
				grid = unsigned(std::min<uint64_t>(uint64_t(maxClusters), std::max<uint32_t>(qtiles, 1))) * cluster;
				residentClusters = maxClusters;
				break;
			}
			cudaGetLastError();
			if (cluster == (pairs ? 2 : 1)) {
				return fail(RXGPU_ERR_SYSTEM, "rxgpu: tensor-core filter kernel cannot be made resident");
			}
			cluster /= 2;
		}
		// Tail grid: clusters of 4 fit 33 times on a B200 (132 of 148 SMs, GPC boundaries).  The kernel is power-bound, so more
		// units at a lower clock is the lever left: 2-CTA clusters of the same kernel take the stranded SMs and scan the last
		// Ct / (2 Cm + Ct) of the row tiles -- twice per main launch, once for each half of its 4 query blocks -- on a second stream.
		// Candidate lists, thresholds and bound lists are per query and global, so both grids feed the same re-rank.
		uint32_t tailClusters = 0, tilesMain = qtiles;
		if (ix->tc_tail && cluster == 4 && int(grid) == residentClusters * cluster && qtiles >= 256) {
			const uint32_t spare = uint32_t(ix->sm_count) - grid;
			tailClusters = spare / 2;
			if (tailClusters) {
				const uint32_t cm = grid / cluster;
				// the tail's share by SM count would be Ct / (2 Cm + Ct) = 10.8 %; measured best at 10 % (80 / 100 / 120 / 140 permille:
			// 71.4 / 72.5 / 70.8 / 68.1 k queries/s on one box): a 2-CTA cluster reads its rows from HBM once per 256 queries, not 512
			uint32_t tilesTail = uint32_t(uint64_t(qtiles) * tailClusters * 15 / ((2ull * cm + tailClusters) * 16));
			static const char* tp = std::getenv("RXGPU_TC_TAIL_PERMILLE");  // tuning aid: the tail's share of the row tiles
			if (tp) {
				tilesTail = uint32_t(uint64_t(qtiles) * uint32_t(std::atoi(tp)) / 1000);
			}
				tilesMain = qtiles - tilesTail;
				if (!ws.tail_stream) {
					RX_CUDA(cudaStreamCreateWithFlags(&ws.tail_stream, cudaStreamNonBlocking));
					RX_CUDA(cudaEventCreateWithFlags(&ws.tail_fork, cudaEventDisableTiming));
					RX_CUDA(cudaEventCreateWithFlags(&ws.tail_join, cudaEventDisableTiming));
				}
				RX_CUDA(cudaEventRecord(ws.tail_fork, st));
				RX_CUDA(cudaStreamWaitEvent(ws.tail_stream, ws.tail_fork, 0));
			}
		}
		const uint32_t rowsMain = uint32_t(std::min<uint64_t>(ix->size, uint64_t(tilesMain) * tileRows));
		for (uint32_t b = 0; b < qblocks; b += cluster) {
			TqArgs a{};
			a.shadow = static_cast<const unsigned char*>(ix->d_shadow);
			a.vnorm = ix->d_vnorm;
			a.vw = ix->d_vw;
			a.vinv = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
			a.qnorm = ws.d_qnorm.p;
			a.qbf = ws.d_qbf.p;
			a.tau = ws.d_tau.p;
			a.ub_list = ws.d_ub_list.p;
			a.ub_lock = ws.d_ub_lock.p;
			a.cand_rows = ws.d_cand_rows.p;
			a.cand_count = ws.d_cand_count.p;
			a.cand_cap = kTcCandCap;
			a.init_rows = uint32_t(std::min<uint64_t>(ix->size, kTcInitRows));
			a.n = rowsMain;
			a.kchunks = kchunks;
			a.pitch_bf = pitchBf;
			a.nq_total = nq;
			a.q0 = b * kTqQueries;
			a.k1 = k1;
			a.stages = stages;
			a.metric = ix->metric;
			{
				static const char* pf = std::getenv("RXGPU_TC_PREFETCH");  // tuning aid: L2 prefetch distance in tiles
				a.prefetch = pf ? uint32_t(std::atoi(pf)) : 0u;
				static const char* si = std::getenv("RXGPU_TC_SINGLE_ISSUER");  // tuning aid for knn_tc_filter_q
				a.single_issuer = si ? uint32_t(std::atoi(si)) : 0u;  // measured: no effect (the ring is not what limits the kernel), off by default
			}
			static DevBuf<unsigned long long> traceBuf;  // profiling aid: RXGPU_TC_TRACE=<file> dumps per-tile timestamps of CTA 0
			const char* tracePath = std::getenv("RXGPU_TC_TRACE");
			if (tracePath && b == 0) {
				RX_CUDA(traceBuf.ensure(256 * 16));
				RX_CUDA(cudaMemsetAsync(traceBuf.p, 0, 256 * 16 * 8, st));
				a.trace = traceBuf.p;
				const char* first = std::getenv("RXGPU_TC_TRACE_FIRST");
				a.trace_first = first ? uint32_t(std::atoi(first)) : 0u;
			}
			cudaEvent_t e0 = nullptr, e1 = nullptr;
			if (g_profile.load(std::memory_order_relaxed)) {
				RX_CUDA(cudaEventCreate(&e0));
				RX_CUDA(cudaEventCreate(&e1));
				RX_CUDA(cudaEventRecord(e0, st));
			}
			cudaLaunchConfig_t cfg{};
			cfg.gridDim = dim3(grid);
			cfg.blockDim = dim3(threads);
			cfg.dynamicSmemBytes = smem;
			cfg.stream = st;
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeClusterDimension;
			attr[0].val.clusterDim.x = unsigned(cluster);
			attr[0].val.clusterDim.y = 1;
			attr[0].val.clusterDim.z = 1;
			cfg.attrs = attr;
			cfg.numAttrs = 1;
			RX_CUDA(cudaLaunchKernelEx(&cfg, kernelOf(cluster), a));
			RX_CUDA(cudaGetLastError());
			if (e0) {
				RX_CUDA(cudaEventRecord(e1, st));
				g_prof_events.emplace_back(e0, e1);
			}
			g_stats.launches += 1;
			g_stats.passes += 1;
			for (uint32_t pb = b; tailClusters && pb < std::min<uint32_t>(b + uint32_t(cluster), qblocks); pb += 2) {
				TqArgs t = a;
				t.trace = nullptr;
				t.row_base = rowsMain;
				t.shadow = a.shadow + size_t(rowsMain) * pitchBf * 2;
				t.vw = a.vw + rowsMain;
				t.n = uint32_t(ix->size) - rowsMain;
				t.q0 = pb * kTqQueries;
				cudaLaunchConfig_t tcfg{};
				tcfg.gridDim = dim3(tailClusters * 2);
				tcfg.blockDim = dim3(threads);
				tcfg.dynamicSmemBytes = smem;
				tcfg.stream = ws.tail_stream;
				cudaLaunchAttribute tattr[1];
				tattr[0].id = cudaLaunchAttributeClusterDimension;
				tattr[0].val.clusterDim.x = 2;
				tattr[0].val.clusterDim.y = 1;
				tattr[0].val.clusterDim.z = 1;
				tcfg.attrs = tattr;
				tcfg.numAttrs = 1;
				RX_CUDA(cudaLaunchKernelEx(&tcfg, kernelOf(2), t));
				RX_CUDA(cudaGetLastError());
				g_stats.launches += 1;
			}
			if (a.trace) {
				std::vector<unsigned long long> h(256 * 16);
				RX_CUDA(cudaStreamSynchronize(st));
				RX_CUDA(cudaMemcpy(h.data(), a.trace, h.size() * 8, cudaMemcpyDeviceToHost));
				if (FILE* f = std::fopen(std::getenv("RXGPU_TC_TRACE"), "w")) {
					for (int i = 0; i < 256; ++i) {
						for (int j = 0; j < 16; ++j) {
							std::fprintf(f, "%llu%c", h[i * 16 + j], j == 15 ? '\n' : ' ');
						}
					}
					std::fclose(f);
				}
			}
		}
		if (tailClusters) {
			RX_CUDA(cudaEventRecord(ws.tail_join, ws.tail_stream));
			RX_CUDA(cudaStreamWaitEvent(st, ws.tail_join, 0));
		}
		g_stats.tc_cluster = uint32_t(cluster);
		g_stats.tc_kernel = pairs ? 5 : 2;
		g_stats.query_tile = uint32_t(kTqQueries * cluster);
		g_stats.algorithmic_bytes += uint64_t((qblocks + cluster - 1) / cluster) * (uint64_t(ix->size) * pitchBf * 2 + uint64_t(ix->size) * 4) +
									 uint64_t(nq) * pitchBf * 2;
		launched = true;
	}
	if (!launched) {
	// Two CTAs per cluster share every row tile (TMA multicast) and own different query blocks: one pass serves 2*nqb queries.
	const int cluster = (ix->tc_variant == 3 || nblocks < 2 || ix->sm_count < 2 || ntiles < 2) ? 1 : 2;
	CUtensorMap mapQ;
	if (int rc = makeBf16Map(&mapQ, ws.d_qbf.p, pitchBf, nqPad, uint64_t(pitchBf) * 2, nqb)) {
		return rc;
	}
	const size_t smem = tc_smem_bytes(nqb, kchunks);
	RX_CUDA(raiseSmemCeilingOnce(knn_tc_filter<1>, ix->device, int(kTcSmemLimit)));
	RX_CUDA(raiseSmemCeilingOnce(knn_tc_filter<2>, ix->device, int(kTcSmemLimit)));
	unsigned grid = std::min<unsigned>(unsigned(ix->sm_count), ntiles * cluster);
	grid -= grid % cluster;
	for (uint32_t b = 0; b < nblocks; b += cluster) {
		TcArgs a{};
		a.shadow = static_cast<const unsigned char*>(ix->d_shadow);
		a.vnorm = ix->d_vnorm;
		a.vinv = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
		a.qnorm = ws.d_qnorm.p;
		a.tau = ws.d_tau.p;
		a.ub_list = ws.d_ub_list.p;
		a.ub_lock = ws.d_ub_lock.p;
		a.init_rows = uint32_t(std::min<uint64_t>(ix->size, kTcInitRows));
		a.cand_rows = ws.d_cand_rows.p;
		a.cand_count = ws.d_cand_count.p;
		a.cand_cap = kTcCandCap;
		a.n = uint32_t(ix->size);
		a.kchunks = kchunks;
		a.nq_block = nqb;
		a.q0 = b * nqb;
		a.nq_total = nq;
		a.k1 = k1;
		a.metric = ix->metric;
		cudaEvent_t e0 = nullptr, e1 = nullptr;
		if (g_profile.load(std::memory_order_relaxed)) {
			RX_CUDA(cudaEventCreate(&e0));
			RX_CUDA(cudaEventCreate(&e1));
			RX_CUDA(cudaEventRecord(e0, st));
		}
		if (cluster == 2) {
			cudaLaunchConfig_t cfg{};
			cfg.gridDim = dim3(grid);
			cfg.blockDim = dim3(kTcThreads);
			cfg.dynamicSmemBytes = smem;
			cfg.stream = st;
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeClusterDimension;
			attr[0].val.clusterDim.x = 2;
			attr[0].val.clusterDim.y = 1;
			attr[0].val.clusterDim.z = 1;
			cfg.attrs = attr;
			cfg.numAttrs = 1;
			RX_CUDA(cudaLaunchKernelEx(&cfg, knn_tc_filter<2>, mapQ, a));
		} else {
			knn_tc_filter<1><<<grid, kTcThreads, smem, st>>>(mapQ, a);
		}
		RX_CUDA(cudaGetLastError());
		if (e0) {
			RX_CUDA(cudaEventRecord(e1, st));
			g_prof_events.emplace_back(e0, e1);
		}
		g_stats.launches += 1;
		g_stats.passes += 1;
	}
	g_stats.tc_cluster = uint32_t(cluster);
	g_stats.tc_kernel = 1;
	g_stats.query_tile = nqb;
	g_stats.algorithmic_bytes += uint64_t((nblocks + cluster - 1) / cluster) * (uint64_t(ix->size) * pitchBf * 2 + uint64_t(ix->size) * 4 +
																			   uint64_t(nqb) * pitchBf * 2);
	}
	// exact re-rank of the candidates with the arithmetic of knn_scan_warp, then decode + labels
	const size_t rsmem = size_t((ix->dim + 127) / 128) * 512 + size_t(kScanWarps) * (k1 + kCandBuf) * 8;
	const float* norms = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
	if (ix->metric == RXGPU_L2) {
		RX_CUDA(raiseSmemCeilingOnce(knn_rerank<true>, ix->device, kScanSmemBudget));
		knn_rerank<true><<<nq, kScanThreads, rsmem, st>>>(ix->d_rows, ix->pitch, ix->dim, norms, d_queries, ws.d_cand_rows.p,
														   ws.d_cand_count.p, kTcCandCap, k1, ws.d_lists.p);
	} else {
		RX_CUDA(raiseSmemCeilingOnce(knn_rerank<false>, ix->device, kScanSmemBudget));
		knn_rerank<false><<<nq, kScanThreads, rsmem, st>>>(ix->d_rows, ix->pitch, ix->dim, norms, d_queries, ws.d_cand_rows.p,
															ws.d_cand_count.p, kTcCandCap, k1, ws.d_lists.p);
	}
	MergeArgs m{};
	m.lists = ws.d_lists.p;
	m.labels = ix->d_labels;
	m.out_dist = d_out_dist;
	m.out_idx = d_out_idx;
	m.out_label = d_out_label;
	m.out_count = d_out_count;
	m.nlists = 1;
	m.qt = nq;
	m.k1 = k1;
	m.q_offset = 0;
	m.out_stride = k1;
	m.out_offset = 0;
	m.mode = kModeTopK;
	knn_merge_lists<<<nq, 256, 0, st>>>(m);
	RX_CUDA(cudaGetLastError());
	g_stats.launches += 2;
	// a query whose candidate list overflowed (pathological data, e.g. masses of near-duplicates) is answered by the exact scan
	RX_CUDA(cudaMemcpyAsync(ws.h_cand_count.p, ws.d_cand_count.p, size_t(nq) * 4, cudaMemcpyDeviceToHost, st));
	RX_CUDA(cudaStreamSynchronize(st));
	uint64_t cands = 0;
	for (uint32_t q = 0; q < nq; ++q) {
		cands += std::min<unsigned>(ws.h_cand_count.p[q], kTcCandCap);
		if (ws.h_cand_count.p[q] > kTcCandCap) {
			g_stats.tc_fallbacks += 1;
			if (int rc = scanTopKExact(ix, ws, st, d_queries + size_t(q) * ix->dim, 1, k1, kModeTopK, 0.f, d_out_dist + size_t(q) * k1,
									   d_out_idx + size_t(q) * k1, d_out_label ? d_out_label + size_t(q) * k1 : nullptr, d_out_count + q)) {
				return rc;
			}
		}
	}
	ws.tc_lists_valid = true;
	ws.tc_lists_nq = nq;
	ws.tc_lists_version = ix->version;
	g_stats.tc_used = 1;
	g_stats.tc_candidates = cands;
	g_stats.algorithmic_bytes += cands * (uint64_t(ix->dim) * 4 + 4);
	return 0;
}

}  // namespace

namespace rxgpu {
int scanTopK(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nq, uint32_t k1, int mode, float bound,
			 float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count) {
	if (tcEligible(ix, nq, k1, mode)) {
		return scanTopKTensorCore(ix, ws, st, d_queries, nq, k1, d_out_dist, d_out_idx, d_out_label, d_out_count);
	}
	if (mode == kModeTopK) {
		ws.tc_lists_valid = false;
	}
	return scanTopKExact(ix, ws, st, d_queries, nq, k1, mode, bound, d_out_dist, d_out_idx, d_out_label, d_out_count);
}

int setRowAt(rxgpu_index* ix, uint32_t idx, uint64_t label, const float* vec) {
	if (idx > ix->size || (idx == ix->size && ix->size >= ix->capacity)) {
		return fail(RXGPU_ERR_LOGIC, "The number of elements exceeds the specified limit\n");
	}
	const uint32_t other = ix->dict.find(label);
	if (other != LabelMap::kNotFound && other != idx) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: label already belongs to another row");
	}
	RX_CUDA(ix->st_rows.ensure(ix->pitch));
	RX_CUDA(cudaMemsetAsync(ix->st_rows.p, 0, size_t(ix->pitch) * 4, ix->stream));
	RX_CUDA(cudaMemcpyAsync(ix->st_rows.p, vec, size_t(ix->dim) * 4, cudaMemcpyHostToDevice, ix->stream));
	RX_CUDA(cudaMemcpyAsync(ix->d_rows + size_t(idx) * ix->pitch, ix->st_rows.p, size_t(ix->pitch) * 4, cudaMemcpyDeviceToDevice, ix->stream));
	RX_CUDA(cudaMemcpyAsync(ix->d_labels + idx, &label, 8, cudaMemcpyHostToDevice, ix->stream));
	if (ix->metric == RXGPU_COS) {
		norm_coef_kernel<<<1, 32, 0, ix->stream>>>(ix->d_rows, ix->pitch, ix->dim, idx, idx + 1, ix->d_norms);
		RX_CUDA(cudaGetLastError());
	}
	RX_CUDA(cudaStreamSynchronize(ix->stream));
	if (idx < ix->size) {
		if (ix->h_labels[idx] != label) {
			ix->dict.erase(ix->h_labels[idx]);
		}
		ix->h_labels[idx] = label;
	} else {
		ix->h_labels.push_back(label);
		ix->size += 1;
	}
	ix->dict.put(label, idx);
	if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
		std::memcpy(ix->h_rows.data() + size_t(idx) * ix->dim, vec, ix->dim * sizeof(float));
	}
	ix->touchRows(idx, uint64_t(idx) + 1);
	ix->version++;
	return 0;
}

int tieRowsAfterScan(const rxgpu_index* ix, Workspace& ws, cudaStream_t st, const float* d_queries, uint32_t nsel, const uint32_t* sel,
					 const float* dstar, uint32_t k, float* d_out_dist, uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count) {
	if (nsel == 0) {
		return 0;
	}
	bool fromLists = ws.tc_lists_valid && ws.tc_lists_version == ix->version && k <= kTcMaxK1;
	for (uint32_t i = 0; fromLists && i < nsel; ++i) {  // a query whose list overflowed was answered by the exact scan: no list
		fromLists = sel[i] < ws.tc_lists_nq && ws.h_cand_count.p[sel[i]] <= kTcCandCap;
	}
	if (!fromLists) {
		for (uint32_t i = 0; i < nsel; ++i) {
			if (int rc = scanTopKExact(ix, ws, st, d_queries + size_t(sel[i]) * ix->dim, 1, k, kModeTieRows, dstar[i], d_out_dist + size_t(i) * k,
									   d_out_idx + size_t(i) * k, d_out_label ? d_out_label + size_t(i) * k : nullptr, d_out_count + i)) {
				return rc;
			}
		}
		g_stats.tie_replays += nsel;
		return 0;
	}
	RX_CUDA(ws.d_sel.ensure(nsel));
	RX_CUDA(ws.d_selbound.ensure(nsel));
	RX_CUDA(ws.d_lists.ensure(size_t(nsel) * k));
	RX_CUDA(cudaMemcpyAsync(ws.d_sel.p, sel, size_t(nsel) * 4, cudaMemcpyHostToDevice, st));
	RX_CUDA(cudaMemcpyAsync(ws.d_selbound.p, dstar, size_t(nsel) * 4, cudaMemcpyHostToDevice, st));
	const size_t rsmem = size_t((ix->dim + 127) / 128) * 512 + size_t(kScanWarps) * (k + kCandBuf) * 8;
	const float* norms = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
	if (ix->metric == RXGPU_L2) {
		RX_CUDA(raiseSmemCeilingOnce(knn_rerank<true>, ix->device, kScanSmemBudget));
		knn_rerank<true><<<nsel, kScanThreads, rsmem, st>>>(ix->d_rows, ix->pitch, ix->dim, norms, d_queries, ws.d_cand_rows.p, ws.d_cand_count.p,
															 kTcCandCap, k, ws.d_lists.p, ws.d_sel.p, ws.d_selbound.p);
	} else {
		RX_CUDA(raiseSmemCeilingOnce(knn_rerank<false>, ix->device, kScanSmemBudget));
		knn_rerank<false><<<nsel, kScanThreads, rsmem, st>>>(ix->d_rows, ix->pitch, ix->dim, norms, d_queries, ws.d_cand_rows.p, ws.d_cand_count.p,
															  kTcCandCap, k, ws.d_lists.p, ws.d_sel.p, ws.d_selbound.p);
	}
	MergeArgs m{};
	m.lists = ws.d_lists.p;
	m.labels = ix->d_labels;
	m.out_dist = d_out_dist;
	m.out_idx = d_out_idx;
	m.out_label = d_out_label;
	m.out_count = d_out_count;
	m.nlists = 1;
	m.qt = nsel;
	m.k1 = k;
	m.q_offset = 0;
	m.out_stride = k;
	m.out_offset = 0;
	m.mode = kModeTieRows;
	knn_merge_lists<<<nsel, 256, 0, st>>>(m);
	RX_CUDA(cudaGetLastError());
	g_stats.launches += 2;
	g_stats.tie_replays += nsel;
	g_stats.tie_from_lists += nsel;
	return 0;
}
}  // namespace rxgpu

namespace {

int normsForRange(rxgpu_index* ix, uint64_t begin, uint64_t end) {
	if (ix->metric != RXGPU_COS || begin >= end) {
		return 0;
	}
	const uint64_t rows = end - begin;
	const unsigned blocks = unsigned((rows * 32 + 255) / 256);
	norm_coef_kernel<<<blocks, 256, 0, ix->stream>>>(ix->d_rows, ix->pitch, ix->dim, uint32_t(begin), uint32_t(end), ix->d_norms);
	RX_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace

extern "C" {

const char* rxgpu_last_error(void) { return g_err.c_str(); }
int rxgpu_abi_version(void) { return RXGPU_ABI_VERSION; }
int rxgpu_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int rxgpu_index_create(rxgpu_index** out, rxgpu_metric metric, uint32_t dim, uint64_t capacity, int device, uint32_t flags) {
	if (!out) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null output handle");
	}
	*out = nullptr;
	if (dim == 0 || dim > 65536) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: vector dimension must be in [1, 65536]");
	}
	if (int(metric) < 0 || int(metric) > 2) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: unknown vector metric");
	}
	if (capacity >= (1ull << 31)) {  // the reference scans with an `int` index (bruteforce.cc:116)
		return fail(RXGPU_ERR_PARAMS, "rxgpu: capacity must be below 2^31 rows per shard");
	}
	if (rxgpu_device_count() <= device || device < 0) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: no usable CUDA device (this library has no CPU fallback)");
	}
	RX_CUDA(cudaSetDevice(device));
	std::unique_ptr<rxgpu_index> ix;
	try {
		ix = std::make_unique<rxgpu_index>();
		ix->metric = int(metric);
		ix->dim = dim;
		ix->pitch = (dim + 3u) & ~3u;
		ix->capacity = capacity;
		ix->device = device;
		ix->flags = flags;
		ix->h_labels.reserve(capacity);
		ix->dict.reserve(capacity);
		if (flags & RXGPU_FLAG_HOST_MIRROR) {
			ix->h_rows.resize(size_t(capacity) * dim);
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "Not enough memory: BruteforceSearch failed to allocate data");
	}
	cudaDeviceProp prop{};
	RX_CUDA(cudaGetDeviceProperties(&prop, device));
	ix->sm_count = prop.multiProcessorCount;
	RX_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
	if (int rc = allocDevice(ix.get(), capacity, &ix->d_rows, &ix->d_labels, &ix->d_norms)) {
		return rc;
	}
	*out = ix.release();
	return 0;
}

int rxgpu_index_clone(rxgpu_index** out, const rxgpu_index* src, uint64_t new_capacity) {
	if (!out || !src) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null handle");
	}
	const uint64_t cap = std::max(src->capacity, new_capacity);  // bruteforce.cc:22
	rxgpu_index* ix = nullptr;
	if (int rc = rxgpu_index_create(&ix, rxgpu_metric(src->metric), src->dim, cap, src->device, src->flags)) {
		return rc;
	}
	std::unique_ptr<rxgpu_index> guard(ix);
	try {
		ix->h_labels = src->h_labels;
		ix->dict = src->dict;
		if (src->flags & RXGPU_FLAG_HOST_MIRROR) {
			std::memcpy(ix->h_rows.data(), src->h_rows.data(), size_t(src->size) * src->dim * sizeof(float));
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "Not enough memory: BruteforceSearch failed to allocate data");
	}
	ix->size = src->size;
	ix->qt_override = src->qt_override;
	ix->tc_mode = src->tc_mode;
	ix->tc_variant = src->tc_variant;
	ix->tc_cluster_max = src->tc_cluster_max;
	RX_CUDA(cudaMemcpyAsync(ix->d_rows, src->d_rows, size_t(src->size) * src->pitch * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
	RX_CUDA(cudaMemcpyAsync(ix->d_labels, src->d_labels, size_t(src->size) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ix->stream));
	if (src->d_norms) {
		RX_CUDA(cudaMemcpyAsync(ix->d_norms, src->d_norms, size_t(src->size) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
	}
	RX_CUDA(cudaStreamSynchronize(ix->stream));
	*out = guard.release();
	return 0;
}

void rxgpu_index_destroy(rxgpu_index* ix) { delete ix; }

int rxgpu_index_resize(rxgpu_index* ix, uint64_t new_capacity) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (new_capacity < ix->size) {
		return fail(RXGPU_ERR_LOGIC, "Cannot resize, max element is less than the current number of elements");
	}
	if (new_capacity >= (1ull << 31)) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: capacity must be below 2^31 rows per shard");
	}
	if (new_capacity == ix->capacity) {
		return 0;
	}
	float *rows = nullptr, *norms = nullptr;
	uint64_t* labels = nullptr;
	if (int rc = allocDevice(ix, new_capacity, &rows, &labels, &norms)) {
		cudaFree(rows);
		cudaFree(labels);
		cudaFree(norms);
		return fail(RXGPU_ERR_SYSTEM, "Not enough memory: resizeIndex failed to allocate data");
	}
	RX_CUDA(cudaMemcpyAsync(rows, ix->d_rows, size_t(ix->size) * ix->pitch * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
	RX_CUDA(cudaMemcpyAsync(labels, ix->d_labels, size_t(ix->size) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ix->stream));
	if (norms) {
		RX_CUDA(cudaMemcpyAsync(norms, ix->d_norms, size_t(ix->size) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
	}
	RX_CUDA(cudaStreamSynchronize(ix->stream));
	cudaFree(ix->d_rows);
	cudaFree(ix->d_labels);
	cudaFree(ix->d_norms);
	ix->d_rows = rows;
	ix->d_labels = labels;
	ix->d_norms = norms;
	ix->capacity = new_capacity;
	if (ix->d_shadow) {  // rebuilt lazily at the new capacity
		cudaFree(ix->d_shadow);
		cudaFree(ix->d_vnorm);
		cudaFree(ix->d_vw);
		ix->d_vw = nullptr;
		ix->d_shadow = nullptr;
		ix->d_vnorm = nullptr;
	}
	try {
		if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
			ix->h_rows.resize(size_t(new_capacity) * ix->dim);
		}
		ix->h_labels.reserve(new_capacity);
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "Not enough memory: resizeIndex failed to allocate data");
	}
	return 0;
}

int rxgpu_index_upsert_batch(rxgpu_index* ix, uint64_t n, const uint64_t* labels, const float* vecs) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (n == 0) {
		return 0;
	}
	if (!labels || !vecs) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null labels / vectors");
	}
	// Resolve destinations sequentially, exactly like n AddPointNoLock calls (bruteforce.cc:44-64) -- into temporaries: the
	// dictionary, h_labels, the host mirror and size are committed only after every device operation of the batch succeeded, so a
	// failed call (staging allocation, copy or launch error) leaves the index exactly as it was.
	try {
		std::vector<uint32_t> dst(n);
		std::vector<uint64_t> fresh;  // new labels in arrival order; fresh[i] goes to row size + i
		std::unordered_map<uint64_t, uint32_t> pending;  // new label -> row, for repeats of a new label inside the batch
		uint64_t newSize = ix->size;
		bool pureAppend = true;
		uint64_t accepted = n;
		for (uint64_t i = 0; i < n; ++i) {
			uint32_t idx = ix->dict.find(labels[i]);
			if (idx == LabelMap::kNotFound) {
				const auto it = pending.find(labels[i]);
				if (it != pending.end()) {
					idx = it->second;
					pureAppend = false;
				} else {
					if (newSize >= ix->capacity) {
						accepted = i;  // rows before i are applied, like the reference's sequential calls
						break;
					}
					idx = uint32_t(newSize++);
					pending.emplace(labels[i], idx);
					fresh.push_back(labels[i]);
				}
			} else {
				pureAppend = false;
			}
			dst[i] = idx;
		}
		const uint64_t m = accepted;
		if (m) {
			if (pureAppend && ix->pitch == ix->dim) {
				RX_CUDA(cudaMemcpyAsync(ix->d_rows + size_t(ix->size) * ix->pitch, vecs, size_t(m) * ix->dim * sizeof(float),
										cudaMemcpyHostToDevice, ix->stream));
				RX_CUDA(cudaMemcpyAsync(ix->d_labels + ix->size, labels, size_t(m) * sizeof(uint64_t), cudaMemcpyHostToDevice, ix->stream));
				if (int rc = normsForRange(ix, ix->size, ix->size + m)) {
					return rc;
				}
			} else {
				// stage + scatter in bounded slices; the staging buffers live in the index (no cudaMalloc / cudaFree per call)
				const uint64_t slice = std::max<uint64_t>(1, (64ull << 20) / (ix->dim * sizeof(float)));
				RX_CUDA(ix->st_rows.ensure(size_t(std::min(slice, m)) * ix->dim));
				RX_CUDA(ix->st_dst.ensure(size_t(std::min(slice, m))));
				RX_CUDA(ix->st_labels.ensure(size_t(std::min(slice, m))));
				for (uint64_t off = 0; off < m; off += slice) {
					const uint64_t cnt = std::min(slice, m - off);
					RX_CUDA(cudaMemcpyAsync(ix->st_rows.p, vecs + off * ix->dim, size_t(cnt) * ix->dim * sizeof(float), cudaMemcpyHostToDevice,
											ix->stream));
					RX_CUDA(cudaMemcpyAsync(ix->st_dst.p, dst.data() + off, size_t(cnt) * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
					RX_CUDA(cudaMemcpyAsync(ix->st_labels.p, labels + off, size_t(cnt) * sizeof(uint64_t), cudaMemcpyHostToDevice, ix->stream));
					// duplicates of one label inside a slice must apply in order: one launch per row when present
					bool dup = false;
					if (cnt > 1) {
						std::vector<uint32_t> sorted(dst.begin() + off, dst.begin() + off + cnt);
						std::sort(sorted.begin(), sorted.end());
						dup = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
					}
					if (!dup) {
						scatter_rows_kernel<<<unsigned(cnt), 128, 0, ix->stream>>>(ix->st_rows.p, ix->st_dst.p, ix->st_labels.p, uint32_t(cnt), ix->dim,
																				   ix->pitch, ix->d_rows, ix->d_labels);
						RX_CUDA(cudaGetLastError());
					} else {
						for (uint64_t i = 0; i < cnt; ++i) {
							scatter_rows_kernel<<<1, 128, 0, ix->stream>>>(ix->st_rows.p + i * ix->dim, ix->st_dst.p + i, ix->st_labels.p + i, 1, ix->dim,
																		   ix->pitch, ix->d_rows, ix->d_labels);
						}
						RX_CUDA(cudaGetLastError());
					}
					if (ix->metric == RXGPU_COS) {
						for (uint64_t i = 0; i < cnt;) {  // norms for maximal runs of consecutive destinations
							uint64_t j = i + 1;
							while (j < cnt && dst[off + j] == dst[off + j - 1] + 1) {
								++j;
							}
							if (int rc = normsForRange(ix, dst[off + i], uint64_t(dst[off + j - 1]) + 1)) {
								return rc;
							}
							i = j;
						}
					}
					if (off + slice < m) {
						RX_CUDA(cudaStreamSynchronize(ix->stream));  // the staging buffers are reused by the next slice
					}
				}
			}
			RX_CUDA(cudaStreamSynchronize(ix->stream));
			// ---- commit (host state only; nothing below can fail except by std::bad_alloc, which reserve() at create/resize precludes)
			for (size_t i = 0; i < fresh.size(); ++i) {
				ix->dict.put(fresh[i], uint32_t(ix->size + i));
				ix->h_labels.push_back(fresh[i]);
			}
			if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
				for (uint64_t i = 0; i < m; ++i) {
					std::memcpy(ix->h_rows.data() + size_t(dst[i]) * ix->dim, vecs + i * ix->dim, ix->dim * sizeof(float));
				}
			}
			for (uint64_t i = 0; i < m;) {  // maximal runs of consecutive destinations (an appended batch is one range)
				uint64_t j = i + 1;
				while (j < m && dst[j] == dst[j - 1] + 1) {
					++j;
				}
				ix->touchRows(dst[i], uint64_t(dst[j - 1]) + 1);
				i = j;
			}
			ix->size = newSize;
			ix->version++;
		}
		if (accepted < n) {
			return fail(RXGPU_ERR_LOGIC, "The number of elements exceeds the specified limit\n");
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_index_upsert(rxgpu_index* ix, uint64_t label, const float* vec) { return rxgpu_index_upsert_batch(ix, 1, &label, vec); }

int rxgpu_index_remove(rxgpu_index* ix, uint64_t label) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	const uint32_t cur = ix->dict.find(label);
	if (cur == LabelMap::kNotFound) {
		return 0;  // bruteforce.cc:72-74
	}
	ix->dict.erase(label);
	ix->version++;
	const uint64_t last = ix->size - 1;
	if (cur != last) {  // move the last row into the hole (bruteforce.cc:78-82)
		ix->touchRows(cur, uint64_t(cur) + 1);
		const uint64_t lastLabel = ix->h_labels[last];
		ix->dict.put(lastLabel, cur);
		ix->h_labels[cur] = lastLabel;
		RX_CUDA(cudaMemcpyAsync(ix->d_rows + size_t(cur) * ix->pitch, ix->d_rows + size_t(last) * ix->pitch, ix->pitch * sizeof(float),
								cudaMemcpyDeviceToDevice, ix->stream));
		RX_CUDA(cudaMemcpyAsync(ix->d_labels + cur, ix->d_labels + last, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ix->stream));
		if (ix->d_norms) {
			RX_CUDA(cudaMemcpyAsync(ix->d_norms + cur, ix->d_norms + last, sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
		}
		if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
			std::memcpy(ix->h_rows.data() + size_t(cur) * ix->dim, ix->h_rows.data() + size_t(last) * ix->dim, ix->dim * sizeof(float));
		}
		RX_CUDA(cudaStreamSynchronize(ix->stream));
	}
	ix->h_labels.pop_back();
	ix->size--;
	return 0;
}

int rxgpu_index_get(const rxgpu_index* ix, uint64_t label, const float** host_row) {
	if (!ix || !host_row) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null handle");
	}
	const uint32_t idx = ix->dict.find(label);
	if (idx == LabelMap::kNotFound) {
		return fail(RXGPU_ERR_NOT_FOUND, "Label not found");
	}
	if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
		*host_row = ix->h_rows.data() + size_t(idx) * ix->dim;
		return 0;
	}
	RX_CUDA(cudaSetDevice(ix->device));
	g_row_scratch.resize(ix->dim);
	RX_CUDA(cudaMemcpy(g_row_scratch.data(), ix->d_rows + size_t(idx) * ix->pitch, ix->dim * sizeof(float), cudaMemcpyDeviceToHost));
	*host_row = g_row_scratch.data();
	return 0;
}

uint64_t rxgpu_index_size(const rxgpu_index* ix) { return ix ? ix->size : 0; }
uint64_t rxgpu_index_capacity(const rxgpu_index* ix) { return ix ? ix->capacity : 0; }
uint64_t rxgpu_index_element_size(const rxgpu_index* ix) { return ix ? uint64_t(ix->dim) * 4 + 8 : 0; }
uint64_t rxgpu_index_device_bytes(const rxgpu_index* ix) {
	if (!ix) {
		return 0;
	}
	const uint64_t cap = ix->capacity ? ix->capacity : 1;
	return cap * ix->pitch * 4 + cap * 8 + (ix->d_norms ? cap * 4 : 0);
}
uint32_t rxgpu_index_dim(const rxgpu_index* ix) { return ix ? ix->dim : 0; }
int rxgpu_index_metric(const rxgpu_index* ix) { return ix ? ix->metric : -1; }
int rxgpu_index_device(const rxgpu_index* ix) { return ix ? ix->device : -1; }

int rxgpu_set_query_tile(rxgpu_index* ix, uint32_t qt) {
	if (!ix) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null handle");
	}
	ix->qt_override = qt;
	return 0;
}
int rxgpu_set_tensor_core_filter(rxgpu_index* ix, int mode) {
	if (!ix || mode < 0 || mode > 17 || mode == 7 || mode == 8 || (mode >= 10 && mode <= 13)) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: tensor-core filter mode must be 0..6, 9 or 14..17");
	}
	ix->tc_mode = uint32_t(mode >= 3 ? 1 : mode);
	ix->tc_variant = (mode == 3 || mode == 4) ? uint32_t(mode) : ((mode >= 14 && mode <= 16) ? 14u : 0u);
	ix->tc_tail = mode == 17 ? 0u : 1u;
	ix->tc_cluster_max = mode == 5 ? 1u : ((mode == 6 || mode == 14) ? 4u : ((mode == 9 || mode == 16) ? 8u : (mode == 15 ? 2u : 0u)));
	return 0;
}
int rxgpu_set_profile(int on) {
	g_profile.store(on ? 1 : 0);
	return 0;
}
void rxgpu_last_search_stats(rxgpu_search_stats* out) {
	if (out) {
		*out = g_stats;
	}
}

// ---------------------------------------------------------------------------------------------------------------- search
int rxgpu_search_knn_device(const rxgpu_index* ix, uint32_t nq, const float* d_queries, uint32_t k1, float* d_out_dist,
							uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count, void* stream) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	g_stats = rxgpu_search_stats{};
	if (nq == 0) {
		return 0;
	}
	if (k1 == 0 || k1 > kMaxSearchK1) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: k must be in [1, 65535]");
	}
	WsLease lease(ix);
	Workspace& ws = *lease.ws;
	cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
	if (ix->size == 0) {
		RX_CUDA(cudaMemsetAsync(d_out_count, 0, nq * sizeof(uint32_t), st));
		RX_CUDA(cudaStreamSynchronize(st));
		return 0;
	}
	if (int rc = scanTopK(ix, ws, st, d_queries, nq, k1, kModeTopK, 0.f, d_out_dist, d_out_idx, d_out_label, d_out_count)) {
		return rc;
	}
	RX_CUDA(cudaStreamSynchronize(st));  // the workspace goes back to the pool: nothing of this call may still be running
	collectProfile();
	return 0;
}

int rxgpu_search_tie_rows_device(const rxgpu_index* ix, const float* d_query, float dstar, uint32_t k, float* d_out_dist,
								 uint32_t* d_out_idx, uint64_t* d_out_label, uint32_t* d_out_count, void* stream) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (k == 0 || k > kMaxSearchK1) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: k must be in [1, 65535]");
	}
	WsLease lease(ix);
	cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
	if (ix->size == 0) {
		RX_CUDA(cudaMemsetAsync(d_out_count, 0, sizeof(uint32_t), st));
		RX_CUDA(cudaStreamSynchronize(st));
		return 0;
	}
	if (int rc = scanTopK(ix, *lease.ws, st, d_query, 1, k, kModeTieRows, dstar, d_out_dist, d_out_idx, d_out_label, d_out_count)) {
		return rc;
	}
	RX_CUDA(cudaStreamSynchronize(st));
	g_stats.tie_replays += 1;
	return 0;
}

int rxgpu_merge_shards(uint32_t nshards, uint32_t nq, uint32_t k, uint32_t k1, const float* dist, const uint32_t* idx,
					   const uint64_t* label, const uint32_t* count, const uint64_t* shard_base, float* out_dist, uint64_t* out_gidx,
					   uint64_t* out_label, uint32_t* out_count, uint8_t* need_tie) {
	if (!dist || !idx || !label || !count || !shard_base || !out_dist || !out_gidx || !out_label || !out_count || !need_tie) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	try {
		std::vector<Hit> all;
		for (uint32_t q = 0; q < nq; ++q) {
			all.clear();
			for (uint32_t s = 0; s < nshards; ++s) {
				const size_t b = (size_t(s) * nq + q) * k1;
				const uint32_t c = std::min(count[size_t(s) * nq + q], k1);
				for (uint32_t j = 0; j < c; ++j) {
					all.push_back(Hit{dist[b + j], shard_base[s] + idx[b + j], label[b + j]});
				}
			}
			std::sort(all.begin(), all.end(), hitLessByIndex);
			const uint32_t n = uint32_t(std::min<size_t>(all.size(), k));
			// a tie straddling the k-th place: the reference's survivors depend on arrival order and labels
			need_tie[q] = all.size() > k && k > 0 && !(all[k - 1].dist < all[k].dist) ? 1 : 0;
			std::vector<Hit> top(all.begin(), all.begin() + n);
			orderTiesByLabel(top);
			for (uint32_t j = 0; j < n; ++j) {
				out_dist[size_t(q) * k + j] = top[j].dist;
				out_gidx[size_t(q) * k + j] = top[j].gidx;
				out_label[size_t(q) * k + j] = top[j].label;
			}
			out_count[q] = n;
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_tie_replay(uint32_t k, float dstar, uint32_t n_lower, const float* lower_dist, const uint64_t* lower_gidx,
					 const uint64_t* lower_label, uint32_t n_first, const float* first_dist, const uint64_t* first_gidx,
					 const uint64_t* first_label, float* out_dist, uint64_t* out_label, uint32_t* out_count) {
	try {
		std::vector<Hit> lower(n_lower), first(n_first);
		for (uint32_t i = 0; i < n_lower; ++i) {
			lower[i] = Hit{lower_dist[i], lower_gidx[i], lower_label[i]};
		}
		for (uint32_t i = 0; i < n_first; ++i) {
			first[i] = Hit{first_dist[i], first_gidx[i], first_label[i]};
		}
		const auto res = tieReplay(k, dstar, lower, first);
		for (size_t i = 0; i < res.size(); ++i) {
			out_dist[i] = res[i].dist;
			out_label[i] = res[i].label;
		}
		*out_count = uint32_t(res.size());
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

static int searchKnnHost(const rxgpu_index* ix, uint32_t nq, const float* queries, uint32_t k, std::vector<std::vector<Hit>>& results) {
	results.assign(nq, {});
	g_stats = rxgpu_search_stats{};
	if (nq == 0 || k == 0 || ix->size == 0) {
		return 0;  // bruteforce.cc:106-108
	}
	const uint32_t kEff = uint32_t(std::min<uint64_t>(k, ix->size));           // bruteforce.cc:111
	const uint32_t k1 = uint32_t(std::min<uint64_t>(uint64_t(kEff) + 1, ix->size));  // one extra row exposes a tie at the k-th place
	if (k1 > kMaxSearchK1) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: k must be in [1, 65535]");
	}
	WsLease lease(ix);
	Workspace& ws = *lease.ws;
	if (!ws.stream) {
		RX_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
	}
	cudaStream_t st = ws.stream;
	const size_t qn = size_t(nq) * ix->dim, on = size_t(nq) * k1;
	RX_CUDA(ws.d_queries.ensure(qn));
	RX_CUDA(ws.h_queries.ensure(qn));
	RX_CUDA(ws.d_out_dist.ensure(on));
	RX_CUDA(ws.d_out_idx.ensure(on));
	RX_CUDA(ws.d_out_label.ensure(on));
	RX_CUDA(ws.d_out_count.ensure(nq));
	RX_CUDA(ws.h_out_dist.ensure(on));
	RX_CUDA(ws.h_out_idx.ensure(on));
	RX_CUDA(ws.h_out_label.ensure(on));
	RX_CUDA(ws.h_out_count.ensure(nq));
	std::memcpy(ws.h_queries.p, queries, qn * sizeof(float));
	RX_CUDA(cudaMemcpyAsync(ws.d_queries.p, ws.h_queries.p, qn * sizeof(float), cudaMemcpyHostToDevice, st));
	if (int rc = scanTopK(ix, ws, st, ws.d_queries.p, nq, k1, kModeTopK, 0.f, ws.d_out_dist.p, ws.d_out_idx.p, ws.d_out_label.p,
						  ws.d_out_count.p)) {
		return rc;
	}
	RX_CUDA(cudaMemcpyAsync(ws.h_out_dist.p, ws.d_out_dist.p, on * sizeof(float), cudaMemcpyDeviceToHost, st));
	RX_CUDA(cudaMemcpyAsync(ws.h_out_idx.p, ws.d_out_idx.p, on * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	RX_CUDA(cudaMemcpyAsync(ws.h_out_label.p, ws.d_out_label.p, on * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	RX_CUDA(cudaMemcpyAsync(ws.h_out_count.p, ws.d_out_count.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	RX_CUDA(cudaStreamSynchronize(st));
	collectProfile();

	for (uint32_t q = 0; q < nq; ++q) {
		const uint32_t cnt = std::min(ws.h_out_count.p[q], k1);
		const float* d = ws.h_out_dist.p + size_t(q) * k1;
		const uint32_t* ii = ws.h_out_idx.p + size_t(q) * k1;
		const uint64_t* ll = ws.h_out_label.p + size_t(q) * k1;
		std::vector<Hit>& res = results[q];
		const uint32_t n = std::min(cnt, kEff);
		const bool tie = cnt > kEff && !(d[kEff - 1] < d[kEff]);
		if (!tie) {
			res.reserve(n);
			for (uint32_t j = 0; j < n; ++j) {
				res.push_back(Hit{d[j], ii[j], ll[j]});
			}
			orderTiesByLabel(res);
			continue;
		}
		// replay the reference's heap tie rule: fetch the first kEff rows (internal order) with dist <= dstar -- from the filter's
		// candidate lists when this batch went through the tensor-core path (no second pass over the rows)
		const float dstar = d[kEff - 1];
		std::vector<Hit> lower;
		for (uint32_t j = 0; j < kEff && d[j] < dstar; ++j) {
			lower.push_back(Hit{d[j], ii[j], ll[j]});
		}
		RX_CUDA(ws.d_tie_dist.ensure(kEff));
		RX_CUDA(ws.d_tie_idx.ensure(kEff));
		RX_CUDA(ws.d_tie_label.ensure(kEff));
		RX_CUDA(ws.d_tie_count.ensure(1));
		if (int rc = tieRowsAfterScan(ix, ws, st, ws.d_queries.p, 1, &q, &dstar, kEff, ws.d_tie_dist.p, ws.d_tie_idx.p, ws.d_tie_label.p,
									  ws.d_tie_count.p)) {
			return rc;
		}
		std::vector<float> td(kEff);
		std::vector<uint32_t> ti(kEff);
		std::vector<uint64_t> tl(kEff);
		uint32_t tc = 0;
		RX_CUDA(cudaMemcpyAsync(td.data(), ws.d_tie_dist.p, kEff * sizeof(float), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(ti.data(), ws.d_tie_idx.p, kEff * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(tl.data(), ws.d_tie_label.p, kEff * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(&tc, ws.d_tie_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		std::vector<Hit> first;
		for (uint32_t j = 0; j < std::min(tc, kEff); ++j) {
			first.push_back(Hit{td[j], ti[j], tl[j]});
		}
		res = tieReplay(kEff, dstar, lower, first);
	}
	return 0;
}

int rxgpu_search_knn(const rxgpu_index* ix, uint32_t nq, const float* queries, uint32_t k, float* out_dist, uint64_t* out_label,
					 uint32_t* out_count) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (nq && (!queries || !out_count || (k && (!out_dist || !out_label)))) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	try {
		std::vector<std::vector<Hit>> results;
		if (int rc = searchKnnHost(ix, nq, queries, k, results)) {
			return rc;
		}
		for (uint32_t q = 0; q < nq; ++q) {
			const auto& r = results[q];
			for (size_t j = 0; j < r.size(); ++j) {
				out_dist[size_t(q) * k + j] = r[j].dist;
				out_label[size_t(q) * k + j] = r[j].label;
			}
			out_count[q] = uint32_t(r.size());
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

static int searchRangeHost(const rxgpu_index* ix, const float* query, float radius, std::vector<Hit>& res) {
	res.clear();
	g_stats = rxgpu_search_stats{};
	if (ix->size == 0) {
		return 0;
	}
	WsLease lease(ix);
	Workspace& ws = *lease.ws;
	if (!ws.stream) {
		RX_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
	}
	cudaStream_t st = ws.stream;
	RX_CUDA(ws.d_queries.ensure(ix->dim));
	RX_CUDA(ws.d_range_count.ensure(1));
	RX_CUDA(cudaMemcpyAsync(ws.d_queries.p, query, ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
	// result buffer: up to 4M matches (32 MB) without a rescan; a larger result grows the buffer and scans once more
	uint64_t cap = std::max<uint64_t>(ws.d_range.n, std::min<uint64_t>(std::max<uint64_t>(ix->size, 1), 1u << 22));
	for (;;) {
		RX_CUDA(ws.d_range.ensure(cap));
		RX_CUDA(cudaMemsetAsync(ws.d_range_count.p, 0, sizeof(unsigned long long), st));
		ScanArgs a{};
		a.rows = ix->d_rows;
		a.norm_coefs = ix->metric == RXGPU_COS ? ix->d_norms : nullptr;
		a.queries = ws.d_queries.p;
		a.pitch = ix->pitch;
		a.dim = ix->dim;
		a.row_begin = 0;
		a.row_end = uint32_t(ix->size);
		a.nq = 1;
		a.k1 = 1;
		a.mode = kModeRange;
		a.bound = radius;
		a.range_out = ws.d_range.p;
		a.range_count = ws.d_range_count.p;
		a.range_cap = cap;
		unsigned grid = 0;
		RX_CUDA(launchScan(ix, 1, a, &grid, st));
		g_stats.launches += 1;
		g_stats.passes += 1;
		unsigned long long total = 0;
		RX_CUDA(cudaMemcpyAsync(&total, ws.d_range_count.p, sizeof(total), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		if (total > cap) {  // buffer too small: grow and rescan (results are a set, the scan is deterministic)
			cap = total;
			continue;
		}
		RX_CUDA(ws.h_range.ensure(std::max<uint64_t>(total, 1)));
		RX_CUDA(cudaMemcpyAsync(ws.h_range.p, ws.d_range.p, total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		res.reserve(total);
		for (unsigned long long i = 0; i < total; ++i) {
			const uint64_t key = ws.h_range.p[i];
			const uint32_t row = uint32_t(key);
			res.push_back(Hit{ord_float(uint32_t(key >> 32)), row, ix->h_labels[row]});
		}
		break;
	}
	g_stats.query_tile = 1;
	g_stats.algorithmic_bytes = uint64_t(ix->size) * ix->dim * 4 + (ix->metric == RXGPU_COS ? uint64_t(ix->size) * 4 : 0) + ix->dim * 4 +
								res.size() * 8;
	std::sort(res.begin(), res.end(), hitLessByLabel);  // the order in which the reference's heap drains backwards
	return 0;
}

int rxgpu_search_range(const rxgpu_index* ix, const float* query, float radius, uint64_t max_out, float* out_dist, uint64_t* out_label,
					   uint64_t* out_n) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (!query || !out_n) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	try {
		std::vector<Hit>& res = g_range_result;
		if (int rc = searchRangeHost(ix, query, radius, res)) {
			return rc;
		}
		const uint64_t n = std::min<uint64_t>(res.size(), max_out);
		for (uint64_t i = 0; i < n; ++i) {
			out_dist[i] = res[i].dist;
			out_label[i] = res[i].label;
		}
		*out_n = res.size();
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_last_range_results(uint64_t offset, uint64_t n, float* out_dist, uint64_t* out_label) {
	if (offset > g_range_result.size() || n > g_range_result.size() - offset || (n && (!out_dist || !out_label))) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: range outside the retained result of this thread's last rxgpu_search_range");
	}
	for (uint64_t i = 0; i < n; ++i) {
		out_dist[i] = g_range_result[offset + i].dist;
		out_label[i] = g_range_result[offset + i].label;
	}
	return 0;
}

int rxgpu_select_postprocess(int metric, const rxgpu_select_params* p, uint64_t n, const float* dist, const uint64_t* label, int32_t* out_row_ids,
							 float* out_ranks, uint64_t* out_n) {
	if (!p || !out_n || (n && (!dist || !label || !out_row_ids || !out_ranks)) || metric < 0 || metric > 2) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: bad argument");
	}
	try {
		std::vector<Hit> res(n);
		for (uint64_t i = 0; i < n; ++i) {
			res[i] = Hit{dist[i], 0, label[i]};
		}
		SelectParams sp;
		sp.metric = metric;
		sp.needSort = p->need_sort != 0;
		sp.isArray = p->is_array != 0;
		sp.raw = p->raw != 0;
		sp.hasK = p->k != 0;
		sp.k = p->k;
		sp.hasRadius = p->has_radius != 0;
		std::vector<int32_t> ids;
		std::vector<float> ranks;
		selectPostprocess(sp, res, ids, ranks);
		for (size_t i = 0; i < ids.size(); ++i) {
			out_row_ids[i] = ids[i];
			out_ranks[i] = ranks[i];
		}
		*out_n = ids.size();
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_select_knn(const rxgpu_index* ix, const float* query, const rxgpu_select_params* p, uint64_t max_out, int32_t* out_row_ids,
					 float* out_ranks, uint64_t* out_n) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (!query || !p || !out_n) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	if (p->k == 0 && !p->has_radius) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: KNN query needs k or radius");
	}
	try {
		// HnswIndexBase::search, hnsw_index.cc:160-191
		std::vector<float> normalized;
		const float* keyData = query;
		if (ix->metric == RXGPU_COS) {
			normalized.resize(ix->dim);
			normalizeCopyVector(query, int32_t(ix->dim), normalized.data());
			keyData = normalized.data();
		}
		std::vector<Hit> res;
		if (p->has_radius) {
			if (int rc = searchRangeHost(ix, keyData, ix->metric == RXGPU_L2 ? p->radius : -p->radius, res)) {
				return rc;
			}
		} else {
			std::vector<std::vector<Hit>> results;
			if (int rc = searchKnnHost(ix, 1, keyData, p->k, results)) {
				return rc;
			}
			res = std::move(results[0]);
		}
		SelectParams sp;
		sp.metric = ix->metric;
		sp.needSort = p->need_sort != 0;
		sp.isArray = p->is_array != 0;
		sp.raw = p->raw != 0;
		sp.hasK = p->k != 0;
		sp.k = p->k;
		sp.hasRadius = p->has_radius != 0;
		std::vector<int32_t> ids;
		std::vector<float> ranks;
		selectPostprocess(sp, res, ids, ranks);
		const uint64_t n = std::min<uint64_t>(ids.size(), max_out);
		for (uint64_t i = 0; i < n; ++i) {
			out_row_ids[i] = ids[i];
			out_ranks[i] = ranks[i];
		}
		*out_n = ids.size();
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------- bench support
int rxgpu_index_append_synth(rxgpu_index* ix, uint64_t seed, uint64_t first_row, uint64_t n) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (ix->flags & RXGPU_FLAG_HOST_MIRROR) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: device-side synthetic fill is not available with a host mirror");
	}
	if (ix->size + n > ix->capacity) {
		return fail(RXGPU_ERR_LOGIC, "The number of elements exceeds the specified limit\n");
	}
	if (n == 0) {
		return 0;
	}
	try {
		ix->h_labels.reserve(ix->size + n);
		ix->dict.reserve(ix->size + n);
		for (uint64_t r = 0; r < n; ++r) {
			const uint64_t label = (first_row + r) << 32;
			if (ix->dict.find(label) != LabelMap::kNotFound) {
				return fail(RXGPU_ERR_LOGIC, "rxgpu: synthetic rows must have fresh labels");
			}
			ix->dict.put(label, uint32_t(ix->size + r));
			ix->h_labels.push_back(label);
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	synth_rows_kernel<<<unsigned(ix->sm_count) * 8, 256, 0, ix->stream>>>(ix->d_rows, ix->d_labels, ix->pitch, ix->dim, uint32_t(ix->size),
																		 seed, first_row, n);
	RX_CUDA(cudaGetLastError());
	if (int rc = normsForRange(ix, ix->size, ix->size + n)) {
		return rc;
	}
	RX_CUDA(cudaStreamSynchronize(ix->stream));
	ix->touchRows(ix->size, ix->size + n);
	ix->size += n;
	ix->version++;
	return 0;
}

int rxgpu_synth_fill_device(float* d_out, uint64_t seed, uint64_t first_index, uint64_t count, int device, void* stream) {
	if (rxgpu_device_count() <= device || device < 0) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: no usable CUDA device (this library has no CPU fallback)");
	}
	RX_CUDA(cudaSetDevice(device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	synth_fill_kernel<<<1184, 256, 0, st>>>(d_out, seed, first_index, count);
	RX_CUDA(cudaGetLastError());
	RX_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------- IVF
}  // extern "C"

struct rxgpu_ivf_device {
	uint32_t nlist = 0;
	uint64_t index_version = 0;
	DevBuf<float> centroids;       // [nlist][pitch]
	DevBuf<float> cnorm;           // Cosine: 1/||centroid|| (IndexFlatCosine's norm coefficients)
	DevBuf<uint32_t> list_begin;   // [nlist + 1] rows of list l = [list_begin[l], list_begin[l + 1])
	std::mutex mtx;                // one IVF batch at a time per index (scratch below)
	DevBuf<float> d_q, d_dist;
	DevBuf<uint4> d_work;
	DevBuf<uint64_t> d_lists, d_label;
	DevBuf<uint32_t> d_idx, d_count;
	DevBuf<uint64_t> d_range;
	DevBuf<unsigned long long> d_range_count;
	// mutable lists (rxgpu_ivf_create / _add / _remove): every list owns a region [begin, begin + cap) of a row slab; size <= cap
	bool own = false;
	float* rows = nullptr;       // [slab_rows][pitch]
	uint64_t* labels = nullptr;  // [slab_rows]
	float* norms = nullptr;      // [slab_rows] (Cosine)
	uint64_t slab_rows = 0, high_water = 0, live = 0, dead = 0;
	std::vector<uint32_t> begin, size, cap;
	DevBuf<uint32_t> list_end;   // begin + size (list_begin holds begin)
	std::vector<uint64_t> h_slab_labels;                                   // host mirror of `labels`
	std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> where;     // id -> (list, offset), faiss::DirectMap::Hashtable
	DevBuf<float> st_rows;
	DevBuf<uint32_t> st_dst;
	DevBuf<uint64_t> st_labels;
	uint64_t relocations = 0, compactions = 0;
	~rxgpu_ivf_device() {
		cudaFree(rows);
		cudaFree(labels);
		cudaFree(norms);
	}
};
namespace rxgpu {
void ivfRelease(rxgpu_ivf_device* p) { delete p; }
}  // namespace rxgpu

extern "C" {

int rxgpu_ivf_import(rxgpu_index* ix, uint32_t nlist, const float* centroids, const uint64_t* list_sizes) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (!centroids || !list_sizes || nlist == 0) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	if (nlist > 16384) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: at most 16384 IVF centroids on the device path");
	}
	try {
		std::vector<uint32_t> begin(size_t(nlist) + 1, 0u);
		uint64_t total = 0;
		for (uint32_t l = 0; l < nlist; ++l) {
			total += list_sizes[l];
			if (total > ix->size) {
				break;
			}
			begin[l + 1] = uint32_t(total);
		}
		if (total != ix->size) {
			return fail(RXGPU_ERR_LOGIC, "rxgpu: IVF list sizes do not add up to the number of rows in the index");
		}
		auto h = std::make_unique<rxgpu_ivf_device>();
		h->nlist = nlist;
		RX_CUDA(h->centroids.ensure(size_t(nlist) * ix->pitch));
		RX_CUDA(h->list_begin.ensure(size_t(nlist) + 1));
		RX_CUDA(cudaMemset(h->centroids.p, 0, size_t(nlist) * ix->pitch * sizeof(float)));
		RX_CUDA(cudaMemcpy2D(h->centroids.p, size_t(ix->pitch) * 4, centroids, size_t(ix->dim) * 4, size_t(ix->dim) * 4, nlist,
							 cudaMemcpyHostToDevice));
		RX_CUDA(cudaMemcpy(h->list_begin.p, begin.data(), begin.size() * 4, cudaMemcpyHostToDevice));
		if (ix->metric == RXGPU_COS) {
			RX_CUDA(h->cnorm.ensure(nlist));
			norm_coef_kernel<<<(nlist * 32 + 255) / 256, 256, 0, ix->stream>>>(h->centroids.p, ix->pitch, ix->dim, 0, nlist, h->cnorm.p);
			RX_CUDA(cudaGetLastError());
			RX_CUDA(cudaStreamSynchronize(ix->stream));
		}
		h->index_version = ix->version;
		if (ix->ivf) {
			ivfRelease(ix->ivf);
		}
		ix->ivf = h.release();
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_ivf_search_knn(const rxgpu_index* ix, uint32_t nq, const float* queries, uint32_t k, uint32_t nprobe, float* out_dist,
						 uint64_t* out_label, uint32_t* out_count) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	g_stats = rxgpu_search_stats{};
	if (nq == 0) {
		return 0;
	}
	if (!queries || !out_count || (k && (!out_dist || !out_label))) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	rxgpu_ivf_device* h = ix->ivf;
	if (!h) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: no IVF lists imported into this index");
	}
	if (h->index_version != ix->version) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: the index changed after the IVF lists were imported");
	}
	if (k == 0 || k > kMaxFusedK1) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: IVF search needs k in [1, 256]");
	}
	nprobe = std::max(1u, std::min(nprobe, h->nlist));  // faiss::IndexIVF::search clamps nprobe to nlist
	if (nprobe > 256u * kMergeOwn) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: nprobe exceeds the merge fan-in (1024)");
	}
	const uint32_t nch = (ix->dim + 127u) / 128u;
	const size_t coarseSmem = size_t(nch) * 512 + size_t(h->nlist) * 8;
	if (coarseSmem > 200 * 1024) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: dimension / centroid count exceeds the coarse quantiser's shared memory");
	}
	std::lock_guard<std::mutex> lck(h->mtx);
	cudaStream_t st = ix->stream;
	const size_t nwork = size_t(nq) * nprobe;
	RX_CUDA(h->d_q.ensure(size_t(nq) * ix->dim));
	RX_CUDA(h->d_work.ensure(nwork));
	RX_CUDA(h->d_lists.ensure(nwork * k));
	RX_CUDA(h->d_dist.ensure(size_t(nq) * k));
	RX_CUDA(h->d_idx.ensure(size_t(nq) * k));
	RX_CUDA(h->d_label.ensure(size_t(nq) * k));
	RX_CUDA(h->d_count.ensure(nq));
	RX_CUDA(cudaMemcpyAsync(h->d_q.p, queries, size_t(nq) * ix->dim * 4, cudaMemcpyHostToDevice, st));
	if (ix->metric == RXGPU_L2) {
		RX_CUDA(raiseSmemCeilingOnce(ivf_coarse_kernel<true>, ix->device, 200 * 1024));
		ivf_coarse_kernel<true><<<nq, kScanThreads, coarseSmem, st>>>(h->centroids.p, ix->pitch, ix->dim, h->nlist, h->d_q.p, nq, nprobe,
																	   h->list_begin.p, h->own ? h->list_end.p : nullptr, nullptr, h->d_work.p);
	} else {
		RX_CUDA(raiseSmemCeilingOnce(ivf_coarse_kernel<false>, ix->device, 200 * 1024));
		ivf_coarse_kernel<false><<<nq, kScanThreads, coarseSmem, st>>>(h->centroids.p, ix->pitch, ix->dim, h->nlist, h->d_q.p, nq, nprobe,
																		h->list_begin.p, h->own ? h->list_end.p : nullptr, ix->metric == RXGPU_COS ? h->cnorm.p : nullptr, h->d_work.p);
	}
	RX_CUDA(cudaGetLastError());
	// list scans: the exact scan kernel in work-item mode, one CTA per (query, probed list), fused top-k per CTA
	ScanArgs a{};
	a.rows = h->own ? h->rows : ix->d_rows;
	a.norm_coefs = ix->metric != RXGPU_COS ? nullptr : h->own ? h->norms : ix->d_norms;
	a.queries = h->d_q.p;
	a.pitch = ix->pitch;
	a.dim = ix->dim;
	a.nq = 1;
	a.k1 = k;
	a.mode = kModeTopK;
	a.lists = h->d_lists.p;
	a.work = h->d_work.p;
	a.nwork = uint32_t(nwork);
	if (scan_smem_bytes(1, ix->dim, k) > 100 * 1024) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: dimension/k combination exceeds the fused top-k shared-memory budget");
	}
	unsigned grid = 0;
	RX_CUDA(launchScan(ix, 1, a, &grid, st));
	MergeArgs m{};
	m.lists = h->d_lists.p;
	m.labels = h->own ? h->labels : ix->d_labels;
	m.out_dist = h->d_dist.p;
	m.out_idx = h->d_idx.p;
	m.out_label = h->d_label.p;
	m.out_count = h->d_count.p;
	m.nlists = nprobe;
	m.qt = nq;  // lists are probe-major: list of (probe p, query q) = p * nq + q
	m.k1 = k;
	m.q_offset = 0;
	m.out_stride = k;
	m.out_offset = 0;
	m.mode = kModeTopK;
	knn_merge_lists<<<nq, 256, 0, st>>>(m);
	RX_CUDA(cudaGetLastError());
	g_stats.launches = 3;
	g_stats.passes = 1;
	try {
		std::vector<float> hd(size_t(nq) * k);
		std::vector<uint64_t> hl(size_t(nq) * k);
		std::vector<uint32_t> hi(size_t(nq) * k), hc(nq);
		RX_CUDA(cudaMemcpyAsync(hd.data(), h->d_dist.p, hd.size() * 4, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(hl.data(), h->d_label.p, hl.size() * 8, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(hi.data(), h->d_idx.p, hi.size() * 4, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(hc.data(), h->d_count.p, hc.size() * 4, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		std::vector<Hit> hits;
		for (uint32_t q = 0; q < nq; ++q) {
			hits.clear();
			for (uint32_t j = 0; j < std::min(hc[q], k); ++j) {
				hits.push_back(Hit{hd[size_t(q) * k + j], hi[size_t(q) * k + j], hl[size_t(q) * k + j]});
			}
			orderTiesByLabel(hits);  // FAISS' heap leaves bit-equal distances in no particular order: (distance, label) here
			for (size_t j = 0; j < hits.size(); ++j) {
				out_dist[size_t(q) * k + j] = hits[j].dist;
				out_label[size_t(q) * k + j] = hits[j].label;
			}
			out_count[q] = uint32_t(hits.size());
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

int rxgpu_ivf_search_range(const rxgpu_index* ix, const float* query, float radius, uint32_t nprobe, uint64_t max_out, float* out_dist,
						   uint64_t* out_label, uint64_t* out_n) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (!query || !out_n || (max_out && (!out_dist || !out_label))) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	*out_n = 0;
	g_stats = rxgpu_search_stats{};
	rxgpu_ivf_device* h = ix->ivf;
	if (!h) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: no IVF lists imported into this index");
	}
	if (h->index_version != ix->version) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: the index changed after the IVF lists were imported");
	}
	nprobe = std::max(1u, std::min(nprobe, h->nlist));
	const uint32_t nch = (ix->dim + 127u) / 128u;
	const size_t coarseSmem = size_t(nch) * 512 + size_t(h->nlist) * 8;
	if (coarseSmem > 200 * 1024) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: dimension / centroid count exceeds the coarse quantiser's shared memory");
	}
	std::lock_guard<std::mutex> lck(h->mtx);
	cudaStream_t st = ix->stream;
	RX_CUDA(h->d_q.ensure(ix->dim));
	RX_CUDA(h->d_work.ensure(nprobe));
	RX_CUDA(h->d_range_count.ensure(1));
	RX_CUDA(cudaMemcpyAsync(h->d_q.p, query, size_t(ix->dim) * 4, cudaMemcpyHostToDevice, st));
	if (ix->metric == RXGPU_L2) {
		RX_CUDA(raiseSmemCeilingOnce(ivf_coarse_kernel<true>, ix->device, 200 * 1024));
		ivf_coarse_kernel<true><<<1, kScanThreads, coarseSmem, st>>>(h->centroids.p, ix->pitch, ix->dim, h->nlist, h->d_q.p, 1, nprobe,
																	  h->list_begin.p, h->own ? h->list_end.p : nullptr, nullptr, h->d_work.p);
	} else {
		RX_CUDA(raiseSmemCeilingOnce(ivf_coarse_kernel<false>, ix->device, 200 * 1024));
		ivf_coarse_kernel<false><<<1, kScanThreads, coarseSmem, st>>>(h->centroids.p, ix->pitch, ix->dim, h->nlist, h->d_q.p, 1, nprobe,
																	   h->list_begin.p, h->own ? h->list_end.p : nullptr,
																	   ix->metric == RXGPU_COS ? h->cnorm.p : nullptr, h->d_work.p);
	}
	RX_CUDA(cudaGetLastError());
	g_stats.launches = 1;
	uint64_t cap = std::max<uint64_t>(h->d_range.n, 1u << 14);
	unsigned long long total = 0;
	for (;;) {  // like the brute-force range search: grow the result buffer and rescan when it was too small
		RX_CUDA(h->d_range.ensure(cap));
		RX_CUDA(cudaMemsetAsync(h->d_range_count.p, 0, sizeof(unsigned long long), st));
		ScanArgs a{};
		a.rows = h->own ? h->rows : ix->d_rows;
		a.norm_coefs = ix->metric != RXGPU_COS ? nullptr : h->own ? h->norms : ix->d_norms;
		a.queries = h->d_q.p;
		a.pitch = ix->pitch;
		a.dim = ix->dim;
		a.nq = 1;
		a.k1 = 1;
		a.mode = kModeRange;
		a.bound = radius;
		a.range_out = h->d_range.p;
		a.range_count = h->d_range_count.p;
		a.range_cap = cap;
		a.work = h->d_work.p;
		a.nwork = nprobe;
		unsigned grid = 0;
		RX_CUDA(launchScan(ix, 1, a, &grid, st));
		g_stats.launches += 1;
		RX_CUDA(cudaMemcpyAsync(&total, h->d_range_count.p, sizeof(total), cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		if (total <= cap) {
			break;
		}
		cap = total;
	}
	*out_n = total;
	try {
		std::vector<uint64_t> keys(total);
		if (total) {
			RX_CUDA(cudaMemcpy(keys.data(), h->d_range.p, total * sizeof(uint64_t), cudaMemcpyDeviceToHost));
		}
		std::vector<Hit> res(total);
		for (unsigned long long i = 0; i < total; ++i) {
			const uint32_t row = uint32_t(keys[i]);
			res[i] = Hit{ord_float(uint32_t(keys[i] >> 32)), row, h->own ? h->h_slab_labels[row] : ix->h_labels[row]};
		}
		std::sort(res.begin(), res.end(), hitLessByLabel);  // IvfIndex sorts the range result by distance (ivf_index.cc:220-224); ties by label here
		const uint64_t nout = std::min<uint64_t>(total, max_out);
		for (uint64_t j = 0; j < nout; ++j) {
			out_dist[j] = res[j].dist;
			out_label[j] = res[j].label;
		}
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

}  // extern "C"

// ---- mutable IVF lists -----------------------------------------------------------------------------------------------------------
namespace {
// grows the slab to at least `want` rows, keeping its contents
int ivfGrowSlab(rxgpu_index* ix, rxgpu_ivf_device* h, uint64_t want) {
	if (want <= h->slab_rows) {
		return 0;
	}
	if (want > 0xFFFFFFF0ull) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: IVF row slab exceeds 2^32 rows");
	}
	const uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(want, h->slab_rows + h->slab_rows / 2), 0xFFFFFFF0ull);
	float* rows = nullptr;
	uint64_t* labels = nullptr;
	float* norms = nullptr;
	RX_CUDA(cudaMalloc(reinterpret_cast<void**>(&rows), cap * ix->pitch * sizeof(float)));
	cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&labels), cap * sizeof(uint64_t));
	if (e == cudaSuccess && ix->metric == RXGPU_COS) {
		e = cudaMalloc(reinterpret_cast<void**>(&norms), cap * sizeof(float));
	}
	if (e != cudaSuccess) {
		cudaFree(rows);
		cudaFree(labels);
		return fail(RXGPU_ERR_SYSTEM, std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaMalloc (IVF slab)");
	}
	cudaStream_t st = ix->stream;
	if (h->high_water) {
		cudaMemcpyAsync(rows, h->rows, h->high_water * ix->pitch * sizeof(float), cudaMemcpyDeviceToDevice, st);
		cudaMemcpyAsync(labels, h->labels, h->high_water * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
		if (norms) {
			cudaMemcpyAsync(norms, h->norms, h->high_water * sizeof(float), cudaMemcpyDeviceToDevice, st);
		}
	}
	e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) {
		cudaFree(rows);
		cudaFree(labels);
		cudaFree(norms);
		return fail(RXGPU_ERR_SYSTEM, std::string("CUDA error: ") + cudaGetErrorString(e) + " at the IVF slab copy");
	}
	cudaFree(h->rows);
	cudaFree(h->labels);
	cudaFree(h->norms);
	h->rows = rows;
	h->labels = labels;
	h->norms = norms;
	h->slab_rows = cap;
	h->h_slab_labels.resize(cap);
	return 0;
}
// moves list l to the end of the slab with room for `need` rows (amortised x1.5 growth); the old region becomes dead space
int ivfRelocate(rxgpu_index* ix, rxgpu_ivf_device* h, uint32_t l, uint32_t need) {
	const uint32_t newCap = std::max<uint32_t>(32u, (need + need / 2 + 31u) & ~31u);
	if (int rc = ivfGrowSlab(ix, h, h->high_water + newCap)) {
		return rc;
	}
	cudaStream_t st = ix->stream;
	const uint64_t src = h->begin[l], dst = h->high_water, n = h->size[l];
	if (n) {
		RX_CUDA(cudaMemcpyAsync(h->rows + dst * ix->pitch, h->rows + src * ix->pitch, n * ix->pitch * sizeof(float), cudaMemcpyDeviceToDevice, st));
		RX_CUDA(cudaMemcpyAsync(h->labels + dst, h->labels + src, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
		if (h->norms) {
			RX_CUDA(cudaMemcpyAsync(h->norms + dst, h->norms + src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
		}
		std::copy(h->h_slab_labels.begin() + src, h->h_slab_labels.begin() + src + n, h->h_slab_labels.begin() + dst);
	}
	h->dead += h->cap[l];
	h->begin[l] = uint32_t(dst);
	h->cap[l] = newCap;
	h->high_water += newCap;
	h->relocations++;
	return 0;
}
// rewrites the slab without the dead regions (every list keeps 25 % slack)
int ivfCompact(rxgpu_index* ix, rxgpu_ivf_device* h) {
	uint64_t total = 0;
	std::vector<uint32_t> newCap(h->nlist);
	for (uint32_t l = 0; l < h->nlist; ++l) {
		newCap[l] = std::max<uint32_t>(32u, (h->size[l] + h->size[l] / 4 + 31u) & ~31u);
		total += newCap[l];
	}
	float* rows = nullptr;
	uint64_t* labels = nullptr;
	float* norms = nullptr;
	RX_CUDA(cudaMalloc(reinterpret_cast<void**>(&rows), total * ix->pitch * sizeof(float)));
	cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&labels), total * sizeof(uint64_t));
	if (e == cudaSuccess && ix->metric == RXGPU_COS) {
		e = cudaMalloc(reinterpret_cast<void**>(&norms), total * sizeof(float));
	}
	if (e != cudaSuccess) {
		cudaFree(rows);
		cudaFree(labels);
		return fail(RXGPU_ERR_SYSTEM, std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaMalloc (IVF slab)");
	}
	cudaStream_t st = ix->stream;
	std::vector<uint64_t> hl(total);
	uint64_t at = 0;
	for (uint32_t l = 0; l < h->nlist; ++l) {
		const uint64_t src = h->begin[l], n = h->size[l];
		if (n) {
			cudaMemcpyAsync(rows + at * ix->pitch, h->rows + src * ix->pitch, n * ix->pitch * sizeof(float), cudaMemcpyDeviceToDevice, st);
			cudaMemcpyAsync(labels + at, h->labels + src, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
			if (norms) {
				cudaMemcpyAsync(norms + at, h->norms + src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
			}
			std::copy(h->h_slab_labels.begin() + src, h->h_slab_labels.begin() + src + n, hl.begin() + at);
		}
		h->begin[l] = uint32_t(at);
		h->cap[l] = newCap[l];
		at += newCap[l];
	}
	e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) {
		cudaFree(rows);
		cudaFree(labels);
		cudaFree(norms);
		return fail(RXGPU_ERR_SYSTEM, std::string("CUDA error: ") + cudaGetErrorString(e) + " at the IVF slab copy");
	}
	cudaFree(h->rows);
	cudaFree(h->labels);
	cudaFree(h->norms);
	h->rows = rows;
	h->labels = labels;
	h->norms = norms;
	h->slab_rows = total;
	h->high_water = total;
	h->dead = 0;
	h->h_slab_labels.swap(hl);
	h->compactions++;
	return 0;
}
int ivfPushBounds(rxgpu_index* ix, rxgpu_ivf_device* h) {
	std::vector<uint32_t> end(h->nlist);
	for (uint32_t l = 0; l < h->nlist; ++l) {
		end[l] = h->begin[l] + h->size[l];
	}
	RX_CUDA(cudaMemcpyAsync(h->list_begin.p, h->begin.data(), size_t(h->nlist) * 4, cudaMemcpyHostToDevice, ix->stream));
	RX_CUDA(cudaMemcpyAsync(h->list_end.p, end.data(), size_t(h->nlist) * 4, cudaMemcpyHostToDevice, ix->stream));
	RX_CUDA(cudaStreamSynchronize(ix->stream));
	return 0;
}
}  // namespace

extern "C" {

int rxgpu_ivf_create(rxgpu_index* ix, uint32_t nlist, const float* centroids) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (ix->size != 0) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: rxgpu_ivf_create needs an empty index (the rows live in the lists)");
	}
	const std::vector<uint64_t> zeros(nlist ? nlist : 1, 0);
	if (int rc = rxgpu_ivf_import(ix, nlist, centroids, zeros.data())) {
		return rc;
	}
	rxgpu_ivf_device* h = ix->ivf;
	try {
		h->own = true;
		h->begin.assign(nlist, 0u);
		h->size.assign(nlist, 0u);
		h->cap.assign(nlist, 0u);
		RX_CUDA(h->list_end.ensure(nlist));
		return ivfPushBounds(ix, h);
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
}

int rxgpu_ivf_add(rxgpu_index* ix, uint64_t n, const uint32_t* list_nos, const uint64_t* labels, const float* vecs) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	rxgpu_ivf_device* h = ix->ivf;
	if (!h || !h->own) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: rxgpu_ivf_add needs lists made by rxgpu_ivf_create");
	}
	if (n == 0) {
		return 0;
	}
	if (!list_nos || !labels || !vecs) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	std::lock_guard<std::mutex> lck(h->mtx);
	try {
		std::unordered_set<uint64_t> seen;
		std::vector<uint32_t> adds(h->nlist, 0u);
		for (uint64_t i = 0; i < n; ++i) {
			if (list_nos[i] >= h->nlist) {
				return fail(RXGPU_ERR_PARAMS, "rxgpu: IVF list number out of range");
			}
			if (h->where.count(labels[i]) || !seen.insert(labels[i]).second) {
				return fail(RXGPU_ERR_LOGIC, "rxgpu: the id is already in the IVF lists");
			}
			adds[list_nos[i]]++;
		}
		if (h->dead > h->live + n + 4096) {  // more dead space than rows: rewrite the slab before growing it further
			if (int rc = ivfCompact(ix, h)) {
				return rc;
			}
		}
		for (uint32_t l = 0; l < h->nlist; ++l) {
			if (adds[l] && h->size[l] + adds[l] > h->cap[l]) {
				if (int rc = ivfRelocate(ix, h, l, h->size[l] + adds[l])) {
					return rc;
				}
			}
		}
		cudaStream_t st = ix->stream;
		const uint64_t slice = std::max<uint64_t>(1, (uint64_t(64) << 20) / (size_t(ix->dim) * 4));
		std::vector<uint32_t> dst;
		for (uint64_t off = 0; off < n; off += slice) {
			const uint64_t cnt = std::min(slice, n - off);
			dst.resize(cnt);
			for (uint64_t i = 0; i < cnt; ++i) {
				const uint32_t l = list_nos[off + i];
				const uint32_t row = h->begin[l] + h->size[l];
				dst[i] = row;
				h->where.emplace(labels[off + i], std::make_pair(l, h->size[l]));
				h->h_slab_labels[row] = labels[off + i];
				h->size[l]++;
			}
			RX_CUDA(h->st_rows.ensure(cnt * ix->dim));
			RX_CUDA(h->st_dst.ensure(cnt));
			RX_CUDA(h->st_labels.ensure(cnt));
			RX_CUDA(cudaMemcpyAsync(h->st_rows.p, vecs + off * ix->dim, cnt * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
			RX_CUDA(cudaMemcpyAsync(h->st_dst.p, dst.data(), cnt * 4, cudaMemcpyHostToDevice, st));
			RX_CUDA(cudaMemcpyAsync(h->st_labels.p, labels + off, cnt * 8, cudaMemcpyHostToDevice, st));
			scatter_rows_kernel<<<unsigned(cnt), 128, 0, st>>>(h->st_rows.p, h->st_dst.p, h->st_labels.p, uint32_t(cnt), ix->dim, ix->pitch, h->rows,
															   h->labels);
			if (h->norms) {
				norm_coef_at_kernel<<<unsigned((cnt * 32 + 255) / 256), 256, 0, st>>>(h->rows, ix->pitch, ix->dim, h->st_dst.p, uint32_t(cnt), h->norms);
			}
			RX_CUDA(cudaGetLastError());
			RX_CUDA(cudaStreamSynchronize(st));  // `dst` is reused by the next slice
		}
		h->live += n;
		return ivfPushBounds(ix, h);
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
}

int rxgpu_ivf_remove(rxgpu_index* ix, uint64_t label) {
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	rxgpu_ivf_device* h = ix->ivf;
	if (!h || !h->own) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: rxgpu_ivf_remove needs lists made by rxgpu_ivf_create");
	}
	std::lock_guard<std::mutex> lck(h->mtx);
	const auto it = h->where.find(label);
	if (it == h->where.end()) {
		return fail(RXGPU_ERR_NOT_FOUND, "rxgpu: the id is not in the IVF lists");
	}
	// InvertedLists swap-remove (faiss DirectMap::remove_ids, Hashtable flavour): the list's last entry fills the hole
	const uint32_t l = it->second.first, pos = it->second.second, last = h->size[l] - 1;
	const uint64_t at = uint64_t(h->begin[l]) + pos, from = uint64_t(h->begin[l]) + last;
	cudaStream_t st = ix->stream;
	if (pos != last) {
		RX_CUDA(cudaMemcpyAsync(h->rows + at * ix->pitch, h->rows + from * ix->pitch, size_t(ix->pitch) * sizeof(float), cudaMemcpyDeviceToDevice, st));
		RX_CUDA(cudaMemcpyAsync(h->labels + at, h->labels + from, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
		if (h->norms) {
			RX_CUDA(cudaMemcpyAsync(h->norms + at, h->norms + from, sizeof(float), cudaMemcpyDeviceToDevice, st));
		}
		const uint64_t moved = h->h_slab_labels[from];
		h->h_slab_labels[at] = moved;
		h->where[moved].second = pos;
	}
	h->where.erase(it);
	h->size[l] = last;
	h->live--;
	const uint32_t end = h->begin[l] + last;
	RX_CUDA(cudaMemcpyAsync(h->list_end.p + l, &end, 4, cudaMemcpyHostToDevice, st));
	RX_CUDA(cudaStreamSynchronize(st));
	return 0;
}

uint64_t rxgpu_ivf_size(const rxgpu_index* ix) { return ix && ix->ivf ? (ix->ivf->own ? ix->ivf->live : ix->size) : 0; }
int rxgpu_ivf_list_stats(const rxgpu_index* ix, uint64_t* slab_rows, uint64_t* dead_rows, uint64_t* relocations, uint64_t* compactions) {
	if (!ix || !ix->ivf) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: no IVF lists in this index");
	}
	const rxgpu_ivf_device* h = ix->ivf;
	if (slab_rows) *slab_rows = h->slab_rows;
	if (dead_rows) *dead_rows = h->dead;
	if (relocations) *relocations = h->relocations;
	if (compactions) *compactions = h->compactions;
	return 0;
}

}  // extern "C"
