// Multi-GPU brute-force KNN behind the C ABI: the namespace is sharded by internal-row range, one shard per GPU / rank.
// One call = local fused scan + top-(k+1) on this rank's shard, ONE ncclAllGather of the per-shard lists over NVLink / NVSwitch, a
// device-side k-way merge under the reference's comparator, and -- only when bit-equal distances straddle the k-th place -- the
// reference's heap tie rule (bruteforce.cc:103-127) replayed globally from the filter's per-query candidate lists (every row at or
// below the k-th distance is in them, so no shard is scanned a second time) with a second, tiny all-gather.
// NCCL is resolved with dlopen at the first rxgpu_comm_* call: librxgpu.so itself keeps loading on a box without NCCL or a GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rxgpu.h"
#include "../host/knn_select.h"
#include "internal.h"
#include "common.cuh"

using namespace rxgpu;

namespace {

struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
const NcclApi& nccl() {
	static NcclApi api = [] {
		NcclApi a;
		void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) {
			h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		}
		if (!h) {
			return a;
		}
		a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
		a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
		a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
		a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
		a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
		a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
		a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GetErrorString;
		return a;
	}();
	return api;
}
#define RX_NCCL(expr)                                                                                               \
	do {                                                                                                            \
		ncclResult_t r_ = (expr);                                                                                   \
		if (r_ != ncclSuccess) {                                                                                    \
			return fail(RXGPU_ERR_SYSTEM, std::string("NCCL error: ") + nccl().GetErrorString(r_) + " at " #expr); \
		}                                                                                                           \
	} while (0)

// Layout of one rank's contribution to the exchange (all sections 16-byte aligned):
//   [dist f32 nq*k1][idx u32 nq*k1][label u64 nq*k1][count u32 nq][size u64, pad]
struct PayloadLayout {
	size_t off_dist, off_idx, off_label, off_count, off_size, bytes;
	PayloadLayout(uint32_t nq, uint32_t k1) {
		auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
		const size_t n = size_t(nq) * k1;
		off_dist = 0;
		off_idx = up(off_dist + n * 4);
		off_label = up(off_idx + n * 4);
		off_count = up(off_label + n * 8);
		off_size = up(off_count + size_t(nq) * 4);
		bytes = up(off_size + 16);
	}
};

// One warp per query: lane s walks shard s's list (already ascending under (dist, internal row)); k rounds of a warp-wide minimum
// under (dist, global row) give the global top-k; one more round looks at the (k+1)-th candidate to flag a straddling tie.
// Map-space distances compare as floats (-0 == +0, like the reference's float compare and the host merge).
__global__ void shard_merge_kernel(const unsigned char* all, uint32_t nshards, uint32_t nq, uint32_t k, uint32_t k1, PayloadLayout lay,
								   float* out_dist, uint64_t* out_gidx, uint64_t* out_label, uint32_t* out_count, uint8_t* need_tie) {
	const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
	const int lane = threadIdx.x & 31;
	if (q >= nq) {
		return;
	}
	// shard bases = prefix sums of the shard sizes (rows are appended shard by shard)
	uint64_t base = 0;
	for (uint32_t s = 0; s < nshards && s < uint32_t(lane); ++s) {
		base += *reinterpret_cast<const uint64_t*>(all + size_t(s) * lay.bytes + lay.off_size);
	}
	const unsigned char* mine = all + size_t(lane) * lay.bytes;
	const bool have = uint32_t(lane) < nshards;
	const uint32_t cnt = have ? min(reinterpret_cast<const uint32_t*>(mine + lay.off_count)[q], k1) : 0u;
	const float* d = reinterpret_cast<const float*>(mine + lay.off_dist) + size_t(q) * k1;
	const uint32_t* ix = reinterpret_cast<const uint32_t*>(mine + lay.off_idx) + size_t(q) * k1;
	const uint64_t* lb = reinterpret_cast<const uint64_t*>(mine + lay.off_label) + size_t(q) * k1;
	uint32_t head = 0, n = 0;
	float last = 0.f;
	bool tie = false;
	for (uint32_t r = 0; r <= k; ++r) {
		float hd = head < cnt ? d[head] : INFINITY;
		uint64_t hg = head < cnt ? base + ix[head] : ~0ull;
		bool valid = head < cnt;
		int owner = lane;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			const float od = __shfl_xor_sync(0xffffffffu, hd, off);
			const uint64_t og = __shfl_xor_sync(0xffffffffu, hg, off);
			const bool ov = __shfl_xor_sync(0xffffffffu, valid, off);
			const int oo = __shfl_xor_sync(0xffffffffu, owner, off);
			const bool better = ov && (!valid || od < hd || (!(hd < od) && og < hg));
			if (better) {
				hd = od;
				hg = og;
				valid = ov;
				owner = oo;
			}
		}
		if (!valid) {
			break;
		}
		if (r == k) {
			tie = !(last < hd);  // the k-th and the (k+1)-th distance are bit-equal (as floats): the reference's tie rule decides
			break;
		}
		if (lane == owner) {
			out_dist[size_t(q) * k + r] = d[head];
			out_gidx[size_t(q) * k + r] = hg;
			out_label[size_t(q) * k + r] = lb[head];
			++head;
		}
		last = hd;
		++n;
	}
	if (lane == 0) {
		out_count[q] = n;
		need_tie[q] = tie ? 1 : 0;
	}
}

}  // namespace

// Ranks that live in ONE process (one host thread per rank; the GPUs may differ or be the same): the exchange goes through pinned host
// memory behind a rendezvous instead of NCCL, which refuses two ranks on one device.  This is what a single reindexer process driving
// several GPUs uses, and what lets a one-GPU box exercise every cross-shard code path.
struct LocalGroup {
	int n = 0;
	std::mutex m;
	std::condition_variable cv;
	int arrived = 0;
	uint64_t generation = 0;
	std::vector<std::vector<unsigned char>> contrib;  // per rank
	std::vector<unsigned char> result;
	template <class F>
	void rendezvous(F&& leader_work) {  // the last rank to arrive runs leader_work, then everybody leaves
		std::unique_lock<std::mutex> lck(m);
		if (++arrived == n) {
			leader_work();
			arrived = 0;
			++generation;
			cv.notify_all();
		} else {
			const uint64_t g = generation;
			cv.wait(lck, [&] { return generation != g; });
		}
	}
};

struct rxgpu_comm {
	ncclComm_t comm = nullptr;
	std::shared_ptr<LocalGroup> local;  // set: the ranks of this communicator are threads of this process
	int nranks = 1, rank = 0, device = 0;
	cudaStream_t stream = nullptr;
	std::mutex mtx;  // one collective call at a time per communicator
	DevBuf<float> d_queries;
	DevBuf<unsigned char> d_send, d_recv;
	DevBuf<float> d_m_dist;
	DevBuf<uint64_t> d_m_gidx, d_m_label;
	DevBuf<uint32_t> d_m_count;
	DevBuf<uint8_t> d_m_tie;
	PinBuf<float> h_m_dist;
	PinBuf<uint64_t> h_m_gidx, h_m_label;
	PinBuf<uint32_t> h_m_count;
	PinBuf<uint8_t> h_m_tie;
	PinBuf<unsigned char> h_tie_recv;
	PinBuf<uint64_t> h_size;
	~rxgpu_comm() {
		cudaSetDevice(device);
		if (comm && nccl().ok) {
			nccl().CommDestroy(comm);
		}
		if (stream) {
			cudaStreamDestroy(stream);
		}
	}
};

// ---- collectives for the other translation units (declared in internal.h): in place on device buffers, on the caller's stream ------
namespace rxgpu {
int commRank(const rxgpu_comm* c) { return c ? c->rank : 0; }
int commSize(const rxgpu_comm* c) { return c ? c->nranks : 1; }
int commDevice(const rxgpu_comm* c) { return c ? c->device : 0; }
std::mutex& commMutex(rxgpu_comm* c) { return c->mtx; }

int commAllReduce(rxgpu_comm* c, void* d_buf, size_t count, CommOp op, cudaStream_t st) {
	if (!c || c->nranks == 1 || count == 0) {
		return 0;
	}
	const size_t esz = op == CommOp::SumU64 ? 8 : 4;
	if (c->local) {
		LocalGroup& g = *c->local;
		std::vector<unsigned char>& mine = g.contrib[size_t(c->rank)];
		mine.resize(count * esz);
		RX_CUDA(cudaMemcpyAsync(mine.data(), d_buf, count * esz, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		g.rendezvous([&] {
			g.result = g.contrib[0];
			for (int r = 1; r < g.n; ++r) {
				for (size_t i = 0; i < count; ++i) {
					if (op == CommOp::SumU64) {
						reinterpret_cast<unsigned long long*>(g.result.data())[i] += reinterpret_cast<const unsigned long long*>(g.contrib[size_t(r)].data())[i];
					} else if (op == CommOp::SumU32) {
						reinterpret_cast<uint32_t*>(g.result.data())[i] += reinterpret_cast<const uint32_t*>(g.contrib[size_t(r)].data())[i];
					} else {
						uint32_t& a = reinterpret_cast<uint32_t*>(g.result.data())[i];
						a = std::max(a, reinterpret_cast<const uint32_t*>(g.contrib[size_t(r)].data())[i]);
					}
				}
			}
		});
		RX_CUDA(cudaMemcpyAsync(d_buf, g.result.data(), count * esz, cudaMemcpyHostToDevice, st));
		RX_CUDA(cudaStreamSynchronize(st));
		g.rendezvous([] {});  // nobody starts the next exchange (which rewrites `result`) before everybody has copied this one
		return 0;
	}
	const ncclDataType_t ty = op == CommOp::SumU64 ? ncclUint64 : ncclUint32;
	RX_NCCL(nccl().AllReduce(d_buf, d_buf, count, ty, op == CommOp::MaxU32 ? ncclMax : ncclSum, c->comm, st));
	return 0;
}

int commAllGather(rxgpu_comm* c, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st) {
	if (bytes == 0) {
		return 0;
	}
	if (!c || c->nranks == 1) {
		RX_CUDA(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, st));
		return 0;
	}
	if (c->local) {
		LocalGroup& g = *c->local;
		std::vector<unsigned char>& mine = g.contrib[size_t(c->rank)];
		mine.resize(bytes);
		RX_CUDA(cudaMemcpyAsync(mine.data(), d_send, bytes, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		g.rendezvous([&] {
			g.result.resize(bytes * size_t(g.n));
			for (int r = 0; r < g.n; ++r) {
				std::memcpy(g.result.data() + bytes * size_t(r), g.contrib[size_t(r)].data(), bytes);
			}
		});
		RX_CUDA(cudaMemcpyAsync(d_recv, g.result.data(), bytes * size_t(g.n), cudaMemcpyHostToDevice, st));
		RX_CUDA(cudaStreamSynchronize(st));
		g.rendezvous([] {});
		return 0;
	}
	RX_NCCL(nccl().AllGather(d_send, d_recv, bytes, ncclChar, c->comm, st));
	return 0;
}
}  // namespace rxgpu

extern "C" {

int rxgpu_comm_unique_id(void* out128) {
	if (!out128) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	if (!nccl().ok) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: libnccl.so.2 could not be loaded (multi-GPU sharding needs NCCL)");
	}
	static_assert(sizeof(ncclUniqueId) == RXGPU_COMM_ID_BYTES, "unique id size");
	ncclUniqueId id;
	RX_NCCL(nccl().GetUniqueId(&id));
	std::memcpy(out128, &id, sizeof(id));
	return 0;
}

int rxgpu_comm_create(rxgpu_comm** out, int nranks, int rank, const void* id128, int device) {
	if (!out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: bad communicator arguments");
	}
	*out = nullptr;
	if (nranks > 32) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: at most 32 shards per communicator (one warp lane per shard in the merge)");
	}
	if (rxgpu_device_count() <= device || device < 0) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: no usable CUDA device (this library has no CPU fallback)");
	}
	RX_CUDA(cudaSetDevice(device));
	auto c = std::make_unique<rxgpu_comm>();
	c->nranks = nranks;
	c->rank = rank;
	c->device = device;
	RX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	if (nranks > 1) {
		if (!nccl().ok) {
			return fail(RXGPU_ERR_SYSTEM, "rxgpu: libnccl.so.2 could not be loaded (multi-GPU sharding needs NCCL)");
		}
		ncclUniqueId id;
		std::memcpy(&id, id128, sizeof(id));
		RX_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank));
	}
	*out = c.release();
	return 0;
}

int rxgpu_comm_create_local(rxgpu_comm** out, int nranks, const int* devices) {
	if (!out || nranks < 1 || nranks > 32) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: bad communicator arguments (1..32 ranks)");
	}
	for (int r = 0; r < nranks; ++r) {
		out[r] = nullptr;
	}
	auto group = std::make_shared<LocalGroup>();
	group->n = nranks;
	group->contrib.resize(size_t(nranks));
	std::vector<std::unique_ptr<rxgpu_comm>> made;
	for (int r = 0; r < nranks; ++r) {
		const int device = devices ? devices[r] : 0;
		if (rxgpu_device_count() <= device || device < 0) {
			return fail(RXGPU_ERR_SYSTEM, "rxgpu: no usable CUDA device (this library has no CPU fallback)");
		}
		RX_CUDA(cudaSetDevice(device));
		auto c = std::make_unique<rxgpu_comm>();
		c->nranks = nranks;
		c->rank = r;
		c->device = device;
		c->local = group;
		RX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		made.push_back(std::move(c));
	}
	for (int r = 0; r < nranks; ++r) {
		out[r] = made[size_t(r)].release();
	}
	return 0;
}

void rxgpu_comm_destroy(rxgpu_comm* c) { delete c; }
int rxgpu_comm_rank(const rxgpu_comm* c) { return c ? c->rank : -1; }
int rxgpu_comm_size(const rxgpu_comm* c) { return c ? c->nranks : 0; }

int rxgpu_merge_shards_device(uint32_t nshards, uint32_t nq, uint32_t k, uint32_t k1, const void* d_payloads, float* d_out_dist,
							  uint64_t* d_out_gidx, uint64_t* d_out_label, uint32_t* d_out_count, uint8_t* d_need_tie, void* stream) {
	if (nshards == 0 || nshards > 32 || k == 0 || k1 < k) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: bad shard merge arguments");
	}
	if (nq == 0) {
		return 0;
	}
	const PayloadLayout lay(nq, k1);
	const unsigned blocks = unsigned((uint64_t(nq) * 32 + 255) / 256);
	shard_merge_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char*>(d_payloads), nshards, nq, k, k1, lay,
																			 d_out_dist, d_out_gidx, d_out_label, d_out_count, d_need_tie);
	RX_CUDA(cudaGetLastError());
	return 0;
}

uint64_t rxgpu_shard_payload_bytes(uint32_t nq, uint32_t k1) { return PayloadLayout(nq, k1).bytes; }

int rxgpu_sharded_search_knn(rxgpu_comm* c, const rxgpu_index* ix, uint32_t nq, const float* queries, int queries_on_device, uint32_t k,
							 float* out_dist, uint64_t* out_label, uint32_t* out_count) {
	if (!c) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null communicator");
	}
	if (int rc = checkIndex(ix)) {
		return rc;
	}
	if (ix->device != c->device) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: the shard lives on another device than its communicator");
	}
	if (c->local && c->nranks > 1) {
		return fail(RXGPU_ERR_LOGIC, "rxgpu: the sharded KNN search exchanges over NCCL (one process per GPU); in-process rank groups serve the ft_fast merge");
	}
	if (nq && (!queries || !out_count || (k && (!out_dist || !out_label)))) {
		return fail(RXGPU_ERR_PARAMS, "rxgpu: null argument");
	}
	g_stats = rxgpu_search_stats{};
	if (nq == 0) {
		return 0;
	}
	if (k == 0) {
		std::memset(out_count, 0, size_t(nq) * 4);
		return 0;
	}
	const uint32_t k1 = k + 1;  // the same on every rank (shard sizes differ): one extra row exposes a tie at the k-th place
	if (k1 > 65536u) {  // kMaxSearchK1 of the scan
		return fail(RXGPU_ERR_PARAMS, "rxgpu: k must be in [1, 65535]");
	}
	std::lock_guard<std::mutex> lck(c->mtx);
	try {
		cudaStream_t st = c->stream;
		WsLease lease(ix);
		Workspace& ws = *lease.ws;
		const PayloadLayout lay(nq, k1);
		const uint32_t R = uint32_t(c->nranks);
		const float* d_q = queries;
		if (!queries_on_device) {
			RX_CUDA(c->d_queries.ensure(size_t(nq) * ix->dim));
			RX_CUDA(cudaMemcpyAsync(c->d_queries.p, queries, size_t(nq) * ix->dim * 4, cudaMemcpyHostToDevice, st));
			d_q = c->d_queries.p;
		}
		RX_CUDA(c->d_send.ensure(lay.bytes));
		RX_CUDA(c->d_recv.ensure(lay.bytes * R));
		RX_CUDA(c->h_size.ensure(2));
		unsigned char* snd = R > 1 ? c->d_send.p : c->d_recv.p;  // a single shard merges its own payload in place
		float* s_dist = reinterpret_cast<float*>(snd + lay.off_dist);
		uint32_t* s_idx = reinterpret_cast<uint32_t*>(snd + lay.off_idx);
		uint64_t* s_label = reinterpret_cast<uint64_t*>(snd + lay.off_label);
		uint32_t* s_count = reinterpret_cast<uint32_t*>(snd + lay.off_count);
		c->h_size.p[0] = ix->size;
		c->h_size.p[1] = 0;
		RX_CUDA(cudaMemcpyAsync(snd + lay.off_size, c->h_size.p, 16, cudaMemcpyHostToDevice, st));
		// ---- 1. this shard's top-(k+1) under (dist, internal row), straight into the send buffer
		if (ix->size == 0) {
			RX_CUDA(cudaMemsetAsync(s_count, 0, size_t(nq) * 4, st));
			ws.tc_lists_valid = false;
		} else if (int rc = scanTopK(ix, ws, st, d_q, nq, k1, kModeTopK, 0.f, s_dist, s_idx, s_label, s_count)) {
			return rc;
		}
		const rxgpu_search_stats scanStats = g_stats;  // what the roofline figure describes: the shard scan, not the rare tie pass
		// ---- 2. one all-gather, 3. device merge
		if (R > 1) {
			RX_NCCL(nccl().AllGather(c->d_send.p, c->d_recv.p, lay.bytes, ncclChar, c->comm, st));
		}
		const size_t on = size_t(nq) * k;
		RX_CUDA(c->d_m_dist.ensure(on));
		RX_CUDA(c->d_m_gidx.ensure(on));
		RX_CUDA(c->d_m_label.ensure(on));
		RX_CUDA(c->d_m_count.ensure(nq));
		RX_CUDA(c->d_m_tie.ensure(nq));
		RX_CUDA(c->h_m_dist.ensure(on));
		RX_CUDA(c->h_m_gidx.ensure(on));
		RX_CUDA(c->h_m_label.ensure(on));
		RX_CUDA(c->h_m_count.ensure(nq));
		RX_CUDA(c->h_m_tie.ensure(nq));
		if (int rc = rxgpu_merge_shards_device(R, nq, k, k1, c->d_recv.p, c->d_m_dist.p, c->d_m_gidx.p, c->d_m_label.p, c->d_m_count.p,
											   c->d_m_tie.p, st)) {
			return rc;
		}
		RX_CUDA(cudaMemcpyAsync(c->h_m_dist.p, c->d_m_dist.p, on * 4, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(c->h_m_gidx.p, c->d_m_gidx.p, on * 8, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(c->h_m_label.p, c->d_m_label.p, on * 8, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(c->h_m_count.p, c->d_m_count.p, size_t(nq) * 4, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaMemcpyAsync(c->h_m_tie.p, c->d_m_tie.p, nq, cudaMemcpyDeviceToHost, st));
		RX_CUDA(cudaStreamSynchronize(st));
		collectProfile();
		// ---- 4. host: runs of bit-equal distances ordered by label (the drain order of the reference's heap), output
		std::vector<uint32_t> tieQ;
		std::vector<float> tieD;
		std::vector<Hit> top;
		for (uint32_t q = 0; q < nq; ++q) {
			const uint32_t n = c->h_m_count.p[q];
			top.resize(n);
			bool anyEqual = false;
			for (uint32_t j = 0; j < n; ++j) {
				top[j] = Hit{c->h_m_dist.p[size_t(q) * k + j], c->h_m_gidx.p[size_t(q) * k + j], c->h_m_label.p[size_t(q) * k + j]};
				anyEqual |= j && !(top[j - 1].dist < top[j].dist);
			}
			if (anyEqual) {
				orderTiesByLabel(top);
			}
			for (uint32_t j = 0; j < n; ++j) {
				out_dist[size_t(q) * k + j] = top[j].dist;
				out_label[size_t(q) * k + j] = top[j].label;
			}
			out_count[q] = n;
			if (c->h_m_tie.p[q] && n == k) {
				tieQ.push_back(q);
				tieD.push_back(c->h_m_dist.p[size_t(q) * k + k - 1]);
			}
		}
		// ---- 5. rare: a tie straddles the global k-th place of some queries (the same set on every rank) -> replay the reference's rule
		if (!tieQ.empty()) {
			const uint32_t nt = uint32_t(tieQ.size());
			const PayloadLayout tlay(nt, k);
			RX_CUDA(c->d_send.ensure(tlay.bytes));
			RX_CUDA(c->d_recv.ensure(tlay.bytes * R));
			unsigned char* tsnd = R > 1 ? c->d_send.p : c->d_recv.p;
			RX_CUDA(cudaMemcpyAsync(tsnd + tlay.off_size, c->h_size.p, 16, cudaMemcpyHostToDevice, st));
			if (ix->size == 0) {
				RX_CUDA(cudaMemsetAsync(tsnd + tlay.off_count, 0, size_t(nt) * 4, st));
			} else if (int rc = tieRowsAfterScan(ix, ws, st, d_q, nt, tieQ.data(), tieD.data(), k, reinterpret_cast<float*>(tsnd + tlay.off_dist),
												 reinterpret_cast<uint32_t*>(tsnd + tlay.off_idx), reinterpret_cast<uint64_t*>(tsnd + tlay.off_label),
												 reinterpret_cast<uint32_t*>(tsnd + tlay.off_count))) {
				return rc;
			}
			if (R > 1) {
				RX_NCCL(nccl().AllGather(c->d_send.p, c->d_recv.p, tlay.bytes, ncclChar, c->comm, st));
			}
			RX_CUDA(c->h_tie_recv.ensure(tlay.bytes * R));
			RX_CUDA(cudaMemcpyAsync(c->h_tie_recv.p, c->d_recv.p, tlay.bytes * R, cudaMemcpyDeviceToHost, st));
			RX_CUDA(cudaStreamSynchronize(st));
			std::vector<uint64_t> base(R, 0);
			for (uint32_t s = 1; s < R; ++s) {
				base[s] = base[s - 1] + *reinterpret_cast<const uint64_t*>(c->h_tie_recv.p + size_t(s - 1) * tlay.bytes + tlay.off_size);
			}
			std::vector<Hit> lower, first;
			for (uint32_t t = 0; t < nt; ++t) {
				const uint32_t q = tieQ[t];
				const float dstar = tieD[t];
				lower.clear();
				first.clear();
				for (uint32_t j = 0; j < k && c->h_m_dist.p[size_t(q) * k + j] < dstar; ++j) {
					lower.push_back(Hit{c->h_m_dist.p[size_t(q) * k + j], c->h_m_gidx.p[size_t(q) * k + j], c->h_m_label.p[size_t(q) * k + j]});
				}
				for (uint32_t s = 0; s < R; ++s) {  // shards in rank order = global internal order; each list is already in internal order
					const unsigned char* p = c->h_tie_recv.p + size_t(s) * tlay.bytes;
					const uint32_t cnt = std::min(reinterpret_cast<const uint32_t*>(p + tlay.off_count)[t], k);
					for (uint32_t j = 0; j < cnt && first.size() < k; ++j) {
						first.push_back(Hit{reinterpret_cast<const float*>(p + tlay.off_dist)[size_t(t) * k + j],
											base[s] + reinterpret_cast<const uint32_t*>(p + tlay.off_idx)[size_t(t) * k + j],
											reinterpret_cast<const uint64_t*>(p + tlay.off_label)[size_t(t) * k + j]});
					}
				}
				const std::vector<Hit> res = tieReplay(k, dstar, lower, first);
				for (size_t j = 0; j < res.size(); ++j) {
					out_dist[size_t(q) * k + j] = res[j].dist;
					out_label[size_t(q) * k + j] = res[j].label;
				}
				out_count[q] = uint32_t(res.size());
			}
		}
		// report the shard scan's figures (kernel, tile, launches timed) plus what the tie pass added
		const rxgpu_search_stats after = g_stats;
		g_stats.query_tile = scanStats.query_tile;
		g_stats.tc_used = scanStats.tc_used;
		g_stats.tc_cluster = scanStats.tc_cluster;
		g_stats.tc_kernel = scanStats.tc_kernel;
		g_stats.tc_candidates = scanStats.tc_candidates;
		g_stats.tc_fallbacks = scanStats.tc_fallbacks;
		g_stats.launches = after.launches + 1 + (tieQ.empty() ? 0 : 0);
		g_stats.tie_replays = uint32_t(tieQ.size());
	} catch (const std::bad_alloc&) {
		return fail(RXGPU_ERR_SYSTEM, "rxgpu: out of host memory");
	}
	return 0;
}

}  // extern "C"
