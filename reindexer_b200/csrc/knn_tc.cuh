// Tensor-core candidate filter for large query batches (sm_100a: TMA + tcgen05.mma + TMEM), exact results.
//
// knn_scan_warp is HBM-bound only while <= ~16 queries share a pass; a batch of 1024 queries is FMA-bound there.  This kernel
// computes APPROXIMATE scores for a block of NQ queries against every row with bf16 operands on the 5th-gen tensor cores and
// keeps, per query, only the rows that can still be among the k best under a CERTIFIED error bound; the survivors (a few hundred
// per query) are then re-ranked with the exact fp32 routine of knn_scan_warp, so the final result is identical to the exact scan.
//
//   error bound   |q~.v~ - q.v| <= c * ||q|| * ||v||,  c = 2^-8 + 2^-18 (two bf16 roundings, unit roundoff 2^-9 each)
//                                                        + dim * 2^-23 (fp32 accumulation in the MMA); c = 0.0042 leaves 5% slack
//                                                        at 768 dims and 1% at 2048 dims, the largest dimension the filter accepts
//   lower bound   lb = d~ - e, upper bound ub = d~ + e in map space (smaller is better)
//   threshold     tau_q = k1-th smallest ub over all DISTINCT rows seen so far by any CTA (one small list per query in HBM,
//                 updated under a per-query lock -- only O(k log n) successful inserts per query over a whole pass) => a valid
//                 upper bound of the final k1-th best TRUE distance; a row is a candidate iff lb <= tau_q.  tau starts from an
//                 exact scan of the first rows (tc_init_tau) and only decreases.
//
// Roles (192 threads, 1 CTA per SM, persistent over 128-row tiles):
//   warp 0    producer: bf16 shadow rows, 128 x 64 tiles (16 KB) through a 4-stage mbarrier ring.  The shadow is stored TILED and
//             PRE-SWIZZLED in HBM ([tile of 64 rows][K chunk][64 x 128 B in the SWIZZLE_128B pattern]) so a stage is two
//             contiguous 8 KB cp.async.bulk copies (row-major fp32 stays the source of truth; the shadow is private, derived)
//   warp 1    allocates TMEM (512 columns), issues tcgen05.mma (M=128 rows, N=NQ queries, K=16) from shared-memory descriptors;
//             the query block (NQ x dim bf16) is loaded once by TMA and stays resident in shared memory
//   warps 2-5 epilogue: tcgen05.ld the 128 x NQ fp32 accumulators (double buffered in TMEM so the next tile's MMAs overlap),
//             apply the metric, test against tau, append candidates to per-query lists in HBM, tighten tau
// Bound: HBM (bf16 shadow: n*dim*2 bytes per pass of NQ queries); MMA time per tile is ~4x below the tile's HBM time.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "knn_scan.cuh"

namespace rxgpu {

constexpr int kTcThreads = 192;
constexpr int kTcTileRows = 128;     // UMMA M
constexpr int kTcChunkK = 64;        // bf16 elements per 128-byte swizzle row
constexpr int kTcStages = 4;
constexpr int kTcStageBytes = kTcTileRows * kTcChunkK * 2;  // 16 KB
constexpr uint32_t kTcMaxK1 = 128;  // k + 1 <= 128: the bound list is scanned linearly under the per-query lock and ~25-35 (k + 1) candidates
                                    // per query must fit the 4096-entry lists (k = 10: 330; k = 63: 1600; an overflowing query takes the exact scan)
constexpr uint32_t kTcQueueCap = 512;
constexpr float kTcErrCoef = 0.0042f;  // see header comment

struct TcArgs {
	const unsigned char* shadow;  // bf16 shadow, [tile of 64 rows][K chunk][64 rows x 128 B, SWIZZLE_128B pattern pre-applied]
	const float* vnorm;        // [n] ||row||_2 (fp32)
	const float* vinv;         // [n] 1/||row|| (Cosine) or nullptr
	const float* qnorm;        // [nq_total] ||q||_2
	unsigned int* tau;         // [nq_total] ordered-uint of the current threshold (map space), shared by all CTAs
	float* ub_list;            // [nq_total][kTcMaxK1] the k1 smallest upper bounds over ALL rows seen by any CTA (guarded by ub_lock)
	unsigned int* ub_lock;     // [nq_total]
	uint32_t init_rows;        // rows [0, init_rows) are already represented in ub_list by tc_init_tau (never insert them twice)
	uint32_t* cand_rows;       // [nq_total][cand_cap]
	unsigned int* cand_count;  // [nq_total]
	uint32_t cand_cap;
	uint32_t n;                // rows
	uint32_t kchunks;          // padded dim / 64
	uint32_t nq_block;         // UMMA N (multiple of 32, <= 256): queries resident in this launch
	uint32_t nq_total;         // queries in the whole batch
	uint32_t q0;               // first query of this launch; CTA rank r of a cluster owns queries [q0 + r*nq_block, +nq_block)
	uint32_t k1;
	int metric;                // kL2 / kIP / kCos
};

__host__ __device__ inline size_t tc_smem_bytes(uint32_t nq_block, uint32_t kchunks) {
	return 1024 /*align slack*/ + size_t(nq_block) * kchunks * 128 + size_t(kTcStages) * kTcStageBytes + 256 /*barriers*/ +
		   size_t(nq_block) * (4 + 4 + 8) + size_t(kTcQueueCap) * 8 + 64;
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t x, int32_t y) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
					 smem_u32(dst)),
				 "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
				 : "memory");
}
// 1-D bulk copies (TMA engine, no tensor map): the shadow is stored pre-swizzled, so a stage is a verbatim contiguous copy
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
				 "r"(bytes), "r"(smem_u32(bar))
				 : "memory");
}
// warm the L2 with a block this SM will copy a few tiles from now: the later bulk copy then pays the L2 latency, not the HBM one,
// which a shared-memory ring of only one or two tiles cannot hide
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
					 smem_u32(dst)),
				 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
				 : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t x, int32_t y, uint16_t mask) {
	asm volatile(
		"cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::
			"r"(smem_u32(dst)),
		"l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask)
		: "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
				 "h"(mask)
				 : "memory");
}
// true in exactly one lane of a converged warp (elect.sync): the single-thread tcgen05 / TMA instructions are issued under this
// predicate from warp-uniform code, so their operands stay in uniform registers
__device__ __forceinline__ bool elect_one_sync() {
	uint32_t p;
	asm volatile(
		"{\n"
		".reg .pred P;\n"
		"elect.sync _|P, 0xffffffff;\n"
		"selp.u32 %0, 1, 0, P;\n"
		"}\n"
		: "=r"(p));
	return p != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), one swizzle atom along K (LBO unused)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
	uint64_t d = 0;
	d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
	d |= uint64_t(0) << 16;                      // leading byte offset (unused for swizzled K-major)
	d |= uint64_t(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
	d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
	d |= uint64_t(2) << 61;                      // layout type: SWIZZLE_128B
	return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, dense
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
	return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
		  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
		  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
		  "=r"(r[31])
		: "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// The candidate test  lb = d~ - e <= tau  rewritten as ONE fused multiply-add and one compare on the raw accumulator s = q~.v~:
//   IP      d = -s,               e = qe*vn            <=>  s >= -tau - qe*vn                       P = -qe        R = -tau      w = 0
//   Cosine  d = -s/vn,            e = qe               <=>  s >= (-tau - qe) * vn                   P = -tau - qe  R = 0         w = 0
//   L2      d = qn2 + vn2 - 2s,   e = 2qe*vn + eps*(qn2+vn2)
//                                                      <=>  s - (1-eps)/2*vn2 >= -qe*vn + ((1-eps)*qn2 - tau)/2   (w = (1-eps)/2*vn2 per row)
// The error coefficient carries 5% slack, which also covers the one-ulp differences of these rearrangements.
constexpr float kTcL2Eps = 1e-5f;
__device__ __forceinline__ float2 tc_make_pr(int metric, float tau, float qe) {
	if (metric == kIP) {
		return make_float2(-qe, -tau);
	}
	if (metric == kCos) {
		return make_float2(-tau - qe, 0.f);
	}
	const float qn = qe * (1.f / kTcErrCoef);
	return make_float2(-qe, 0.5f * ((1.f - kTcL2Eps) * qn * qn - tau));
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
		  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
		  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
		  "=r"(r[31])
		: "r"(taddr));
}

// ---- the filter kernel -----------------------------------------------------------------------------------------------------------
// kCluster == 2: a cluster of two CTAs walks the same row tiles with DIFFERENT query blocks; each CTA fetches half of every
// stage (64 rows) and TMA-multicasts it into both CTAs' shared memory, so one pass over the shadow serves 2 x nq_block queries.
template <int kCluster>
__global__ void __launch_bounds__(kTcThreads, 1)
	knn_tc_filter(const __grid_constant__ CUtensorMap map_queries, const TcArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset arithmetic keeps the shared window
	const uint32_t qchunk_bytes = a.nq_block * 128;                   // one K-chunk of the query block: nq_block rows x 128 B
	unsigned char* s_q = base;                                        // [kchunks][nq_block][128 B], swizzled by TMA
	unsigned char* s_rows = s_q + size_t(a.kchunks) * qchunk_bytes;   // [stages][128][128 B]   (1024-aligned: nq_block % 8 == 0)
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_rows + size_t(kTcStages) * kTcStageBytes);
	uint64_t* full_bar = bars;                   // [stages] TMA -> MMA
	uint64_t* empty_bar = bars + kTcStages;      // [stages] MMA -> TMA
	uint64_t* q_bar = bars + 2 * kTcStages;      // queries resident
	uint64_t* acc_full = q_bar + 1;              // [2] MMA -> epilogue
	uint64_t* acc_empty = acc_full + 2;          // [2] epilogue -> MMA
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 2);
	float* s_thr = reinterpret_cast<float*>(bars + 32);                // [nq_block] current tau (map space)
	float* s_qe = s_thr + a.nq_block;                                  // [nq_block] c * ||q||
	float2* s_pr = reinterpret_cast<float2*>(s_qe + a.nq_block);      // [nq_block] (P, R): candidate iff s - w_row >= fma(P, ||v||, R)
	uint2* s_queue = reinterpret_cast<uint2*>(s_pr + a.nq_block);     // (query, ub bits)
	uint32_t* s_qcount = reinterpret_cast<uint32_t*>(s_queue + kTcQueueCap);

	const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
	const uint32_t ntiles = (a.n + kTcTileRows - 1) / kTcTileRows;
	const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
	const uint32_t cid = blockIdx.x / kCluster, ncl = gridDim.x / kCluster;  // tile walkers
	const uint32_t q0 = a.q0 + crank * a.nq_block;
	const uint32_t nq_valid = q0 < a.nq_total ? min(a.nq_block, a.nq_total - q0) : 0u;

	if (threadIdx.x == 0) {
		for (int s = 0; s < kTcStages; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], kCluster);  // every consumer CTA of the stage must release it
		}
		mbar_init(q_bar, 1);
		for (int s = 0; s < 2; ++s) {
			mbar_init(&acc_full[s], 1);
			mbar_init(&acc_empty[s], 4);  // one arrive per epilogue warp
		}
		*s_qcount = 0;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (uint32_t i = threadIdx.x; i < a.nq_block; i += blockDim.x) {
		const bool valid = i < nq_valid;
		s_thr[i] = valid ? ord_float(a.tau[q0 + i]) : -INFINITY;
		s_qe[i] = valid ? kTcErrCoef * a.qnorm[q0 + i] : 0.f;
		s_pr[i] = valid ? tc_make_pr(a.metric, s_thr[i], s_qe[i]) : make_float2(0.f, INFINITY);  // padding queries never match
	}
	if (warp == 1) {  // TMEM: 512 columns = 2 accumulator buffers of up to 256 fp32 columns
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if constexpr (kCluster > 1) {
		cluster_sync_all();  // the peer's barriers exist before anything of ours can signal them
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = *s_tmem;

	if (warp == 0) {
		// ===== TMA producer: the whole warp walks the loop, one elected lane issues (operands stay in uniform registers) =====
		if (elect_one_sync()) {
			mbar_expect_tx(q_bar, a.kchunks * qchunk_bytes);
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				tma_load_2d(s_q + size_t(kc) * qchunk_bytes, &map_queries, q_bar, int32_t(kc * kTcChunkK), int32_t(q0));
			}
		}
		__syncwarp();
		uint32_t stage = 0, phase = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl) {
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				mbar_wait(&empty_bar[stage], phase ^ 1);
				// a 128-row stage = the 64-row shadow sub-tiles 2t and 2t+1 of this K chunk, 8 KB each, contiguous in HBM
				const unsigned char* src = a.shadow + (size_t(2 * t) * a.kchunks + kc) * 8192u;
				unsigned char* dst = s_rows + size_t(stage) * kTcStageBytes;
				if (elect_one_sync()) {
					mbar_expect_tx(&full_bar[stage], kTcStageBytes);
					if constexpr (kCluster > 1) {  // my half of the stage, delivered to both CTAs
						bulk_load_mc(dst + crank * 8192u, src + size_t(crank) * a.kchunks * 8192u, 8192u, &full_bar[stage],
									 uint16_t((1u << kCluster) - 1u));
					} else {
						bulk_load(dst, src, 8192u, &full_bar[stage]);
						bulk_load(dst + 8192u, src + size_t(a.kchunks) * 8192u, 8192u, &full_bar[stage]);
					}
				}
				__syncwarp();
				if (++stage == kTcStages) {
					stage = 0;
					phase ^= 1;
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer (same structure: convergent loop, elected issue) =====
		const uint32_t idesc = umma_idesc_bf16(kTcTileRows, a.nq_block);
		mbar_wait(q_bar, 0);
		uint32_t stage = 0, phase = 0, it = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
			mbar_wait(&acc_empty[acc], acc_phase ^ 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const uint32_t tmem_d = tmem_base + acc * 256;
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				mbar_wait(&full_bar[stage], phase);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t a_addr = smem_u32(s_rows + size_t(stage) * kTcStageBytes);
				const uint32_t b_addr = smem_u32(s_q + size_t(kc) * qchunk_bytes);
				if (elect_one_sync()) {
#pragma unroll
					for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
						umma_bf16(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kc | k) != 0);
					}
					if constexpr (kCluster > 1) {  // the stage is rewritten by BOTH producers: release it in both CTAs
						umma_commit_mc(&empty_bar[stage], uint16_t((1u << kCluster) - 1u));
					} else {
						umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
					}
					if (kc + 1 == a.kchunks) {
						umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
					}
				}
				__syncwarp();
				if (++stage == kTcStages) {
					stage = 0;
					phase ^= 1;
				}
			}
		}
	} else {
		// ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
		const uint32_t quad = warp & 3;
		const uint32_t row_in_tile = quad * 32 + lane;
		uint32_t it = 0;
		unsigned int tau_ahead0 = float_ord(INFINITY), tau_ahead1 = float_ord(INFINITY);
		float vn_ahead = 0.f, vinv_ahead = 1.f;
		{
			const uint32_t frow = cid * kTcTileRows + row_in_tile;
			if (cid < ntiles && frow < a.n) {
				vn_ahead = a.vnorm[frow];
				vinv_ahead = a.vinv ? a.vinv[frow] : 1.f;
			}
		}
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
			const uint32_t row = t * kTcTileRows + row_in_tile;
			const bool row_ok = row < a.n;
			// refresh tau from the other CTAs (thread i of the epilogue group owns queries i and i + 128).  The global loads were
			// issued during the PREVIOUS tile, so their latency is hidden; the ones for the next tile are issued now.
			{
				const uint32_t i0 = threadIdx.x - 64, i1 = i0 + 128;
				if (i0 < nq_valid) {
					s_thr[i0] = fminf(s_thr[i0], ord_float(tau_ahead0));
					s_pr[i0] = tc_make_pr(a.metric, s_thr[i0], s_qe[i0]);
					tau_ahead0 = a.tau[q0 + i0];
				}
				if (i1 < nq_valid) {
					s_thr[i1] = fminf(s_thr[i1], ord_float(tau_ahead1));
					s_pr[i1] = tc_make_pr(a.metric, s_thr[i1], s_qe[i1]);
					tau_ahead1 = a.tau[q0 + i1];
				}
			}
			const float vn = vn_ahead, vinv = vinv_ahead;
			const float w_row = a.metric == kL2 ? 0.5f * (1.f - kTcL2Eps) * vn * vn : 0.f;
			const float vn_t = fmaxf(vn, 1e-30f);  // keeps -inf * ||v|| = -inf (tau still +inf) for all-zero rows
			{
				const uint32_t nrow = (t + ncl) * kTcTileRows + row_in_tile;
				const bool nok = t + ncl < ntiles && nrow < a.n;
				vn_ahead = nok ? a.vnorm[nrow] : 0.f;
				vinv_ahead = (nok && a.vinv) ? a.vinv[nrow] : 1.f;
			}
			asm volatile("bar.sync 1, 128;" ::: "memory");
			mbar_wait(&acc_full[acc], acc_phase);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			for (uint32_t c0 = 0; c0 < a.nq_block; c0 += 32) {
				uint32_t v[32];
				tmem_ld32(tmem_base + acc * 256 + c0 + ((quad * 32) << 16), v);
				// hot path: one broadcast LDS.64, one FFMA, one compare per element; hits are collected in a bit mask
				uint32_t hits = 0;
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const float2 pr = s_pr[c0 + j];
					const float sv = __uint_as_float(v[j]) - w_row;
					hits |= uint32_t(sv >= fmaf(pr.x, vn_t, pr.y)) << j;
				}
				hits = row_ok ? hits : 0u;
				const unsigned any_hits = __reduce_or_sync(0xffffffffu, hits);
				if (any_hits) {  // rare path: exact bounds, candidate append, threshold tightening (static indices into v[])
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						if (!(any_hits & (1u << j)) || !(hits & (1u << j))) {
							continue;
						}
						const uint32_t q = c0 + j;
						const float s = __uint_as_float(v[j]);
						float d, e;
						if (a.metric == kL2) {
							const float qn = s_qe[q] * (1.f / kTcErrCoef);
							d = fmaf(-2.f, s, fmaf(qn, qn, vn * vn));
							e = 2.f * s_qe[q] * vn + kTcL2Eps * (qn * qn + vn * vn);
						} else if (a.metric == kCos) {
							d = -s * vinv;
							e = s_qe[q] * vn * vinv;
						} else {
							d = -s;
							e = s_qe[q] * vn;
						}
						const unsigned pos = atomicAdd(&a.cand_count[q0 + q], 1u);
						if (pos < a.cand_cap) {
							a.cand_rows[size_t(q0 + q) * a.cand_cap + pos] = row;
						}
						const float ub = d + e;
						if (ub < s_thr[q] && row >= a.init_rows) {
							const uint32_t slot = atomicAdd(s_qcount, 1u);
							if (slot < kTcQueueCap) {
								s_queue[slot] = make_uint2(q, __float_as_uint(ub));
							}
						}
					}
				}
			}
			// accumulator drained: hand the TMEM buffer back to the MMA warp
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(&acc_empty[acc]);
			}
			asm volatile("bar.sync 1, 128;" ::: "memory");
			// tighten tau: one thread folds the queued upper bounds into the per-query global lists (rare after the first tiles)
			if (threadIdx.x == 64) {
				const uint32_t cnt = min(*s_qcount, kTcQueueCap);
				for (uint32_t i = 0; i < cnt; ++i) {
					const uint32_t q = s_queue[i].x;
					const float ub = __uint_as_float(s_queue[i].y);
					const uint32_t gq = q0 + q;
					if (!(ub < ord_float(*reinterpret_cast<volatile unsigned int*>(&a.tau[gq])))) {
						continue;
					}
					while (atomicCAS(&a.ub_lock[gq], 0u, 1u) != 0u) {
					}
					__threadfence();
					volatile float* list = a.ub_list + size_t(gq) * kTcMaxK1;
					uint32_t mi = 0;
					float mx = list[0];
					for (uint32_t j = 1; j < a.k1; ++j) {
						const float v = list[j];
						if (v > mx) {
							mx = v;
							mi = j;
						}
					}
					if (ub < mx) {
						list[mi] = ub;
						float nmx = list[0];
						for (uint32_t j = 1; j < a.k1; ++j) {
							nmx = fmaxf(nmx, list[j]);
						}
						atomicMin(&a.tau[gq], float_ord(nmx));
						s_thr[q] = fminf(s_thr[q], nmx);
						s_pr[q] = tc_make_pr(a.metric, s_thr[q], s_qe[q]);
					}
					__threadfence();
					atomicExch(&a.ub_lock[gq], 0u);
				}
				*s_qcount = 0;
			}
			asm volatile("bar.sync 1, 128;" ::: "memory");
		}
	}
	__syncthreads();
	if constexpr (kCluster > 1) {
		cluster_sync_all();  // nobody leaves while the peer may still multicast into this CTA or signal its barriers
	}
	if (warp == 1) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
	}
}

// ---- exact re-rank of the candidates ----------------------------------------------------------------------------------------------
// One CTA per query: its 8 warps stream the candidate rows of that query (gathered 128-bit coalesced loads), compute the exact fp32
// distance with the SAME per-row arithmetic sequence as knn_scan_warp (so distances are bit-identical to the exact scan), keep the
// best k1 keys per warp, merge in the CTA and write one ascending list [k1] per query.
template <bool kIsL2>
__global__ void __launch_bounds__(kScanThreads) knn_rerank(const float* rows, uint32_t pitch, uint32_t dim, const float* norm_coefs,
															const float* queries, const uint32_t* cand_rows, const unsigned int* cand_count,
															uint32_t cand_cap, uint32_t k1, uint64_t* lists /* [gridDim.x][k1] */,
															const uint32_t* qsel = nullptr, const float* tie_bound = nullptr) {
	// qsel: CTA b serves query qsel[b] (default: query b).  tie_bound != nullptr = tie mode (kModeTieRows): among the candidates
	// with dist <= tie_bound[b], the first k1 in internal row order (key = row << 32 | ord(dist)) -- every row at or below the k-th
	// distance is a candidate, so this replaces a second scan of the whole shard when the reference's tie rule must be replayed.
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t q = qsel ? qsel[blockIdx.x] : blockIdx.x;
	const bool tie = tie_bound != nullptr;
	const float bound = tie ? tie_bound[blockIdx.x] : 0.f;
	const uint32_t nch = (dim + 127u) / 128u, dp4 = nch * 32u, pitch4 = pitch >> 2;
	const uint32_t m = k1 + kCandBuf;
	float4* sq4 = reinterpret_cast<float4*>(smem_raw);
	uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + size_t(dp4) * 16);  // [8 warps][m]
	{
		float* sq = reinterpret_cast<float*>(sq4);
		for (uint32_t i = threadIdx.x; i < dp4 * 4; i += blockDim.x) {
			sq[i] = i < dim ? queries[size_t(q) * dim + i] : 0.f;
		}
		for (uint32_t i = threadIdx.x; i < kScanWarps * m; i += blockDim.x) {
			skeys[i] = kKeyNone;
		}
	}
	__syncthreads();
	uint64_t* wkeys = skeys + size_t(warp) * m;
	uint64_t thr = kKeyNone;
	uint32_t cnt = 0;
	const uint32_t ncand = min(cand_count[q], cand_cap);
	const float4* rows4 = reinterpret_cast<const float4*>(rows);
	const uint32_t* my = cand_rows + size_t(q) * cand_cap;
	for (uint32_t i = warp; i < ncand; i += kScanWarps) {
		const uint32_t row = my[i];
		float s = 0.f;
		for (uint32_t c = 0; c < nch; ++c) {
			const uint32_t f4 = c * 32u + lane;
			const float4 db = f4 < pitch4 ? ldg_stream(rows4 + size_t(row) * pitch4 + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
			const float4 qv = sq4[f4];
			if constexpr (kIsL2) {
				float d;
				d = qv.x - db.x;
				s = fmaf(d, d, s);
				d = qv.y - db.y;
				s = fmaf(d, d, s);
				d = qv.z - db.z;
				s = fmaf(d, d, s);
				d = qv.w - db.w;
				s = fmaf(d, d, s);
			} else {
				s = fmaf(qv.x, db.x, s);
				s = fmaf(qv.y, db.y, s);
				s = fmaf(qv.z, db.z, s);
				s = fmaf(qv.w, db.w, s);
			}
		}
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			s += __shfl_xor_sync(0xffffffffu, s, off);
		}
		float dist = kIsL2 ? s : -s;
		if (!kIsL2 && norm_coefs != nullptr) {
			dist *= norm_coefs[row];
		}
		const uint64_t key = !tie ? make_key(dist, row) : (dist <= bound ? ((uint64_t(row) << 32) | float_ord(dist)) : kKeyNone);
		if (key < thr) {  // warp-uniform
			if (lane == 0) {
				wkeys[k1 + cnt] = key;
			}
			++cnt;
			__syncwarp();
			if (cnt == kCandBuf) {
				warp_select(wkeys, k1 + cnt, k1, lane);
				thr = wkeys[k1 - 1];
				cnt = 0;
			}
		}
	}
	if (cnt) {
		warp_select(wkeys, k1 + cnt, k1, lane);
	}
	__syncthreads();
	if (warp == 0) {  // CTA merge: strictly increasing selection over the 8 warp lists (keys are unique)
		uint64_t last = 0;
		bool first = true;
		for (uint32_t r = 0; r < k1; ++r) {
			uint64_t best = kKeyNone;
			for (uint32_t i = lane; i < kScanWarps * k1; i += 32) {
				const uint32_t w = i / k1, j = i - w * k1;
				const uint64_t kx = skeys[size_t(w) * m + j];
				if ((first || kx > last) && kx < best) {
					best = kx;
				}
			}
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) {
				const uint64_t ok = __shfl_xor_sync(0xffffffffu, best, off);
				best = ok < best ? ok : best;
			}
			if (lane == 0) {
				lists[size_t(blockIdx.x) * k1 + r] = best;
			}
			last = best;
			first = false;
		}
	}
}

// per-row constants of the filter epilogue, one float2 per row: (max(||v||, tiny), w) with w = the row's share of the L2 expansion
// (0 for IP / Cosine); rows beyond n are zero.  A 64-row tile's pairs are one contiguous 512-byte block (one cp.async.bulk).
__global__ void tc_make_vw(const float* vnorm, uint32_t n, uint32_t padded, int metric, float2* vw) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < padded) {
		const float vn = i < n ? vnorm[i] : 0.f;
		vw[i] = make_float2(fmaxf(vn, 1e-30f), metric == kL2 ? 0.5f * (1.f - kTcL2Eps) * vn * vn : 0.f);
	}
}

// ---- helpers: bf16 shadow, norms, query preparation, threshold init ---------------------------------------------------------------
// rows fp32 [n][pitch] -> bf16 shadow + ||row||_2.  Shadow layout: [tile of 64 rows][K chunk of 64][64 rows x 128 bytes], and inside
// every 8 KB block the 16-byte units of row r are XOR-permuted with (r % 8) -- the SWIZZLE_128B pattern tcgen05.mma expects in
// shared memory -- so that a plain contiguous cp.async.bulk brings a ready-to-multiply operand tile.
__global__ void tc_convert_rows(const float* rows, uint32_t pitch, uint32_t dim, uint32_t row_begin, uint32_t row_end, __nv_bfloat16* shadow,
								uint32_t kchunks, float* vnorm) {
	const uint32_t row = row_begin + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
	const int lane = threadIdx.x & 31;
	if (row >= row_end) {
		return;
	}
	const float* p = rows + size_t(row) * pitch;
	const uint32_t tile = row / 64u, r = row % 64u;
	float s = 0.f;
	for (uint32_t c = lane; c < kchunks * kTcChunkK; c += 32) {
		const float v = c < dim ? p[c] : 0.f;
		s = fmaf(v, v, s);
		const uint32_t kc = c / kTcChunkK, cc = c % kTcChunkK;
		const uint32_t unit = (cc >> 3) ^ (r & 7u);  // 16-byte unit = 8 bf16
		shadow[(size_t(tile) * kchunks + kc) * 4096u + r * 64u + unit * 8u + (cc & 7u)] = __float2bfloat16_rn(v);
	}
	for (int off = 16; off > 0; off >>= 1) {
		s += __shfl_xor_sync(0xffffffffu, s, off);
	}
	if (lane == 0 && vnorm) {
		vnorm[row] = sqrtf(s);
	}
}

// tau_init[q] = upper bound of the k1-th best distance among the first `nrows` (<= 1024) rows, fp32 dot products.  Any upper bound is
// valid; a small relative slack covers the difference to the arithmetic order of knn_scan_warp.  One block serves kTcInitQ queries
// (staged in shared memory, zero padded to the row pitch) so the rows come from L2 once per kTcInitQ queries; a warp keeps four rows
// (128-bit loads) in flight; the k1 smallest distances of a query are then picked by one warp (k1 rounds of a warp-wide argmin).
constexpr int kTcInitQ = 4;
constexpr uint32_t kTcInitRows = 1024;
__host__ __device__ inline size_t tc_init_smem_bytes(uint32_t pitch) { return size_t(kTcInitQ) * (pitch + kTcInitRows) * sizeof(float); }
__global__ void __launch_bounds__(256) tc_init_tau(const float* rows, uint32_t pitch, uint32_t dim, const float* norm_coefs, uint32_t nrows,
												   const float* queries, uint32_t nq, uint32_t k1, int metric, unsigned int* tau, float* ub_list,
												   unsigned int* ub_lock) {
	extern __shared__ __align__(16) float s_init[];
	float* s_q = s_init;                     // [kTcInitQ][pitch]
	float* s_d = s_init + kTcInitQ * pitch;  // [kTcInitQ][kTcInitRows]
	const uint32_t q0 = blockIdx.x * kTcInitQ;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint32_t i = threadIdx.x; i < kTcInitQ * pitch; i += blockDim.x) {
		const uint32_t qi = i / pitch, c = i % pitch;
		s_q[i] = (q0 + qi < nq && c < dim) ? queries[size_t(q0 + qi) * dim + c] : 0.f;
	}
	__syncthreads();
	const uint32_t pitch4 = pitch / 4;
	constexpr uint32_t kRows = 4;  // rows in flight per warp
	for (uint32_t r0 = warp * kRows; r0 < kTcInitRows; r0 += 8 * kRows) {
		float acc[kRows][kTcInitQ];
#pragma unroll
		for (uint32_t x = 0; x < kRows; ++x) {
#pragma unroll
			for (int qi = 0; qi < kTcInitQ; ++qi) {
				acc[x][qi] = 0.f;
			}
		}
		if (r0 < nrows) {
			const float4* p4[kRows];
#pragma unroll
			for (uint32_t x = 0; x < kRows; ++x) {
				p4[x] = reinterpret_cast<const float4*>(rows + size_t(min(r0 + x, nrows - 1)) * pitch);
			}
#pragma unroll 3
			for (uint32_t c = lane; c < pitch4; c += 32) {
				float4 v[kRows];
#pragma unroll
				for (uint32_t x = 0; x < kRows; ++x) {
					v[x] = __ldg(p4[x] + c);
				}
#pragma unroll
				for (int qi = 0; qi < kTcInitQ; ++qi) {
					const float4 qq = reinterpret_cast<const float4*>(s_q + qi * pitch)[c];
#pragma unroll
					for (uint32_t x = 0; x < kRows; ++x) {
						if (metric == kL2) {
							const float a0 = qq.x - v[x].x, a1 = qq.y - v[x].y, a2 = qq.z - v[x].z, a3 = qq.w - v[x].w;
							acc[x][qi] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, acc[x][qi]))));
						} else {
							acc[x][qi] = fmaf(qq.x, v[x].x, fmaf(qq.y, v[x].y, fmaf(qq.z, v[x].z, fmaf(qq.w, v[x].w, acc[x][qi]))));
						}
					}
				}
			}
		}
#pragma unroll
		for (uint32_t x = 0; x < kRows; ++x) {
#pragma unroll
			for (int qi = 0; qi < kTcInitQ; ++qi) {
				float s = acc[x][qi];
				for (int off = 16; off > 0; off >>= 1) {
					s += __shfl_xor_sync(0xffffffffu, s, off);
				}
				if (lane == uint32_t(qi)) {
					const uint32_t r = r0 + x;
					float d = INFINITY;
					if (r < nrows) {
						d = metric == kL2 ? s : -s;
						if (metric == kCos) {
							d *= norm_coefs[r];
						}
						d += 1e-4f * fabsf(d) + 1e-6f;
					}
					s_d[qi * kTcInitRows + r] = d;
				}
			}
		}
	}
	__syncthreads();
	if (warp < kTcInitQ && q0 + warp < nq) {  // warp w: the k1 smallest of query q0 + w (rows < nrows sort first, the rest are +inf)
		const uint32_t q = q0 + warp;
		float* d = s_d + warp * kTcInitRows;
		float last = INFINITY;
		for (uint32_t round = 0; round < kTcMaxK1; ++round) {
			float best = INFINITY;
			uint32_t at = lane;
			if (round < k1) {
				for (uint32_t j = lane; j < kTcInitRows; j += 32) {
					const float y = d[j];
					if (y < best) {
						best = y;
						at = j;
					}
				}
				for (int off = 16; off > 0; off >>= 1) {
					const float ob = __shfl_xor_sync(0xffffffffu, best, off);
					const uint32_t oa = __shfl_xor_sync(0xffffffffu, at, off);
					if (ob < best || (ob == best && oa < at)) {
						best = ob;
						at = oa;
					}
				}
				if (lane == 0) {
					d[at] = INFINITY;
				}
				__syncwarp();
				last = best;
			}
			if (lane == 0) {
				ub_list[size_t(q) * kTcMaxK1 + round] = round < k1 ? best : -INFINITY;
			}
		}
		if (lane == 0) {
			tau[q] = float_ord(last);  // +inf while fewer than k1 rows exist
			ub_lock[q] = 0;
		}
	}
}

// queries fp32 [nq][dim] -> bf16 [nq_pad][pitch_bf] (zero padded) + ||q||
__global__ void tc_prepare_queries(const float* queries, uint32_t nq, uint32_t nq_pad, uint32_t dim, uint32_t pitch_bf, __nv_bfloat16* out,
								   float* qnorm) {
	const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
	const int lane = threadIdx.x & 31;
	if (q >= nq_pad) {
		return;
	}
	float s = 0.f;
	for (uint32_t c = lane; c < pitch_bf; c += 32) {
		const float v = (q < nq && c < dim) ? queries[size_t(q) * dim + c] : 0.f;
		s = fmaf(v, v, s);
		out[size_t(q) * pitch_bf + c] = __float2bfloat16_rn(v);
	}
	for (int off = 16; off > 0; off >>= 1) {
		s += __shfl_xor_sync(0xffffffffu, s, off);
	}
	if (lane == 0 && q < nq) {
		qnorm[q] = sqrtf(s);
	}
}

}  // namespace rxgpu
