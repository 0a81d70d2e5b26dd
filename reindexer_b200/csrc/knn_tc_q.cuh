// Tensor-core candidate filter, second generation: the QUERY block lives in tensor memory (the A operand of tcgen05.mma read
// from TMEM), so all of shared memory is a deep TMA ring for the bf16 shadow rows, and a cluster of C CTAs shares every row tile
// through TMA multicast while each CTA owns a different block of 128 queries.
//
//   one pass over the shadow (n * dim * 2 bytes from HBM) now serves C * 128 queries           (knn_tc_filter: 96)
//   bytes in flight per SM: up to 24 stages x 8 KB = 192 KB                                    (knn_tc_filter: 64 KB)
//   per SM and pass: tiles_of_cluster x 48 MMAs (M=128 queries, N=64 rows, K=16)  -> with C=4 the kernel is balanced between
//   the tensor pipe and HBM; same certified-bound candidate logic as knn_tc.cuh (see there), so results stay exact after re-rank.
//
// TMEM map (512 columns): [0, dim_padded/2) the 128 x dim bf16 query block (two K elements per 32-bit column, lane = query),
//                         [384, 448) and [448, 512) two fp32 accumulators D[128 queries x 64 rows] (double buffered).
// Roles (352 threads): warp 0 TMA producer (its 64/C-row slice of every stage, multicast to the cluster); warps 1 and 6 issue the
// MMAs of the even / odd tiles into accumulator 0 / 1; TWO epilogue groups of four warps, warps 2-5 drain the even tiles
// (accumulator 0) and warps 7-10 the odd ones (accumulator 1), so a slow tile (the rare candidate path takes global atomics)
// never delays the hand-back of the other accumulator: thread = one query (TMEM lane) with P/R/tau of its query in registers,
// per-row terms broadcast from shared memory, one FFMA + compare per (query, row).
// Requires dim_padded <= 768.
#pragma once
#include "knn_tc.cuh"

namespace rxgpu {

constexpr int kTqTileRows = 64;                           // UMMA N
constexpr int kTqSubBytes = kTqTileRows * 128;            // 8 KB: 64 rows x 64 bf16 (one 128-byte swizzle atom wide)
constexpr int kTqSubsPerStage = 4;                        // K chunks per stage
constexpr int kTqStageBytes = kTqSubsPerStage * kTqSubBytes;  // a stage = 64 rows x 256 bf16 = 32 KB behind ONE mbarrier
constexpr int kTqThreads = 352;                           // producer, issuer A, epilogue group 0 (4 warps), issuer B, epilogue group 1 (4 warps)
constexpr int kTqQueries = 128;                           // UMMA M = queries per CTA
constexpr uint32_t kTqAccCol0 = 384;                      // first accumulator column
constexpr uint32_t kTqMaxKchunks = 12;                    // 768 / 64

struct TqArgs {
	const unsigned char* shadow;  // see tc_convert_rows: [tile64][K chunk][8 KB pre-swizzled]
	const float* vnorm;
	const float2* vw;         // tc_make_vw: [tiles x 64] (||v||, w)
	const float* vinv;
	const float* qnorm;
	const uint16_t* qbf;       // [nq_pad][pitch_bf] bf16 queries (row-major, zero padded)
	unsigned int* tau;
	float* ub_list;
	unsigned int* ub_lock;
	uint32_t* cand_rows;
	unsigned int* cand_count;
	uint32_t cand_cap;
	uint32_t init_rows;
	uint32_t n;
	uint32_t kchunks;
	uint32_t pitch_bf;
	uint32_t nq_total;
	uint32_t q0;               // first query of this launch; CTA rank r owns [q0 + 128 r, +128)
	uint32_t k1;
	uint32_t stages;
	int metric;
	unsigned long long* trace;  // profiling aid (RXGPU_TC_TRACE): per-tile timestamps of CTA 0, or nullptr
	uint32_t trace_first;       // first local tile index recorded (RXGPU_TC_TRACE_FIRST)
	uint32_t prefetch;          // L2 prefetch distance of the producers in tiles (0 = off)
	uint32_t single_issuer;     // knn_tc_filter_q: 1 = warp 1 issues every tile in order (stages are released at the full pipe rate)
	uint32_t row_base;          // global internal row of this launch's tile 0 (shadow / vw point at it, n counts the rows from there): the tail
	                            // grid on the SMs a cluster-of-4 grid strands scans its own row range
};

constexpr uint32_t kTqVwSlots = 16;  // ring of per-tile (||v||, w) blocks, filled kTqVwAhead tiles ahead by the epilogue itself
constexpr uint32_t kTqVwAhead = 4;   // slots >= ahead + 12: see the reuse argument in the epilogue
__host__ __device__ inline size_t tq_smem_bytes(uint32_t stages) {
	return 1024 + size_t(stages) * kTqStageBytes + kTqVwSlots * kTqTileRows * 8 + (2 * size_t(stages) + 8 + kTqVwSlots) * 8 + 64;
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
		"{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(
			taddr),
		"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
		"r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
		"r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
		: "memory");
}

__device__ __forceinline__ unsigned long long tq_clock() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define TQ_TRACE(slot, idx)                                                        \
	do {                                                                           \
		if (a.trace && blockIdx.x == 0 && uint32_t((idx) - a.trace_first) < 256u) { \
			a.trace[((idx) - a.trace_first) * 16 + (slot)] = tq_clock();           \
		}                                                                          \
	} while (0)

// The rare path of the epilogue: row `row` passed the test for query `my_q` (raw accumulator s, row norm vn).  Appends the candidate,
// and when its upper bound beats the current threshold, inserts it into the query's global bound list (under the per-query lock;
// this thread is the only one of its CTA that handles my_q, other clusters contend) and tightens tau.  Returns the new threshold.
// Kept out of line: it runs for ~350 of 10M rows per query, and inlining it 64 times per tile only costs instruction cache.
struct TqCandCtx {
	unsigned int* cand_count;
	uint32_t* cand_rows;
	unsigned int* ub_lock;
	float* ub_list;
	unsigned int* tau;
	uint32_t cand_cap, init_rows, k1;
	int metric;
};
__device__ __noinline__ float tq_candidate(const TqCandCtx c, uint32_t my_q, uint32_t row, float s, float vn, float qe, float tau) {
	float d, e;
	if (c.metric == kL2) {
		const float qn = qe * (1.f / kTcErrCoef);
		d = fmaf(-2.f, s, fmaf(qn, qn, vn * vn));
		e = 2.f * qe * vn + kTcL2Eps * (qn * qn + vn * vn);
	} else if (c.metric == kCos) {
		const float vinv = 1.f / vn;  // within 1e-5 of the stored coefficient (normalize.cc shortcut), inside the slack
		d = -s * vinv;
		e = qe * 1.0001f;
	} else {
		d = -s;
		e = qe * vn;
	}
	const unsigned pos = atomicAdd(&c.cand_count[my_q], 1u);
	if (pos < c.cand_cap) {
		c.cand_rows[size_t(my_q) * c.cand_cap + pos] = row;
	}
	const float ub = d + e;
	if (ub < tau && row >= c.init_rows) {
		while (atomicCAS(&c.ub_lock[my_q], 0u, 1u) != 0u) {
		}
		__threadfence();
		volatile float* list = c.ub_list + size_t(my_q) * kTcMaxK1;
		uint32_t mi = 0;
		float mx = list[0];
		for (uint32_t x = 1; x < c.k1; ++x) {
			const float y = list[x];
			if (y > mx) {
				mx = y;
				mi = x;
			}
		}
		if (ub < mx) {
			list[mi] = ub;
			float nmx = list[0];
			for (uint32_t x = 1; x < c.k1; ++x) {
				nmx = fmaxf(nmx, list[x]);
			}
			atomicMin(&c.tau[my_q], float_ord(nmx));
			tau = fminf(tau, nmx);
		} else {
			tau = fminf(tau, mx);
		}
		__threadfence();
		atomicExch(&c.ub_lock[my_q], 0u);
	}
	return tau;
}

template <int kCluster>
__global__ void __launch_bounds__(kTqThreads, 1) knn_tc_filter_q(const TqArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_rows = base;  // [stages][64 rows][128 B]
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kTqStageBytes);  // [kTqVwSlots][64] per-row (||v||, w)
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTqVwSlots * kTqTileRows);
	uint64_t* full_bar = bars;
	uint64_t* empty_bar = bars + a.stages;
	uint64_t* acc_full = bars + 2 * a.stages;   // [2]
	uint64_t* acc_empty = acc_full + 2;          // [2]
	uint64_t* q_ready = acc_empty + 2;           // queries stored in TMEM
	uint64_t* vw_full = q_ready + 1;             // [kTqVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTqVwSlots);

	const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp index provably uniform
	const uint32_t ntiles = (a.n + kTqTileRows - 1) / kTqTileRows;
	const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
	const uint32_t cid = blockIdx.x / kCluster, ncl = gridDim.x / kCluster;
	const uint32_t q0 = a.q0 + crank * kTqQueries;

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < a.stages; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], kCluster);
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(&acc_full[s], 1);
			mbar_init(&acc_empty[s], 4);
		}
		mbar_init(q_ready, 4);
		for (uint32_t s = 0; s < kTqVwSlots; ++s) {
			mbar_init(&vw_full[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if constexpr (kCluster > 1) {
		cluster_sync_all();
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = *s_tmem;

	if (warp == 0) {
		// ===== producer: contiguous 8 KB bulk copies; in a cluster CTA r fetches the K chunks with (chunk % C == r) of every stage and
		// multicasts them to all CTAs.  The WHOLE warp walks the loop (convergent code: addresses and counters live in uniform
		// registers) and one elected lane issues the copies -- inside an `if (lane == 0)` region ptxas cannot prove the operands
		// warp-uniform and wraps every UBLKCP / UTCHMMA into an ELECT + R2UR waterfall loop. =====
		const uint32_t kpairs = (a.kchunks + kTqSubsPerStage - 1) / kTqSubsPerStage;
		uint32_t stage = 0, phase = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl) {
			const unsigned char* tile_src = a.shadow + size_t(t) * a.kchunks * kTqSubBytes;
			for (uint32_t kp = 0; kp < kpairs; ++kp) {
				const uint32_t nsub = min(uint32_t(kTqSubsPerStage), a.kchunks - kTqSubsPerStage * kp);
				if (kp == 0 && lane == 0) {
					TQ_TRACE(9, (t - cid) / ncl);
				}
				mbar_wait(&empty_bar[stage], phase ^ 1);
				if (kp == 0 && lane == 0) {
					TQ_TRACE(10, (t - cid) / ncl);
				}
				unsigned char* dst = s_rows + size_t(stage) * kTqStageBytes;
				const unsigned char* src = tile_src + size_t(kTqSubsPerStage * kp) * kTqSubBytes;
				if (elect_one_sync()) {
					mbar_expect_tx(&full_bar[stage], nsub * kTqSubBytes);
					if constexpr (kCluster > kTqSubsPerStage) {  // 8 CTAs: every 8 KB K chunk of the stage is fetched in two halves
						constexpr uint32_t kParts = kCluster / kTqSubsPerStage, kPart = kTqSubBytes / kParts;
						const uint32_t sub = crank / kParts, off = sub * kTqSubBytes + (crank % kParts) * kPart;
						if (sub < nsub) {
							bulk_load_mc(dst + off, src + off, kPart, &full_bar[stage], uint16_t((1u << kCluster) - 1u));
						}
					} else if constexpr (kCluster > 1) {
						for (uint32_t sub = crank; sub < nsub; sub += kCluster) {
							bulk_load_mc(dst + sub * kTqSubBytes, src + size_t(sub) * kTqSubBytes, kTqSubBytes, &full_bar[stage],
										 uint16_t((1u << kCluster) - 1u));
						}
					} else {
						bulk_load(dst, src, nsub * kTqSubBytes, &full_bar[stage]);  // the K chunks of a tile are contiguous
					}
					if (a.prefetch && uint64_t(t) + uint64_t(a.prefetch) * ncl < ntiles) {  // my share of the same stage, a.prefetch tiles ahead
						const unsigned char* ahead = src + size_t(a.prefetch) * ncl * a.kchunks * kTqSubBytes;
						for (uint32_t sub = crank; sub < nsub; sub += kCluster) {
							bulk_prefetch_l2(ahead + size_t(sub) * kTqSubBytes, kTqSubBytes);
						}
					}
				}
				__syncwarp();
				if (++stage == a.stages) {
					stage = 0;
					phase ^= 1;
				}
			}
		}
	} else if (warp == 1 || warp == 6) {
		// ===== MMA issuers: warp 1 -> even tiles / accumulator 0, warp 6 -> odd tiles / accumulator 1 =====
		// D[128 queries x 64 rows] += A(TMEM) x B(smem stage)^T.  Whole warp in the loop, one elected lane issues (see the producer).
		const bool single = a.single_issuer != 0;
		const uint32_t parity = warp == 1 ? 0u : 1u;
		const uint32_t idesc = umma_idesc_bf16(kTqQueries, kTqTileRows);
		const uint32_t kpairs = (a.kchunks + kTqSubsPerStage - 1) / kTqSubsPerStage;
		mbar_wait(q_ready, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		// two issuers: stages are consumed in tile order alternately, mine are [it * kpairs, (it + 1) * kpairs) for it = parity, parity + 2, ...
		// one issuer (a.single_issuer): warp 1 walks every tile in order, the accumulators alternate; warp 6 has nothing to do
		const uint32_t step = single ? 1u : 2u;
		const uint32_t first = single ? 0u : parity;
		uint32_t stage = (first * kpairs) % a.stages, phase = ((first * kpairs) / a.stages) & 1;
		for (uint32_t it = first, t = cid + first * ncl; t < ntiles && !(single && parity); it += step, t += step * ncl) {
			const uint32_t acc = single ? (it & 1u) : parity;
			const uint32_t tmem_d = tmem_base + kTqAccCol0 + acc * kTqTileRows;
			if (lane == 0) {
				TQ_TRACE(0, it);
			}
			mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (lane == 0) {
				TQ_TRACE(1, it);
			}
			for (uint32_t kp = 0; kp < kpairs; ++kp) {
				const uint32_t nsub = min(uint32_t(kTqSubsPerStage), a.kchunks - kTqSubsPerStage * kp);
				mbar_wait(&full_bar[stage], phase);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (kp == 0 && lane == 0) {
					TQ_TRACE(2, it);
				}
				const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kTqStageBytes);
				const uint32_t a_col = tmem_base + kTqSubsPerStage * kp * 32;
				if (elect_one_sync()) {
					for (uint32_t sub = 0; sub < nsub; ++sub) {
#pragma unroll
						for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {  // K = 16 bf16 = 8 TMEM columns of A, 32 bytes of the B swizzle row
							umma_bf16_ts(tmem_d, a_col + (sub * 4 + k) * 8, umma_desc_sw128(b_addr + sub * kTqSubBytes + k * 32), idesc,
										 (kp | sub | k) != 0);
						}
					}
					if constexpr (kCluster > 1) {
						umma_commit_mc(&empty_bar[stage], uint16_t((1u << kCluster) - 1u));
					} else {
						umma_commit(&empty_bar[stage]);
					}
					if (kp + 1 == kpairs) {
						umma_commit(&acc_full[acc]);
					}
				}
				__syncwarp();
				if (++stage == a.stages) {
					stage = 0;
					phase ^= 1;
				}
			}
			if (lane == 0) {
				TQ_TRACE(3, it);
			}
			for (uint32_t kp = 0; kp < kpairs && !single; ++kp) {  // skip the other issuer's stages
				if (++stage == a.stages) {
					stage = 0;
					phase ^= 1;
				}
			}
		}
	} else {
		// ===== epilogue: group 0 = warps 2..5 (even tiles, accumulator 0), group 1 = warps 7..10 (odd tiles, accumulator 1);
		// thread = query (TMEM lane quadrant = warp % 4) =====
		const uint32_t quad = warp & 3;
		const uint32_t grp = warp >= 7 ? 1u : 0u;
		const bool leader = warp == 2 || warp == 7;        // issues the group's (||v||, w) copies
		const uint32_t my_q = q0 + quad * 32 + lane;       // global query index of this TMEM lane
		const bool q_ok = my_q < a.nq_total;
		// 1. my query -> TMEM (A operand): 32 columns (64 bf16) per store (group 0 only)
		if (grp == 0) {
			const uint4* src = reinterpret_cast<const uint4*>(a.qbf + size_t(q_ok ? my_q : 0) * a.pitch_bf);
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				uint32_t r[32];
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					const uint4 x = q_ok ? src[kc * 8 + i] : make_uint4(0, 0, 0, 0);
					r[4 * i] = x.x;
					r[4 * i + 1] = x.y;
					r[4 * i + 2] = x.z;
					r[4 * i + 3] = x.w;
				}
				tmem_st32(tmem_base + kc * 32 + ((quad * 32) << 16), r);
			}
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(q_ready);
			}
		}
		const TqCandCtx cc{a.cand_count, a.cand_rows, a.ub_lock, a.ub_list, a.tau, a.cand_cap, a.init_rows, a.k1, a.metric};
		const float qe = q_ok ? kTcErrCoef * a.qnorm[my_q] : 0.f;
		float tau = q_ok ? ord_float(a.tau[my_q]) : -INFINITY;
		float2 pr = q_ok ? tc_make_pr(a.metric, tau, qe) : make_float2(0.f, INFINITY);
		// per-row terms (||v||, w): one 512-byte bulk copy per tile into a ring of kTqVwSlots slots, issued kTqVwAhead tiles (two
		// of the group's own) ahead by the group's first warp, so no global-load latency and no CTA barrier sits on the epilogue path.
		// Slot reuse is safe without an "empty" barrier: the group's tiles are it, it + 2, ...; when its leader starts tile `it` it has
		// passed acc_full(it - 2); those MMAs waited for the acc_empty arrivals of tile it - 4 from all four warps of the group, so
		// every warp is at least inside tile it - 4 and has finished tile it - 6 -- slot (it + 4) % 16 was last read for tile it - 12.
		static_assert(kTqVwSlots >= kTqVwAhead + 12, "vw ring reuse distance");
		auto issue_vw = [&](uint32_t j) {
			const uint64_t t = uint64_t(cid) + uint64_t(j) * ncl;
			if (t < ntiles) {
				const uint32_t slot = j % kTqVwSlots;
				mbar_expect_tx(&vw_full[slot], kTqTileRows * 8);
				bulk_load(reinterpret_cast<unsigned char*>(s_vw + slot * kTqTileRows),
						  reinterpret_cast<const unsigned char*>(a.vw + t * kTqTileRows), kTqTileRows * 8, &vw_full[slot]);
			}
		};
		if (leader) {  // convergent, one elected lane issues the copies of the group's first two tiles
			if (elect_one_sync()) {
				issue_vw(grp);
				issue_vw(grp + 2);
			}
			__syncwarp();
		}
		unsigned int tau_ahead = q_ok ? a.tau[my_q] : 0u;
		for (uint32_t it = grp, t = cid + grp * ncl; t < ntiles; it += 2, t += 2 * ncl) {
			const uint32_t acc = grp, acc_phase = (it >> 1) & 1;
			const uint32_t rows_valid = min(uint32_t(kTqTileRows), a.n - t * kTqTileRows);
			if (leader) {
				if (elect_one_sync()) {
					issue_vw(it + kTqVwAhead);
				}
				__syncwarp();
			}
			const float2* vw_tile = s_vw + (it % kTqVwSlots) * kTqTileRows;
			if (q_ok) {  // the threshold other CTAs tightened: loaded one tile ago, consumed now
				const float tn = ord_float(tau_ahead);
				if (tn < tau) {
					tau = tn;
					pr = tc_make_pr(a.metric, tau, qe);
				}
				tau_ahead = a.tau[my_q];
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(4, it);
			}
			mbar_wait(&acc_full[acc], acc_phase);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (threadIdx.x == 64) {
				TQ_TRACE(5, it);
			}
			// pull the whole 128 x 64 accumulator into registers and hand the TMEM buffer back to its issuer BEFORE looking at the
			// values: the buffer's turn-around time, not the compare loop, is on the critical path of the tensor pipe
			uint32_t vall[2][32];
			tmem_ld32_nowait(tmem_base + kTqAccCol0 + acc * kTqTileRows + ((quad * 32) << 16), vall[0]);
			tmem_ld32_nowait(tmem_base + kTqAccCol0 + acc * kTqTileRows + 32 + ((quad * 32) << 16), vall[1]);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(&acc_empty[acc]);
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(6, it);
			}
			mbar_wait(&vw_full[it % kTqVwSlots], (it / kTqVwSlots) & 1);
#pragma unroll
			for (uint32_t ch = 0; ch < 2; ++ch) {
				const uint32_t c0 = ch * 32;
				uint32_t (&v)[32] = vall[ch];
				uint32_t hits = 0;
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const float2 vw = vw_tile[c0 + j];
					hits |= uint32_t(__uint_as_float(v[j]) - vw.y >= fmaf(pr.x, vw.x, pr.y)) << j;
				}
				const uint32_t nv = rows_valid > c0 ? min(32u, rows_valid - c0) : 0u;
				hits &= nv >= 32 ? 0xFFFFFFFFu : ((1u << nv) - 1u);
				const unsigned any_hits = __reduce_or_sync(0xffffffffu, hits);
				if (any_hits) {
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						if (!(any_hits & (1u << j)) || !(hits & (1u << j))) {
							continue;
						}
						const float nt = tq_candidate(cc, my_q, a.row_base + t * kTqTileRows + c0 + j, __uint_as_float(v[j]), vw_tile[c0 + j].x, qe, tau);
						if (nt < tau) {
							tau = nt;
							pr = tc_make_pr(a.metric, tau, qe);
						}
					}
				}
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(7, it);
			}
			__syncwarp();  // the rare path diverges (per-lane lock loops): reconverge before the .aligned tcgen05 ops of the next tile
			if (threadIdx.x == 64) {
				TQ_TRACE(8, it);
			}
		}
	}
	__syncthreads();
	if constexpr (kCluster > 1) {
		cluster_sync_all();
	}
	if (warp == 1) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
	}
}

}  // namespace rxgpu
