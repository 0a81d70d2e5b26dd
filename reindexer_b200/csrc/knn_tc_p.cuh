// Tensor-core candidate filter, fifth generation: CTA PAIRS multiply as one (tcgen05.mma.cta_group::2, M = 256 queries, N = 128 rows),
// each SM stages only HALF of every row tile, and the pairs of a cluster share the row stream through TMA multicast.
//
// What bounded the earlier generations (measured, round 2): every SM of knn_tc_filter_q / _n / _k has to pull the WHOLE row stream
// of its cluster into its own shared memory -- 2 KB per N = 64 MMA, 4 KB per N = 128 MMA -- and the fill rate of one SM's shared
// memory from L2 tops out near 32-36 B/cycle whatever the tile shape, cluster size or ring depth (q: 57 cycles per 2 KB, n: 133 and
// k: 124 cycles per 4 KB, cluster of 8: 29 B/cycle): the tensor pipe waits for operand bytes, not for issue slots.  A CTA pair halves
// that: the B operand of a cta_group::2 MMA is read from BOTH SMs' shared memory (rows 0..63 of the tile from the even CTA, rows
// 64..127 from the odd one), so an SM ingests 2 KB per N = 128 MMA = 64 cycles of tensor work.
//
//   A  [256 queries x K]  = the two CTAs' query blocks, each whole in its own tensor memory (384 columns, lane = query)
//   B  [128 rows    x K]  = one 8 KB pre-swizzled shadow block (64 rows x 64 K) per K chunk and SM
//   D  [256 x 128] fp32   = each CTA's 128 queries x all 128 rows in its own tensor memory, columns 384..511: ONE accumulator, handed
//                           back at once by the eight epilogue warps of each CTA (two tcgen05.ld.32x32b.x32 each, then arrive)
//
// Cluster of C = 2 P CTAs = P pairs (default C = 4): CTA rank r belongs to pair r / 2 and stages row half h = r % 2; the P CTAs with the
// same h split the K chunks of every stage between them and multicast them to each other, so HBM is read once per cluster pass
// (C x 128 queries) and an SM's ingest is half a row stream.  A stage is reused when the MMAs of ALL pairs have consumed it (every
// leader's tcgen05.commit is multicast to the whole cluster).
// Roles (320 threads, 1 CTA / SM): warp 0 producer; warp 1 = MMA issuer in the even CTA of a pair, "my half has landed" relay in the
// odd one; warps 2-5 / 6-9 epilogue of accumulator columns 0..63 / 64..127 (thread = one query = one TMEM lane; one FFMA + compare
// per (query, row); the rare hit path is shared with knn_tc_filter_q).  Single-thread instructions are issued from warp-uniform code
// under elect.sync.  Requires padded dim <= 768.  Same certified-bound candidate logic as knn_tc.cuh: results stay exact after re-rank.
#pragma once
#include "knn_tc_q.cuh"

namespace rxgpu {

constexpr int kTpTileRows = 128;                                // UMMA N over the pair
constexpr int kTpHalfRows = 64;                                 // rows staged per CTA (= one shadow block per K chunk)
constexpr int kTpSubsPerStage = 4;                              // K chunks per stage
constexpr int kTpStageBytes = kTpSubsPerStage * kTqSubBytes;    // 32 KB behind one mbarrier
constexpr uint32_t kTpAccCol0 = 384;
constexpr uint32_t kTpVwSlots = 8, kTpVwAhead = 4;
constexpr int kTpThreads = 320;

__host__ __device__ inline size_t tp_smem_bytes(uint32_t stages) {
	return 1024 + size_t(stages) * kTpStageBytes + kTpVwSlots * kTpTileRows * 8 + (2 * size_t(stages) + 3 + kTpVwSlots) * 8 + 64;
}

__device__ __forceinline__ void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
				 "h"(mask)
				 : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
	asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local barrier whose arrivals come from other CTAs of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
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

template <int kCluster>
__global__ void __launch_bounds__(kTpThreads, 1) knn_tc_filter_p(const TqArgs a) {
	static_assert(kCluster == 2 || kCluster == 4 || kCluster == 8, "whole pairs");
	constexpr uint32_t kPairs = kCluster / 2;
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_rows = base;  // [stages][4 K chunks][64 rows][128 B]: my half of the row tiles
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kTpStageBytes);  // [kTpVwSlots][128] per-row (||v||, w)
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTpVwSlots * kTpTileRows);
	uint64_t* full_bar = bars;                 // even CTA: own bytes + the odd CTA's relay; odd CTA: own bytes
	uint64_t* empty_bar = bars + a.stages;     // one multicast commit per pair and use
	uint64_t* acc_full = bars + 2 * a.stages;  // one commit per tile, multicast to the pair
	uint64_t* acc_empty = acc_full + 1;        // even CTA only: 8 epilogue warps of each CTA of the pair
	uint64_t* q_ready = acc_empty + 1;         // even CTA only: 8 + 8 warps
	uint64_t* vw_full = q_ready + 1;           // [kTpVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTpVwSlots);

	const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp index provably uniform
	const uint32_t ntiles = (a.n + kTpTileRows - 1) / kTpTileRows;
	const uint32_t crank = cluster_ctarank();
	const uint32_t rhalf = crank & 1u;          // which 64 rows of every tile I stage (0 = MMA leader of my pair)
	const uint32_t leader = crank & ~1u;        // cluster rank of my pair's leader
	const uint32_t pidx = crank >> 1;           // my pair
	const uint32_t cid = blockIdx.x / kCluster, ncl = gridDim.x / kCluster;
	const uint32_t q0 = a.q0 + crank * kTqQueries;
	const uint32_t kstages = (a.kchunks + kTpSubsPerStage - 1) / kTpSubsPerStage;  // stages per tile

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < a.stages; ++s) {
			mbar_init(&full_bar[s], rhalf == 0 ? 2 : 1);
			mbar_init(&empty_bar[s], kPairs);
		}
		mbar_init(acc_full, 1);
		mbar_init(acc_empty, 16);
		mbar_init(q_ready, 16);
		for (uint32_t s = 0; s < kTpVwSlots; ++s) {
			mbar_init(&vw_full[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {  // one warp of EACH CTA of the pair, same warp id, same destination offset
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	cluster_sync_all();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = *s_tmem;

	if (warp == 0) {
		// ===== producer: my row half of every K chunk; the kPairs CTAs with the same half take every kPairs-th chunk of a stage and
		// multicast it to each other =====
		uint16_t mask = 0;
		for (uint32_t j = 0; j < kPairs; ++j) {
			mask |= uint16_t(1u << (2 * j + rhalf));
		}
		uint32_t stage = 0, phase = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl) {
			const unsigned char* tile_src = a.shadow + size_t(2 * t + rhalf) * a.kchunks * kTqSubBytes;
			for (uint32_t ks = 0; ks < kstages; ++ks) {
				const uint32_t nsub = min(uint32_t(kTpSubsPerStage), a.kchunks - kTpSubsPerStage * ks);
				if (ks == 0 && lane == 0) {
					TQ_TRACE(9, (t - cid) / ncl);
				}
				mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
				if (ks == 0 && lane == 0) {
					TQ_TRACE(10, (t - cid) / ncl);
				}
				unsigned char* dst = s_rows + size_t(stage) * kTpStageBytes;
				const unsigned char* src = tile_src + size_t(kTpSubsPerStage * ks) * kTqSubBytes;
				if (elect_one_sync()) {
					mbar_expect_tx(&full_bar[stage], nsub * kTqSubBytes);
					for (uint32_t sub = pidx; sub < nsub; sub += kPairs) {
						if constexpr (kPairs > 1) {
							bulk_load_mc(dst + sub * kTqSubBytes, src + size_t(sub) * kTqSubBytes, kTqSubBytes, &full_bar[stage], mask);
						} else {
							bulk_load(dst + sub * kTqSubBytes, src + size_t(sub) * kTqSubBytes, kTqSubBytes, &full_bar[stage]);
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
	} else if (warp == 1) {
		if (rhalf == 0) {
			// ===== MMA issuer of the pair: D[256 queries x 128 rows] += A(TMEM of both CTAs) x B(my stage | the odd CTA's stage)^T =====
			const uint32_t idesc = umma_idesc_bf16(2 * kTqQueries, kTpTileRows);  // M = 256 over the pair, N = 128
			const uint32_t tmem_d = tmem_base + kTpAccCol0;
			const uint16_t pair_mask = uint16_t(3u << crank), all_mask = uint16_t((1u << kCluster) - 1u);
			mbar_wait_cluster(q_ready, 0);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			uint32_t stage = 0, phase = 0, it = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
				unsigned long long stall = 0;
				if (lane == 0) {
					TQ_TRACE(0, it);
				}
				mbar_wait_cluster(&full_bar[stage], phase);  // the tile's first stage, while the epilogues still drain the previous tile
				mbar_wait_cluster(acc_empty, (it & 1) ^ 1);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (lane == 0) {
					TQ_TRACE(1, it);
				}
				for (uint32_t ks = 0; ks < kstages; ++ks) {
					const uint32_t nsub = min(uint32_t(kTpSubsPerStage), a.kchunks - kTpSubsPerStage * ks);
					if (ks != 0) {
						const unsigned long long w0 = a.trace ? tq_clock() : 0ull;
						mbar_wait_cluster(&full_bar[stage], phase);
						if (a.trace) {
							stall += tq_clock() - w0;
						}
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					}
					if (ks == 0 && lane == 0) {
						TQ_TRACE(2, it);
					}
					const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kTpStageBytes);
					const uint32_t a_col = tmem_base + kTpSubsPerStage * ks * 32;
					if (elect_one_sync()) {
						for (uint32_t sub = 0; sub < nsub; ++sub) {
#pragma unroll
							for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {  // K = 16 bf16 = 8 TMEM columns of A, 32 bytes of the B swizzle row
								umma2_bf16_ts(tmem_d, a_col + (sub * 4 + k) * 8, umma_desc_sw128(b_addr + sub * kTqSubBytes + k * 32), idesc,
											  (ks | sub | k) != 0);
							}
						}
						umma2_commit_mc(&empty_bar[stage], all_mask);
						if (ks + 1 == kstages) {
							umma2_commit_mc(acc_full, pair_mask);
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
					if (a.trace && blockIdx.x == 0 && uint32_t(it - a.trace_first) < 256u) {
						a.trace[(it - a.trace_first) * 16 + 11] = stall;
					}
				}
			}
		} else {
			// ===== relay of the odd CTA: tell the leader when my half of a stage has landed =====
			uint32_t stage = 0, phase = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl) {
				for (uint32_t ks = 0; ks < kstages; ++ks) {
					mbar_wait(&full_bar[stage], phase);
					if (elect_one_sync()) {
						mbar_arrive_cluster(&full_bar[stage], leader);
					}
					__syncwarp();
					if (++stage == a.stages) {
						stage = 0;
						phase ^= 1;
					}
				}
			}
		}
	} else {
		// ===== epilogue: warps 2..5 = accumulator columns 0..63 (the rows the even CTA staged), warps 6..9 = columns 64..127; thread =
		// query (TMEM lane quadrant = warp % 4) =====
		const uint32_t quad = warp & 3;
		const uint32_t half = warp >= 6 ? 1u : 0u;
		const bool vw_leader = warp == 2;                    // issues the (||v||, w) copies of the CTA
		const uint32_t my_q = q0 + quad * 32 + lane;         // global query index of this TMEM lane
		const bool q_ok = my_q < a.nq_total;
		// 1. my query -> TMEM (A operand): 32 columns (64 bf16) per store; the two warps of a quadrant take alternate K chunks
		{
			const uint4* src = reinterpret_cast<const uint4*>(a.qbf + size_t(q_ok ? my_q : 0) * a.pitch_bf);
			for (uint32_t kc = half; kc < a.kchunks; kc += 2) {
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
				mbar_arrive_cluster(q_ready, leader);
			}
		}
		const TqCandCtx cc{a.cand_count, a.cand_rows, a.ub_lock, a.ub_list, a.tau, a.cand_cap, a.init_rows, a.k1, a.metric};
		const float qe = q_ok ? kTcErrCoef * a.qnorm[my_q] : 0.f;
		float tau = q_ok ? ord_float(a.tau[my_q]) : -INFINITY;
		float2 pr = q_ok ? tc_make_pr(a.metric, tau, qe) : make_float2(0.f, INFINITY);
		// per-row terms (||v||, w): one 1 KB bulk copy per tile into a ring of kTpVwSlots slots, issued kTpVwAhead tiles ahead by warp 2.
		// Slot reuse is safe without an "empty" barrier: when warp 2 starts tile `it` it has passed acc_full(it - 1); those MMAs waited
		// for the acc_empty arrivals of tile it - 2 from all sixteen warps of the pair, so every warp of this CTA has its tile it - 2
		// values in registers and is done with the slot of tile it - 3 -- slot (it + ahead) % slots was last read for tile <= it - 3.
		static_assert(kTpVwSlots >= kTpVwAhead + 3, "vw ring reuse distance");
		auto issue_vw = [&](uint32_t j) {
			const uint64_t t = uint64_t(cid) + uint64_t(j) * ncl;
			if (t < ntiles) {
				const uint32_t slot = j % kTpVwSlots;
				mbar_expect_tx(&vw_full[slot], kTpTileRows * 8);
				bulk_load(reinterpret_cast<unsigned char*>(s_vw + slot * kTpTileRows),
						  reinterpret_cast<const unsigned char*>(a.vw + t * kTpTileRows), kTpTileRows * 8, &vw_full[slot]);
			}
		};
		if (vw_leader) {
			if (elect_one_sync()) {
				for (uint32_t j = 0; j < kTpVwAhead; ++j) {
					issue_vw(j);
				}
			}
			__syncwarp();
		}
		unsigned int tau_ahead = q_ok ? a.tau[my_q] : 0u;
		const uint32_t acc_addr = tmem_base + kTpAccCol0 + half * 64 + ((quad * 32) << 16);
		uint32_t it = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t row0 = t * kTpTileRows + half * 64;
			const uint32_t rows_valid = a.n > row0 ? min(64u, a.n - row0) : 0u;
			if (vw_leader) {
				if (elect_one_sync()) {
					issue_vw(it + kTpVwAhead);
				}
				__syncwarp();
			}
			const float2* vw_tile = s_vw + (it % kTpVwSlots) * kTpTileRows + half * 64;
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
			mbar_wait_cluster(acc_full, it & 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (threadIdx.x == 64) {
				TQ_TRACE(5, it);
			}
			// my 64 columns of the accumulator -> registers, and the buffer goes back to the issuer BEFORE the values are looked at
			uint32_t vall[2][32];
			tmem_ld32_nowait(acc_addr, vall[0]);
			tmem_ld32_nowait(acc_addr + 32, vall[1]);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive_cluster(acc_empty, leader);
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(6, it);
			}
			mbar_wait(&vw_full[it % kTpVwSlots], (it / kTpVwSlots) & 1);
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
						const float nt = tq_candidate(cc, my_q, a.row_base + row0 + c0 + j, __uint_as_float(v[j]), vw_tile[c0 + j].x, qe, tau);
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
		}
	}
	__syncthreads();
	cluster_sync_all();  // the leader's MMAs read my shared memory and write my tensor memory until the very end
	if (warp == 1) {
		asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
	}
}

}  // namespace rxgpu
