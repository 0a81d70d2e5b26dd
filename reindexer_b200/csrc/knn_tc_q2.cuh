// Tensor-core candidate filter, third generation: the two CTAs of a pair multiply as ONE (tcgen05.mma.cta_group::2).
//
// knn_tc_filter_q keeps a whole 64-row tile (96 KB at 768 dims) in every SM's shared memory, so its 192 KB ring holds two tiles and
// both are being multiplied at any time: a released stage is refilled only after the HBM round trip, and the tensor pipe idles
// (measured: tensor pipe 46 % active, HBM 47 %, issuers wait for data).  Here a CTA PAIR executes each MMA together:
//   A  [256 queries x K]  = the two CTAs' query blocks, each in its own tensor memory (as before)
//   B  [ 64 rows    x K]  = rows 0..31 from the even CTA's shared memory, rows 32..63 from the odd CTA's  -> each SM stages HALF a tile
//   D  [256 x 64]         = each CTA's 128 queries x all 64 rows, in its own tensor memory (as before)
// so the same shared memory holds 4.5 tiles per SM, the pair reads every row tile from HBM once (no multicast needed), and one
// thread issues the MMAs of both SMs (half the issue work per SM).
//
// Roles (224 threads per CTA): warp 0 producer (my 32-row half of every stage, plain bulk copies into my own shared memory);
// warps 1 and 6 of the EVEN CTA issue the MMAs (even / odd tiles -> accumulator 0 / 1) and commit with a multicast to both CTAs;
// warp 1 of the ODD CTA relays "my half of stage s has landed" to the even CTA's full barrier; warps 2-5 epilogue exactly as in
// knn_tc_filter_q (thread = query; the accumulator hand-back and the query-ready signal go to the even CTA's barriers).
#pragma once
#include "knn_tc_q.cuh"

namespace rxgpu {

constexpr int kT2HalfRows = kTqTileRows / 2;                    // rows staged per CTA
constexpr int kT2SubBytes = kT2HalfRows * 128;                  // 4 KB: my half of one K chunk of a tile
constexpr int kT2SubsPerStage = 6;                              // K chunks per stage
constexpr int kT2StageBytes = kT2SubsPerStage * kT2SubBytes;    // 24 KB behind one mbarrier

__host__ __device__ inline size_t t2_smem_bytes(uint32_t stages) {
	return 1024 + size_t(stages) * kT2StageBytes + kTqVwSlots * kTqTileRows * 8 + (2 * size_t(stages) + 8 + kTqVwSlots) * 8 + 64;
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
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// same, without release semantics: for hand-offs whose payload is tensor memory (ordered by the tcgen05 fences around the
// arrive), so that the arrive does not wait for this thread's outstanding global stores and atomics (measured: 0.8 us per tile)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
	asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local barrier whose arrivals come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}

__global__ void __launch_bounds__(kTqThreads, 1) knn_tc_filter_q2(const TqArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_rows = base;  // [stages][6 K chunks][32 rows][128 B]
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kT2StageBytes);  // [kTqVwSlots][64]
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTqVwSlots * kTqTileRows);
	uint64_t* full_bar = bars;                   // even CTA: own bytes + the odd CTA's relay; odd CTA: own bytes
	uint64_t* empty_bar = bars + a.stages;       // one multicast commit per use
	uint64_t* acc_full = bars + 2 * a.stages;    // [2] one multicast commit per tile
	uint64_t* acc_empty = acc_full + 2;          // [2] even CTA only: 4 epilogue warps of each CTA
	uint64_t* q_ready = acc_empty + 2;           // even CTA only: 4 + 4 warps
	uint64_t* vw_full = q_ready + 1;             // [kTqVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTqVwSlots);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t ntiles = (a.n + kTqTileRows - 1) / kTqTileRows;
	const uint32_t crank = cluster_ctarank();    // 0 = even CTA = MMA leader
	const uint32_t cid = blockIdx.x / 2, ncl = gridDim.x / 2;
	const uint32_t q0 = a.q0 + crank * kTqQueries;
	const uint32_t kparts = (a.kchunks + kT2SubsPerStage - 1) / kT2SubsPerStage;  // stages per tile

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < a.stages; ++s) {
			mbar_init(&full_bar[s], crank == 0 ? 2 : 1);
			mbar_init(&empty_bar[s], 1);
		}
		for (int s = 0; s < 2; ++s) {
			mbar_init(&acc_full[s], 1);
			mbar_init(&acc_empty[s], 8);
		}
		mbar_init(q_ready, 8);
		for (uint32_t s = 0; s < kTqVwSlots; ++s) {
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
		// ===== producer: my 32 rows of every K chunk (the first / second 4 KB of the tile's 8 KB chunk blocks) =====
		if (lane == 0) {
			uint32_t stage = 0, phase = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl) {
				const unsigned char* tile_src = a.shadow + size_t(t) * a.kchunks * kTqSubBytes + size_t(crank) * kT2SubBytes;
				for (uint32_t kp = 0; kp < kparts; ++kp) {
					const uint32_t nsub = min(uint32_t(kT2SubsPerStage), a.kchunks - kT2SubsPerStage * kp);
					if (kp == 0) {
						TQ_TRACE(9, (t - cid) / ncl);
					}
					mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
					if (kp == 0) {
						TQ_TRACE(10, (t - cid) / ncl);
					}
					mbar_expect_tx(&full_bar[stage], nsub * kT2SubBytes);
					unsigned char* dst = s_rows + size_t(stage) * kT2StageBytes;
					const unsigned char* src = tile_src + size_t(kT2SubsPerStage * kp) * kTqSubBytes;
					for (uint32_t sub = 0; sub < nsub; ++sub) {
						bulk_load(dst + sub * kT2SubBytes, src + size_t(sub) * kTqSubBytes, kT2SubBytes, &full_bar[stage]);
					}
					if (a.prefetch && uint64_t(t) + uint64_t(a.prefetch) * ncl < ntiles) {
						const unsigned char* ahead = src + size_t(a.prefetch) * ncl * a.kchunks * kTqSubBytes;
						for (uint32_t sub = 0; sub < nsub; ++sub) {
							bulk_prefetch_l2(ahead + size_t(sub) * kTqSubBytes, kT2SubBytes);
						}
					}
					if (++stage == a.stages) {
						stage = 0;
						phase ^= 1;
					}
				}
			}
		}
	} else if (warp == 1 || warp == 6) {
		if (crank == 0) {
			// ===== MMA issuers of the pair: warp 1 -> even tiles / accumulator 0, warp 6 -> odd tiles / accumulator 1 =====
			if (lane == 0) {
				const uint32_t parity = warp == 1 ? 0u : 1u;
				const uint32_t idesc = umma_idesc_bf16(2 * kTqQueries, kTqTileRows);  // M = 256 over the pair, N = 64
				mbar_wait_cluster(q_ready, 0);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t tmem_d = tmem_base + kTqAccCol0 + parity * kTqTileRows;
				for (uint32_t it = parity, t = cid + parity * ncl; t < ntiles; it += 2, t += 2 * ncl) {
					TQ_TRACE(0, it);
					mbar_wait_cluster(&acc_empty[parity], ((it >> 1) & 1) ^ 1);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					TQ_TRACE(1, it);
					const uint32_t sidx0 = it * kparts;
					for (uint32_t kp = 0; kp < kparts; ++kp) {
						const uint32_t sidx = sidx0 + kp;
						const uint32_t stage = sidx % a.stages, phase = (sidx / a.stages) & 1;
						const uint32_t nsub = min(uint32_t(kT2SubsPerStage), a.kchunks - kT2SubsPerStage * kp);
						mbar_wait_cluster(&full_bar[stage], phase);
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						if (kp == 0) {
							TQ_TRACE(2, it);
						}
						const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kT2StageBytes);
						for (uint32_t sub = 0; sub < nsub; ++sub) {
#pragma unroll
							for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {
								umma2_bf16_ts(tmem_d, tmem_base + ((kT2SubsPerStage * kp + sub) * 4 + k) * 8,
											  umma_desc_sw128(b_addr + sub * kT2SubBytes + k * 32), idesc, (kp | sub | k) != 0);
							}
						}
						umma2_commit_mc(&empty_bar[stage], 3);
					}
					umma2_commit_mc(&acc_full[parity], 3);
					TQ_TRACE(3, it);
				}
			}
		} else if (warp == 1) {
			// ===== relay of the odd CTA: tell the leader when my half of a stage has landed =====
			if (lane == 0) {
				uint32_t stage = 0, phase = 0;
				for (uint32_t t = cid; t < ntiles; t += ncl) {
					for (uint32_t kp = 0; kp < kparts; ++kp) {
						mbar_wait(&full_bar[stage], phase);
						mbar_arrive_cluster(&full_bar[stage], 0);
						if (++stage == a.stages) {
							stage = 0;
							phase ^= 1;
						}
					}
				}
			}
		}
	} else {
		// ===== epilogue warps 2..5: thread = query (TMEM lane quadrant = warp % 4) =====
		const uint32_t quad = warp & 3;
		const uint32_t et = threadIdx.x - 64;
		const uint32_t my_q = q0 + quad * 32 + lane;
		const bool q_ok = my_q < a.nq_total;
		{
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
				mbar_arrive_cluster(q_ready, 0);
			}
		}
		const TqCandCtx cc{a.cand_count, a.cand_rows, a.ub_lock, a.ub_list, a.tau, a.cand_cap, a.init_rows, a.k1, a.metric};
		const float qe = q_ok ? kTcErrCoef * a.qnorm[my_q] : 0.f;
		float tau = q_ok ? ord_float(a.tau[my_q]) : -INFINITY;
		float2 pr = q_ok ? tc_make_pr(a.metric, tau, qe) : make_float2(0.f, INFINITY);
		// per-row (||v||, w) ring, see knn_tc_filter_q: the slot reuse argument holds per CTA (acc_empty needs all 8 warps of the pair)
		static_assert(kTqVwSlots >= kTqVwAhead + 4, "vw ring reuse distance");
		auto issue_vw = [&](uint32_t j) {
			const uint64_t t = uint64_t(cid) + uint64_t(j) * ncl;
			if (t < ntiles) {
				const uint32_t slot = j % kTqVwSlots;
				mbar_expect_tx(&vw_full[slot], kTqTileRows * 8);
				bulk_load(reinterpret_cast<unsigned char*>(s_vw + slot * kTqTileRows),
						  reinterpret_cast<const unsigned char*>(a.vw + t * kTqTileRows), kTqTileRows * 8, &vw_full[slot]);
			}
		};
		if (et == 0) {
			for (uint32_t j = 0; j < kTqVwAhead; ++j) {
				issue_vw(j);
			}
		}
		unsigned int tau_ahead = q_ok ? a.tau[my_q] : 0u;
		uint32_t it = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
			const uint32_t rows_valid = min(uint32_t(kTqTileRows), a.n - t * kTqTileRows);
			if (et == 0) {
				issue_vw(it + kTqVwAhead);
			}
			const float2* vw_tile = s_vw + (it % kTqVwSlots) * kTqTileRows;
			if (q_ok) {
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
			mbar_wait_cluster(&acc_full[acc], acc_phase);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (threadIdx.x == 64) {
				TQ_TRACE(5, it);
			}
			uint32_t vall[2][32];
			tmem_ld32_nowait(tmem_base + kTqAccCol0 + acc * kTqTileRows + ((quad * 32) << 16), vall[0]);
			tmem_ld32_nowait(tmem_base + kTqAccCol0 + acc * kTqTileRows + 32 + ((quad * 32) << 16), vall[1]);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive_cluster_relaxed(&acc_empty[acc], 0);
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
						const float nt = tq_candidate(cc, my_q, t * kTqTileRows + c0 + j, __uint_as_float(v[j]), vw_tile[c0 + j].x, qe, tau);
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
			__syncwarp();
		}
	}
	__syncthreads();
	cluster_sync_all();  // the leader's MMAs read my shared memory and write my tensor memory until the very end
	if (warp == 1) {
		asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
	}
}

}  // namespace rxgpu
