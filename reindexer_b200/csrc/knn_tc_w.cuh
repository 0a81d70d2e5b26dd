// Tensor-core candidate filter, wide-tile variant: UMMA N = 128 rows per MMA.
//
// Measured on B200 (tests/cuda/mma_rate.cu): an M = 128 tcgen05.mma.kind::f16 never takes less than 48 cycles, so N = 64 runs the
// tensor pipe at 67 % of its peak at best, while N = 96 / 128 / 256 reach 100 % (48 / 64 / 128 cycles).  knn_tc_filter_q uses
// N = 64 because two fp32 accumulators of 64 columns are all that fits next to the 384 columns of the query block in tensor memory.
// This variant spends the same 128 columns on ONE accumulator of N = 128: every MMA runs at the full rate, and the price is that
// the next tile's MMAs cannot start before the epilogue has pulled the accumulator into registers (4 x tcgen05.ld.x32, ~0.3 us
// against 1.7 us of MMAs per tile).
//
// Everything else is knn_tc_filter_q: query block (128 queries x dim bf16) in TMEM as the A operand, shared memory = ring of
// 32 KB stages (2 K chunks x 128 rows) filled with contiguous 8 KB bulk copies of the pre-swizzled shadow (a 128-row tile = two
// consecutive 64-row shadow tiles), CTA r of a cluster fetches every C-th copy and multicasts it, epilogue thread = query with the
// single-FMA candidate test and the per-row (||v||, w) ring.
// Roles (192 threads): warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue.
#pragma once
#include "knn_tc_q.cuh"

namespace rxgpu {

constexpr int kTwTileRows = 128;                               // UMMA N
constexpr int kTwSubBytes = kTwTileRows * 128;                 // 16 KB: 128 rows x one 64-element K chunk
constexpr int kTwSubsPerStage = 2;
constexpr int kTwStageBytes = kTwSubsPerStage * kTwSubBytes;   // 32 KB
constexpr int kTwThreads = 192;
constexpr uint32_t kTwVwSlots = 8, kTwVwAhead = 4;             // slots >= ahead + 3 (single accumulator, see the epilogue)

__host__ __device__ inline size_t tw_smem_bytes(uint32_t stages) {
	return 1024 + size_t(stages) * kTwStageBytes + kTwVwSlots * kTwTileRows * 8 + (2 * size_t(stages) + 8 + kTwVwSlots) * 8 + 64;
}

template <int kCluster>
__global__ void __launch_bounds__(kTwThreads, 1) knn_tc_filter_w(const TqArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_rows = base;  // [stages][2 K chunks][128 rows][128 B]
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kTwStageBytes);  // [kTwVwSlots][128]
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTwVwSlots * kTwTileRows);
	uint64_t* full_bar = bars;
	uint64_t* empty_bar = bars + a.stages;
	uint64_t* acc_full = bars + 2 * a.stages;
	uint64_t* acc_empty = acc_full + 1;
	uint64_t* q_ready = acc_empty + 1;
	uint64_t* vw_full = q_ready + 1;  // [kTwVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTwVwSlots);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t ntiles = (a.n + kTwTileRows - 1) / kTwTileRows;
	const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
	const uint32_t cid = blockIdx.x / kCluster, ncl = gridDim.x / kCluster;
	const uint32_t q0 = a.q0 + crank * kTqQueries;
	const uint32_t kparts = (a.kchunks + kTwSubsPerStage - 1) / kTwSubsPerStage;  // stages per tile

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < a.stages; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], kCluster);
		}
		mbar_init(acc_full, 1);
		mbar_init(acc_empty, 4);
		mbar_init(q_ready, 4);
		for (uint32_t s = 0; s < kTwVwSlots; ++s) {
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
		// ===== producer: a stage = 2 K chunks x (two 64-row shadow blocks of 8 KB) = 4 copies; CTA r issues copies r, r + C, ... =====
		if (lane == 0) {
			uint32_t stage = 0, phase = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl) {
				const unsigned char* tile_src = a.shadow + size_t(2 * t) * a.kchunks * kTqSubBytes;  // shadow tile 2t; 2t + 1 follows
				for (uint32_t kp = 0; kp < kparts; ++kp) {
					const uint32_t nsub = min(uint32_t(kTwSubsPerStage), a.kchunks - kTwSubsPerStage * kp);
					if (kp == 0) {
						TQ_TRACE(9, (t - cid) / ncl);
					}
					mbar_wait(&empty_bar[stage], phase ^ 1);
					if (kp == 0) {
						TQ_TRACE(10, (t - cid) / ncl);
					}
					mbar_expect_tx(&full_bar[stage], nsub * kTwSubBytes);
					unsigned char* dst = s_rows + size_t(stage) * kTwStageBytes;
					for (uint32_t i = crank; i < 2 * nsub; i += kCluster) {
						const uint32_t sub = i >> 1, half = i & 1;
						const unsigned char* src = tile_src + (size_t(half) * a.kchunks + kTwSubsPerStage * kp + sub) * kTqSubBytes;
						unsigned char* d = dst + sub * kTwSubBytes + half * kTqSubBytes;
						if constexpr (kCluster > 1) {
							bulk_load_mc(d, src, kTqSubBytes, &full_bar[stage], uint16_t((1u << kCluster) - 1u));
						} else {
							bulk_load(d, src, kTqSubBytes, &full_bar[stage]);
						}
						if (a.prefetch && uint64_t(t) + uint64_t(a.prefetch) * ncl < ntiles) {  // the same block, a.prefetch tiles ahead
							bulk_prefetch_l2(src + size_t(a.prefetch) * ncl * 2 * a.kchunks * kTqSubBytes, kTqSubBytes);
						}
					}
					if (++stage == a.stages) {
						stage = 0;
						phase ^= 1;
					}
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer: D[128 queries x 128 rows] += A(TMEM) x B(smem stage)^T =====
		if (lane == 0) {
			const uint32_t idesc = umma_idesc_bf16(kTqQueries, kTwTileRows);
			mbar_wait(q_ready, 0);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const uint32_t tmem_d = tmem_base + kTqAccCol0;
			uint32_t stage = 0, phase = 0, it = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
				TQ_TRACE(0, it);
				mbar_wait(acc_empty, (it & 1) ^ 1);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				TQ_TRACE(1, it);
				unsigned long long waited = 0;  // trace only: time spent waiting for the later stages of the tile
				for (uint32_t kp = 0; kp < kparts; ++kp) {
					const uint32_t nsub = min(uint32_t(kTwSubsPerStage), a.kchunks - kTwSubsPerStage * kp);
					const unsigned long long w0 = a.trace ? tq_clock() : 0ull;
					mbar_wait(&full_bar[stage], phase);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					if (kp == 0) {
						TQ_TRACE(2, it);
					} else if (a.trace) {
						waited += tq_clock() - w0;
					}
					const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kTwStageBytes);
					for (uint32_t sub = 0; sub < nsub; ++sub) {
#pragma unroll
						for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {
							umma_bf16_ts(tmem_d, tmem_base + ((kTwSubsPerStage * kp + sub) * 4 + k) * 8,
										 umma_desc_sw128(b_addr + sub * kTwSubBytes + k * 32), idesc, (kp | sub | k) != 0);
						}
					}
					if constexpr (kCluster > 1) {
						umma_commit_mc(&empty_bar[stage], uint16_t((1u << kCluster) - 1u));
					} else {
						umma_commit(&empty_bar[stage]);
					}
					if (++stage == a.stages) {
						stage = 0;
						phase ^= 1;
					}
				}
				umma_commit(acc_full);
				TQ_TRACE(3, it);
				if (a.trace && blockIdx.x == 0 && uint32_t(it - a.trace_first) < 256u) {
					a.trace[(it - a.trace_first) * 16 + 11] = waited;
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
				mbar_arrive(q_ready);
			}
		}
		const TqCandCtx cc{a.cand_count, a.cand_rows, a.ub_lock, a.ub_list, a.tau, a.cand_cap, a.init_rows, a.k1, a.metric};
		const float qe = q_ok ? kTcErrCoef * a.qnorm[my_q] : 0.f;
		float tau = q_ok ? ord_float(a.tau[my_q]) : -INFINITY;
		float2 pr = q_ok ? tc_make_pr(a.metric, tau, qe) : make_float2(0.f, INFINITY);
		// per-row (||v||, w) ring (1 KB per tile).  Slot reuse without an "empty" barrier: when the first epilogue thread starts tile
		// `it` it has passed acc_full(it - 1); those MMAs waited for the acc_empty arrivals of tile it - 2 from all four warps, which
		// every warp issues after finishing the compare loop of tile it - 3 -- so a slot last read by tile it + ahead - slots <= it - 3
		// is free.
		static_assert(kTwVwSlots >= kTwVwAhead + 3, "vw ring reuse distance");
		auto issue_vw = [&](uint32_t j) {
			const uint64_t t = uint64_t(cid) + uint64_t(j) * ncl;
			if (t < ntiles) {
				const uint32_t slot = j % kTwVwSlots;
				mbar_expect_tx(&vw_full[slot], kTwTileRows * 8);
				bulk_load(reinterpret_cast<unsigned char*>(s_vw + slot * kTwTileRows),
						  reinterpret_cast<const unsigned char*>(a.vw + t * kTwTileRows), kTwTileRows * 8, &vw_full[slot]);
			}
		};
		if (et == 0) {
			for (uint32_t j = 0; j < kTwVwAhead; ++j) {
				issue_vw(j);
			}
		}
		unsigned int tau_ahead = q_ok ? a.tau[my_q] : 0u;
		uint32_t it = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t rows_valid = min(uint32_t(kTwTileRows), a.n - t * kTwTileRows);
			if (et == 0) {
				issue_vw(it + kTwVwAhead);
			}
			const float2* vw_tile = s_vw + (it % kTwVwSlots) * kTwTileRows;
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
			mbar_wait(acc_full, it & 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (threadIdx.x == 64) {
				TQ_TRACE(5, it);
			}
			// the whole 128 x 128 accumulator into registers, then hand it back at once: the next tile's MMAs wait for this
			uint32_t vall[4][32];
#pragma unroll
			for (uint32_t ch = 0; ch < 4; ++ch) {
				tmem_ld32_nowait(tmem_base + kTqAccCol0 + ch * 32 + ((quad * 32) << 16), vall[ch]);
			}
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(acc_empty);
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(6, it);
			}
			mbar_wait(&vw_full[it % kTwVwSlots], (it / kTwVwSlots) & 1);
#pragma unroll
			for (uint32_t ch = 0; ch < 4; ++ch) {
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
						const float nt = tq_candidate(cc, my_q, t * kTwTileRows + c0 + j, __uint_as_float(v[j]), vw_tile[c0 + j].x, qe, tau);
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
			__syncwarp();  // the rare path diverges: reconverge before the .aligned tcgen05 ops of the next tile
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
