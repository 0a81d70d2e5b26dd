// Tensor-core candidate filter, four-issuer variant of knn_tc_filter_q.
//
// Measured on B200 (tests/cuda/mma_rate.cu, mma_contention.cu): an M = 128 tcgen05.mma never takes less than 48 cycles, and ONE
// issuing thread whose stream contains tcgen05.commit sustains only one MMA per ~110 cycles, whatever N <= 128 is; several issuing
// threads overlap (the two issuers of knn_tc_filter_q reach ~70 cycles per MMA together).  Here every tile has TWO issuers that
// split its K range (even / odd stages), so four threads feed the tensor pipe: warps 1 and 6 as before (even / odd tiles,
// accumulator 0 / 1), warps 7 and 8 the other half of the stages of the same tiles.  Two threads accumulate into one accumulator
// in no particular order, so no MMA may be the "first" (overwriting) one: all accumulate, and the epilogue clears the accumulator
// with tcgen05.st after pulling it (before handing it back).  Stages are 16 KB (2 K chunks) so that both halves have work.
#pragma once
#include "knn_tc_q.cuh"

namespace rxgpu {

constexpr int kT4SubsPerStage = 2;
constexpr int kT4StageBytes = kT4SubsPerStage * kTqSubBytes;  // 16 KB
constexpr int kT4Threads = 288;                               // producer, issuer A0, 4 epilogue warps, issuer B0, issuers A1, B1

__host__ __device__ inline size_t t4_smem_bytes(uint32_t stages) {
	return 1024 + size_t(stages) * kT4StageBytes + kTqVwSlots * kTqTileRows * 8 + (2 * size_t(stages) + 8 + kTqVwSlots) * 8 + 64;
}

template <int kCluster>
__global__ void __launch_bounds__(kT4Threads, 1) knn_tc_filter_q4(const TqArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_rows = base;  // [stages][64 rows][128 B]
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kT4StageBytes);  // [kTqVwSlots][64] per-row (||v||, w)
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTqVwSlots * kTqTileRows);
	uint64_t* full_bar = bars;
	uint64_t* empty_bar = bars + a.stages;
	uint64_t* acc_full = bars + 2 * a.stages;   // [2]
	uint64_t* acc_empty = acc_full + 2;          // [2]
	uint64_t* q_ready = acc_empty + 2;           // queries stored in TMEM
	uint64_t* vw_full = q_ready + 1;             // [kTqVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTqVwSlots);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
			mbar_init(&acc_full[s], 2);  // both issuers of the tile commit
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
		// multicasts them to all CTAs =====
		if (lane == 0) {
			const uint32_t kpairs = (a.kchunks + kT4SubsPerStage - 1) / kT4SubsPerStage;
			uint32_t stage = 0, phase = 0;
			for (uint32_t t = cid; t < ntiles; t += ncl) {
				const unsigned char* tile_src = a.shadow + size_t(t) * a.kchunks * kTqSubBytes;
				for (uint32_t kp = 0; kp < kpairs; ++kp) {
					const uint32_t nsub = min(uint32_t(kT4SubsPerStage), a.kchunks - kT4SubsPerStage * kp);
					if (kp == 0) {
						TQ_TRACE(9, (t - cid) / ncl);
					}
					mbar_wait(&empty_bar[stage], phase ^ 1);
					if (kp == 0) {
						TQ_TRACE(10, (t - cid) / ncl);
					}
					mbar_expect_tx(&full_bar[stage], nsub * kTqSubBytes);
					unsigned char* dst = s_rows + size_t(stage) * kT4StageBytes;
					const unsigned char* src = tile_src + size_t(kT4SubsPerStage * kp) * kTqSubBytes;
					if constexpr (kCluster > 1) {
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
					if (++stage == a.stages) {
						stage = 0;
						phase ^= 1;
					}
				}
			}
		}
	} else if (warp == 1 || warp >= 6) {
		// ===== MMA issuers: warps 1 / 6 -> even / odd tiles, even stages; warps 7 / 8 -> even / odd tiles, odd stages =====
		// D[128 queries x 64 rows] += A(TMEM) x B(smem stage)^T, always accumulating (the epilogue hands back a cleared accumulator)
		if (lane == 0) {
			const uint32_t parity = (warp == 1 || warp == 7) ? 0u : 1u;
			const uint32_t part = warp >= 7 ? 1u : 0u;
			const uint32_t idesc = umma_idesc_bf16(kTqQueries, kTqTileRows);
			const uint32_t kpairs = (a.kchunks + kT4SubsPerStage - 1) / kT4SubsPerStage;
			mbar_wait(q_ready, 0);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const uint32_t tmem_d = tmem_base + kTqAccCol0 + parity * kTqTileRows;
			for (uint32_t it = parity, t = cid + parity * ncl; t < ntiles; it += 2, t += 2 * ncl) {
				if (part == 0) {
					TQ_TRACE(0, it);
				}
				mbar_wait(&acc_empty[parity], ((it >> 1) & 1) ^ 1);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (part == 0) {
					TQ_TRACE(1, it);
				}
				const uint32_t sidx0 = it * kpairs;
				for (uint32_t kp = part; kp < kpairs; kp += 2) {
					const uint32_t sidx = sidx0 + kp;
					const uint32_t stage = sidx % a.stages, phase = (sidx / a.stages) & 1;
					const uint32_t nsub = min(uint32_t(kT4SubsPerStage), a.kchunks - kT4SubsPerStage * kp);
					mbar_wait(&full_bar[stage], phase);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					if (kp == 0) {
						TQ_TRACE(2, it);
					}
					const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kT4StageBytes);
					for (uint32_t sub = 0; sub < nsub; ++sub) {
#pragma unroll
						for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {
							umma_bf16_ts(tmem_d, tmem_base + ((kT4SubsPerStage * kp + sub) * 4 + k) * 8,
										 umma_desc_sw128(b_addr + sub * kTqSubBytes + k * 32), idesc, 1u);
						}
					}
					if constexpr (kCluster > 1) {
						umma_commit_mc(&empty_bar[stage], uint16_t((1u << kCluster) - 1u));
					} else {
						umma_commit(&empty_bar[stage]);
					}
				}
				umma_commit(&acc_full[parity]);
				if (part == 0) {
					TQ_TRACE(3, it);
				}
			}
		}
	} else {
		// ===== epilogue warps 2..5: thread = query (TMEM lane quadrant = warp % 4) =====
		const uint32_t quad = warp & 3;
		const uint32_t et = threadIdx.x - 64;              // 0..127 inside the epilogue group
		const uint32_t my_q = q0 + quad * 32 + lane;       // global query index of this TMEM lane
		const bool q_ok = my_q < a.nq_total;
		// 1. my query -> TMEM (A operand): 32 columns (64 bf16) per store
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
			{  // both accumulators start cleared: every MMA accumulates
				uint32_t z[32];
#pragma unroll
				for (int i = 0; i < 32; ++i) {
					z[i] = 0;
				}
#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					tmem_st32(tmem_base + kTqAccCol0 + c * 32 + ((quad * 32) << 16), z);
				}
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
		// per-row terms (||v||, w): one 512-byte bulk copy per tile into a ring of kTqVwSlots slots, issued kTqVwAhead tiles ahead by
		// the first epilogue thread, so no global-load latency and no CTA barrier sits on the epilogue path.  Slot reuse is safe
		// without an "empty" barrier: when this thread starts tile `it` it has passed acc_full(it - 1); those MMAs waited for the
		// acc_empty arrivals of tile it - 3 from all four warps, which every warp issues after finishing tile it - 4 -- the last
		// reader of slot (it + kTqVwAhead) % kTqVwSlots.
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
			{  // clear my quadrant of the accumulator before handing it back
				uint32_t z[32];
#pragma unroll
				for (int i = 0; i < 32; ++i) {
					z[i] = 0;
				}
				tmem_st32(tmem_base + kTqAccCol0 + acc * kTqTileRows + ((quad * 32) << 16), z);
				tmem_st32(tmem_base + kTqAccCol0 + acc * kTqTileRows + 32 + ((quad * 32) << 16), z);
				asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			}
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
