// Tensor-core candidate filter, third generation: UMMA N = 128 row tiles with a K-SPLIT query block.
//
// knn_tc_filter_q keeps the whole 128 x 768 bf16 query block in tensor memory (384 of the 512 columns), which leaves room for two
// accumulators of only 64 rows -- and an M = 128 tcgen05.mma never takes less than 48 cycles, so N = 64 caps the tensor pipe at 67 %
// (profiles/r1_microbench_tcgen05.txt).  Here the first 512 dimensions of the query block live in TMEM (256 columns, A operand of
// .ts MMAs), the remaining <= 256 dimensions live in shared memory (64 KB, SWIZZLE_128B K-major, A operand of .ss MMAs that
// accumulate into the same D), and the freed 128 columns make two accumulators of 128 rows: every MMA carries 64 cycles of work
// (100 % rate) and reads the A operand half as often per flop.
//
// TMEM map (512 columns): [0, 256) query dims 0..511 (two bf16 per column, lane = query), [256, 384) and [384, 512) the two fp32
//                         accumulators D[128 queries x 128 rows].
// Shared memory: [A tail: 4 K chunks x 16 KB][row ring: stages x 16 KB (one K chunk of a 128-row tile per stage)][(||v||, w) ring]
// Roles (320 threads, 1 CTA / SM): warp 0 producer (bulk copies of the pre-swizzled shadow; in a cluster of C CTAs each fetches 1/C of
// every chunk and multicasts it), warp 1 MMA issuer (tiles alternate between the accumulators), warps 2-5 / 6-9 epilogue of the even / odd tiles (thread = one
// query = one TMEM lane; one FFMA + compare per (query, row); the rare hit path is shared with knn_tc_filter_q).
// All single-thread instructions are issued from warp-uniform code under elect.sync (operands in uniform registers).
// Requires padded dim <= 768.  Same certified-bound candidate logic as knn_tc.cuh: results stay exact after the re-rank.
#pragma once
#include "knn_tc_q.cuh"

namespace rxgpu {

constexpr int kTkTileRows = 128;                           // UMMA N
constexpr int kTkChunkBytes = kTkTileRows * 128;           // 16 KB: 128 rows x 64 bf16 = one stage
constexpr uint32_t kTkTmemChunks = 8;                      // K chunks of the query block held in TMEM (256 columns)
constexpr uint32_t kTkTailChunks = kTqMaxKchunks - kTkTmemChunks;  // 4: K chunks of the query block held in shared memory
constexpr uint32_t kTkTailBytes = kTkTailChunks * kTqQueries * 128;  // 64 KB
constexpr uint32_t kTkAccCol0 = 256;
constexpr uint32_t kTkVwSlots = 16, kTkVwAhead = 4;
constexpr int kTkThreads = 320;                            // producer, issuer, two epilogue groups of 4 warps (even / odd tiles)

__host__ __device__ inline size_t tk_smem_bytes(uint32_t stages) {
	return 1024 + kTkTailBytes + size_t(stages) * kTkChunkBytes + kTkVwSlots * kTkTileRows * 8 + (2 * size_t(stages) + 8 + kTkVwSlots) * 8 + 64;
}

template <int kCluster>
__global__ void __launch_bounds__(kTkThreads, 1) knn_tc_filter_k(const TqArgs a) {
	extern __shared__ unsigned char smem_raw[];
	unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char* s_tail = base;                                      // [4 chunks][128 queries][128 B], SWIZZLE_128B
	unsigned char* s_rows = s_tail + kTkTailBytes;                     // [stages][128 rows][128 B]
	float2* s_vw = reinterpret_cast<float2*>(s_rows + size_t(a.stages) * kTkChunkBytes);  // [kTkVwSlots][128] per-row (||v||, w)
	uint64_t* bars = reinterpret_cast<uint64_t*>(s_vw + kTkVwSlots * kTkTileRows);
	uint64_t* full_bar = bars;
	uint64_t* empty_bar = bars + a.stages;
	uint64_t* acc_full = bars + 2 * a.stages;   // [2]
	uint64_t* acc_empty = acc_full + 2;          // [2]
	uint64_t* q_ready = acc_empty + 2;           // queries stored in TMEM / shared memory
	uint64_t* vw_full = q_ready + 1;             // [kTkVwSlots]
	uint32_t* s_tmem = reinterpret_cast<uint32_t*>(vw_full + kTkVwSlots);

	const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp index provably uniform
	const uint32_t ntiles = (a.n + kTkTileRows - 1) / kTkTileRows;
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
		for (uint32_t s = 0; s < kTkVwSlots; ++s) {
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
		// ===== producer: one stage = one K chunk of a 128-row tile = the 64-row shadow sub-tiles 2t and 2t+1 of that chunk (8 KB each).
		// A cluster of C CTAs splits the 16 KB into C pieces; CTA r fetches piece r and multicasts it to all =====
		uint32_t stage = 0, phase = 0;
		constexpr uint32_t kPiece = kTkChunkBytes / kCluster;  // 16 / 8 / 4 KB
		for (uint32_t t = cid; t < ntiles; t += ncl) {
			const unsigned char* tile_src = a.shadow + size_t(2 * t) * a.kchunks * kTqSubBytes;
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				if (kc == 0 && lane == 0) {
					TQ_TRACE(9, (t - cid) / ncl);
				}
				mbar_wait(&empty_bar[stage], phase ^ 1);
				if (kc == 0 && lane == 0) {
					TQ_TRACE(10, (t - cid) / ncl);
				}
				unsigned char* dst = s_rows + size_t(stage) * kTkChunkBytes;
				const unsigned char* src0 = tile_src + size_t(kc) * kTqSubBytes;                     // rows 0..63 of the tile
				const unsigned char* src1 = src0 + size_t(a.kchunks) * kTqSubBytes;                    // rows 64..127
				if (elect_one_sync()) {
					mbar_expect_tx(&full_bar[stage], kTkChunkBytes);
					if constexpr (kCluster > 1) {
						const uint32_t off = crank * kPiece;  // byte offset of my piece inside the 16 KB stage
						const unsigned char* src = (off < uint32_t(kTqSubBytes) ? src0 + off : src1 + (off - kTqSubBytes));
						bulk_load_mc(dst + off, src, kPiece, &full_bar[stage], uint16_t((1u << kCluster) - 1u));
						if (a.prefetch && uint64_t(t) + uint64_t(a.prefetch) * ncl < ntiles) {
							bulk_prefetch_l2(src + size_t(a.prefetch) * ncl * 2 * a.kchunks * kTqSubBytes, kPiece);
						}
					} else {
						bulk_load(dst, src0, kTqSubBytes, &full_bar[stage]);
						bulk_load(dst + kTqSubBytes, src1, kTqSubBytes, &full_bar[stage]);
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
		// ===== MMA issuer: D[128 queries x 128 rows] += A x B(smem stage)^T with A from TMEM (K chunks 0..7) or from shared memory
		// (K chunks 8..11); tiles alternate between the two accumulators.  ONE issuer: with the elected, uniform-register issue path a
		// tile's 48 instructions cost a few hundred cycles against 3072 cycles of tensor work, and a second issuer starting
		// kchunks > stages chunks ahead would wait on a ring barrier more than one phase ahead of it (parity waits cannot tell). =====
		const uint32_t idesc = umma_idesc_bf16(kTqQueries, kTkTileRows);
		mbar_wait(q_ready, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const uint32_t tail_addr = smem_u32(s_tail);
		uint32_t stage = 0, phase = 0, it = 0;
		for (uint32_t t = cid; t < ntiles; t += ncl, ++it) {
			const uint32_t acc = it & 1;
			const uint32_t tmem_d = tmem_base + kTkAccCol0 + acc * kTkTileRows;
			if (lane == 0) {
				TQ_TRACE(0, it);
			}
			mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if (lane == 0) {
				TQ_TRACE(1, it);
			}
			for (uint32_t kc = 0; kc < a.kchunks; ++kc) {
				mbar_wait(&full_bar[stage], phase);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (kc == 0 && lane == 0) {
					TQ_TRACE(2, it);
				}
				const uint32_t b_addr = smem_u32(s_rows + size_t(stage) * kTkChunkBytes);
				if (elect_one_sync()) {
					if (kc < kTkTmemChunks) {
						const uint32_t a_col = tmem_base + kc * 32;
#pragma unroll
						for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {  // K = 16 bf16 = 8 TMEM columns of A, 32 bytes of the B swizzle row
							umma_bf16_ts(tmem_d, a_col + k * 8, umma_desc_sw128(b_addr + k * 32), idesc, (kc | k) != 0);
						}
					} else {
						const uint32_t a_addr = tail_addr + (kc - kTkTmemChunks) * (kTqQueries * 128);
#pragma unroll
						for (uint32_t k = 0; k < kTcChunkK / 16; ++k) {
							umma_bf16(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, 1u);
						}
					}
					if constexpr (kCluster > 1) {
						umma_commit_mc(&empty_bar[stage], uint16_t((1u << kCluster) - 1u));
					} else {
						umma_commit(&empty_bar[stage]);
					}
					if (kc + 1 == a.kchunks) {
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
		}
	} else {
		// ===== epilogue: group 0 = warps 2..5 (even tiles, accumulator 0), group 1 = warps 6..9 (odd tiles, accumulator 1);
		// thread = query (TMEM lane quadrant = warp % 4) =====
		const uint32_t quad = warp & 3;
		const uint32_t grp = warp >= 6 ? 1u : 0u;
		const bool leader = warp == 2 || warp == 6;
		const uint32_t qrow = quad * 32 + lane;             // query row inside the block = TMEM lane
		const uint32_t my_q = q0 + qrow;                    // global query index
		const bool q_ok = my_q < a.nq_total;
		// 1. my query -> TMEM (K chunks 0..7, 32 columns = 64 bf16 per store) and -> shared memory (K chunks 8.., SWIZZLE_128B K-major:
		//    8-row groups 1024 B apart, the 16-byte units of row r XOR-permuted with r % 8); group 0 only
		if (grp == 0) {
			const uint4* src = reinterpret_cast<const uint4*>(a.qbf + size_t(q_ok ? my_q : 0) * a.pitch_bf);
			const uint32_t ntm = min(a.kchunks, kTkTmemChunks);
			for (uint32_t kc = 0; kc < ntm; ++kc) {
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
			for (uint32_t kc = kTkTmemChunks; kc < a.kchunks; ++kc) {
				unsigned char* rowp = s_tail + size_t(kc - kTkTmemChunks) * (kTqQueries * 128) + (qrow >> 3) * 1024 + (qrow & 7) * 128;
#pragma unroll
				for (uint32_t u = 0; u < 8; ++u) {
					const uint4 x = q_ok ? src[kc * 8 + u] : make_uint4(0, 0, 0, 0);
					*reinterpret_cast<uint4*>(rowp + ((u ^ (qrow & 7)) << 4)) = x;
				}
			}
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores above -> visible to the MMA's async-proxy reads
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
		// per-row terms (||v||, w): one 1 KB bulk copy per tile into a ring, issued kTkVwAhead tiles (two of the group's own) ahead by the
		// group's first warp (slot reuse argument as in knn_tc_filter_q: slots >= ahead + 12)
		static_assert(kTkVwSlots >= kTkVwAhead + 12, "vw ring reuse distance");
		auto issue_vw = [&](uint32_t j) {
			const uint64_t t = uint64_t(cid) + uint64_t(j) * ncl;
			if (t < ntiles) {
				const uint32_t slot = j % kTkVwSlots;
				mbar_expect_tx(&vw_full[slot], kTkTileRows * 8);
				bulk_load(reinterpret_cast<unsigned char*>(s_vw + slot * kTkTileRows),
						  reinterpret_cast<const unsigned char*>(a.vw + t * kTkTileRows), kTkTileRows * 8, &vw_full[slot]);
			}
		};
		if (leader) {
			if (elect_one_sync()) {
				issue_vw(grp);
				issue_vw(grp + 2);
			}
			__syncwarp();
		}
		unsigned int tau_ahead = q_ok ? a.tau[my_q] : 0u;
		for (uint32_t it = grp, t = cid + grp * ncl; t < ntiles; it += 2, t += 2 * ncl) {
			const uint32_t acc = grp, acc_phase = (it >> 1) & 1;
			const uint32_t rows_valid = min(uint32_t(kTkTileRows), a.n - t * kTkTileRows);
			if (leader) {
				if (elect_one_sync()) {
					issue_vw(it + kTkVwAhead);
				}
				__syncwarp();
			}
			const float2* vw_tile = s_vw + (it % kTkVwSlots) * kTkTileRows;
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
			mbar_wait(&vw_full[it % kTkVwSlots], (it / kTkVwSlots) & 1);
			const uint32_t acc_addr = tmem_base + kTkAccCol0 + acc * kTkTileRows + ((quad * 32) << 16);
#pragma unroll
			for (uint32_t half = 0; half < 2; ++half) {
				// 64 accumulator columns at a time; the TMEM buffer goes back to its issuer as soon as the second half is in registers
				uint32_t vall[2][32];
				tmem_ld32_nowait(acc_addr + half * 64, vall[0]);
				tmem_ld32_nowait(acc_addr + half * 64 + 32, vall[1]);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				if (half == 1) {
					asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
					__syncwarp();
					if (lane == 0) {
						mbar_arrive(&acc_empty[acc]);
					}
					if (threadIdx.x == 64) {
						TQ_TRACE(6, it);
					}
				}
#pragma unroll
				for (uint32_t ch = 0; ch < 2; ++ch) {
					const uint32_t c0 = half * 64 + ch * 32;
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
							const float nt = tq_candidate(cc, my_q, t * kTkTileRows + c0 + j, __uint_as_float(v[j]), vw_tile[c0 + j].x, qe, tau);
							if (nt < tau) {
								tau = nt;
								pr = tc_make_pr(a.metric, tau, qe);
							}
						}
					}
				}
				__syncwarp();  // the rare path diverges (per-lane lock loops): reconverge before the next .aligned tcgen05 op
			}
			if (threadIdx.x == 64) {
				TQ_TRACE(7, it);
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
