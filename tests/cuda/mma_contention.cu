// Micro-benchmark (not product, not a test): does concurrent traffic slow tcgen05.mma down?
//   base   : one thread issues TS MMAs (M=128, N, K=16), B rotating over a 192 KB shared-memory ring
//   +tma   : a second warp keeps cp.async.bulk copies (8 KB each, HBM -> the same ring) in flight
//   +ld    : four warps keep reading a (different) accumulator region with tcgen05.ld
// Reports SM cycles per MMA (clock64) and wall ns per MMA (globaltimer) -- the two differ when the clock drops under load.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
	return uint64_t((a & 0x3FFFFu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(uint32_t m, uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }
__device__ __forceinline__ unsigned long long gtime() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__global__ void __launch_bounds__(192, 1) k(uint32_t n, uint32_t tiles, int with_tma, int with_ld, uint32_t commit_every, int fence_too, const unsigned char* src, size_t src_bytes,
											 unsigned long long* out) {
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t bar_done, bar_tma[24], bar_dummy;
	__shared__ uint32_t s_tmem;
	__shared__ volatile int s_stop;
	unsigned char* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 192 KB
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint32_t i = threadIdx.x; i < 192u * 1024u / 4; i += blockDim.x) {
		reinterpret_cast<uint32_t*>(ring)[i] = 0;
	}
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar_dummy)));
		for (int i = 0; i < 24; ++i) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma[i])));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		s_stop = 0;
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tm = s_tmem;
	if (warp == 0 && lane == 0) {
		const uint32_t idesc = idesc_bf16(128, n);
		const uint32_t rows_bytes = n * 128;  // one K chunk of a tile
		const long long c0 = clock64();
		const unsigned long long g0 = gtime();
		uint32_t off = 0;
		for (uint32_t t = 0; t < tiles; ++t) {
			for (uint32_t kc = 0; kc < 12; ++kc) {
#pragma unroll
				for (uint32_t kk = 0; kk < 4; ++kk) {
					asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tm + 384),
								 "r"(tm + kc * 32 + kk * 8), "l"(desc_sw128(smem_u32(ring + off) + kk * 32)), "r"(idesc), "r"(uint32_t((kc | kk) != 0))
								 : "memory");
				}
				off += rows_bytes;
				if (off + rows_bytes > 192u * 1024u) {
					off = 0;
				}
				if (commit_every && (kc + 1) % commit_every == 0) {
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_dummy)) : "memory");
					if (fence_too) {
						uint32_t ok;
						asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar_dummy)) : "memory");
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					}
				}
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
		asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar_done))
					 : "memory");
		const long long c1 = clock64();
		const unsigned long long g1 = gtime();
		s_stop = 1;
		if (blockIdx.x == 0) {
			out[0] = (unsigned long long)(c1 - c0);
			out[1] = g1 - g0;
		}
	} else if (warp == 1 && lane == 0 && with_tma) {
		// 24 x 8 KB slots, refilled as fast as they complete (the data is never used: only the traffic matters)
		size_t pos = size_t(blockIdx.x) * 8192 * 24;
		uint32_t phase[24] = {0};
		for (int i = 0; i < 24; ++i) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma[i])), "r"(8192) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + i * 8192)),
						 "l"(src + (pos % src_bytes)), "r"(8192), "r"(smem_u32(&bar_tma[i]))
						 : "memory");
			pos += 8192;
		}
		while (!s_stop) {
			for (int i = 0; i < 24 && !s_stop; ++i) {
				uint32_t ok = 0;
				asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
							 : "=r"(ok)
							 : "r"(smem_u32(&bar_tma[i])), "r"(phase[i])
							 : "memory");
				if (ok) {
					phase[i] ^= 1;
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma[i])), "r"(8192) : "memory");
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
									 smem_u32(ring + i * 8192)),
								 "l"(src + (pos % src_bytes)), "r"(8192), "r"(smem_u32(&bar_tma[i]))
								 : "memory");
					pos += size_t(gridDim.x) * 8192;
				}
			}
		}
		for (int i = 0; i < 24; ++i) {  // drain
			asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(
							 smem_u32(&bar_tma[i])),
						 "r"(phase[i])
						 : "memory");
		}
	} else if (warp >= 2 && with_ld) {
		uint32_t acc = 0;
		while (!s_stop) {
			uint32_t r[32];
			asm volatile(
				"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
				: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
				  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
				  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
				  "=r"(r[31])
				: "r"(tm + 384 + (((warp - 2) * 32) << 16)));
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			for (int i = 0; i < 32; ++i) {
				acc += r[i];
			}
			__nanosleep(with_ld);
		}
		if (acc == 0x12345678u) {
			out[7] = acc;
		}
	}
	__syncthreads();
	if (warp == 0) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
	}
}

int main() {
	unsigned long long* d;
	cudaMalloc(&d, 64);
	unsigned char* src;
	const size_t src_bytes = size_t(8) << 30;
	cudaMalloc(&src, src_bytes + (1 << 20));
	cudaMemset(src, 0, src_bytes + (1 << 20));
	const size_t smem = 1024 + 192 * 1024;
	cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	const uint32_t tiles = 4000;
	for (uint32_t n : {64u, 128u}) {
		for (uint32_t ce : {0u, 12u, 4u, 2u, 1u}) {
			for (int fence : {0, 1}) {
				if (!ce && fence) {
					continue;
				}
				unsigned long long h[2] = {0, 0};
				for (int it = 0; it < 2; ++it) {
					k<<<148, 192, smem>>>(n, tiles, 1, 1000, ce, fence, src, src_bytes, d);
					cudaError_t e = cudaDeviceSynchronize();
					if (e != cudaSuccess) {
						printf("error %s\n", cudaGetErrorString(e));
						return 1;
					}
				}
				cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
				const double mmas = double(tiles) * 48;
				printf("N=%3u commit every %2u K chunks (%2u MMAs)%s: %.1f cycles/MMA, %.1f ns/MMA, tile every %.2f us\n", n, ce, ce * 4,
					   fence ? " + wait + fence" : "               ", h[0] / mmas, h[1] / mmas, h[1] / double(tiles) / 1000.0);
			}
		}
	}
	return 0;
}
