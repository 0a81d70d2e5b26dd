// Micro-benchmark (not product, not a test): issue rate of tcgen05.mma.kind::f16 for the operand shapes the filter kernels use.
//   A from tensor memory (TS) or shared memory (SS), M = 128, N in {64, 128, 256}, K = 16, one CTA per SM, one issuing thread.
// Prints cycles per MMA and the implied fraction of the dense bf16 peak (4096 MAC / cycle / SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
	uint64_t d = 0;
	d |= uint64_t((a & 0x3FFFFu) >> 4);
	d |= uint64_t(1024 >> 4) << 32;
	d |= uint64_t(1) << 46;
	d |= uint64_t(2) << 61;
	return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(uint32_t m, uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }

template <bool kTS>
__global__ void __launch_bounds__(128, 1) rate(uint32_t n, uint32_t reps, uint32_t kblocks, unsigned long long* out) {
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t s_tmem;
	unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
	for (uint32_t i = threadIdx.x; i < (64u + 16u) * 1024u / 4; i += blockDim.x) {
		reinterpret_cast<uint32_t*>(base)[i] = 0;
	}
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tm = s_tmem;
	if (threadIdx.x == 0) {
		const uint32_t idesc = idesc_bf16(128, n);
		const uint32_t b_addr = smem_u32(base);              // B: up to 256 rows x 64 K (32 KB), SW128
		const uint32_t a_addr = smem_u32(base + 64 * 1024);  // A (SS): 128 rows x 64 K (16 KB)
		const uint32_t tmem_d = tm + 256;                    // D: up to 256 columns
		const long long t0 = clock64();
		for (uint32_t r = 0; r < reps; ++r) {
			for (uint32_t kb = 0; kb < kblocks; ++kb) {  // kblocks x 4 MMAs accumulate into D like one K block of the real kernels
#pragma unroll
				for (uint32_t k = 0; k < 4; ++k) {
					const uint64_t bd = desc_sw128(b_addr + k * 32);
					const uint32_t acc = (kb | k) != 0;
					if constexpr (kTS) {
						asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
									 "r"(tm + (kb % 24) * 32 + k * 8), "l"(bd), "r"(idesc), "r"(acc)
									 : "memory");
					} else {
						const uint64_t ad = desc_sw128(a_addr + k * 32);
						asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad),
									 "l"(bd), "r"(idesc), "r"(acc)
									 : "memory");
					}
				}
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile(
			"{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar))
			: "memory");
		const long long t1 = clock64();
		if (blockIdx.x == 0) {
			out[0] = (unsigned long long)(t1 - t0);
		}
	}
	__syncthreads();
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
	}
}

int main() {
	unsigned long long* d;
	cudaMalloc(&d, 8);
	const size_t smem = 1024 + 80 * 1024;
	cudaFuncSetAttribute(rate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	cudaFuncSetAttribute(rate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	const uint32_t reps = 200, kblocks = 12;
	for (int ts = 1; ts >= 0; --ts) {
		for (uint32_t n : {32u, 64u, 96u, 128u, 192u, 256u}) {
			for (int grid : {1, 148}) {
				unsigned long long h = 0;
				for (int it = 0; it < 2; ++it) {
					if (ts) {
						rate<true><<<grid, 128, smem>>>(n, reps, kblocks, d);
					} else {
						rate<false><<<grid, 128, smem>>>(n, reps, kblocks, d);
					}
					cudaError_t e = cudaDeviceSynchronize();
					if (e != cudaSuccess) {
						printf("error %s\n", cudaGetErrorString(e));
						return 1;
					}
				}
				cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
				const double per = double(h) / (double(reps) * kblocks * 4);
				printf("%s M=128 N=%3u grid=%3d: %.1f cycles/MMA  -> %.0f MAC/cycle/SM (%.0f%% of 4096)\n", ts ? "TS" : "SS", n, grid, per,
					   128.0 * n * 16 / per, 100.0 * 128.0 * n * 16 / per / 4096.0);
			}
		}
	}
	return 0;
}
