// SQ8 state attached to an index (sq8.cu) and the pieces hnsw.cu shares with it.
#pragma once
#include <mutex>

#include "internal.h"
#include "common.cuh"

struct rxgpu_sq8_device {
	rxgpu_sq8_params params{};
	uint32_t code_pitch = 0;  // bytes per row, multiple of 16, zero padded
	uint64_t n = 0;
	uint64_t index_version = 0;
	rxgpu::DevBuf<uint8_t> codes;  // [n][code_pitch]
	rxgpu::DevBuf<float> corr;     // [n] corrective offsets (DistCalculator::correctiveOffsets_)
	std::mutex mtx;                // one SQ8 batch at a time per index (scratch below)
	rxgpu::DevBuf<uint8_t> d_q;
	rxgpu::DevBuf<float> d_qcorr, d_qcoef, d_dist;
	rxgpu::DevBuf<uint64_t> d_lists, d_label;
	rxgpu::DevBuf<uint32_t> d_idx, d_count;
};

namespace rxgpu {
float sq8QuantizeHost(const rxgpu_sq8_device* s, int metric, uint32_t dim, const float* v, float scale, uint8_t* codes);

// in-place selection of the best `want` keys of arr[0, total) into arr[0, want) (ascending); one warp (as knn_scan.cuh:warp_select)
__device__ __forceinline__ void warp_select_keys(uint64_t* arr, uint32_t total, uint32_t want, int lane) {
	for (uint32_t r = 0; r < want; ++r) {
		uint64_t best = kKeyNone;
		uint32_t bpos = r;
		for (uint32_t i = r + lane; i < total; i += 32) {
			const uint64_t kx = arr[i];
			if (kx < best) {
				best = kx;
				bpos = i;
			}
		}
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			const uint64_t ok = __shfl_xor_sync(0xffffffffu, best, off);
			const uint32_t op = __shfl_xor_sync(0xffffffffu, bpos, off);
			if (ok < best || (ok == best && op < bpos)) {
				best = ok;
				bpos = op;
			}
		}
		if (lane == 0 && bpos != r) {
			const uint64_t tmp = arr[r];
			arr[r] = best;
			arr[bpos] = tmp;
		}
		__syncwarp();
	}
}
}  // namespace rxgpu
