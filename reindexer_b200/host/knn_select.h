// Host-side (C++) pieces of the float_vector KNN path that sit between the device scan and the reference's callers.
// Product code: no dependency on oracle/.  Each function names the reference code whose behaviour it reproduces
// (paths relative to /root/reference/cpp_src).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <unordered_set>
#include <vector>

namespace rxgpu {

struct Hit {
	float dist;     // map-space distance (smaller is better)
	uint64_t gidx;  // global internal row index (insertion order with swap-deletes)
	uint64_t label; // FloatVectorId::AsNumber()
};

// std::less<std::pair<float, labeltype>> -- the comparator of the reference's result heap
// (core/index/float_vector/hnswlib/hnsw_interface.h:14, bruteforce.cc:114)
inline bool hitLessByLabel(const Hit& a, const Hit& b) noexcept { return a.dist < b.dist || (!(b.dist < a.dist) && a.label < b.label); }
inline bool hitLessByIndex(const Hit& a, const Hit& b) noexcept { return a.dist < b.dist || (!(b.dist < a.dist) && a.gidx < b.gidx); }

// tools/normalize.cc:10-23 (calculateL2Module): 1/sqrt(sum x^2) with the "already normalised" shortcut
inline float calculateL2Module(const float* x, int32_t d) noexcept {
	float normL2Sqr = 0.f;
	for (int32_t i = 0; i < d; ++i) {
		normL2Sqr += x[i] * x[i];
	}
	float normL2K = 1.f;
	if (normL2Sqr > 0.f && std::abs(1.0f - normL2Sqr) > 0.00001f) {
		normL2K = float(1.0 / std::sqrt(normL2Sqr));
	}
	return normL2K;
}
// tools/normalize.h:16-20 (NormalizeCopyVector)
inline float normalizeCopyVector(const float* x, int32_t d, float* out) noexcept {
	std::memcpy(out, x, size_t(d) * sizeof(float));
	const float k = calculateL2Module(out, d);
	for (int32_t i = 0; i < d; ++i) {
		out[i] *= k;
	}
	return k;
}

// The reference's heap tie rule in closed form (bruteforce.cc:103-127; derivation in DESIGN.md §tie rule).
//   lower : the m < k rows with dist < dstar
//   first : the first min(k, #) rows, in global internal order, with dist <= dstar
// Let S = first.  Every row of S is accepted by the reference's heap (it is either pushed during the initial fill or
// replaces a top that is > dstar).  Once the heap holds k rows <= dstar, later ties are rejected (strict <) and each later
// `lower` row evicts the heap top = the tie with the largest label.  Hence the survivors are
//   lower  U  (ties in S) minus the E largest labels,  E = #{lower rows not in S}.
// Result: best-first, equal distances by ascending label (the order in which the max-heap drains backwards).
inline std::vector<Hit> tieReplay(uint32_t k, float dstar, const std::vector<Hit>& lower, const std::vector<Hit>& first) {
	std::vector<Hit> ties;
	std::vector<uint64_t> firstIdx;
	firstIdx.reserve(first.size());
	for (const Hit& h : first) {
		firstIdx.push_back(h.gidx);
		if (!(h.dist < dstar)) {
			ties.push_back(h);
		}
	}
	std::sort(firstIdx.begin(), firstIdx.end());
	size_t evict = 0;
	for (const Hit& h : lower) {
		if (!std::binary_search(firstIdx.begin(), firstIdx.end(), h.gidx)) {
			++evict;
		}
	}
	std::sort(ties.begin(), ties.end(), [](const Hit& a, const Hit& b) noexcept { return a.label < b.label; });
	ties.resize(ties.size() > evict ? ties.size() - evict : 0);
	std::vector<Hit> out(lower);
	out.insert(out.end(), ties.begin(), ties.end());
	std::sort(out.begin(), out.end(), hitLessByLabel);
	if (out.size() > k) {
		out.resize(k);
	}
	return out;
}

// The device returns rows ordered by (dist, internal index); the reference drains its heap by (dist, label).
inline void orderTiesByLabel(std::vector<Hit>& hits) { std::stable_sort(hits.begin(), hits.end(), hitLessByLabel); }

struct SelectParams {
	int metric = 0;  // 0 L2, 1 IP, 2 Cosine
	bool needSort = true;
	bool isArray = false;
	bool raw = false;
	bool hasK = false;
	size_t k = 0;
	bool hasRadius = false;
};

// HnswIndexBase<Map>::select / selectRaw post-processing (core/index/float_vector/hnsw_index.cc:206-229, 232-288),
// removeDuplicateRowId (core/index/float_vector/float_vector_index.h:141-160), removeOverK (hnsw_index.cc:194-203).
// in: results best-first in map space.  out: row ids (label >> 32) and user-visible ranks.
inline void selectPostprocess(const SelectParams& p, const std::vector<Hit>& res, std::vector<int32_t>& rowIds, std::vector<float>& ranks) {
	const size_t n = res.size();
	rowIds.assign(n, 0);
	ranks.assign(n, 0.f);
	if (n == 0) {
		return;
	}
	size_t lastSameDist = n - 1;
	for (size_t i = n; i > 0;) {  // the reference pops worst-first and fills slot i = n-1 ... 0
		--i;
		ranks[i] = p.metric == 0 ? res[i].dist : -res[i].dist;
		rowIds[i] = int32_t(res[i].label >> 32);
		if (!p.raw && p.needSort) {
			const bool newDist = p.metric == 0 ? (ranks[lastSameDist] > ranks[i]) : (ranks[lastSameDist] < ranks[i]);
			if (newDist) {
				std::sort(rowIds.begin() + i + 1, rowIds.begin() + lastSameDist + 1);
				lastSameDist = i;
			}
		}
	}
	if (!p.raw && p.needSort) {
		std::sort(rowIds.begin(), rowIds.begin() + lastSameDist + 1);
	}
	if (p.isArray) {
		size_t to = 0;
		std::unordered_set<int32_t> seen;  // the reference dedups through a hash set too (float_vector_index.h:141-160)
		seen.reserve(n);
		for (size_t from = 0; from < n; ++from) {
			if (seen.insert(rowIds[from]).second) {
				rowIds[to] = rowIds[from];
				ranks[to] = ranks[from];
				++to;
			}
		}
		rowIds.resize(to);
		ranks.resize(to);
	}
	if (p.hasK && p.hasRadius && rowIds.size() > p.k) {
		rowIds.resize(p.k);
		ranks.resize(p.k);
	}
}

}  // namespace rxgpu
