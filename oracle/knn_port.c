/* ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle_port.h.  Never part of the product path.
 *
 * Plain-C restatement of the reference's brute-force float_vector KNN:
 *   distances        cpp_src/tools/distances/l2_dist.cc:13-28 (scalar L2Sqr), ip_dist.cc:11-23 (scalar InnerProduct)
 *   norms            cpp_src/tools/normalize.cc:10-23 (calculateL2Module), :25-32 (normalizeVector), normalize.h:16-22
 *   metric wrapper   cpp_src/core/index/float_vector/hnswlib/hnswlib.h:147-165,192-197 (DistCalculator<float>)
 *   map              cpp_src/core/index/float_vector/hnswlib/bruteforce.cc:44-64 (add), :70-86 (remove), :88-101 (resize),
 *                    :103-127 (SearchKnn), :129-143 (SearchRange); heap = priority_queue.h with std::less<pair<float,u64>>
 *   select           cpp_src/core/index/float_vector/hnsw_index.cc:194-203, 206-229, 232-288; float_vector_index.h:141-160
 * Summation is strictly sequential fp32 (no reassociation: built with -ffp-contract=off -fno-fast-math), which is ONE of the
 * orders the reference itself may use (its SSE/AVX/AVX-512 variants all differ) -- hence distances are
 * compared at 1e-4 relative while ids/order are compared exactly wherever distances are separated by more than fp noise. */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ synthetic data (ours; mirrored by csrc/synth.cuh) */
static inline uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
/* sum of four 16-bit uniforms (Irwin-Hall, exact in integers) scaled to sigma = 0.25 like the reference's test generator
 * N(0, 0.25) (cpp_src/gtests/tools.h:120-129); one fp32 multiply => identical bits on host and device. */
float port_synth_value(uint64_t seed, uint64_t index) {
	const uint64_t h = mix64(seed ^ (index * 0xD1342543DE82EF95ull));
	const int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) + (int32_t)(h >> 48);
	return (float)(s - 131070) * 6.6072488e-06f; /* 0.25 / 37837.23 */
}
void port_synth_fill(uint64_t seed, uint64_t first_index, uint64_t count, float* out) {
	for (uint64_t i = 0; i < count; ++i) {
		out[i] = port_synth_value(seed, first_index + i);
	}
}

/* ------------------------------------------------------------------ distances */
float port_l2sqr(const float* a, const float* b, size_t d) { /* l2_dist.cc:13-28 */
	float res = 0.f;
	for (size_t i = 0; i < d; ++i) {
		const float t = a[i] - b[i];
		res += t * t;
	}
	return res;
}
float port_ip(const float* a, const float* b, size_t d) { /* ip_dist.cc:11-23 */
	float res = 0.f;
	for (size_t i = 0; i < d; ++i) {
		res += a[i] * b[i];
	}
	return res;
}
float port_calc_l2_module(const float* x, int32_t d) { /* normalize.cc:10-23 */
	float normL2Sqr = 0.f;
	for (int32_t i = 0; i < d; ++i) {
		normL2Sqr += x[i] * x[i];
	}
	float normL2K = 1.f;
	if (normL2Sqr > 0.f && fabsf(1.0f - normL2Sqr) > 0.00001f) {
		normL2K = (float)(1.0 / sqrtf(normL2Sqr)); /* `1.0 / std::sqrt(float)` : sqrt in fp32, divide in fp64, round to fp32 */
	}
	return normL2K;
}
float port_normalize_copy(const float* x, int32_t d, float* out) { /* normalize.h:16-20 + normalize.cc:25-32 */
	memcpy(out, x, (size_t)d * sizeof(float));
	const float k = port_calc_l2_module(out, d);
	for (int32_t i = 0; i < d; ++i) {
		out[i] *= k;
	}
	return k;
}

/* ------------------------------------------------------------------ label -> idx map (stand-in for tsl::hopscotch_sc_map) */
typedef struct {
	uint64_t* keys;
	uint32_t* vals;
	uint8_t* used;
	size_t cap; /* power of two */
	size_t n;
} lmap;
static size_t lmap_slot(const lmap* m, uint64_t key) { return (size_t)(mix64(key) & (m->cap - 1)); }
static void lmap_init(lmap* m, size_t cap) {
	size_t c = 16;
	while (c < cap * 2) {
		c <<= 1;
	}
	m->cap = c;
	m->n = 0;
	m->keys = (uint64_t*)calloc(c, sizeof(uint64_t));
	m->vals = (uint32_t*)calloc(c, sizeof(uint32_t));
	m->used = (uint8_t*)calloc(c, 1);
}
static void lmap_free(lmap* m) {
	free(m->keys);
	free(m->vals);
	free(m->used);
}
static int64_t lmap_find(const lmap* m, uint64_t key) {
	for (size_t s = lmap_slot(m, key); m->used[s]; s = (s + 1) & (m->cap - 1)) {
		if (m->keys[s] == key) {
			return (int64_t)s;
		}
	}
	return -1;
}
static void lmap_put(lmap* m, uint64_t key, uint32_t val);
static void lmap_grow(lmap* m) {
	lmap old = *m;
	lmap_init(m, old.cap);
	for (size_t s = 0; s < old.cap; ++s) {
		if (old.used[s]) {
			lmap_put(m, old.keys[s], old.vals[s]);
		}
	}
	lmap_free(&old);
}
static void lmap_put(lmap* m, uint64_t key, uint32_t val) {
	if ((m->n + 1) * 2 > m->cap) {
		lmap_grow(m);
	}
	size_t s = lmap_slot(m, key);
	while (m->used[s] && m->keys[s] != key) {
		s = (s + 1) & (m->cap - 1);
	}
	if (!m->used[s]) {
		m->used[s] = 1;
		m->keys[s] = key;
		m->n++;
	}
	m->vals[s] = val;
}
static void lmap_erase(lmap* m, uint64_t key) { /* backward-shift deletion */
	int64_t f = lmap_find(m, key);
	if (f < 0) {
		return;
	}
	size_t hole = (size_t)f;
	m->used[hole] = 0;
	m->n--;
	for (size_t s = (hole + 1) & (m->cap - 1); m->used[s]; s = (s + 1) & (m->cap - 1)) {
		const size_t home = lmap_slot(m, m->keys[s]);
		const int between = hole <= s ? (home > hole && home <= s) : (home > hole || home <= s);
		if (!between) {
			m->keys[hole] = m->keys[s];
			m->vals[hole] = m->vals[s];
			m->used[hole] = 1;
			m->used[s] = 0;
			hole = s;
		}
	}
}

/* ------------------------------------------------------------------ BruteforceSearch */
struct port_bf {
	int metric;
	size_t dim;
	size_t max_elements;
	size_t cur;
	float* rows;      /* max_elements x dim (the reference interleaves the label after each row: bruteforce.h:47-48) */
	uint64_t* labels; /* max_elements */
	float* norm_coefs; /* Cosine only: hnswlib.h:33-35 */
	lmap dict;
};

port_bf* port_bf_create(int metric, size_t dim, size_t capacity) {
	port_bf* b = (port_bf*)calloc(1, sizeof(port_bf));
	b->metric = metric;
	b->dim = dim;
	b->max_elements = capacity;
	b->rows = (float*)malloc((capacity ? capacity : 1) * dim * sizeof(float));
	b->labels = (uint64_t*)malloc((capacity ? capacity : 1) * sizeof(uint64_t));
	b->norm_coefs = metric == PORT_COS ? (float*)calloc(capacity ? capacity : 1, sizeof(float)) : NULL;
	lmap_init(&b->dict, capacity);
	return b;
}
port_bf* port_bf_clone(const port_bf* src, size_t new_capacity) { /* bruteforce.cc:20-34 */
	const size_t cap = src->max_elements > new_capacity ? src->max_elements : new_capacity;
	port_bf* b = port_bf_create(src->metric, src->dim, cap);
	memcpy(b->rows, src->rows, src->cur * src->dim * sizeof(float));
	memcpy(b->labels, src->labels, src->cur * sizeof(uint64_t));
	if (b->norm_coefs) {
		memcpy(b->norm_coefs, src->norm_coefs, src->cur * sizeof(float));
	}
	b->cur = src->cur;
	for (size_t i = 0; i < src->cur; ++i) {
		lmap_put(&b->dict, src->labels[i], (uint32_t)i);
	}
	return b;
}
void port_bf_destroy(port_bf* b) {
	if (!b) {
		return;
	}
	free(b->rows);
	free(b->labels);
	free(b->norm_coefs);
	lmap_free(&b->dict);
	free(b);
}
size_t port_bf_size(const port_bf* b) { return b->cur; }
size_t port_bf_capacity(const port_bf* b) { return b->max_elements; }
size_t port_bf_element_size(const port_bf* b) { return b->dim * sizeof(float) + sizeof(uint64_t); } /* bruteforce.h:47-48 */

int port_bf_add(port_bf* b, const float* vec, uint64_t label) { /* bruteforce.cc:44-64 */
	size_t idx;
	const int64_t f = lmap_find(&b->dict, label);
	if (f >= 0) {
		idx = b->dict.vals[f];
	} else {
		if (b->cur >= b->max_elements) {
			return 1; /* "The number of elements exceeds the specified limit" */
		}
		idx = b->cur;
		lmap_put(&b->dict, label, (uint32_t)idx);
		b->cur++;
	}
	if (b->norm_coefs) {
		b->norm_coefs[idx] = port_calc_l2_module(vec, (int32_t)b->dim); /* DistCalculator::AddNorm hnswlib.h:80-92 */
	}
	memcpy(b->rows + idx * b->dim, vec, b->dim * sizeof(float));
	b->labels[idx] = label;
	return 0;
}
int port_bf_remove(port_bf* b, uint64_t label) { /* bruteforce.cc:70-86 */
	const int64_t f = lmap_find(&b->dict, label);
	if (f < 0) {
		return 0;
	}
	const size_t cur_c = b->dict.vals[f];
	lmap_erase(&b->dict, label);
	if (cur_c + 1 != b->cur) {
		const size_t last = b->cur - 1;
		lmap_put(&b->dict, b->labels[last], (uint32_t)cur_c);
		memcpy(b->rows + cur_c * b->dim, b->rows + last * b->dim, b->dim * sizeof(float));
		b->labels[cur_c] = b->labels[last];
		if (b->norm_coefs) {
			b->norm_coefs[cur_c] = b->norm_coefs[last]; /* MoveNorm */
		}
	}
	b->cur--;
	return 0;
}
int port_bf_resize(port_bf* b, size_t new_capacity) { /* bruteforce.cc:88-101 */
	if (new_capacity < b->cur) {
		return 1; /* "Cannot resize, max element is less than the current number of elements" */
	}
	const size_t c = new_capacity ? new_capacity : 1;
	b->rows = (float*)realloc(b->rows, c * b->dim * sizeof(float));
	b->labels = (uint64_t*)realloc(b->labels, c * sizeof(uint64_t));
	if (b->norm_coefs) {
		b->norm_coefs = (float*)realloc(b->norm_coefs, c * sizeof(float));
	}
	b->max_elements = new_capacity;
	return 0;
}
const float* port_bf_get(const port_bf* b, uint64_t label) { /* bruteforce.cc:36-42 */
	const int64_t f = lmap_find(&b->dict, label);
	return f < 0 ? NULL : b->rows + (size_t)b->dict.vals[f] * b->dim;
}

/* DistCalculator<float>::operator()(q, row, id): hnswlib.h:147-165 with l2/ip of :192-197 (alpha2 = 1, corrective offsets = 0) */
static float port_dist(const port_bf* b, const float* q, size_t idx) {
	const float* row = b->rows + idx * b->dim;
	float dist;
	if (b->metric == PORT_L2) {
		dist = 1.f * port_l2sqr(q, row, b->dim) + 0.f + 0.f;
	} else {
		dist = -(1.f * port_ip(q, row, b->dim) + 0.f + 0.f);
	}
	if (b->metric == PORT_COS) {
		dist *= b->norm_coefs[idx];
	}
	return dist;
}

/* max-heap over (dist, label) under std::less<std::pair<float,uint64_t>> (hnsw_interface.h:14, priority_queue.h) */
typedef struct {
	float d;
	uint64_t l;
} dl_pair;
static int pair_less(dl_pair a, dl_pair b) { return a.d < b.d || (!(b.d < a.d) && a.l < b.l); }
static void sift_down(dl_pair* h, size_t n, size_t i) {
	for (;;) {
		size_t big = i, l = 2 * i + 1, r = l + 1;
		if (l < n && pair_less(h[big], h[l])) {
			big = l;
		}
		if (r < n && pair_less(h[big], h[r])) {
			big = r;
		}
		if (big == i) {
			return;
		}
		dl_pair t = h[i];
		h[i] = h[big];
		h[big] = t;
		i = big;
	}
}
static void sift_up(dl_pair* h, size_t i) {
	while (i > 0) {
		const size_t p = (i - 1) / 2;
		if (!pair_less(h[p], h[i])) {
			return;
		}
		dl_pair t = h[i];
		h[i] = h[p];
		h[p] = t;
		i = p;
	}
}
/* pop everything: slots n-1..0, as HnswIndexBase::select drains the queue (hnsw_index.cc:258-276) */
static void heap_drain(dl_pair* h, size_t n, size_t max_out, float* dists, uint64_t* labels) {
	for (size_t i = n; i > 0;) {
		--i;
		if (i < max_out) {
			dists[i] = h[0].d;
			labels[i] = h[0].l;
		}
		h[0] = h[i];
		sift_down(h, i, 0);
	}
}

int64_t port_bf_search_knn(const port_bf* b, const float* query, size_t k, float* dists, uint64_t* labels) { /* bruteforce.cc:103-127 */
	if (b->cur == 0 || k == 0) {
		return 0;
	}
	if (k > b->cur) {
		k = b->cur;
	}
	dl_pair* heap = (dl_pair*)malloc(k * sizeof(dl_pair));
	for (size_t i = 0; i < k; ++i) { /* first k pushed unconditionally */
		heap[i].d = port_dist(b, query, i);
		heap[i].l = b->labels[i];
		sift_up(heap, i);
	}
	float lastdist = heap[0].d;
	for (size_t i = k; i < b->cur; ++i) {
		const float dist = port_dist(b, query, i);
		if (dist < lastdist) { /* strict */
			heap[0].d = dist;
			heap[0].l = b->labels[i];
			sift_down(heap, k, 0);
			lastdist = heap[0].d;
		}
	}
	heap_drain(heap, k, k, dists, labels);
	free(heap);
	return (int64_t)k;
}

int64_t port_bf_search_range(const port_bf* b, const float* query, float radius, size_t max_out, float* dists,
							 uint64_t* labels) { /* bruteforce.cc:129-143 */
	size_t n = 0, cap = 64;
	dl_pair* heap = (dl_pair*)malloc(cap * sizeof(dl_pair));
	for (size_t i = 0; i < b->cur; ++i) {
		const float dist = port_dist(b, query, i);
		if (dist < radius) {
			if (n == cap) {
				cap *= 2;
				heap = (dl_pair*)realloc(heap, cap * sizeof(dl_pair));
			}
			heap[n].d = dist;
			heap[n].l = b->labels[i];
			sift_up(heap, n);
			n++;
		}
	}
	heap_drain(heap, n, max_out, dists, labels);
	free(heap);
	return (int64_t)n;
}

/* ------------------------------------------------------------------ select post-processing */
static int cmp_i32(const void* a, const void* b) {
	const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
	return (x > y) - (x < y);
}
size_t port_select_postprocess(const port_select_opts* o, size_t n, const float* dists, const uint64_t* labels, int32_t* row_ids,
							   float* ranks) {
	if (n == 0) {
		return 0;
	}
	/* hnsw_index.cc:258-276 (select) / :209-222 (selectRaw): walk i = n-1 .. 0 exactly as the heap pops worst-first */
	size_t lastSameDist = n - 1;
	for (size_t i = n; i > 0;) {
		--i;
		ranks[i] = o->metric == PORT_L2 ? dists[i] : -dists[i]; /* sign flip for IP / Cosine :261-270 */
		row_ids[i] = (int32_t)(labels[i] >> 32);                 /* FloatVectorId::RowId() */
		if (!o->raw && o->need_sort) {                           /* sortSameDist lambda :240-257 */
			const int newDist = o->metric == PORT_L2 ? (ranks[lastSameDist] > ranks[i]) : (ranks[lastSameDist] < ranks[i]);
			if (newDist) {
				qsort(row_ids + i + 1, lastSameDist - i, sizeof(int32_t), cmp_i32);
				lastSameDist = i;
			}
		}
	}
	if (!o->raw && o->need_sort) {
		qsort(row_ids, lastSameDist + 1, sizeof(int32_t), cmp_i32); /* :274-276 */
	}
	size_t cnt = n;
	if (o->is_array) { /* removeDuplicateRowId, float_vector_index.h:141-160: keep first (best) occurrence */
		size_t to = 0;
		for (size_t from = 0; from < n; ++from) {
			int seen = 0;
			for (size_t j = 0; j < to; ++j) {
				if (row_ids[j] == row_ids[from]) {
					seen = 1;
					break;
				}
			}
			if (!seen) {
				row_ids[to] = row_ids[from];
				ranks[to] = ranks[from];
				++to;
			}
		}
		cnt = to;
	}
	if (o->has_k && o->has_radius && cnt > o->k) { /* removeOverK :194-203 */
		cnt = o->k;
	}
	return cnt;
}
