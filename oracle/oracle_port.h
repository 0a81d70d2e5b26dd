/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
 *
 * Plain-C restatement of the reference's CPU algorithm for the float_vector brute-force KNN hot path
 * (and, in ft_port.c, the ft_fast BM25 merge).  Every function cites the reference file:line it follows.
 * Pinned against the reference's own code (oracle/_ref, built in place from /root/reference) by
 * tests/test_oracle_pin.py and against the committed fixtures under tests/golden/.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PORT_L2 = 0, PORT_IP = 1, PORT_COS = 2 }; /* reindexer::VectorMetric, cpp_src/core/enums.h:101 */

/* ---- synthetic data (our own definition, shared bit-for-bit with the device generator in csrc/synth.cuh) */
float port_synth_value(uint64_t seed, uint64_t index);
void port_synth_fill(uint64_t seed, uint64_t first_index, uint64_t count, float* out);

/* ---- distance primitives: scalar forms, cpp_src/tools/distances/l2_dist.cc:13-28, ip_dist.cc:11-23 */
float port_l2sqr(const float* a, const float* b, size_t d);
float port_ip(const float* a, const float* b, size_t d);
/* cpp_src/tools/normalize.cc:10-23 and :25-32 / normalize.h:16-22 */
float port_calc_l2_module(const float* x, int32_t d);
float port_normalize_copy(const float* x, int32_t d, float* out);

/* ---- hnswlib::BruteforceSearch restated (cpp_src/core/index/float_vector/hnswlib/bruteforce.{h,cc}) */
typedef struct port_bf port_bf;
port_bf* port_bf_create(int metric, size_t dim, size_t capacity);
port_bf* port_bf_clone(const port_bf* src, size_t new_capacity);
void port_bf_destroy(port_bf*);
size_t port_bf_size(const port_bf*);
size_t port_bf_capacity(const port_bf*);
size_t port_bf_element_size(const port_bf*);
int port_bf_add(port_bf*, const float* vec, uint64_t label);      /* 0 ok, 1 = capacity exceeded */
int port_bf_remove(port_bf*, uint64_t label);                      /* swap-with-last */
int port_bf_resize(port_bf*, size_t new_capacity);                 /* 1 = smaller than current size */
const float* port_bf_get(const port_bf*, uint64_t label);          /* NULL = label not found */
/* results best-first (the reference's heap drained into slots n-1..0); map-space sign convention */
int64_t port_bf_search_knn(const port_bf*, const float* query, size_t k, float* dists, uint64_t* labels);
int64_t port_bf_search_range(const port_bf*, const float* query, float radius, size_t max_out, float* dists, uint64_t* labels);

/* ---- HnswIndexBase<Map>::search + select / selectRaw post-processing
 *      (cpp_src/core/index/float_vector/hnsw_index.cc:160-191, 194-203, 206-229, 232-288; float_vector_index.h:141-160) */
typedef struct {
	int metric;       /* PORT_L2 / PORT_IP / PORT_COS */
	int need_sort;    /* KnnCtx::NeedSort(): no explicit ORDER BY */
	int is_array;     /* index over an array field: dedup by rowId */
	int raw;          /* selectRaw(): no tie sort */
	int has_k;        /* params.K() set */
	size_t k;
	int has_radius;   /* params.Radius() or index-level radius set */
} port_select_opts;
/* in: n results best-first in map space; out: row ids (label >> 32) and user-visible ranks; returns count */
size_t port_select_postprocess(const port_select_opts*, size_t n, const float* dists, const uint64_t* labels, int32_t* row_ids,
							   float* ranks);

#ifdef __cplusplus
}
#endif
#endif
