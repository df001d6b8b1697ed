/* b200hnsw.h -- C ABI of libb200hnsw.so, the B200-native (sm_100a) replacement for the hnswlib engine
 * vendored in hiozings/Research-New-HNSW.
 *
 * This is the drop-in boundary of SURVEY.md 8(b): the reference's consumers include "hnswlib/hnswlib.h"
 * (index_builder/build.cpp:8, hnsw_service/main.cpp:5, test.cpp:1); our header shim of the same name
 * (research_new_hnsw_b200/hnswlib/hnswlib.h) forwards every engine call to the entry points below.  Each entry
 * point cites the reference interface it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - plain C types only; all pointers are HOST pointers owned by the caller unless the name ends in _device;
 *   - every function returns 0 on success or a negative b200hnsw_status; the message of the last failure on the
 *     calling thread is b200hnsw_last_error(); no C++ exception crosses this boundary;
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with B200HNSW_E_CUDA;
 *   - result rows are closest-first, k entries per query, padded with label = UINT64_MAX, dist = +inf;
 *   - handles are opaque and library-owned; an index handle may be searched from several host threads at once
 *     (reference: searchKnn is const and thread-safe, hnswalg.h:1270, visited_list_pool.h:50-68).
 */
#ifndef B200HNSW_H_
#define B200HNSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200HNSW_ABI_VERSION 2

typedef enum {
    B200HNSW_OK = 0,
    B200HNSW_E_CUDA = -1,      /* CUDA runtime / no device / kernel failure */
    B200HNSW_E_NOMEM = -2,     /* "Not enough memory" (hnswalg.h:128) */
    B200HNSW_E_OPEN = -3,      /* "Cannot open file" (hnswalg.h:720) */
    B200HNSW_E_CORRUPT = -4,   /* "Index seems to be corrupted or unsupported" (hnswalg.h:758,768) */
    B200HNSW_E_CAPACITY = -5,  /* "The number of elements exceeds the specified limit" (hnswalg.h:1178) */
    B200HNSW_E_LABEL = -6,     /* "Label not found" (hnswalg.h:833,860) */
    B200HNSW_E_ARG = -7,       /* invalid argument (null pointer, dim mismatch, unsupported dim/ef) */
    B200HNSW_E_CAND = -8,      /* "cand error": neighbour id beyond max_elements (hnswalg.h:1292) */
    B200HNSW_E_STATE = -9,     /* call not valid in this state (e.g. already-deleted label, hnswalg.h:880) */
    B200HNSW_E_UNSUPPORTED = -10
} b200hnsw_status;

typedef enum { B200HNSW_L2 = 0, B200HNSW_IP = 1 } b200hnsw_metric;        /* space_l2.h:207 / space_ip.h:343 */
typedef enum { B200HNSW_F32 = 0, B200HNSW_BF16 = 1 } b200hnsw_storage;    /* device copy of the vectors */

typedef struct b200hnsw_index b200hnsw_index; /* HierarchicalNSW<float> (hnswalg.h:17) */
typedef struct b200bf_index b200bf_index;     /* BruteforceSearch<float> (bruteforce.h:10) */

/* Constructor arguments of HierarchicalNSW (hnswalg.h:89-95) plus what the GPU engine needs to know. */
typedef struct {
    int32_t metric;                /* b200hnsw_metric */
    int32_t storage;               /* b200hnsw_storage */
    int32_t device;                /* CUDA device ordinal; -1 = current device */
    int32_t allow_replace_deleted; /* hnswalg.h:94 */
    uint64_t dim;                  /* space dimension (space_l2.h:230) */
    uint64_t max_elements;         /* hnswalg.h:91; for load: max(arg, file) as in hnswalg.h:732-735 */
    uint64_t M;                    /* hnswalg.h:92, capped at 10000 */
    uint64_t ef_construction;      /* hnswalg.h:93, raised to M */
    uint64_t random_seed;          /* hnswalg.h:94: level generator seed (default 100) */
} b200hnsw_params;

/* Mirror of the public data members consumers read directly: cur_element_count,
 * maxlevel_, enterpoint_node_ (build.cpp:24-36, test.cpp:21-23, main.cpp:56,88) and the header of saveIndex. */
typedef struct {
    uint64_t cur_element_count, max_elements, num_deleted;
    uint64_t dim, M, maxM, maxM0, ef_construction, ef;
    uint64_t size_data_per_element, size_links_per_element, size_links_level0, offset_data, label_offset;
    double mult;
    int32_t maxlevel;
    uint32_t enterpoint_node;
    int32_t metric, storage, device, reserved;
} b200hnsw_info;

/* Work counters of the most recent search/build on this handle (sums over the batch); they are the numerators
 * of SURVEY.md 8(d)'s algorithmic-bytes figure and correspond to metric_distance_computations / metric_hops
 * (hnswalg.h:65-66). */
typedef struct {
    uint64_t queries;
    uint64_t dist_evals;     /* D: vectors read and compared, all layers */
    uint64_t hops_base;      /* H0: level-0 expansions (brute force: 1 if the tensor-core path produced the result) */
    uint64_t hops_upper;     /* Hup: upper-layer list scans */
    uint64_t visited_resets; /* times a per-query visited table was rebuilt (re-evaluations possible, results unchanged) */
    uint64_t kernel_launches;
    double last_kernel_ms;   /* CUDA-event time of the dominant kernel of the last call */
    uint64_t dropped_reverse_edges; /* build: reverse edges not offered to a neighbour because more than 32 new points
                                       of ONE batch selected it (the reference would have re-pruned it once per edge,
                                       hnswalg.h:590-612); 0 on every configuration measured so far */
} b200hnsw_stats;

const char *b200hnsw_last_error(void);
int b200hnsw_abi_version(void);
/* Number of CUDA devices visible, or a negative status. */
int b200hnsw_device_count(void);

/* ---- HierarchicalNSW<float> ------------------------------------------------------------------------------ */
/* Build constructor, hnswalg.h:89-144. */
int b200hnsw_create(const b200hnsw_params *params, b200hnsw_index **out);
/* Load constructor / loadIndex, hnswalg.h:78-86, 716-822.  params supplies metric, dim, storage, device and the
 * optional max_elements; the graph parameters come from the file. Resets ef to 10 (hnswalg.h:795). */
int b200hnsw_load(const char *path, const b200hnsw_params *params, b200hnsw_index **out);
/* saveIndex, hnswalg.h:685-713: byte-identical format. Flushes staged insertions first. */
int b200hnsw_save(b200hnsw_index *h, const char *path);
/* ~HierarchicalNSW / clear(), hnswalg.h:147-162. */
void b200hnsw_destroy(b200hnsw_index *h);
/* setEf, hnswalg.h:173-175: default ef used when a search passes ef = 0. */
int b200hnsw_set_ef(b200hnsw_index *h, size_t ef);
/* addPoint(const void*, labeltype), hnswalg.h:954-964 -> 1153-1267, batched: n rows of dim floats and n labels.
 * Levels, element count, entry point and max level are assigned immediately in row order with the reference's
 * level generator (hnswalg.h:207-211,1187-1198,1255-1265); graph linking may be deferred until b200hnsw_flush
 * (or any call that reads the graph).  labels == NULL means labels cur_element_count .. +n-1.
 * A label that already exists is UPDATED (hnswalg.h:1157-1174 -> updatePoint, :995-1139): new vector, delete mark
 * cleared, every old neighbour re-pruned over the 1-hop + 2-hop set (:1009-1069) and the point re-linked
 * (repairConnectionsForUpdate, :1075-1139), all on the GPU.  Up to 2048 updates per call are applied one after the other
 * like the reference does; larger calls in groups that see the graph as of the start of their group. */
int b200hnsw_add_batch(b200hnsw_index *h, const float *X, const uint64_t *labels, size_t n);
/* addPoint(data, label, replace_deleted = true), hnswalg.h:954-992: while deleted elements exist, each row takes the
 * place of one of them (its label and vector replaced, delete mark cleared, re-linked like an update); otherwise like
 * b200hnsw_add_batch.  Fails with B200HNSW_E_STATE unless the index was created with allow_replace_deleted. */
int b200hnsw_add_batch_replace_deleted(b200hnsw_index *h, const float *X, const uint64_t *labels, size_t n);
/* Links every staged point into the graph on the GPU and refreshes the host mirror. */
int b200hnsw_flush(b200hnsw_index *h);
/* searchKnn, hnswalg.h:1270-1324, batched: nq queries of dim floats; ef = 0 -> the setEf value; the engine uses
 * max(ef, k) as the reference does (hnswalg.h:1309).  labels_out/dists_out are [nq][k]; counts_out (nullable)
 * receives the number of valid results per query; work_out (nullable) receives [nq][4] = {D, H0, Hup, resets}.
 * Limits (B200HNSW_E_UNSUPPORTED beyond them; the reference has none): max(ef, k) <= 4096, dim <= 1024, and while
 * elements are marked deleted or a filter is given at most 2^30 elements (one id bit carries the mark). */
int b200hnsw_search_batch(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef, uint64_t *labels_out,
                          float *dists_out, uint32_t *counts_out, uint32_t *work_out);
/* Asynchronous form for a serving loop that keeps more than one batch in flight: with PAGE-LOCKED Q / labels_out /
 * dists_out / counts_out the launch is enqueued (the kernel reads the queries and stores the rows over PCIe itself) and
 * the call returns a ticket at once; b200hnsw_search_batch_wait blocks until that batch is complete.  The buffers must
 * stay valid and untouched until then.  At most four batches may be in flight per index (one more submit fails with
 * B200HNSW_E_STATE); pageable buffers make submit complete the search synchronously and return ticket 0 (wait(0) is a no-op).
 * Calls that change the index wait for the launches in flight. */
int b200hnsw_search_batch_submit(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef, uint64_t *labels_out,
                                 float *dists_out, uint32_t *counts_out, uint64_t *ticket_out);
int b200hnsw_search_batch_wait(b200hnsw_index *h, uint64_t ticket);
/* Same, with DEVICE pointers on the index's device, enqueued on cuda_stream (a cudaStream_t; NULL = legacy
 * default stream) without host synchronisation. */
int b200hnsw_search_batch_device(b200hnsw_index *h, const float *dQ, size_t nq, size_t k, size_t ef,
                                 uint64_t *d_labels_out, float *d_dists_out, uint32_t *d_counts_out,
                                 uint32_t *d_work_out, void *cuda_stream);
/* searchKnn with a BaseFilterFunctor (hnswlib.h:128-132, hnswalg.h:1270,1306-1313): the functor is a host callback, so
 * the caller evaluates it once per stored label and passes the verdicts -- allowed[i] != 0 for INTERNAL id i,
 * cur_element_count bytes (NULL = no filter).  A node that is not allowed is traversed but never returned, exactly like
 * a deleted one (hnswalg.h:406-407). */
int b200hnsw_search_batch_filtered(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef,
                                   const uint8_t *allowed, uint64_t *labels_out, float *dists_out, uint32_t *counts_out);
/* getExternalLabel for every internal id 0 .. cur_element_count-1 (bulk form of hnswalg.h:186-190). */
int b200hnsw_get_labels(b200hnsw_index *h, uint64_t *labels_out, size_t capacity);
/* Public fields / accessors the consumers touch. */
int b200hnsw_get_info(b200hnsw_index *h, b200hnsw_info *out);
/* element_levels_ (hnswalg.h:52): pointer to cur_element_count ints, valid until the next mutating call. */
int b200hnsw_get_levels(b200hnsw_index *h, const int32_t **levels_out);
/* get_linklist_at_level (hnswalg.h:501-503): pointer to the reference-layout list header (u16 count in the low
 * half-word, neighbours from ptr+1) in the host mirror; flushes staged insertions first. */
int b200hnsw_get_linklist(b200hnsw_index *h, uint32_t internal_id, int level, const uint32_t **ptr_out);
/* getExternalLabel (hnswalg.h:186-190) / getDataByInternalId (hnswalg.h:202-204) */
int b200hnsw_get_label(b200hnsw_index *h, uint32_t internal_id, uint64_t *label_out);
int b200hnsw_get_data(b200hnsw_index *h, uint32_t internal_id, const float **vec_out);
/* getDataByLabel (hnswalg.h:825-847): copies dim floats. */
int b200hnsw_get_data_by_label(b200hnsw_index *h, uint64_t label, float *vec_out);
/* markDelete / unmarkDelete (hnswalg.h:853-917). */
int b200hnsw_mark_delete(b200hnsw_index *h, uint64_t label);
int b200hnsw_unmark_delete(b200hnsw_index *h, uint64_t label);
/* resizeIndex (hnswalg.h:633-656). */
int b200hnsw_resize(b200hnsw_index *h, size_t new_max_elements);
/* indexFileSize (hnswalg.h:658-683). */
int b200hnsw_index_file_size(b200hnsw_index *h, uint64_t *bytes_out);
int b200hnsw_get_stats(b200hnsw_index *h, b200hnsw_stats *out);

/* Merge of per-shard results: in = [shards][nq][k] rows gathered from every rank (device
 * pointers), each row closest first as every search returns it (ascending distance, padding last; ties in any order);
 * out = [nq][k] smallest (dist, label) pairs, closest first.  Enqueued on cuda_stream. */
int b200hnsw_merge_topk_device(const uint64_t *d_labels_in, const float *d_dists_in, size_t shards, size_t nq,
                               size_t k, uint64_t *d_labels_out, float *d_dists_out, void *cuda_stream);

/* Same merge for blocks packed per shard as [nq*k labels (u64) | nq*k dists (f32) | padding] of block_bytes each
 * (block_bytes a multiple of 8): a shard writes its search results straight into its block and ONE all_gather moves
 * both arrays. */
int b200hnsw_merge_topk_packed_device(const void *d_blocks, size_t block_bytes, size_t shards, size_t nq, size_t k,
                                      uint64_t *d_labels_out, float *d_dists_out, void *cuda_stream);

/* ---- sharded index: one process, one sub-index per GPU ---------------------------------------
 * The reference has no sharding; north_star prescribes it: the data set is split into n_shards HierarchicalNSW
 * sub-indexes, shard s on CUDA device devices[s] (devices may repeat), every query is searched on every shard and the
 * per-shard top-k are merged by (dist, label) on devices[0].  Inside one process the exchange needs no collective:
 * with peer access every shard's search kernel stores its rows straight into the root device's packed buffer over
 * NVLink, otherwise the block is copied device to device.  Each shard is an ordinary index: its file is a reference
 * saveIndex file that the reference (or b200hnsw_load) opens on its own.  Labels are global; a label lives on shard
 * label % n_shards, so re-adding a label updates it where it is.  params->max_elements is PER SHARD. */
typedef struct b200hnsw_sharded b200hnsw_sharded;
int b200hnsw_sharded_create(const b200hnsw_params *params, const int *devices, size_t n_shards, b200hnsw_sharded **out);
/* paths[s] = saveIndex file of shard s (hnswalg.h:716-822 per shard). */
int b200hnsw_sharded_load(const char *const *paths, const b200hnsw_params *params, const int *devices, size_t n_shards,
                          b200hnsw_sharded **out);
int b200hnsw_sharded_save(b200hnsw_sharded *h, const char *const *paths);
void b200hnsw_sharded_destroy(b200hnsw_sharded *h);
int b200hnsw_sharded_num_shards(b200hnsw_sharded *h, size_t *n_out);
/* Borrowed handle of one shard (valid until b200hnsw_sharded_destroy): every b200hnsw_* call applies to it. */
int b200hnsw_sharded_get_shard(b200hnsw_sharded *h, size_t shard, b200hnsw_index **out);
int b200hnsw_sharded_count(b200hnsw_sharded *h, uint64_t *count_out);
/* addPoint, routed by label (labels == NULL: consecutive labels continuing the running count). */
int b200hnsw_sharded_add_batch(b200hnsw_sharded *h, const float *X, const uint64_t *labels, size_t n);
int b200hnsw_sharded_flush(b200hnsw_sharded *h);
/* searchKnn over all shards: rows as b200hnsw_search_batch; ef is applied to every shard. */
int b200hnsw_sharded_search_batch(b200hnsw_sharded *h, const float *Q, size_t nq, size_t k, size_t ef,
                                  uint64_t *labels_out, float *dists_out, uint32_t *counts_out);
/* CUDA-event time of the last sharded search (H2D of the queries on every device .. merge kernel), milliseconds. */
int b200hnsw_sharded_last_ms(b200hnsw_sharded *h, double *ms_out);

/* ---- result exchange between processes, one per GPU, without a collective kernel ----------------
 * Every rank owns a receive area in its HBM, published as a B200HNSW_EXCHANGE_DESC_BYTES descriptor (CUDA IPC handles)
 * that the host gathers from all ranks by whatever means it has (the Python form uses torch.distributed once at
 * set-up).  Per step s = 1, 2, ...: the rank's search writes its packed block ([nq*k labels | nq*k dists], as for
 * b200hnsw_merge_topk_packed_device) into *my_block_out of b200hnsw_exchange_slot(s); b200hnsw_exchange_step(s, stream)
 * then pushes it into every peer's area with copy-engine copies over NVLink, raises the peers' flags with stream memory
 * writes and makes `stream` wait (stream memory waits, no SM) until the blocks of all peers for step s have landed in
 * *all_blocks_out, which the merge kernel reads in place.  Three areas are cycled, so a rank may run one step ahead. */
#define B200HNSW_EXCHANGE_DESC_BYTES 128
typedef struct b200hnsw_exchange b200hnsw_exchange;
int b200hnsw_exchange_create(int device, size_t world, size_t rank, size_t block_bytes, b200hnsw_exchange **out,
                             void *desc_out /* B200HNSW_EXCHANGE_DESC_BYTES */);
int b200hnsw_exchange_connect(b200hnsw_exchange *x, const void *all_descs /* world descriptors, rank order */);
int b200hnsw_exchange_slot(b200hnsw_exchange *x, uint32_t step, void **my_block_out, void **all_blocks_out);
int b200hnsw_exchange_step(b200hnsw_exchange *x, uint32_t step, void *cuda_stream);
void b200hnsw_exchange_destroy(b200hnsw_exchange *x);

/* ---- BruteforceSearch<float> (bruteforce.h) -------------------------------------------------------------- */
/* BruteforceSearch(space, maxElements), bruteforce.h:48-59 */
int b200bf_create(const b200hnsw_params *params, b200bf_index **out);
/* BruteforceSearch(space, location) / loadIndex, bruteforce.h:36-45,152-171 */
int b200bf_load(const char *path, const b200hnsw_params *params, b200bf_index **out);
/* saveIndex, bruteforce.h:138-149 (byte-identical) */
int b200bf_save(b200bf_index *h, const char *path);
void b200bf_destroy(b200bf_index *h);
/* addPoint, bruteforce.h:64-83 (existing label -> row overwritten) */
int b200bf_add_batch(b200bf_index *h, const float *X, const uint64_t *labels, size_t n);
/* removePoint, bruteforce.h:86-103 (last row swapped in) */
int b200bf_remove(b200bf_index *h, uint64_t label);
/* searchKnn, bruteforce.h:106-135, batched: the k lexicographically smallest (dist, label) pairs per query. */
int b200bf_search_batch(b200bf_index *h, const float *Q, size_t nq, size_t k, uint64_t *labels_out,
                        float *dists_out, uint32_t *counts_out);
/* searchKnn with a BaseFilterFunctor (bruteforce.h:114,121: rows the functor rejects are skipped).  As for the graph
 * index the caller evaluates the host callback once per stored ROW and passes the verdicts: allowed[i] != 0 for row i
 * (rows in storage order, b200bf_get_labels gives their labels), b200bf_count bytes. */
int b200bf_search_batch_filtered(b200bf_index *h, const float *Q, size_t nq, size_t k, const uint8_t *allowed,
                                 uint64_t *labels_out, float *dists_out, uint32_t *counts_out);
/* label of every stored row, in storage order (bruteforce.h:51-52: the label sits behind the vector). */
int b200bf_get_labels(b200bf_index *h, uint64_t *labels_out, size_t capacity);
int b200bf_search_batch_device(b200bf_index *h, const float *dQ, size_t nq, size_t k, uint64_t *d_labels_out,
                               float *d_dists_out, uint32_t *d_counts_out, void *cuda_stream);
int b200bf_count(b200bf_index *h, uint64_t *count_out);
int b200bf_get_stats(b200bf_index *h, b200hnsw_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* B200HNSW_H_ */
