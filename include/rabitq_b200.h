/* rabitq_b200.h -- C ABI of the B200-native IVF-RaBitQ query path (librabitq_b200.so).
 *
 * Drop-in boundary for ONE path of kemingy/rabitq v0.2.2: `RaBitQ::load_from_dir` + `RaBitQ::query`
 * (reference src/rabitq.rs:84-125 and :268-367, callers crates/cli/src/main.rs:55,71,82).  The reference has
 * no FFI of its own; these entry points are what a Rust `extern "C"` block in src/rabitq.rs would bind
 * (see INTEGRATION.md for the stub).  Plain pointers and sizes only; no C++/torch types.
 *
 * Error behaviour: the reference panics (`expect`/`assert!`, abort in release, Cargo.toml:43).  Here every
 * entry returns 0 on success and a non-zero RABITQ_E* code otherwise; rabitq_last_error() gives the message the
 * reference would have panicked with.  There is NO CPU fallback: without a CUDA device every compute entry
 * fails with RABITQ_ECUDA.
 *
 * Threading: calls on one handle are serialised internally (one stream per handle); distinct handles are
 * independent.  Matches `&self` + relaxed atomics in the reference (src/metrics.rs).
 */
#ifndef RABITQ_B200_H
#define RABITQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rabitq_index rabitq_index; /* opaque; owns host + device memory (struct RaBitQ, src/rabitq.rs:57-68) */

enum {
    RABITQ_OK = 0,
    RABITQ_EIO = 1,      /* file missing / short read ("open ... error", src/rabitq.rs:85-106)            */
    RABITQ_EINVAL = 2,   /* assertion of the reference violated (dim % 64, query length, topk == 0, ...)    */
    RABITQ_ECUDA = 3,    /* CUDA runtime error or no device                                                 */
    RABITQ_ENOMEM = 4,   /* device or host allocation failed                                                */
    RABITQ_EUNSUPPORTED = 5
};

/* ---- construction ------------------------------------------------------------------------------------- */

/* RaBitQ::load_from_dir(path) -- src/rabitq.rs:84-125.  Reads the six-file layout (base.fvecs,
 * orthogonal.fvecs, centroids.fvecs, offsets_ids.ivecs, factors.fvecs, x_binary_vec.u64vecs) and uploads it
 * to CUDA device `device`. */
int rabitq_load_from_dir(const char* dir, int device, rabitq_index** out);

/* Same, keeping only shard `shard_rank` of `shard_count`: a contiguous range of cluster ids balanced by vector
 * count.  Centroids, P and the offsets table are replicated; other clusters look empty to this handle.
 * (Multi-GPU data parallelism, one process per GPU; the reference is single-process.) */
int rabitq_load_from_dir_sharded(const char* dir, int device, int shard_rank, int shard_count, rabitq_index** out);

/* The in-memory equivalent of the `RaBitQ { .. }` struct literal at src/rabitq.rs:114-124 / :254-264: adopt
 * (copy) already-built arrays.  `ptr_on_device` != 0 means every pointer is a device pointer on `device`.
 *   base       n x dim f32, cluster-sorted, UNROTATED, row u contiguous          (rabitq.rs:110-112)
 *   orthogonal dim x dim f32, row r = P[r,:]                                      (orthogonal.fvecs)
 *   centroids  k x dim f32, ROTATED, centroid c contiguous                        (faer dim x k col-major)
 *   offsets    k+1 u32;  map_ids n u32;  codes n x dim/64 u64;  factors n x 4 f32 (ip, ppc, err, cds) */
int rabitq_from_arrays(uint32_t dim, size_t n, size_t k, const float* base, const float* orthogonal,
                       const float* centroids, const uint32_t* offsets, const uint32_t* map_ids,
                       const uint64_t* codes, const float* factors, int ptr_on_device, int device,
                       int shard_rank, int shard_count, rabitq_index** out);

/* ---- index training (the step before the query path) ------------------------------------- */

/* RaBitQ::from_path(base_path, centroid_path) -- src/rabitq.rs:159-265, on the device.  `seed` makes the random
 * orthogonal matrix reproducible (the reference draws it from an unseeded RNG, src/utils.rs:16-20). */
int rabitq_from_path(const char* base_path, const char* centroid_path, uint64_t seed, int device, rabitq_index** out);

/* Same on in-memory arrays: base n x len, centroids k x len (original space, host or device pointers).  `orthogonal`
 * (dim x dim, row r = P[r,:], dim = len rounded up to 64) may be NULL to generate it from `seed`. */
int rabitq_build(const float* base, size_t n, size_t len, const float* centroids, size_t k, const float* orthogonal,
                 uint64_t seed, int ptr_on_device, int device, rabitq_index** out);

/* RaBitQ::dump_to_dir(path) -- src/rabitq.rs:128-156: writes the six-file layout the reference reads back. */
int rabitq_dump_to_dir(rabitq_index* idx, const char* dir);

/* Copy the arrays of `struct RaBitQ` (src/rabitq.rs:57-68) out of an unsharded handle into caller buffers (host, or
 * device when ptr_on_device != 0); shapes as in rabitq_from_arrays, any pointer may be NULL. */
int rabitq_export_arrays(rabitq_index* idx, float* base, float* orthogonal, float* centroids, uint32_t* offsets,
                         uint32_t* map_ids, uint64_t* codes, float* factors, int ptr_on_device);

/* A new handle on the same device holding shard `shard_rank` of `shard_count` of an unsharded handle (device-to-device
 * copies; the source handle stays valid and is usually freed afterwards).  Lets every rank train once and keep its part. */
int rabitq_reshard(rabitq_index* idx, int shard_rank, int shard_count, rabitq_index** out);

void rabitq_free(rabitq_index* idx);

uint32_t rabitq_dim(const rabitq_index* idx);     /* padded D (multiple of 64)        */
size_t rabitq_num_vectors(const rabitq_index* idx); /* vectors held by this handle/shard */
size_t rabitq_num_clusters(const rabitq_index* idx);

/* ---- query ---------------------------------------------------------------------------------------------- */

/* RaBitQ::query(&self, query, probe, topk, heuristic_rank) -> Vec<(f32, u32)> -- src/rabitq.rs:268-333.
 * `len` must satisfy ceil(len/64)*64 == dim (the assert at :275).  Writes up to `topk` (exact squared L2,
 * original id) pairs, ascending by distance (the reference returns heap order), and the count. */
int rabitq_query(rabitq_index* idx, const float* query, size_t len, size_t probe, size_t topk,
                 int heuristic_rank, float* out_dist, uint32_t* out_ids, uint32_t* out_count);

/* The CLI loop (crates/cli/src/main.rs:69-75) as one call: nq queries, row-major nq x len, HOST pointers.
 * out_dist / out_ids are nq x topk (unused tail: +inf / 0xFFFFFFFF), out_count is nq. */
int rabitq_query_batch(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe,
                       size_t topk, int heuristic_rank, float* out_dist, uint32_t* out_ids, uint32_t* out_count);

/* rabitq_query_batch for a serving loop that already holds its NEXT batch (the reference CLI holds every query before its
 * timed loop starts, crates/cli/src/main.rs:63): answers `queries` exactly like rabitq_query_batch and, while the kernels
 * run, uploads `next_queries` (HOST pointer, same nq x len; NULL = nothing) on a copy stream.  The following call whose
 * `queries` is that same pointer (same nq, len) finds them resident and skips its upload.  The caller must not modify the
 * next batch between the two calls.  Results are identical to rabitq_query_batch. */
int rabitq_query_batch_pipelined(rabitq_index* idx, const float* queries, const float* next_queries, size_t nq, size_t len,
                                 size_t probe, size_t topk, int heuristic_rank, float* out_dist, uint32_t* out_ids,
                                 uint32_t* out_count);

/* Same with DEVICE pointers for queries and outputs (inputs already resident in HBM; read in place). */
int rabitq_query_batch_device(rabitq_index* idx, const float* d_queries, size_t nq, size_t len, size_t probe,
                              size_t topk, int heuristic_rank, float* d_out_dist, uint32_t* d_out_ids,
                              uint32_t* d_out_count);

/* Shard geometry used by the sharded constructors (host-only, no device needed): rows [row_lo, row_hi) of the
 * cluster-sorted arrays belong to shard `shard_rank`; boundaries fall on cluster boundaries. */
int rabitq_shard_range(const uint32_t* offsets, size_t k, int shard_rank, int shard_count, size_t* row_lo, size_t* row_hi);

/* Merge `n_lists` per-shard results (each nq x topk, device pointers laid out back to back: list s starts at
 * d_dist + s*nq*topk) into one ascending nq x topk result.  Used after the NCCL all-gather of (dist, id). */
int rabitq_merge_topk_device(int device, const float* d_dist, const uint32_t* d_ids, int n_lists, size_t nq,
                             size_t topk, float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count,
                             void* cuda_stream);

/* ---- distributed pipeline: one process per GPU, index sharded by cluster range, results IDENTICAL to the reference --------
 *
 * The reference is single-process; its filter threshold is sequential state (src/rerank.rs:83-101).  To keep its results
 * while the clusters live on different GPUs, the visit order is cut into the same two rounds on every shard and the
 * survivors of the second round are replayed in order on the query's HOME rank:
 *
 *   rabitq_dist_front   home:   pad, rotate, centroid distances, probe selection of this rank's nq_local queries
 *        -- caller: all-gather of the per-rank chunks (NCCL) --
 *   rabitq_dist_round1  source: local slots + query records for the whole batch; round 1 = the first 128-vector chunk of
 *                               the query's nearest non-empty cluster, replayed on the shard that owns it; records of the
 *                               candidates the reference reranks are written into the home rank's inbox (peer memory)
 *        -- caller: all-reduce(min) of d_thr (every shard now holds the reference's threshold after round 1) --
 *   rabitq_dist_round2  source: everything else, filtered with that (frozen) threshold; exact distances of all survivors,
 *                               shipped to the home inbox by the kernel that computes them (stores over NVLink)
 *        -- caller: all-reduce(max) of d_status (doubles as the barrier that makes the records visible) --
 *   rabitq_dist_finish  home:   HeapReRanker::rank_batch replayed over the union, in the reference's visit order
 *
 * All pointers are DEVICE pointers; everything is launched on the handle's stream (rabitq_set_stream) so the caller's
 * collectives on the same stream order the phases.  Every rank passes the same nq_local / probe / topk. */
int rabitq_dist_init(rabitq_index* idx, int rank, int world, size_t nq_local, size_t probe, size_t topk,
                     size_t records_per_query, size_t* inbox_bytes);
/* CUDA IPC handle (64 bytes) of this rank's inbox, to be opened by the other ranks' processes ... */
int rabitq_dist_ipc_handle(rabitq_index* idx, unsigned char out_handle[64]);
/* ... or its raw device pointer, when the ranks are handles inside ONE process (tests on a single GPU). */
int rabitq_dist_inbox_ptr(rabitq_index* idx, void** out);
/* Register the inbox of rank `peer_rank`: an IPC handle from that rank's process (opened here), or a raw pointer. */
int rabitq_dist_set_peer(rabitq_index* idx, int peer_rank, const unsigned char* ipc_handle, void* raw_ptr);
/* Unmaps every peer inbox this rank has open.  Before an inbox is re-created (rabitq_dist_init with other sizes) EVERY rank calls
 * this, then a barrier, then rabitq_dist_init: an exported allocation must not be freed while a peer still maps it. */
int rabitq_dist_close_peers(rabitq_index* idx);
/* 4-byte words one rank contributes to the all-gather: nq_local x (len + dim + 2*probe + 1). */
size_t rabitq_dist_chunk_words(const rabitq_index* idx, size_t len);
int rabitq_dist_front(rabitq_index* idx, const float* d_queries, size_t len, void* d_send);
int rabitq_dist_round1(rabitq_index* idx, const void* d_gathered, float* d_thr /* world*nq_local */);
/* The same front end in two halves, so that the all-gather of the big part overlaps the centroid scan: `front_rotate` (pad, rotate)
 * fills d_send_qy = [q | y] (rabitq_dist_chunk_words_qy words per rank), the caller starts its all-gather asynchronously,
 * `front_select` (centroid distances, probe selection) fills d_send_meta = [probe ids | probe distances | first non-empty rank]
 * (rabitq_dist_chunk_words_meta words), second all-gather, then `round1_split` takes the two gathered buffers. */
size_t rabitq_dist_chunk_words_qy(const rabitq_index* idx, size_t len);
size_t rabitq_dist_chunk_words_meta(const rabitq_index* idx, size_t len);
/* d_send_qy == NULL (and d_gathered_qy == NULL in round1_split): PUSH mode -- instead of handing the [q | y] chunk to a
 * collective, front_rotate copies it into slot `rank` of every peer's inbox with copy-engine transfers over NVLink on a side
 * stream (no SM, no collective); front_select orders the main stream behind those copies, so the meta all-gather that follows
 * tells every peer that its slot is complete. */
int rabitq_dist_front_rotate(rabitq_index* idx, const float* d_queries, size_t len, void* d_send_qy);
int rabitq_dist_front_select(rabitq_index* idx, void* d_send_meta);
int rabitq_dist_round1_split(rabitq_index* idx, const void* d_gathered_qy, const void* d_gathered_meta, float* d_thr);
int rabitq_dist_round2(rabitq_index* idx, uint32_t* d_status /* 1 word */);
/* d_status bits after the step: 1 = a (home, source) record region overflowed, 2 = this home saw an overflowed segment;
 * non-zero on any rank => repeat the step after rabitq_dist_init with a larger records_per_query. */
int rabitq_dist_finish(rabitq_index* idx, float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count, uint32_t* d_status);
/* The status word as rabitq_dist_finish left it (it is copied to the host by that call's own synchronisation), so the caller
 * need not read `d_status` back itself.  0 = results valid; non-zero = an inbox region overflowed somewhere: grow and repeat. */
int rabitq_dist_last_status(const rabitq_index* idx, uint32_t* out_status);
/* dst[i] = min(dst[i], src[i]) on the device: the all-reduce(min) when all ranks live in one process. */
int rabitq_min_f32_device(int device, float* d_dst, const float* d_src, size_t n, void* cuda_stream);

/* METRICS (src/metrics.rs:30-41): out = {query, rough, precise, cache miss}. */
void rabitq_metrics(const rabitq_index* idx, uint64_t out[4]);
void rabitq_metrics_reset(rabitq_index* idx);

const char* rabitq_last_error(void); /* thread-local */

/* ---- tuning / measurement --------------------------------------------------------------------------------- */

/* Probe-rank boundaries of the rerank rounds (DESIGN.md "exact replay"): rounds[0] = 0 < rounds[1] < ... ;
 * the last round always extends to `probe`.  Default {0, 1}. */
int rabitq_set_rounds(rabitq_index* idx, const uint32_t* rounds, int n);

/* Integer tuning knobs (results never depend on them; tests sweep them):
 *   "first_chunks"     128-vector chunks of the nearest cluster that form the first rerank round (default 1, 0 = whole cluster);
 *   "scan_slices"      shared-memory record slices per scan work item (default 1; hot clusters are cut into several items);
 *   "scan_stages" / "scan_sub"  ring depth / record passes per stage of the scan (0 = by dimension); "scan_mode" 0..2 forces 1, 2, 4
 *                      record tiles (n8) per consumer warp (-1 = by dimension; the tests run every instantiation);
 *   "rerank_mode"      1 (default) = one warp-specialised CTA per query (rerank_cta_kernel), 0 = one warp per query (rerank_kernel);
 *   "rerank_rows"      rows per rerank wave (0 = by dimension; CTA form: <= 8); "rerank_stages" / "rerank_warps" / "rerank_nc" = row
 *                      buffers, compute warps and candidates per eight-lane group of the CTA form (0 = by dimension);
 *   "rerank_prefetch"  L2 prefetch of survivor rows ahead of the gather, warp form (default 0: measured slower on B200);
 *   "debug_rerank"     1 keeps per-query rerank statistics for rabitq_debug_rerank_stats;
 *   "prefilter"        1 (default) lets the centroid scan run as a TF32 tensor-core prefilter + exact recheck of the candidates when
 *                      k >= 512 and probe <= k/8 (probe lists stay bit-identical), 0 = always all k exact distances;
 *   "prefilter_gemm"   1 (default) = the key GEMM on tcgen05 / TMEM / TMA (tc5_gemm.cuh), 0 = the mma.sync form;
 *   "prefilter_cap"    candidates per query the prefilter may certify (<= 1024; a query above it sends its batch to the exact path on
 *                      the device); "prefilter_mode" = 1 plain TF32 keys, 3 = 3xTF32 split, 0 = off (the handle moves 1 -> 3 -> 0 by
 *                      itself when batches cannot be certified);
 *   "speculative_sizing"  1 (default) sizes a batch's survivor slots from earlier batches and checks the capacity on the device (no host
 *                      round trip in the middle of the batch; a batch that does not fit is repeated with exact sizes), 0 = always read
 *                      the totals back first; "spec_words_per_query_milli" overrides the high-water mark (tests);
 *   "dist_r2_seq"      1 (default) = the frozen round of the distributed pipeline runs the source-side sequential filter
 *                      (rerank_cta_kernel<.., SINK = 2>), 0 = compaction + flat exact distances for every survivor. */
int rabitq_set_option(rabitq_index* idx, const char* name, long value);

/* Byte position of dimension d inside a K3 query record (the tensor-core fragment order K4 reads, kernels.cuh rec_pos); host-only,
 * lets the CPU tests check the layout algebra without a GPU. */
int rabitq_debug_rec_pos(int d);

/* After a batch run with "debug_rerank" = 1: out[nq][2 rounds][8] = {SM cycles the query's warp spent in K5, waves, exact
 * distances computed, cycles inside enqueue (including the waves processed there), cycles waiting for gathered rows, in the exact distances, in the replay, staging
 * survivor words} of the last sub-batch.  Tuning aid; not part of the reference's surface. */
int rabitq_debug_rerank_stats(rabitq_index* idx, uint32_t* out, size_t nq);

/* The reference's OTHER query quantiser.  On a host without AVX2, `scalar_quantize` falls back to
 * `scalar_quantize_raw` (src/utils.rs:194-209): q = ((r - lo) * (1/delta) + rand_bias[i]) as u8 -- truncation plus a random
 * bias drawn at load time (src/rabitq.rs:119, never persisted), instead of the AVX2 path's round-half-even without bias.
 * `bias` = dim floats (host pointer) switches K3 to that formula with the caller's bias, so both sides can be given the same
 * noise; NULL restores the AVX2 semantics (the default and the parity target). */
int rabitq_set_quantize_bias(rabitq_index* idx, const float* bias);

/* Launch everything on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream) so
 * the caller can bracket calls with its own CUDA events; NULL restores the handle's private stream. */
int rabitq_set_stream(rabitq_index* idx, void* cuda_stream);

/* CUDA-event timings (ms) of the stages of the LAST rabitq_query_batch* call on this handle, summed over
 * sub-batches and rounds: [0] H2D+pad, [1] rotate, [2] centroid distances, [3] probe select, [4] quantize,
 * [5] bucket/inverted lists, [6] code scan, [7] rerank replay, [8] D2H, [9] total (first to last event).
 * counts: [0] pairs scanned, [1] survivors emitted by the scan, [2] exact distances computed (incl.
 * speculative), [3] precise (reference counter), [4] scan kernel launches, [5] total kernel launches. */
int rabitq_last_timings(const rabitq_index* idx, float ms[10], uint64_t counts[6]);

/* ---- stage-level entries (parity tests against the oracle; host pointers) ----------------------------------- */

/* project (src/utils.rs:237-258): y = q * P, nq x dim. */
int rabitq_stage_rotate(rabitq_index* idx, const float* queries, size_t nq, size_t len, float* out_y);
/* centroid scan + select (src/rabitq.rs:283-297): out_centroid_dist nq x k (may be NULL), probe lists nq x P'
 * with P' = min(probe, k). */
int rabitq_stage_probe(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe,
                       float* out_centroid_dist, uint32_t* out_probe_ids, float* out_probe_dist);
/* min_max_residual + scalar_quantize + vector_binarize_query (src/rabitq.rs:305-317) per (query, probe):
 * out_lo/out_delta nq x P', out_sum nq x P' (u32), out_planes nq x P' x 4*(dim/64) u64. */
int rabitq_stage_quantize(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe,
                          float* out_lo, float* out_delta, uint32_t* out_sum, uint64_t* out_planes);
/* calculate_rough_distance (src/rabitq.rs:336-367) for every (query, probed cluster, vector) in visit order
 * with NO filter: out_rough/out_abdp hold pair_capacity entries; out_pair_start is nq+1 prefix offsets. */
int rabitq_stage_scan(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe,
                      size_t pair_capacity, float* out_rough, uint32_t* out_abdp, uint64_t* out_pair_start);

#ifdef __cplusplus
}
#endif
#endif /* RABITQ_B200_H */
