// rabitq_capi.cu -- host side of librabitq_b200.so: the C ABI declared in include/rabitq_b200.h, the on-disk
// loader (six-file layout of RaBitQ::dump_to_dir, src/rabitq.rs:128-156) and the batch pipeline that strings
// the sm_100a kernels of kernels.cuh together on one CUDA stream.  No CPU fallback exists: every compute
// entry fails with RABITQ_ECUDA when no device is present.
#include "../../include/rabitq_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "build_index.cuh"
#include "prefilter.cuh"
#include "tc5_gemm.cuh"
#include <sys/stat.h>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? RABITQ_ENOMEM : RABITQ_ECUDA,                      \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
};

// ---- vecs files (src/utils.rs:280-330): record = u32 count, then count elements -------------------------------
template <typename T>
bool read_vecs(const std::string& path, std::vector<T>& flat, std::vector<size_t>& rec_len) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    flat.clear();
    rec_len.clear();
    flat.reserve(size_t(sz) / sizeof(T));
    uint32_t cnt;
    while (std::fread(&cnt, 4, 1, f) == 1) {
        size_t old = flat.size();
        flat.resize(old + cnt);
        if (cnt && std::fread(flat.data() + old, sizeof(T), cnt, f) != cnt) {
            std::fclose(f);
            return false;
        }
        rec_len.push_back(cnt);
    }
    std::fclose(f);
    return true;
}

enum Stage { ST_H2D = 0, ST_ROTATE, ST_CDIST, ST_SELECT, ST_QUANT, ST_BUCKET, ST_SCAN, ST_RERANK, ST_D2H, ST_TOTAL, ST_N };

}  // namespace

// State of the distributed pipeline on one rank.  The inbox is where OTHER ranks' kernels write the
// survivor records of the queries homed here; it is a plain cudaMalloc block so that it can be exported with CUDA IPC.
struct DistState {
    bool ready = false;
    int world = 0, rank = 0;
    size_t nq_l = 0, topk = 0, len = 0;
    int P = 0;
    uint32_t r1cap = 0, cap2 = 0;
    unsigned char* inbox = nullptr;
    size_t inbox_bytes = 0, off_r1cnt = 0, off_r1rec = 0, off_r2tab = 0, off_r2rec = 0;
    size_t off_q = 0, off_y = 0;  // flat [world * nq_l][D] regions: padded queries and rotated queries, PUSHED block by block by their owners
    static constexpr int NCOPY = 3;
    cudaEvent_t ev_rot = nullptr, ev_pushes[NCOPY] = {nullptr, nullptr, nullptr};
    cudaStream_t copy_streams[NCOPY] = {nullptr, nullptr, nullptr};
    bool push_pending = false, pushed = false;
    std::vector<unsigned char*> peers_h;
    std::vector<char> opened;  // peers_h[r] came from cudaIpcOpenMemHandle
    unsigned char** peers_d = nullptr;
    int phase = 0;             // 0 idle, 1 front done, 2 round 1 done, 3 round 2 done
    rq::ScanArgs sa;
    rq::RerankArgs ra;
    float* d_thr = nullptr;
};

// Inverted probe lists + scan work list of ONE rerank round (K4's input).  One set per round: the sets of all rounds are built
// on a side stream while K3 runs, so no round waits for its lists.
struct ListSet {
    DevBuf cl_count, cl_start, item_start, cl_cursor, cl_items, work, work_ctl;
    DevBuf qrec;            // this round's records (K3), in list order: a cluster's records are contiguous
    size_t cap_items = 0;   // upper bound of the list's length (queries x ranks of the window)
    uint32_t MS = 0, ch_min = 0;
    cudaEvent_t ready = nullptr;   // inverted list built (K3 and the work-item pass wait for it)
    cudaEvent_t qready = nullptr;  // records written by K3 on the side stream (later rounds only; the scan of the round waits for it)
    cudaEvent_t wready = nullptr;  // flat work items written on the second side stream
    bool w_on_aux = false;
    bool q_on_aux = false;
    void release() {
        for (DevBuf* b : {&cl_count, &cl_start, &item_start, &cl_cursor, &cl_items, &work, &work_ctl, &qrec}) b->release();
        if (ready) cudaEventDestroy(ready);
        if (qready) cudaEventDestroy(qready);
        if (wready) cudaEventDestroy(wready);
        ready = qready = wready = nullptr;
    }
};

struct rabitq_index {
    int device = 0;
    uint32_t D = 0;
    size_t n = 0, K = 0;  // vectors held by this shard, clusters (all)
    int shard_rank = 0, shard_count = 1;
    int sm_count = 148;
    uint32_t max_cluster = 0;
    // resident index (HBM)
    float* base = nullptr;        // n x D
    float* P = nullptr;           // D x D rows
    float* PT = nullptr;          // P transposed (PT[c][r] = P[r][c]): the operand layout of K1
    float* cent = nullptr;        // K x D
    uint32_t* offsets = nullptr;  // K+1, local rows
    // tensor-core prefilter of the centroid scan (prefilter.cuh): mean centroid, centred TF32 centroids, their norms
    float *pf_mu = nullptr, *pf_chat = nullptr, *pf_chat_lo = nullptr, *pf_cnorm = nullptr, *pf_cnorm2 = nullptr, *pf_cnorm_max = nullptr;
    int prefilter = 1;            // 0 = always the classic exact scan + select
    int pf_mode = 1;              // 1 = plain TF32 keys, 3 = 3xTF32 split (tighter bound), 0 = gave up (exact path); adapts to the data
    int pf_strikes = 0;           // batches the current mode could not certify
    bool pf_pending = false;      // a fallback flag is on its way to h_pin[6]
    // tot_blk (device, 8 words) mirrors h_pin[0..8): [0] slot words of the batch, [1] speculative-sizing overflow flag, [2..3] pairs,
    // [6] prefilter fallback flag, [7] prefilter candidates (statistics): ONE small copy brings all of it to the host
    size_t pf_last_nb = 0;
    int pf_gemm = 1;              // 1 = tcgen05 / TMEM / TMA key GEMM (tc5_gemm.cuh), 0 = mma.sync key GEMM (prefilter.cuh)
    CUtensorMap tm_chat, tm_chat_lo, tm_yhat, tm_yhat_lo;
    bool tm_c_ready = false;
    int prefilter_cap = 1024;     // candidates per query the prefilter may certify (tests lower it to force the fallback)
    float* quant_bias = nullptr;  // D: non-NULL switches K3 to the reference's non-AVX2 quantiser (rabitq_set_quantize_bias)
    uint32_t* goffsets = nullptr; // K+1, rows of the WHOLE index (equal to offsets on an unsharded handle)
    uint32_t* row_bounds = nullptr;  // shard_count+1, global rows where the shards begin
    uint32_t* map_ids = nullptr;  // n
    uint32_t* codes = nullptr;    // n x D/32
    float4* factors = nullptr;    // n
    // scan-layout copy of codes + Factors (kernels.cuh scan_layout_kernel): per cluster, chunks of 128 vectors in the image K4 stages
    uint2* scan_codes = nullptr;      // chunks x [D/64][128]
    float4* scan_fac = nullptr;       // chunks x 128
    uint32_t* chunk_start = nullptr;  // K+1: first chunk of every cluster
    size_t n_chunks = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_totals = nullptr;  // marks the arrival of the slot totals in h_pin
    cudaStream_t aux_stream = nullptr;  // side stream: the rounds' inverted lists are built here, concurrently with K3
    cudaStream_t aux2_stream = nullptr; // second side stream: round windows and the scan's flat work items (off the main stream's chain of tiny launches)
    cudaEvent_t ev_win = nullptr;       // round windows written (the first replay waits for it)
    // speculative survivor-slot sizing: slots sized from the largest words-per-query seen so far, verified on the device
    bool spec_enabled = true;           // (RABITQ_SPEC=0: always read the totals back in the middle of the batch)
    bool spec_on = false;               // this sub-batch runs speculatively
    uint32_t spec_cap = 0;              // slot words it may use
    double hw_wpq = 0.0;                // high-water mark of slot words per query
    bool use_aux2 = true;               // (A/B switch RABITQ_AUX2=0: the same work on the main stream)
    bool state_reset_done = false;      // thresholds / counters of this batch were reset on the side stream already
    cudaEvent_t ev_fork = nullptr;      // recorded on `stream` after K2b: everything the lists depend on is done
    std::vector<ListSet> lists;
    std::mutex mu;
    std::vector<uint32_t> rounds{0};
    int rerank_rows = 0;   // rows per rerank wave; 0 = by dimension
    int rerank_mode = 1;   // 1: one CTA per query, warp-specialised (rerank_cta_kernel); 0: one warp per query (rerank_kernel)
    bool dist_sink_cta = true; // distributed round 1 on the CTA form of K5 (RABITQ_DIST_SINK_CTA=0: the warp form)
    bool dist_r2_seq = true;   // distributed frozen round: source-side sequential filter (rerank_cta_kernel<.., 2>) instead of the flat exact pass (RABITQ_DIST_R2_SEQ=0)
    int rerank_nc = 0;     // candidates per eight-lane group of the CTA form (1 or 2; 0 = by wave size)
    int pf_threads = 0;    // 256: always 256 threads per query in prefilter_select_kernel (A/B switch)
    int rerank_stages = 0; // row buffers in the CTA form's ring (0 = 4)
    int rerank_warps = 0;  // compute warps of the CTA form (0 = 3)
    int debug_rerank = 0;     // per-query rerank statistics (rabitq_debug_rerank_stats)
    int rerank_prefetch = 0;  // L2 prefetch of survivor rows in the rerank stream (rabitq_set_option("rerank_prefetch"))
    int first_chunks = 1;  // 128-vector chunks of the nearest cluster in the first round (0 = the whole cluster)
    int scan_slices = 1;   // shared-memory record slices per scan work item (hot clusters are cut into several items)
    // work buffers
    DevBuf qraw, qpad, y, cdist, probe_ids, probe_dist, slot_local, q_words, q_pairs, q_p0, q_wbase, q_pbase, thr, heap_dist, heap_ids, heap_cnt, q_precise, h_recent, h_wcount, bitmap,
        entries, counters, out_all, rr_dbg, round_win, r2_cnt, r2_off, home_tot, cand, pf_yhat, pf_yhat_lo, pf_ynorm, sel_scratch, tot_blk;
    DistState dist;
    const float* q_in = nullptr;   // the sub-batch's raw queries (nb x len) on the device: ix->qraw, or the caller's device pointer
    const float* y_all = nullptr;  // rotated queries K3 reads: ix->y, or (distributed push mode) the inbox region every rank pushed its block into
    const float* q_pad = nullptr;  // the same, zero-padded to D: ix->qpad, or q_in itself when len == D (nothing to pad, nothing copied)
    // double-buffered upload (rabitq_query_batch_pipelined): the NEXT batch's queries travel on the copy stream while this one runs
    DevBuf qstage;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_staged = nullptr;
    const float* staged_src = nullptr;
    size_t staged_nq = 0, staged_len = 0;
    const float* stage_next = nullptr;  // upload still to be started for this call
    size_t stage_bytes = 0;
    uint32_t* h_pin = nullptr;  // small pinned staging (totals, counters)
    float* ovr_dist = nullptr;  // device-pointer call answered in one sub-batch: K5 writes straight into the caller's tensors
    uint32_t *ovr_ids = nullptr, *ovr_count = nullptr;
    uint32_t* h_out = nullptr;  // pinned staging of a sub-batch's results (host-pointer calls)
    size_t h_out_cap = 0;
    // metrics (src/metrics.rs)
    std::atomic<uint64_t> m_query{0}, m_rough{0}, m_precise{0};  // relaxed atomics like the reference's METRICS: read without the handle lock
    // last-call measurements
    float ms[ST_N] = {0};
    uint64_t counts[6] = {0};
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stage;
    size_t ev_used = 0;
    bool timings_pending = false;  // the last call's event chain has not been turned into ms[] yet
    int scan_blocks_per_sm = 0;
    bool scan_attr_done[2][5] = {{false, false, false, false, false}, {false, false, false, false, false}};
    int scan_stages = 0;   // ring depth of the scan (0 = by dimension)
    int scan_sub = 0;      // consumer passes per stage (0 = by dimension)
    size_t max_items = 0;  // capacity of a round's scan work list for the current sub-batch (setup_rounds)
    int scan_mode = -1;  // -1 = default; RABITQ_SCAN_MODE overrides (tuning)

    ~rabitq_index() {
        cudaSetDevice(device);
        for (void* p : {(void*)base, (void*)P, (void*)cent, (void*)offsets, (void*)goffsets, (void*)row_bounds, (void*)map_ids, (void*)codes,
                        (void*)factors, (void*)scan_codes, (void*)scan_fac, (void*)chunk_start, (void*)PT, (void*)dist.peers_d, (void*)quant_bias, (void*)pf_mu, (void*)pf_chat, (void*)pf_chat_lo, (void*)pf_cnorm, (void*)pf_cnorm2,
                        (void*)pf_cnorm_max})
            if (p) cudaFree(p);
        for (size_t r = 0; r < dist.peers_h.size(); r++)
            if (dist.opened[r] && dist.peers_h[r]) cudaIpcCloseMemHandle(dist.peers_h[r]);
        if (dist.inbox) cudaFree(dist.inbox);
        if (dist.ev_rot) cudaEventDestroy(dist.ev_rot);
        for (int c = 0; c < DistState::NCOPY; c++) {
            if (dist.copy_streams[c]) cudaStreamDestroy(dist.copy_streams[c]);
            if (dist.ev_pushes[c]) cudaEventDestroy(dist.ev_pushes[c]);
        }
        for (DevBuf* b : {&qraw, &qpad, &y, &cdist, &probe_ids, &probe_dist, &slot_local, &q_words, &q_pairs, &q_p0, &q_wbase, &q_pbase,
                          &thr, &heap_dist,
                          &heap_ids, &heap_cnt, &q_precise, &h_recent, &h_wcount, &bitmap, &entries, &counters, &out_all, &rr_dbg, &round_win,
                          &r2_cnt, &r2_off, &home_tot, &cand, &pf_yhat, &pf_yhat_lo, &pf_ynorm, &sel_scratch, &tot_blk})
            b->release();
        if (h_pin) cudaFreeHost(h_pin);
        if (h_out) cudaFreeHost(h_out);
        for (auto e : ev_pool) cudaEventDestroy(e);
        if (ev_totals) cudaEventDestroy(ev_totals);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& l : lists) l.release();
        qstage.release();
        if (ev_staged) cudaEventDestroy(ev_staged);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (aux2_stream) cudaStreamDestroy(aux2_stream);
        if (ev_win) cudaEventDestroy(ev_win);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
};

namespace {

using namespace rq;

int tick(rabitq_index* ix, int stage) {
    if (ix->ev_used == ix->ev_pool.size()) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ix->ev_pool.push_back(e);
        ix->ev_stage.push_back(0);
    }
    ix->ev_stage[ix->ev_used] = stage;
    CU(cudaEventRecord(ix->ev_pool[ix->ev_used], ix->stream));
    ix->ev_used++;
    return 0;
}

// PT = P^T, once per index (K1 reads a column's rows contiguously)
int make_pt(rabitq_index* ix) {
    const int D = (int)ix->D;
    if (!ix->PT) {
        CU(cudaMalloc((void**)&ix->PT, (size_t)D * D * 4));
        CU(cudaFuncSetAttribute(rotate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ROT_SMEM_BYTES));
    }
    transpose_square_kernel<<<dim3(D / 32, D / 32), dim3(32, 8)>>>(ix->P, ix->PT, D);
    CU(cudaGetLastError());
    return 0;
}

// K1 for `rows` padded vectors (queries, centroids or base rows at build time)
int launch_rotate(rabitq_index* ix, const float* in, float* out, size_t rows, cudaStream_t st) {
    dim3 grid((unsigned)(ix->D / ROT_TC), (unsigned)((rows + ROT_TQ - 1) / ROT_TQ));
    rotate_kernel<<<grid, ROT_THREADS, ROT_SMEM_BYTES, st>>>(in, ix->PT, out, (int)rows, (int)ix->D);
    CU(cudaGetLastError());
    return 0;
}

// Stage timings are read from the event chain LAZILY (rabitq_last_timings): ~20 cudaEventElapsedTime calls cost tens of
// microseconds of host time, which a serving loop that never asks for them should not pay between batches.
int collect_timings(rabitq_index* ix) {
    CU(cudaStreamSynchronize(ix->stream));
    ix->timings_pending = true;
    return 0;
}

int resolve_timings(rabitq_index* ix) {
    if (!ix->timings_pending) return 0;
    ix->timings_pending = false;
    static const bool trace = std::getenv("RABITQ_TRACE") != nullptr;
    for (size_t i = 1; i < ix->ev_used; i++) {
        float t = 0;
        if (trace || ix->ev_stage[i] >= 0) CU(cudaEventElapsedTime(&t, ix->ev_pool[i - 1], ix->ev_pool[i]));
        if (trace) std::fprintf(stderr, "[rabitq trace r%d] ev %2zu stage %2d  %.4f ms\n", ix->shard_rank, i, ix->ev_stage[i], t);
        if (ix->ev_stage[i] < 0) continue;  // a restart marker: gaps spent in the caller's collectives are not counted
        ix->ms[ix->ev_stage[i]] += t;
    }
    if (ix->ev_used >= 2) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, ix->ev_pool[0], ix->ev_pool[ix->ev_used - 1]));
        ix->ms[ST_TOTAL] = t;
    }
    return 0;
}

// ---- shard geometry: contiguous cluster-id ranges balanced by EXPECTED SCAN WORK -------------------------------------------
// The pairs a shard scans are sum_c n_c * (queries probing c).  Queries follow the data, so a big cluster is probed more often,
// but with P probes per query most of a cluster's visitors come from its NEIGHBOURS, which averages the effect out: measured on
// C2 at two shards, ranges balanced by n_c alone were 5.6 % off one way, ranges balanced by n_c * (n_c + mean) / 2 were 12 % off
// the other way.  The weight n_c * (0.2 n_c + 0.8 mean) sits between; memory stays within a small factor of even.
void shard_rows(const uint32_t* offsets, size_t K, int rank, int count, size_t* row_lo, size_t* row_hi) {
    const size_t N = offsets[K];
    const double mean = K ? (double)N / (double)K : 0.0;
    std::vector<double> cum(K + 1, 0.0);
    for (size_t c = 0; c < K; c++) {
        const double n = (double)(offsets[c + 1] - offsets[c]);
        cum[c + 1] = cum[c] + n * (0.2 * n + 0.8 * mean);
    }
    auto bound = [&](int r) -> size_t {
        if (r <= 0) return 0;
        if (r >= count) return N;
        const double target = cum[K] * (double)r / (double)count;
        const size_t c = (size_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        return offsets[std::min(c, K)];
    };
    *row_lo = bound(rank);
    *row_hi = bound(rank + 1);
}

// common tail of every constructor: device properties, stream, pinned staging, kernel attributes, env knobs
int finish_index(rabitq_index* ix) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ix->device));
    ix->sm_count = prop.multiProcessorCount;
    if (make_pt(ix)) return RABITQ_ECUDA;
    if (const char* e = std::getenv("RABITQ_SCAN_MODE")) ix->scan_mode = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_FIRST_CHUNKS")) ix->first_chunks = std::max(0, std::atoi(e));
    if (const char* e = std::getenv("RABITQ_RR_ROWS")) ix->rerank_rows = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_RR_PREFETCH")) ix->rerank_prefetch = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_RR_MODE")) ix->rerank_mode = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_RR_NC")) ix->rerank_nc = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_PF_THREADS")) ix->pf_threads = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_PF_GEMM")) ix->pf_gemm = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_AUX2")) ix->use_aux2 = std::atoi(e) != 0;
    if (const char* e = std::getenv("RABITQ_SPEC")) ix->spec_enabled = std::atoi(e) != 0;
    CU(cudaFuncSetAttribute(approx_gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC5_SMEM_BYTES));
    if (const char* e = std::getenv("RABITQ_RR_STAGES")) ix->rerank_stages = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_RR_WARPS")) ix->rerank_warps = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_SCAN_SLICES")) ix->scan_slices = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RABITQ_SCAN_STAGES")) ix->scan_stages = std::max(0, std::atoi(e));
    if (const char* e = std::getenv("RABITQ_SCAN_SUB")) ix->scan_sub = std::max(0, std::atoi(e));
    CU(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    ix->stream = ix->own_stream;
    CU(cudaStreamCreateWithFlags(&ix->aux_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ix->aux2_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ix->ev_win, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ix->ev_staged, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ix->ev_fork, cudaEventDisableTiming));
    CU(cudaMallocHost((void**)&ix->h_pin, 256));
    std::memset(ix->h_pin, 0, 256);
    CU(cudaFuncSetAttribute(rerank_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (const char* e = std::getenv("RABITQ_DIST_SINK_CTA")) ix->dist_sink_cta = std::atoi(e) != 0;
    if (const char* e = std::getenv("RABITQ_DIST_R2_SEQ")) ix->dist_r2_seq = std::atoi(e) != 0;
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(rerank_cta_kernel<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(select_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CU(cudaFuncSetAttribute(approx_gemm_tf32_kernel<128, 128, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * PF_PITCH * 4));
    if (const char* e = std::getenv("RABITQ_PREFILTER")) ix->prefilter = std::atoi(e);
    if (const char* e = std::getenv("RABITQ_PREFILTER_MODE")) ix->pf_mode = std::atoi(e);
    {   // scan-layout copy of the codes and Factors: what K4's producer warp stages with one TMA bulk copy per chunk
        const size_t K = ix->K, JJ = ix->D / 64;
        std::vector<uint32_t> off_h(K + 1), cs(K + 1);
        CU(cudaMemcpy(off_h.data(), ix->offsets, (K + 1) * 4, cudaMemcpyDeviceToHost));
        size_t tot = 0;
        for (size_t c = 0; c < K; c++) {
            cs[c] = (uint32_t)tot;
            tot += ((size_t)(off_h[c + 1] - off_h[c]) + SCAN_THREADS - 1) / SCAN_THREADS;
        }
        cs[K] = (uint32_t)tot;
        if (tot >= ((size_t)1 << 32)) return fail(RABITQ_EUNSUPPORTED, "too many scan chunks");
        ix->n_chunks = tot;
        CU(cudaMalloc((void**)&ix->chunk_start, (K + 1) * 4));
        CU(cudaMemcpy(ix->chunk_start, cs.data(), (K + 1) * 4, cudaMemcpyHostToDevice));
        CU(cudaMalloc((void**)&ix->scan_codes, std::max<size_t>(tot * JJ * 128 * 8, 16)));
        CU(cudaMalloc((void**)&ix->scan_fac, std::max<size_t>(tot * 128 * 16, 16)));
        if (tot) {
            scan_layout_kernel<<<(unsigned)tot, 128>>>(ix->codes, ix->factors, ix->offsets, ix->chunk_start, (int)K, (int)JJ, ix->scan_codes, ix->scan_fac);
            CU(cudaGetLastError());
        }
    }
    {   // prefilter operands: mu, c^ = tf32(c - mu), ||c - mu||, ||c - mu||^2, max norm
        const size_t K = ix->K, D = ix->D;
        CU(cudaMalloc((void**)&ix->pf_mu, D * 4));
        CU(cudaMalloc((void**)&ix->pf_chat, K * D * 4));
        CU(cudaMalloc((void**)&ix->pf_chat_lo, K * D * 4));
        CU(cudaMalloc((void**)&ix->pf_cnorm, K * 4));
        CU(cudaMalloc((void**)&ix->pf_cnorm2, K * 4));
        CU(cudaMalloc((void**)&ix->pf_cnorm_max, 4));
        CU(cudaMemset(ix->pf_cnorm_max, 0, 4));
        centroid_mean_kernel<<<(unsigned)((D + 127) / 128), 128>>>(ix->cent, (int)K, (int)D, ix->pf_mu);
        centroid_center_kernel<<<(unsigned)((K + 3) / 4), 128>>>(ix->cent, ix->pf_mu, (int)K, (int)D, ix->pf_chat, ix->pf_chat_lo, ix->pf_cnorm, ix->pf_cnorm2,
                                                                  ix->pf_cnorm_max);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
    }
    return 0;
}

// Is the tensor-core prefilter worth it (and within its candidate capacity) for this (K, P)?
bool use_prefilter(const rabitq_index* ix, int P) {
    return ix->prefilter && ix->pf_mode != 0 && ix->K >= 512 && (size_t)P * 8 <= ix->K && P <= 384;
}

// Host-side adaptation, called wherever the stream has just been synchronised: a batch the prefilter could not certify (it was
// then answered by the exact kernels on the device) moves the handle to the tighter 3xTF32 keys, and a second one to the
// exact path for good.  Results never depend on the mode.
void prefilter_adapt(rabitq_index* ix) {
    if (!ix->pf_pending) return;
    ix->pf_pending = false;
    static const bool pf_trace = std::getenv("RABITQ_TRACE") != nullptr;
    if (pf_trace && ix->pf_last_nb) std::fprintf(stderr, "[rabitq trace] prefilter mode %d: %.1f candidates rechecked per query\n", ix->pf_mode, (double)ix->h_pin[7] / (double)ix->pf_last_nb);
    if (ix->h_pin[6] == 0u) return;
    if (std::getenv("RABITQ_TRACE")) std::fprintf(stderr, "[rabitq trace] prefilter mode %d could not certify a batch (reason bits %u)\n", ix->pf_mode, ix->h_pin[6]);
    ix->h_pin[6] = 0u;
    if (ix->prefilter_cap < PS_CAP) return;  // a test lowered the capacity to force the fallback: not the data's fault
    if (++ix->pf_strikes >= 2) {
        ix->pf_mode = ix->pf_mode == 1 ? 3 : 0;
        ix->pf_strikes = 0;
    }
}

// Build a handle from host or device arrays (full index); keeps only this shard's rows.
int make_index(uint32_t dim, size_t n_total, size_t K, const float* base, const float* orth, const float* cent,
               const uint32_t* offsets_h /* host copy, always */, const uint32_t* map_ids, const uint64_t* codes,
               const float* factors, bool on_device, int device, int shard_rank, int shard_count, rabitq_index** out) {
    if (dim == 0 || dim % 64 != 0) return fail(RABITQ_EINVAL, "assertion failed: dim % 64 == 0");
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(RABITQ_EINVAL, "bad shard rank/count");
    if (K == 0) return fail(RABITQ_EINVAL, "offsets is empty");
    if (offsets_h[K] != n_total) return fail(RABITQ_EINVAL, "offsets[k] != number of vectors");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RABITQ_ECUDA, "no CUDA device: rabitq_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(RABITQ_EINVAL, "bad device ordinal");
    CU(cudaSetDevice(device));
    auto ix = new rabitq_index();
    ix->device = device;
    ix->D = dim;
    ix->K = K;
    ix->shard_rank = shard_rank;
    ix->shard_count = shard_count;
    size_t row_lo, row_hi;
    shard_rows(offsets_h, K, shard_rank, shard_count, &row_lo, &row_hi);
    ix->n = row_hi - row_lo;
    std::vector<uint32_t> loc(K + 1);
    uint32_t mx = 0;
    for (size_t c = 0; c <= K; c++) {
        size_t o = std::min(std::max((size_t)offsets_h[c], row_lo), row_hi) - row_lo;
        loc[c] = (uint32_t)o;
        if (c) mx = std::max(mx, loc[c] - loc[c - 1]);
    }
    ix->max_cluster = mx;
    const size_t D = dim, W32 = D / 32, n = ix->n;
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
#define ALLOC_COPY(dst, type, count, src)                                                          \
    do {                                                                                           \
        cudaError_t e_ = cudaMalloc((void**)&(dst), std::max<size_t>((count) * sizeof(type), 16)); \
        if (e_ != cudaSuccess) { delete ix; return fail(RABITQ_ENOMEM, std::string("cudaMalloc ") + #dst + ": " + cudaGetErrorString(e_)); } \
        if ((count) > 0) {                                                                         \
            e_ = cudaMemcpy((dst), (src), (count) * sizeof(type), kind);                           \
            if (e_ != cudaSuccess) { delete ix; return fail(RABITQ_ECUDA, std::string("cudaMemcpy ") + #dst + ": " + cudaGetErrorString(e_)); } \
        }                                                                                          \
    } while (0)
    ALLOC_COPY(ix->base, float, n * D, base + row_lo * D);
    ALLOC_COPY(ix->P, float, D * D, orth);
    ALLOC_COPY(ix->cent, float, K * D, cent);
    ALLOC_COPY(ix->map_ids, uint32_t, n, map_ids + row_lo);
    ALLOC_COPY(ix->codes, uint32_t, n * W32, reinterpret_cast<const uint32_t*>(codes) + row_lo * W32);
    ALLOC_COPY(ix->factors, float4, n, reinterpret_cast<const float4*>(factors) + row_lo);
#undef ALLOC_COPY
    {
        cudaError_t e_ = cudaMalloc((void**)&ix->offsets, (K + 1) * 4);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(ix->offsets, loc.data(), (K + 1) * 4, cudaMemcpyHostToDevice);
        if (e_ != cudaSuccess) { delete ix; return fail(RABITQ_ECUDA, std::string("offsets upload: ") + cudaGetErrorString(e_)); }
    }
    {
        std::vector<uint32_t> rb(shard_count + 1);
        for (int r = 0; r <= shard_count; r++) {
            size_t lo_r, hi_r;
            shard_rows(offsets_h, K, std::min(r, shard_count - 1), shard_count, &lo_r, &hi_r);
            rb[r] = (uint32_t)(r < shard_count ? lo_r : hi_r);
        }
        cudaError_t e_ = cudaMalloc((void**)&ix->goffsets, (K + 1) * 4);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(ix->goffsets, offsets_h, (K + 1) * 4, cudaMemcpyHostToDevice);
        if (e_ == cudaSuccess) e_ = cudaMalloc((void**)&ix->row_bounds, rb.size() * 4);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(ix->row_bounds, rb.data(), rb.size() * 4, cudaMemcpyHostToDevice);
        if (e_ != cudaSuccess) { delete ix; return fail(RABITQ_ECUDA, std::string("global offsets upload: ") + cudaGetErrorString(e_)); }
    }
    { int rc_ = finish_index(ix); if (rc_) { delete ix; return rc_; } }
    *out = ix;
    return RABITQ_OK;
}

int load_dir(const char* dir, int device, int shard_rank, int shard_count, rabitq_index** out) {
    if (!dir || !out) return fail(RABITQ_EINVAL, "null argument");
    std::string d(dir);
    std::vector<float> orth, cent_dk, fac, base;
    std::vector<uint32_t> offs;
    std::vector<uint64_t> xb;
    std::vector<size_t> rl_orth, rl_cent, rl_fac, rl_base, rl_offs, rl_xb;
    if (!read_vecs(d + "/orthogonal.fvecs", orth, rl_orth) || rl_orth.empty()) return fail(RABITQ_EIO, "read vecs error: orthogonal.fvecs");
    if (!read_vecs(d + "/centroids.fvecs", cent_dk, rl_cent) || rl_cent.empty()) return fail(RABITQ_EIO, "read vecs error: centroids.fvecs");
    if (!read_vecs(d + "/offsets_ids.ivecs", offs, rl_offs)) return fail(RABITQ_EIO, "open offsets_ids error");
    if (rl_offs.empty()) return fail(RABITQ_EINVAL, "offsets is empty");
    if (!read_vecs(d + "/factors.fvecs", fac, rl_fac)) return fail(RABITQ_EIO, "open factors error");
    if (!read_vecs(d + "/x_binary_vec.u64vecs", xb, rl_xb)) return fail(RABITQ_EIO, "open x_binary_vec error");
    if (!read_vecs(d + "/base.fvecs", base, rl_base)) return fail(RABITQ_EIO, "read vecs error: base.fvecs");
    const size_t D = rl_orth.size();  // orthogonal.nrows(), rabitq.rs:108
    if (D % 64 != 0) return fail(RABITQ_EINVAL, "assertion failed: dim % 64 == 0");
    if (orth.size() != D * D) return fail(RABITQ_EINVAL, "orthogonal.fvecs is not square");
    if (rl_cent.size() != D) return fail(RABITQ_EINVAL, "centroids.fvecs must hold dim records (matrix is dim x k)");
    const size_t K = rl_cent[0];
    if (cent_dk.size() != D * K) return fail(RABITQ_EINVAL, "centroids.fvecs records differ in length");
    // offsets = first record, map_ids = last record (rabitq.rs:90-91)
    const size_t off_len = rl_offs.front(), ids_len = rl_offs.back();
    if (off_len != K + 1) return fail(RABITQ_EINVAL, "offsets length != k + 1");
    const uint32_t* offsets = offs.data();
    const uint32_t* map_ids = offs.data() + (offs.size() - ids_len);
    const size_t N = ids_len;
    if (base.size() != N * D) return fail(RABITQ_EINVAL, "base.fvecs size != n * dim");
    if (fac.size() != N * 4) return fail(RABITQ_EINVAL, "factors.fvecs size != 4 * n");
    if (xb.size() != N * (D / 64)) return fail(RABITQ_EINVAL, "x_binary_vec.u64vecs size != n * dim / 64");
    // file rows are components (dim x k); the device wants each rotated centroid contiguous (k x dim)
    std::vector<float> cent(K * D);
    for (size_t dd = 0; dd < D; dd++)
        for (size_t c = 0; c < K; c++) cent[c * D + dd] = cent_dk[dd * K + c];
    return make_index((uint32_t)D, N, K, base.data(), orth.data(), cent.data(), offsets, map_ids, xb.data(), fac.data(), false,
                      device, shard_rank, shard_count, out);
}

// ---- index training on the device: RaBitQ::from_path (src/rabitq.rs:159-265) ----------------------------------------
struct TmpBufs {  // frees whatever was allocated, on every exit path
    std::vector<void*> ptrs;
    ~TmpBufs() { for (void* p : ptrs) if (p) cudaFree(p); }
    cudaError_t alloc(void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

int build_impl(const float* base, size_t n, size_t len, const float* centroids, size_t k, const float* orthogonal, uint64_t seed,
               bool on_device, int device, rabitq_index** out) {
    if (!base || !centroids || !out) return fail(RABITQ_EINVAL, "null argument");
    if (n == 0 || k == 0 || len == 0) return fail(RABITQ_EINVAL, "empty base or centroids");
    if (n >= ((size_t)1 << 32)) return fail(RABITQ_EUNSUPPORTED, "more than 2^32-1 vectors");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RABITQ_ECUDA, "no CUDA device: rabitq_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(RABITQ_EINVAL, "bad device ordinal");
    CU(cudaSetDevice(device));
    const size_t D = (len + 63) / 64 * 64, W32 = D / 32, K = k;  // padding to 64, rabitq.rs:167-179
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    TmpBufs tmp;
    const float* d_base_in = base;
    const float* d_cent_in = centroids;
    if (!on_device) {
        float *b = nullptr, *c = nullptr;
        CU(tmp.alloc((void**)&b, n * len * 4));
        CU(tmp.alloc((void**)&c, K * len * 4));
        CU(cudaMemcpy(b, base, n * len * 4, kind));
        CU(cudaMemcpy(c, centroids, K * len * 4, kind));
        d_base_in = b;
        d_cent_in = c;
    }
    auto ix = new rabitq_index();
    struct Guard { rabitq_index* p; ~Guard() { delete p; } } guard{ix};
    ix->device = device;
    ix->D = (uint32_t)D;
    ix->K = K;
    ix->n = n;
    CU(cudaMalloc((void**)&ix->P, D * D * 4));
    CU(cudaMalloc((void**)&ix->cent, K * D * 4));
    CU(cudaMalloc((void**)&ix->offsets, (K + 1) * 4));
    // --- P: given, or Q of a seeded standard-normal matrix (utils.rs:16-20; fp64 Gram-Schmidt applied twice) ---
    if (orthogonal) {
        CU(cudaMemcpy(ix->P, orthogonal, D * D * 4, kind));
    } else {
        double *A = nullptr, *dots = nullptr;
        CU(tmp.alloc((void**)&A, D * D * 8));
        CU(tmp.alloc((void**)&dots, D * 8));
        gauss_fill_kernel<<<(unsigned)((D * D + 255) / 256), 256>>>(A, D * D, seed * 0x9e3779b97f4a7c15ull + 12345);
        for (int j = 0; j < (int)D; j++) {
            for (int pass = 0; pass < 2 && j > 0; pass++) {
                gs_dots_kernel<<<(j + 127) / 128, 128>>>(A, (int)D, j, dots);
                gs_update_kernel<<<((int)D + 127) / 128, 128>>>(A, (int)D, j, dots);
            }
            gs_normalize_kernel<<<1, 256>>>(A, (int)D, j);
        }
        f64_to_f32_kernel<<<(unsigned)((D * D + 255) / 256), 256>>>(A, ix->P, D * D);
        CU(cudaGetLastError());
    }
    // --- rotated centroids (rabitq.rs:189) ---
    {
        float* cpad = nullptr;
        CU(tmp.alloc((void**)&cpad, K * D * 4));
        pad_queries_kernel<<<(unsigned)((K * D + 255) / 256), 256>>>(d_cent_in, cpad, K, (int)len, (int)D);
        if (make_pt(ix) || launch_rotate(ix, cpad, ix->cent, K, 0)) return RABITQ_ECUDA;
    }
    // --- per-vector label, code, Factor; chunks keep the distance tile around 1 GiB ---
    uint32_t *label = nullptr, *codes_u = nullptr, *vals_in = nullptr, *vals_out = nullptr, *counts = nullptr;
    float* min_dist = nullptr;
    float4* fac_u = nullptr;
    unsigned long long *key_in = nullptr, *key_out = nullptr;
    CU(tmp.alloc((void**)&label, n * 4));
    CU(tmp.alloc((void**)&min_dist, n * 4));
    CU(tmp.alloc((void**)&codes_u, n * W32 * 4));
    CU(tmp.alloc((void**)&fac_u, n * 16));
    CU(tmp.alloc((void**)&key_in, n * 8));
    CU(tmp.alloc((void**)&key_out, n * 8));
    CU(tmp.alloc((void**)&vals_in, n * 4));
    CU(tmp.alloc((void**)&vals_out, n * 4));
    CU(tmp.alloc((void**)&counts, K * 4));
    {
        const size_t CH = std::max<size_t>(1024, std::min<size_t>(65536, ((size_t)1 << 28) / K)) / 32 * 32;
        float *xpad = nullptr, *xp = nullptr, *dist = nullptr;
        CU(tmp.alloc((void**)&xpad, CH * D * 4));
        CU(tmp.alloc((void**)&xp, CH * D * 4));
        CU(tmp.alloc((void**)&dist, CH * K * 4));
        for (size_t s0 = 0; s0 < n; s0 += CH) {
            const size_t rows = std::min(CH, n - s0);
            pad_queries_kernel<<<(unsigned)((rows * D + 255) / 256), 256>>>(d_base_in + s0 * len, xpad, rows, (int)len, (int)D);
            if (launch_rotate(ix, xpad, xp, rows, 0)) return RABITQ_ECUDA;  // rabitq.rs:188
            const unsigned g2 = (unsigned)(((K + CD_TC - 1) / CD_TC) * ((rows + CD_QG * CD_TQ - 1) / (CD_QG * CD_TQ)));
            centroid_dist_kernel<<<g2, CD_THREADS>>>(ix->cent, xp, dist, (int)rows, (int)K, (int)D, nullptr);
            argmin_rows_kernel<<<(unsigned)((rows + 3) / 4), 128>>>(dist, rows, (int)K, label + s0, min_dist + s0);
            encode_kernel<<<(unsigned)((rows + 3) / 4), 128>>>(xp, ix->cent, label + s0, min_dist + s0, rows, (int)D, codes_u + s0 * W32,
                                                               fac_u + s0, key_in + s0);
            CU(cudaGetLastError());
        }
    }
    // --- clusters in id order, ascending distance inside, stable (rabitq.rs:231-252): one stable radix sort ---
    iota_kernel<<<(unsigned)((n + 255) / 256), 256>>>(vals_in, n);
    {
        size_t tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, vals_in, vals_out, (int64_t)n, 0, 64));
        void* sort_tmp = nullptr;
        CU(tmp.alloc(&sort_tmp, tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(sort_tmp, tmp_bytes, key_in, key_out, vals_in, vals_out, (int64_t)n, 0, 64));
    }
    CU(cudaMemset(counts, 0, K * 4));
    label_histogram_kernel<<<(unsigned)((n + 255) / 256), 256>>>(label, n, counts);
    offsets_scan_kernel<<<1, 1024>>>(counts, (int)K, ix->offsets);
    CU(cudaMalloc((void**)&ix->base, n * D * 4));
    CU(cudaMalloc((void**)&ix->codes, n * W32 * 4));
    CU(cudaMalloc((void**)&ix->factors, n * 16));
    CU(cudaMalloc((void**)&ix->map_ids, n * 4));
    permute_kernel<<<(unsigned)((n + 3) / 4), 128>>>(d_base_in, (int)len, (int)D, vals_out, codes_u, fac_u, n, ix->base, ix->codes,
                                                     ix->factors, ix->map_ids);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    {
        std::vector<uint32_t> cnt_h(K);
        CU(cudaMemcpy(cnt_h.data(), counts, K * 4, cudaMemcpyDeviceToHost));
        ix->max_cluster = *std::max_element(cnt_h.begin(), cnt_h.end());
    }
    {
        const uint32_t rb[2] = {0u, (uint32_t)n};
        CU(cudaMalloc((void**)&ix->goffsets, (K + 1) * 4));
        CU(cudaMemcpy(ix->goffsets, ix->offsets, (K + 1) * 4, cudaMemcpyDeviceToDevice));
        CU(cudaMalloc((void**)&ix->row_bounds, sizeof(rb)));
        CU(cudaMemcpy(ix->row_bounds, rb, sizeof(rb), cudaMemcpyHostToDevice));
    }
    int rc = finish_index(ix);
    if (rc) return rc;
    guard.p = nullptr;
    *out = ix;
    return RABITQ_OK;
}

// RaBitQ::dump_to_dir (src/rabitq.rs:128-156): the six-file layout, byte for byte.
template <typename T>
bool write_records(const std::string& path, const T* data, size_t rows, size_t cols) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t c = (uint32_t)cols;
    bool ok = true;
    for (size_t r = 0; r < rows && ok; r++)
        ok = std::fwrite(&c, 4, 1, f) == 1 && (cols == 0 || std::fwrite(data + r * cols, sizeof(T), cols, f) == cols);
    return std::fclose(f) == 0 && ok;
}

int dump_impl(rabitq_index* ix, const char* dir) {
    if (!ix || !dir) return fail(RABITQ_EINVAL, "null argument");
    if (ix->shard_count != 1) return fail(RABITQ_EUNSUPPORTED, "dump_to_dir needs an unsharded handle");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    const std::string d(dir);
    ::mkdir(d.c_str(), 0777);  // create_dir_all for the leaf; parents are the caller's business
    const size_t D = ix->D, K = ix->K, n = ix->n, W64 = D / 64;
    {   // base.fvecs: n records of D floats, streamed in chunks
        FILE* f = std::fopen((d + "/base.fvecs").c_str(), "wb");
        if (!f) return fail(RABITQ_EIO, "write base error");
        const size_t CH = std::max<size_t>(1, ((size_t)64 << 20) / (D * 4));
        std::vector<float> buf(CH * D);
        const uint32_t c = (uint32_t)D;
        for (size_t s0 = 0; s0 < n; s0 += CH) {
            const size_t rows = std::min(CH, n - s0);
            CU(cudaMemcpy(buf.data(), ix->base + s0 * D, rows * D * 4, cudaMemcpyDeviceToHost));
            for (size_t r = 0; r < rows; r++) {
                std::fwrite(&c, 4, 1, f);
                std::fwrite(buf.data() + r * D, 4, D, f);
            }
        }
        if (std::fclose(f) != 0) return fail(RABITQ_EIO, "write base error");
    }
    std::vector<float> P(D * D), cent(K * D), cdk(D * K);
    CU(cudaMemcpy(P.data(), ix->P, D * D * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cent.data(), ix->cent, K * D * 4, cudaMemcpyDeviceToHost));
    for (size_t dd = 0; dd < D; dd++)
        for (size_t c = 0; c < K; c++) cdk[dd * K + c] = cent[c * D + dd];  // the matrix is D x K, written row-wise (rabitq.rs:133)
    if (!write_records(d + "/orthogonal.fvecs", P.data(), D, D)) return fail(RABITQ_EIO, "write orthogonal error");
    if (!write_records(d + "/centroids.fvecs", cdk.data(), D, K)) return fail(RABITQ_EIO, "write centroids error");
    std::vector<uint32_t> off(K + 1), ids(n);
    CU(cudaMemcpy(off.data(), ix->offsets, (K + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ids.data(), ix->map_ids, n * 4, cudaMemcpyDeviceToHost));
    {
        FILE* f = std::fopen((d + "/offsets_ids.ivecs").c_str(), "wb");
        if (!f) return fail(RABITQ_EIO, "write offsets_ids error");
        uint32_t c = (uint32_t)(K + 1);
        std::fwrite(&c, 4, 1, f);
        std::fwrite(off.data(), 4, K + 1, f);
        c = (uint32_t)n;
        std::fwrite(&c, 4, 1, f);
        std::fwrite(ids.data(), 4, n, f);
        if (std::fclose(f) != 0) return fail(RABITQ_EIO, "write offsets_ids error");
    }
    std::vector<float> fac(n * 4);
    CU(cudaMemcpy(fac.data(), ix->factors, n * 16, cudaMemcpyDeviceToHost));
    if (!write_records(d + "/factors.fvecs", fac.data(), 1, n * 4)) return fail(RABITQ_EIO, "write factors error");
    std::vector<uint64_t> codes(n * W64);
    CU(cudaMemcpy(codes.data(), ix->codes, n * W64 * 8, cudaMemcpyDeviceToHost));
    if (!write_records(d + "/x_binary_vec.u64vecs", codes.data(), 1, n * W64)) return fail(RABITQ_EIO, "write x_binary_vec error");
    return RABITQ_OK;
}

// ---- scan launch ------------------------------------------------------------------------------------------------
// Geometry of the scan's shared-memory ring for this dimension: NT record tiles (of 8) per consumer pass, `sub` passes per
// stage (records per stage = work-item size = 8 * NT * sub), `stages` ring slots.  Small dimensions take deep rings of big
// stages (the epilogue dominates and every stage re-copies the chunk's codes); large ones shrink until one CTA fits.
// "scan_mode" 0..2 forces NT = 1, 2, 4 (tests run every instantiation).
struct ScanGeom { int nt, sub, stages; };
ScanGeom scan_geom(const rabitq_index* ix) {
    const int D = (int)ix->D;
    // shared memory per CTA that still lets 3 / 2 / 1 CTAs share an SM (227 KB usable, 1 KB reserved per CTA)
    const size_t tiers[3] = {(size_t)74 * 1024, (size_t)112 * 1024, (size_t)227 * 1024};
    // measured on B200 (C1: D=128, C2: D=960): residency beats tile size -- three CTAs of 16 records per stage outrun two of 32 --
    // and a third ring slot only pays where stages are small
    static const int shapes[3][2] = {{2, 3}, {1, 3}, {1, 2}};  // (sub, stages), in order of preference
    ScanGeom g{1, 1, 1};
    bool found = false;
    const int forced_nt = (ix->scan_mode >= 0 && ix->scan_mode <= 2) ? (1 << ix->scan_mode) : 0;
    for (int tier = 0; tier < 3 && !found; tier++)
        for (int nt : {4, 2, 1}) {
            if (forced_nt && nt != forced_nt) continue;
            if (!forced_nt && tier == 0 && nt == 1) continue;  // 8 records per pass rebuild the A fragments too often: rather two CTAs
            for (const auto& sh : shapes)
                if (scan_smem_bytes(D, 8 * nt * sh[0], sh[1]) <= tiers[tier]) { g = ScanGeom{nt, sh[0], sh[1]}; found = true; break; }
            if (found) break;
        }
    if (!found) g = ScanGeom{forced_nt ? forced_nt : 1, 1, 1};  // one stage, no overlap (dim > ~4600)
    const size_t cap = tiers[2];
    if (ix->scan_sub > 0) g.sub = std::min(8, ix->scan_sub);
    if (ix->scan_stages > 0) g.stages = std::min(8, ix->scan_stages);
    while (g.sub > 1 && scan_smem_bytes(D, 8 * g.nt * g.sub, 1) > cap) g.sub--;
    while (g.stages > 1 && scan_smem_bytes(D, 8 * g.nt * g.sub, g.stages) > cap) g.stages--;
    return g;
}
int scan_qs(const rabitq_index* ix) {  // records per stage
    const ScanGeom g = scan_geom(ix);
    return 8 * g.nt * g.sub;
}

template <int NT, bool DENSE, int MINB>
int launch_scan_nt(rabitq_index* ix, ScanArgs& a, const ScanGeom& g) {
    const size_t smem = scan_smem_bytes((int)ix->D, 8 * NT * g.sub, g.stages);
    if (smem > (size_t)227 * 1024) return fail(RABITQ_EUNSUPPORTED, "dim too large for the code scan's shared-memory staging (dim <= 8192)");
    auto kern = scan_mma_kernel<NT, DENSE, MINB>;
    bool& attr_done = ix->scan_attr_done[DENSE ? 1 : 0][MINB == 4 ? 4 : MINB == 3 ? 3 : NT == 4 ? 2 : NT == 2 ? 1 : 0];
    if (!attr_done) {  // once per handle (= per device) and instantiation; always the maximum, other handles share the function
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    int bps = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SCAN_BLOCK, smem));
    bps = std::max(1, bps);
    ix->scan_blocks_per_sm = bps;
    kern<<<ix->sm_count * bps, SCAN_BLOCK, smem, ix->stream>>>(a);
    CU(cudaGetLastError()); ix->counts[5]++;
    return 0;
}

template <bool DENSE>
int launch_scan(rabitq_index* ix, ScanArgs& a) {
    const int D = (int)ix->D;
    const ScanGeom g = scan_geom(ix);
    a.D = D;
    a.rec_pitch = scan_rec_pitch(D);
    a.stages = g.stages;
    a.sub = g.sub;
    if constexpr (!DENSE) {  // three CTAs per SM want <= 128 registers; where shared memory allows two anyway, the compiler gets 200
        if (g.nt == 4 && scan_smem_bytes(D, 8 * g.nt * g.sub, g.stages) <= (size_t)74 * 1024) return launch_scan_nt<4, false, 3>(ix, a, g);
        if (g.nt == 2 && scan_smem_bytes(D, 8 * g.nt * g.sub, g.stages) <= (size_t)55 * 1024) return launch_scan_nt<2, false, 4>(ix, a, g);
    }
    switch (g.nt) {
        case 4: return launch_scan_nt<4, DENSE, 2>(ix, a, g);
        case 2: return launch_scan_nt<2, DENSE, 2>(ix, a, g);
        default: return launch_scan_nt<1, DENSE, 2>(ix, a, g);
    }
}

enum StopAfter { STOP_ROTATE, STOP_PROBE, STOP_QUANT, STOP_SCAN_DENSE, STOP_NONE };

struct BatchOut {  // device pointers of the sub-batch products
    uint32_t P = 0;
    uint32_t total_words = 0;
    uint64_t total_pairs = 0;
    bool speculative = false;  // total_words is the speculative capacity: the real totals arrive with the batch's final synchronisation
};

// ---- pieces of the batch pipeline (shared by the single-GPU path and the distributed phases) ------------------------------
struct Pos { int p, ch; };  // a visit position: (effective probe rank, 128-vector chunk)

// K0-K2b for nb queries already in ix->qraw (nb x len): pad, rotate, centroid distances, probe selection.  `global_view`:
// pair counts and the first non-empty rank refer to the whole index (distributed front end) instead of this shard's rows.
int run_front_select(rabitq_index* ix, size_t nb, int P, bool global_view);

int run_front(rabitq_index* ix, size_t nb, size_t len, int P, bool stop_after_rotate, bool global_view) {
    const int D = (int)ix->D;
    cudaStream_t st = ix->stream;
    ix->state_reset_done = false;
    CU(ix->y.ensure(nb * D * 4));
    if (!ix->q_in) ix->q_in = ix->qraw.as<float>();
    if ((int)len == D) {
        ix->q_pad = ix->q_in;  // src/rabitq.rs:277-280 pads only when the query is shorter than the index dimension
    } else {
        CU(ix->qpad.ensure(nb * D * 4));
        size_t tot = nb * (size_t)D;
        pad_queries_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ix->q_in, ix->qpad.as<float>(), nb, (int)len, D);
        CU(cudaGetLastError()); ix->counts[5]++;
        ix->q_pad = ix->qpad.as<float>();
    }
    if (tick(ix, ST_H2D)) return RABITQ_ECUDA;
    ix->y_all = ix->y.as<float>();
    if (launch_rotate(ix, ix->q_pad, ix->y.as<float>(), nb, st)) return RABITQ_ECUDA;
    ix->counts[5]++;
    if (tick(ix, ST_ROTATE)) return RABITQ_ECUDA;
    if (stop_after_rotate) return 0;
    return run_front_select(ix, nb, P, global_view);
}

// K2p / K2 / K2b for nb rotated queries in ix->y (the second half of the front end)
int run_front_select(rabitq_index* ix, size_t nb, int P, bool global_view) {
    const int D = (int)ix->D, K = (int)ix->K;
    cudaStream_t st = ix->stream;
    CU(ix->tot_blk.ensure(32));
    CU(cudaMemsetAsync(ix->tot_blk.p, 0, 32, st));  // totals, speculative-sizing flag, prefilter flag + statistic of this batch
    CU(ix->cdist.ensure(nb * (size_t)K * 4));
    CU(ix->probe_ids.ensure(nb * P * 4));
    CU(ix->probe_dist.ensure(nb * P * 4));
    CU(ix->slot_local.ensure(nb * P * 4));
    CU(ix->q_words.ensure(nb * 4));
    CU(ix->q_pairs.ensure(nb * 4));
    CU(ix->q_p0.ensure(nb * 4));
    const bool pf = use_prefilter(ix, P);
    const uint32_t* run_if = nullptr;
    if (pf) {
        // tensor-core prefilter (prefilter.cuh): approximate keys for all K, exact distances for the few candidates
        CU(ix->pf_yhat.ensure(nb * (size_t)D * 4));
        CU(ix->pf_ynorm.ensure(nb * 4));

        const bool split = ix->pf_mode == 3;
        if (split) CU(ix->pf_yhat_lo.ensure(nb * (size_t)D * 4));
        query_center_kernel<<<(unsigned)((nb + 3) / 4), 128, 0, st>>>(ix->y.as<float>(), ix->pf_mu, (int)nb, D, ix->pf_yhat.as<float>(),
                                                                      split ? ix->pf_yhat_lo.as<float>() : nullptr, ix->pf_ynorm.as<float>());
        CU(cudaGetLastError()); ix->counts[5]++;
        const size_t tiles128 = ((nb + 127) / 128) * (size_t)((K + 127) / 128);
        // tcgen05 form: tensor maps of the operands (centroid side once per index, query side per batch: a host-side encode)
        bool tc5 = ix->pf_gemm == 1;
        if (tc5 && !ix->tm_c_ready) {
            if (tc5_make_tmap(&ix->tm_chat, ix->pf_chat, (size_t)K, (size_t)D, TC5_BN) || tc5_make_tmap(&ix->tm_chat_lo, ix->pf_chat_lo, (size_t)K, (size_t)D, TC5_BN)) {
                ix->pf_gemm = 0;  // no tensor-map encoder in this driver: the mma.sync form computes the same keys
                tc5 = false;
            } else {
                ix->tm_c_ready = true;
            }
        }
        if (tc5) {
            if (tc5_make_tmap(&ix->tm_yhat, ix->pf_yhat.as<float>(), nb, (size_t)D, TC5_BM) ||
                (split && tc5_make_tmap(&ix->tm_yhat_lo, ix->pf_yhat_lo.as<float>(), nb, (size_t)D, TC5_BM)))
                return fail(RABITQ_ECUDA, "cuTensorMapEncodeTiled failed for the query operand");
        }
        auto gemm = [&](const float* ya, const float* cb, int accumulate) {
            if (tc5) {
                dim3 grid((K + TC5_BN - 1) / TC5_BN, (unsigned)((nb + TC5_BM - 1) / TC5_BM));
                approx_gemm_tc5_kernel<<<grid, TC5_THREADS, TC5_SMEM_BYTES, st>>>(ya == ix->pf_yhat.as<float>() ? ix->tm_yhat : ix->tm_yhat_lo,
                                                                                  cb == ix->pf_chat ? ix->tm_chat : ix->tm_chat_lo, ix->pf_cnorm2,
                                                                                  (int)nb, K, D, ix->cdist.as<float>(), accumulate);
            } else if (tiles128 >= (size_t)ix->sm_count) {
                dim3 grid((K + 127) / 128, (unsigned)((nb + 127) / 128));
                approx_gemm_tf32_kernel<128, 128, 2, 4><<<grid, 256, 2 * 256 * PF_PITCH * 4, st>>>(ya, cb, ix->pf_cnorm2, (int)nb, K, D,
                                                                                                    ix->cdist.as<float>(), accumulate);
            } else {
                dim3 grid((K + 63) / 64, (unsigned)((nb + 63) / 64));
                approx_gemm_tf32_kernel<64, 64, 2, 2><<<grid, 128, 2 * 128 * PF_PITCH * 4, st>>>(ya, cb, ix->pf_cnorm2, (int)nb, K, D,
                                                                                                  ix->cdist.as<float>(), accumulate);
            }
            ix->counts[5]++;
        };
        gemm(ix->pf_yhat.as<float>(), ix->pf_chat, 0);
        if (split) {  // 3xTF32: hi*hi + hi*lo + lo*hi
            gemm(ix->pf_yhat.as<float>(), ix->pf_chat_lo, 1);
            gemm(ix->pf_yhat_lo.as<float>(), ix->pf_chat, 1);
        }
        CU(cudaGetLastError());
        if (tick(ix, ST_CDIST)) return RABITQ_ECUDA;
        // candidate capacity: 8 per probe (at least 512); the arrays live in shared memory, and a smaller CTA footprint means
        // more queries per SM in flight
        const int ps_cap = std::max(2, std::min(std::min(ix->prefilter_cap, PS_CAP), std::max(K >= 32768 ? 1024 : 512, 8 * P)) & ~1);
        const size_t ps_smem = (size_t)D * 4 + (size_t)ps_cap * 12 + (size_t)P * 4;
#define PS_ARGS ix->cdist.as<float>(), ix->pf_ynorm.as<float>(), ix->pf_cnorm, ix->pf_cnorm_max, ix->y.as<float>(), ix->cent, K, P, D, ix->offsets, \
            global_view ? ix->goffsets : nullptr, ix->probe_ids.as<uint32_t>(), ix->probe_dist.as<float>(), ix->slot_local.as<uint32_t>(), \
            ix->q_words.as<uint32_t>(), ix->q_pairs.as<uint32_t>(), ix->q_p0.as<uint32_t>(), ix->tot_blk.as<uint32_t>() + 6, ps_cap, \
            split ? (float)(2 * D + 64) * 4.76837158203125e-07f : 0.00390625f, (float)(D / 8 + 8) * 1.1920928955078125e-07f
        // one CTA per query; 128 threads instead of 256 when 256-thread CTAs would need several waves (the kernel is a chain of
        // latencies: twice the queries in flight per SM beat twice the threads per query)
        if (nb > (size_t)ix->sm_count * 8 && ix->pf_threads != 256) prefilter_select_kernel<128><<<(unsigned)nb, 128, ps_smem, st>>>(PS_ARGS);
        else prefilter_select_kernel<256><<<(unsigned)nb, 256, ps_smem, st>>>(PS_ARGS);
#undef PS_ARGS
        ix->pf_last_nb = nb;
        ix->pf_pending = true;
        CU(cudaGetLastError()); ix->counts[5]++;
        run_if = ix->tot_blk.as<uint32_t>() + 6;  // the classic kernels below run only if some query could not be certified
    }
    {
        const size_t tiles = (size_t)((K + CD_TC - 1) / CD_TC) * ((nb + CD_QG * CD_TQ - 1) / (CD_QG * CD_TQ));
        const unsigned grid = (unsigned)std::min<size_t>(tiles, run_if ? (size_t)ix->sm_count * 2 : ((size_t)1 << 30));
        centroid_dist_kernel<<<grid, CD_THREADS, 0, st>>>(ix->cent, ix->y.as<float>(), ix->cdist.as<float>(), (int)nb, K, D, run_if);
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    if (!pf && tick(ix, ST_CDIST)) return RABITQ_ECUDA;
    CU(ix->probe_ids.ensure(nb * P * 4));
    CU(ix->probe_dist.ensure(nb * P * 4));
    CU(ix->slot_local.ensure(nb * P * 4));
    CU(ix->q_words.ensure(nb * 4));
    CU(ix->q_pairs.ensure(nb * 4));
    CU(ix->q_p0.ensure(nb * 4));
    CU(ix->q_wbase.ensure((nb + 1) * 4));
    CU(ix->q_pbase.ensure((nb + 1) * 8));
    {
        if (P > 4096) {  // no cap in the reference (src/rabitq.rs:294): full sort of the K keys in a global scratch row per query
            int Kpow2 = 1;
            while (Kpow2 < K) Kpow2 <<= 1;
            CU(ix->sel_scratch.ensure(nb * (size_t)Kpow2 * 8));
            select_probe_large_kernel<<<(unsigned)nb, SEL_THREADS, 0, st>>>(
                ix->cdist.as<float>(), K, Kpow2, P, ix->offsets, global_view ? ix->goffsets : nullptr, ix->sel_scratch.as<unsigned long long>(),
                ix->probe_ids.as<uint32_t>(), ix->probe_dist.as<float>(), ix->slot_local.as<uint32_t>(), ix->q_words.as<uint32_t>(),
                ix->q_pairs.as<uint32_t>(), ix->q_p0.as<uint32_t>());
            CU(cudaGetLastError()); ix->counts[5]++;
        } else {
        int Ppow2 = 1;
        while (Ppow2 < P) Ppow2 <<= 1;
        const bool pivot_ok = K >= 16 * P && K >= 4 * SEL_SAMPLE;  // kernels.cuh: sampling-pivot path, keys stay in L2/HBM
        const int cache_keys = (!pivot_ok && K <= 16384) ? 1 : 0;  // radix path: keys of one query in shared memory (<= 64 KB)
        const size_t sel_smem = (size_t)std::max(Ppow2, SEL_CAND) * 8 + SEL_SAMPLE * 4 + (cache_keys ? (size_t)K * 4 : 0);
        select_probe_kernel<<<(unsigned)nb, SEL_THREADS, sel_smem, st>>>(
            ix->cdist.as<float>(), K, P, Ppow2, cache_keys, ix->offsets, global_view ? ix->goffsets : nullptr, ix->probe_ids.as<uint32_t>(),
            ix->probe_dist.as<float>(), ix->slot_local.as<uint32_t>(), ix->q_words.as<uint32_t>(), ix->q_pairs.as<uint32_t>(),
            ix->q_p0.as<uint32_t>(), run_if);
        CU(cudaGetLastError()); ix->counts[5]++;
        }
        query_base_scan_kernel<<<1, 1024, 0, st>>>(ix->q_words.as<uint32_t>(), ix->q_pairs.as<uint32_t>(), (int)nb,
                                                   ix->q_wbase.as<uint32_t>(), ix->q_pbase.as<unsigned long long>(),
                                                   ix->spec_on ? ix->spec_cap : 0u, ix->spec_on ? ix->tot_blk.as<uint32_t>() + 1 : nullptr,
                                                   ix->tot_blk.as<uint32_t>());
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    return 0;
}

// totals of the slot layout to the host (they size the survivor slots).  post_totals enqueues the copy, wait_totals blocks
// the host until it has landed: the caller launches K3 in between, so the GPU is never idle while the host sleeps.
int post_totals(rabitq_index* ix, size_t nb) {
    cudaStream_t st = ix->stream;
    // one 32-byte copy of the totals block; a speculatively sized batch does not need it before its end (query_batch_impl copies it
    // with the results), so nothing sits between the probe selection and K3 on the main stream
    if (!ix->spec_on) CU(cudaMemcpyAsync(ix->h_pin, ix->tot_blk.p, 32, cudaMemcpyDeviceToHost, st));
    if (!ix->ev_totals) CU(cudaEventCreateWithFlags(&ix->ev_totals, cudaEventDisableTiming));
    CU(cudaEventRecord(ix->ev_totals, st));
    CU(cudaEventRecord(ix->ev_fork, st));
    if (tick(ix, ST_SELECT)) return RABITQ_ECUDA;
    return 0;
}

int wait_totals(rabitq_index* ix, BatchOut* bo) {
    CU(cudaEventSynchronize(ix->ev_totals));
    prefilter_adapt(ix);
    bo->total_words = ix->h_pin[0];
    std::memcpy(&bo->total_pairs, ix->h_pin + 2, 8);
    return 0;
}

// records per work item of the scan = per shared-memory stage; "scan_slices" > 1 cuts items smaller (partial stages: a test knob)
uint32_t scan_ms(const rabitq_index* ix) { return (uint32_t)std::max(1, scan_qs(ix) / std::max(1, ix->scan_slices)); }

// K3 over the inverted list of round `set` (built on the side stream: the main stream waits for it here); on a shard, items
// whose cluster lives elsewhere are skipped
int run_quantize_list(rabitq_index* ix, int P, size_t set, bool on_aux) {
    const int D = (int)ix->D, W32 = D / 32, RS = (D + REC_META_BYTES) / 4;  // record words: D code bytes + scalars
    // the first round's records on the main stream (its scan comes next); later rounds' on the side stream, right behind their
    // list: they are written while the first round's scan and (latency-bound, SM-starved) replay run
    cudaStream_t st = on_aux ? ix->aux_stream : ix->stream;
    ListSet& L = ix->lists[set];
    const int pitch = scan_rec_pitch(D);
    CU(L.qrec.ensure(std::max<size_t>(L.cap_items, 1) * (size_t)pitch));
    L.q_on_aux = on_aux;
    if (!on_aux) CU(cudaStreamWaitEvent(st, L.ready, 0));
    const uint32_t* skip = ix->shard_count > 1 ? ix->offsets : nullptr;
    if (L.cap_items == 0) { L.q_on_aux = false; return 0; }
#define QUANT_COMMON L.cl_items.as<uint2>(), L.cl_start.as<uint32_t>() + ix->K, ix->probe_dist.as<float>(), ix->slot_local.as<uint32_t>(), \
                     ix->q_wbase.as<uint32_t>(), skip, ix->quant_bias, L.qrec.as<unsigned char>(), pitch, P
    if (W32 == 2 || W32 == 4 || W32 == 6 || W32 == 8) {
        // eight lanes per record; consecutive list slots per group (centroid in registers across the slots of one cluster): as
        // many as keep >= ~48 warps per SM in the grid
        int pch = 1;
        while (pch < 8 && L.cap_items / (size_t)(8 * pch) >= (size_t)ix->sm_count * 48) pch *= 2;
        const size_t groups = (L.cap_items + pch - 1) / pch;
        const unsigned qgrid = (unsigned)((groups + 15) / 16);
        const size_t qsmem = (size_t)16 * RS * 4;  // one record per group, staged for the coalesced store
        switch (W32) {
            case 2: quantize_small_kernel<2><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, pch); break;
            case 4: quantize_small_kernel<4><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, pch); break;
            case 6: quantize_small_kernel<6><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, pch); break;
            default: quantize_small_kernel<8><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, pch); break;
        }
    } else {
        // a warp per record; up to dim 1536 the residual stays in registers (one pass over the two rows), beyond that two passes
        const int pch = 1;
        const size_t warps = L.cap_items;
        const unsigned qgrid = (unsigned)((warps + 3) / 4);
        const size_t qsmem = (size_t)4 * RS * 4;  // one record per warp, staged for the coalesced store
        switch (W32) {
#define QK(w) case w: quantize_kernel<w><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, D, pch); break;
            QK(10) QK(12) QK(14) QK(16) QK(20) QK(24) QK(30) QK(32) QK(40) QK(48)
#undef QK
            default: quantize_kernel<0><<<qgrid, 128, qsmem, st>>>(ix->y_all, ix->cent, QUANT_COMMON, D, pch); break;
        }
    }
#undef QUANT_COMMON
    if (on_aux) {
        if (!L.qready) CU(cudaEventCreateWithFlags(&L.qready, cudaEventDisableTiming));
        CU(cudaEventRecord(L.qready, st));
    }
    CU(cudaGetLastError()); ix->counts[5]++;
    return 0;
}

// work buffers + kernel argument blocks of the scan / rerank rounds
int setup_rounds(rabitq_index* ix, size_t nb, int P, size_t topk, const BatchOut* bo, ScanArgs* sa_out, RerankArgs* ra_out) {
    const int D = (int)ix->D, K = (int)ix->K;
    cudaStream_t st = ix->stream;
    // survivor slots: one bitmap word + 32 (rough, j) entries per 32 vectors of every probed cluster
    const size_t words = std::max<uint32_t>(bo->total_words, 1);
    CU(ix->bitmap.ensure(words * 4));
    CU(ix->entries.ensure(words * 32 * 8));
    const uint32_t MS = scan_ms(ix);
    ix->max_items = ix->n / SCAN_THREADS + (size_t)K + 2 + ((size_t)bo->total_words / 4 + nb * (size_t)P) / MS;
    CU(ix->thr.ensure(nb * 4));
    CU(ix->heap_dist.ensure(nb * topk * 4));
    CU(ix->heap_ids.ensure(nb * topk * 4));
    CU(ix->heap_cnt.ensure(nb * 4));
    CU(ix->q_precise.ensure(nb * 4));
    CU(ix->h_recent.ensure(nb * 4));
    CU(ix->h_wcount.ensure(nb * 4));
    CU(ix->counters.ensure(64));
    CU(ix->out_all.ensure(nb * topk * 8 + nb * 4));  // [dist nb x topk | ids nb x topk | count nb]: one D2H for the three results
    if (!ix->state_reset_done) {
        CU(cudaMemsetAsync(ix->counters.p, 0, 64, st));
        fill_f32_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(ix->thr.as<float>(), nb, 3.402823466e+38f);
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    ix->state_reset_done = false;

    ScanArgs sa;
    sa.scan_codes = ix->scan_codes;
    sa.scan_fac = ix->scan_fac;
    sa.qrec = nullptr;  // per round (run_round_scan)
    sa.thr = ix->thr.as<float>();
    sa.q_p0 = ix->q_p0.as<uint32_t>();
    sa.bitmap = ix->bitmap.as<uint32_t>();
    sa.entries = ix->entries.as<float2>();
    sa.counters = ix->counters.as<unsigned long long>();
    sa.P = P;
    sa.MS = MS;

    RerankArgs ra;
    std::memset(&ra, 0, sizeof(ra));
    ra.qpad = ix->q_pad;
    ra.base = ix->base;
    ra.map_ids = ix->map_ids;
    ra.q_wbase = ix->q_wbase.as<uint32_t>();
    ra.slot_local = ix->slot_local.as<uint32_t>();
    ra.q_p0 = ix->q_p0.as<uint32_t>();
    ra.bitmap = ix->bitmap.as<uint32_t>();
    ra.entries = ix->entries.as<float2>();
    ra.heap_dist = ix->heap_dist.as<float>();
    ra.heap_ids = ix->heap_ids.as<uint32_t>();
    ra.heap_cnt = ix->heap_cnt.as<uint32_t>();
    ra.thr = ix->thr.as<float>();
    ra.q_precise = ix->q_precise.as<uint32_t>();
    ra.h_recent = ix->h_recent.as<float>();
    ra.h_wcount = ix->h_wcount.as<uint32_t>();
    ra.counters = ix->counters.as<unsigned long long>();
    ra.out_dist = ix->out_all.as<float>();
    ra.out_ids = ix->out_all.as<uint32_t>() + nb * topk;
    ra.out_count = ix->out_all.as<uint32_t>() + 2 * nb * topk;
    if (ix->ovr_dist) { ra.out_dist = ix->ovr_dist; ra.out_ids = ix->ovr_ids; if (ix->ovr_count) ra.out_count = ix->ovr_count; }
    ra.nq = (int)nb;
    ra.P = P;
    ra.D = D;
    ra.topk = (int)topk;
    // rows per wave (two wave buffers per warp): small waves keep shared memory low and occupancy high; the gathers are
    // issued at enqueue time, so a wave's rows are usually resident before it is replayed.  When the batch is small enough
    // the wave is shrunk further until every query-warp is co-resident (one wave of CTAs instead of two).
    auto rr_smem = [&](int R) {
        return (int)((16 + (size_t)D * 4 + (size_t)2 * R * (D + 8) * 4 + 2 * topk * 4 + 2 * 64 * 4 + 128 * 8 + 16 + 15) / 16 * 16);
    };
    auto rr_warps_per_block = [&](int smem) { return smem > 12 * 1024 ? 1 : 4; };
    auto rr_resident = [&](int R) {
        const int smem = rr_smem(R), wpb = rr_warps_per_block(smem);
        const int by_smem = (int)((227 * 1024) / ((size_t)wpb * smem + 1024)), by_warps = 32 / wpb;
        return (size_t)std::max(1, std::min(std::min(by_smem, by_warps), 32)) * wpb * ix->sm_count;
    };
    ra.R = (int)std::max<size_t>(2, std::min<size_t>(8, 16384 / ((size_t)D * 4)));
    if (ix->rerank_rows > 0) ra.R = std::max(1, std::min(32, ix->rerank_rows));
    else
        for (int R = ra.R; R >= 2; R--)
            if (rr_resident(R) >= nb) { ra.R = R; break; }
    {   // <= 80 KB of row buffers per warp, and the whole per-warp state (query row, k-slot result buffer, rows) inside 200 KB
        const size_t fixed = (size_t)D * 4 + 2 * topk * 4 + 2048;
        const size_t budget = std::min<size_t>(80 * 1024, (size_t)200 * 1024 > fixed ? (size_t)200 * 1024 - fixed : 0);
        ra.R = (int)std::max<size_t>(1, std::min<size_t>(ra.R, budget / (2 * (size_t)(D + 8) * 4)));
    }
    ra.smem_per_warp = rr_smem(ra.R);
    ra.prefetch = ix->rerank_prefetch;
    ra.win = nullptr;
    ra.ns = 2;
    if (ix->debug_rerank) {
        CU(ix->rr_dbg.ensure(nb * 16 * 4));
        CU(cudaMemsetAsync(ix->rr_dbg.p, 0, nb * 16 * 4, st));
        ra.dbg = ix->rr_dbg.as<uint32_t>();
    }
    *sa_out = sa;
    *ra_out = ra;
    return 0;
}

enum RoundKind { ROUND_REPLAY, ROUND_SINK1, ROUND_SINK2 };

// inverted probe list (cluster -> records) of one window [lo, hi) of visit positions into list set `set`, on stream `st`;
// also counts the window's scan work items.  Needs nothing but K2b's products: built for every round before K3.
int build_lists(rabitq_index* ix, size_t nb, int P, uint32_t MS, Pos lo, Pos hi, size_t set, cudaStream_t st) {
    const int K = (int)ix->K;
    if (ix->lists.size() <= set) ix->lists.resize(set + 1);
    ListSet& L = ix->lists[set];
    const int p_lo = lo.p, p_hi_incl = std::min(P, hi.p + (hi.ch > 0 ? 1 : 0));  // ranks that have items in this round
    const bool single_rank = p_hi_incl == p_lo + 1;
    const uint32_t ch_min = single_rank ? (uint32_t)lo.ch : 0u;
    const uint32_t ch_max = (single_rank && hi.ch > 0) ? (uint32_t)hi.ch : 0xffffffffu;
    const size_t items = nb * (size_t)(p_hi_incl - p_lo);
    const uint32_t* foreign = ix->shard_count > 1 ? ix->offsets : nullptr;  // on a shard, clusters that live elsewhere get no items
    L.cap_items = items;
    L.MS = MS;
    L.ch_min = ch_min;
    CU(L.cl_count.ensure((size_t)K * 4));
    CU(L.cl_start.ensure((size_t)(K + 1) * 4));
    CU(L.item_start.ensure((size_t)(K + 1) * 4));
    CU(L.cl_cursor.ensure((size_t)K * 4));
    CU(L.cl_items.ensure(std::max<size_t>(items, 1) * 8));
    CU(L.work_ctl.ensure(16));
    if (!L.ready) CU(cudaEventCreateWithFlags(&L.ready, cudaEventDisableTiming));
    CU(cudaMemsetAsync(L.cl_count.p, 0, (size_t)K * 4, st));
    if (items) {
        bucket_count_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(ix->probe_ids.as<uint32_t>(), ix->q_p0.as<uint32_t>(), foreign, nb, P, p_lo, p_hi_incl,
                                                                              L.cl_count.as<uint32_t>());
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    bucket_scan_kernel<<<1, 1024, 0, st>>>(L.cl_count.as<uint32_t>(), ix->offsets, K, SCAN_THREADS, MS, ch_min, ch_max,
                                           L.cl_start.as<uint32_t>(), L.item_start.as<uint32_t>(), L.cl_cursor.as<uint32_t>(),
                                           L.work_ctl.as<uint32_t>(), ix->spec_on ? ix->tot_blk.as<uint32_t>() + 1 : nullptr);
    CU(cudaGetLastError()); ix->counts[5]++;
    if (items) {
        bucket_fill_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(ix->probe_ids.as<uint32_t>(), ix->q_p0.as<uint32_t>(), foreign, nb, P, p_lo, p_hi_incl,
                                                                             L.cl_start.as<uint32_t>(), L.cl_cursor.as<uint32_t>(),
                                                                             L.cl_items.as<uint2>());
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    CU(cudaEventRecord(L.ready, st));
    return 0;
}

// scan of one window [lo, hi) of visit positions over list set `set` (its inverted list and records exist: build_lists +
// run_quantize_list).  The flat work items are written here: their buffer is sized from the slot totals (setup_rounds).
int run_round_scan(rabitq_index* ix, size_t nb, int P, ScanArgs& sa, Pos lo, Pos hi, bool dense, size_t set = 0) {
    cudaStream_t st = ix->stream;
    const int K = (int)ix->K;
    ListSet& L = ix->lists[set];
    CU(cudaStreamWaitEvent(st, L.ready, 0));
    if (L.q_on_aux) CU(cudaStreamWaitEvent(st, L.qready, 0));
    if (L.w_on_aux) {
        CU(cudaStreamWaitEvent(st, L.wready, 0));
        L.w_on_aux = false;
    } else {
        CU(L.work.ensure(ix->max_items * sizeof(ScanItem)));
        work_items_kernel<<<(K + 255) / 256, 256, 0, st>>>(L.item_start.as<uint32_t>(), L.cl_count.as<uint32_t>(), L.cl_start.as<uint32_t>(), ix->offsets,
                                                           ix->chunk_start, K, L.MS, L.ch_min, L.work.as<ScanItem>(),
                                                           ix->spec_on ? ix->tot_blk.as<uint32_t>() + 1 : nullptr);
        CU(cudaGetLastError()); ix->counts[5]++;
    }
    sa.cl_items = L.cl_items.as<uint2>();
    sa.work = L.work.as<ScanItem>();
    sa.work_ctl = L.work_ctl.as<uint32_t>();
    sa.qrec = L.qrec.as<unsigned char>();
    sa.MS = L.MS;
    if (tick(ix, ST_BUCKET)) return RABITQ_ECUDA;
    sa.p_lo = lo.p; sa.ch_lo = lo.ch; sa.p_hi = hi.p; sa.ch_hi = hi.ch;
    int rc = dense ? launch_scan<true>(ix, sa) : launch_scan<false>(ix, sa);
    if (rc) return rc;
    if (tick(ix, ST_SCAN)) return RABITQ_ECUDA;
    ix->counts[4] += 1;
    return 0;
}

int run_round_rerank(rabitq_index* ix, size_t nb, const RerankArgs& ra, Pos lo, Pos hi, bool is_first, bool is_last, bool heuristic,
                     RoundKind kind) {
    cudaStream_t st = ix->stream;
    const int rr_wpb = ra.smem_per_warp > 12 * 1024 ? 1 : 4;
    const dim3 rgrid((unsigned)((nb + rr_wpb - 1) / rr_wpb)), rblock(rr_wpb * 32);
    const size_t rsmem = (size_t)rr_wpb * ra.smem_per_warp;
    const int f = is_first ? 1 : 0, l = is_last ? 1 : 0;
    if ((ix->rerank_mode == 1 && (kind != ROUND_SINK1 || ix->dist_sink_cta)) || kind == ROUND_SINK2) {
        // one CTA per query: producer + replay + compute warps (rerank_cta_kernel).  A wave is one compute warp's job (4 rows, or 8
        // with two candidates per eight-lane group for short rows); `ns` row buffers keep that many gathers in flight
        const size_t D = (size_t)ra.D;
        int R = ix->rerank_rows > 0 ? std::min(8, ix->rerank_rows) : (D <= 256 ? 8 : 4);
        if (ix->rerank_nc == 1) R = std::min(R, 4);
        const int nc = R > 4 || ix->rerank_nc == 2 ? 2 : 1;
        // a row buffer is always consumed by the same compute warp (waves go round-robin over the warps, buffers round-robin over
        // the waves: ns must be a multiple of ncw, or a warp would wait on a barrier whose phases it has not followed)
        int ncw = std::max(1, std::min(8, ix->rerank_warps > 0 ? ix->rerank_warps : (nc == 2 ? 1 : (D > 1024 ? 2 : 3))));  // (three CTAs per SM stay resident)
        int bpw = std::max(1, (std::min(8, ix->rerank_stages > 0 ? ix->rerank_stages : (nc == 2 || D > 1024 ? 2 : 3)) + ncw - 1) / ncw);  // buffers per warp
        while (ncw * bpw > 8) bpw--;
        if (bpw < 1) { bpw = 1; ncw = 8; }
        int ns = ncw * bpw;
        auto need = [&]() { return rerank_cta_smem((int)D, ra.topk, R, ns, nc) + (kind == ROUND_SINK2 ? (size_t)ra.P * 4 : 0); };
        while (need() > (size_t)200 * 1024 && bpw > 1) { bpw--; ns = ncw * bpw; }
        while (need() > (size_t)200 * 1024 && ncw > 1) { ncw--; ns = ncw * bpw; }
        while (need() > (size_t)200 * 1024 && R > 1) R--;
        const size_t smem = need();
        if (smem <= (size_t)200 * 1024) {
            RerankArgs rc = ra;
            rc.R = R;
            rc.ns = ns;
            const dim3 grid((unsigned)nb), block((unsigned)(2 + ncw) * 32);
            if (kind == ROUND_SINK2) {  // distributed frozen round, source side: sequential local filter, records to the home ranks' inboxes
                if (nc == 2) rerank_cta_kernel<false, 2, 2><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
                else rerank_cta_kernel<false, 1, 2><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
            } else if (kind == ROUND_SINK1) {  // distributed round 1: records go to the home ranks' inboxes
                if (nc == 2) rerank_cta_kernel<false, 2, 1><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
                else rerank_cta_kernel<false, 1, 1><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
            } else if (heuristic) {
                if (nc == 2) rerank_cta_kernel<true, 2><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
                else rerank_cta_kernel<true, 1><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
            } else {
                if (nc == 2) rerank_cta_kernel<false, 2><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
                else rerank_cta_kernel<false, 1><<<grid, block, smem, st>>>(rc, lo.p, lo.ch, hi.p, hi.ch, f, l);
            }
            CU(cudaGetLastError()); ix->counts[5]++;
            if (tick(ix, ST_RERANK)) return RABITQ_ECUDA;
            return 0;
        }
    }
    if (kind == ROUND_SINK2) return fail(RABITQ_EUNSUPPORTED, "the source-side sequential filter does not fit shared memory at this dim / topk / probe");
    if (kind == ROUND_SINK1) rerank_kernel<false, 1><<<rgrid, rblock, rsmem, st>>>(ra, lo.p, lo.ch, hi.p, hi.ch, f, l);
    else if (heuristic) rerank_kernel<true, 0><<<rgrid, rblock, rsmem, st>>>(ra, lo.p, lo.ch, hi.p, hi.ch, f, l);
    else rerank_kernel<false, 0><<<rgrid, rblock, rsmem, st>>>(ra, lo.p, lo.ch, hi.p, hi.ch, f, l);
    CU(cudaGetLastError()); ix->counts[5]++;
    if (tick(ix, ST_RERANK)) return RABITQ_ECUDA;
    return 0;
}

// the next batch's queries (rabitq_query_batch_pipelined): on the copy stream, concurrent with this batch's kernels
int start_staging(rabitq_index* ix) {
    const float* src = ix->stage_next;
    ix->stage_next = nullptr;
    CU(ix->qstage.ensure(ix->stage_bytes));
    CU(cudaMemcpyAsync(ix->qstage.p, src, ix->stage_bytes, cudaMemcpyHostToDevice, ix->copy_stream));
    CU(cudaEventRecord(ix->ev_staged, ix->copy_stream));
    ix->staged_src = src;
    return 0;
}

// round windows of visit positions (probe rank, 128-vector chunk): the first round covers only the first `first_chunks` chunks
// of the nearest cluster, so that everything after it is filtered with a real threshold
std::vector<Pos> round_bounds(const rabitq_index* ix, int P, bool dense) {
    std::vector<Pos> bounds;
    if (dense) return {{0, 0}, {P, 0}};
    bounds.push_back({0, 0});
    if (ix->first_chunks > 0) bounds.push_back({0, ix->first_chunks});
    for (uint32_t r : ix->rounds)
        if ((int)r > 0 && (int)r < P && (int)r > bounds.back().p) bounds.push_back({(int)r, 0});
    bounds.push_back({P, 0});
    return bounds;
}

// inverted lists of every round on the side stream (they depend on K2b only), K3 over each list on the main stream as soon as
// the list exists
int run_lists_and_quantize(rabitq_index* ix, size_t nb, int P, const std::vector<Pos>& bounds) {
    const uint32_t MS = scan_ms(ix);
    CU(cudaStreamWaitEvent(ix->aux_stream, ix->ev_fork, 0));
    {   // second side stream: the batch's thresholds / counters are reset and the word windows of every round written (K5 reads one
        // pair per query instead of resolving the bounds through three dependent loads) -- none of it in the main stream's chain
        cudaStream_t a2 = ix->use_aux2 ? ix->aux2_stream : ix->stream;
        CU(cudaStreamWaitEvent(a2, ix->ev_fork, 0));
        CU(ix->thr.ensure(nb * 4));
        CU(ix->counters.ensure(64));
        CU(cudaMemsetAsync(ix->counters.p, 0, 64, a2));
        fill_f32_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, a2>>>(ix->thr.as<float>(), nb, 3.402823466e+38f);
        CU(cudaGetLastError()); ix->counts[5]++;
        ix->state_reset_done = true;
        RoundBounds rb;
        rb.n = (int)bounds.size();
        for (int i = 0; i < rb.n; i++) { rb.p[i] = bounds[i].p; rb.ch[i] = bounds[i].ch; }
        CU(ix->round_win.ensure((size_t)(rb.n - 1) * nb * 8));
        round_windows_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, a2>>>(ix->q_wbase.as<uint32_t>(), ix->slot_local.as<uint32_t>(),
                                                                                        ix->q_p0.as<uint32_t>(), (int)nb, P, rb, ix->round_win.as<uint2>(),
                                                               ix->spec_on ? ix->tot_blk.as<uint32_t>() + 1 : nullptr);
        CU(cudaGetLastError()); ix->counts[5]++;
        CU(cudaEventRecord(ix->ev_win, a2));
    }
    int rc;
    static const bool k3_aux = std::getenv("RABITQ_K3_MAIN") == nullptr;
    for (size_t r = 0; r + 1 < bounds.size(); r++) {
        if ((rc = build_lists(ix, nb, P, MS, bounds[r], bounds[r + 1], r, ix->aux_stream))) return rc;
        if (r > 0 && k3_aux && (rc = run_quantize_list(ix, P, r, true))) return rc;
    }
    if (tick(ix, ST_BUCKET)) return RABITQ_ECUDA;
    for (size_t r = 0; r + 1 < bounds.size(); r++)
        if ((r == 0 || !k3_aux) && (rc = run_quantize_list(ix, P, r, false))) return rc;
    if (tick(ix, ST_QUANT)) return RABITQ_ECUDA;
    return 0;
}

// One sub-batch of nb queries, already on the device in ix->qraw (nb x len).  Runs up to `stop`.
int run_sub_batch(rabitq_index* ix, size_t nb, size_t len, size_t probe, size_t topk, bool heuristic, StopAfter stop, BatchOut* bo,
                  bool allow_spec = false) {
    const int P = (int)std::min(probe, ix->K);
    bo->P = P;
    // Speculative slot sizing: the survivor slots (one bitmap word + 32 entries per 32 vectors of every probed cluster) are sized from
    // the largest words-per-query any earlier batch needed (+25 %), so the host does not wait for this batch's totals between the
    // probe selection and the scan; the device checks the real total against the capacity (query_base_scan_kernel) and a batch that
    // does not fit is neutralised on the device and repeated by the caller with the exact sizes.
    ix->spec_on = false;
    if (allow_spec && stop == STOP_NONE && ix->spec_enabled && ix->rerank_mode == 1 && ix->hw_wpq > 0.0 && ix->shard_count == 1) {
        const double cap = ix->hw_wpq * (double)nb * 1.25 + 4096.0;
        if (cap * 260.0 <= 16.0 * 1073741824.0 && cap < 4.0e9) {
            ix->spec_on = true;
            ix->spec_cap = (uint32_t)cap;
        }
    }
    int rc = run_front(ix, nb, len, P, stop == STOP_ROTATE, false);
    if (rc || stop == STOP_ROTATE) return rc;
    if ((rc = post_totals(ix, nb))) return rc;
    if (stop == STOP_PROBE) return wait_totals(ix, bo);
    const std::vector<Pos> bounds = round_bounds(ix, P, stop == STOP_SCAN_DENSE);
    if (bounds.size() > 17) return fail(RABITQ_EUNSUPPORTED, "more than 16 rerank rounds");
    if ((rc = run_lists_and_quantize(ix, nb, P, bounds))) return rc;
    if (ix->stage_next && (rc = start_staging(ix))) return rc;
    if (ix->spec_on) {
        bo->speculative = true;
        bo->total_words = ix->spec_cap;
        bo->total_pairs = 0;
    } else if ((rc = wait_totals(ix, bo)) || stop == STOP_QUANT) {
        return rc;
    }
    ScanArgs sa;
    RerankArgs ra;
    if ((rc = setup_rounds(ix, nb, P, topk, bo, &sa, &ra))) return rc;
    if (stop != STOP_SCAN_DENSE) {
        // the flat work items of every round on the second side stream (their buffers are sized now; the lists were built long ago)
        cudaStream_t a2 = ix->use_aux2 ? ix->aux2_stream : ix->stream;
        for (size_t r = 0; r + 1 < bounds.size(); r++) {
            ListSet& L = ix->lists[r];
            CU(L.work.ensure(ix->max_items * sizeof(ScanItem)));
            CU(cudaStreamWaitEvent(a2, L.ready, 0));
            work_items_kernel<<<((int)ix->K + 255) / 256, 256, 0, a2>>>(L.item_start.as<uint32_t>(), L.cl_count.as<uint32_t>(),
                                                                                       L.cl_start.as<uint32_t>(), ix->offsets, ix->chunk_start, (int)ix->K,
                                                                                       L.MS, L.ch_min, L.work.as<ScanItem>(),
                                                                                       ix->spec_on ? ix->tot_blk.as<uint32_t>() + 1 : nullptr);
            CU(cudaGetLastError()); ix->counts[5]++;
            if (!L.wready) CU(cudaEventCreateWithFlags(&L.wready, cudaEventDisableTiming));
            CU(cudaEventRecord(L.wready, a2));
            L.w_on_aux = true;
        }
    }
    CU(cudaStreamWaitEvent(ix->stream, ix->ev_win, 0));
    for (size_t r = 0; r + 1 < bounds.size(); r++) {
        if ((rc = run_round_scan(ix, nb, P, sa, bounds[r], bounds[r + 1], stop == STOP_SCAN_DENSE, r))) return rc;
        if (stop == STOP_SCAN_DENSE) return 0;
        ra.win = ix->round_win.as<uint2>() + r * nb;
        if ((rc = run_round_rerank(ix, nb, ra, bounds[r], bounds[r + 1], r == 0, r + 2 == bounds.size(), heuristic, ROUND_REPLAY))) return rc;
    }
    return 0;
}

int validate_query_args(const rabitq_index* ix, size_t len, size_t probe, size_t topk, int heuristic) {
    if (!ix) return fail(RABITQ_EINVAL, "null index");
    if ((len + 63) / 64 * 64 != ix->D)
        return fail(RABITQ_EINVAL, "assertion `left == right` failed: dim != query.len().div_ceil(64) * 64");
    if (probe == 0) return fail(RABITQ_EINVAL, "probe must be >= 1 (the reference underflows on `length - 1`)");
    if (topk == 0) return fail(RABITQ_EINVAL, "topk must be >= 1 (the reference panics on an empty heap peek)");
    // the reference caps neither probe (`probe.min(k)`, src/rabitq.rs:294) nor topk (BinaryHeap::with_capacity, src/rerank.rs:69-78).
    // probe: any.  topk: the k-slot result buffer of a query lives in shared memory next to its row buffers (K5), which bounds
    // topk * 8 + dim * 12 bytes by one SM's 200 KB -- 16384 results per query at dim <= 1024; beyond that is not an ANN call.
    if (topk * 8 + (size_t)ix->D * 12 + 4096 > (size_t)200 * 1024) return fail(RABITQ_EUNSUPPORTED, "topk * 8 + dim * 12 exceeds the on-chip rerank state (200 KB)");
    return 0;
}

size_t pick_sub_batch(const rabitq_index* ix, size_t nq, size_t probe) {
    const size_t K = ix->K, P = std::min(probe, K);
    // survivor slots: 260 B per 32-vector word (2x headroom on the average cluster size); centroid-distance matrix: 4 B per (query, centroid)
    const double words_per_q = double(P) * (double(ix->n) / double(K) / 32.0 + 1.0) * 2.0;
    auto fit = [&](double slot_bytes, double cdist_bytes) {
        size_t nb = 65536;
        nb = std::min(nb, std::max<size_t>(1, (size_t)(cdist_bytes / (double(K) * 4.0))));
        nb = std::min(nb, (size_t)std::max(1.0, (slot_bytes / 260.0) / words_per_q));
        nb = std::min(nb, (size_t)0xffffffffu / std::max<size_t>(P, 1));
        return std::max<size_t>(1, nb);
    };
    // default budgets: 12 GiB of slots, 2 GiB of centroid distances -- enough for every batch the small configurations send
    size_t nb = fit(12.0 * 1073741824.0, 2.0 * 1073741824.0);
    if (nb < nq) {
        // a big batch on a big index (config 5: 65536 queries x 65536 clusters): the more queries share a sub-batch, the more
        // records each scanned cluster chunk serves (the scan's tiles fill up), so take what the part's free memory allows --
        // half of it, counting the work buffers this handle already holds, for the slots and a sixth for the distance matrix
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const double avail = double(free_b) + double(ix->bitmap.cap) + double(ix->entries.cap) + double(ix->cdist.cap);
            nb = std::max(nb, fit(std::min(64.0 * 1073741824.0, avail * 0.5), std::min(8.0 * 1073741824.0, avail / 6.0)));
        } else {
            cudaGetLastError();
        }
    }
    return std::max<size_t>(1, std::min(nb, nq));
}

int query_batch_impl(rabitq_index* ix, const float* queries, bool on_device, size_t nq, size_t len, size_t probe, size_t topk,
                     int heuristic, float* out_dist, uint32_t* out_ids, uint32_t* out_count, const float* next_host = nullptr) {
    int rc = validate_query_args(ix, len, probe, topk, heuristic);
    if (rc) return rc;
    if (nq == 0) return RABITQ_OK;
    if (!queries || !out_dist || !out_ids) return fail(RABITQ_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    std::memset(ix->ms, 0, sizeof(ix->ms));
    ix->timings_pending = false;
    std::memset(ix->counts, 0, sizeof(ix->counts));
    ix->ev_used = 0;
    size_t nbmax = pick_sub_batch(ix, nq, probe);
    nbmax = (nq + (nq + nbmax - 1) / nbmax - 1) / ((nq + nbmax - 1) / nbmax);  // equal sub-batches, not one full and one tiny
    cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (tick(ix, -1)) return RABITQ_ECUDA;
    for (size_t q0 = 0; q0 < nq; q0 += nbmax) {
        const size_t nb = std::min(nbmax, nq - q0);
        if (on_device) {
            ix->q_in = queries + q0 * len;  // read in place: nothing is copied
        } else {
            if (nb == nq && ix->staged_src == queries && ix->staged_nq == nq && ix->staged_len == len) {
                // uploaded by the previous pipelined call while that batch was being answered
                CU(cudaStreamWaitEvent(ix->stream, ix->ev_staged, 0));
                std::swap(ix->qstage, ix->qraw);
            } else {
                CU(ix->qraw.ensure(nb * len * 4));
                CU(cudaMemcpyAsync(ix->qraw.p, queries + q0 * len, nb * len * 4, kin, ix->stream));
            }
            ix->staged_src = nullptr;
            ix->staged_nq = nq; ix->staged_len = len;
            ix->q_in = ix->qraw.as<float>();
            ix->stage_next = (next_host && nb == nq) ? next_host : nullptr;
            ix->stage_bytes = nq * len * 4;
            static const bool late = std::getenv("RABITQ_STAGE_LATE") != nullptr;
            if (ix->stage_next && !late) {
                int rc2 = start_staging(ix);
                if (rc2) return rc2;
            }
        }
        BatchOut bo;
        const bool direct = on_device && nb == nq;
        for (int attempt = 0;; attempt++) {  // (a second pass only when a speculatively sized batch did not fit)
        bo = BatchOut();
        if (direct) { ix->ovr_dist = out_dist; ix->ovr_ids = out_ids; ix->ovr_count = out_count; }
        rc = run_sub_batch(ix, nb, len, probe, topk, heuristic != 0, STOP_NONE, &bo, attempt == 0);
        ix->ovr_dist = nullptr; ix->ovr_ids = nullptr; ix->ovr_count = nullptr;
        if (rc) return rc;
        const uint32_t* oa = ix->out_all.as<uint32_t>();
        const size_t out_words = nb * topk * 2 + nb;
        if (direct) {
        } else if (on_device) {
            CU(cudaMemcpyAsync(out_dist + q0 * topk, oa, nb * topk * 4, kout, ix->stream));
            CU(cudaMemcpyAsync(out_ids + q0 * topk, oa + nb * topk, nb * topk * 4, kout, ix->stream));
            if (out_count) CU(cudaMemcpyAsync(out_count + q0, oa + 2 * nb * topk, nb * 4, kout, ix->stream));
        } else {  // ONE copy into pinned staging (the caller's buffers may be pageable); scattered to them after the sync below
            if (ix->h_out_cap < out_words * 4) {
                if (ix->h_out) cudaFreeHost(ix->h_out);
                ix->h_out = nullptr;
                ix->h_out_cap = 0;
                CU(cudaMallocHost((void**)&ix->h_out, out_words * 4 + out_words / 2));
                ix->h_out_cap = out_words * 4 + out_words / 2;
            }
            CU(cudaMemcpyAsync(ix->h_out, oa, out_words * 4, cudaMemcpyDeviceToHost, ix->stream));
        }
        if (bo.speculative) CU(cudaMemcpyAsync(ix->h_pin, ix->tot_blk.p, 32, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaMemcpyAsync(ix->h_pin + 8, ix->counters.p, 32, cudaMemcpyDeviceToHost, ix->stream));
        if (tick(ix, ST_D2H)) return RABITQ_ECUDA;
        CU(cudaStreamSynchronize(ix->stream));
        if (bo.speculative) {  // the totals (and the prefilter's verdict) arrived with everything else
            prefilter_adapt(ix);
            const uint32_t real_words = ix->h_pin[0];
            std::memcpy(&bo.total_pairs, ix->h_pin + 2, 8);
            ix->hw_wpq = std::max(ix->hw_wpq, (double)real_words / (double)nb);
            ix->spec_on = false;
            if (ix->h_pin[1] != 0u) continue;  // did not fit: neutralised on the device, repeated with the exact sizes
        } else {
            ix->hw_wpq = std::max(ix->hw_wpq, (double)bo.total_words / (double)nb);
        }
        break;
        }
        if (!on_device) {
            std::memcpy(out_dist + q0 * topk, ix->h_out, nb * topk * 4);
            std::memcpy(out_ids + q0 * topk, ix->h_out + nb * topk, nb * topk * 4);
            if (out_count) std::memcpy(out_count + q0, ix->h_out + 2 * nb * topk, nb * 4);
        }
        unsigned long long c[4];
        std::memcpy(c, ix->h_pin + 8, 32);
        ix->counts[0] += bo.total_pairs;
        ix->counts[1] += c[0];
        ix->counts[2] += c[1];
        ix->counts[3] += c[2];
        ix->m_rough += bo.total_pairs;
        ix->m_precise += c[2];
        ix->m_query += nb;
    }
    return collect_timings(ix);
}

// helper for the stage entries: upload host queries and run the front of the pipeline
int stage_prefix(rabitq_index* ix, const float* queries, size_t nq, size_t len, size_t probe, StopAfter stop, BatchOut* bo) {
    int rc = validate_query_args(ix, len, probe, 1, 0);
    if (rc) return rc;
    if (!queries || nq == 0) return fail(RABITQ_EINVAL, "null/empty queries");
    if (nq > 65536 || nq * ix->K > ((size_t)1 << 29)) return fail(RABITQ_EUNSUPPORTED, "stage entries take small batches only");
    CU(cudaSetDevice(ix->device));
    ix->ev_used = 0;
    std::memset(ix->ms, 0, sizeof(ix->ms));
    ix->timings_pending = false;
    std::memset(ix->counts, 0, sizeof(ix->counts));
    CU(ix->qraw.ensure(nq * len * 4));
    CU(cudaMemcpyAsync(ix->qraw.p, queries, nq * len * 4, cudaMemcpyHostToDevice, ix->stream));
    ix->q_in = ix->qraw.as<float>();
    if (tick(ix, -1)) return RABITQ_ECUDA;
    rc = run_sub_batch(ix, nq, len, probe, 1, false, stop, bo);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ix->stream));
    CU(cudaStreamSynchronize(ix->aux_stream));  // later rounds' records are written there
    CU(cudaStreamSynchronize(ix->aux2_stream));
    return 0;
}


// ---- distributed pipeline: phases between the caller's collectives ---------------------------------
size_t dist_chunk_words(const rabitq_index* ix, size_t nq_l, size_t len, int P) { return dist_chunk_layout(nq_l, len, ix->D, (size_t)P).words; }

int dist_init_impl(rabitq_index* ix, int rank, int world, size_t nq_l, size_t probe, size_t topk, size_t records_per_query) {
    if (!ix) return fail(RABITQ_EINVAL, "null index");
    if (world < 1 || world > 32 || rank < 0 || rank >= world) return fail(RABITQ_EINVAL, "bad rank/world (world <= 32)");
    if (world != ix->shard_count || rank != ix->shard_rank) return fail(RABITQ_EINVAL, "rank/world differ from the handle's shard rank/count");
    if (nq_l == 0) return fail(RABITQ_EINVAL, "nq_local must be >= 1");
    int rc = validate_query_args(ix, ix->D, probe, topk, 0);
    if (rc) return rc;
    if (std::min(probe, ix->K) > 4096 || topk > 1024)
        return fail(RABITQ_EUNSUPPORTED, "the sharded pipeline keeps probe <= 4096 and topk <= 1024 (record capacities of the inboxes)");
    if (nq_l * world * std::min(probe, ix->K) > 0xffffffffu) return fail(RABITQ_EUNSUPPORTED, "nq_total * probe exceeds 2^32");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    DistState& d = ix->dist;
    // Peers' inboxes must be unmapped by every rank BEFORE their owners free them (CUDA IPC: freeing exported memory that a peer
    // still has open is undefined).  The caller closes them first on every rank (rabitq_dist_close_peers), runs a barrier, and only
    // then re-initialises; whatever is still open here is closed as a last resort.
    for (size_t r = 0; r < d.peers_h.size(); r++)
        if (d.opened[r] && d.peers_h[r]) cudaIpcCloseMemHandle(d.peers_h[r]);
    CU(cudaStreamSynchronize(ix->stream));
    if (ix->copy_stream) CU(cudaStreamSynchronize(ix->copy_stream));
    if (d.inbox) cudaFree(d.inbox);
    if (d.peers_d) cudaFree(d.peers_d);
    if (d.ev_rot) cudaEventDestroy(d.ev_rot);
    for (int c = 0; c < DistState::NCOPY; c++) {
        if (d.copy_streams[c]) { cudaStreamSynchronize(d.copy_streams[c]); cudaStreamDestroy(d.copy_streams[c]); }
        if (d.ev_pushes[c]) cudaEventDestroy(d.ev_pushes[c]);
    }
    d = DistState();
    CU(cudaEventCreateWithFlags(&d.ev_rot, cudaEventDisableTiming));
    for (int c = 0; c < DistState::NCOPY; c++) {
        CU(cudaStreamCreateWithFlags(&d.copy_streams[c], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&d.ev_pushes[c], cudaEventDisableTiming));
    }
    d.world = world; d.rank = rank; d.nq_l = nq_l; d.topk = topk;
    d.P = (int)std::min(probe, ix->K);
    d.r1cap = (uint32_t)(SCAN_THREADS * std::max(1, ix->first_chunks));
    const size_t c2 = nq_l * std::max<size_t>(records_per_query, 8);
    if (c2 > 0x7fffffffu) return fail(RABITQ_EUNSUPPORTED, "records_per_query too large");
    d.cap2 = (uint32_t)c2;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    d.off_r1cnt = 0;
    d.off_r1rec = up(d.off_r1cnt + (size_t)world * nq_l * 4);
    d.off_r2tab = up(d.off_r1rec + (size_t)world * nq_l * d.r1cap * sizeof(SurvRec));
    d.off_r2rec = up(d.off_r2tab + (size_t)world * nq_l * 8);
    d.off_q = up(d.off_r2rec + (size_t)world * d.cap2 * sizeof(SurvRec));
    d.off_y = up(d.off_q + (size_t)world * nq_l * ix->D * 4);
    d.inbox_bytes = up(d.off_y + (size_t)world * nq_l * ix->D * 4);
    CU(cudaMalloc((void**)&d.inbox, d.inbox_bytes));
    CU(cudaMemset(d.inbox, 0, d.off_r1rec));
    CU(cudaMalloc((void**)&d.peers_d, sizeof(void*) * world));
    d.peers_h.assign(world, nullptr);
    d.opened.assign(world, 0);
    d.peers_h[rank] = d.inbox;
    CU(cudaMemcpy(d.peers_d, d.peers_h.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
    d.ready = true;
    return RABITQ_OK;
}

int dist_set_peer_impl(rabitq_index* ix, int r, const unsigned char* ipc_handle, void* raw) {
    if (!ix || !ix->dist.ready) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    DistState& d = ix->dist;
    if (r < 0 || r >= d.world) return fail(RABITQ_EINVAL, "bad peer rank");
    if (r == d.rank) return RABITQ_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    if (d.opened[r] && d.peers_h[r]) { cudaIpcCloseMemHandle(d.peers_h[r]); d.opened[r] = 0; }
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
        std::memcpy(&h, ipc_handle, sizeof(h));
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        d.peers_h[r] = static_cast<unsigned char*>(ptr);
        d.opened[r] = 1;
    } else {
        if (!raw) return fail(RABITQ_EINVAL, "null peer pointer");
        d.peers_h[r] = static_cast<unsigned char*>(raw);
    }
    CU(cudaMemcpy(d.peers_d, d.peers_h.data(), sizeof(void*) * d.world, cudaMemcpyHostToDevice));
    return RABITQ_OK;
}

// Phase 1 (home): front end of this rank's nq_l queries -> its chunk(s) of the all-gather.  Two halves, so that the caller can
// start the all-gather of the big [q | y] part while the centroid scan and the probe selection still run:
//   1a  pad + rotate          -> d_send_qy   (words_a of dist_chunk_layout)
//   1b  K2p / K2 / K2b        -> d_send_meta (words_b)
int dist_front_rotate_impl(rabitq_index* ix, const float* d_queries, size_t len, void* d_send_qy) {
    if (!ix || !ix->dist.ready) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    if (!d_queries) return fail(RABITQ_EINVAL, "null argument");
    if (!d_send_qy)  // push mode: the chunk goes straight into every peer's inbox
        for (int r = 0; r < ix->dist.world; r++)
            if (!ix->dist.peers_h[r]) return fail(RABITQ_EINVAL, "peer inbox of rank " + std::to_string(r) + " not set (rabitq_dist_set_peer)");
    DistState& d = ix->dist;
    if ((len + 63) / 64 * 64 != ix->D) return fail(RABITQ_EINVAL, "assertion `left == right` failed: dim != query.len().div_ceil(64) * 64");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    std::memset(ix->ms, 0, sizeof(ix->ms));
    ix->timings_pending = false;
    std::memset(ix->counts, 0, sizeof(ix->counts));
    ix->ev_used = 0;
    d.len = len;
    const size_t nq_l = d.nq_l, D = ix->D;
    cudaStream_t st = ix->stream;
    if (tick(ix, -1)) return RABITQ_ECUDA;
    CU(cudaMemsetAsync(d.inbox + d.off_r1cnt, 0, (size_t)d.world * nq_l * 4, st));  // before the all-gather = before any owner writes
    CU(ix->qraw.ensure(nq_l * len * 4));
    CU(cudaMemcpyAsync(ix->qraw.p, d_queries, nq_l * len * 4, cudaMemcpyDeviceToDevice, st));
    ix->q_in = ix->qraw.as<float>();
    int rc = run_front(ix, nq_l, len, d.P, true, true);
    if (rc) return rc;
    const DistChunk L = dist_chunk_layout(nq_l, len, D, (size_t)d.P);
    if (d_send_qy) {  // the caller all-gathers the chunk
        uint32_t* send = static_cast<uint32_t*>(d_send_qy);
        CU(cudaMemcpyAsync(send + L.a_q, ix->qraw.p, nq_l * len * 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(send + L.a_y, ix->y.p, nq_l * D * 4, cudaMemcpyDeviceToDevice, st));
        d.push_pending = false;
    } else {
        // PUSH: this rank's block of padded queries and of rotated queries into rows [rank * nq_l, (rank + 1) * nq_l) of the flat
        // Q and Y regions of every inbox (its own included) with copy-engine transfers over NVLink, spread over a few copy streams:
        // no collective and no SM is involved, so the centroid scan and the probe selection that follow on the main stream run
        // undisturbed, and nothing has to be unpacked afterwards.  Peers may read the rows once they have seen this rank's part of
        // the NEXT collective, which the main stream enters only after the pushes (front_select waits for the push events).
        CU(cudaEventRecord(d.ev_rot, st));
        const size_t blk = nq_l * D * 4;
        for (int c = 0; c < DistState::NCOPY; c++) CU(cudaStreamWaitEvent(d.copy_streams[c], d.ev_rot, 0));
        for (int k = 0; k < d.world; k++) {
            const int r = (d.rank + 1 + k) % d.world;  // stagger the targets: rank i starts with i+1, itself last
            cudaStream_t cs = d.copy_streams[k % DistState::NCOPY];
            CU(cudaMemcpyAsync(d.peers_h[r] + d.off_q + (size_t)d.rank * blk, ix->q_pad, blk, cudaMemcpyDeviceToDevice, cs));
            CU(cudaMemcpyAsync(d.peers_h[r] + d.off_y + (size_t)d.rank * blk, ix->y.p, blk, cudaMemcpyDeviceToDevice, cs));
        }
        for (int c = 0; c < DistState::NCOPY; c++) CU(cudaEventRecord(d.ev_pushes[c], d.copy_streams[c]));
        d.push_pending = true;
    }
    if (tick(ix, ST_ROTATE)) return RABITQ_ECUDA;
    d.phase = 10;
    return RABITQ_OK;
}

int dist_front_select_impl(rabitq_index* ix, void* d_send_meta) {
    if (!ix || !ix->dist.ready || ix->dist.phase != 10) return fail(RABITQ_EINVAL, "rabitq_dist_front_rotate first");
    if (!d_send_meta) return fail(RABITQ_EINVAL, "null argument");
    DistState& d = ix->dist;
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    const size_t nq_l = d.nq_l, D = ix->D;
    const int P = d.P;
    cudaStream_t st = ix->stream;
    int rc = run_front_select(ix, nq_l, P, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ix->h_pin + 18, ix->q_pbase.as<unsigned long long>() + nq_l, 8, cudaMemcpyDeviceToHost, st));  // `rough` of the home queries
    uint32_t* send = static_cast<uint32_t*>(d_send_meta);
    const DistChunk L = dist_chunk_layout(nq_l, d.len, D, (size_t)P);
    CU(cudaMemcpyAsync(send + L.b_ids, ix->probe_ids.p, nq_l * P * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(send + L.b_dist, ix->probe_dist.p, nq_l * P * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(send + L.b_p0, ix->q_p0.p, nq_l * 4, cudaMemcpyDeviceToDevice, st));
    if (d.push_pending)  // the collective that follows carries the "my pushes are done" edge
        for (int c = 0; c < DistState::NCOPY; c++) CU(cudaStreamWaitEvent(st, d.ev_pushes[c], 0));
    if (tick(ix, ST_SELECT)) return RABITQ_ECUDA;
    d.phase = 1;
    return RABITQ_OK;
}

// the two halves into ONE chunk (combined layout: part B follows part A)
int dist_front_impl(rabitq_index* ix, const float* d_queries, size_t len, void* d_send) {
    int rc = dist_front_rotate_impl(ix, d_queries, len, d_send);
    if (rc) return rc;
    const DistChunk L = dist_chunk_layout(ix->dist.nq_l, len, ix->D, (size_t)ix->dist.P);
    return dist_front_select_impl(ix, static_cast<uint32_t*>(d_send) + L.words_a);
}

// Phase 2 (source): the whole batch's products -> local slots, query records, round 1 (first chunk of the nearest non-empty
// cluster) on the shard that owns it; the round-1 threshold of every owned query lands in d_thr (others keep +FLT_MAX).
int dist_round1_impl(rabitq_index* ix, const void* d_gathered_a, size_t stride_a, const void* d_gathered_b, size_t stride_b, float* d_thr) {
    if (!ix || !ix->dist.ready || ix->dist.phase != 1) return fail(RABITQ_EINVAL, "rabitq_dist_front first");
    if (!d_gathered_b || !d_thr) return fail(RABITQ_EINVAL, "null argument");
    DistState& d = ix->dist;
    const bool pushed = !d_gathered_a;  // push mode: every rank's rows already sit in the flat Q / Y regions of this rank's inbox
    if (pushed) {
        if (!d.push_pending) return fail(RABITQ_EINVAL, "no gathered [q | y] buffer and nothing was pushed");
        d.push_pending = false;
    }
    for (int r = 0; r < d.world; r++)
        if (!d.peers_h[r]) return fail(RABITQ_EINVAL, "peer inbox of rank " + std::to_string(r) + " not set (rabitq_dist_set_peer)");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    ix->spec_on = false;  // (the distributed phases always read the slot totals back: a repeat would be a collective act)
    const size_t nq = d.nq_l * d.world, D = ix->D;
    const int P = d.P;
    if (!pushed) {
        CU(ix->qpad.ensure(nq * D * 4));
        CU(ix->y.ensure(nq * D * 4));
    }
    CU(ix->probe_ids.ensure(nq * P * 4));
    CU(ix->probe_dist.ensure(nq * P * 4));
    CU(ix->slot_local.ensure(nq * P * 4));
    CU(ix->q_words.ensure(nq * 4));
    CU(ix->q_pairs.ensure(nq * 4));
    CU(ix->q_p0.ensure(nq * 4));
    CU(ix->q_wbase.ensure((nq + 1) * 4));
    CU(ix->q_pbase.ensure((nq + 1) * 8));
    CU(ix->r2_cnt.ensure(nq * 4));
    CU(ix->r2_off.ensure(nq * 4));
    CU(ix->home_tot.ensure(32 * 4));
    CU(ix->cand.ensure((size_t)d.world * d.cap2 * sizeof(Cand)));
    // all ranks' queries: unpacked from the gathered chunks below, or already in place (pushed)
    ix->q_pad = pushed ? reinterpret_cast<const float*>(d.inbox + d.off_q) : ix->qpad.as<float>();
    ix->y_all = pushed ? reinterpret_cast<const float*>(d.inbox + d.off_y) : ix->y.as<float>();
    if (tick(ix, -1)) return RABITQ_ECUDA;  // the all-gather sits between the phases: not ours to time
    dist_unpack_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(static_cast<const uint32_t*>(d_gathered_a), stride_a,
                                                          static_cast<const uint32_t*>(d_gathered_b), stride_b, d.world, (int)d.nq_l, (int)d.len, (int)D, P,
                                                          pushed ? nullptr : ix->qpad.as<float>(), pushed ? nullptr : ix->y.as<float>(), ix->probe_ids.as<uint32_t>(),
                                                          ix->probe_dist.as<float>(), ix->q_p0.as<uint32_t>());
    CU(cudaGetLastError()); ix->counts[5]++;
    if (tick(ix, ST_H2D)) return RABITQ_ECUDA;
    slot_layout_kernel<<<(unsigned)nq, SEL_THREADS, (size_t)P * 4, st>>>(ix->probe_ids.as<uint32_t>(), ix->offsets, P, ix->slot_local.as<uint32_t>(),
                                                                          ix->q_words.as<uint32_t>(), ix->q_pairs.as<uint32_t>());
    CU(cudaGetLastError()); ix->counts[5]++;
    query_base_scan_kernel<<<1, 1024, 0, st>>>(ix->q_words.as<uint32_t>(), ix->q_pairs.as<uint32_t>(), (int)nq, ix->q_wbase.as<uint32_t>(),
                                               ix->q_pbase.as<unsigned long long>(), 0u, nullptr, ix->tot_blk.as<uint32_t>());
    CU(cudaGetLastError()); ix->counts[5]++;
    BatchOut bo;
    bo.P = P;
    int rc = post_totals(ix, nq);
    if (rc) return rc;
    const int fc = (int)(d.r1cap / SCAN_THREADS);
    // both rounds' inverted lists on the side stream (they depend on the unpacked probe lists only), K3 over each behind it
    if ((rc = run_lists_and_quantize(ix, nq, P, {Pos{0, 0}, Pos{0, fc}, Pos{P, 0}}))) return rc;
    if ((rc = wait_totals(ix, &bo))) return rc;
    ix->counts[0] += bo.total_pairs;
    if ((rc = setup_rounds(ix, nq, P, d.topk, &bo, &d.sa, &d.ra))) return rc;
    CU(cudaStreamWaitEvent(st, ix->ev_win, 0));  // the batch's counters were reset on the second side stream (run_lists_and_quantize)
    fill_f32_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_thr, nq, 3.402823466e+38f);
    CU(cudaGetLastError()); ix->counts[5]++;
    d.d_thr = d_thr;
    d.sa.thr = d_thr;
    d.ra.thr = d_thr;
    d.ra.peers = d.peers_d;
    d.ra.off_r1cnt = d.off_r1cnt; d.ra.off_r1rec = d.off_r1rec;
    d.ra.world = d.world; d.ra.rank = d.rank; d.ra.nq_local = (int)d.nq_l; d.ra.r1cap = (int)d.r1cap;
    if ((rc = run_round_scan(ix, nq, P, d.sa, Pos{0, 0}, Pos{0, fc}, false, 0))) return rc;
    if ((rc = run_round_rerank(ix, nq, d.ra, Pos{0, 0}, Pos{0, fc}, true, false, false, ROUND_SINK1))) return rc;
    d.phase = 2;
    return RABITQ_OK;
}

// Phase 3 (source), after the all-reduce(min) of d_thr: everything after round 1, filtered with the frozen threshold; every
// survivor's exact distance is computed here, next to its base row, and shipped to the home rank's inbox.
int dist_round2_impl(rabitq_index* ix, uint32_t* d_status) {
    if (!ix || !ix->dist.ready || ix->dist.phase != 2) return fail(RABITQ_EINVAL, "rabitq_dist_round1 first");
    if (!d_status) return fail(RABITQ_EINVAL, "null argument");
    DistState& d = ix->dist;
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const size_t nq = d.nq_l * d.world;
    const int P = d.P, fc = (int)(d.r1cap / SCAN_THREADS);
    if (tick(ix, -1)) return RABITQ_ECUDA;
    CU(cudaMemsetAsync(d_status, 0, 4, st));
    int rc = run_round_scan(ix, nq, P, d.sa, Pos{0, fc}, Pos{P, 0}, false, 1);
    if (rc) return rc;
    r2_count_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(ix->bitmap.as<uint32_t>(), ix->q_wbase.as<uint32_t>(), ix->slot_local.as<uint32_t>(),
                                                              ix->q_p0.as<uint32_t>(), (int)nq, P, 0, fc, ix->r2_cnt.as<uint32_t>());
    CU(cudaGetLastError()); ix->counts[5]++;
    r2_offsets_kernel<<<d.world, 1024, 0, st>>>(ix->r2_cnt.as<uint32_t>(), ix->r2_off.as<uint32_t>(), (int)d.nq_l, d.cap2, d.peers_d, d.off_r2tab,
                                                 d.rank, d_status, ix->home_tot.as<uint32_t>());
    CU(cudaGetLastError()); ix->counts[5]++;
    const bool r2_seq = ix->dist_r2_seq && P <= 4096 &&
                        rerank_cta_smem((int)ix->D, (int)d.topk, 1, 2, 1) + (size_t)P * 4 <= (size_t)200 * 1024;
    if (r2_seq) {
        // source-side sequential filter: every query's candidates on this shard replayed against a local threshold that is never
        // below the reference's (rerank_cta_kernel<.., 2>); only what it computes is shipped, exact distances included
        if (tick(ix, ST_BUCKET)) return RABITQ_ECUDA;
        RerankArgs ra2 = d.ra;
        ra2.win = nullptr;
        ra2.r2_cnt = ix->r2_cnt.as<uint32_t>();
        ra2.r2_off = ix->r2_off.as<uint32_t>();
        ra2.off_r2rec = d.off_r2rec;
        ra2.off_r2tab = d.off_r2tab;
        ra2.cap2 = d.cap2;
        if ((rc = run_round_rerank(ix, nq, ra2, Pos{0, fc}, Pos{P, 0}, true, false, false, ROUND_SINK2))) return rc;
        d.phase = 3;
        return RABITQ_OK;
    }
    r2_compact_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(ix->bitmap.as<uint32_t>(), ix->entries.as<float2>(), ix->q_wbase.as<uint32_t>(),
                                                                ix->slot_local.as<uint32_t>(), ix->q_p0.as<uint32_t>(), ix->r2_cnt.as<uint32_t>(),
                                                                ix->r2_off.as<uint32_t>(), ix->home_tot.as<uint32_t>(), (int)nq, (int)d.nq_l, P, 0, fc,
                                                                ix->cand.as<Cand>());
    CU(cudaGetLastError()); ix->counts[5]++;
    if (tick(ix, ST_BUCKET)) return RABITQ_ECUDA;
    r2_exact_kernel<<<ix->sm_count * 8, 256, 0, st>>>(ix->cand.as<Cand>(), ix->home_tot.as<uint32_t>(), d.world, ix->q_pad, ix->base,
                                                      ix->map_ids, (int)ix->D, (int)d.nq_l, d.rank, d.cap2, d.peers_d, d.off_r2rec,
                                                      ix->counters.as<unsigned long long>());
    CU(cudaGetLastError()); ix->counts[5]++;
    if (tick(ix, ST_RERANK)) return RABITQ_ECUDA;
    d.phase = 3;
    return RABITQ_OK;
}

// Phase 4 (home), after a barrier that makes every shard's records visible: the sequential replay of the home queries.
int dist_finish_impl(rabitq_index* ix, float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count, uint32_t* d_status) {
    if (!ix || !ix->dist.ready || ix->dist.phase != 3) return fail(RABITQ_EINVAL, "rabitq_dist_round2 first");
    if (!d_out_dist || !d_out_ids || !d_out_count || !d_status) return fail(RABITQ_EINVAL, "null argument");
    DistState& d = ix->dist;
    std::lock_guard<std::mutex> lk(ix->mu);
    CU(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    if (tick(ix, -1)) return RABITQ_ECUDA;
    HomeArgs h;
    h.inbox = d.inbox;
    h.off_r1cnt = d.off_r1cnt; h.off_r1rec = d.off_r1rec; h.off_r2tab = d.off_r2tab; h.off_r2rec = d.off_r2rec;
    h.probe_ids = ix->probe_ids.as<uint32_t>();
    h.q_p0 = ix->q_p0.as<uint32_t>();
    h.goffsets = ix->goffsets;
    h.row_bounds = ix->row_bounds;
    h.out_dist = d_out_dist; h.out_ids = d_out_ids; h.out_count = d_out_count;
    h.counters = ix->counters.as<unsigned long long>();
    h.status = d_status;
    h.world = d.world; h.rank = d.rank; h.nq_l = (int)d.nq_l; h.P = d.P; h.topk = (int)d.topk; h.r1cap = (int)d.r1cap; h.cap2 = d.cap2;
    CU(cudaMemsetAsync(ix->counters.as<unsigned long long>() + 2, 0, 8, st));  // [2] = precise of the HOME queries from here on
    home_replay_kernel<<<(unsigned)((d.nq_l + 3) / 4), 128, 4 * 2 * d.topk * 4, st>>>(h);
    CU(cudaGetLastError()); ix->counts[5]++;
    if (tick(ix, ST_RERANK)) return RABITQ_ECUDA;
    CU(cudaMemcpyAsync(ix->h_pin + 8, ix->counters.p, 32, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ix->h_pin + 16, d_status, 4, cudaMemcpyDeviceToHost, st));  // rabitq_dist_last_status: no second sync for the caller
    CU(cudaStreamSynchronize(st));
    prefilter_adapt(ix);
    unsigned long long c[4], rough_home;
    std::memcpy(c, ix->h_pin + 8, 32);
    std::memcpy(&rough_home, ix->h_pin + 18, 8);
    ix->counts[1] += c[0];
    ix->counts[2] += c[1];
    ix->counts[3] += c[2];
    ix->m_query += d.nq_l;
    ix->m_rough += rough_home;
    ix->m_precise += c[2];
    d.phase = 0;
    ix->timings_pending = true;  // every phase restarted the event chain with a marker; resolved on demand
    return RABITQ_OK;
}

}  // namespace

// ===================================================================================================================
extern "C" {

int rabitq_dist_init(rabitq_index* idx, int rank, int world, size_t nq_local, size_t probe, size_t topk, size_t records_per_query,
                     size_t* inbox_bytes) {
    int rc = dist_init_impl(idx, rank, world, nq_local, probe, topk, records_per_query);
    if (rc == 0 && inbox_bytes) *inbox_bytes = idx->dist.inbox_bytes;
    return rc;
}

int rabitq_dist_ipc_handle(rabitq_index* idx, unsigned char out_handle[64]) {
    if (!idx || !idx->dist.ready || !out_handle) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    CU(cudaSetDevice(idx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, idx->dist.inbox));
    std::memcpy(out_handle, &h, 64);
    return RABITQ_OK;
}

int rabitq_dist_inbox_ptr(rabitq_index* idx, void** out) {
    if (!idx || !idx->dist.ready || !out) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    *out = idx->dist.inbox;
    return RABITQ_OK;
}

int rabitq_dist_close_peers(rabitq_index* idx) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    CU(cudaSetDevice(idx->device));
    DistState& d = idx->dist;
    CU(cudaStreamSynchronize(idx->stream));
    for (int c = 0; c < DistState::NCOPY; c++)
        if (d.copy_streams[c]) CU(cudaStreamSynchronize(d.copy_streams[c]));
    for (size_t r = 0; r < d.peers_h.size(); r++) {
        if (d.opened[r] && d.peers_h[r]) cudaIpcCloseMemHandle(d.peers_h[r]);
        if ((int)r != d.rank) d.peers_h[r] = nullptr;
        d.opened[r] = 0;
    }
    return RABITQ_OK;
}

int rabitq_dist_set_peer(rabitq_index* idx, int peer_rank, const unsigned char* ipc_handle, void* raw_ptr) {
    return dist_set_peer_impl(idx, peer_rank, ipc_handle, raw_ptr);
}

size_t rabitq_dist_chunk_words(const rabitq_index* idx, size_t len) {
    if (!idx || !idx->dist.ready) return 0;
    return dist_chunk_words(idx, idx->dist.nq_l, len, idx->dist.P);
}

int rabitq_dist_front(rabitq_index* idx, const float* d_queries, size_t len, void* d_send) { return dist_front_impl(idx, d_queries, len, d_send); }
int rabitq_dist_round1(rabitq_index* idx, const void* d_gathered, float* d_thr) {
    if (!idx || !idx->dist.ready) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    const DistChunk L = dist_chunk_layout(idx->dist.nq_l, idx->dist.len, idx->D, (size_t)idx->dist.P);
    return dist_round1_impl(idx, d_gathered, L.words, d_gathered ? static_cast<const uint32_t*>(d_gathered) + L.words_a : nullptr, L.words, d_thr);
}
size_t rabitq_dist_chunk_words_qy(const rabitq_index* idx, size_t len) {
    if (!idx || !idx->dist.ready) return 0;
    return dist_chunk_layout(idx->dist.nq_l, len, idx->D, (size_t)idx->dist.P).words_a;
}
size_t rabitq_dist_chunk_words_meta(const rabitq_index* idx, size_t len) {
    if (!idx || !idx->dist.ready) return 0;
    return dist_chunk_layout(idx->dist.nq_l, len, idx->D, (size_t)idx->dist.P).words_b;
}
int rabitq_dist_front_rotate(rabitq_index* idx, const float* d_queries, size_t len, void* d_send_qy) {
    return dist_front_rotate_impl(idx, d_queries, len, d_send_qy);
}
int rabitq_dist_front_select(rabitq_index* idx, void* d_send_meta) { return dist_front_select_impl(idx, d_send_meta); }
int rabitq_dist_round1_split(rabitq_index* idx, const void* d_gathered_qy, const void* d_gathered_meta, float* d_thr) {
    if (!idx || !idx->dist.ready) return fail(RABITQ_EINVAL, "rabitq_dist_init first");
    const DistChunk L = dist_chunk_layout(idx->dist.nq_l, idx->dist.len, idx->D, (size_t)idx->dist.P);
    return dist_round1_impl(idx, d_gathered_qy /* NULL: pushed */, L.words_a, d_gathered_meta, L.words_b, d_thr);
}
int rabitq_dist_round2(rabitq_index* idx, uint32_t* d_status) { return dist_round2_impl(idx, d_status); }
int rabitq_dist_finish(rabitq_index* idx, float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count, uint32_t* d_status) {
    return dist_finish_impl(idx, d_out_dist, d_out_ids, d_out_count, d_status);
}

int rabitq_dist_last_status(const rabitq_index* idx, uint32_t* out_status) {
    if (!idx || !out_status) return fail(RABITQ_EINVAL, "null argument");
    *out_status = idx->h_pin[16];
    return RABITQ_OK;
}

int rabitq_min_f32_device(int device, float* d_dst, const float* d_src, size_t n, void* cuda_stream) {
    if (!d_dst || !d_src) return fail(RABITQ_EINVAL, "null argument");
    if (n == 0) return RABITQ_OK;
    CU(cudaSetDevice(device));
    rq::min_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_dst, d_src, n);
    CU(cudaGetLastError());
    return RABITQ_OK;
}

const char* rabitq_last_error(void) { return g_err.c_str(); }

int rabitq_debug_rec_pos(int d) { return rq::rec_pos(d); }

int rabitq_load_from_dir(const char* dir, int device, rabitq_index** out) { return load_dir(dir, device, 0, 1, out); }

int rabitq_load_from_dir_sharded(const char* dir, int device, int shard_rank, int shard_count, rabitq_index** out) {
    return load_dir(dir, device, shard_rank, shard_count, out);
}

int rabitq_from_arrays(uint32_t dim, size_t n, size_t k, const float* base, const float* orthogonal, const float* centroids,
                       const uint32_t* offsets, const uint32_t* map_ids, const uint64_t* codes, const float* factors,
                       int ptr_on_device, int device, int shard_rank, int shard_count, rabitq_index** out) {
    if (!base || !orthogonal || !centroids || !offsets || !map_ids || !codes || !factors || !out)
        return fail(RABITQ_EINVAL, "null argument");
    std::vector<uint32_t> off_h(k + 1);
    if (ptr_on_device) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(RABITQ_ECUDA, "no CUDA device: rabitq_b200 has no CPU fallback");
        }
        CU(cudaSetDevice(device));
        CU(cudaMemcpy(off_h.data(), offsets, (k + 1) * 4, cudaMemcpyDeviceToHost));
    } else {
        std::memcpy(off_h.data(), offsets, (k + 1) * 4);
    }
    return make_index(dim, n, k, base, orthogonal, centroids, off_h.data(), map_ids, codes, factors, ptr_on_device != 0, device,
                      shard_rank, shard_count, out);
}

int rabitq_build(const float* base, size_t n, size_t len, const float* centroids, size_t k, const float* orthogonal, uint64_t seed,
                 int ptr_on_device, int device, rabitq_index** out) {
    return build_impl(base, n, len, centroids, k, orthogonal, seed, ptr_on_device != 0, device, out);
}

int rabitq_from_path(const char* base_path, const char* centroid_path, uint64_t seed, int device, rabitq_index** out) {
    if (!base_path || !centroid_path || !out) return fail(RABITQ_EINVAL, "null argument");
    std::vector<float> base, cent;
    std::vector<size_t> rl_b, rl_c;
    if (!read_vecs(base_path, base, rl_b) || rl_b.empty()) return fail(RABITQ_EIO, "read vecs error: base");
    if (!read_vecs(centroid_path, cent, rl_c) || rl_c.empty()) return fail(RABITQ_EIO, "read vecs error: centroids");
    const size_t len = rl_b[0];
    if (rl_c[0] != len) return fail(RABITQ_EINVAL, "assertion failed: dim == centroids.ncols()");
    if (base.size() != rl_b.size() * len || cent.size() != rl_c.size() * len) return fail(RABITQ_EINVAL, "ragged fvecs records");
    return build_impl(base.data(), rl_b.size(), len, cent.data(), rl_c.size(), nullptr, seed, false, device, out);
}

int rabitq_dump_to_dir(rabitq_index* idx, const char* dir) { return dump_impl(idx, dir); }

int rabitq_reshard(rabitq_index* idx, int shard_rank, int shard_count, rabitq_index** out) {
    if (!idx || !out) return fail(RABITQ_EINVAL, "null argument");
    if (idx->shard_count != 1) return fail(RABITQ_EUNSUPPORTED, "reshard needs an unsharded handle");
    std::lock_guard<std::mutex> lk(idx->mu);
    CU(cudaSetDevice(idx->device));
    CU(cudaStreamSynchronize(idx->stream));
    std::vector<uint32_t> off_h(idx->K + 1);
    CU(cudaMemcpy(off_h.data(), idx->offsets, (idx->K + 1) * 4, cudaMemcpyDeviceToHost));
    return make_index(idx->D, idx->n, idx->K, idx->base, idx->P, idx->cent, off_h.data(), idx->map_ids,
                      reinterpret_cast<const uint64_t*>(idx->codes), reinterpret_cast<const float*>(idx->factors), true, idx->device, shard_rank,
                      shard_count, out);
}

int rabitq_export_arrays(rabitq_index* idx, float* base, float* orthogonal, float* centroids, uint32_t* offsets, uint32_t* map_ids,
                         uint64_t* codes, float* factors, int ptr_on_device) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    if (idx->shard_count != 1) return fail(RABITQ_EUNSUPPORTED, "export needs an unsharded handle");
    std::lock_guard<std::mutex> lk(idx->mu);
    CU(cudaSetDevice(idx->device));
    const cudaMemcpyKind kind = ptr_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const size_t D = idx->D, K = idx->K, n = idx->n;
    if (base) CU(cudaMemcpy(base, idx->base, n * D * 4, kind));
    if (orthogonal) CU(cudaMemcpy(orthogonal, idx->P, D * D * 4, kind));
    if (centroids) CU(cudaMemcpy(centroids, idx->cent, K * D * 4, kind));
    if (offsets) CU(cudaMemcpy(offsets, idx->offsets, (K + 1) * 4, kind));
    if (map_ids) CU(cudaMemcpy(map_ids, idx->map_ids, n * 4, kind));
    if (codes) CU(cudaMemcpy(codes, idx->codes, n * (D / 64) * 8, kind));
    if (factors) CU(cudaMemcpy(factors, idx->factors, n * 16, kind));
    return RABITQ_OK;
}

void rabitq_free(rabitq_index* idx) { delete idx; }

uint32_t rabitq_dim(const rabitq_index* idx) { return idx ? idx->D : 0; }
size_t rabitq_num_vectors(const rabitq_index* idx) { return idx ? idx->n : 0; }
size_t rabitq_num_clusters(const rabitq_index* idx) { return idx ? idx->K : 0; }

int rabitq_query(rabitq_index* idx, const float* query, size_t len, size_t probe, size_t topk, int heuristic_rank,
                 float* out_dist, uint32_t* out_ids, uint32_t* out_count) {
    return query_batch_impl(idx, query, false, 1, len, probe, topk, heuristic_rank, out_dist, out_ids, out_count);
}

int rabitq_query_batch(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe, size_t topk,
                       int heuristic_rank, float* out_dist, uint32_t* out_ids, uint32_t* out_count) {
    return query_batch_impl(idx, queries, false, nq, len, probe, topk, heuristic_rank, out_dist, out_ids, out_count);
}

int rabitq_query_batch_pipelined(rabitq_index* idx, const float* queries, const float* next_queries, size_t nq, size_t len, size_t probe,
                                 size_t topk, int heuristic_rank, float* out_dist, uint32_t* out_ids, uint32_t* out_count) {
    return query_batch_impl(idx, queries, false, nq, len, probe, topk, heuristic_rank, out_dist, out_ids, out_count, next_queries);
}

int rabitq_query_batch_device(rabitq_index* idx, const float* d_queries, size_t nq, size_t len, size_t probe, size_t topk,
                              int heuristic_rank, float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count) {
    return query_batch_impl(idx, d_queries, true, nq, len, probe, topk, heuristic_rank, d_out_dist, d_out_ids, d_out_count);
}

int rabitq_shard_range(const uint32_t* offsets, size_t k, int shard_rank, int shard_count, size_t* row_lo, size_t* row_hi) {
    if (!offsets || !row_lo || !row_hi || k == 0 || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
        return fail(RABITQ_EINVAL, "bad argument");
    shard_rows(offsets, k, shard_rank, shard_count, row_lo, row_hi);
    return RABITQ_OK;
}

int rabitq_merge_topk_device(int device, const float* d_dist, const uint32_t* d_ids, int n_lists, size_t nq, size_t topk,
                             float* d_out_dist, uint32_t* d_out_ids, uint32_t* d_out_count, void* cuda_stream) {
    if (!d_dist || !d_ids || !d_out_dist || !d_out_ids || !d_out_count || n_lists < 1 || topk == 0)
        return fail(RABITQ_EINVAL, "bad argument");
    if (nq == 0) return RABITQ_OK;
    CU(cudaSetDevice(device));
    rq::merge_topk_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_dist, d_ids, n_lists, nq, (int)topk,
                                                                                                       d_out_dist, d_out_ids, d_out_count);
    CU(cudaGetLastError());
    return RABITQ_OK;  // stream-ordered: the caller synchronises (or keeps launching on the same stream)
}

void rabitq_metrics(const rabitq_index* idx, uint64_t out[4]) {
    out[0] = idx ? idx->m_query.load(std::memory_order_relaxed) : 0;
    out[1] = idx ? idx->m_rough.load(std::memory_order_relaxed) : 0;
    out[2] = idx ? idx->m_precise.load(std::memory_order_relaxed) : 0;
    out[3] = 0;  // cache miss: the disk/S3 cache of crates/disk is out of scope; base vectors are HBM-resident
}

void rabitq_metrics_reset(rabitq_index* idx) {
    if (idx) { idx->m_query = 0; idx->m_rough = 0; idx->m_precise = 0; }
}

int rabitq_set_rounds(rabitq_index* idx, const uint32_t* rounds, int n) {
    if (!idx || !rounds || n < 1 || rounds[0] != 0) return fail(RABITQ_EINVAL, "rounds must start at 0");
    for (int i = 1; i < n; i++)
        if (rounds[i] <= rounds[i - 1]) return fail(RABITQ_EINVAL, "rounds must be strictly increasing");
    std::lock_guard<std::mutex> lk(idx->mu);
    idx->rounds.assign(rounds, rounds + n);
    return RABITQ_OK;
}

int rabitq_set_option(rabitq_index* idx, const char* name, long value) {
    if (!idx || !name) return fail(RABITQ_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    const std::string n(name);
    if (n == "first_chunks") idx->first_chunks = (int)std::max(0L, value);
    else if (n == "scan_mode") idx->scan_mode = (int)value;
    else if (n == "rerank_rows") idx->rerank_rows = (int)value;
    else if (n == "rerank_prefetch") idx->rerank_prefetch = (int)value;
    else if (n == "rerank_mode") idx->rerank_mode = (int)value;
    else if (n == "rerank_nc") idx->rerank_nc = (int)value;
    else if (n == "rerank_stages") idx->rerank_stages = (int)value;
    else if (n == "rerank_warps") idx->rerank_warps = (int)value;
    else if (n == "debug_rerank") idx->debug_rerank = (int)value;
    else if (n == "scan_slices") idx->scan_slices = (int)std::max(1L, value);
    else if (n == "scan_stages") idx->scan_stages = (int)std::max(0L, value);
    else if (n == "scan_sub") idx->scan_sub = (int)std::max(0L, value);
    else if (n == "prefilter") idx->prefilter = (int)value;
    else if (n == "prefilter_mode") { idx->pf_mode = (int)value; idx->pf_strikes = 0; }
    else if (n == "prefilter_cap") idx->prefilter_cap = (int)value;
    else if (n == "prefilter_gemm") idx->pf_gemm = (int)value;
    else if (n == "speculative_sizing") idx->spec_enabled = value != 0;
    else if (n == "dist_r2_seq") idx->dist_r2_seq = value != 0;
    else if (n == "spec_words_per_query_milli") idx->hw_wpq = (double)value / 1000.0;  // (tests: a tiny value makes the next batch overflow its speculative slots)
    else return fail(RABITQ_EINVAL, "unknown option: " + n);
    return RABITQ_OK;
}

int rabitq_debug_rerank_stats(rabitq_index* idx, uint32_t* out, size_t nq) {
    if (!idx || !out) return fail(RABITQ_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    CU(cudaSetDevice(idx->device));
    if (!idx->rr_dbg.p || idx->rr_dbg.cap < nq * 16 * 4) return fail(RABITQ_EINVAL, "set_option(\"debug_rerank\", 1) and run a batch of at least nq queries first");
    CU(cudaStreamSynchronize(idx->stream));
    CU(cudaMemcpy(out, idx->rr_dbg.p, nq * 16 * 4, cudaMemcpyDeviceToHost));
    return RABITQ_OK;
}

int rabitq_set_quantize_bias(rabitq_index* idx, const float* bias) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    CU(cudaSetDevice(idx->device));
    CU(cudaStreamSynchronize(idx->stream));
    if (!bias) {
        if (idx->quant_bias) cudaFree(idx->quant_bias);
        idx->quant_bias = nullptr;
        return RABITQ_OK;
    }
    if (!idx->quant_bias) CU(cudaMalloc((void**)&idx->quant_bias, (size_t)idx->D * 4));
    CU(cudaMemcpy(idx->quant_bias, bias, (size_t)idx->D * 4, cudaMemcpyHostToDevice));
    return RABITQ_OK;
}

int rabitq_set_stream(rabitq_index* idx, void* cuda_stream) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    idx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : idx->own_stream;
    return RABITQ_OK;
}

int rabitq_last_timings(const rabitq_index* idx, float ms[10], uint64_t counts[6]) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    {
        rabitq_index* ix = const_cast<rabitq_index*>(idx);
        std::lock_guard<std::mutex> lk(ix->mu);
        CU(cudaSetDevice(ix->device));
        if (resolve_timings(ix)) return RABITQ_ECUDA;
    }
    if (ms) std::memcpy(ms, idx->ms, sizeof(float) * 10);
    if (counts) std::memcpy(counts, idx->counts, sizeof(uint64_t) * 6);
    return RABITQ_OK;
}

// ---- stage entries --------------------------------------------------------------------------------------------------
int rabitq_stage_rotate(rabitq_index* idx, const float* queries, size_t nq, size_t len, float* out_y) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    BatchOut bo;
    int rc = stage_prefix(idx, queries, nq, len, 1, STOP_ROTATE, &bo);
    if (rc) return rc;
    CU(cudaMemcpy(out_y, idx->y.p, nq * idx->D * 4, cudaMemcpyDeviceToHost));
    return RABITQ_OK;
}

int rabitq_stage_probe(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe, float* out_centroid_dist,
                       uint32_t* out_probe_ids, float* out_probe_dist) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    BatchOut bo;
    const int pf_saved = idx->prefilter;
    if (out_centroid_dist) idx->prefilter = 0;  // the full exact distance matrix only exists on the classic path
    int rc = stage_prefix(idx, queries, nq, len, probe, STOP_PROBE, &bo);
    idx->prefilter = pf_saved;
    if (rc) return rc;
    if (out_centroid_dist) CU(cudaMemcpy(out_centroid_dist, idx->cdist.p, nq * idx->K * 4, cudaMemcpyDeviceToHost));
    if (out_probe_ids) CU(cudaMemcpy(out_probe_ids, idx->probe_ids.p, nq * bo.P * 4, cudaMemcpyDeviceToHost));
    if (out_probe_dist) CU(cudaMemcpy(out_probe_dist, idx->probe_dist.p, nq * bo.P * 4, cudaMemcpyDeviceToHost));
    return RABITQ_OK;
}

int rabitq_stage_quantize(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe, float* out_lo,
                          float* out_delta, uint32_t* out_sum, uint64_t* out_planes) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    BatchOut bo;
    int rc = stage_prefix(idx, queries, nq, len, probe, STOP_QUANT, &bo);
    if (rc) return rc;
    const size_t D = idx->D, W32 = D / 32, K = idx->K, items = nq * bo.P, pitch = (size_t)scan_rec_pitch((int)D);
    const size_t n_lists = round_bounds(idx, (int)bo.P, false).size() - 1;
    std::vector<uint32_t> planes(4 * W32);
    std::vector<char> seen(items, 0);
    for (size_t l = 0; l < n_lists; l++) {  // every (q, p) record is in the list of the round(s) that visit it
        ListSet& L = idx->lists[l];
        uint32_t n_l = 0;
        CU(cudaMemcpy(&n_l, L.cl_start.as<uint32_t>() + K, 4, cudaMemcpyDeviceToHost));
        std::vector<uint2> ids(n_l);
        std::vector<unsigned char> rec((size_t)n_l * pitch);
        if (n_l) {
            CU(cudaMemcpy(ids.data(), L.cl_items.p, (size_t)n_l * 8, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(rec.data(), L.qrec.p, (size_t)n_l * pitch, cudaMemcpyDeviceToHost));
        }
        for (size_t s = 0; s < n_l; s++) {
            const size_t i = ids[s].x;
            if (i >= items) return fail(RABITQ_ECUDA, "corrupt inverted list");
            seen[i] = 1;
            const unsigned char* cb = rec.data() + s * pitch;
            const uint32_t* r = reinterpret_cast<const uint32_t*>(cb);
            if (out_planes) {  // vector_binarize_query (src/simd.rs:83-107) of the record's 4-bit codes: [4][W64] u64 little-endian == [4][W32] u32
                std::fill(planes.begin(), planes.end(), 0u);
                for (size_t d = 0; d < D; d++) {
                    const uint32_t q4 = (uint32_t)cb[rec_pos((int)d)] >> (3 - (d & 3));
                    for (size_t b = 0; b < 4; b++) planes[b * W32 + d / 32] |= ((q4 >> b) & 1u) << (d & 31);
                }
                std::memcpy(out_planes + i * 2 * W32, planes.data(), 4 * W32 * 4);
            }
            if (out_lo) std::memcpy(&out_lo[i], r + D / 4 + 0, 4);
            if (out_delta) std::memcpy(&out_delta[i], r + D / 4 + 1, 4);
            if (out_sum) out_sum[i] = r[D / 4 + 5];
        }
    }
    for (size_t i = 0; i < items; i++)
        if (!seen[i] && idx->shard_count == 1) return fail(RABITQ_ECUDA, "a (query, probe) record is in no round's list");
    return RABITQ_OK;
}

int rabitq_stage_scan(rabitq_index* idx, const float* queries, size_t nq, size_t len, size_t probe, size_t pair_capacity,
                      float* out_rough, uint32_t* out_abdp, uint64_t* out_pair_start) {
    if (!idx) return fail(RABITQ_EINVAL, "null index");
    std::lock_guard<std::mutex> lk(idx->mu);
    BatchOut bo;
    int rc = stage_prefix(idx, queries, nq, len, probe, STOP_SCAN_DENSE, &bo);
    if (rc) return rc;
    const size_t P = bo.P, K = idx->K;
    if (bo.total_pairs > pair_capacity) return fail(RABITQ_EINVAL, "pair_capacity too small");
    std::vector<uint32_t> wbase(nq + 1), slot(nq * P), pids(nq * P), off(K + 1);
    std::vector<float2> ent((size_t)bo.total_words * 32);
    CU(cudaMemcpy(wbase.data(), idx->q_wbase.p, (nq + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(slot.data(), idx->slot_local.p, nq * P * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pids.data(), idx->probe_ids.p, nq * P * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(off.data(), idx->offsets, (K + 1) * 4, cudaMemcpyDeviceToHost));
    if (!ent.empty()) CU(cudaMemcpy(ent.data(), idx->entries.p, ent.size() * 8, cudaMemcpyDeviceToHost));
    size_t t = 0;
    for (size_t q = 0; q < nq; q++) {
        if (out_pair_start) out_pair_start[q] = t;
        for (size_t p = 0; p < P; p++) {
            uint32_t c = pids[q * P + p], n_c = off[c + 1] - off[c];
            size_t w = (size_t)wbase[q] + slot[q * P + p];
            for (uint32_t v = 0; v < n_c; v++, t++) {
                float2 e = ent[w * 32 + v];
                if (out_rough) out_rough[t] = e.x;
                if (out_abdp) std::memcpy(&out_abdp[t], &e.y, 4);
            }
        }
    }
    if (out_pair_start) out_pair_start[nq] = t;
    return RABITQ_OK;
}

}  // extern "C"
